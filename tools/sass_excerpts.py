#!/usr/bin/env python
"""SASS evidence for the contraction kernels of the built library (runs without a GPU): per kernel the mnemonic counts that prove
the Blackwell-native path (UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG / UTMAREDG = TMA, LDTM / STTM = tcgen05.ld / st, UTCBAR =
tcgen05.commit, no HMMA) and a listing excerpt around the first tcgen05.mma issue loop.  Writes profiles/r2_sass_excerpts.txt."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "easy_vitpose_b200", "csrc", "libvitpose_b200.so")
WANT = ["gemm_chain_tcgen05ILi256", "gemm_bf16_tcgen05ILi256ELi2", "gemm_bf16_tcgen05ILi256ELi0", "gemm_bf16_tcgen05ILi32ELi4", "attention_pack_tcgen05ILi64ELi8", "attention_tcgen05ILi64ELi0",
        "attention_tcgen05ILi80ELi0", "decode_heatmaps", "frame_to_patch_rows"]
MNEM = ["UTCHMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "LDTM", "STTM", "HMMA", "MUFU.EX2", "MUFU.TANH", "FMNMX3", "SYNCS", "ELECT", "BRA.U.ANY",
        "ACQBULK", "UCGABAR", "LDG.E.128", "STG.E.128", "RED", "NANOSLEEP"]


def main() -> None:
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs = {}
    cur = None
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            funcs[cur] = []
        elif cur is not None:
            funcs[cur].append(ln)
    out = ["round 2: cuobjdump -sass easy_vitpose_b200/csrc/libvitpose_b200.so (sm_100a); instruction text only, encodings stripped", ""]
    for want in WANT:
        names = [f for f in funcs if want in f]
        if not names:
            out.append(f"== {want}: not in the library")
            continue
        body = [re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", l) for l in funcs[names[0]] if re.search(r"/\*[0-9a-f]{4,6}\*/", l)]
        text = "\n".join(body)
        counts = {k: len(re.findall(r"\b" + re.escape(k), text)) for k in MNEM}
        out.append(f"== {names[0]}  ({len(body)} instructions)")
        out.append("   " + "  ".join(f"{k}={v}" for k, v in counts.items() if v))
        idx = next((i for i, l in enumerate(body) if "UTCHMMA" in l), None)
        if idx is not None:
            out.append("   -- excerpt around the first tcgen05.mma (wait on the full barrier, four K=16 UTCHMMA of one k-block, tcgen05.commit):")
            for l in body[max(0, idx - 10): idx + 14]:
                out.append("   " + re.sub(r"^\s*/\*([0-9a-f]{4,6})\*/\s*", r"\1  ", l).rstrip())
        out.append("")
    path = os.path.join(ROOT, "profiles", "r2_sass_excerpts.txt")
    with open(path, "w") as fh:
        fh.write("\n".join(out) + "\n")
    print("\n".join(out[:70]))


if __name__ == "__main__":
    main()
