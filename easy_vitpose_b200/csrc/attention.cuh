// Fused multi-head attention for one ViTPose crop: T = 192 tokens, head_dim 32 / 64 / 80 (ViT-S / B,L / H), on chip.
//
// Work item = (crop b, head h); a CTA walks items blockIdx.x, +gridDim.x, ...  Two CTAs share an SM (<= 92 KB smem,
// 256 TMEM columns each), so one CTA's softmax overlaps the other's tensor-core phases.  Per item, for each of the
// two 128-row M tiles (tokens 0..127, then 128..191):
//   TMA    Q,K,V [192 x hd] bf16 boxes straight out of the qkv activation [M, 3D] -> swizzled smem (once per item)
//   UMMA   S = Q K^T              M=128 x N=192, fp32 -> TMEM columns [0,192)                (q arrives pre-scaled)
//   SIMT   row softmax out of TMEM (one thread per row): max, exp2, sum; P is written back IN PLACE as packed bf16
//          (tcgen05.st, columns [0,96)) -- P never touches shared memory
//   UMMA   O = P V                A = P from TMEM, B = V as an MN-major smem operand, i.e. exactly the [token][dim]
//          box TMA delivered (no transpose); fp32 -> TMEM columns [O_COL0, O_COL0+hd)
//   SIMT   O / rowsum -> bf16 -> attn_out[b*192 + t, h*hd + d]
// Operand tiles: head_dim 64 -> one 128-byte-swizzled box per operand; 32 -> one 64-byte-swizzled box; 80 -> a
// 128B-swizzled box of 64 dims plus a 32B-swizzled box of the last 16 (QK^T: 4+1 K steps; PV: an N=64 and an N=16 MMA).
// The second M tile only has 64 live rows.  Even items take A rows 128..255 (live rows in TMEM lanes 0..63, the rest
// reads past Q into what follows: UMMA rows are independent, they only feed lanes nobody reads); odd items take A rows
// 64..191 (live rows in lanes 64..127), so the half-tile work alternates between warps 0-1 and warps 2-3.
//
// Warps 0..3: softmax / epilogue (warp w owns TMEM lane quarter w).  Warp 4, one thread: TMA + MMA issue.
// Q/K of the next item are fetched as soon as the item's last S is done, V as soon as its last PV is done.
#pragma once
#include <cuda.h>

#include "ptx.cuh"

namespace vpb {

constexpr int ATT_T = 192;
constexpr int ATT_WORKERS = 128;
constexpr int ATT_THREADS = ATT_WORKERS + 32;
constexpr int ATT_TMEM_COLS = 256;

template <int HD>
struct AttCfg {
  static_assert(HD == 32 || HD == 64 || HD == 80, "head_dim");
  static constexpr int MAIN = HD == 32 ? 32 : 64;             // dims in the main box
  static constexpr int TAIL = HD - MAIN;                      // 0 or 16 dims in the 32B-swizzled tail box
  static constexpr int MAIN_ROW = MAIN * 2;                   // bytes per row = swizzle span (128 or 64)
  static constexpr int MAIN_BYTES = ATT_T * MAIN_ROW;         // 24576 / 12288
  static constexpr int TAIL_BYTES = TAIL ? ATT_T * 32 : 0;    // 6144
  static constexpr int OPER_BYTES = MAIN_BYTES + TAIL_BYTES;  // one of Q / K / V
  static constexpr int OUT_PITCH = HD * 2 + 16;               // staging row pitch (bytes): conflict-free 16-byte accesses
  static constexpr int OUT_STAGE = 4 * 32 * OUT_PITCH;        // 4 worker warps x 32 rows
  static constexpr int SMEM = 3 * OPER_BYTES + OUT_STAGE + 1024 + 128;
  // O accumulator: behind S when it fits in 256 columns, else inside the dead upper half of S (P only needs [0,96))
  static constexpr int O_COL0 = HD <= 64 ? 192 : 96;
  static constexpr bool O_IN_S = O_COL0 < 192;
};

struct AttnParams {
  int batch;              // crops
  int heads;
  int dim;                // D = heads * head_dim
  __nv_bfloat16* out;     // [batch*192, D]
  long long* dbg;         // debug cycle counters per CTA [8] or nullptr: 0 lifetime, 1 worker wait S, 2 softmax, 3 worker wait O,
                          //   4 epilogue, 5 ctl wait P, 6 ctl wait O, 7 ctl wait loads
};

// P (A operand) from TMEM, V (B operand) from smem
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(static_cast<uint32_t>(accumulate))
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// tmap_main: box [192 rows x MAIN cols] (swizzle = MAIN*2 bytes); tmap_tail: box [192 x 16] (32B swizzle), hd 80 only.
template <int HD>
__global__ void __launch_bounds__(ATT_THREADS, 2)
attention_tcgen05(const __grid_constant__ CUtensorMap tmap_main, const __grid_constant__ CUtensorMap tmap_tail, const AttnParams p) {
  using Cfg = AttCfg<HD>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                       // each operand: main tile, then tail tile
  uint8_t* sK = sQ + Cfg::OPER_BYTES;
  uint8_t* sV = sK + Cfg::OPER_BYTES;
  uint8_t* sOut = sV + Cfg::OPER_BYTES;                      // per-warp output staging (coalesced global stores)
  uint64_t* bar_qk = reinterpret_cast<uint64_t*>(sOut + Cfg::OUT_STAGE);
  uint64_t* bar_v = bar_qk + 1;
  uint64_t* bar_s = bar_qk + 2;      // S tile complete            (MMA commit -> workers)
  uint64_t* bar_p = bar_qk + 3;      // P written, S consumed      (128 workers -> MMA thread)
  uint64_t* bar_o = bar_qk + 4;      // O tile complete            (MMA commit -> workers, MMA thread)
  uint64_t* bar_e = bar_qk + 5;      // O drained from TMEM        (128 workers -> MMA thread)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_qk + 6);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int items = p.batch * p.heads;
  const long long t_cta0 = p.dbg ? clock64() : 0;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_main);
    if constexpr (Cfg::TAIL > 0) tma_prefetch_desc(&tmap_tail);
    mbar_init(bar_qk, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, ATT_WORKERS);
    mbar_init(bar_o, 1);
    mbar_init(bar_e, ATT_WORKERS);
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, ATT_TMEM_COLS);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();                                               // qkv from the previous GEMM is complete

  if (warp == 4) {
    if (lane == 0) {
      // ------------------------------------------------------------------ TMA + MMA issue (one thread)
      auto load_oper = [&](uint8_t* dst, uint64_t* bar, int col0, int row0) {
        tma_load_2d(dst, &tmap_main, bar, col0, row0);
        if constexpr (Cfg::TAIL > 0) tma_load_2d(dst + Cfg::MAIN_BYTES, &tmap_tail, bar, col0 + Cfg::MAIN, row0);
      };
      auto load_qk = [&](int item) {
        const int b = item / p.heads, h = item % p.heads;
        mbar_expect_tx(bar_qk, 2 * Cfg::OPER_BYTES);
        load_oper(sQ, bar_qk, h * HD, b * ATT_T);
        load_oper(sK, bar_qk, p.dim + h * HD, b * ATT_T);
      };
      auto load_v = [&](int item) {
        const int b = item / p.heads, h = item % p.heads;
        mbar_expect_tx(bar_v, Cfg::OPER_BYTES);
        load_oper(sV, bar_v, 2 * p.dim + h * HD, b * ATT_T);
      };
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, ATT_T);
      constexpr uint32_t idesc_o_main = umma_idesc_bf16(128, Cfg::MAIN, /*b_mn_major=*/true);
      constexpr uint32_t idesc_o_tail = umma_idesc_bf16(128, 16, /*b_mn_major=*/true);
      if (blockIdx.x < items) { load_qk(blockIdx.x); load_v(blockIdx.x); }
      uint32_t step = 0;                                    // tile steps done so far: parity of bar_s/p/o/e
      uint32_t it = 0;                                      // items done so far: parity of bar_qk/bar_v
      long long c_wp = 0, c_wo = 0, c_wl = 0, c0;
      for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
        const int next = item + gridDim.x;
        c0 = clock64();
        mbar_wait(bar_qk, it & 1);
        c_wl += clock64() - c0;
        tc_fence_after_sync();
        for (int mt = 0; mt < 2; ++mt, ++step) {
          const uint32_t par = step & 1;
          if (Cfg::O_IN_S && step > 0) { mbar_wait(bar_e, (step - 1) & 1); tc_fence_after_sync(); }   // O lives inside S here
          // S = Q K^T.  The S/P columns are free: the previous tile's PV (their last reader) was waited for below.
          const int q_row0 = mt == 0 ? 0 : ((it & 1) ? 64 : 128);
          const uint64_t qd = umma_desc_rows<Cfg::MAIN_ROW>(smem_u32(sQ) + q_row0 * Cfg::MAIN_ROW);
          const uint64_t kd = umma_desc_rows<Cfg::MAIN_ROW>(smem_u32(sK));
#pragma unroll
          for (int k = 0; k < Cfg::MAIN / 16; ++k) umma_bf16(tmem_base, qd + 2 * k, kd + 2 * k, idesc_s, k != 0);
          if constexpr (Cfg::TAIL > 0)
            umma_bf16(tmem_base, umma_desc_rows<32>(smem_u32(sQ) + Cfg::MAIN_BYTES + q_row0 * 32),
                      umma_desc_rows<32>(smem_u32(sK) + Cfg::MAIN_BYTES), idesc_s, true);
          umma_commit(bar_s);
          if (mt == 1) {                                    // last S of the item done -> Q/K smem can be refilled
            mbar_wait(bar_s, par);
            if (next < items) load_qk(next);
          }
          // O = P V once P is in TMEM and the previous O has been drained
          c0 = clock64();
          mbar_wait(bar_p, par);
          if (!Cfg::O_IN_S && step > 0) mbar_wait(bar_e, (step - 1) & 1);
          c_wp += clock64() - c0;
          c0 = clock64();
          if (mt == 0) mbar_wait(bar_v, it & 1);
          c_wl += clock64() - c0;
          tc_fence_after_sync();
#pragma unroll
          for (int kk = 0; kk < ATT_T / 16; ++kk) {
            // P: 16 bf16 of K = 8 packed TMEM columns per step.  V (MN-major): 16 tokens = two 8-row groups.
            umma_bf16_ts(tmem_base + Cfg::O_COL0, tmem_base + kk * 8,
                         umma_desc_rows<Cfg::MAIN_ROW>(smem_u32(sV) + kk * 16 * Cfg::MAIN_ROW), idesc_o_main, kk != 0);
            if constexpr (Cfg::TAIL > 0)
              umma_bf16_ts(tmem_base + Cfg::O_COL0 + Cfg::MAIN, tmem_base + kk * 8,
                           umma_desc_rows<32>(smem_u32(sV) + Cfg::MAIN_BYTES + kk * 16 * 32), idesc_o_tail, kk != 0);
          }
          umma_commit(bar_o);
          c0 = clock64();
          mbar_wait(bar_o, par);                            // PV retired: P columns and (after the item's last tile) V are free
          c_wo += clock64() - c0;
          if (mt == 1 && next < items) load_v(next);
        }
      }
      if (p.dbg) { p.dbg[blockIdx.x * 8 + 5] = c_wp; p.dbg[blockIdx.x * 8 + 6] = c_wo; p.dbg[blockIdx.x * 8 + 7] = c_wl; }
    }
  } else {
    // -------------------------------------------------------------------- softmax + epilogue (warps 0..3)
    const int quarter = warp;                               // TMEM lane quarter
    const int tl = quarter * 32 + lane;                     // TMEM lane = row of the M tile
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    uint32_t step = 0, it = 0;
    long long w_s = 0, w_sm = 0, w_o = 0, w_ep = 0, c0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
      const int b = item / p.heads, h = item % p.heads;
      for (int mt = 0; mt < 2; ++mt, ++step) {
        const uint32_t par = step & 1;
        // token handled by this thread; in the half tile only one pair of warps holds live rows
        int token;
        if (mt == 0) token = tl;
        else if (it & 1) token = 64 + tl;                   // A rows 64..191 -> lanes 64..127 hold tokens 128..191
        else token = 128 + tl;                              // A rows 128..255 -> lanes 0..63 hold tokens 128..191
        const bool live_warp = (mt == 0) || ((it & 1) ? quarter >= 2 : quarter < 2);   // warp-uniform

        c0 = clock64();
        mbar_wait(bar_s, par);
        w_s += clock64() - c0;
        c0 = clock64();
        tc_fence_after_sync();
        float sum = 1.0f;
        if (live_warp) {
          // Both passes keep one TMEM load in flight while the previous chunk is processed (ra / rb ping-pong).
          uint32_t ra[32], rb[32];
          float mx = -INFINITY;
          tmem_ld32(lane_addr, ra);
#pragma unroll
          for (int c = 0; c < ATT_T; c += 64) {
            tmem_ld_wait();
            tmem_ld32(lane_addr + c + 32, rb);
#pragma unroll
            for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(ra[j]));
            tmem_ld_wait();
            if (c + 64 < ATT_T) tmem_ld32(lane_addr + c + 64, ra);
#pragma unroll
            for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(rb[j]));
          }
          const float mscaled = mx * 1.4426950408889634f;
          sum = 0.0f;
          auto exp_chunk = [&](const uint32_t (&r)[32], int c) {
            uint32_t pk[16];
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const float e0 = ex2_approx(fmaf(__uint_as_float(r[j]), 1.4426950408889634f, -mscaled));
              const float e1 = ex2_approx(fmaf(__uint_as_float(r[j + 1]), 1.4426950408889634f, -mscaled));
              sum += e0 + e1;
              pk[j >> 1] = pack_bf16(e0, e1);
            }
            tmem_st16(lane_addr + (c >> 1), pk);            // P columns [c/2, c/2+16) trail the S read frontier
          };
          tmem_ld32(lane_addr, ra);
#pragma unroll
          for (int c = 0; c < ATT_T; c += 64) {
            tmem_ld_wait();
            tmem_ld32(lane_addr + c + 32, rb);              // S columns ahead of every P column written so far
            exp_chunk(ra, c);
            tmem_ld_wait();
            if (c + 64 < ATT_T) tmem_ld32(lane_addr + c + 64, ra);
            exp_chunk(rb, c + 32);
          }
          tmem_st_wait();
        }
        tc_fence_before_sync();
        mbar_arrive(bar_p);
        w_sm += clock64() - c0;

        c0 = clock64();
        mbar_wait(bar_o, par);
        w_o += clock64() - c0;
        c0 = clock64();
        tc_fence_after_sync();
        constexpr int OCH = HD / 16;                        // 16-column chunks of O
        uint32_t o[OCH][16];
        if (live_warp) {
#pragma unroll
          for (int q = 0; q < OCH; ++q) tmem_ld16(lane_addr + Cfg::O_COL0 + 16 * q, o[q]);
          tmem_ld_wait();
        }
        tc_fence_before_sync();
        mbar_arrive(bar_e);                                 // O is in registers: the next MMA may overwrite it
        if (live_warp) {
          // O rows -> this warp's smem staging (row pitch hd*2+16 B: conflict-free), then the warp writes its 32 rows
          // with consecutive lanes on consecutive 16-byte chunks of a row (full 128-byte lines instead of 32 rows per store).
          const float inv = 1.0f / sum;
          uint8_t* stage = sOut + quarter * 32 * Cfg::OUT_PITCH;
          uint8_t* srow = stage + lane * Cfg::OUT_PITCH;
#pragma unroll
          for (int q = 0; q < OCH; ++q) {
#pragma unroll
            for (int v = 0; v < 2; ++v) {
              uint4 w;
              w.x = pack_bf16(__uint_as_float(o[q][8 * v + 0]) * inv, __uint_as_float(o[q][8 * v + 1]) * inv);
              w.y = pack_bf16(__uint_as_float(o[q][8 * v + 2]) * inv, __uint_as_float(o[q][8 * v + 3]) * inv);
              w.z = pack_bf16(__uint_as_float(o[q][8 * v + 4]) * inv, __uint_as_float(o[q][8 * v + 5]) * inv);
              w.w = pack_bf16(__uint_as_float(o[q][8 * v + 6]) * inv, __uint_as_float(o[q][8 * v + 7]) * inv);
              *reinterpret_cast<uint4*>(srow + 32 * q + 16 * v) = w;
            }
          }
          __syncwarp();
          constexpr int CPR = HD / 8;                       // 16-byte chunks per row
          const int token0 = token - lane;                  // token of this warp's row 0 (rows are consecutive tokens)
          __nv_bfloat16* obase = p.out + (static_cast<size_t>(b) * ATT_T + token0) * p.dim + h * HD;
#pragma unroll
          for (int i = lane; i < 32 * CPR; i += 32) {
            const int rr = i / CPR, ch = i % CPR;
            const uint4 w = *reinterpret_cast<const uint4*>(stage + rr * Cfg::OUT_PITCH + ch * 16);
            *reinterpret_cast<uint4*>(obase + static_cast<size_t>(rr) * p.dim + ch * 8) = w;
          }
          __syncwarp();                                     // staging is reused by this warp's next tile
        }
        w_ep += clock64() - c0;
      }
    }
    if (p.dbg && threadIdx.x == 0) {
      p.dbg[blockIdx.x * 8 + 1] = w_s; p.dbg[blockIdx.x * 8 + 2] = w_sm; p.dbg[blockIdx.x * 8 + 3] = w_o; p.dbg[blockIdx.x * 8 + 4] = w_ep;
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, ATT_TMEM_COLS);
  if (p.dbg && threadIdx.x == 0) p.dbg[blockIdx.x * 8 + 0] = clock64() - t_cta0;
}

}  // namespace vpb
