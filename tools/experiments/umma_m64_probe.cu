// Hardware probe (not part of the library): where does a cta_group::1 tcgen05.mma with M = 64 put / read its TMEM rows?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I easy_vitpose_b200/csrc -o tools/experiments/umma_m64_probe tools/experiments/umma_m64_probe.cu
// Question behind it (attention): the 64-row second tile of every (crop, head) wastes half of the softmax
// lanes.  Two such half tiles could share one 128-lane pass if (1) an M = 64 accumulator can be placed at lane offset 16 of
// every 32-lane sub-partition (next to another one at offset 0) and (2) the A operand of a TS-MMA (P from TMEM) can be READ
// from lane offset 16 as well.  The vendored CUTLASS headers document (1) ("Interleaved" accumulator fragments) but build A
// fragments "NonInterleaved", which suggests (2) is not supported.  This program answers both on the device.
//
// D[m][n] = (m + 1) + (n + 1) / 64 for operand set 0 and (m + 101) + (n + 1) / 64 for operand set 1: the value read from a TMEM
// (lane, column) names the matrix element that landed there.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "attention.cuh"

using namespace vpb;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t sw128_off(int r, int k) {   // K-major, 64 bf16 per row, 128-byte swizzle
  return (r / 8) * 1024 + (r % 8) * 128 + (((k / 8) ^ (r % 8)) * 16) + (k % 8) * 2;
}
__device__ __forceinline__ float a_val(int set, int m, int k) { return k == 0 ? static_cast<float>(m + 1 + 100 * set) : (k == 1 ? 1.0f : 0.0f); }
__device__ __forceinline__ float b_val(int n, int k) { return k == 0 ? 1.0f : (k == 1 ? (n + 1) / 64.0f : 0.0f); }

// out: [5 tests][128 lanes][64 cols]
__global__ void __launch_bounds__(128, 1) probe(float* out, int only) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;            // [64 x 64] bf16, operand set 0
  uint8_t* sB = smem + 8192;     // [64 x 64] bf16
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 16384 + 64);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 64 * 64; i += 128) {
    const int r = i / 64, k = i % 64;
    *reinterpret_cast<__nv_bfloat16*>(sA + sw128_off(r, k)) = __float2bfloat16(a_val(0, r, k));
    *reinterpret_cast<__nv_bfloat16*>(sB + sw128_off(r, k)) = __float2bfloat16(b_val(r, k));
  }
  if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  fence_proxy_async_smem();
  if (warp == 0) tmem_alloc(slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t base = *slot;
  const uint32_t lane_base = base + (static_cast<uint32_t>(warp * 32) << 16);
  // zero every column we will dump, so that untouched lanes read 0 and not stale data
  {
    uint32_t z[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) z[i] = 0;
    for (int c = 0; c < 512; c += 16) tmem_st16(lane_base + c, z);
    tmem_st_wait();
  }
  // A operands in TMEM (packed bf16 pairs, 32 columns for K = 64) at columns [128,160): lanes 0..15 of every sub-partition
  // carry operand set 0 (row m = 16 * quarter + lane), lanes 16..31 operand set 1 (same rows)
  {
    const int set = (lane >= 16) ? 1 : 0;
    const int m = warp * 16 + (lane & 15);
    uint32_t pk[16];
    for (int half = 0; half < 2; ++half) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int k = (half * 16 + j) * 2;
        pk[j] = pack_bf16(a_val(set, m, k), a_val(set, m, k + 1));
      }
      tmem_st16(lane_base + 128 + half * 16, pk);
    }
    tmem_st_wait();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  constexpr uint32_t idesc = umma_idesc_bf16(64, 64);
  uint32_t phase = 0;
  auto run = [&](int test) {
    if (tid == 0) {
      const uint64_t adesc = umma_desc_sw128(smem_u32(sA), 1024), bdesc = umma_desc_sw128(smem_u32(sB), 1024);
      for (int k = 0; k < 4; ++k) {
        if (test == 0) umma_bf16(base + 0, adesc + 2 * k, bdesc + 2 * k, idesc, k != 0);                        // SS, D at lane 0
        if (test == 1) umma_bf16(base + (16u << 16) + 64, adesc + 2 * k, bdesc + 2 * k, idesc, k != 0);         // SS, D at lane 16
        if (test == 2) umma_bf16_ts(base + 192, base + 128 + 8 * k, bdesc + 2 * k, idesc, k != 0);              // TS, A lane 0, D lane 0
        if (test == 3) umma_bf16_ts(base + (16u << 16) + 256, base + (16u << 16) + 128 + 8 * k, bdesc + 2 * k, idesc, k != 0);   // TS, A lane 16, D lane 16
        if (test == 4) umma_bf16_ts(base + 320, base + (16u << 16) + 128 + 8 * k, bdesc + 2 * k, idesc, k != 0);                 // TS, A lane 16, D lane 0
      }
      umma_commit(bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after_sync();
  };
  const int dcol[5] = {0, 64, 192, 256, 320};
  for (int test = 0; test < 5; ++test) {
    if (only >= 0 && test != only) continue;
    run(test);
    for (int c = 0; c < 64; c += 32) {
      uint32_t r[32];
      tmem_ld32(lane_base + dcol[test] + c, r);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) out[(test * 128 + tid) * 64 + c + j] = __uint_as_float(r[j]);
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(base, 512);
}

int main(int argc, char** argv) {      // optional argument: run only that test (one process per test: an illegal operand address kills the context)
  const int only = argc > 1 ? atoi(argv[1]) : -1;
  float* d_out = nullptr;
  const size_t n = 5 * 128 * 64;
  CK(cudaMalloc(&d_out, n * sizeof(float)));
  CK(cudaMemset(d_out, 0, n * sizeof(float)));
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 20480));
  probe<<<1, 128, 20480>>>(d_out, only);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> h(n);
  CK(cudaMemcpy(h.data(), d_out, n * sizeof(float), cudaMemcpyDeviceToHost));
  const char* names[5] = {"SS  M=64, D at lane offset 0", "SS  M=64, D at lane offset 16", "TS  M=64, A lane 0  -> D lane 0",
                          "TS  M=64, A lane 16 -> D lane 16", "TS  M=64, A lane 16 -> D lane 0"};
  for (int t = 0; t < 5; ++t) {
    if (only >= 0 && t != only) continue;
    printf("== test %d: %s\n", t, names[t]);
    int used = 0, consistent = 0;
    for (int l = 0; l < 128; ++l) {
      const float* row = &h[(t * 128 + l) * 64];
      bool any = false;
      for (int c = 0; c < 64; ++c) any = any || row[c] != 0.0f;
      if (!any) continue;
      ++used;
      const int m0 = static_cast<int>(row[0]);          // (m + 1 [+ 100]) + 1/64
      bool ok = true;
      for (int c = 0; c < 64; ++c) ok = ok && row[c] == static_cast<float>(m0) + (c + 1) / 64.0f;
      consistent += ok;
      if (l < 40 || !ok) printf("   lane %3d: row value %d (%s), col0 %.4f col63 %.4f\n", l, m0, ok ? "all 64 columns as expected" : "COLUMNS UNEXPECTED", row[0], row[63]);
    }
    printf("   lanes holding data: %d, of which fully consistent rows: %d\n", used, consistent);
  }
  cudaFree(d_out);
  return 0;
}
