// Micro-benchmark for the attention softmax stage (tools/, not part of the product): per-SM throughput of
//   ex2.approx, cvt.rn.bf16x2.f32, both, an integer RNE pack, FFMA, tcgen05.ld 32x32b.x32, and tcgen05.ld overlapped with ex2
// from other warps.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_softmax_pipes tools/ubench_softmax_pipes.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t cvt2(float a, float b) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a)); return r; }
__device__ __forceinline__ uint32_t pack_int(float a, float b) {      // round-to-nearest-even by integer arithmetic
  uint32_t x = __float_as_uint(a), y = __float_as_uint(b);
  x += 0x7fffu + ((x >> 16) & 1u); y += 0x7fffu + ((y >> 16) & 1u);
  return __byte_perm(x, y, 0x7632);
}

// mode: 0 ex2, 1 cvt, 2 ex2+ex2+cvt (softmax inner loop), 3 ex2+ex2+int pack, 4 ffma, 5 ex2 x2 + ffma x2 + cvt (full inner loop)
__global__ void alu_bench(int mode, int iters, float* out, long long* cycles) {
  float a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = threadIdx.x * 1e-3f + j;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      if (mode == 0) { a[j] = ex2(a[j]); a[j + 1] = ex2(a[j + 1]); }
      else if (mode == 1) { acc ^= cvt2(a[j], a[j + 1]); a[j] += 1.0f; }
      else if (mode == 2) { const float e0 = ex2(a[j]), e1 = ex2(a[j + 1]); acc ^= cvt2(e0, e1); a[j] = e0; a[j + 1] = e1; }
      else if (mode == 3) { const float e0 = ex2(a[j]), e1 = ex2(a[j + 1]); acc ^= pack_int(e0, e1); a[j] = e0; a[j + 1] = e1; }
      else if (mode == 4) { a[j] = fmaf(a[j], 1.0001f, 0.5f); a[j + 1] = fmaf(a[j + 1], 1.0001f, 0.5f); }
      else { const float e0 = ex2(fmaf(a[j], 1.44f, -3.f)), e1 = ex2(fmaf(a[j + 1], 1.44f, -3.f)); acc ^= cvt2(e0, e1); a[j] = e0 + e1; a[j + 1] = e1; }
    }
  }
  const long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += a[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + __uint_as_float(acc & 0x7fffff);
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// TMEM: warps [0, ld_warps) stream tcgen05.ld.32x32b.x32 from their lane quarter; warps [ld_warps, ld_warps + ex_warps) run ex2
__global__ void tmem_bench(int ld_warps, int ex_warps, int iters, float* out, long long* cycles) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = slot;
  float keep = 0;
  const long long t0 = clock64();
  if (warp < ld_warps) {
    const uint32_t addr = base + ((uint32_t)((warp & 3) * 32) << 16);
    for (int i = 0; i < iters; ++i) {
      uint32_t r[32];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                     "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                     "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                     "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                   : "r"(addr + (i & 7) * 32));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      keep += __uint_as_float(r[i & 31] & 0x3fffffff);
    }
  } else if (warp < ld_warps + ex_warps) {
    float a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = threadIdx.x * 1e-3f + j;
    for (int i = 0; i < iters; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = ex2(a[j]);      // 8 ex2 per iteration per thread = 32 B/thread of "work" vs 128 B per ld
#pragma unroll
    for (int j = 0; j < 8; ++j) keep += a[j];
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = keep;
  if ((threadIdx.x & 31) == 0) cycles[blockIdx.x * 16 + warp] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base));
}

int main() {
  float* out; long long* cyc;
  CK(cudaMalloc(&out, 148 * 512 * sizeof(float)));
  CK(cudaMalloc(&cyc, 148 * 16 * sizeof(long long)));
  long long h[148 * 16];
  const char* names[] = {"ex2", "cvt.bf16x2", "2 ex2 + cvt", "2 ex2 + int pack", "ffma", "2 ffma + 2 ex2 + cvt + fadd"};
  const double ops[] = {8, 4, 12, 8, 8, 8};      // counted ops per iteration per thread (mode 2: 8 ex2 + 4 cvt; mode 3/5: 8 ex2)
  for (int threads : {128, 256}) {
    for (int mode = 0; mode < 6; ++mode) {
      const int iters = 4096;
      alu_bench<<<148, threads>>>(mode, iters, out, cyc);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(h, cyc, 148 * sizeof(long long), cudaMemcpyDeviceToHost));
      double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
      printf("alu threads=%d %-28s %8.0f clk  -> %.1f counted ops/clk/SM\n", threads, names[mode], c, ops[mode] * iters * threads / c);
    }
  }
  for (int cfg = 0; cfg < 5; ++cfg) {
    const int ldw[] = {4, 8, 4, 0, 8}, exw[] = {0, 0, 4, 4, 4};
    const int iters = 4096;
    tmem_bench<<<148, 32 * (ldw[cfg] + exw[cfg])>>>(ldw[cfg], exw[cfg], iters, out, cyc);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, cyc, 148 * 16 * sizeof(long long), cudaMemcpyDeviceToHost));
    double cl = 0, ce = 0;
    for (int i = 0; i < 148; ++i) { if (ldw[cfg]) cl += h[i * 16]; if (exw[cfg]) ce += h[i * 16 + ldw[cfg]]; }
    cl /= 148; ce /= 148;
    printf("tmem ld_warps=%d ex_warps=%d:", ldw[cfg], exw[cfg]);
    if (ldw[cfg]) printf("  ld %8.0f clk -> %.1f B/clk/SM", cl, 128.0 * 32 * ldw[cfg] * iters / cl);
    if (exw[cfg]) printf("  ex2 %8.0f clk -> %.1f ex2/clk/SM", ce, 8.0 * 32 * exw[cfg] * iters / ce);
    printf("\n");
  }
  return 0;
}
