// One persistent launch for a CHAIN of dependent GEMMs of a transformer block, with the LayerNorms between them:
//
//     [proj (+= x)] -> LN -> [fc1 + GELU] -> [fc2 (+= x)] -> LN -> [qkv of the next block]        (backbone/vit.py:202-205,164-180)
//
// Round 1 ran these as 4 GEMM launches + 2 LayerNorm launches.  Each tensor-core launch paid ~2 us of head (barrier init,
// TMEM alloc, first TMA round trip) and ~3 us of tail (last tile's epilogue, idle SMs in the last wave) that programmatic
// dependent launch cannot hide (a 227 KB CTA leaves no room for the next grid), and every LayerNorm was a separate pass of
// the fp32 stream through all SMs with nothing else running.  Here the tiles of all phases form ONE list, handed out
// statically (tile g -> cluster g mod #clusters, phase-major, n fastest), and what used to be a kernel boundary is a counter:
//
//   * every 128-row block `mt` of an activation has a counter that the producing epilogue warps bump once their TMA stores /
//     reduce-adds of a tile have COMPLETED (cp.async.bulk.wait_group 0, cross-proxy fence, release);
//   * the TMA-producer warp of a consuming tile spins (acquire) on the counter of its A rows before its first load;
//   * LayerNorm runs on four dedicated warps of every CTA, concurrently with that CTA's tensor-core tiles: 16-row jobs handed
//     out statically (job j -> CTA j mod grid), each waiting for its row block's residual adds to complete; they read the fp32
//     stream out of L2 (ld.global.cg) and write the bf16 operand rows; the gpu-scope part of a job -- polling the residual
//     counter, the acquire, the cross-proxy fence and the release that bumps the "normalised rows ready" counter -- runs on a
//     control warp (ln_ctl, default) so that the LayerNorm warps only load, normalise and store.
//
// Dependencies only point to tiles that come earlier in the list (and LN jobs only to tiles), all CTAs are resident (one per
// SM, grid <= #SMs), every role walks its own list in order (see "Tile list" in the kernel for the order): the smallest unfinished tile can always run, so the waits cannot
// deadlock; a counter that never arrives traps (VPB_HANG_TRAP_SPINS) instead of hanging the GPU.
// The arithmetic of every phase is that of gemm.cuh's kernels and of layernorm_f32_to_bf16, in the same order: results are
// bit-identical to the unchained path (tests/test_gpu_engine.py::test_chain_is_bit_identical).
//
//   warps 0..7   epilogue (tcgen05.ld -> bias / GELU -> swizzled smem staging -> TMA store or TMA reduce-add)
//   warp  8      TMEM allocator            warp 10  TMA producer (+ dependency waits)      warp 11  tcgen05.mma issuer (leader CTA)
//   warp  9      LayerNorm control (ln_ctl): polls the residual counters / publishes the jobs for warps 12..15
//   warps 12..15 LayerNorm jobs
#pragma once
#include "gemm.cuh"

namespace vpb {

constexpr int CHAIN_MAX_PHASES = 4;
constexpr int CHAIN_MAX_LN = 2;
constexpr int CHAIN_THREADS = 512;
constexpr int CHAIN_LN_WARPS = 4;
constexpr int CHAIN_LN_JOB_ROWS = 16;      // default rows per LayerNorm job (ChainParams::ln_job_rows: 8 or 16): 4 per warp

struct ChainPhase {
  int N, K;                 // W is [N,K]; N % BN == 0, K % 64 == 0
  int epi;                  // EPI_BF16 | EPI_BF16_GELU | EPI_BF16_GELU_ERF | EPI_F32_ADD
  const float* bias;        // [N]
  const int* a_ready;       // per 128-row block: A rows are complete once a_ready[mt] >= target (nullptr: produced by an earlier launch)
  int a_target;             // > 0: that many arrivals (GEMM-produced A: one per column tile of the producing phase);
                            // 0: LayerNorm-produced A: one arrival per 16-row job of the block
  int* out_done;            // per 128-row block, += 1 per tile once the CTA's stores of it have completed (nullptr: nobody waits)
};
struct ChainLn {
  const int* src_done;      // out_done of the residual phase that completes the fp32 rows
  int src_target;           // column tiles of that phase
  const float* gamma;       // [D]
  const float* beta;        // [D]
  int* ready;               // per 128-row block, += 1 per finished job
};
struct ChainParams {
  int M;                    // rows (tokens) of every phase
  int D;                    // LayerNorm width = row pitch of x / xn
  int num_phases, num_ln;
  const float* x;           // fp32 stream [M, D]
  __nv_bfloat16* xn;        // LayerNorm output [M, D]
  float eps;
  int wave_lag[2];          // tile order: lag (in 256-row pairs) of the second phase behind the first inside wavefronts {0,1} and {2,3}
  int ln_job_rows;          // rows per LayerNorm job: 16 (two 2-row iterations per warp) or 8 (one)
  int ln_ctl;               // 1 = the counter polls / publishes of the LayerNorm jobs run on a control warp (warp 9), 0 = on warp 12
  int rmw;                  // fp32 residual phases: 1 = load + add + store in the generic proxy (gemm.cuh: epilogue_f32_rmw), 0 = TMA reduce-add
  int dbg_nowait;           // measurement only (results may be wrong): publish tiles without waiting for their stores to complete
  long long* dbg;           // measurement: per cluster [CHAIN_MAX_PHASES][12] cycle counters (leader CTA) or nullptr (8..10: LayerNorm
                            //   stage s under phase 2s, warp 12: wait for the residual rows, busy, jobs):
                            //   0 mma busy+wait total, 1 mma wait full (operands), 2 mma wait acc_empty (epilogue), 3 producer dependency wait,
                            //   4 producer wait empty (ring), 5 epilogue(warp 0) busy, 6 epilogue wait acc_full, 7 tiles
  ChainPhase ph[CHAIN_MAX_PHASES];
  ChainLn ln[CHAIN_MAX_LN];
};
struct alignas(64) ChainMaps {
  CUtensorMap a[CHAIN_MAX_PHASES], w[CHAIN_MAX_PHASES], out[CHAIN_MAX_PHASES];
};

template <int BN>
struct ChainCfg {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_SLICE = BN * GEMM_BK * 2 / GEMM_CL;
  static constexpr int STAGE_BYTES = A_BYTES + B_SLICE;
  static constexpr int STAGING = GEMM_EPI_WARPS * GEMM_STAGE_TILE;
  static constexpr int STAGES_RAW = (227 * 1024 - 2048 - STAGING) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int ACC_STRIDE = BN <= 128 ? 128 : 256;
  static constexpr int TMEM_COLS = 2 * ACC_STRIDE;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int HALF = BN / 2;
  static_assert(BN == 256 || BN == 128, "chain tiles");
};

// ---------------------------------------------------------------- counters in global memory
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// generic proxy <-> async proxy (TMA) ordering for global memory
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ int ld_relaxed_gpu(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Whole warp spins (one coalesced load per probe) until *p >= target.  The probes are RELAXED loads and one acquire fence
// follows the successful one: an acquire load at gpu scope is compiled to LDG + CCTL.IVALL, i.e. every probe of every waiting
// warp invalidated the SM's L1 under the epilogue warps' bias loads (ncu source view of the first version: 8 % of all
// stall samples on that CCTL, and the GELU epilogue three times over its pipe-bound time).
__device__ __forceinline__ void wait_counter(const int* p, int target) {
  uint32_t spins = 0;
  while (ld_relaxed_gpu(p) < target) {
    __nanosleep(64);
    if (++spins > (VPB_HANG_TRAP_SPINS >> 3)) __trap();
  }
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
}
__host__ __device__ __forceinline__ int chain_ln_jobs_in_block(int M, int mt, int job_rows) {
  const int rows = M - mt * GEMM_BM < GEMM_BM ? M - mt * GEMM_BM : GEMM_BM;
  return (rows + job_rows - 1) / job_rows;
}

// ---------------------------------------------------------------- LayerNorm rows (same arithmetic, same order as
// layernorm_f32_to_bf16 in pointwise.cuh: fp32 mean, centred biased variance, rsqrtf; backbone/vit.py:190,198,304)
template <int V>
struct LnRow {
  float4 v[V];
  float rstd;
  __device__ __forceinline__ void load(const float* __restrict__ xrow, int lane) {
#pragma unroll
    for (int i = 0; i < V; ++i) v[i] = __ldcg(reinterpret_cast<const float4*>(xrow) + i * 32 + lane);    // L2: written by other SMs
  }
  // centres v in place and leaves 1/sqrt(var + eps) in rstd
  __device__ __forceinline__ void stats(float eps) {
    constexpr int D = 128 * V;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
      q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    rstd = rsqrtf(q * (1.0f / D) + eps);
  }
  __device__ __forceinline__ uint2 out(int i, const float4& g, const float4& b) const {
    uint2 o;
    o.x = pack_bf16(v[i].x * rstd * g.x + b.x, v[i].y * rstd * g.y + b.y);
    o.y = pack_bf16(v[i].z * rstd * g.z + b.z, v[i].w * rstd * g.w + b.w);
    return o;
  }
};
// rows [r0, r1) of one warp's share of a job; two rows in flight while they fit in registers (D <= 768), else one.
// gamma / beta are fetched once per 128-column slice and shared by the rows in flight.
template <int V>
__device__ __forceinline__ void chain_ln_rows(const ChainParams& p, const ChainLn& ln, int r0, int r1, int lane) {
  constexpr int D = 128 * V;
  constexpr int R = V <= 6 ? 2 : 1;
  for (int r = r0; r < r1; r += R) {
    LnRow<V> row[R];
    const bool two = R == 2 && r + 1 < r1;
    row[0].load(p.x + static_cast<size_t>(r) * D, lane);
    if constexpr (R == 2) { if (two) row[1].load(p.x + static_cast<size_t>(r + 1) * D, lane); }
    // gamma / beta are requested together with the rows (one exposed memory latency per job instead of two: the SM's L1 is
    // invalidated by every acquire fence of the CTA, so these are L2 round trips more often than not)
    float4 g[V], b[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
      g[i] = __ldg(reinterpret_cast<const float4*>(ln.gamma) + i * 32 + lane);
      b[i] = __ldg(reinterpret_cast<const float4*>(ln.beta) + i * 32 + lane);
    }
    row[0].stats(p.eps);
    if constexpr (R == 2) { if (two) row[1].stats(p.eps); }
    uint2* y0 = reinterpret_cast<uint2*>(p.xn + static_cast<size_t>(r) * D);
    uint2* y1 = reinterpret_cast<uint2*>(p.xn + static_cast<size_t>(r + 1) * D);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      y0[i * 32 + lane] = row[0].out(i, g[i], b[i]);
      if constexpr (R == 2) { if (two) y1[i * 32 + lane] = row[1].out(i, g[i], b[i]); }
    }
  }
}

// ---------------------------------------------------------------- one tile's epilogue (the TMA epilogues of gemm.cuh)
template <int BN, int EPI>
__device__ __forceinline__ void chain_epilogue_tile(uint32_t t_row, int half, int n0, int row0, const float* __restrict__ bias,
                                                    uint8_t* stile, int lane, const CUtensorMap* tmap_out) {
  constexpr int HALF = BN / 2;
  constexpr int COLS = (EPI == EPI_F32_ADD) ? 32 : 64;       // one 128-byte staging row per round
  const int sw = lane & 7;
#pragma unroll 1
  for (int c = 0; c < HALF; c += COLS) {
    const int col = half * HALF + c;
    const int n = n0 + col;
    if (elect_one()) tma_store_wait_read<0>();                // previous store has finished reading the staging tile
    __syncwarp();
#pragma unroll
    for (int sub = 0; sub < COLS; sub += 32) {
      uint32_t r[32];
      tmem_ld32(t_row + col + sub, r);
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + n + sub + j));
        v[j] = __uint_as_float(r[j]) + b4.x; v[j + 1] = __uint_as_float(r[j + 1]) + b4.y;
        v[j + 2] = __uint_as_float(r[j + 2]) + b4.z; v[j + 3] = __uint_as_float(r[j + 3]) + b4.w;
      }
      if constexpr (EPI == EPI_BF16_GELU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = gelu_fast(v[j]);
      }
      if constexpr (EPI == EPI_BF16_GELU_ERF) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = gelu_erf_as(v[j]);
      }
      uint8_t* srow = stile + lane * 128;                     // staging row = lane, 16-byte chunk index XOR (lane % 8): SWIZZLE_128B
      if constexpr (EPI == EPI_F32_ADD) {
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<float4*>(srow + ((q ^ sw) << 4)) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 w;
          w.x = pack_bf16(v[8 * q], v[8 * q + 1]); w.y = pack_bf16(v[8 * q + 2], v[8 * q + 3]);
          w.z = pack_bf16(v[8 * q + 4], v[8 * q + 5]); w.w = pack_bf16(v[8 * q + 6], v[8 * q + 7]);
          *reinterpret_cast<uint4*>(srow + ((((sub >> 3) + q) ^ sw) << 4)) = w;
        }
      }
    }
    fence_proxy_async_smem();                                 // staging writes -> visible to the TMA engine
    __syncwarp();
    if (elect_one()) {
      if constexpr (EPI == EPI_F32_ADD) tma_reduce_add_2d(tmap_out, stile, n, row0);
      else tma_store_2d(tmap_out, stile, n, row0);            // rows past M are clipped by the tensor map
      tma_store_commit();
    }
  }
}

template <int BN>
__global__ void __cluster_dims__(GEMM_CL, 1, 1) __launch_bounds__(CHAIN_THREADS, 1)
gemm_chain_tcgen05(const __grid_constant__ ChainMaps maps, const __grid_constant__ ChainParams p) {
  using Cfg = ChainCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = smem;
  uint8_t* staging = smem + Cfg::STAGES * Cfg::STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + Cfg::STAGING);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* acc_full = empty_bar + Cfg::STAGES;     // [2]
  uint64_t* acc_empty = acc_full + 2;               // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  uint64_t* ln_done = reinterpret_cast<uint64_t*>(tmem_slot + 2);    // [2] LayerNorm warps -> control warp: the job's rows are written

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int cta_rank = static_cast<int>(cluster_ctarank());
  const int cluster = static_cast<int>(cluster_id_x());
  const int num_clusters = static_cast<int>(cluster_count_x());
  const int num_m = (p.M + GEMM_BM - 1) / GEMM_BM;
  const int num_mp = (num_m + GEMM_CL - 1) / GEMM_CL;
  // Tile list.  Default (lag >= #pairs): PHASE-MAJOR, inside a phase tile = mp * num_n + nb (n fastest).  The general form pairs the
  // phases into two wavefronts, {0, 1} then {2, 3}: slot s holds the tiles of the first phase for row-block pair s and those of the
  // second phase for pair s - lag, so that a cluster alternates between a reduce-add phase (proj, fc2: epilogues bound by the
  // L2's fp32 adds) and its consumer.  Dependencies only point backwards for any lag (tests/test_chain_order.py), but every lag
  // tried was SLOWER than phase-major (engine.cu: VPB_CHAIN_LAG0/1), so this is kept as a measured experiment only.
  int n_of[CHAIN_MAX_PHASES];
#pragma unroll
  for (int i = 0; i < CHAIN_MAX_PHASES; ++i) n_of[i] = i < p.num_phases ? p.ph[i].N / BN : 0;
  const int lag0 = p.wave_lag[0] < num_mp ? p.wave_lag[0] : num_mp, lag1 = p.wave_lag[1] < num_mp ? p.wave_lag[1] : num_mp;
  const int wave0_tiles = num_mp * (n_of[0] + n_of[1]);
  const int total_tiles = wave0_tiles + num_mp * (n_of[2] + n_of[3]);
  auto locate = [&](int g, int& ph, int& mp, int& nb) {
    const bool w1 = g >= wave0_tiles;
    const int gg = w1 ? g - wave0_tiles : g;
    const int na = w1 ? n_of[2] : n_of[0], nbb = w1 ? n_of[3] : n_of[1];
    const int lag = nbb == 0 ? num_mp : (w1 ? lag1 : lag0);
    const int pa = w1 ? 2 : 0;
    const int head = na * lag;                               // slots [0, lag): first phase only
    const int mid = (num_mp - lag) * (na + nbb);             // slots [lag, num_mp): both
    if (gg < head) { ph = pa; mp = gg / na; nb = gg % na; }
    else if (gg < head + mid) {
      const int q = gg - head, s = lag + q / (na + nbb), r = q % (na + nbb);
      if (r < na) { ph = pa; mp = s; nb = r; }
      else { ph = pa + 1; mp = s - lag; nb = r - na; }
    } else {
      const int q = gg - head - mid;                         // slots [num_mp, num_mp + lag): second phase only
      ph = pa + 1; mp = num_mp - lag + q / nbb; nb = q % nbb;
    }
  };
  constexpr uint16_t kAllCtas = (1u << GEMM_CL) - 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.num_phases; ++i) {
      tma_prefetch_desc(&maps.a[i]);
      tma_prefetch_desc(&maps.w[i]);
      tma_prefetch_desc(&maps.out[i]);
    }
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], GEMM_CL * GEMM_EPI_WARPS);
      mbar_init(&ln_done[s], CHAIN_LN_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc_pair(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before_sync();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem_base = uniform_u32(*tmem_slot);
  pdl_launch_dependents();
  pdl_wait();                                       // the previous kernel's outputs (first phase's A operand, x) are complete

  if (warp == 10) {
    // ------------------------------------------------------------ TMA producer (+ the waits that replace kernel boundaries)
    int stage = 0;
    uint32_t phase = 0;
    int prev_ph = -1;
    for (int g = cluster; g < total_tiles; g += num_clusters) {
      int ph, mp, nb;
      locate(g, ph, mp, nb);
      const ChainPhase& P = p.ph[ph];
      const int mt = mp * GEMM_CL + cta_rank;
      const int m0 = mt * GEMM_BM, n0 = nb * BN;
      const bool first_of_phase = ph != prev_ph;
      prev_ph = ph;
      const int num_kb = P.K / GEMM_BK;
      long long c0 = p.dbg ? clock64() : 0, w_dep = 0, w_ring = 0;
      if (P.a_ready != nullptr && mt < num_m) {
        wait_counter(P.a_ready + mt, P.a_target > 0 ? P.a_target : chain_ln_jobs_in_block(p.M, mt, p.ln_job_rows));
        fence_proxy_async_all();                    // the rows were written through the generic / async proxy of other SMs
      }
      if (p.dbg) w_dep = clock64() - c0;
      for (int kb = 0; kb < num_kb; ++kb) {
        if (p.dbg) c0 = clock64();
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (p.dbg) w_ring += clock64() - c0;
        uint8_t* sa = ring + stage * Cfg::STAGE_BYTES;
        if (elect_one()) {
          if (cta_rank == 0) mbar_expect_tx(&full_bar[stage], GEMM_CL * Cfg::STAGE_BYTES);
          tma_load_2d_pair(sa, &maps.a[ph], &full_bar[stage], kb * GEMM_BK, m0);
          tma_load_2d_pair(sa + Cfg::A_BYTES, &maps.w[ph], &full_bar[stage], kb * GEMM_BK, n0 + cta_rank * (BN / GEMM_CL));
        }
        __syncwarp();
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
      }
      if (p.dbg && cta_rank == 0 && lane == 0) {
        p.dbg[(cluster * CHAIN_MAX_PHASES + ph) * 12 + 3] += w_dep;
        if (first_of_phase) p.dbg[(cluster * CHAIN_MAX_PHASES + ph) * 12 + 11] += w_dep;      // of which: this cluster's first tile of the phase
        p.dbg[(cluster * CHAIN_MAX_PHASES + ph) * 12 + 4] += w_ring;
      }
    }
  } else if (warp == 11 && cta_rank == 0) {
    // ------------------------------------------------------------ MMA issuer (leader CTA of the pair only)
    constexpr uint32_t idesc = umma_idesc_bf16(GEMM_CL * GEMM_BM, BN);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int g = cluster; g < total_tiles; g += num_clusters, ++it) {
      int ph, mp, nb;
      locate(g, ph, mp, nb);
      const int num_kb = p.ph[ph].K / GEMM_BK;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const long long t0 = p.dbg ? clock64() : 0;
      long long w_full = 0, c0 = 0;
      mbar_wait(&acc_empty[acc], acc_phase ^ 1);
      const long long w_acc = p.dbg ? clock64() - t0 : 0;
      tc_fence_after_sync();
      const uint32_t d_tmem = tmem_base + acc * Cfg::ACC_STRIDE;
      for (int kb = 0; kb < num_kb; ++kb) {
        if (p.dbg) c0 = clock64();
        mbar_wait(&full_bar[stage], phase);
        if (p.dbg) w_full += clock64() - c0;
        tc_fence_after_sync();
        const uint32_t sa = smem_u32(ring + stage * Cfg::STAGE_BYTES);
        const uint64_t adesc = umma_desc_sw128(sa, 1024);
        const uint64_t bdesc = umma_desc_sw128(sa + Cfg::A_BYTES, 1024);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) umma_bf16_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          umma_commit_pair(&empty_bar[stage], kAllCtas);
          if (kb == num_kb - 1) umma_commit_pair(&acc_full[acc], kAllCtas);
        }
        __syncwarp();
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
      }
      if (p.dbg && lane == 0) {
        long long* d = p.dbg + (cluster * CHAIN_MAX_PHASES + ph) * 12;
        d[0] += clock64() - t0; d[1] += w_full; d[2] += w_acc; d[7] += 1;
      }
    }
  } else if (warp < GEMM_EPI_WARPS) {
    // ------------------------------------------------------------ epilogue
    const int quarter = warp & 3;
    const int half = warp >> 2;
    uint8_t* stile = staging + warp * GEMM_STAGE_TILE;
    int it = 0;
    for (int g = cluster; g < total_tiles; g += num_clusters, ++it) {
      int ph, mp, nb;
      locate(g, ph, mp, nb);
      const ChainPhase& P = p.ph[ph];
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int mt = mp * GEMM_CL + cta_rank;
      const int row0 = mt * GEMM_BM + quarter * 32;
      const long long e0 = p.dbg ? clock64() : 0;
      // load + add + store form of the residual phases: the warp's first 32 x 32 box of x.  Rows last written by an earlier launch
      // (no residual phase before this one in the launch) are requested before the accumulator is ready.
      const bool rmw = p.rmw != 0 && P.epi == EPI_F32_ADD;
      const int x_col0 = nb * BN + half * Cfg::HALF;
      bool x_early = rmw;
      for (int j = 0; j < ph; ++j) x_early = x_early && p.ph[j].epi != EPI_F32_ADD;
      float4 xr[8];
      if (x_early) rmw_load_box(xr, p.x, p.D, row0, lane, p.M, x_col0);
      mbar_wait(&acc_full[acc], acc_phase);
      const long long e1 = p.dbg ? clock64() : 0;
      tc_fence_after_sync();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * Cfg::ACC_STRIDE;
      if (rmw) {
        if (!x_early) rmw_load_box(xr, p.x, p.D, row0, lane, p.M, x_col0);
        epilogue_f32_rmw<Cfg::HALF / 32>(t_row + half * Cfg::HALF, x_col0, row0, p.M, P.bias, const_cast<float*>(p.x), p.D, xr, stile, lane);
      } else if (P.epi == EPI_F32_ADD) chain_epilogue_tile<BN, EPI_F32_ADD>(t_row, half, nb * BN, row0, P.bias, stile, lane, &maps.out[ph]);
      else if (P.epi == EPI_BF16_GELU) chain_epilogue_tile<BN, EPI_BF16_GELU>(t_row, half, nb * BN, row0, P.bias, stile, lane, &maps.out[ph]);
      else if (P.epi == EPI_BF16_GELU_ERF) chain_epilogue_tile<BN, EPI_BF16_GELU_ERF>(t_row, half, nb * BN, row0, P.bias, stile, lane, &maps.out[ph]);
      else chain_epilogue_tile<BN, EPI_BF16>(t_row, half, nb * BN, row0, P.bias, stile, lane, &maps.out[ph]);
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(&acc_empty[acc], 0);   // TMEM is free for the MMA thread; the stores may still be in flight
      if (P.out_done != nullptr) {
        // publish the tile ONCE per CTA: every epilogue warp waits until its own stores / reduce-adds have been PERFORMED (not
        // just read out of smem) and fences them across the proxies; the eight warps meet on a named barrier; one thread
        // does the gpu-scope release on the row block's counter.  (First version: fence + release per warp = eight
        // MEMBAR.GPU / ERRBAR / CCTL.IVALL sequences per tile -- 15 % of the kernel's stall samples.)
        if (elect_one()) {
          if (!p.dbg_nowait) tma_store_wait_all<0>();
          fence_proxy_async_all();
        }
        __syncwarp();
        asm volatile("bar.sync 3, 256;" ::: "memory");
        if (warp == 0 && mt < num_m && elect_one()) red_release_gpu_add(P.out_done + mt, 1);
      }
      if (p.dbg && warp == 0 && lane == 0 && cta_rank == 0) {
        long long* d = p.dbg + (cluster * CHAIN_MAX_PHASES + ph) * 12;
        d[5] += clock64() - e1; d[6] += e1 - e0;
      }
    }
    if (elect_one()) tma_store_wait_all<0>();
  } else if (warp == 9 && p.ln_ctl != 0) {
    // ------------------------------------------------------------ LayerNorm control warp
    // The gpu-scope fences around a job -- the acquire after the poll of the residual phase's counter, fence.proxy.async +
    // red.release to publish the normalised rows -- cost the LayerNorm warps ~3 k of the ~10.5 k cycles a job takes them
    // (tools/chain_diag.py), and the LayerNorm stages are what the consumer phases wait for (fc1 / qkv dependency waits: a quarter
    // of a chained launch).  This warp takes both over: it walks the CTA's job list (job j -> CTA j mod grid, stage-major), polls
    // job i + 1 while the LayerNorm warps work on job i and publishes job i when they have arrived on ln_done.  Two slots each
    // way (slot = job sequence number & 1; "ready": named barriers 4 / 5, "done": mbarriers); "ready" for job i + 2 is only
    // signalled after "done" of job i was seen, so neither barrier can run a phase ahead.  Neither probe blocks: the first job of stage 1 waits for fc2 tiles that may
    // themselves wait for this CTA's last job of stage 0, which must be publishable in the meantime.
    const int jobs = (p.M + p.ln_job_rows - 1) / p.ln_job_rows;
    const int first = static_cast<int>(blockIdx.x), step = static_cast<int>(gridDim.x);
    const int end_s = first < jobs ? p.num_ln : 0;
    int a_s = 0, a_job = first, a_seq = 0;            // poll cursor
    int b_s = 0, b_job = first, b_seq = 0;            // publish cursor
    uint32_t idle = 0;
    while (b_s < end_s) {
      bool progressed = false;
      // neither probe blocks: a job that is done is published even while the next one's source rows are still outstanding
      if (a_s < end_s && a_seq <= b_seq + 1) {
        const ChainLn& L = p.ln[a_s];
        if (__shfl_sync(0xffffffffu, ld_relaxed_gpu(L.src_done + (a_job * p.ln_job_rows) / GEMM_BM), 0) >= L.src_target) {
          asm volatile("fence.acq_rel.gpu;" ::: "memory");
          __syncwarp();
          // "ready" is a NAMED barrier (ids 4 / 5 by slot, 4 LayerNorm warps + this one): the LayerNorm warps sleep in
          // bar.sync without taking issue slots.  First version: an mbarrier they polled with try_wait -- four more spinning
          // warps per SM, the epilogue warps sharing their schedulers lost 15 % (fc1 epilogue 8.8 k -> 10.3 k cycles per tile)
          // and the MMA thread's accumulator waits rose from 5 % to 19 % of the fc1 phase, which ate the gain.
          if (a_seq & 1) asm volatile("bar.arrive 5, 160;" ::: "memory");
          else asm volatile("bar.arrive 4, 160;" ::: "memory");
          ++a_seq;
          a_job += step;
          if (a_job >= jobs) { ++a_s; a_job = first; }
          progressed = true;
        }
      }
      if (b_seq < a_seq && __shfl_sync(0xffffffffu, mbar_test_wait(&ln_done[b_seq & 1], (b_seq >> 1) & 1) ? 1 : 0, 0)) {
        if (lane == 0) {
          fence_proxy_async_all();                    // consumed by TMA loads (async proxy) of other SMs
          red_release_gpu_add(p.ln[b_s].ready + (b_job * p.ln_job_rows) / GEMM_BM, 1);
        }
        __syncwarp();
        ++b_seq;
        b_job += step;
        if (b_job >= jobs) { ++b_s; b_job = first; }
        progressed = true;
      }
      if (progressed) idle = 0;
      else {
        __nanosleep(96);                              // this warp shares a scheduler with two epilogue warps
        if (++idle > (VPB_HANG_TRAP_SPINS >> 3)) __trap();
      }
    }
  } else if (warp >= 12) {
    // ------------------------------------------------------------ LayerNorm jobs
    const int lw = warp - 12;
    const int jobs = (p.M + p.ln_job_rows - 1) / p.ln_job_rows;
    const int ROWS_PER_WARP = p.ln_job_rows / CHAIN_LN_WARPS;
    const bool ctl = p.ln_ctl != 0;                  // polls / publishes on the control warp (warp 9)
    uint32_t seq = 0;                                 // job sequence number of this CTA over both stages (mbarrier slot / parity)
    for (int s = 0; s < p.num_ln; ++s) {
      const ChainLn& L = p.ln[s];
      for (int job = blockIdx.x; job < jobs; job += gridDim.x, ++seq) {
        const int mt = (job * p.ln_job_rows) / GEMM_BM;
        const long long l0 = p.dbg ? clock64() : 0;
        if (ctl) {
          // the control warp has acquired the rows at gpu scope and arrived on the slot's named barrier
          if (seq & 1) asm volatile("bar.sync 5, 160;" ::: "memory");
          else asm volatile("bar.sync 4, 160;" ::: "memory");
        } else {
          // one warp polls the counter, the other three sleep on a named barrier (bar.sync carries the acquired state over)
          if (lw == 0) wait_counter(L.src_done + mt, L.src_target);
          asm volatile("bar.sync 4, 128;" ::: "memory");
        }
        const long long l1 = p.dbg ? clock64() : 0;
        const int r0 = job * p.ln_job_rows + lw * ROWS_PER_WARP;
        const int r1 = min(r0 + ROWS_PER_WARP, p.M);
        switch (p.D) {
          case 384: chain_ln_rows<3>(p, L, r0, r1, lane); break;
          case 768: chain_ln_rows<6>(p, L, r0, r1, lane); break;
          case 1024: chain_ln_rows<8>(p, L, r0, r1, lane); break;
          default: chain_ln_rows<10>(p, L, r0, r1, lane); break;     // 1280
        }
        // one gpu-scope release per job and CTA (a MEMBAR.GPU on an SM with TMA traffic in flight costs thousands of cycles):
        // the four warps meet on a named barrier (orders their row stores before the releasing thread), warp 12 publishes
        const long long l2 = p.dbg ? clock64() : 0;
        if (ctl) {
          __syncwarp();                                                // every lane's row stores precede the arrive
          if (lane == 0) mbar_arrive(&ln_done[seq & 1]);              // the control warp publishes the job
        } else {
          asm volatile("bar.sync 5, 128;" ::: "memory");
        }
        const long long l3 = p.dbg ? clock64() : 0;
        if (!ctl && lw == 0 && lane == 0) {
          fence_proxy_async_all();                                     // consumed by TMA loads (async proxy) of other SMs
          red_release_gpu_add(L.ready + mt, 1);
        }
        if (p.dbg && lw == 0 && lane == 0 && cta_rank == 0) {
          long long* d = p.dbg + (cluster * CHAIN_MAX_PHASES + 2 * s) * 12;
          const long long l4 = clock64();
          d[8] += l1 - l0; d[9] += l4 - l1; d[10] += 1;
          long long* e = d + 12;                                       // breakdown of the busy part, stored under the next phase's slots 8..10
          e[8] += l2 - l1; e[9] += l3 - l2; e[10] += l4 - l3;         // own rows (loads, statistics, stores) | wait for the other warps | publish
        }
      }
    }
  }

  tc_fence_before_sync();
  cluster_sync_all();
  if (warp == 8) tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
}

}  // namespace vpb
