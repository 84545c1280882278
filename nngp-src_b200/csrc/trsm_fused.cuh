// trsm_fused.cuh -- the prediction-side triangular solve as ONE persistent kernel (K10 + K11):
//
//     V = K_* L^-T   (row-major block buffer B, rows x N, overwritten in place)
//     var[i] = K(x_i,x_i) - sum_j V[i][j]^2
//
// Work item (r, J) = row tile r (128 test rows) x column block J (64 columns):
//     acc  = -(B[r, J] - V[r, 0:64J] * L[J, 0:64J]^T) = -R      TMA + DMMA main loop (mma_mainloop)
//     V[r, J] = R * W_J^T,  W_J = inv(L_JJ)                      4 more k-tiles on the SAME tensor pipe: R goes from
//                                                                the accumulators into the (drained) A stages of the
//                                                                ring in fragment layout, W_J arrives by TMA from the
//                                                                per-block inverses computed once per fit
//     ssq[row] += |V[row, J]|^2                                  (var written with the last column block)
// The diagonal step used to be a one-thread-per-row forward substitution on the FP64 CUDA cores.  Those DFMAs
// share the FP64 pipe with the co-resident CTA's DMMAs and queue behind them one by one: ncu showed 14 % of all
// warp samples parked at the barrier that ends the substitution.  As a 128 x 64 x 64 DMMA product it costs 4
// k-tiles (the average item has 254).  inv(L_JJ) of a 64 x 64 Cholesky diagonal block is benign: the error of
// R * W^T is eps * cond(L_JJ) <= eps * sqrt(cond(K + lambda I)).
// Items are claimed in J-major order from a global counter, so a CTA that needs V[r, 0:64J] finds the
// item (r, J-1) already claimed by a running CTA: it acquire-spins on progress[r] (no deadlock, no
// co-residency assumption).  Row tiles never interact, the per-row reduction order is fixed, so the
// result of a row is bitwise independent of the grid, of the block it sits in and of the GPU count.
// Compared with one GEMM launch + one solve launch per column block this removes the per-launch
// tail (1.73 waves of 296 CTAs -> 13.5 % idle), 2 N/64 launches, and one full re-read of V for the variance.
#pragma once
#include "dense_kernels.cuh"
#include "gemm_nt.cuh"

namespace nngp {

constexpr int TF_STAGES = 4;
constexpr int TF_RING_BYTES = TF_STAGES * (GEMM_A_STAGE_BYTES + GEMM_B_STAGE_BYTES);  // 96 KiB
static_assert(TF_STAGES * GEMM_BK == NB, "the diagonal step stages all 64 columns of R in the ring's A stages");
constexpr int TF_SMEM_BYTES = TF_RING_BYTES + 2 * TF_STAGES * 8 + 16 + 1024;

struct TrsmFusedParams {
  double* B;            // rows x N block buffer: K_* on entry, V on exit
  long long ldb;
  int rows;
  const double* L;      // N x N lower factor (tensor map tmL)
  long long ldl;        // (the per-block inverses inv(L_JJ) come through tensor map tmW: 64 rows x 64 per block)
  int N;
  int row_tiles, col_blocks;
  int* counter;         // zeroed before launch (dynamic schedule only)
  int static_sched;     // 1: CTA c takes items c, c+G, c+2G, ... (used when row_tiles % G == 0: every
                        //    dependency is then the CTA's own previous item -> no cross-CTA waiting)
  int* progress;        // [row_tiles], zeroed before launch
  double* ssq;          // [rows] running sum of squares (not required to be zeroed)
  const double* kss;    // [rows] K(x,x)
  double* var;          // [rows] output; nullptr: plain solve, no variance bookkeeping
  uint32_t zero;        // always 0 (value-initialised): see mma_mainloop's release dependence
  int upper_start;      // 1: B is upper triangular on entry (the identity, when the fit builds L^-T for latency mode), so
                        //    V[r, 0:128 r] stays zero: items left of the diagonal are skipped and every k loop starts at
                        //    column 128 r -- N^3/3 flop instead of N^3
};

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// PIPE = false: an item waits (once, up front) until all the column blocks it consumes are published -- the variant
//   for blocks with at least as many row tiles as CTAs, where that is practically always already the case.
// PIPE = true: every operand-tile load is gated on its producer's progress, so the CTAs working on one row tile
//   pipeline along J -- the variant for few row tiles (a handful of queries up to ~37 000).
template <bool PIPE>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
trsm_fused_kernel(const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmL,
                  const __grid_constant__ CUtensorMap tmW, const TrsmFusedParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
  uint8_t* ring = smem_raw + pad;
  uint8_t* ringA = ring;
  uint8_t* ringB = ring + TF_STAGES * GEMM_A_STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + TF_RING_BYTES);
  uint64_t* empty_bar = full_bar + TF_STAGES;
  int* s_item = reinterpret_cast<int*>(empty_bar + TF_STAGES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tid = threadIdx.x;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < TF_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], GEMM_CONSUMER_WARPS);
    }
    fence_mbar_init();
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmL);
    tma_prefetch_desc(&tmW);
  }

  const int total = p.row_tiles * p.col_blocks;
  int stage = 0;        // ring position: advances identically in every thread
  uint32_t phase = 0;
  const int wm = warp >> 1, wn = warp & 1;
  const int g = lane >> 2, t = lane & 3;

  int next_static = blockIdx.x;
  for (;;) {
    if (tid == 0) {
      const int it = p.static_sched ? next_static : atomicAdd(p.counter, 1);
      s_item[0] = it;
    }
    next_static += gridDim.x;
    __syncthreads();  // publishes the item; also: nobody still uses the ring / the scratch of the last item
    const int item = s_item[0];
    if (item >= total) break;
    const int J = item / p.row_tiles;
    const int r = item - J * p.row_tiles;
    const int col0 = J * NB;
    const int nb = min(NB, p.N - col0);
    const int row0 = r * GEMM_BM;
    // upper-triangular right-hand side: nothing to do left of the diagonal, and the k loop starts at the row tile's
    // own first column (block-uniform branch; a skipped item publishes nothing -- nobody ever waits for it)
    if (p.upper_start && col0 + NB <= row0) { __syncthreads(); continue; }   // (everyone has read s_item)
    const int kt0 = p.upper_start ? r * (GEMM_BM / GEMM_BK) : 0;
    const int ktiles = J * (NB / GEMM_BK) - kt0;

    TileSrc src;
    src.tmA = &tmB; src.tmB = &tmL;
    src.a_col0 = kt0 * GEMM_BK; src.a_row = row0; src.b_col0 = kt0 * GEMM_BK; src.b_row = col0;
    // The V operand of k-tile kt is column block kt/4 of this row tile, produced by item (r, kt/4): possibly by
    // another CTA, possibly still in flight.  Thread 0 gates every TMA issue on progress[r] > kt/4 (cached: one
    // acquire per newly needed block), so an item starts consuming the blocks that exist instead of waiting for all
    // of them -- with few row tiles the items of one row tile form a software pipeline along J across the CTAs.
    // Claims are handed out in J-major order, so whatever this CTA waits for is owned by a running CTA.
    int known = 0;   // progress[r] as last seen by thread 0
    auto gate = [&](int kt) {
      const int need = (kt + kt0) / (NB / GEMM_BK) + 1;
      if (known < need) {
        for (unsigned spin = 0; (known = ld_acquire_gpu(p.progress + r)) < need; ++spin) {
          __nanosleep(64);
          if (spin > (1u << 27)) __trap();   // bounded like mbar_wait: a scheduling bug traps instead of hanging
        }
        fence_proxy_async_all();  // V was written through the generic proxy by other CTAs -> read by TMA
      }
    };
    if (tid == 0) {
      if (!PIPE) {   // one up-front wait for item (r, J-1) (+ proxy fence), then a plain prologue
        if (ktiles > 0) gate(ktiles - 1);
        ring_prologue<TF_STAGES>(src, ringA, ringB, full_bar, stage, ktiles);
      } else {       // gated prologue: issue what exists, block by block
        const int n0 = ktiles < TF_STAGES ? ktiles : TF_STAGES;
        int st = stage;
        for (int i = 0; i < n0; ++i) {
          gate(i);
          ring_issue<TF_STAGES>(src, ringA, ringB, full_bar, st, i);
          if (++st == TF_STAGES) st = 0;
        }
      }
    }

    double acc[4][4][2];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
      const int lr = wm * 32 + mi * 8 + g;
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        const int lc = wn * 32 + ni * 8 + 2 * t;
        double v0 = 0.0, v1 = 0.0;
        if (row0 + lr < p.rows) {
          const double* src_c = p.B + (long long)(row0 + lr) * p.ldb + col0 + lc;
          if (lc + 1 < nb) {
            const double2 v = __ldcg(reinterpret_cast<const double2*>(src_c));
            v0 = v.x; v1 = v.y;
          } else if (lc < nb) {
            v0 = __ldcg(src_c);
          }
        }
        acc[mi][ni][0] = -v0;
        acc[mi][ni][1] = -v1;
      }
    }

    // PIPE (few rows): a 32-row slab entirely beyond the valid rows does no arithmetic (results are never stored)
    const bool slab_active = !PIPE || row0 + wm * 32 < p.rows;
    if (PIPE)
      mma_mainloop<TF_STAGES>(acc, src, ringA, ringB, full_bar, empty_bar, stage, phase, ktiles, wm, wn, lane, p.zero,
                              gate, slab_active);
    else
      mma_mainloop<TF_STAGES>(acc, src, ringA, ringB, full_bar, empty_bar, stage, phase, ktiles, wm, wn, lane, p.zero);

    // ---- diagonal step: V[r, J] = R * W_J^T on the tensor pipe -------------------------------------------
    __syncthreads();  // every warp has left the ring (each warp's last release waited for its fragment loads)
    if (tid == 0) {   // W_J (64 x 64) -> the B halves of the next 4 ring slots, one 16-wide k-tile each
      int st = stage;
#pragma unroll
      for (int kt = 0; kt < TF_STAGES; ++kt) {
        mbar_arrive_expect_tx(&full_bar[st], GEMM_B_STAGE_BYTES);
        tma_load_2d(ringB + st * GEMM_B_STAGE_BYTES, &tmW, kt * GEMM_BK, col0, &full_bar[st]);
        if (++st == TF_STAGES) st = 0;
      }
    }
    // R = -acc -> the A halves of the same slots, in the swizzled fragment layout mma_mainloop reads:
    // columns [16 s, 16 s + 16) in slot (stage + s) % 4, row lr at 128 lr, 16-byte chunk c at c ^ (lr & 7)
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
      const int lr = wm * 32 + mi * 8 + g;
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        int slot = stage + wn * 2 + (ni >> 1);
        if (slot >= TF_STAGES) slot -= TF_STAGES;
        const int chunk = ((ni & 1) * 4 + t) ^ g;   // lr & 7 == g
        double2* dst = reinterpret_cast<double2*>(ringA + slot * GEMM_A_STAGE_BYTES + lr * 128 + chunk * 16);
        *dst = make_double2(-acc[mi][ni][0], -acc[mi][ni][1]);
        acc[mi][ni][0] = 0.0;
        acc[mi][ni][1] = 0.0;
      }
    }
    __syncthreads();  // R is visible to every warp (generic proxy on both sides)
    mma_mainloop<TF_STAGES>(acc, src, ringA, ringB, full_bar, empty_bar, stage, phase, TF_STAGES, wm, wn, lane, p.zero,
                            NoGate(), slab_active);
    __syncthreads();  // ring drained again: its first 2 KiB are the row-sum scratch of the epilogue
    {
      DiagOut o;
      o.B = p.B + col0; o.ldb = p.ldb; o.rows = p.rows; o.ssq = p.ssq; o.kss = p.kss; o.var = p.var;
      o.J = J; o.col_blocks = p.col_blocks;
      diag_epilogue(acc, o, row0, nb, wm, wn, g, t, reinterpret_cast<double*>(ringA));
    }
    fence_proxy_async_all();
    __threadfence();
    __syncthreads();
    if (tid == 0) st_release_gpu(p.progress + r, J + 1);
  }
}

}  // namespace nngp
