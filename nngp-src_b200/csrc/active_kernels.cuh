// active_kernels.cuh -- the variance-driven query selection of the active-learning loop, on the device
// (reference: active/ActiveLearner.py:43-55).
//
//     score_i = sqrt(var_i) / max_j(mean_j)                                   (ActiveLearner.py:46-47)
//     mode TOPK   : the `budget` largest scores            == np.argsort(score)[-budget:]            (:54)
//     mode SAMPLE : `budget` draws without replacement, p_i = score_i / sum(score)                   (:49-53)
//                   as Gumbel-top-k: the `budget` largest  log(score_i) + G_i,  G_i = -log(-log u_i)
//                   (the same construction jax.random.choice uses; the uniform stream here is a counter-based
//                   splitmix64 of (seed, i), not JAX's threefry -- that branch is unpinned)
//
// Selection is an exact radix select over the 96-bit composite (order-preserving 64-bit image of the key, 32-bit
// row index): 12 passes of an 8-bit histogram restricted to the current prefix, each followed by a one-thread
// "pick the bin" step that stays on the device, then one compaction.  Ties are therefore broken by row index
// (the larger index wins), i.e. the result is the tail of a stable ascending sort -- one of the orders
// np.argsort may return.  HBM-bound and tiny next to the prediction that produces mean / var.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nngp {

struct SelectState {       // lives in device memory, zero-initialised before the first pass
  unsigned long long prefix_key;   // decided high digits of the threshold key
  unsigned int prefix_idx;         // decided high digits of the threshold index
  unsigned int pad;
  long long k_remaining;           // how many elements of the current prefix class are still to be taken
  unsigned int hist[256];
  unsigned int out_count;
  int bad;                         // SAMPLE mode: a negative / non-finite score was seen
};

// max over a vector (np.max(pred_mean, 0)); NaN propagates like numpy.
__global__ void max_reduce_kernel(const double* __restrict__ v, long long n, double* out) {
  __shared__ double red[32];
  __shared__ int any_nan;
  if (threadIdx.x == 0) any_nan = 0;
  __syncthreads();
  double m = -INFINITY;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const double x = v[i];
    if (x != x) any_nan = 1;
    m = fmax(m, x);
  }
  for (int o = 16; o; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : -INFINITY;
    for (int o = 16; o; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (threadIdx.x == 0) *out = any_nan ? NAN : m;
  }
}

__device__ __forceinline__ unsigned long long orderable_u64(double x) {
  if (x != x) return ~0ull;  // NaN sorts last, like numpy
  const unsigned long long b = (unsigned long long)__double_as_longlong(x);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

__device__ __forceinline__ double splitmix_uniform(unsigned long long seed, unsigned long long i) {
  unsigned long long z = seed + (i + 1ull) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return ((double)(z >> 12) + 0.5) * 0x1.0p-52;  // 52 bits, so the +0.5 is exact: strictly inside (0, 1)
}

// score_i = sqrt(var_i) / maxmean;  key_i = image of score_i (TOPK) or of log(score_i) + Gumbel_i (SAMPLE)
__global__ void score_key_kernel(const double* __restrict__ var, const double* __restrict__ maxmean, long long n,
                                 int mode, unsigned long long seed, double* __restrict__ score,
                                 unsigned long long* __restrict__ key, SelectState* st) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double s = sqrt(var[i]) / *maxmean;
  score[i] = s;
  if (mode == 0) {
    key[i] = orderable_u64(s);
  } else {
    if (!(s >= 0.0) || isinf(s)) st->bad = 1;
    const double u = splitmix_uniform(seed, (unsigned long long)i);
    const double gk = (s > 0.0) ? log(s) - log(-log(u)) : -INFINITY;  // p_i == 0 is never drawn
    key[i] = orderable_u64(gk);
  }
}

// digit `pass` (0 = most significant) of the composite (key:64, idx:32): 12 digits of 8 bits
__device__ __forceinline__ unsigned int composite_digit(unsigned long long key, unsigned int idx, int pass) {
  return pass < 8 ? (unsigned int)(key >> (56 - 8 * pass)) & 255u : (idx >> (24 - 8 * (pass - 8))) & 255u;
}
__device__ __forceinline__ bool composite_prefix_matches(unsigned long long key, unsigned int idx, int pass,
                                                         unsigned long long pk, unsigned int pi) {
  if (pass == 0) return true;
  if (pass <= 8) return pass == 8 ? key == pk : (key >> (64 - 8 * pass)) == (pk >> (64 - 8 * pass));
  return key == pk && (idx >> (32 - 8 * (pass - 8))) == (pi >> (32 - 8 * (pass - 8)));
}

__global__ void select_hist_kernel(const unsigned long long* __restrict__ key, long long n, int pass, SelectState* st) {
  __shared__ unsigned int h[256];
  for (int b = threadIdx.x; b < 256; b += blockDim.x) h[b] = 0;
  __syncthreads();
  const unsigned long long pk = st->prefix_key;
  const unsigned int pi = st->prefix_idx;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const unsigned long long k = key[i];
    if (composite_prefix_matches(k, (unsigned int)i, pass, pk, pi)) atomicAdd(&h[composite_digit(k, (unsigned int)i, pass)], 1u);
  }
  __syncthreads();
  for (int b = threadIdx.x; b < 256; b += blockDim.x)
    if (h[b]) atomicAdd(&st->hist[b], h[b]);
}

// one thread: walk the bins from the top until the k-th largest element of the prefix class is covered
__global__ void select_pick_kernel(int pass, SelectState* st) {
  long long need = st->k_remaining;
  int b = 255;
  for (; b > 0; --b) {
    const long long c = st->hist[b];
    if (c >= need) break;
    need -= c;
  }
  st->k_remaining = need;
  if (pass < 8) st->prefix_key |= (unsigned long long)b << (56 - 8 * pass);
  else st->prefix_idx |= (unsigned int)b << (24 - 8 * (pass - 8));
  for (int i = 0; i < 256; ++i) st->hist[i] = 0;
}

// everything >= the threshold composite: exactly k rows (composites are distinct)
__global__ void select_compact_kernel(const unsigned long long* __restrict__ key, long long n, SelectState* st,
                                      unsigned long long* __restrict__ out_key, unsigned int* __restrict__ out_idx) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long k = key[i];
  const unsigned long long pk = st->prefix_key;
  if (k > pk || (k == pk && (unsigned int)i >= st->prefix_idx)) {
    const unsigned int pos = atomicAdd(&st->out_count, 1u);
    out_key[pos] = k;
    out_idx[pos] = (unsigned int)i;
  }
}

}  // namespace nngp
