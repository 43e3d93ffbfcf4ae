// Epsilon threshold of Hundman et al. as the reference applies it to the anomaly scores (eval_methods.py:186-236
// find_epsilon; called from prediction.py via epsilon_eval on A_Score_Global) on the device, where the single-pass
// scorer leaves the scores.  For the 19 candidates z = 2.5, 3.0, ..., 11.5:
//   eps = mean + z * sd;  pruned = scores < eps;  anomalies = scores >= eps, dilated by +-49 indices and clipped;
//   score(z) = ((mean - mean(pruned))/mean + (sd - sd(pruned))/sd) / denom(reg_level, #dilated)
// and the last z with score >= running maximum and #dilated < N/2 wins (max(scores) when none qualifies).
// np.mean / np.std semantics: population standard deviation, double accumulation.
#include "common.cuh"
#include "../../include/mtadgat.h"

namespace {
constexpr int NZ = 19, HALO = 49, TB = 256;

// scratch doubles: [0] sum, [1] sumsq, [2] max (as double);  per candidate c at 4 + 4c: cnt_below, sum_below, sumsq_below, n_dilated
__global__ void __launch_bounds__(TB) eps_moments_kernel(const float* __restrict__ e, long long N, double* __restrict__ sc) {
  __shared__ double red[3][TB / 32];
  double s = 0.0, q = 0.0, m = -INFINITY;
  for (long long i = (long long)blockIdx.x * TB + threadIdx.x; i < N; i += (long long)gridDim.x * TB) {
    const double v = (double)__ldg(e + i);
    s += v; q += v * v; m = fmax(m, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { red[0][w] = s; red[1][w] = q; red[2][w] = m; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ts = 0.0, tq = 0.0, tm = -INFINITY;
    for (int i = 0; i < TB / 32; ++i) { ts += red[0][i]; tq += red[1][i]; tm = fmax(tm, red[2][i]); }
    atomicAdd(sc + 0, ts); atomicAdd(sc + 1, tq);
    // max through an ordered-bit atomic on the double's bits (scores are finite; handles negatives)
    unsigned long long bits = (unsigned long long)__double_as_longlong(tm);
    bits = (bits >> 63) ? ~bits : (bits | 0x8000000000000000ull);
    atomicMax(reinterpret_cast<unsigned long long*>(sc + 2), bits);
  }
}

__device__ __forceinline__ double decode_max(double raw) {
  unsigned long long bits = (unsigned long long)__double_as_longlong(raw);
  bits = (bits >> 63) ? (bits & 0x7FFFFFFFFFFFFFFFull) : ~bits;
  return __longlong_as_double((long long)bits);
}

// grid (chunks, NZ): block = TB consecutive indices + HALO on both sides in shared memory
__global__ void __launch_bounds__(TB) eps_candidates_kernel(const float* __restrict__ e, long long N, double* __restrict__ sc) {
  __shared__ float tile[TB + 2 * HALO];
  __shared__ double red[4][TB / 32];
  const int c = blockIdx.y;
  const double mean = sc[0] / (double)N;
  const double var = fmax(sc[1] / (double)N - mean * mean, 0.0);
  const double eps = mean + sqrt(var) * (2.5 + 0.5 * c);
  const long long base = (long long)blockIdx.x * TB;
  for (int t = threadIdx.x; t < TB + 2 * HALO; t += TB) {
    const long long i = base + t - HALO;
    tile[t] = (i >= 0 && i < N) ? __ldg(e + i) : -INFINITY;           // -inf is never an anomaly
  }
  __syncthreads();
  const long long i = base + threadIdx.x;
  double cnt = 0.0, s = 0.0, q = 0.0, dil = 0.0;
  if (i < N) {
    const double v = (double)tile[threadIdx.x + HALO];
    if (v < eps) { cnt = 1.0; s = v; q = v * v; }
    bool any = false;
#pragma unroll 11
    for (int t = 0; t < 2 * HALO + 1; ++t) any |= ((double)tile[threadIdx.x + t] >= eps);
    dil = any ? 1.0 : 0.0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o); s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o); dil += __shfl_xor_sync(0xffffffffu, dil, o);
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { red[0][w] = cnt; red[1][w] = s; red[2][w] = q; red[3][w] = dil; }
  __syncthreads();
  if (threadIdx.x < 4) {
    double t = 0.0;
    for (int k = 0; k < TB / 32; ++k) t += red[threadIdx.x][k];
    if (t != 0.0) atomicAdd(sc + 4 + 4 * c + threadIdx.x, t);
  }
}

__global__ void eps_select_kernel(const double* __restrict__ sc, long long N, int reg_level, float* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double mean = sc[0] / (double)N;
  const double sd = sqrt(fmax(sc[1] / (double)N - mean * mean, 0.0));
  double best = NAN, max_score = -10000000.0; int best_c = -1;
  for (int c = 0; c < NZ; ++c) {
    const double cnt = sc[4 + 4 * c], s = sc[5 + 4 * c], q = sc[6 + 4 * c], nd = sc[7 + 4 * c];
    if (!(nd > 0.0)) continue;
    const double pm = s / cnt;                                         // NaN when nothing is below eps, as np.mean([])
    const double psd = sqrt(fmax(q / cnt - pm * pm, 0.0));
    const double denom = reg_level == 0 ? 1.0 : (reg_level == 1 ? nd : nd * nd);
    const double score = ((mean - pm) / mean + (sd - psd) / sd) / denom;
    if (score >= max_score && nd < 0.5 * (double)N) { max_score = score; best = mean + sd * (2.5 + 0.5 * c); best_c = c; }
  }
  if (best_c < 0) best = decode_max(sc[2]);
  out[0] = (float)best; out[1] = best_c < 0 ? -1.f : (float)(2.5 + 0.5 * best_c); out[2] = (float)max_score;
}
}  // namespace

extern "C" long long mtadgat_find_epsilon_scratch_doubles(void) { return 4 + 4 * NZ; }

extern "C" int mtadgat_find_epsilon(const float* scores, long long n_scores, int reg_level, float* out, double* scratch,
                                    void* stream) {
  MG_CHECK_ARG(scores && out && scratch && n_scores > 0 && reg_level >= 0 && reg_level <= 2, "find_epsilon: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  MG_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double) * (4 + 4 * NZ), s));
  eps_moments_kernel<<<(int)min((long long)592, (n_scores + TB - 1) / TB), TB, 0, s>>>(scores, n_scores, scratch);
  MG_COUNT_LAUNCH();
  eps_candidates_kernel<<<dim3((unsigned)((n_scores + TB - 1) / TB), NZ), TB, 0, s>>>(scores, n_scores, scratch);
  MG_COUNT_LAUNCH();
  eps_select_kernel<<<1, 32, 0, s>>>(scratch, n_scores, reg_level, out);
  MG_COUNT_LAUNCH();
  MG_CHECK_LAUNCH("find_epsilon");
  return MTADGAT_OK;
}
