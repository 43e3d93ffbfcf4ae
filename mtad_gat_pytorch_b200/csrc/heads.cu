// Linear (+ReLU +dropout) used by Forecasting_Model (reference modules.py:286-311) and the reconstruction
// head's fc (modules.py:282).  y = dropout(act(x W^T + b)); dropout multipliers are regenerated from the
// Philox stream in backward, so only y is kept.
#include "gemm.cuh"
#include "../../include/mtadgat.h"

namespace {

struct StLinear {
  float* y; int O; const float* bias; int act; float p, inv_keep; const unsigned long long* seed; uint32_t stream;
  __device__ __forceinline__ void operator()(int, int m, int o, float v, bool) const {
    v += __ldg(bias + o);
    if (act == ACT_RELU) v = fmaxf(v, 0.f);
    if (p > 0.f) v *= dropout_mult(seed, stream, (unsigned long long)m * O + o, p, inv_keep);
    y[(long long)m * O + o] = v;
  }
};

// dpre[m][o] = dy * mult * (act ? y>0 : 1)
struct DpreLin {
  const float* dy; const float* y; int O; int act; float p, inv_keep; const unsigned long long* seed; uint32_t stream;
  __device__ __forceinline__ float get(int m, int o) const {
    long long i = (long long)m * O + o;
    float v = __ldg(dy + i);
    if (act == ACT_RELU && !(__ldg(y + i) > 0.f)) return 0.f;
    if (p > 0.f) v *= dropout_mult(seed, stream, (unsigned long long)i, p, inv_keep);
    return v;
  }
};
struct DpreA : DpreLin {   // A(m, kk=o)
  static constexpr bool fast_second = true;
  __device__ __forceinline__ float operator()(int, int m, int o) const { return get(m, o); }
};
struct DpreAT : DpreLin {  // A(m=o, kk=row)
  static constexpr bool fast_second = false;
  __device__ __forceinline__ float operator()(int, int o, int m) const { return get(m, o); }
};

}  // namespace

extern "C" int mtadgat_linear_fwd(const float* x, const float* w, const float* b, float* y, int M, int I, int O,
                                  int act, float p_drop, const unsigned long long* seed, unsigned int rng_stream,
                                  void* stream) {
  MG_CHECK_ARG(x && w && b && y, "linear_fwd: null pointer");
  MG_CHECK_ARG(M > 0 && I > 0 && O > 0, "linear_fwd: bad shape");
  MG_CHECK_ARG(p_drop >= 0.f && p_drop < 1.f && (p_drop == 0.f || seed), "linear_fwd: bad dropout arguments");
  cudaStream_t s = (cudaStream_t)stream;
  // A(m,kk) = x[m*I+kk] ; B(kk, o) = w[o*I + kk]
  launch_gemm_batched(1, M, O, I, Strided2<true>{x, 0, I, 1}, Strided2<false>{w, 0, 1, I},
                      StLinear{y, O, b, act, p_drop, 1.f / (1.f - p_drop), seed, rng_stream}, s);
  MG_CHECK_LAUNCH("linear_fwd");
  return MTADGAT_OK;
}

extern "C" int mtadgat_linear_bwd(const float* x, const float* w, const float* y, const float* dy, float* dx,
                                  int dx_accumulate, float* dw, float* db, int M, int I, int O, int act, float p_drop,
                                  const unsigned long long* seed, unsigned int rng_stream, void* stream) {
  MG_CHECK_ARG(x && w && y && dy && dw && db, "linear_bwd: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  const float inv_keep = 1.f / (1.f - p_drop);
  DpreA A; A.dy = dy; A.y = y; A.O = O; A.act = act; A.p = p_drop; A.inv_keep = inv_keep; A.seed = seed; A.stream = rng_stream;
  DpreAT At; At.dy = dy; At.y = y; At.O = O; At.act = act; At.p = p_drop; At.inv_keep = inv_keep; At.seed = seed; At.stream = rng_stream;
  if (dx) {
    // dx = dpre W : B(kk=o, n=i) = w[o*I + i]
    launch_gemm_batched(1, M, I, O, A, Strided2<true>{w, 0, I, 1}, StStrided{dx, 0, I, 1, nullptr, ACT_NONE, dx_accumulate}, s);
  }
  MG_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)O * I, s));
  MG_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * (size_t)O, s));
  // dw[o][i] = sum_m dpre[m][o] x[m][i]
  launch_gemm_splitk(O, I, M, At, Strided2<true>{x, 0, I, 1}, StStrided{dw, 0, I, 1, nullptr, ACT_NONE, 0}, s);
  launch_colsum(M, O, A, db, s);
  MG_CHECK_LAUNCH("linear_bwd");
  return MTADGAT_OK;
}
