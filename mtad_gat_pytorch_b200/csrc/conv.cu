// ConvLayer (reference modules.py:5-22): zero-pad (ks-1)/2 each side, Conv1d over time mixing all k
// features, ReLU -- computed as an implicit GEMM that reads and writes the (B,n,k) window layout directly
// (the reference's two permutes and the padded copy never exist).
#include "gemm.cuh"
#include "../../include/mtadgat.h"

namespace {

// A(m=(b,t), kk=(tau,c)) = src[b, t + sgn*(tau - pad), c]  (0 outside the window), optionally masked by y>0
template <bool MASKED>
struct ConvShiftLoad {
  static constexpr bool fast_second = true;
  const float* src; const float* y; int n, k, pad, sgn;
  __device__ __forceinline__ float operator()(int, int m, int kk) const {
    int b = m / n, t = m - b * n;
    int tau = kk / k, c = kk - tau * k;
    int ts = t + sgn * (tau - pad);
    if (ts < 0 || ts >= n) return 0.f;
    long long o = ((long long)b * n + ts) * k + c;
    float v = __ldg(src + o);
    if (MASKED) v = (__ldg(y + o) > 0.f) ? v : 0.f;
    return v;
  }
};

// B(kk=(tau,ci), n=co) = w[co, ci, tau]           (forward)
struct ConvWFwd {
  static constexpr bool fast_second = false;
  const float* w; int k, ks;
  __device__ __forceinline__ float operator()(int, int kk, int co) const {
    int tau = kk / k, ci = kk - tau * k;
    return __ldg(w + ((long long)co * k + ci) * ks + tau);
  }
};
// B(kk=(tau,co), n=ci) = w[co, ci, tau]           (data gradient)
struct ConvWBwd {
  static constexpr bool fast_second = false;
  const float* w; int k, ks;
  __device__ __forceinline__ float operator()(int, int kk, int ci) const {
    int tau = kk / k, co = kk - tau * k;
    return __ldg(w + ((long long)co * k + ci) * ks + tau);
  }
};
// A(m=co, kk=(b,t)) = dy[(b,t),co] * (y>0)        (weight gradient)
struct DpreT {
  static constexpr bool fast_second = false;
  const float* dy; const float* y; int k;
  __device__ __forceinline__ float operator()(int, int co, int kk) const {
    long long o = (long long)kk * k + co;
    return (__ldg(y + o) > 0.f) ? __ldg(dy + o) : 0.f;
  }
};
// B(kk=(b,t), n=(tau,ci)) = xpad[b, t+tau-pad, ci]
struct ConvXCols {
  static constexpr bool fast_second = true;
  const float* x; int n, k, pad;
  __device__ __forceinline__ float operator()(int, int kk, int nn) const {
    int b = kk / n, t = kk - b * n;
    int tau = nn / k, ci = nn - tau * k;
    int ts = t + tau - pad;
    if (ts < 0 || ts >= n) return 0.f;
    return __ldg(x + ((long long)b * n + ts) * k + ci);
  }
};
// C(m=co, n=(tau,ci)) -> dw[co, ci, tau]
struct StoreDW {
  float* dw; int k, ks;
  __device__ __forceinline__ void operator()(int, int co, int nn, float v, bool) const {
    int tau = nn / k, ci = nn - tau * k;
    atomicAdd(dw + ((long long)co * k + ci) * ks + tau, v);
  }
};
struct DpreCols {
  static constexpr bool fast_second = true;
  const float* dy; const float* y; int k;
  __device__ __forceinline__ float operator()(int, int m, int c) const {
    long long o = (long long)m * k + c;
    return (__ldg(y + o) > 0.f) ? __ldg(dy + o) : 0.f;
  }
};

}  // namespace

extern "C" int mtadgat_conv_relu_fwd(const float* x, const float* w, const float* bias, float* y, int B, int n,
                                     int k, int ks, void* stream) {
  MG_CHECK_ARG(x && w && bias && y, "conv_relu_fwd: null pointer");
  MG_CHECK_ARG(B > 0 && n > 0 && k > 0 && ks > 0 && (ks & 1), "conv_relu_fwd: need B,n,k>0 and odd kernel_size (got %d)", ks);
  cudaStream_t s = (cudaStream_t)stream;
  ConvShiftLoad<false> A{x, nullptr, n, k, (ks - 1) / 2, +1};
  ConvWFwd Bw{w, k, ks};
  StStrided C{y, 0, k, 1, bias, ACT_RELU, 0};
  launch_gemm_batched(1, B * n, k, ks * k, A, Bw, C, s);
  MG_CHECK_LAUNCH("conv_relu_fwd");
  return MTADGAT_OK;
}

extern "C" int mtadgat_conv_relu_bwd(const float* x, const float* w, const float* y, const float* dy, float* dx,
                                     float* dw, float* db, int B, int n, int k, int ks, void* stream) {
  MG_CHECK_ARG(x && w && y && dy && dw && db, "conv_relu_bwd: null pointer");
  MG_CHECK_ARG(B > 0 && n > 0 && k > 0 && ks > 0 && (ks & 1), "conv_relu_bwd: bad shape");
  cudaStream_t s = (cudaStream_t)stream;
  const int pad = (ks - 1) / 2;
  if (dx) {
    // dx[b,t',ci] = sum_{tau,co} dpre[b, t'-tau+pad, co] * w[co,ci,tau]
    ConvShiftLoad<true> A{dy, y, n, k, pad, -1};
    ConvWBwd Bw{w, k, ks};
    StStrided C{dx, 0, k, 1, nullptr, ACT_NONE, 0};
    launch_gemm_batched(1, B * n, k, ks * k, A, Bw, C, s);
  }
  MG_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)k * k * ks, s));
  MG_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * (size_t)k, s));
  {
    DpreT A{dy, y, k};
    ConvXCols Bx{x, n, k, pad};
    StoreDW C{dw, k, ks};
    launch_gemm_splitk(k, ks * k, B * n, A, Bx, C, s);
  }
  launch_colsum(B * n, k, DpreCols{dy, y, k}, db, s);
  MG_CHECK_LAUNCH("conv_relu_bwd");
  return MTADGAT_OK;
}
