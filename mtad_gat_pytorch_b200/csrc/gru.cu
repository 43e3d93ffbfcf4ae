// GRULayer / RNNDecoder recurrences (reference modules.py:220-257; torch.nn.GRU gate order r,z,n, h0 = 0).
//
// Baseline path: the input projection of all n steps is one GEMM (reading the three (B,n,k) tensors that the
// reference concatenates at mtad_gat.py:71 as three K-slices -- the concat never exists), then a persistent
// recurrent kernel keeps a tile of BT windows' hidden state in shared memory for all n steps.
// The decoder's input is the reference's scrambled repeat of h_end (modules.py:279):
//   rep[b,t,c] = h_end[b,(t*H+c)//n]   =>   W_ih rep[b,t,:] = sum_j h_end[b,m0(t)+j] * S[t,j,:]
// with S[t,j,g] = sum_{c:(t*H+c)//n = m0(t)+j} W_ih[g,c]  -- J = O(H/n + 2) FMAs per gate instead of H.
#include "gemm.cuh"
#include "../../include/mtadgat.h"

// tensor-core recurrence (gru_tc.cu)
int mtadgat_gru_tc_supported(int H);
int mtadgat_gru_tc_fwd_launch(const float* gi, const float* S, const float* hsrc, const float* b_ih, int J, int Hs,
                              const float* w_hh, const float* b_hh, float* out, float* h_last, float* gates, int B,
                              int n, int H, cudaStream_t s);
int mtadgat_gru_tc_bwd_supported(int H);
int mtadgat_gru_tc_bwd_launch(const float* gates, const float* out, const float* w_hh, const float* dout,
                              const float* dh_last, unsigned int* gmax_bits, float* dgi, float* dghn, int B, int n,
                              int H, cudaStream_t s);
static int g_gru_impl = 1;   // 0 = fp32 SIMT recurrence, 1 = tcgen05 fp16-operand recurrence (fp32 accumulate/state)

namespace {

constexpr int BT = 8;   // windows per CTA in the recurrent kernels

// ---- functors ----------------------------------------------------------------------------------
// A(m=(b,t), kk) over three column slices
struct Cat3A {
  static constexpr bool fast_second = true;
  const float *x0, *x1, *x2; int k0, k1, k2;
  __device__ __forceinline__ float operator()(int, int m, int kk) const {
    if (kk < k0) return __ldg(x0 + (long long)m * k0 + kk);
    kk -= k0;
    if (kk < k1) return __ldg(x1 + (long long)m * k1 + kk);
    kk -= k1;
    return __ldg(x2 + (long long)m * k2 + kk);
  }
};
// B(kk, n=g) = W[g, kk]  for a row-major (G, I) weight
struct WT {
  static constexpr bool fast_second = false;
  const float* w; int I;
  __device__ __forceinline__ float operator()(int, int kk, int g) const { return __ldg(w + (long long)g * I + kk); }
};
// store into three column slices (data gradient of the concatenated input)
struct StCat3 {
  float *d0, *d1, *d2; int k0, k1, k2; int acc0, acc1, acc2;
  __device__ __forceinline__ void operator()(int, int m, int kk, float v, bool) const {
    float* q; int acc;
    if (kk < k0) { q = d0 ? d0 + (long long)m * k0 + kk : nullptr; acc = acc0; }
    else if (kk < k0 + k1) { q = d1 ? d1 + (long long)m * k1 + (kk - k0) : nullptr; acc = acc1; }
    else { q = d2 ? d2 + (long long)m * k2 + (kk - k0 - k1) : nullptr; acc = acc2; }
    if (!q) return;
    *q = acc ? (*q + v) : v;
  }
};
// A(m=g, kk=(b,t)) = dgi[(b,t), g]           (for dW_ih)
struct DgiT {
  static constexpr bool fast_second = false;
  const float* dgi; int G;
  __device__ __forceinline__ float operator()(int, int g, int kk) const { return __ldg(dgi + (long long)kk * G + g); }
};
// A(m=g, kk=(b,t)) = dgh[(b,t), g] : first 2H columns from dgi, last H from dghn   (for dW_hh)
struct DghT {
  static constexpr bool fast_second = false;
  const float* dgi; const float* dghn; int H;
  __device__ __forceinline__ float operator()(int, int g, int kk) const {
    return g < 2 * H ? __ldg(dgi + (long long)kk * 3 * H + g) : __ldg(dghn + (long long)kk * H + (g - 2 * H));
  }
};
struct DghCols {
  static constexpr bool fast_second = true;
  const float* dgi; const float* dghn; int H;
  __device__ __forceinline__ float operator()(int, int m, int g) const {
    return g < 2 * H ? __ldg(dgi + (long long)m * 3 * H + g) : __ldg(dghn + (long long)m * H + (g - 2 * H));
  }
};
// B(kk=(b,t), n) = x_cat[(b,t), n]
struct Cat3B {
  static constexpr bool fast_second = true;
  const float *x0, *x1, *x2; int k0, k1, k2;
  __device__ __forceinline__ float operator()(int, int m, int kk) const {
    if (kk < k0) return __ldg(x0 + (long long)m * k0 + kk);
    kk -= k0;
    if (kk < k1) return __ldg(x1 + (long long)m * k1 + kk);
    kk -= k1;
    return __ldg(x2 + (long long)m * k2 + kk);
  }
};
// B(kk=(b,t), n=u) = h_{t-1}[b,u] = out[b,t-1,u] (0 at t = 0)
struct HprevB {
  static constexpr bool fast_second = true;
  const float* out; int n, H;
  __device__ __forceinline__ float operator()(int, int kk, int u) const {
    int t = kk % n;
    return t == 0 ? 0.f : __ldg(out + (long long)(kk - 1) * H + u);
  }
};

__global__ void transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int R, int C) {
  // dst[c][r] = src[r][c]
  __shared__ float tile[32][33];
  int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int r = r0 + i, c = c0 + threadIdx.x;
    if (r < R && c < C) tile[i][threadIdx.x] = src[(long long)r * C + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i, r = r0 + threadIdx.x;
    if (r < R && c < C) dst[(long long)c * R + r] = tile[threadIdx.x][i];
  }
}

// ---- decoder input shortcut -----------------------------------------------------------------------
// m0[t] = (t*Hs)//n ;  S[t][j][g] = sum_{c in [0,Hs): (t*Hs+c)//n == m0[t]+j} w_ih[g][c]
__global__ void rep_build_S_kernel(const float* __restrict__ w_ih, int n, int Hs, int G, int J, float* __restrict__ S) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * J * G) return;
  int g = idx % G, j = (idx / G) % J, t = idx / (G * J);
  long long base = (long long)t * Hs;
  int m = (int)(base / n) + j;
  // c range with (base + c) / n == m  <=>  m*n <= base + c < (m+1)*n
  long long lo = (long long)m * n - base, hi = (long long)(m + 1) * n - base;
  int clo = (int)max(lo, 0LL), chi = (int)min(hi, (long long)Hs);
  float acc = 0.f;
  for (int c = clo; c < chi; ++c) acc += w_ih[(long long)g * Hs + c];
  S[idx] = acc;
}
// dW_ih[g][c] = sum_t dS[t][(t*Hs+c)//n - m0[t]][g]
__global__ void rep_dw_kernel(const float* __restrict__ dS, int n, int Hs, int G, int J, float* __restrict__ dw_ih) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= G * Hs) return;
  int c = idx % Hs, g = idx / Hs;
  float acc = 0.f;
  for (int t = 0; t < n; ++t) {
    long long base = (long long)t * Hs;
    int j = (int)((base + c) / n) - (int)(base / n);
    acc += dS[((long long)t * J + j) * G + g];
  }
  dw_ih[idx] = acc;
}
// dS[t][j][g] = sum_b dgi[b,t,g] * hsrc[b, m0[t]+j]
__global__ void rep_dS_kernel(const float* __restrict__ dgi, const float* __restrict__ hsrc, int B, int n, int Hs, int G,
                              int J, float* __restrict__ dS) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * J * G) return;
  int g = idx % G, j = (idx / G) % J, t = idx / (G * J);
  int m = (int)(((long long)t * Hs) / n) + j;
  float acc = 0.f;
  if (m < Hs)
    for (int b = 0; b < B; ++b) acc += dgi[((long long)b * n + t) * G + g] * hsrc[(long long)b * Hs + m];
  dS[idx] = acc;
}
// dhsrc[b][m] (+)= sum_{t,j: m0[t]+j == m} sum_g dgi[b,t,g] S[t][j][g]      one warp per (b,m)
__global__ void rep_dh_kernel(const float* __restrict__ dgi, const float* __restrict__ S, int B, int n, int Hs, int G,
                              int J, float* __restrict__ dh, int accumulate) {
  int wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (wid >= B * Hs) return;
  int m = wid % Hs, b = wid / Hs;
  // t such that m0[t] <= m <= m0[t]+J-1, i.e. (t*Hs)//n in [m-J+1, m]
  float acc = 0.f;
  int tlo = (int)(((long long)max(m - J + 1, 0) * n) / Hs);
  int thi = (int)min((long long)n - 1, (((long long)(m + 1) * n) / Hs));
  for (int t = tlo; t <= thi; ++t) {
    int j = m - (int)(((long long)t * Hs) / n);
    if (j < 0 || j >= J) continue;
    const float* dg = dgi + ((long long)b * n + t) * G;
    const float* s = S + ((long long)t * J + j) * G;
    for (int g = lane; g < G; g += 32) acc = fmaf(dg[g], s[g], acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) dh[wid] = accumulate ? dh[wid] + acc : acc;
}

// ---- recurrent forward ------------------------------------------------------------------------------
struct GruFwdParams {
  const float* gi;        // (B,n,3H) incl. b_ih, or nullptr in rep mode
  const float* S; const float* hsrc; const float* b_ih; int J, Hs;   // rep mode
  const float* wt;        // (H, 3H) = W_hh^T
  const float* b_hh;
  float* out; float* h_last; float* gates;   // out (B,n,H) / gates (B,n,4H) may be null
  int B, n, H;
};

__global__ void __launch_bounds__(1024) gru_fwd_kernel(GruFwdParams P) {
  extern __shared__ __align__(16) float smem[];
  const int H = P.H, G = 3 * H, n = P.n;
  float* hs = smem;                 // [H][BT]
  float* ghs = hs + (size_t)H * BT; // [G][BT+1]
  const int tid = threadIdx.x, nth = blockDim.x;
  const int b0 = blockIdx.x * BT;
  for (int i = tid; i < H * BT; i += nth) hs[i] = 0.f;
  __syncthreads();
  for (int t = 0; t < n; ++t) {
    // phase 1: gh = W_hh h + b_hh
    for (int g = tid; g < G; g += nth) {
      float acc[BT];
      float bb = __ldg(P.b_hh + g);
#pragma unroll
      for (int w = 0; w < BT; ++w) acc[w] = bb;
      const float* wcol = P.wt + g;
#pragma unroll 4
      for (int kk = 0; kk < H; ++kk) {
        float wv = __ldg(wcol + (size_t)kk * G);
        float4 h0 = *reinterpret_cast<const float4*>(hs + kk * BT);
        float4 h1 = *reinterpret_cast<const float4*>(hs + kk * BT + 4);
        acc[0] = fmaf(wv, h0.x, acc[0]); acc[1] = fmaf(wv, h0.y, acc[1]);
        acc[2] = fmaf(wv, h0.z, acc[2]); acc[3] = fmaf(wv, h0.w, acc[3]);
        acc[4] = fmaf(wv, h1.x, acc[4]); acc[5] = fmaf(wv, h1.y, acc[5]);
        acc[6] = fmaf(wv, h1.z, acc[6]); acc[7] = fmaf(wv, h1.w, acc[7]);
      }
#pragma unroll
      for (int w = 0; w < BT; ++w) ghs[g * (BT + 1) + w] = acc[w];
    }
    __syncthreads();
    // phase 2: gates + state update
    for (int idx = tid; idx < H * BT; idx += nth) {
      int u = idx % H, w = idx / H;
      int b = b0 + w;
      if (b >= P.B) continue;
      float gr, gz, gn;
      if (P.gi) {
        const float* gp = P.gi + ((size_t)b * n + t) * G;
        gr = __ldg(gp + u); gz = __ldg(gp + H + u); gn = __ldg(gp + 2 * H + u);
      } else {
        gr = __ldg(P.b_ih + u); gz = __ldg(P.b_ih + H + u); gn = __ldg(P.b_ih + 2 * H + u);
        int m0 = (int)(((long long)t * P.Hs) / n);
        for (int j = 0; j < P.J; ++j) {
          int m = m0 + j;
          if (m >= P.Hs) break;
          float hv = __ldg(P.hsrc + (size_t)b * P.Hs + m);
          const float* sp = P.S + ((size_t)t * P.J + j) * G;
          gr = fmaf(hv, __ldg(sp + u), gr); gz = fmaf(hv, __ldg(sp + H + u), gz); gn = fmaf(hv, __ldg(sp + 2 * H + u), gn);
        }
      }
      float r = sigmoidf_(gr + ghs[u * (BT + 1) + w]);
      float z = sigmoidf_(gz + ghs[(H + u) * (BT + 1) + w]);
      float hn = ghs[(2 * H + u) * (BT + 1) + w];
      float nn = tanhf_(gn + r * hn);
      float hp = hs[u * BT + w];
      float hnew = (1.f - z) * nn + z * hp;
      hs[u * BT + w] = hnew;
      size_t o = (size_t)b * n + t;
      if (P.out) P.out[o * H + u] = hnew;
      if (P.gates) {
        float* gp = P.gates + o * 4 * H;
        gp[u] = r; gp[H + u] = z; gp[2 * H + u] = nn; gp[3 * H + u] = hn;
      }
      if (t == n - 1 && P.h_last) P.h_last[(size_t)b * H + u] = hnew;
    }
    __syncthreads();
  }
}

// ---- recurrent backward (BPTT) ------------------------------------------------------------------------
struct GruBwdParams {
  const float* gates; const float* out; const float* w_hh;   // w_hh (3H,H) native layout
  const float* dout; const float* dh_last;                    // either may be null
  float* dgi; float* dghn;                                    // (B,n,3H), (B,n,H)
  int B, n, H;
};

__global__ void __launch_bounds__(1024) gru_bwd_kernel(GruBwdParams P) {
  extern __shared__ __align__(16) float smem[];
  const int H = P.H, G = 3 * H, n = P.n;
  float* dhs = smem;                  // [H][BT]
  float* dg = dhs + (size_t)H * BT;   // [G][BT]
  const int tid = threadIdx.x, nth = blockDim.x;
  const int b0 = blockIdx.x * BT;
  for (int idx = tid; idx < H * BT; idx += nth) {
    int u = idx % H, w = idx / H, b = b0 + w;
    dhs[u * BT + w] = (P.dh_last && b < P.B) ? __ldg(P.dh_last + (size_t)b * H + u) : 0.f;
  }
  __syncthreads();
  for (int t = n - 1; t >= 0; --t) {
    for (int idx = tid; idx < H * BT; idx += nth) {
      int u = idx % H, w = idx / H, b = b0 + w;
      float dpr = 0.f, dpz = 0.f, dghn_ = 0.f, dhz = 0.f;
      if (b < P.B) {
        size_t o = (size_t)b * n + t;
        float dh = dhs[u * BT + w];
        if (P.dout) dh += __ldg(P.dout + o * H + u);
        const float* gp = P.gates + o * 4 * H;
        float r = __ldg(gp + u), z = __ldg(gp + H + u), nn = __ldg(gp + 2 * H + u), hn = __ldg(gp + 3 * H + u);
        float hp = t > 0 ? __ldg(P.out + (o - 1) * H + u) : 0.f;
        float dn = dh * (1.f - z);
        float dz = dh * (hp - nn);
        float dpn = dn * (1.f - nn * nn);
        dpz = dz * z * (1.f - z);
        float dr = dpn * hn;
        dpr = dr * r * (1.f - r);
        dghn_ = dpn * r;
        dhz = dh * z;
        float* q = P.dgi + o * G;
        q[u] = dpr; q[H + u] = dpz; q[2 * H + u] = dpn;
        P.dghn[o * H + u] = dghn_;
      }
      dg[u * BT + w] = dpr; dg[(H + u) * BT + w] = dpz; dg[(2 * H + u) * BT + w] = dghn_;
      dhs[u * BT + w] = dhz;
    }
    __syncthreads();
    // dh_prev[u] += sum_g dgh[g] W_hh[g][u]; thread = (part, u), part splits the g range in 3
    for (int it = tid; it < 3 * H; it += nth) {
      int part = it / H, u = it - part * H;
      float acc[BT];
#pragma unroll
      for (int w = 0; w < BT; ++w) acc[w] = 0.f;
      const float* wp = P.w_hh + (size_t)part * H * H + u;
      const float* dgp = dg + (size_t)part * H * BT;
#pragma unroll 4
      for (int g = 0; g < H; ++g) {
        float wv = __ldg(wp + (size_t)g * H);
        float4 d0 = *reinterpret_cast<const float4*>(dgp + g * BT);
        float4 d1 = *reinterpret_cast<const float4*>(dgp + g * BT + 4);
        acc[0] = fmaf(wv, d0.x, acc[0]); acc[1] = fmaf(wv, d0.y, acc[1]);
        acc[2] = fmaf(wv, d0.z, acc[2]); acc[3] = fmaf(wv, d0.w, acc[3]);
        acc[4] = fmaf(wv, d1.x, acc[4]); acc[5] = fmaf(wv, d1.y, acc[5]);
        acc[6] = fmaf(wv, d1.z, acc[6]); acc[7] = fmaf(wv, d1.w, acc[7]);
      }
#pragma unroll
      for (int w = 0; w < BT; ++w) atomicAdd(dhs + u * BT + w, acc[w]);
    }
    __syncthreads();
  }
}

static int rep_J(int n, int Hs) {
  int J = 1;
  for (int t = 0; t < n; ++t) {
    long long base = (long long)t * Hs;
    int span = (int)((base + Hs - 1) / n) - (int)(base / n) + 1;
    if (span > J) J = span;
  }
  return J;
}

static int launch_gru_fwd(GruFwdParams& P, cudaStream_t s) {
  int G = 3 * P.H;
  int threads = min(1024, ((G + 31) / 32) * 32);
  size_t smem = sizeof(float) * ((size_t)P.H * BT + (size_t)G * (BT + 1));
  if (smem > 200 * 1024) { mtadgat_set_error("gru_fwd: hidden size %d too large", P.H); return MTADGAT_ERR_UNSUPPORTED; }
  cudaFuncSetAttribute(gru_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  gru_fwd_kernel<<<cdiv(P.B, BT), threads, smem, s>>>(P);
  MG_COUNT_LAUNCH();
  return MTADGAT_OK;
}
static int launch_gru_bwd(GruBwdParams& P, cudaStream_t s) {
  int G = 3 * P.H;
  int threads = min(1024, ((G + 31) / 32) * 32);
  size_t smem = sizeof(float) * ((size_t)P.H * BT + (size_t)G * BT);
  if (smem > 200 * 1024) { mtadgat_set_error("gru_bwd: hidden size %d too large", P.H); return MTADGAT_ERR_UNSUPPORTED; }
  cudaFuncSetAttribute(gru_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  gru_bwd_kernel<<<cdiv(P.B, BT), threads, smem, s>>>(P);
  MG_COUNT_LAUNCH();
  return MTADGAT_OK;
}
static void launch_transpose(const float* src, float* dst, int R, int C, cudaStream_t s) {
  transpose_kernel<<<dim3(cdiv(C, 32), cdiv(R, 32)), dim3(32, 8), 0, s>>>(src, dst, R, C);
  MG_COUNT_LAUNCH();
}

}  // namespace

extern "C" int mtadgat_set_gru_impl(int impl) {
  MG_CHECK_ARG(impl == 0 || impl == 1, "set_gru_impl: 0 (fp32 SIMT) or 1 (tcgen05)");
  g_gru_impl = impl;
  return MTADGAT_OK;
}
extern "C" int mtadgat_get_gru_impl(void) { return g_gru_impl; }

// saved (floats): wt (H*3H) | gates (B*n*4H, only if save) ; gi scratch is separate
extern "C" long long mtadgat_gru_saved_floats(int B, int n, int H, int save) {
  return (long long)((size_t)3 * H * H + (save ? (size_t)B * n * 4 * H : 0));
}
extern "C" long long mtadgat_gru_fwd_scratch_floats(int B, int n, int H) { return (long long)((size_t)B * n * 3 * H); }
extern "C" long long mtadgat_gru_bwd_scratch_floats(int B, int n, int H) { return (long long)((size_t)B * n * 4 * H + 4); }

extern "C" int mtadgat_gru_fwd(const float* x0, const float* x1, const float* x2, int k0, int k1, int k2,
                               const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                               float* out, float* h_last, float* saved, float* scratch, int B, int n, int H, int save,
                               void* stream) {
  MG_CHECK_ARG(x0 && w_ih && w_hh && b_ih && b_hh && saved && scratch, "gru_fwd: null pointer");
  MG_CHECK_ARG(B > 0 && n > 0 && H > 0 && k0 > 0 && k1 >= 0 && k2 >= 0, "gru_fwd: bad shape");
  MG_CHECK_ARG((k1 == 0 || x1) && (k2 == 0 || x2), "gru_fwd: missing input slice");
  MG_CHECK_ARG(!save || out, "gru_fwd: training mode needs the per-step outputs");
  cudaStream_t s = (cudaStream_t)stream;
  const int I = k0 + k1 + k2, G = 3 * H;
  float* wt = saved; float* gates = save ? saved + (size_t)3 * H * H : nullptr;
  float* gi = scratch;
  launch_gemm_batched(1, B * n, G, I, Cat3A{x0, x1, x2, k0, k1, k2}, WT{w_ih, I},
                      StStrided{gi, 0, G, 1, b_ih, ACT_NONE, 0}, s);
  if (g_gru_impl == 1 && mtadgat_gru_tc_supported(H)) {
    mtadgat_gru_tc_fwd_launch(gi, nullptr, nullptr, nullptr, 0, 0, w_hh, b_hh, out, h_last, gates, B, n, H, s);
    MG_CHECK_LAUNCH("gru_fwd(tc)");
    return MTADGAT_OK;
  }
  launch_transpose(w_hh, wt, G, H, s);
  GruFwdParams P;
  P.gi = gi; P.S = nullptr; P.hsrc = nullptr; P.b_ih = nullptr; P.J = 0; P.Hs = 0; P.wt = wt; P.b_hh = b_hh;
  P.out = out; P.h_last = h_last; P.gates = gates; P.B = B; P.n = n; P.H = H;
  int rc = launch_gru_fwd(P, s);
  if (rc) return rc;
  MG_CHECK_LAUNCH("gru_fwd");
  return MTADGAT_OK;
}

extern "C" int mtadgat_gru_bwd(const float* x0, const float* x1, const float* x2, int k0, int k1, int k2,
                               const float* w_ih, const float* w_hh, const float* out, const float* saved,
                               const float* dout, const float* dh_last, float* scratch, float* dx0, float* dx1,
                               float* dx2, int acc0, int acc1, int acc2, float* dw_ih, float* dw_hh, float* db_ih,
                               float* db_hh, int B, int n, int H, void* stream) {
  MG_CHECK_ARG(x0 && w_ih && w_hh && out && saved && scratch && dw_ih && dw_hh && db_ih && db_hh, "gru_bwd: null pointer");
  MG_CHECK_ARG(dout || dh_last, "gru_bwd: need dout and/or dh_last");
  cudaStream_t s = (cudaStream_t)stream;
  const int I = k0 + k1 + k2, G = 3 * H;
  const float* gates = saved + (size_t)3 * H * H;
  float* dgi = scratch; float* dghn = scratch + (size_t)B * n * G;
  GruBwdParams P;
  P.gates = gates; P.out = out; P.w_hh = w_hh; P.dout = dout; P.dh_last = dh_last; P.dgi = dgi; P.dghn = dghn;
  P.B = B; P.n = n; P.H = H;
  if (g_gru_impl == 1 && mtadgat_gru_tc_bwd_supported(H)) {
    unsigned int* gmax = reinterpret_cast<unsigned int*>(scratch + (size_t)B * n * 4 * H);
    mtadgat_gru_tc_bwd_launch(gates, out, w_hh, dout, dh_last, gmax, dgi, dghn, B, n, H, s);
  } else {
    int rc = launch_gru_bwd(P, s);
    if (rc) return rc;
  }
  MG_CUDA(cudaMemsetAsync(dw_ih, 0, sizeof(float) * (size_t)G * I, s));
  MG_CUDA(cudaMemsetAsync(dw_hh, 0, sizeof(float) * (size_t)G * H, s));
  MG_CUDA(cudaMemsetAsync(db_ih, 0, sizeof(float) * (size_t)G, s));
  MG_CUDA(cudaMemsetAsync(db_hh, 0, sizeof(float) * (size_t)G, s));
  launch_gemm_splitk(G, I, B * n, DgiT{dgi, G}, Cat3B{x0, x1, x2, k0, k1, k2}, StAtomic2{dw_ih, I}, s);
  launch_gemm_splitk(G, H, B * n, DghT{dgi, dghn, H}, HprevB{out, n, H}, StAtomic2{dw_hh, H}, s);
  launch_colsum(B * n, G, Strided2<true>{dgi, 0, G, 1}, db_ih, s);
  launch_colsum(B * n, G, DghCols{dgi, dghn, H}, db_hh, s);
  if (dx0 || dx1 || dx2) {
    // dx = dgi W_ih : A(m,kk=g) = dgi[m,g] ; B(kk=g, n=i) = w_ih[g, i]
    launch_gemm_batched(1, B * n, I, G, Strided2<true>{dgi, 0, G, 1}, Strided2<true>{w_ih, 0, I, 1},
                        StCat3{dx0, dx1, dx2, k0, k1, k2, acc0, acc1, acc2}, s);
  }
  MG_CHECK_LAUNCH("gru_bwd");
  return MTADGAT_OK;
}

// ---- decoder GRU on the scrambled repeat of h_src (modules.py:279) -------------------------------------
extern "C" int mtadgat_rep_J(int n, int Hs) { return rep_J(n, Hs); }
// saved (floats): wt (R*3R) | S (n*J*3R) | gates (B*n*4R if save)
extern "C" long long mtadgat_gru_rep_saved_floats(int B, int n, int Hs, int R, int save) {
  int J = rep_J(n, Hs);
  return (long long)((size_t)3 * R * R + (size_t)n * J * 3 * R + (save ? (size_t)B * n * 4 * R : 0));
}
// scratch for bwd: dgi (B*n*3R) | dghn (B*n*R) | dS (n*J*3R)
extern "C" long long mtadgat_gru_rep_bwd_scratch_floats(int B, int n, int Hs, int R) {
  int J = rep_J(n, Hs);
  return (long long)((size_t)B * n * 4 * R + (size_t)n * J * 3 * R + 4);
}

extern "C" int mtadgat_gru_rep_fwd(const float* h_src, const float* w_ih, const float* w_hh, const float* b_ih,
                                   const float* b_hh, float* out, float* saved, int B, int n, int Hs, int R, int save,
                                   void* stream) {
  MG_CHECK_ARG(h_src && w_ih && w_hh && b_ih && b_hh && out && saved, "gru_rep_fwd: null pointer");
  MG_CHECK_ARG(B > 0 && n > 0 && Hs > 0 && R > 0, "gru_rep_fwd: bad shape");
  cudaStream_t s = (cudaStream_t)stream;
  const int G = 3 * R, J = rep_J(n, Hs);
  float* wt = saved; float* S = saved + (size_t)3 * R * R;
  float* gates = save ? S + (size_t)n * J * G : nullptr;
  rep_build_S_kernel<<<cdiv((long long)n * J * G, 256), 256, 0, s>>>(w_ih, n, Hs, G, J, S);
  MG_COUNT_LAUNCH();
  if (g_gru_impl == 1 && mtadgat_gru_tc_supported(R)) {
    mtadgat_gru_tc_fwd_launch(nullptr, S, h_src, b_ih, J, Hs, w_hh, b_hh, out, nullptr, gates, B, n, R, s);
    MG_CHECK_LAUNCH("gru_rep_fwd(tc)");
    return MTADGAT_OK;
  }
  launch_transpose(w_hh, wt, G, R, s);
  GruFwdParams P;
  P.gi = nullptr; P.S = S; P.hsrc = h_src; P.b_ih = b_ih; P.J = J; P.Hs = Hs; P.wt = wt; P.b_hh = b_hh;
  P.out = out; P.h_last = nullptr; P.gates = gates; P.B = B; P.n = n; P.H = R;
  int rc = launch_gru_fwd(P, s);
  if (rc) return rc;
  MG_CHECK_LAUNCH("gru_rep_fwd");
  return MTADGAT_OK;
}

extern "C" int mtadgat_gru_rep_bwd(const float* h_src, const float* w_ih, const float* w_hh, const float* out,
                                   const float* saved, const float* dout, float* scratch, float* dh_src,
                                   int dh_accumulate, float* dw_ih, float* dw_hh, float* db_ih, float* db_hh, int B,
                                   int n, int Hs, int R, void* stream) {
  MG_CHECK_ARG(h_src && w_ih && w_hh && out && saved && dout && scratch && dh_src && dw_ih && dw_hh && db_ih && db_hh,
               "gru_rep_bwd: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  const int G = 3 * R, J = rep_J(n, Hs);
  const float* S = saved + (size_t)3 * R * R;
  const float* gates = S + (size_t)n * J * G;
  float* dgi = scratch; float* dghn = dgi + (size_t)B * n * G; float* dS = dghn + (size_t)B * n * R;
  GruBwdParams P;
  P.gates = gates; P.out = out; P.w_hh = w_hh; P.dout = dout; P.dh_last = nullptr; P.dgi = dgi; P.dghn = dghn;
  P.B = B; P.n = n; P.H = R;
  if (g_gru_impl == 1 && mtadgat_gru_tc_bwd_supported(R)) {
    unsigned int* gmax = reinterpret_cast<unsigned int*>(dS + (size_t)n * J * G);
    mtadgat_gru_tc_bwd_launch(gates, out, w_hh, dout, nullptr, gmax, dgi, dghn, B, n, R, s);
  } else {
    int rc = launch_gru_bwd(P, s);
    if (rc) return rc;
  }
  MG_CUDA(cudaMemsetAsync(dw_hh, 0, sizeof(float) * (size_t)G * R, s));
  MG_CUDA(cudaMemsetAsync(db_ih, 0, sizeof(float) * (size_t)G, s));
  MG_CUDA(cudaMemsetAsync(db_hh, 0, sizeof(float) * (size_t)G, s));
  launch_gemm_splitk(G, R, B * n, DghT{dgi, dghn, R}, HprevB{out, n, R}, StAtomic2{dw_hh, R}, s);
  launch_colsum(B * n, G, Strided2<true>{dgi, 0, G, 1}, db_ih, s);
  launch_colsum(B * n, G, DghCols{dgi, dghn, R}, db_hh, s);
  rep_dS_kernel<<<cdiv((long long)n * J * G, 256), 256, 0, s>>>(dgi, h_src, B, n, Hs, G, J, dS);
  MG_COUNT_LAUNCH();
  rep_dw_kernel<<<cdiv((long long)G * Hs, 256), 256, 0, s>>>(dS, n, Hs, G, J, dw_ih);
  MG_COUNT_LAUNCH();
  rep_dh_kernel<<<cdiv((long long)B * Hs, 8), 256, 0, s>>>(dgi, S, B, n, Hs, G, J, dh_src, dh_accumulate);
  MG_COUNT_LAUNCH();
  MG_CHECK_LAUNCH("gru_rep_bwd");
  return MTADGAT_OK;
}
