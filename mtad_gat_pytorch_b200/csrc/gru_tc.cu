// Persistent tensor-core GRU recurrence for sm_100a (tcgen05 / TMEM).
//
// One CTA owns a tile of NB = 16 windows for all n steps.  W_hh (fp16, 141 KB at H = 150) is staged ONCE
// into shared memory in the canonical K-major no-swizzle operand layout and stays resident; per step one elected
// thread issues the  gh^T = W_hh h^T  product as tcgen05.mma tiles  (M = 128 hidden units of one gate,
// N = 16 windows, K = 16 per instruction)  accumulating in TMEM, commits to an mbarrier, and 256 epilogue threads
// (thread = hidden unit, TMEM lane = unit) pull their gate pre-activations with tcgen05.ld, apply the gate
// math in fp32 registers (the fp32 master copy of h never leaves registers), and write h_t back as the fp16 B
// operand of the next step.  Gate order r,z,n as torch.nn.GRU (reference modules.py:233,253).
//
// Row layout of the A operand: rows [g*Hp8, g*Hp8 + H) hold gate g (Hp8 = H rounded up to 8).  The M=128 tile
// of unit block `blk` of gate g starts at row g*Hp8 + 128*blk; for the last (partial) block the instruction
// reads on into the next gate's rows -- those accumulator lanes are never read.
#include "tc.cuh"
#include "../../include/mtadgat.h"

namespace {

constexpr int NB = 16;          // windows per CTA  (MMA N)
constexpr int EPI_THREADS = 256;
constexpr int MMA_WARPS = 2;    // warp 8 issues unit block 0, warp 9 unit block 1 (issue is per-thread serial)
constexpr int TC_THREADS = EPI_THREADS + 32 * MMA_WARPS;

struct GruTcDims {
  int H, Hp8, Kp, KC, Mtot, NBLK, lboA, lboB;
  size_t a_bytes, b_bytes, smem;
};
static GruTcDims tc_dims(int H) {
  GruTcDims d;
  d.H = H; d.Hp8 = (H + 7) & ~7; d.Kp = (H + 15) & ~15; d.KC = d.Kp / 8; d.Mtot = 3 * d.Hp8;
  d.NBLK = (H + 127) / 128;
  d.lboA = d.Mtot * 16; d.lboB = NB * 16 + 16;
  d.a_bytes = (size_t)d.KC * d.lboA;
  // the partial last block reads up to 128 rows past its start: keep that inside the allocation
  size_t over = (size_t)(2 * d.Hp8 + 128 * (d.NBLK - 1) + 128 - d.Mtot) * 16;
  d.b_bytes = (size_t)d.KC * d.lboB;
  size_t tail = d.b_bytes + 64;
  if (tail < over + 64) tail = over + 64;
  d.smem = d.a_bytes + tail + 128;
  return d;
}

struct GruTcParams {
  const float* gi;                                    // (B,n,3H) incl. b_ih, or nullptr in rep mode
  const float* S; const float* hsrc; const float* b_ih; int J, Hs;
  const float* w_hh; const float* b_hh;
  float* out; float* h_last; float* gates;
  int B, n, H;
};

__device__ __forceinline__ float sigm(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) { return 2.f * __fdividef(1.f, 1.f + __expf(-2.f * x)) - 1.f; }

__global__ void __launch_bounds__(TC_THREADS, 1) gru_tc_fwd_kernel(GruTcParams P) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int H = P.H, G = 3 * H, n = P.n;
  const int Hp8 = (H + 7) & ~7, Kp = (H + 15) & ~15, KC = Kp / 8, Mtot = 3 * Hp8, NBLK = (H + 127) / 128;
  const int lboA = Mtot * 16, lboB = NB * 16 + 16;
  uint8_t* sA = smem_raw;
  uint8_t* sB = sA + (size_t)KC * lboA;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (((size_t)KC * lboB + 15) & ~(size_t)15));
  uint64_t* acc_bar = bars;        // MMA -> epilogue
  uint64_t* h_bar = bars + 1;      // epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b0 = blockIdx.x * NB;

  // ---- one-time staging: W_hh -> fp16 canonical layout; h_0 = 0 ----
  for (int idx = tid; idx < Mtot * Kp; idx += TC_THREADS) {
    int row = idx / Kp, k = idx - row * Kp;
    int g = row / Hp8, u = row - g * Hp8;
    float v = (u < H && k < H) ? __ldg(P.w_hh + ((size_t)g * H + u) * H + k) : 0.f;
    *reinterpret_cast<__half*>(sA + (size_t)(k >> 3) * lboA + (size_t)row * 16 + (k & 7) * 2) = __float2half_rn(v);
  }
  for (int idx = tid; idx < (KC * lboB) / 4; idx += TC_THREADS) reinterpret_cast<uint32_t*>(sB)[idx] = 0u;
  if (tid == 0) {
    tc::mbar_init(acc_bar, NBLK);          // one commit per issuing warp
    tc::mbar_init(h_bar, EPI_THREADS);
    tc::fence_mbar_init();
  }
  if (warp == 8) tc::tmem_alloc(tmem_slot, 128);
  tc::fence_proxy_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tbase = *tmem_slot;

  if (warp >= 8) {
    // ================= MMA issuers: warp 8 -> unit block 0, warp 9 -> unit block 1 =================
    const int blk = warp - 8;
    if (lane == 0 && blk < NBLK) {
      const uint32_t idesc = tc::make_idesc_f16(128, NB, 0);
      const uint32_t aBase = tc::smem_u32(sA), bBase = tc::smem_u32(sB);
      const uint64_t bd0 = tc::make_smem_desc(bBase, lboB, 128);
      const uint32_t blo0 = (uint32_t)bd0, bhi = (uint32_t)(bd0 >> 32);
      const uint32_t ainc = (uint32_t)(2 * lboA) >> 4, binc = (uint32_t)(2 * lboB) >> 4;
      uint32_t alo0[3], ahi = 0, dt[3];
#pragma unroll
      for (int g = 0; g < 3; ++g) {
        uint64_t ad = tc::make_smem_desc(aBase + (uint32_t)(g * Hp8 + blk * 128) * 16, lboA, 128);
        alo0[g] = (uint32_t)ad; ahi = (uint32_t)(ad >> 32);
        dt[g] = tbase + (uint32_t)((blk * 3 + g) * NB);
      }
      const int nkc = Kp / 16;
      for (int t = 0; t < n; ++t) {
        if (t > 0) tc::mbar_wait(h_bar, (t - 1) & 1);
        tc::tc_fence_after();
#pragma unroll
        for (int g = 0; g < 3; ++g) {
          uint32_t alo = alo0[g], blo = blo0;
          tc::mma_f16_ss_lohi(dt[g], alo, ahi, blo, bhi, idesc, 0u);
          for (int kc = 1; kc < nkc; ++kc) {
            alo += ainc; blo += binc;
            tc::mma_f16_ss_lohi(dt[g], alo, ahi, blo, bhi, idesc, 1u);
          }
        }
        tc::mma_commit(acc_bar);
      }
    }
  } else {
    // ================= epilogue: thread = hidden unit =================
    const int blk = warp >> 2, q = warp & 3;
    const int u = blk * 128 + q * 32 + lane;
    const bool valid = u < H;
    const float bhr = valid ? __ldg(P.b_hh + u) : 0.f, bhz = valid ? __ldg(P.b_hh + H + u) : 0.f,
                bhn = valid ? __ldg(P.b_hh + 2 * H + u) : 0.f;
    float bir = 0.f, biz = 0.f, bin = 0.f;
    if (!P.gi && valid) { bir = __ldg(P.b_ih + u); biz = __ldg(P.b_ih + H + u); bin = __ldg(P.b_ih + 2 * H + u); }
    float h[NB];
#pragma unroll
    for (int w = 0; w < NB; ++w) h[w] = 0.f;
    uint8_t* hdst = sB + (size_t)(u >> 3) * lboB + (u & 7) * 2;
    const uint32_t tlane = tbase + ((uint32_t)(q * 32) << 16) + (uint32_t)(blk * 3 * NB);

    for (int t = 0; t < n; ++t) {
      // input-side pre-activations for this step (independent of the recurrence: issued before the wait)
      float gr[NB], gz[NB], gn[NB];
      if (P.gi) {
#pragma unroll
        for (int w = 0; w < NB; ++w) {
          int b = b0 + w;
          if (valid && b < P.B) {
            const float* gp = P.gi + ((size_t)b * n + t) * G + u;
            gr[w] = __ldg(gp); gz[w] = __ldg(gp + H); gn[w] = __ldg(gp + 2 * H);
          } else { gr[w] = 0.f; gz[w] = 0.f; gn[w] = 0.f; }
        }
      } else {
        const int m0 = (int)(((long long)t * P.Hs) / n);
#pragma unroll
        for (int w = 0; w < NB; ++w) { gr[w] = bir; gz[w] = biz; gn[w] = bin; }
        for (int j = 0; j < P.J; ++j) {
          int m = m0 + j;
          if (m >= P.Hs) break;
          float sr = 0.f, sz = 0.f, sn = 0.f;
          if (valid) {
            const float* sp = P.S + ((size_t)t * P.J + j) * G + u;
            sr = __ldg(sp); sz = __ldg(sp + H); sn = __ldg(sp + 2 * H);
          }
#pragma unroll
          for (int w = 0; w < NB; ++w) {
            int b = b0 + w;
            float hv = (b < P.B) ? __ldg(P.hsrc + (size_t)b * P.Hs + m) : 0.f;
            gr[w] = fmaf(hv, sr, gr[w]); gz[w] = fmaf(hv, sz, gz[w]); gn[w] = fmaf(hv, sn, gn[w]);
          }
        }
      }
      tc::mbar_wait(acc_bar, t & 1);
      tc::tc_fence_after();
      float ar[NB], az[NB], an[NB];
      if (blk < NBLK) {
        tc::tmem_ld16(tlane, ar);
        tc::tmem_ld16(tlane + NB, az);
        tc::tmem_ld16(tlane + 2 * NB, an);
        tc::tmem_ld_wait();
      }
      if (valid) {
#pragma unroll
        for (int w = 0; w < NB; ++w) {
          float r = sigm(gr[w] + ar[w] + bhr);
          float z = sigm(gz[w] + az[w] + bhz);
          float hn = an[w] + bhn;
          float nn = tanh_fast(gn[w] + r * hn);
          float hnew = (1.f - z) * nn + z * h[w];
          h[w] = hnew;
          *reinterpret_cast<__half*>(hdst + w * 16) = __float2half_rn(hnew);
          int b = b0 + w;
          if (b < P.B) {
            size_t o = (size_t)b * n + t;
            if (P.out) P.out[o * H + u] = hnew;
            if (P.gates) {
              float* gp = P.gates + o * 4 * H + u;
              gp[0] = r; gp[H] = z; gp[2 * H] = nn; gp[3 * H] = hn;
            }
          }
        }
      }
      tc::fence_proxy_async_smem();
      tc::tc_fence_before();
      tc::mbar_arrive(h_bar);
    }
    if (valid && P.h_last) {
#pragma unroll
      for (int w = 0; w < NB; ++w)
        if (b0 + w < P.B) P.h_last[(size_t)(b0 + w) * H + u] = h[w];
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 8) tc::tmem_dealloc(tbase, 128);
}


// =====================================================================================================
// BPTT on the tensor cores:  dh_{t-1} = dh_t (.) z_t + W_hh^T dgh_t.
// A operand = W_hh^T (rows = hidden units, K = the 3H gate rows) resident in shared memory as fp16; B operand =
// dgh_t (16 windows x 3H) rewritten every step by the epilogue threads, scaled by a power of two so that the
// gradients sit in fp16's normal range (scale derived on the device from max|dout|,|dh_last|); accumulators in
// TMEM; dgi / dgh_n for the weight-gradient GEMMs leave in fp32.
// =====================================================================================================
struct GruTcBwdParams {
  const float* gates; const float* out; const float* w_hh;
  const float* dout; const float* dh_last;
  const unsigned int* gmax_bits;      // max |dout|,|dh_last| as float bits (device)
  float* dgi; float* dghn;
  int B, n, H;
};

struct GruTcBwdDims { int Hp8, G, Kp, KC, NBLK, lboA, lboB; size_t smem; };
static GruTcBwdDims tc_bwd_dims(int H) {
  GruTcBwdDims d;
  d.Hp8 = (H + 7) & ~7; d.G = 3 * H; d.Kp = (d.G + 15) & ~15; d.KC = d.Kp / 8; d.NBLK = (H + 127) / 128;
  d.lboA = d.Hp8 * 16; d.lboB = NB * 16 + 16;
  size_t a_bytes = (size_t)d.KC * d.lboA;
  size_t over = (size_t)(128 * d.NBLK - d.Hp8 > 0 ? 128 * d.NBLK - d.Hp8 : 0) * 16;
  size_t b_bytes = (size_t)d.KC * d.lboB;
  size_t tail = b_bytes + 64 > over + 64 ? b_bytes + 64 : over + 64;
  d.smem = a_bytes + tail + 128;
  return d;
}

__global__ void absmax_kernel(const float* __restrict__ a, long long na, const float* __restrict__ b, long long nb,
                              unsigned int* __restrict__ out_bits) {
  float m = 0.f;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
  if (a) for (long long j = i; j < na; j += stride) m = fmaxf(m, fabsf(a[j]));
  if (b) for (long long j = i; j < nb; j += stride) m = fmaxf(m, fabsf(b[j]));
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out_bits, __float_as_uint(m));
}

__device__ __forceinline__ __half to_half_sat(float v) { return __float2half_rn(fminf(fmaxf(v, -60000.f), 60000.f)); }

__global__ void __launch_bounds__(TC_THREADS, 1) gru_tc_bwd_kernel(GruTcBwdParams P) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int H = P.H, G = 3 * H, n = P.n;
  const int Hp8 = (H + 7) & ~7, Kp = (G + 15) & ~15, KC = Kp / 8, NBLK = (H + 127) / 128;
  const int lboA = Hp8 * 16, lboB = NB * 16 + 16;
  uint8_t* sA = smem_raw;
  uint8_t* sB = sA + (size_t)KC * lboA;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (((size_t)KC * lboB + 15) & ~(size_t)15));
  uint64_t* acc_bar = bars;
  uint64_t* h_bar = bars + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b0 = blockIdx.x * NB;

  // A[u][g] = W_hh[g][u]  (fp16), zero padded
  for (int idx = tid; idx < Hp8 * Kp; idx += TC_THREADS) {
    int g = idx / Hp8, u = idx - g * Hp8;           // u fastest: coalesced reads of W_hh rows
    float v = (u < H && g < G) ? __ldg(P.w_hh + (size_t)g * H + u) : 0.f;
    *reinterpret_cast<__half*>(sA + (size_t)(g >> 3) * lboA + (size_t)u * 16 + (g & 7) * 2) = __float2half_rn(v);
  }
  for (int idx = tid; idx < (KC * lboB) / 4; idx += TC_THREADS) reinterpret_cast<uint32_t*>(sB)[idx] = 0u;
  if (tid == 0) {
    tc::mbar_init(acc_bar, NBLK);
    tc::mbar_init(h_bar, EPI_THREADS);
    tc::fence_mbar_init();
  }
  if (warp == 8) tc::tmem_alloc(tmem_slot, 32);
  tc::fence_proxy_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tbase = *tmem_slot;

  if (warp >= 8) {
    const int blk = warp - 8;
    if (lane == 0 && blk < NBLK) {
      const uint32_t idesc = tc::make_idesc_f16(128, NB, 0);
      const uint64_t bd0 = tc::make_smem_desc(tc::smem_u32(sB), lboB, 128);
      const uint64_t ad0 = tc::make_smem_desc(tc::smem_u32(sA) + (uint32_t)(blk * 128) * 16, lboA, 128);
      const uint32_t blo0 = (uint32_t)bd0, bhi = (uint32_t)(bd0 >> 32), alo0 = (uint32_t)ad0, ahi = (uint32_t)(ad0 >> 32);
      const uint32_t ainc = (uint32_t)(2 * lboA) >> 4, binc = (uint32_t)(2 * lboB) >> 4;
      const uint32_t dt = tbase + (uint32_t)(blk * NB);
      const int nkc = Kp / 16;
      for (int it = 0; it < n - 1; ++it) {
        tc::mbar_wait(h_bar, it & 1);
        tc::tc_fence_after();
        uint32_t alo = alo0, blo = blo0;
        tc::mma_f16_ss_lohi(dt, alo, ahi, blo, bhi, idesc, 0u);
        for (int kc = 1; kc < nkc; ++kc) {
          alo += ainc; blo += binc;
          tc::mma_f16_ss_lohi(dt, alo, ahi, blo, bhi, idesc, 1u);
        }
        tc::mma_commit(acc_bar);
      }
    }
  } else {
    const int blk = warp >> 2, q = warp & 3;
    const int u = blk * 128 + q * 32 + lane;
    const bool valid = u < H;
    const float gmax = __uint_as_float(*P.gmax_bits);
    // power-of-two scale putting max|grad| at ~2^6 (fp16 overflows at 65504; values are clamped on conversion)
    const float scale = gmax > 0.f ? exp2f(floorf(log2f(64.f / gmax))) : 1.f;
    const float inv_scale = 1.f / scale;
    float dhz[NB];
#pragma unroll
    for (int w = 0; w < NB; ++w) dhz[w] = 0.f;
    uint8_t* d0 = sB + (size_t)(u >> 3) * lboB + (u & 7) * 2;
    uint8_t* d1 = sB + (size_t)((H + u) >> 3) * lboB + ((H + u) & 7) * 2;
    uint8_t* d2 = sB + (size_t)((2 * H + u) >> 3) * lboB + ((2 * H + u) & 7) * 2;
    const uint32_t tlane = tbase + ((uint32_t)(q * 32) << 16) + (uint32_t)(blk * NB);

    for (int t = n - 1; t >= 0; --t) {
      const int it = n - 1 - t;
      float dh[NB];
      // (1) contributions that do not depend on the recurrence
#pragma unroll
      for (int w = 0; w < NB; ++w) {
        int b = b0 + w;
        float v = 0.f;
        if (valid && b < P.B) {
          if (P.dout) v = __ldg(P.dout + ((size_t)b * n + t) * H + u);
          if (it == 0 && P.dh_last) v += __ldg(P.dh_last + (size_t)b * H + u);
        }
        dh[w] = v;
      }
      // (2) recurrent part from the previous iteration's MMA
      if (it > 0) {
        tc::mbar_wait(acc_bar, (it - 1) & 1);
        tc::tc_fence_after();
        float acc[NB];
        if (blk < NBLK) {
          tc::tmem_ld16(tlane, acc);
          tc::tmem_ld_wait();
#pragma unroll
          for (int w = 0; w < NB; ++w) dh[w] += dhz[w] + acc[w] * inv_scale;
        }
      }
      // (3) gate gradients
      if (valid) {
#pragma unroll
        for (int w = 0; w < NB; ++w) {
          int b = b0 + w;
          float dpr = 0.f, dpz = 0.f, dgn = 0.f;
          if (b < P.B) {
            size_t o = (size_t)b * n + t;
            const float* gp = P.gates + o * 4 * H + u;
            float r = __ldg(gp), z = __ldg(gp + H), nn = __ldg(gp + 2 * H), hn = __ldg(gp + 3 * H);
            float hp = t > 0 ? __ldg(P.out + (o - 1) * H + u) : 0.f;
            float d = dh[w];
            float dn = d * (1.f - z);
            float dz = d * (hp - nn);
            float dpn = dn * (1.f - nn * nn);
            dpz = dz * z * (1.f - z);
            dpr = dpn * hn * r * (1.f - r);
            dgn = dpn * r;
            dhz[w] = d * z;
            float* qd = P.dgi + o * G + u;
            qd[0] = dpr; qd[H] = dpz; qd[2 * H] = dpn;
            P.dghn[o * H + u] = dgn;
          } else {
            dhz[w] = 0.f;
          }
          if (t > 0) {
            *reinterpret_cast<__half*>(d0 + w * 16) = to_half_sat(dpr * scale);
            *reinterpret_cast<__half*>(d1 + w * 16) = to_half_sat(dpz * scale);
            *reinterpret_cast<__half*>(d2 + w * 16) = to_half_sat(dgn * scale);
          }
        }
      }
      if (t > 0) {
        tc::fence_proxy_async_smem();
        tc::tc_fence_before();
        tc::mbar_arrive(h_bar);
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 8) tc::tmem_dealloc(tbase, 32);
}

// ---- probe: D(128 x N) = A[row0 .. row0+128) (Mtot x K) * B (N x K)^T through the same operand layout ----
__global__ void __launch_bounds__(128, 1) tc_probe_kernel(const float* __restrict__ A, const float* __restrict__ Bm,
                                                         float* __restrict__ D, int Mtot, int row0, int K, int N) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int KC = K / 8, lboA = Mtot * 16, lboB = N * 16 + 16;
  uint8_t* sA = smem_raw;
  uint8_t* sB = sA + (size_t)KC * lboA;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + (((size_t)KC * lboB + 15) & ~(size_t)15));
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int idx = tid; idx < Mtot * K; idx += 128) {
    int r = idx / K, k = idx - r * K;
    *reinterpret_cast<__half*>(sA + (size_t)(k >> 3) * lboA + (size_t)r * 16 + (k & 7) * 2) = __float2half_rn(A[idx]);
  }
  for (int idx = tid; idx < N * K; idx += 128) {
    int r = idx / K, k = idx - r * K;
    *reinterpret_cast<__half*>(sB + (size_t)(k >> 3) * lboB + (size_t)r * 16 + (k & 7) * 2) = __float2half_rn(Bm[idx]);
  }
  uint32_t ncols = 32;
  while ((int)ncols < N) ncols <<= 1;
  if (tid == 0) { tc::mbar_init(bar, 1); tc::fence_mbar_init(); }
  if (warp == 0) tc::tmem_alloc(slot, ncols);
  tc::fence_proxy_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tbase = *slot;
  if (tid == 0) {
    const uint32_t idesc = tc::make_idesc_f16(128, N, 0);
    for (int kc = 0; kc < K / 16; ++kc) {
      uint64_t ad = tc::make_smem_desc(tc::smem_u32(sA) + (uint32_t)row0 * 16 + (uint32_t)(2 * kc) * lboA, lboA, 128);
      uint64_t bd = tc::make_smem_desc(tc::smem_u32(sB) + (uint32_t)(2 * kc) * lboB, lboB, 128);
      tc::mma_f16_ss(tbase, ad, bd, idesc, kc > 0 ? 1u : 0u);
    }
    tc::mma_commit(bar);
  }
  tc::mbar_wait(bar, 0);
  tc::tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    tc::tmem_ld16(tbase + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
    tc::tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) D[(size_t)(warp * 32 + lane) * N + c0 + i] = v[i];
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tbase, ncols);
}

}  // namespace

// Called from gru.cu ------------------------------------------------------------------------------------------
int mtadgat_gru_tc_supported(int H) {
  if (H < 8 || H > 256) return 0;
  return tc_dims(H).smem <= 220 * 1024;
}

int mtadgat_gru_tc_fwd_launch(const float* gi, const float* S, const float* hsrc, const float* b_ih, int J, int Hs,
                              const float* w_hh, const float* b_hh, float* out, float* h_last, float* gates, int B,
                              int n, int H, cudaStream_t s) {
  GruTcDims d = tc_dims(H);
  GruTcParams P;
  P.gi = gi; P.S = S; P.hsrc = hsrc; P.b_ih = b_ih; P.J = J; P.Hs = Hs; P.w_hh = w_hh; P.b_hh = b_hh;
  P.out = out; P.h_last = h_last; P.gates = gates; P.B = B; P.n = n; P.H = H;
  cudaFuncSetAttribute(gru_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d.smem);
  gru_tc_fwd_kernel<<<cdiv(B, NB), TC_THREADS, d.smem, s>>>(P);
  MG_COUNT_LAUNCH();
  return MTADGAT_OK;
}

int mtadgat_gru_tc_bwd_supported(int H) {
  if (H < 8 || H > 256) return 0;
  return tc_bwd_dims(H).smem <= 220 * 1024;
}

// gmax_bits: one device uint (scratch) ; everything else as in the SIMT BPTT kernel
int mtadgat_gru_tc_bwd_launch(const float* gates, const float* out, const float* w_hh, const float* dout,
                              const float* dh_last, unsigned int* gmax_bits, float* dgi, float* dghn, int B, int n,
                              int H, cudaStream_t s) {
  GruTcBwdDims d = tc_bwd_dims(H);
  cudaMemsetAsync(gmax_bits, 0, sizeof(unsigned int), s);
  absmax_kernel<<<148, 256, 0, s>>>(dout, dout ? (long long)B * n * H : 0, dh_last, dh_last ? (long long)B * H : 0, gmax_bits);
  MG_COUNT_LAUNCH();
  GruTcBwdParams P;
  P.gates = gates; P.out = out; P.w_hh = w_hh; P.dout = dout; P.dh_last = dh_last; P.gmax_bits = gmax_bits;
  P.dgi = dgi; P.dghn = dghn; P.B = B; P.n = n; P.H = H;
  cudaFuncSetAttribute(gru_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d.smem);
  gru_tc_bwd_kernel<<<cdiv(B, NB), TC_THREADS, d.smem, s>>>(P);
  MG_COUNT_LAUNCH();
  return MTADGAT_OK;
}

extern "C" int mtadgat_tc_probe(const float* A, const float* Bm, float* D, int Mtot, int row0, int K, int N,
                                void* stream) {
  MG_CHECK_ARG(A && Bm && D, "tc_probe: null pointer");
  MG_CHECK_ARG(K % 16 == 0 && N % 16 == 0 && N >= 16 && N <= 256 && Mtot % 8 == 0 && row0 % 8 == 0 && row0 >= 0,
               "tc_probe: need K%%16==0, N%%16==0, Mtot%%8==0, row0%%8==0");
  const int KC = K / 8;
  size_t a_bytes = (size_t)KC * Mtot * 16;
  size_t over = (row0 + 128 > Mtot) ? (size_t)(row0 + 128 - Mtot) * 16 : 0;
  size_t b_bytes = (size_t)KC * (N * 16 + 16);
  size_t tail = b_bytes + 64 > over + 64 ? b_bytes + 64 : over + 64;
  size_t smem = a_bytes + tail + 128;
  MG_CHECK_ARG(smem <= 220 * 1024, "tc_probe: operands do not fit in shared memory");
  cudaFuncSetAttribute(tc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  tc_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, Bm, D, Mtot, row0, K, N);
  MG_COUNT_LAUNCH();
  MG_CHECK_LAUNCH("tc_probe");
  return MTADGAT_OK;
}
