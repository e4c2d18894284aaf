// Fused ELBO path for the mixture-of-experts models (MMVAE, MMVAE+): HBM-bound elementwise +
// reduction kernels, 128-bit streaming loads, warp-shuffle reductions, no tensor cores.
//
//   mv_moe_lpx_fwd : lpx[c,k,b]  = sum_r rescale_r * sum_d log p_r(x_r[b,d] | recon_r[c,k,b,d])
//   mv_moe_lw_fwd  : latent log-probs, MoE log-mean-exp, lw, IWAE/DReG weights, loss, unit latent grads
//   mv_moe_lpx_bwd : g_recon = g_loss * coef[c,k,b] * rescale * dlogp/drecon
//
// Work decomposition of the two streaming kernels: one warp owns one (cond modality c, sample b,
// D-chunk) and keeps the target row x[b, chunk] in registers while it walks the K importance samples,
// so x is read once per K rows and the only HBM stream is `recon` (read once) / `g_recon` (written once).
#include "common.cuh"

namespace mv {

// ---- per-element terms (torch.distributions.{Normal,Laplace,Bernoulli}.log_prob) ----------------
template <int DIST>
__device__ __forceinline__ float lp_term(float x, float r) {
  if (DIST == MV_DIST_NORMAL) {
    float t = x - r;
    return t * t;
  } else if (DIST == MV_DIST_LAPLACE) {
    return fabsf(x - r);
  } else {  // Bernoulli(logits=r).log_prob(x) = x*r - softplus(r)
    return x * r - (fmaxf(r, 0.f) + log1pf(expf(-fabsf(r))));
  }
}
template <int DIST>
__device__ __forceinline__ float lp_grad(float x, float r, float inv_s) {
  if (DIST == MV_DIST_NORMAL) {
    return (x - r) * inv_s * inv_s;
  } else if (DIST == MV_DIST_LAPLACE) {
    float t = x - r;
    return t > 0.f ? inv_s : (t < 0.f ? -inv_s : 0.f);
  } else {
    return x - 1.f / (1.f + expf(-r));
  }
}

constexpr int kChunkElems = 2560;  // elements of D one warp keeps in registers (80 floats / lane)

template <typename T, int DIST>
__global__ void __launch_bounds__(128) lpx_fwd_kernel(const T* __restrict__ recon, const float* __restrict__ x,
                                                      float* __restrict__ lpx, int C, int K, int B, int64_t D,
                                                      float mul, float add_per_elem, float rescale,
                                                      const uint8_t* __restrict__ mask, int accumulate, int nchunks,
                                                      int ksplit) {
  constexpr int VE = Vec<T>::N;
  constexpr int NV = kChunkElems / (32 * VE);
  const int lane = threadIdx.x & 31;
  int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (warp >= int64_t(C) * B * nchunks * ksplit) return;
  // the K importance samples of a (c, b, chunk) are split over `ksplit` warps when there are too few rows to fill the GPU
  const int ks = int(warp % ksplit);
  warp /= ksplit;
  const int kper = (K + ksplit - 1) / ksplit;
  const int k_begin = ks * kper, k_end = min(K, k_begin + kper);
  const int chunk = int(warp % nchunks);
  const int64_t cb = warp / nchunks;
  const int b = int(cb % B), c = int(cb / B);
  const bool live = mask == nullptr || mask[b] != 0;
  const int64_t nvec = D / VE;
  const int64_t v0 = int64_t(chunk) * NV * 32 + lane;
  float xr[NV][VE];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int64_t v = v0 + j * 32;
    if (live && v < nvec) {
      const float4* xp = reinterpret_cast<const float4*>(x + int64_t(b) * D + v * VE);
#pragma unroll
      for (int q = 0; q < VE / 4; ++q) {
        float4 t = __ldg(xp + q);
        xr[j][4 * q + 0] = t.x; xr[j][4 * q + 1] = t.y; xr[j][4 * q + 2] = t.z; xr[j][4 * q + 3] = t.w;
      }
    }
  }
  int64_t n_here = nvec - int64_t(chunk) * NV * 32;
  n_here = (n_here > NV * 32 ? NV * 32 : n_here) * VE;
  for (int k = k_begin; k < k_end; ++k) {
    const int64_t row = (int64_t(c) * K + k) * B + b;
    float acc = 0.f;
    if (live) {
      uint4 rv[NV];
      const T* rp = recon + row * D;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int64_t v = v0 + j * 32;
        if (v < nvec) rv[j] = ld_stream(rp + v * VE);
      }
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int64_t v = v0 + j * 32;
        if (v < nvec) {
          float r[VE];
          Vec<T>::unpack(rv[j], r);
#pragma unroll
          for (int e = 0; e < VE; ++e) acc += lp_term<DIST>(xr[j][e], r[e]);
        }
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      const float val = live ? rescale * (acc * mul + add_per_elem * float(n_here)) : 0.f;
      if (nchunks == 1) {
        lpx[row] = accumulate ? lpx[row] + val : val;
      } else {
        atomicAdd(lpx + row, val);
      }
    }
  }
}

// High-occupancy forward variant: one block per (c, b) stages the fp32 target row in shared memory once; its four
// warps walk the K importance samples, streaming `recon` with 128-bit loads (4 in flight per lane).  ~40 registers,
// so 12+ blocks are resident per SM and the HBM stream stays saturated (the register-cached variant above holds
// 80 target floats per lane and tops out at 3 blocks per SM).
// Per-modality arguments of the batched (one launch for all reconstructed modalities) variants: blockIdx.y = modality.
constexpr int kMaxBatchMods = 8;
struct LpxBatch {
  const void* recon[kMaxBatchMods];
  const float* x[kMaxBatchMods];
  void* g[kMaxBatchMods];
  const uint8_t* mask[kMaxBatchMods];
  float mul[kMaxBatchMods], add[kMaxBatchMods], inv_s[kMaxBatchMods], rescale[kMaxBatchMods];
};

// accumulate: 0 store, 1 add (single writer), 2 atomic add (several modalities write the same row concurrently)
template <typename T, int DIST>
__device__ __forceinline__ void lpx_fwd_smem_body(const T* __restrict__ recon, const float* __restrict__ x,
                                                  float* __restrict__ lpx, int C, int K, int B, int64_t D, float mul,
                                                  float add_per_elem, float rescale, const uint8_t* __restrict__ mask,
                                                  int accumulate) {
  constexpr int VE = Vec<T>::N;
  extern __shared__ float xs[];
  const int b = blockIdx.x % B, c = blockIdx.x / B;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool live = mask == nullptr || mask[b] != 0;
  if (live) {
    const float4* xp = reinterpret_cast<const float4*>(x + int64_t(b) * D);
    for (int i = threadIdx.x; i < int(D / 4); i += 128) reinterpret_cast<float4*>(xs)[i] = __ldg(xp + i);
  }
  __syncthreads();
  const int nvec = int(D / VE);
  for (int k = warp; k < K; k += 4) {
    const int64_t row = (int64_t(c) * K + k) * B + b;
    float acc = 0.f;
    if (live) {
      const T* rp = recon + row * D;
      int v = lane;
      for (; v + 96 < nvec; v += 128) {
        uint4 rv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) rv[j] = ld_stream(rp + int64_t(v + 32 * j) * VE);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float r[VE];
          Vec<T>::unpack(rv[j], r);
          const float* xv = xs + (v + 32 * j) * VE;
#pragma unroll
          for (int e = 0; e < VE; ++e) acc += lp_term<DIST>(xv[e], r[e]);
        }
      }
      for (; v < nvec; v += 32) {
        float r[VE];
        Vec<T>::unpack(ld_stream(rp + int64_t(v) * VE), r);
        const float* xv = xs + v * VE;
#pragma unroll
        for (int e = 0; e < VE; ++e) acc += lp_term<DIST>(xv[e], r[e]);
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      const float val = live ? rescale * (acc * mul + add_per_elem * float(D)) : 0.f;
      if (accumulate == 2) atomicAdd(lpx + row, val);
      else lpx[row] = accumulate ? lpx[row] + val : val;
    }
  }
}
template <typename T, int DIST>
__global__ void __launch_bounds__(128) lpx_fwd_smem_kernel(const T* __restrict__ recon, const float* __restrict__ x,
                                                           float* __restrict__ lpx, int C, int K, int B, int64_t D, float mul,
                                                           float add_per_elem, float rescale, const uint8_t* __restrict__ mask,
                                                           int accumulate) {
  lpx_fwd_smem_body<T, DIST>(recon, x, lpx, C, K, B, D, mul, add_per_elem, rescale, mask, accumulate);
}
template <typename T, int DIST>
__global__ void __launch_bounds__(128) lpx_fwd_smem_multi_kernel(const __grid_constant__ LpxBatch bt, float* __restrict__ lpx, int C,
                                                                 int K, int B, int64_t D) {
  const int m = blockIdx.y;
  lpx_fwd_smem_body<T, DIST>(static_cast<const T*>(bt.recon[m]), bt.x[m], lpx, C, K, B, D, bt.mul[m], bt.add[m], bt.rescale[m],
                             bt.mask[m], 2);
}

// scalar fallback for rows whose byte length is not a multiple of 16 (e.g. D = 10)
template <typename T, int DIST>
__global__ void __launch_bounds__(128) lpx_fwd_scalar_kernel(const T* __restrict__ recon, const float* __restrict__ x,
                                                             float* __restrict__ lpx, int C, int K, int B, int64_t D,
                                                             float mul, float add_per_elem, float rescale,
                                                             const uint8_t* __restrict__ mask, int accumulate) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (row >= int64_t(C) * K * B) return;
  const int b = int(row % B);
  const bool live = mask == nullptr || mask[b] != 0;
  float acc = 0.f;
  if (live)
    for (int64_t d = lane; d < D; d += 32)
      acc += lp_term<DIST>(x[int64_t(b) * D + d], Vec<T>::load1(recon + row * D + d));
  acc = warp_sum(acc);
  if (lane == 0) {
    const float val = live ? rescale * (acc * mul + add_per_elem * float(D)) : 0.f;
    lpx[row] = accumulate ? lpx[row] + val : val;
  }
}

template <typename T, int DIST>
__device__ __forceinline__ void lpx_bwd_body(const T* __restrict__ recon, const float* __restrict__ x,
                                             const float* __restrict__ coef, const float* __restrict__ g_loss,
                                             T* __restrict__ g_recon, int C, int K, int B, int64_t D, float inv_s, float rescale,
                                             const uint8_t* __restrict__ mask, int nchunks, int ksplit) {
  constexpr int VE = Vec<T>::N;
  constexpr int NV = kChunkElems / (32 * VE);
  const int lane = threadIdx.x & 31;
  int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (warp >= int64_t(C) * B * nchunks * ksplit) return;
  const int ks = int(warp % ksplit);
  warp /= ksplit;
  const int kper = (K + ksplit - 1) / ksplit;
  const int k_begin = ks * kper, k_end = min(K, k_begin + kper);
  const int chunk = int(warp % nchunks);
  const int64_t cb = warp / nchunks;
  const int b = int(cb % B), c = int(cb / B);
  const bool live = mask == nullptr || mask[b] != 0;
  const int64_t nvec = D / VE;
  const int64_t v0 = int64_t(chunk) * NV * 32 + lane;
  const float gl = *g_loss * rescale;
  float xr[NV][VE];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int64_t v = v0 + j * 32;
    if (live && v < nvec) {
      const float4* xp = reinterpret_cast<const float4*>(x + int64_t(b) * D + v * VE);
#pragma unroll
      for (int q = 0; q < VE / 4; ++q) {
        float4 t = __ldg(xp + q);
        xr[j][4 * q + 0] = t.x; xr[j][4 * q + 1] = t.y; xr[j][4 * q + 2] = t.z; xr[j][4 * q + 3] = t.w;
      }
    }
  }
  for (int k = k_begin; k < k_end; ++k) {
    const int64_t row = (int64_t(c) * K + k) * B + b;
    const float cf = live ? coef[row] * gl : 0.f;
    const T* rp = recon + row * D;
    T* gp = g_recon + row * D;
    uint4 rv[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int64_t v = v0 + j * 32;
      if (live && v < nvec) rv[j] = ld_stream(rp + v * VE);
    }
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int64_t v = v0 + j * 32;
      if (v < nvec) {
        float g[VE];
        if (live) {
          float r[VE];
          Vec<T>::unpack(rv[j], r);
#pragma unroll
          for (int e = 0; e < VE; ++e) g[e] = cf * lp_grad<DIST>(xr[j][e], r[e], inv_s);
        } else {
#pragma unroll
          for (int e = 0; e < VE; ++e) g[e] = 0.f;
        }
        st_stream(gp + v * VE, Vec<T>::pack(g));
      }
    }
  }
}
template <typename T, int DIST>
__global__ void __launch_bounds__(128) lpx_bwd_kernel(const T* __restrict__ recon, const float* __restrict__ x,
                                                      const float* __restrict__ coef, const float* __restrict__ g_loss,
                                                      T* __restrict__ g_recon, int C, int K, int B, int64_t D,
                                                      float inv_s, float rescale, const uint8_t* __restrict__ mask,
                                                      int nchunks, int ksplit) {
  lpx_bwd_body<T, DIST>(recon, x, coef, g_loss, g_recon, C, K, B, D, inv_s, rescale, mask, nchunks, ksplit);
}
template <typename T, int DIST>
__global__ void __launch_bounds__(128) lpx_bwd_multi_kernel(const __grid_constant__ LpxBatch bt, const float* __restrict__ coef,
                                                            const float* __restrict__ g_loss, int C, int K, int B, int64_t D,
                                                            int nchunks, int ksplit) {
  const int m = blockIdx.y;
  lpx_bwd_body<T, DIST>(static_cast<const T*>(bt.recon[m]), bt.x[m], coef, g_loss, static_cast<T*>(bt.g[m]), C, K, B, D, bt.inv_s[m],
                        bt.rescale[m], bt.mask[m], nchunks, ksplit);
}

template <typename T, int DIST>
__global__ void __launch_bounds__(128) lpx_bwd_scalar_kernel(const T* __restrict__ recon, const float* __restrict__ x,
                                                             const float* __restrict__ coef,
                                                             const float* __restrict__ g_loss, T* __restrict__ g_recon,
                                                             int C, int K, int B, int64_t D, float inv_s, float rescale,
                                                             const uint8_t* __restrict__ mask) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (row >= int64_t(C) * K * B) return;
  const int b = int(row % B);
  const bool live = mask == nullptr || mask[b] != 0;
  const float cf = live ? coef[row] * *g_loss * rescale : 0.f;
  for (int64_t d = lane; d < D; d += 32) {
    float g = live ? cf * lp_grad<DIST>(x[int64_t(b) * D + d], Vec<T>::load1(recon + row * D + d), inv_s) : 0.f;
    Vec<T>::store1(g_recon + row * D + d, g);
  }
}

// ---- latent log-densities -----------------------------------------------------------------------
template <int KIND>
__device__ __forceinline__ float lat_lp(float x, float mu, float s) {
  if (KIND == MV_LATENT_LAPLACE) return -logf(2.f * s) - fabsf(x - mu) / s;
  float t = (x - mu) / s;
  return -0.5f * t * t - logf(s) - 0.5f * kLog2Pi;
}
// d/dx of lat_lp (d/dmu is the negative)
template <int KIND>
__device__ __forceinline__ float lat_dx(float x, float mu, float s) {
  if (KIND == MV_LATENT_LAPLACE) {
    float t = x - mu;
    return t > 0.f ? -1.f / s : (t < 0.f ? 1.f / s : 0.f);
  }
  return -(x - mu) / (s * s);
}
template <int KIND>
__device__ __forceinline__ float lat_ds(float x, float mu, float s) {
  if (KIND == MV_LATENT_LAPLACE) return -1.f / s + fabsf(x - mu) / (s * s);
  float t = x - mu;
  return -1.f / s + t * t / (s * s * s);
}

constexpr int kMaxC = 8;  // modalities handled by the latent kernel (reference configs: <= 5)

// one block of kLwWarps warps per (conditioning modality, sample); the K importance samples are spread over its warps
constexpr int kLwWarps = 8;
template <int KIND>
__global__ void __launch_bounds__(32 * kLwWarps) moe_lw_kernel(
    const float* __restrict__ u, const float* __restrict__ w, const float* __restrict__ mu_u,
    const float* __restrict__ sig_u, const float* __restrict__ mu_w, const float* __restrict__ sig_w,
    const float* __restrict__ pz_mean, const float* __restrict__ pz_std, const float* __restrict__ lpx,
    const uint8_t* __restrict__ masks, float* __restrict__ lw, float* __restrict__ wk, float* __restrict__ coef,
    float* __restrict__ loss_b, float* __restrict__ g_u, float* __restrict__ g_w, float* __restrict__ g_mu_u,
    float* __restrict__ g_sig_u, float* __restrict__ g_mu_w, float* __restrict__ g_sig_w, float* __restrict__ g_pz_std,
    int C, int K, int B, int L, int Lw, int loss_kind, float beta, int detach_post, int skip_u_prior) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // One block per (conditioning modality c, sample b), its warps take the K importance samples round-robin: C * min(K, 8)
  // times the parallelism of a warp per sample (the kernel is a chain of dependent reductions: it wants warps, not bytes).
  // What the warps of a sample share (its loss, the prior-scale gradient, the posterior-parameter gradients through the MoE
  // term and of the private code) is accumulated with atomics into buffers the host zero-fills before the launch.
  __shared__ float s_mx[kLwWarps];
  const int wi = blockIdx.x;
  const int c = wi / B, b = wi - c * B;
  int nm = 0;
  bool avail[kMaxC];
#pragma unroll
  for (int m = 0; m < kMaxC; ++m) {
    avail[m] = m < C && (masks == nullptr || masks[m * B + b] != 0);
    nm += avail[m] ? 1 : 0;
  }
  const float log_nm = logf(float(nm > 0 ? nm : 1));
  const float inv_nm = nm > 0 ? 1.f / float(nm) : 0.f;
  const int LT = L + Lw;
  float loss_acc = 0.f;
  {
    // ---- pass 1: lw[c,k,b] for all k -----------------------------------------------------------
    float mx = -INFINITY;
    for (int k = warp; k < K; k += kLwWarps) {
      const int64_t row = (int64_t(c) * K + k) * B + b;
      float val = 0.f;
      if (avail[c]) {
        float lpz = 0.f, lqw = 0.f, lq[kMaxC];
#pragma unroll
        for (int m = 0; m < kMaxC; ++m) lq[m] = 0.f;
        for (int l = lane; l < L; l += 32) {
          const float uu = u[row * L + l];
          if (!skip_u_prior) lpz += lat_lp<KIND>(uu, pz_mean[l], pz_std[l]);
#pragma unroll
          for (int m = 0; m < kMaxC; ++m)
            if (m < C && avail[m])
              lq[m] += lat_lp<KIND>(uu, mu_u[(int64_t(m) * B + b) * L + l], sig_u[(int64_t(m) * B + b) * L + l]);
        }
        for (int l = lane; l < Lw; l += 32) {
          const float ww = w[row * Lw + l];
          lpz += lat_lp<KIND>(ww, pz_mean[L + l], pz_std[L + l]);
          lqw += lat_lp<KIND>(ww, mu_w[(int64_t(c) * B + b) * Lw + l], sig_w[(int64_t(c) * B + b) * Lw + l]);
        }
        lpz = warp_sum(lpz);
        lqw = warp_sum(lqw);
        float mq = -INFINITY;
#pragma unroll
        for (int m = 0; m < kMaxC; ++m)
          if (m < C && avail[m]) {
            lq[m] = warp_sum(lq[m]);
            mq = fmaxf(mq, lq[m]);
          }
        float se = 0.f;
#pragma unroll
        for (int m = 0; m < kMaxC; ++m)
          if (m < C && avail[m]) se += expf(lq[m] - mq);
        const float lqu = mq + logf(se) - log_nm;
        val = lpx[row] + beta * (lpz - lqu - lqw);
      }
      if (lane == 0) lw[row] = val;
      mx = fmaxf(mx, val);
    }
    if (lane == 0) s_mx[warp] = mx;
    __syncthreads();   // every warp's lw values (global) and maxima (shared) are visible to the whole block
#pragma unroll
    for (int i = 0; i < kLwWarps; ++i) mx = fmaxf(mx, s_mx[i]);
    // ---- softmax over k ------------------------------------------------------------------------
    // The weights are normalised exactly (wk = e_k / sum e) and the DReG term is evaluated relative to the maximum,
    // sum_k wk*lw = mx + sum_k wk*(lw - mx): with |lw| ~ 1e4 (D = 12288) the textbook form exp(lw - fl(logsumexp)) leaves
    // sum_k wk = 1 +- 5e-4 (half an ulp of lse) and that error multiplies |lw|.
    float se = 0.f;
    for (int k = lane; k < K; k += 32) se += expf(lw[(int64_t(c) * K + k) * B + b] - mx);
    se = warp_sum(se);
    const float inv_se = 1.f / se;
    const float lse = mx + logf(se);
    float term = 0.f;  // sum_k wk*(lw - mx) (DReG)
    for (int k = lane; k < K; k += 32) {
      const int64_t row = (int64_t(c) * K + k) * B + b;
      const float v = lw[row];
      const float wgt = expf(v - mx) * inv_se;
      if (warp == 0) {   // every warp computes the same weights (it needs them below); warp 0 publishes them
        wk[row] = wgt;
        coef[row] = avail[c] ? -wgt * inv_nm : 0.f;
      }
      term += wgt * (v - mx);
    }
    term = mx + warp_sum(term);
    if (loss_kind == MV_LOSS_IWAE) term = lse - logf(float(K));
    if (avail[c] && warp == 0) loss_acc += term;
    // ---- pass 2: unit gradients of the latent terms ----------------------------------------------
    for (int k = warp; k < K; k += kLwWarps) {
      const int64_t row = (int64_t(c) * K + k) * B + b;
      if (!avail[c]) {
        for (int l = lane; l < L; l += 32) g_u[row * L + l] = 0.f;
        for (int l = lane; l < Lw; l += 32) g_w[row * Lw + l] = 0.f;
        continue;
      }
      const float cb = -expf(lw[row] - mx) * inv_se * inv_nm * beta;  // coef[row] * beta = d loss / d (lpz - lqu - lqw)
      // MoE responsibilities sm_m = softmax_m(lq_m)
      float lq[kMaxC];
#pragma unroll
      for (int m = 0; m < kMaxC; ++m) lq[m] = 0.f;
      for (int l = lane; l < L; l += 32) {
        const float uu = u[row * L + l];
#pragma unroll
        for (int m = 0; m < kMaxC; ++m)
          if (m < C && avail[m])
            lq[m] += lat_lp<KIND>(uu, mu_u[(int64_t(m) * B + b) * L + l], sig_u[(int64_t(m) * B + b) * L + l]);
      }
      float mq = -INFINITY;
#pragma unroll
      for (int m = 0; m < kMaxC; ++m)
        if (m < C && avail[m]) {
          lq[m] = warp_sum(lq[m]);
          mq = fmaxf(mq, lq[m]);
        }
      float sq = 0.f;
#pragma unroll
      for (int m = 0; m < kMaxC; ++m)
        if (m < C && avail[m]) {
          lq[m] = expf(lq[m] - mq);
          sq += lq[m];
        }
      const float inv_sq = 1.f / sq;
      for (int l = lane; l < L; l += 32) {
        const float uu = u[row * L + l];
        const float pm = pz_mean[l], ps = pz_std[l];
        float gu = skip_u_prior ? 0.f : lat_dx<KIND>(uu, pm, ps);
        if (!skip_u_prior) atomicAdd(g_pz_std + int64_t(b) * LT + l, cb * lat_ds<KIND>(uu, pm, ps));
#pragma unroll
        for (int m = 0; m < kMaxC; ++m)
          if (m < C && avail[m]) {
            const int64_t pi = (int64_t(m) * B + b) * L + l;
            const float mm = mu_u[pi], ss = sig_u[pi];
            const float r = lq[m] * inv_sq;
            const float dx = lat_dx<KIND>(uu, mm, ss);
            gu -= r * dx;
            if (!detach_post) {
              atomicAdd(g_mu_u + pi, cb * r * dx);  // -(cb) * r * dlq/dmu, dlq/dmu = -dx
              atomicAdd(g_sig_u + pi, -cb * r * lat_ds<KIND>(uu, mm, ss));
            }
          }
        g_u[row * L + l] = cb * gu;
      }
      for (int l = lane; l < Lw; l += 32) {
        const float ww = w[row * Lw + l];
        const float pm = pz_mean[L + l], ps = pz_std[L + l];
        const int64_t pi = (int64_t(c) * B + b) * Lw + l;
        const float mm = mu_w[pi], ss = sig_w[pi];
        const float dx = lat_dx<KIND>(ww, mm, ss);
        g_w[row * Lw + l] = cb * (lat_dx<KIND>(ww, pm, ps) - dx);
        atomicAdd(g_pz_std + int64_t(b) * LT + L + l, cb * lat_ds<KIND>(ww, pm, ps));
        if (!detach_post) {   // the block's warps (different k) share these addresses
          atomicAdd(g_mu_w + pi, cb * dx);
          atomicAdd(g_sig_w + pi, -cb * lat_ds<KIND>(ww, mm, ss));
        }
      }
    }
  }
  if (lane == 0 && warp == 0 && avail[c]) atomicAdd(loss_b + b, -loss_acc * inv_nm);
}

}  // namespace mv

using namespace mv;

static void lp_consts(int dist, float s, float* mul, float* add) {
  if (dist == MV_DIST_NORMAL) {
    *mul = -1.f / (2.f * s * s);
    *add = -logf(s) - 0.5f * kLog2Pi;
  } else if (dist == MV_DIST_LAPLACE) {
    *mul = -1.f / s;
    *add = -logf(2.f * s);
  } else {
    *mul = 1.f;
    *add = 0.f;
  }
}

// split the K samples of a (c, b, chunk) over several warps until ~24 warps per SM are in flight (HBM latency
// hiding): the smallest divisor of K that reaches that, else K itself
static int pick_ksplit(int64_t base_warps, int K) {
  const int64_t want = int64_t(148) * 12;  // each extra split re-reads the fp32 target row (2x a bf16 recon row)
  for (int ks = 1; ks <= K; ++ks)
    if (K % ks == 0 && base_warps * ks >= want) return ks;
  return K;
}

template <typename T>
static int launch_lpx_fwd(const void* recon, const float* x, float* lpx, int C, int K, int B, int64_t D, int dist,
                          float s, float rescale, const uint8_t* mask, int accumulate, cudaStream_t st) {
  float mul, add;
  lp_consts(dist, s, &mul, &add);
  constexpr int VE = Vec<T>::N;
  const T* r = static_cast<const T*>(recon);
  const bool vec_ok = (D % VE == 0) && (reinterpret_cast<uintptr_t>(recon) % 16 == 0) &&
                      (D % 4 == 0) && (reinterpret_cast<uintptr_t>(x) % 16 == 0);
  if (vec_ok && D * 4 <= 48 * 1024) {
    const int blocks = C * B;
    const size_t smem = size_t(D) * 4;
#define L_(DI) lpx_fwd_smem_kernel<T, DI><<<blocks, 128, smem, st>>>(r, x, lpx, C, K, B, D, mul, add, rescale, mask, accumulate)
    if (dist == MV_DIST_NORMAL) L_(MV_DIST_NORMAL);
    else if (dist == MV_DIST_LAPLACE) L_(MV_DIST_LAPLACE);
    else L_(MV_DIST_BERNOULLI);
#undef L_
  } else if (vec_ok) {
    const int nchunks = int((D + kChunkElems - 1) / kChunkElems);
    if (nchunks > 1 && !accumulate) cudaMemsetAsync(lpx, 0, sizeof(float) * size_t(C) * K * B, st);
    const int ksplit = pick_ksplit(int64_t(C) * B * nchunks, K);
    const int64_t warps = int64_t(C) * B * nchunks * ksplit;
    const int blocks = int((warps + 3) / 4);
#define L_(DI) lpx_fwd_kernel<T, DI><<<blocks, 128, 0, st>>>(r, x, lpx, C, K, B, D, mul, add, rescale, mask, accumulate, nchunks, ksplit)
    if (dist == MV_DIST_NORMAL) L_(MV_DIST_NORMAL);
    else if (dist == MV_DIST_LAPLACE) L_(MV_DIST_LAPLACE);
    else L_(MV_DIST_BERNOULLI);
#undef L_
  } else {
    const int64_t rows = int64_t(C) * K * B;
    const int blocks = int((rows + 3) / 4);
#define L_(DI) lpx_fwd_scalar_kernel<T, DI><<<blocks, 128, 0, st>>>(r, x, lpx, C, K, B, D, mul, add, rescale, mask, accumulate)
    if (dist == MV_DIST_NORMAL) L_(MV_DIST_NORMAL);
    else if (dist == MV_DIST_LAPLACE) L_(MV_DIST_LAPLACE);
    else L_(MV_DIST_BERNOULLI);
#undef L_
  }
  MV_CHECK_LAUNCH("mv_moe_lpx_fwd");
  return MV_OK;
}

template <typename T>
static int launch_lpx_bwd(const void* recon, const float* x, const float* coef, const float* g_loss, void* g_recon,
                          int C, int K, int B, int64_t D, int dist, float s, float rescale, const uint8_t* mask,
                          cudaStream_t st) {
  constexpr int VE = Vec<T>::N;
  const T* r = static_cast<const T*>(recon);
  T* g = static_cast<T*>(g_recon);
  const float inv_s = 1.f / s;
  const bool vec_ok = (D % VE == 0) && (reinterpret_cast<uintptr_t>(recon) % 16 == 0) &&
                      (reinterpret_cast<uintptr_t>(g_recon) % 16 == 0) && (D % 4 == 0) &&
                      (reinterpret_cast<uintptr_t>(x) % 16 == 0);
  if (vec_ok) {
    const int nchunks = int((D + kChunkElems - 1) / kChunkElems);
    const int ksplit = pick_ksplit(int64_t(C) * B * nchunks, K);
    const int64_t warps = int64_t(C) * B * nchunks * ksplit;
    const int blocks = int((warps + 3) / 4);
#define L_(DI) lpx_bwd_kernel<T, DI><<<blocks, 128, 0, st>>>(r, x, coef, g_loss, g, C, K, B, D, inv_s, rescale, mask, nchunks, ksplit)
    if (dist == MV_DIST_NORMAL) L_(MV_DIST_NORMAL);
    else if (dist == MV_DIST_LAPLACE) L_(MV_DIST_LAPLACE);
    else L_(MV_DIST_BERNOULLI);
#undef L_
  } else {
    const int64_t rows = int64_t(C) * K * B;
    const int blocks = int((rows + 3) / 4);
#define L_(DI) lpx_bwd_scalar_kernel<T, DI><<<blocks, 128, 0, st>>>(r, x, coef, g_loss, g, C, K, B, D, inv_s, rescale, mask)
    if (dist == MV_DIST_NORMAL) L_(MV_DIST_NORMAL);
    else if (dist == MV_DIST_LAPLACE) L_(MV_DIST_LAPLACE);
    else L_(MV_DIST_BERNOULLI);
#undef L_
  }
  MV_CHECK_LAUNCH("mv_moe_lpx_bwd");
  return MV_OK;
}

extern "C" int mv_moe_lpx_fwd(const void* recon, int recon_dtype, const float* x, float* lpx, int C, int K, int B,
                              int64_t D, int dist, float dist_scale, float rescale, const uint8_t* mask_r,
                              int accumulate, void* stream) {
  MV_CHECK_ARG(recon && x && lpx, "mv_moe_lpx_fwd: null pointer");
  MV_CHECK_ARG(C > 0 && K > 0 && B > 0 && D > 0, "mv_moe_lpx_fwd: bad sizes C=%d K=%d B=%d D=%lld", C, K, B, (long long)D);
  MV_CHECK_ARG(dist >= MV_DIST_NORMAL && dist <= MV_DIST_BERNOULLI, "mv_moe_lpx_fwd: unsupported distribution %d", dist);
  MV_CHECK_ARG(dist == MV_DIST_BERNOULLI || dist_scale > 0.f, "mv_moe_lpx_fwd: scale must be > 0");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (recon_dtype == MV_F32) return launch_lpx_fwd<float>(recon, x, lpx, C, K, B, D, dist, dist_scale, rescale, mask_r, accumulate, st);
  if (recon_dtype == MV_BF16) return launch_lpx_fwd<__nv_bfloat16>(recon, x, lpx, C, K, B, D, dist, dist_scale, rescale, mask_r, accumulate, st);
  mv::set_error("mv_moe_lpx_fwd: unsupported dtype %d", recon_dtype);
  return MV_ERR_UNSUPPORTED;
}

extern "C" int mv_moe_lpx_bwd(const void* recon, int recon_dtype, const float* x, const float* coef,
                              const float* g_loss, void* g_recon, int C, int K, int B, int64_t D, int dist,
                              float dist_scale, float rescale, const uint8_t* mask_r, void* stream) {
  MV_CHECK_ARG(recon && x && coef && g_loss && g_recon, "mv_moe_lpx_bwd: null pointer");
  MV_CHECK_ARG(C > 0 && K > 0 && B > 0 && D > 0, "mv_moe_lpx_bwd: bad sizes");
  MV_CHECK_ARG(dist >= MV_DIST_NORMAL && dist <= MV_DIST_BERNOULLI, "mv_moe_lpx_bwd: unsupported distribution %d", dist);
  MV_CHECK_ARG(dist == MV_DIST_BERNOULLI || dist_scale > 0.f, "mv_moe_lpx_bwd: scale must be > 0");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (recon_dtype == MV_F32) return launch_lpx_bwd<float>(recon, x, coef, g_loss, g_recon, C, K, B, D, dist, dist_scale, rescale, mask_r, st);
  if (recon_dtype == MV_BF16) return launch_lpx_bwd<__nv_bfloat16>(recon, x, coef, g_loss, g_recon, C, K, B, D, dist, dist_scale, rescale, mask_r, st);
  mv::set_error("mv_moe_lpx_bwd: unsupported dtype %d", recon_dtype);
  return MV_ERR_UNSUPPORTED;
}

// ---- all reconstructed modalities in ONE launch (same D / dtype / distribution family): blockIdx.y = modality -------------
template <typename T>
static bool lpx_multi_ok(int n_mod, const void* const* recon, const float* const* x, void* const* g, int64_t D) {
  constexpr int VE = Vec<T>::N;
  if (n_mod < 1 || n_mod > kMaxBatchMods || D % VE != 0 || D % 4 != 0) return false;
  for (int m = 0; m < n_mod; ++m) {
    if (!recon[m] || !x[m] || reinterpret_cast<uintptr_t>(recon[m]) % 16 || reinterpret_cast<uintptr_t>(x[m]) % 16) return false;
    if (g && (!g[m] || reinterpret_cast<uintptr_t>(g[m]) % 16)) return false;
  }
  return true;
}

extern "C" int mv_moe_lpx_fwd_multi(int n_mod, const void* const* recon, int recon_dtype, const float* const* x, float* lpx, int C,
                                    int K, int B, int64_t D, int dist, const float* dist_scale, const float* rescale,
                                    const uint8_t* const* mask_r, void* stream) {
  MV_CHECK_ARG(recon && x && lpx && dist_scale && rescale && n_mod >= 1, "mv_moe_lpx_fwd_multi: null pointer");
  MV_CHECK_ARG(C > 0 && K > 0 && B > 0 && D > 0, "mv_moe_lpx_fwd_multi: bad sizes");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool batched = recon_dtype == MV_BF16 ? lpx_multi_ok<__nv_bfloat16>(n_mod, recon, x, nullptr, D)
                                              : (recon_dtype == MV_F32 && lpx_multi_ok<float>(n_mod, recon, x, nullptr, D));
  if (!batched || D * 4 > 48 * 1024) {   // general case: one launch per modality, accumulating
    for (int m = 0; m < n_mod; ++m) {
      const int rc = mv_moe_lpx_fwd(recon[m], recon_dtype, x[m], lpx, C, K, B, D, dist, dist_scale[m], rescale[m],
                                    mask_r ? mask_r[m] : nullptr, m > 0 ? 1 : 0, stream);
      if (rc != MV_OK) return rc;
    }
    return MV_OK;
  }
  MV_CHECK_ARG(dist >= MV_DIST_NORMAL && dist <= MV_DIST_BERNOULLI, "mv_moe_lpx_fwd_multi: unsupported distribution %d", dist);
  LpxBatch bt{};
  for (int m = 0; m < n_mod; ++m) {
    MV_CHECK_ARG(dist == MV_DIST_BERNOULLI || dist_scale[m] > 0.f, "mv_moe_lpx_fwd_multi: scale must be > 0");
    bt.recon[m] = recon[m]; bt.x[m] = x[m]; bt.mask[m] = mask_r ? mask_r[m] : nullptr; bt.rescale[m] = rescale[m];
    lp_consts(dist, dist_scale[m], &bt.mul[m], &bt.add[m]);
  }
  cudaMemsetAsync(lpx, 0, sizeof(float) * size_t(C) * K * B, st);
  const dim3 grid(C * B, n_mod);
  const size_t smem = size_t(D) * 4;
#define L_(TT, DI) lpx_fwd_smem_multi_kernel<TT, DI><<<grid, 128, smem, st>>>(bt, lpx, C, K, B, D)
#define LT_(DI)                                     \
  do {                                              \
    if (recon_dtype == MV_BF16) L_(__nv_bfloat16, DI); \
    else L_(float, DI);                             \
  } while (0)
  if (dist == MV_DIST_NORMAL) LT_(MV_DIST_NORMAL);
  else if (dist == MV_DIST_LAPLACE) LT_(MV_DIST_LAPLACE);
  else LT_(MV_DIST_BERNOULLI);
#undef LT_
#undef L_
  MV_CHECK_LAUNCH("mv_moe_lpx_fwd_multi");
  return MV_OK;
}

extern "C" int mv_moe_lpx_bwd_multi(int n_mod, const void* const* recon, int recon_dtype, const float* const* x, const float* coef,
                                    const float* g_loss, void* const* g_recon, int C, int K, int B, int64_t D, int dist,
                                    const float* dist_scale, const float* rescale, const uint8_t* const* mask_r, void* stream) {
  MV_CHECK_ARG(recon && x && coef && g_loss && g_recon && dist_scale && rescale && n_mod >= 1, "mv_moe_lpx_bwd_multi: null pointer");
  MV_CHECK_ARG(C > 0 && K > 0 && B > 0 && D > 0, "mv_moe_lpx_bwd_multi: bad sizes");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool batched = recon_dtype == MV_BF16 ? lpx_multi_ok<__nv_bfloat16>(n_mod, recon, x, g_recon, D)
                                              : (recon_dtype == MV_F32 && lpx_multi_ok<float>(n_mod, recon, x, g_recon, D));
  if (!batched) {
    for (int m = 0; m < n_mod; ++m) {
      const int rc = mv_moe_lpx_bwd(recon[m], recon_dtype, x[m], coef, g_loss, g_recon[m], C, K, B, D, dist, dist_scale[m], rescale[m],
                                    mask_r ? mask_r[m] : nullptr, stream);
      if (rc != MV_OK) return rc;
    }
    return MV_OK;
  }
  MV_CHECK_ARG(dist >= MV_DIST_NORMAL && dist <= MV_DIST_BERNOULLI, "mv_moe_lpx_bwd_multi: unsupported distribution %d", dist);
  LpxBatch bt{};
  for (int m = 0; m < n_mod; ++m) {
    MV_CHECK_ARG(dist == MV_DIST_BERNOULLI || dist_scale[m] > 0.f, "mv_moe_lpx_bwd_multi: scale must be > 0");
    bt.recon[m] = recon[m]; bt.x[m] = x[m]; bt.g[m] = g_recon[m]; bt.mask[m] = mask_r ? mask_r[m] : nullptr;
    bt.rescale[m] = rescale[m]; bt.inv_s[m] = dist == MV_DIST_BERNOULLI ? 1.f : 1.f / dist_scale[m];
  }
  const int nchunks = int((D + kChunkElems - 1) / kChunkElems);
  const int ksplit = pick_ksplit(int64_t(C) * B * nchunks * n_mod, K);
  const int64_t warps = int64_t(C) * B * nchunks * ksplit;
  const dim3 grid(unsigned((warps + 3) / 4), n_mod);
#define L_(TT, DI) lpx_bwd_multi_kernel<TT, DI><<<grid, 128, 0, st>>>(bt, coef, g_loss, C, K, B, D, nchunks, ksplit)
#define LT_(DI)                                     \
  do {                                              \
    if (recon_dtype == MV_BF16) L_(__nv_bfloat16, DI); \
    else L_(float, DI);                             \
  } while (0)
  if (dist == MV_DIST_NORMAL) LT_(MV_DIST_NORMAL);
  else if (dist == MV_DIST_LAPLACE) LT_(MV_DIST_LAPLACE);
  else LT_(MV_DIST_BERNOULLI);
#undef LT_
#undef L_
  MV_CHECK_LAUNCH("mv_moe_lpx_bwd_multi");
  return MV_OK;
}

extern "C" int mv_moe_lw_fwd(const float* u, const float* w, const float* mu_u, const float* sig_u, const float* mu_w,
                             const float* sig_w, const float* pz_mean, const float* pz_std, const float* lpx,
                             const uint8_t* masks, float* lw, float* wk, float* coef, float* loss_b, float* g_u,
                             float* g_w, float* g_mu_u, float* g_sig_u, float* g_mu_w, float* g_sig_w, float* g_pz_std,
                             int C, int K, int B, int L, int Lw, int latent_kind, int loss_kind, float beta,
                             int detach_post, int skip_u_prior, void* stream) {
  MV_CHECK_ARG(u && mu_u && sig_u && pz_mean && pz_std && lpx && lw && wk && coef && loss_b && g_u && g_mu_u &&
                   g_sig_u && g_pz_std, "mv_moe_lw_fwd: null pointer");
  MV_CHECK_ARG(Lw == 0 || (w && mu_w && sig_w && g_w && g_mu_w && g_sig_w), "mv_moe_lw_fwd: null private-latent pointer");
  MV_CHECK_ARG(C > 0 && C <= kMaxC, "mv_moe_lw_fwd: 1 <= n_modalities <= %d required, got %d", kMaxC, C);
  MV_CHECK_ARG(K > 0 && B > 0 && L > 0 && Lw >= 0, "mv_moe_lw_fwd: bad sizes");
  MV_CHECK_ARG(loss_kind == MV_LOSS_IWAE || loss_kind == MV_LOSS_DREG, "mv_moe_lw_fwd: unknown loss %d", loss_kind);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // accumulators shared by the C warps of a sample (atomics in the kernel): zero-filled here
  cudaMemsetAsync(loss_b, 0, sizeof(float) * size_t(B), st);
  cudaMemsetAsync(g_pz_std, 0, sizeof(float) * size_t(B) * (L + Lw), st);
  cudaMemsetAsync(g_mu_u, 0, sizeof(float) * size_t(C) * B * L, st);
  cudaMemsetAsync(g_sig_u, 0, sizeof(float) * size_t(C) * B * L, st);
  if (Lw > 0) {
    cudaMemsetAsync(g_mu_w, 0, sizeof(float) * size_t(C) * B * Lw, st);
    cudaMemsetAsync(g_sig_w, 0, sizeof(float) * size_t(C) * B * Lw, st);
  }
  const int blocks = C * B;   // one block per (conditioning modality, sample)
  if (latent_kind == MV_LATENT_LAPLACE)
    moe_lw_kernel<MV_LATENT_LAPLACE><<<blocks, 32 * kLwWarps, 0, st>>>(u, w, mu_u, sig_u, mu_w, sig_w, pz_mean, pz_std, lpx, masks, lw, wk, coef, loss_b, g_u, g_w, g_mu_u, g_sig_u, g_mu_w, g_sig_w, g_pz_std, C, K, B, L, Lw, loss_kind, beta, detach_post, skip_u_prior);
  else if (latent_kind == MV_LATENT_NORMAL)
    moe_lw_kernel<MV_LATENT_NORMAL><<<blocks, 32 * kLwWarps, 0, st>>>(u, w, mu_u, sig_u, mu_w, sig_w, pz_mean, pz_std, lpx, masks, lw, wk, coef, loss_b, g_u, g_w, g_mu_u, g_sig_u, g_mu_w, g_sig_w, g_pz_std, C, K, B, L, Lw, loss_kind, beta, detach_post, skip_u_prior);
  else {
    mv::set_error("mv_moe_lw_fwd: unknown latent kind %d", latent_kind);
    return MV_ERR_UNSUPPORTED;
  }
  MV_CHECK_LAUNCH("mv_moe_lw_fwd");
  return MV_OK;
}
