// Reparameterised sampling of the mixture-of-experts models in ONE launch per direction (MMVAE, MMVAE+, CMVAE):
//
//   sig_u = std(lv_u),  sig_w = std(lv_w)                                 (_log_var_to_std: mmvaePlus_model.py:113-123,
//                                                                          mmvae_model.py:66-74)
//   U[c,k,b,:] = mu_u[c,b,:] + sig_u[c,b,:] * e_u[c,k,b,:]               (K rsample draws of every unimodal posterior:
//   W[c,k,b,:] = mu_w[c,b,:] + sig_w[c,b,:] * e_w[c,k,b,:]                mmvaePlus_model.py:136-160, mmvae_model.py:111-120)
//   Z[r][c,k,b,:] = cat( U[c,k,b,:],  r == c ? W[c,k,b,:]                 (decoder input of modality r conditioned on modality c;
//                                            : pm[r,:] + sp[r,:] * e_x[c,j(r),k,b,:] )   private code from r's prior when r != c:
//                                                                          mmvaePlus_model.py:163-186)
//
// and the matching backward, which also applies the DReG rule (the gradient reaching the posterior samples u, w is multiplied
// once more by the importance weights wk: mmvaePlus_model.py:330-338) and the derivative of std().  This replaces ~30 small ATen
// kernels per modality and direction (softmax / mul / add / stack / cat and their autograd twins).  One warp per (c, b): the
// rows are 20..64 floats, so everything is warp shuffles; the tensors are a few MB: the kernels are bound by launch latency.
// std kinds: 0 = softmax(lv) * dim + 1e-6 (laplace_with_softmax), 1 = exp(lv / 2) (normal), 2 = softplus(lv) + 1e-6.
#include "common.cuh"

namespace mv {

constexpr int kSampleMaxDim = 128;   // latent dimensions handled per row (4 per lane)
constexpr int kSampleMaxC = 8;

// std of one row held as v[j] = lv[lane + 32 j]; returns the softmax normaliser pieces needed by the backward
template <int KIND>
__device__ __forceinline__ void row_std(const float* lv, int dim, int lane, float* sg) {
  if (KIND == 0) {
    float mx = -INFINITY;
    for (int j = 0; j < 4; ++j) {
      const int l = lane + 32 * j;
      if (l < dim) mx = fmaxf(mx, lv[l]);
    }
    mx = warp_max(mx);
    float e[4], se = 0.f;
    for (int j = 0; j < 4; ++j) {
      const int l = lane + 32 * j;
      e[j] = l < dim ? expf(lv[l] - mx) : 0.f;
      se += e[j];
    }
    se = warp_sum(se);
    const float sc = float(dim) / se;
    for (int j = 0; j < 4; ++j) sg[j] = e[j] * sc + 1e-6f;
  } else {
    for (int j = 0; j < 4; ++j) {
      const int l = lane + 32 * j;
      const float x = l < dim ? lv[l] : 0.f;
      sg[j] = KIND == 1 ? expf(0.5f * x) : ((x > 20.f ? x : log1pf(expf(x))) + 1e-6f);   // torch softplus threshold 20
    }
  }
}

// d loss / d lv from d loss / d std for one row (g[j], sg[j] at l = lane + 32 j)
template <int KIND>
__device__ __forceinline__ void row_std_bwd(const float* lv, const float* sg, const float* g, int dim, int lane, float* glv) {
  if (KIND == 0) {
    // std = dim * p + 1e-6 with p = softmax(lv): g_lv = dim * p * (g - sum(g * p))
    float dot = 0.f, p[4];
    for (int j = 0; j < 4; ++j) {
      const int l = lane + 32 * j;
      p[j] = l < dim ? (sg[j] - 1e-6f) / float(dim) : 0.f;
      dot += l < dim ? g[j] * p[j] : 0.f;
    }
    dot = warp_sum(dot);
    for (int j = 0; j < 4; ++j) glv[j] = float(dim) * p[j] * (g[j] - dot);
  } else {
    for (int j = 0; j < 4; ++j) {
      const int l = lane + 32 * j;
      const float x = l < dim ? lv[l] : 0.f;
      glv[j] = KIND == 1 ? g[j] * 0.5f * sg[j] : g[j] * (x > 20.f ? 1.f : 1.f / (1.f + expf(-x)));
    }
  }
}

template <int KIND>
__global__ void __launch_bounds__(128) moe_sample_fwd_kernel(
    const float* __restrict__ mu_u, const float* __restrict__ lv_u, const float* __restrict__ mu_w, const float* __restrict__ lv_w,
    const float* __restrict__ pm, const float* __restrict__ sp, const float* __restrict__ e_u, const float* __restrict__ e_w,
    const float* __restrict__ e_x, float* __restrict__ sig_u, float* __restrict__ sig_w, float* __restrict__ U,
    float* __restrict__ W, float* __restrict__ Z, int C, int K, int B, int L, int Lw) {
  const int lane = threadIdx.x & 31;
  const int wi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (wi >= C * B) return;
  const int c = wi / B, b = wi - c * B;
  const int64_t cb = int64_t(c) * B + b;
  float su[4], sw[4], mu[4], mw[4];
  row_std<KIND>(lv_u + cb * L, L, lane, su);
  for (int j = 0; j < 4; ++j) {
    const int l = lane + 32 * j;
    mu[j] = l < L ? mu_u[cb * L + l] : 0.f;
    if (l < L) sig_u[cb * L + l] = su[j];
  }
  if (Lw > 0) {
    row_std<KIND>(lv_w + cb * Lw, Lw, lane, sw);
    for (int j = 0; j < 4; ++j) {
      const int l = lane + 32 * j;
      mw[j] = l < Lw ? mu_w[cb * Lw + l] : 0.f;
      if (l < Lw) sig_w[cb * Lw + l] = sw[j];
    }
  }
  const int LT = L + Lw;
  for (int k = 0; k < K; ++k) {
    const int64_t row = (int64_t(c) * K + k) * B + b;
    for (int j = 0; j < 4; ++j) {
      const int l = lane + 32 * j;
      if (l < L) {
        const float u = fmaf(su[j], e_u[row * L + l], mu[j]);
        U[row * L + l] = u;
        if (Z)
          for (int r = 0; r < C; ++r) Z[(int64_t(r) * C * K * B + row) * LT + l] = u;
      }
      if (l < Lw) {
        const float w = fmaf(sw[j], e_w[row * Lw + l], mw[j]);
        W[row * Lw + l] = w;
        for (int r = 0, jx = 0; r < C; ++r) {
          float v = w;
          if (r != c) {
            v = fmaf(sp[r * Lw + l], e_x[(((int64_t(c) * (C - 1) + jx) * K + k) * B + b) * Lw + l], pm[r * Lw + l]);
            ++jx;
          }
          Z[(int64_t(r) * C * K * B + row) * LT + L + l] = v;
        }
      }
    }
  }
}

template <int KIND>
__global__ void __launch_bounds__(128) moe_sample_bwd_kernel(
    const float* __restrict__ lv_u, const float* __restrict__ lv_w, const float* __restrict__ sig_u, const float* __restrict__ sig_w,
    const float* __restrict__ e_u, const float* __restrict__ e_w, const float* __restrict__ e_x, const float* __restrict__ g_U,
    const float* __restrict__ g_W, const float* __restrict__ g_Z, const float* __restrict__ g_sig_u, const float* __restrict__ g_sig_w,
    const float* __restrict__ wk, float* __restrict__ g_mu_u, float* __restrict__ g_lv_u, float* __restrict__ g_mu_w,
    float* __restrict__ g_lv_w, float* __restrict__ g_sp, int C, int K, int B, int L, int Lw) {
  const int lane = threadIdx.x & 31;
  const int wi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (wi >= C * B) return;
  const int c = wi / B, b = wi - c * B;
  const int64_t cb = int64_t(c) * B + b;
  const int LT = L + Lw;
  float gmu[4] = {0, 0, 0, 0}, gsu[4] = {0, 0, 0, 0}, gmw[4] = {0, 0, 0, 0}, gsw[4] = {0, 0, 0, 0};
  float gsp[kSampleMaxC][4];
#pragma unroll
  for (int r = 0; r < kSampleMaxC; ++r)
    for (int j = 0; j < 4; ++j) gsp[r][j] = 0.f;
  for (int k = 0; k < K; ++k) {
    const int64_t row = (int64_t(c) * K + k) * B + b;
    const float wgt = wk ? wk[row] : 1.f;   // DReG: the samples' gradient is multiplied once more by the importance weight
    for (int j = 0; j < 4; ++j) {
      const int l = lane + 32 * j;
      if (l < L) {
        float g = g_U ? g_U[row * L + l] : 0.f;
        if (g_Z)
          for (int r = 0; r < C; ++r) g += g_Z[(int64_t(r) * C * K * B + row) * LT + l];
        g *= wgt;
        gmu[j] += g;
        gsu[j] += g * e_u[row * L + l];
      }
      if (l < Lw) {
        float g = g_W ? g_W[row * Lw + l] : 0.f;
#pragma unroll
        for (int r = 0, jx = 0; r < kSampleMaxC; ++r) {
          if (r >= C) break;
          const float gz = g_Z ? g_Z[(int64_t(r) * C * K * B + row) * LT + L + l] : 0.f;
          if (r == c) {
            g += gz;
          } else {
            gsp[r][j] += gz * e_x[(((int64_t(c) * (C - 1) + jx) * K + k) * B + b) * Lw + l];   // prior samples: no DReG factor
            ++jx;
          }
        }
        g *= wgt;
        gmw[j] += g;
        gsw[j] += g * e_w[row * Lw + l];
      }
    }
  }
  float sg[4], glv[4];
  for (int j = 0; j < 4; ++j) {
    const int l = lane + 32 * j;
    sg[j] = l < L ? sig_u[cb * L + l] : 0.f;
    if (l < L) {
      gsu[j] += g_sig_u ? g_sig_u[cb * L + l] : 0.f;
      g_mu_u[cb * L + l] = gmu[j];
    }
  }
  row_std_bwd<KIND>(lv_u + cb * L, sg, gsu, L, lane, glv);
  for (int j = 0; j < 4; ++j) {
    const int l = lane + 32 * j;
    if (l < L) g_lv_u[cb * L + l] = glv[j];
  }
  if (Lw > 0) {
    for (int j = 0; j < 4; ++j) {
      const int l = lane + 32 * j;
      sg[j] = l < Lw ? sig_w[cb * Lw + l] : 0.f;
      if (l < Lw) {
        gsw[j] += g_sig_w ? g_sig_w[cb * Lw + l] : 0.f;
        g_mu_w[cb * Lw + l] = gmw[j];
      }
    }
    row_std_bwd<KIND>(lv_w + cb * Lw, sg, gsw, Lw, lane, glv);
    for (int j = 0; j < 4; ++j) {
      const int l = lane + 32 * j;
      if (l < Lw) {
        g_lv_w[cb * Lw + l] = glv[j];
#pragma unroll
        for (int r = 0; r < kSampleMaxC; ++r)
          if (r < C && r != c && g_sp) atomicAdd(g_sp + r * Lw + l, gsp[r][j]);
      }
    }
  }
}

}  // namespace mv

using namespace mv;

extern "C" int mv_moe_sample_fwd(const float* mu_u, const float* lv_u, const float* mu_w, const float* lv_w, const float* prior_mean,
                                 const float* prior_std, const float* noise_u, const float* noise_w, const float* noise_x, int std_kind,
                                 float* sig_u, float* sig_w, float* U, float* W, float* Z, int C, int K, int B, int L, int Lw,
                                 void* stream) {
  MV_CHECK_ARG(mu_u && lv_u && noise_u && sig_u && U, "mv_moe_sample_fwd: null pointer");
  MV_CHECK_ARG(Lw == 0 || (mu_w && lv_w && noise_w && sig_w && W && Z && prior_mean && prior_std && (C == 1 || noise_x)),
               "mv_moe_sample_fwd: null private-latent pointer");
  MV_CHECK_ARG(C >= 1 && C <= kSampleMaxC && K >= 1 && B >= 1 && L >= 1 && L <= kSampleMaxDim && Lw >= 0 && Lw <= kSampleMaxDim,
               "mv_moe_sample_fwd: bad sizes (C <= %d, latent dims <= %d)", kSampleMaxC, kSampleMaxDim);
  MV_CHECK_ARG(std_kind >= 0 && std_kind <= 2, "mv_moe_sample_fwd: unknown std kind %d", std_kind);
  const int blocks = (C * B + 3) / 4;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define L_(KD) moe_sample_fwd_kernel<KD><<<blocks, 128, 0, st>>>(mu_u, lv_u, mu_w, lv_w, prior_mean, prior_std, noise_u, noise_w, noise_x, sig_u, sig_w, U, W, Z, C, K, B, L, Lw)
  if (std_kind == 0) L_(0);
  else if (std_kind == 1) L_(1);
  else L_(2);
#undef L_
  MV_CHECK_LAUNCH("mv_moe_sample_fwd");
  return MV_OK;
}

extern "C" int mv_moe_sample_bwd(const float* lv_u, const float* lv_w, const float* sig_u, const float* sig_w, const float* noise_u,
                                 const float* noise_w, const float* noise_x, const float* g_U, const float* g_W, const float* g_Z,
                                 const float* g_sig_u, const float* g_sig_w, const float* wk, int std_kind, float* g_mu_u, float* g_lv_u,
                                 float* g_mu_w, float* g_lv_w, float* g_prior_std, int C, int K, int B, int L, int Lw, void* stream) {
  MV_CHECK_ARG(lv_u && sig_u && noise_u && g_mu_u && g_lv_u, "mv_moe_sample_bwd: null pointer");
  MV_CHECK_ARG(Lw == 0 || (lv_w && sig_w && noise_w && g_mu_w && g_lv_w && (C == 1 || noise_x)), "mv_moe_sample_bwd: null private-latent pointer");
  MV_CHECK_ARG(C >= 1 && C <= kSampleMaxC && K >= 1 && B >= 1 && L >= 1 && L <= kSampleMaxDim && Lw >= 0 && Lw <= kSampleMaxDim,
               "mv_moe_sample_bwd: bad sizes");
  MV_CHECK_ARG(std_kind >= 0 && std_kind <= 2, "mv_moe_sample_bwd: unknown std kind %d", std_kind);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (g_prior_std && Lw > 0) cudaMemsetAsync(g_prior_std, 0, sizeof(float) * size_t(C) * Lw, st);
  const int blocks = (C * B + 3) / 4;
#define L_(KD) moe_sample_bwd_kernel<KD><<<blocks, 128, 0, st>>>(lv_u, lv_w, sig_u, sig_w, noise_u, noise_w, noise_x, g_U, g_W, g_Z, g_sig_u, g_sig_w, wk, g_mu_u, g_lv_u, g_mu_w, g_lv_w, g_prior_std, C, K, B, L, Lw)
  if (std_kind == 0) L_(0);
  else if (std_kind == 1) L_(1);
  else L_(2);
#undef L_
  MV_CHECK_LAUNCH("mv_moe_sample_bwd");
  return MV_OK;
}
