// Fused posterior aggregation for the product-of-experts models (MVTCAE, MVAE, MoPoE):
// subset enumeration from a bitmask table, precision-weighted PoE merge (eps or "stable" form, optional
// N(0,I) prior expert), reparameterised sample of the selected subset, analytic KL(q_s || N(0,I)) of every
// subset weighted per sample, and MVTCAE's KL(joint || q_m) terms — one kernel forward, one backward
// (recompute, no saved intermediates).  Elementwise over (b, l) + a block reduction over l; no tensor cores.
//
// Reference: models/base/base_utils.py:122-147 (poe, stable_poe), models/mopoe/mopoe_model.py:108-145,249-350,
// models/mvae/mvae_model.py:53-113, models/mvtcae/mvtcae_model.py:42-108,134-169.
#include "common.cuh"

namespace mv {

constexpr int kMaxM = 8;

struct PoeAcc {
  float st, sm;  // sum of precisions, sum of mu * precision
};

// precision of expert m at this (b,l); 0 when masked (lv = +inf in the reference)
__device__ __forceinline__ float expert_T(float lv, bool stable, float eps) {
  return stable ? 0.f /* stable form: subset_poe works on the log-variances */ : 1.f / (expf(lv) + eps);
}

// PoE of the experts of one subset at one (b, l).  eps form (base_utils.py:122-130): precisions T_m = 1/(exp(lv)+eps).
// stable form (base_utils.py:133-147): ln var = -logsumexp(-lv) with the maximum taken out per subset, so log-variances
// far outside exp()'s fp32 range neither overflow nor underflow the sum.  rr (optional) = T_m / sum_T per expert.
__device__ __forceinline__ void subset_poe(uint32_t bits, bool prior, bool stable, float Tp, const float* T, const float* mm,
                                           const float* lvv, const bool* avail, float* pmu, float* var, float* plv,
                                           float* rr) {
  constexpr int kM = 8;
  float st, sm = 0.f;
  if (stable) {
    float shift = prior ? 0.f : -INFINITY;
#pragma unroll
    for (int m = 0; m < kM; ++m)
      if ((bits >> m & 1u) && avail[m]) shift = fmaxf(shift, -lvv[m]);
    st = prior ? expf(-shift) : 0.f;
    float t[kM];
#pragma unroll
    for (int m = 0; m < kM; ++m) {
      t[m] = ((bits >> m & 1u) && avail[m]) ? expf(-lvv[m] - shift) : 0.f;
      st += t[m];
      sm += mm[m] * t[m];
    }
    *pmu = sm / st;
    *plv = -shift - logf(st);
    *var = expf(*plv);
    if (rr) {
#pragma unroll
      for (int m = 0; m < kM; ++m) rr[m] = t[m] / st;
    }
    return;
  }
  st = prior ? Tp : 0.f;
#pragma unroll
  for (int m = 0; m < kM; ++m)
    if (bits >> m & 1u) {
      st += T[m];
      sm += mm[m] * T[m];
    }
  *pmu = sm / st;
  *var = 1.f / st;
  *plv = logf(*var);
  if (rr) {
#pragma unroll
    for (int m = 0; m < kM; ++m) rr[m] = T[m] * *var;
  }
}

__device__ __forceinline__ bool use_prior(uint32_t bits, int M, int prior_mode) {
  if (prior_mode == MV_PRIOR_ALWAYS_STABLE) return true;
  if (prior_mode == MV_PRIOR_FULL_SUBSET) return __popc(bits) == M;
  return false;
}

__device__ __forceinline__ float block_sum(float v, float* sh) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[wid] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < (blockDim.x >> 5); ++i) t += sh[i];
  return t;
}

__global__ void __launch_bounds__(128) poe_fwd_kernel(
    const float* __restrict__ mu, const float* __restrict__ lv, const uint8_t* __restrict__ masks,
    const uint32_t* __restrict__ subsets, int S, const int32_t* __restrict__ sel, const float* __restrict__ w,
    float w_uniform, const float* __restrict__ noise, int prior_mode, int stable, float eps, float* __restrict__ z,
    float* __restrict__ jmu_out, float* __restrict__ jlv_out, float* __restrict__ kl_b, float* __restrict__ kldm_b,
    int M, int B, int L) {
  __shared__ float sh[4];
  const int b = blockIdx.x;
  const int my_sel = sel ? sel[b] : 0;
  float kl_acc = 0.f;
  float km_acc[kMaxM];
#pragma unroll
  for (int m = 0; m < kMaxM; ++m) km_acc[m] = 0.f;
  bool avail[kMaxM];
#pragma unroll
  for (int m = 0; m < kMaxM; ++m) avail[m] = m < M && (masks == nullptr || masks[m * B + b] != 0);

  for (int l = threadIdx.x; l < L; l += blockDim.x) {
    float T[kMaxM], mm[kMaxM], lvv[kMaxM];
#pragma unroll
    for (int m = 0; m < kMaxM; ++m) {
      if (m < M) {
        const int64_t i = (int64_t(m) * B + b) * L + l;
        mm[m] = mu[i];
        lvv[m] = lv[i];
        T[m] = avail[m] ? expert_T(lvv[m], stable, eps) : 0.f;
      } else {
        T[m] = 0.f; mm[m] = 0.f; lvv[m] = 0.f;
      }
    }
    const float Tp = stable ? 1.f : 1.f / (1.f + eps);  // prior expert N(0,I)
    for (int s = 0; s < S; ++s) {
      const uint32_t bits = subsets[s];
      float pmu, var, plv;
      subset_poe(bits, use_prior(bits, M, prior_mode), stable, Tp, T, mm, lvv, avail, &pmu, &var, &plv, nullptr);
      const float ws = w ? w[int64_t(s) * B + b] : w_uniform;
      kl_acc += ws * (-0.5f * (1.f + plv - pmu * pmu - var));
      if (s == my_sel) {
        const int64_t o = int64_t(b) * L + l;
        if (z) z[o] = pmu + sqrtf(var) * noise[o];
        if (jmu_out) jmu_out[o] = pmu;
        if (jlv_out) jlv_out[o] = plv;
        if (kldm_b) {
#pragma unroll
          for (int m = 0; m < kMaxM; ++m)
            if (m < M && avail[m]) {
              const float iv = expf(-lvv[m]);
              const float d = pmu - mm[m];
              km_acc[m] += -0.5f * (1.f - var * iv - d * d * iv + plv - lvv[m]);
            }
        }
      }
    }
  }
  const float kl = block_sum(kl_acc, sh);
  if (threadIdx.x == 0) kl_b[b] = kl;
  if (kldm_b) {
#pragma unroll
    for (int m = 0; m < kMaxM; ++m)
      if (m < M) {
        const float v = block_sum(km_acc[m], sh);
        if (threadIdx.x == 0) kldm_b[m * B + b] = v;
      }
  }
}

__global__ void __launch_bounds__(128) poe_bwd_kernel(
    const float* __restrict__ mu, const float* __restrict__ lv, const uint8_t* __restrict__ masks,
    const uint32_t* __restrict__ subsets, int S, const int32_t* __restrict__ sel, const float* __restrict__ w,
    float w_uniform, const float* __restrict__ noise, int prior_mode, int stable, float eps,
    const float* __restrict__ g_z, const float* __restrict__ g_kl, const float* __restrict__ g_kldm,
    float* __restrict__ g_mu, float* __restrict__ g_lv, int M, int B, int L) {
  const int b = blockIdx.x;
  const int my_sel = sel ? sel[b] : 0;
  const float gk = g_kl ? g_kl[b] : 0.f;
  bool avail[kMaxM];
  float gkm[kMaxM];
#pragma unroll
  for (int m = 0; m < kMaxM; ++m) {
    avail[m] = m < M && (masks == nullptr || masks[m * B + b] != 0);
    gkm[m] = (g_kldm && m < M) ? g_kldm[m * B + b] : 0.f;
  }
  for (int l = threadIdx.x; l < L; l += blockDim.x) {
    float T[kMaxM], mm[kMaxM], lvv[kMaxM], gm[kMaxM], gT[kMaxM], glv_direct[kMaxM];
#pragma unroll
    for (int m = 0; m < kMaxM; ++m) {
      gm[m] = 0.f; gT[m] = 0.f; glv_direct[m] = 0.f;
      if (m < M) {
        const int64_t i = (int64_t(m) * B + b) * L + l;
        mm[m] = mu[i];
        lvv[m] = lv[i];
        T[m] = avail[m] ? expert_T(lvv[m], stable, eps) : 0.f;
      } else {
        T[m] = 0.f; mm[m] = 0.f; lvv[m] = 0.f;
      }
    }
    const float Tp = stable ? 1.f : 1.f / (1.f + eps);
    const int64_t o = int64_t(b) * L + l;
    for (int s = 0; s < S; ++s) {
      const uint32_t bits = subsets[s];
      float pmu, var, plv, rr[kMaxM];   // rr[m] = T_m / sum of precisions (responsibility of expert m in this subset)
      subset_poe(bits, use_prior(bits, M, prior_mode), stable, Tp, T, mm, lvv, avail, &pmu, &var, &plv, rr);
      const float st = 1.f / var;
      const float a = (w ? w[int64_t(s) * B + b] : w_uniform) * gk;
      float gmu = a * pmu;                       // d KL / d mu
      float gvar = a * 0.5f * (1.f - st);        // d KL / d var = 0.5 * (1 - 1/var)
      if (s == my_sel) {
        if (g_z) {
          const float gz = g_z[o];
          gmu += gz;
          gvar += gz * noise[o] * 0.5f * rsqrtf(var);
        }
        if (g_kldm) {
#pragma unroll
          for (int m = 0; m < kMaxM; ++m)
            if (m < M && avail[m]) {
              const float iv = expf(-lvv[m]);
              const float d = pmu - mm[m];
              gmu += gkm[m] * d * iv;
              gvar += gkm[m] * 0.5f * (iv - st);
              gm[m] += -gkm[m] * d * iv;
              glv_direct[m] += gkm[m] * 0.5f * (1.f - var * iv - d * d * iv);
            }
        }
      }
#pragma unroll
      for (int m = 0; m < kMaxM; ++m)
        if (bits >> m & 1u) {
          gm[m] += gmu * rr[m];
          if (stable) {
            // d/dlv through T = exp(-lv): dT/dlv = -T, written with the responsibilities (no raw exp(-lv) anywhere)
            glv_direct[m] += -(gmu * (mm[m] - pmu) - gvar * var) * rr[m];
          } else {
            gT[m] += gmu * (mm[m] - pmu) * var - gvar * var * var;
          }
        }
    }
#pragma unroll
    for (int m = 0; m < kMaxM; ++m)
      if (m < M) {
        const int64_t i = (int64_t(m) * B + b) * L + l;
        // dT/dlv: stable: -T ; eps form: -exp(lv) * T^2
        const float dT = stable ? 0.f : -expf(lvv[m]) * T[m] * T[m];
        g_mu[i] = avail[m] ? gm[m] : 0.f;
        g_lv[i] = avail[m] ? gT[m] * dT + glv_direct[m] : 0.f;
      }
  }
}

}  // namespace mv

using namespace mv;

extern "C" int mv_poe_fwd(const float* mu, const float* lv, const uint8_t* masks, const uint32_t* subsets, int S,
                          const int32_t* sel, const float* w, float w_uniform, const float* noise, int prior_mode,
                          int stable, float eps, float* z, float* joint_mu, float* joint_lv, float* kl_b,
                          float* kldm_b, int M, int B, int L, void* stream) {
  MV_CHECK_ARG(mu && lv && subsets && kl_b, "mv_poe_fwd: null pointer");
  MV_CHECK_ARG(z == nullptr || noise != nullptr, "mv_poe_fwd: z requested without noise");
  MV_CHECK_ARG(M > 0 && M <= kMaxM, "mv_poe_fwd: 1 <= n_modalities <= %d required, got %d", kMaxM, M);
  MV_CHECK_ARG(S > 0 && B > 0 && L > 0, "mv_poe_fwd: bad sizes");
  MV_CHECK_ARG(prior_mode >= MV_PRIOR_NEVER && prior_mode <= MV_PRIOR_FULL_SUBSET, "mv_poe_fwd: bad prior mode");
  poe_fwd_kernel<<<B, 128, 0, static_cast<cudaStream_t>(stream)>>>(mu, lv, masks, subsets, S, sel, w, w_uniform, noise,
                                                                  prior_mode, stable, eps, z, joint_mu, joint_lv, kl_b,
                                                                  kldm_b, M, B, L);
  MV_CHECK_LAUNCH("mv_poe_fwd");
  return MV_OK;
}

extern "C" int mv_poe_bwd(const float* mu, const float* lv, const uint8_t* masks, const uint32_t* subsets, int S,
                          const int32_t* sel, const float* w, float w_uniform, const float* noise, int prior_mode,
                          int stable, float eps, const float* g_z, const float* g_kl, const float* g_kldm, float* g_mu,
                          float* g_lv, int M, int B, int L, void* stream) {
  MV_CHECK_ARG(mu && lv && subsets && g_mu && g_lv, "mv_poe_bwd: null pointer");
  MV_CHECK_ARG(g_z == nullptr || noise != nullptr, "mv_poe_bwd: g_z given without noise");
  MV_CHECK_ARG(M > 0 && M <= kMaxM, "mv_poe_bwd: 1 <= n_modalities <= %d required, got %d", kMaxM, M);
  MV_CHECK_ARG(S > 0 && B > 0 && L > 0, "mv_poe_bwd: bad sizes");
  poe_bwd_kernel<<<B, 128, 0, static_cast<cudaStream_t>(stream)>>>(mu, lv, masks, subsets, S, sel, w, w_uniform, noise,
                                                                  prior_mode, stable, eps, g_z, g_kl, g_kldm, g_mu, g_lv,
                                                                  M, B, L);
  MV_CHECK_LAUNCH("mv_poe_bwd");
  return MV_OK;
}
