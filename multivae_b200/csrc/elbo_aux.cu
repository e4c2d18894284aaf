// Smaller pieces of the ELBO / likelihood path (HBM-bound or latency-bound elementwise + reduction kernels):
//
//   mv_moe_lpx_cat_fwd/bwd  categorical decoders: lpx[c,k,b] = rescale * sum_p sum_v x[b,p,v] * log_softmax(recon[c,k,b,p,:] + 1e-6)[v]
//                           (models/base/base_utils.py:28-59 cross_entropy_, :81-87)
//   mv_logmeanexp           out[b] = logsumexp_r lw[r,b] - log R   (importance-sampled likelihoods, K = 1000:
//                           models/base/base_ae_model.py:396-442, mmvaePlus_model.py:478-533, mvae_model.py:241-317 ...)
//   mv_gauss_kl_fwd/bwd     KL(N(mu, e^lv) || N(pm, e^plv)) summed over the last dimension (base_utils.py:90-119)
#include "common.cuh"

namespace mv {

// one warp per (row, position): logsumexp over the V classes, then the dot product with the target probabilities
template <typename T>
__global__ void __launch_bounds__(128) lpx_cat_fwd_kernel(const T* __restrict__ recon, const float* __restrict__ x,
                                                          float* __restrict__ lpx, int64_t rows, int B, int P, int V, float eps,
                                                          float rescale, const uint8_t* __restrict__ mask) {
  const int lane = threadIdx.x & 31;
  const int64_t wi = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (wi >= rows * P) return;
  const int64_t row = wi / P;
  const int p = int(wi - row * P);
  const int b = int(row % B);
  if (mask && !mask[b]) return;
  const T* r = recon + (row * P + p) * V;
  const float* xt = x + (int64_t(b) * P + p) * V;
  float mx = -INFINITY;
  for (int v = lane; v < V; v += 32) mx = fmaxf(mx, Vec<T>::load1(r + v) + eps);
  mx = warp_max(mx);
  float se = 0.f, dot = 0.f, xs = 0.f;
  for (int v = lane; v < V; v += 32) {
    const float l = Vec<T>::load1(r + v) + eps;
    se += expf(l - mx);
    const float t = xt[v];
    dot += t * l;
    xs += t;
  }
  se = warp_sum(se);
  dot = warp_sum(dot);
  xs = warp_sum(xs);
  if (lane == 0) atomicAdd(lpx + row, rescale * (dot - xs * (mx + logf(se))));
}

template <typename T>
__global__ void __launch_bounds__(128) lpx_cat_bwd_kernel(const T* __restrict__ recon, const float* __restrict__ x,
                                                          const float* __restrict__ coef, const float* __restrict__ g_loss,
                                                          T* __restrict__ g_recon, int64_t rows, int B, int P, int V, float eps,
                                                          float rescale, const uint8_t* __restrict__ mask) {
  const int lane = threadIdx.x & 31;
  const int64_t wi = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (wi >= rows * P) return;
  const int64_t row = wi / P;
  const int p = int(wi - row * P);
  const int b = int(row % B);
  const T* r = recon + (row * P + p) * V;
  T* g = g_recon + (row * P + p) * V;
  if (mask && !mask[b]) {
    for (int v = lane; v < V; v += 32) Vec<T>::store1(g + v, 0.f);
    return;
  }
  const float* xt = x + (int64_t(b) * P + p) * V;
  float mx = -INFINITY;
  for (int v = lane; v < V; v += 32) mx = fmaxf(mx, Vec<T>::load1(r + v) + eps);
  mx = warp_max(mx);
  float se = 0.f, xs = 0.f;
  for (int v = lane; v < V; v += 32) {
    se += expf(Vec<T>::load1(r + v) + eps - mx);
    xs += xt[v];
  }
  se = warp_sum(se);
  xs = warp_sum(xs);
  const float cf = coef[row] * *g_loss * rescale;
  const float inv = 1.f / se;
  for (int v = lane; v < V; v += 32) {
    const float sm = expf(Vec<T>::load1(r + v) + eps - mx) * inv;
    Vec<T>::store1(g + v, cf * (xt[v] - xs * sm));
  }
}

// one warp per column b
__global__ void __launch_bounds__(128) logmeanexp_kernel(const float* __restrict__ lw, int R, int B, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= B) return;
  float mx = -INFINITY;
  for (int r = lane; r < R; r += 32) mx = fmaxf(mx, lw[int64_t(r) * B + b]);
  mx = warp_max(mx);
  float se = 0.f;
  if (mx > -INFINITY)
    for (int r = lane; r < R; r += 32) se += expf(lw[int64_t(r) * B + b] - mx);
  se = warp_sum(se);
  if (lane == 0) out[b] = (mx > -INFINITY ? mx + logf(se) : -INFINITY) - logf(float(R));
}

// one warp per row; prior parameters either per row (p_rows == rows) or one broadcast row (p_rows == 1)
__global__ void __launch_bounds__(128) gauss_kl_fwd_kernel(const float* __restrict__ mu, const float* __restrict__ lv,
                                                           const float* __restrict__ pm, const float* __restrict__ plv,
                                                           float* __restrict__ out, int64_t rows, int L, int p_bcast) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  const float* pmr = pm + (p_bcast ? 0 : row * L);
  const float* plr = plv + (p_bcast ? 0 : row * L);
  float acc = 0.f;
  for (int l = lane; l < L; l += 32) {
    const float m = mu[row * L + l], v = lv[row * L + l], d = m - pmr[l];
    acc += 0.5f * (plr[l] - v + expf(v - plr[l]) + d * d / expf(plr[l]) - 1.f);
  }
  acc = warp_sum(acc);
  if (lane == 0) out[row] = acc;
}

__global__ void __launch_bounds__(128) gauss_kl_bwd_kernel(const float* __restrict__ mu, const float* __restrict__ lv,
                                                           const float* __restrict__ pm, const float* __restrict__ plv,
                                                           const float* __restrict__ g, float* __restrict__ g_mu,
                                                           float* __restrict__ g_lv, float* __restrict__ g_pm,
                                                           float* __restrict__ g_plv, int64_t rows, int L, int p_bcast) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * L) return;
  const int64_t row = i / L;
  const int l = int(i - row * L);
  const int64_t pi = p_bcast ? l : i;
  const float m = mu[i], v = lv[i], pmv = pm[pi], pv = plv[pi], d = m - pmv, go = g[row];
  const float ipv = expf(-pv), ratio = expf(v - pv);
  g_mu[i] = go * d * ipv;
  g_lv[i] = go * 0.5f * (ratio - 1.f);
  const float gpm = -go * d * ipv;
  const float gpl = go * 0.5f * (1.f - ratio - d * d * ipv);
  if (p_bcast) {   // broadcast prior row: gradients of all rows add up (buffers zero-filled by the host wrapper)
    if (g_pm) atomicAdd(g_pm + l, gpm);
    if (g_plv) atomicAdd(g_plv + l, gpl);
  } else {
    if (g_pm) g_pm[i] = gpm;
    if (g_plv) g_plv[i] = gpl;
  }
}

}  // namespace mv

using namespace mv;

extern "C" int mv_moe_lpx_cat_fwd(const void* recon, int recon_dtype, const float* x, float* lpx, int C, int K, int B, int P, int V,
                                  float rescale, const uint8_t* mask_r, int accumulate, void* stream) {
  MV_CHECK_ARG(recon && x && lpx, "mv_moe_lpx_cat_fwd: null pointer");
  MV_CHECK_ARG(C > 0 && K > 0 && B > 0 && P > 0 && V > 0, "mv_moe_lpx_cat_fwd: bad sizes");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t rows = int64_t(C) * K * B;
  if (!accumulate) cudaMemsetAsync(lpx, 0, sizeof(float) * size_t(rows), st);
  const int64_t warps = rows * P;
  const unsigned blocks = unsigned((warps + 3) / 4);
  if (recon_dtype == MV_F32)
    lpx_cat_fwd_kernel<float><<<blocks, 128, 0, st>>>(static_cast<const float*>(recon), x, lpx, rows, B, P, V, 1e-6f, rescale, mask_r);
  else if (recon_dtype == MV_BF16)
    lpx_cat_fwd_kernel<__nv_bfloat16><<<blocks, 128, 0, st>>>(static_cast<const __nv_bfloat16*>(recon), x, lpx, rows, B, P, V, 1e-6f,
                                                              rescale, mask_r);
  else {
    mv::set_error("mv_moe_lpx_cat_fwd: unsupported dtype %d", recon_dtype);
    return MV_ERR_UNSUPPORTED;
  }
  MV_CHECK_LAUNCH("mv_moe_lpx_cat_fwd");
  return MV_OK;
}

extern "C" int mv_moe_lpx_cat_bwd(const void* recon, int recon_dtype, const float* x, const float* coef, const float* g_loss,
                                  void* g_recon, int C, int K, int B, int P, int V, float rescale, const uint8_t* mask_r,
                                  void* stream) {
  MV_CHECK_ARG(recon && x && coef && g_loss && g_recon, "mv_moe_lpx_cat_bwd: null pointer");
  MV_CHECK_ARG(C > 0 && K > 0 && B > 0 && P > 0 && V > 0, "mv_moe_lpx_cat_bwd: bad sizes");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t rows = int64_t(C) * K * B;
  const unsigned blocks = unsigned((rows * P + 3) / 4);
  if (recon_dtype == MV_F32)
    lpx_cat_bwd_kernel<float><<<blocks, 128, 0, st>>>(static_cast<const float*>(recon), x, coef, g_loss, static_cast<float*>(g_recon),
                                                      rows, B, P, V, 1e-6f, rescale, mask_r);
  else if (recon_dtype == MV_BF16)
    lpx_cat_bwd_kernel<__nv_bfloat16><<<blocks, 128, 0, st>>>(static_cast<const __nv_bfloat16*>(recon), x, coef, g_loss,
                                                              static_cast<__nv_bfloat16*>(g_recon), rows, B, P, V, 1e-6f, rescale, mask_r);
  else {
    mv::set_error("mv_moe_lpx_cat_bwd: unsupported dtype %d", recon_dtype);
    return MV_ERR_UNSUPPORTED;
  }
  MV_CHECK_LAUNCH("mv_moe_lpx_cat_bwd");
  return MV_OK;
}

extern "C" int mv_logmeanexp(const float* lw, int R, int B, float* out, void* stream) {
  MV_CHECK_ARG(lw && out && R > 0 && B > 0, "mv_logmeanexp: bad arguments");
  logmeanexp_kernel<<<(B + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(lw, R, B, out);
  MV_CHECK_LAUNCH("mv_logmeanexp");
  return MV_OK;
}

extern "C" int mv_gauss_kl_fwd(const float* mu, const float* lv, const float* prior_mu, const float* prior_lv, float* out,
                               int64_t rows, int L, int prior_rows, void* stream) {
  MV_CHECK_ARG(mu && lv && prior_mu && prior_lv && out && rows > 0 && L > 0, "mv_gauss_kl_fwd: bad arguments");
  MV_CHECK_ARG(prior_rows == 1 || prior_rows == rows, "mv_gauss_kl_fwd: prior must have 1 row or one per sample");
  gauss_kl_fwd_kernel<<<unsigned((rows + 3) / 4), 128, 0, static_cast<cudaStream_t>(stream)>>>(mu, lv, prior_mu, prior_lv, out, rows, L,
                                                                                              prior_rows == 1 && rows != 1);
  MV_CHECK_LAUNCH("mv_gauss_kl_fwd");
  return MV_OK;
}

extern "C" int mv_gauss_kl_bwd(const float* mu, const float* lv, const float* prior_mu, const float* prior_lv, const float* g_out,
                               float* g_mu, float* g_lv, float* g_prior_mu, float* g_prior_lv, int64_t rows, int L, int prior_rows,
                               void* stream) {
  MV_CHECK_ARG(mu && lv && prior_mu && prior_lv && g_out && g_mu && g_lv && rows > 0 && L > 0, "mv_gauss_kl_bwd: bad arguments");
  MV_CHECK_ARG(prior_rows == 1 || prior_rows == rows, "mv_gauss_kl_bwd: prior must have 1 row or one per sample");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int bc = prior_rows == 1 && rows != 1;
  if (bc) {
    if (g_prior_mu) cudaMemsetAsync(g_prior_mu, 0, sizeof(float) * L, st);
    if (g_prior_lv) cudaMemsetAsync(g_prior_lv, 0, sizeof(float) * L, st);
  }
  const int64_t n = rows * L;
  gauss_kl_bwd_kernel<<<unsigned((n + 127) / 128), 128, 0, st>>>(mu, lv, prior_mu, prior_lv, g_out, g_mu, g_lv, g_prior_mu, g_prior_lv,
                                                                 rows, L, bc);
  MV_CHECK_LAUNCH("mv_gauss_kl_bwd");
  return MV_OK;
}
