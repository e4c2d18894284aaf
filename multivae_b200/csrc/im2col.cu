// Strided / transposed convolutions of the small convolutional networks (models/nn/svhn.py:7-70, mmnist.py:78-110,173-207:
// 4x4 and 3x3 kernels, stride 2, 3..128 channels, < 10 MFLOP per sample) as "gather + tensor-core GEMM":
//
//   mv_im2col   cols[(n, gy, gx), c*T + t] = src[n, gy*s - pad + ky, gx*s - pad + kx, c]   (0 outside the source)
//               the patch matrix of a strided convolution's forward pass, and of a transposed convolution's weight / data
//               gradient (there `src` is the gradient of the transposed convolution's output and (gy, gx) its input grid)
//   mv_col2im   dst[n, y, x, c] = act(bias[c] + sum_{t : (y + pad - ky) % s == 0, ...} cols[(n, (y+pad-ky)/s, (x+pad-kx)/s), c*T + t])
//               the gather form of the scatter-add: a transposed convolution's forward pass (after the GEMM x W that produces
//               `cols`) and a strided convolution's data gradient; bias, ReLU / Sigmoid and the activation-derivative mask of
//               the layer below are fused
//
// The contractions themselves run on the general tcgen05 GEMM (csrc/gemm.cu).  Columns are ordered (channel, tap) like a torch
// Conv2d weight [N, C, kh, kw] flattened, so packed weights and weight gradients need no permutation.  HBM-bound helpers:
// one thread per element, coalesced along the innermost index of whichever side is written.
#include <algorithm>

#include "common.cuh"

namespace mv {

using bf16 = __nv_bfloat16;

struct ConvGeom {
  int n_img, H, W, C;          // the image-side tensor (im2col: source; col2im: destination)
  int kh, kw, stride, pad;
  int Gh, Gw;                  // the grid the patch matrix has one row per position of
  int64_t sn, sy, sx, sc;      // element strides of the image-side tensor (NHWC or NCHW)
  int ld;                      // row pitch of the patch matrix (>= C*kh*kw, multiple of 8; extra columns are written as zeros)
  int tc;                      // column order of the patch matrix: 1 = (tap, channel) [channels contiguous], 0 = (channel, tap)
};

template <typename T>
__global__ void __launch_bounds__(256) im2col_kernel(const T* __restrict__ src, bf16* __restrict__ cols, const ConvGeom g) {
  const int T_ = g.kh * g.kw;
  const int64_t total = int64_t(g.n_img) * g.Gh * g.Gw * g.ld;
  for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += int64_t(gridDim.x) * blockDim.x) {
    const int col = int(e % g.ld);
    const int64_t row = e / g.ld;
    float v = 0.f;
    if (col < g.C * T_) {
      const int c = g.tc ? col % g.C : col / T_, t = g.tc ? col / g.C : col - (col / T_) * T_;
      const int ky = t / g.kw, kx = t - ky * g.kw;
      const int gx = int(row % g.Gw);
      const int64_t r2 = row / g.Gw;
      const int gy = int(r2 % g.Gh), n = int(r2 / g.Gh);
      const int y = gy * g.stride - g.pad + ky, x = gx * g.stride - g.pad + kx;
      if (y >= 0 && y < g.H && x >= 0 && x < g.W) v = Vec<T>::load1(src + n * g.sn + y * g.sy + x * g.sx + c * g.sc);
    }
    cols[e] = __float2bfloat16_rn(v);
  }
}

template <typename TC>
__global__ void __launch_bounds__(256) col2im_kernel(const TC* __restrict__ cols, bf16* __restrict__ dst, const ConvGeom g,
                                                     const float* __restrict__ bias, int act, const bf16* __restrict__ dact,
                                                     float dslope, int c_fastest) {
  const int T_ = g.kh * g.kw;
  const int64_t total = int64_t(g.n_img) * g.H * g.W * g.C;
  for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += int64_t(gridDim.x) * blockDim.x) {
    int n, y, x, c;
    if (c_fastest) {   // NHWC destination
      c = int(e % g.C);
      int64_t r = e / g.C;
      x = int(r % g.W); r /= g.W;
      y = int(r % g.H); n = int(r / g.H);
    } else {           // NCHW destination
      x = int(e % g.W);
      int64_t r = e / g.W;
      y = int(r % g.H); r /= g.H;
      c = int(r % g.C); n = int(r / g.C);
    }
    float acc = bias ? bias[c] : 0.f;
    for (int ky = 0; ky < g.kh; ++ky) {
      const int yy = y + g.pad - ky;
      if (yy < 0 || yy % g.stride) continue;
      const int gy = yy / g.stride;
      if (gy >= g.Gh) continue;
      for (int kx = 0; kx < g.kw; ++kx) {
        const int xx = x + g.pad - kx;
        if (xx < 0 || xx % g.stride) continue;
        const int gx = xx / g.stride;
        if (gx >= g.Gw) continue;
        const int t = ky * g.kw + kx;
        acc += Vec<TC>::load1(cols + ((int64_t(n) * g.Gh + gy) * g.Gw + gx) * g.ld + (g.tc ? t * g.C + c : c * T_ + t));
      }
    }
    if (act == MV_ACT_RELU) acc = fmaxf(acc, 0.f);
    else if (act == MV_ACT_LRELU02) acc = fmaxf(acc, 0.2f * acc);
    else if (act == MV_ACT_SIGMOID) acc = 1.f / (1.f + expf(-acc));
    const int64_t o = n * g.sn + y * g.sy + x * g.sx + c * g.sc;
    if (dact) acc *= __bfloat162float(dact[o]) > 0.f ? 1.f : dslope;
    dst[o] = __float2bfloat16_rn(acc);
  }
}

// Vector paths for NHWC tensors with C % 8 == 0 and (tap, channel) column order: a thread moves 8 consecutive channels
// (16-byte loads / stores on both sides, 32-bit index arithmetic).
__global__ void __launch_bounds__(256) im2col_vec_kernel(const bf16* __restrict__ src, bf16* __restrict__ cols, const ConvGeom g) {
  const int C8 = g.C >> 3, T_ = g.kh * g.kw;
  const int per_row = T_ * C8;
  const int64_t rows = int64_t(g.n_img) * g.Gh * g.Gw;
  const int64_t total = rows * per_row;
  for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += int64_t(gridDim.x) * blockDim.x) {
    const int64_t row = e / per_row;
    const int r = int(e - row * per_row);
    const int t = r / C8, c8 = r - t * C8;
    const int ky = t / g.kw, kx = t - ky * g.kw;
    const int gx = int(row % g.Gw);
    const int r2 = int(row / g.Gw);
    const int gy = r2 % g.Gh, n = r2 / g.Gh;
    const int y = gy * g.stride - g.pad + ky, x = gx * g.stride - g.pad + kx;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (y >= 0 && y < g.H && x >= 0 && x < g.W) v = *reinterpret_cast<const uint4*>(src + ((int64_t(n) * g.H + y) * g.W + x) * g.C + c8 * 8);
    *reinterpret_cast<uint4*>(cols + row * g.ld + t * g.C + c8 * 8) = v;
  }
}

template <typename TC>
__global__ void __launch_bounds__(256) col2im_vec_kernel(const TC* __restrict__ cols, bf16* __restrict__ dst, const ConvGeom g,
                                                         const float* __restrict__ bias, int act, const bf16* __restrict__ dact,
                                                         float dslope) {
  const int C8 = g.C >> 3;
  const int64_t total = int64_t(g.n_img) * g.H * g.W * C8;
  for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += int64_t(gridDim.x) * blockDim.x) {
    const int c8 = int(e % C8);
    const int pix = int(e / C8);
    const int x = pix % g.W;
    const int r = pix / g.W;
    const int y = r % g.H, n = r / g.H;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = bias ? bias[c8 * 8 + j] : 0.f;
    for (int ky = 0; ky < g.kh; ++ky) {
      const int yy = y + g.pad - ky;
      if (yy < 0 || yy % g.stride) continue;
      const int gy = yy / g.stride;
      if (gy >= g.Gh) continue;
      for (int kx = 0; kx < g.kw; ++kx) {
        const int xx = x + g.pad - kx;
        if (xx < 0 || xx % g.stride) continue;
        const int gx = xx / g.stride;
        if (gx >= g.Gw) continue;
        const TC* p = cols + ((int64_t(n) * g.Gh + gy) * g.Gw + gx) * g.ld + (ky * g.kw + kx) * g.C + c8 * 8;
        float f[8];
        if (sizeof(TC) == 4) {
          Vec<float>::unpack(*reinterpret_cast<const uint4*>(p), f);
          Vec<float>::unpack(*reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(p) + 4), f + 4);
        } else {
          Vec<bf16>::unpack(*reinterpret_cast<const uint4*>(p), f);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += f[j];
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (act == MV_ACT_RELU) acc[j] = fmaxf(acc[j], 0.f);
      else if (act == MV_ACT_LRELU02) acc[j] = fmaxf(acc[j], 0.2f * acc[j]);
      else if (act == MV_ACT_SIGMOID) acc[j] = 1.f / (1.f + expf(-acc[j]));
    }
    const int64_t o = int64_t(pix) * g.C + c8 * 8;
    if (dact) {
      float d[8];
      Vec<bf16>::unpack(*reinterpret_cast<const uint4*>(dact + o), d);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] *= d[j] > 0.f ? 1.f : dslope;
    }
    *reinterpret_cast<uint4*>(dst + o) = Vec<bf16>::pack(acc);
  }
}

// Few-channel tensors (the 3-channel images at either end of the networks, NCHW or NHWC), tap-major columns: one thread per
// (patch row, tap) resp. per pixel, looping over the channels, so the index arithmetic (32-bit) is paid once per pixel and the
// tap loops only visit the taps that can contribute (ky = (y + pad) mod stride, +stride, ...).
template <typename T>
__global__ void __launch_bounds__(256) im2col_pix_kernel(const T* __restrict__ src, bf16* __restrict__ cols, const ConvGeom g) {
  const int T_ = g.kh * g.kw;
  const int64_t total = int64_t(g.n_img) * g.Gh * g.Gw * T_;
  for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += int64_t(gridDim.x) * blockDim.x) {
    const int t = int(e % T_);
    const int row = int(e / T_);
    const int ky = t / g.kw, kx = t - ky * g.kw;
    const int gx = row % g.Gw;
    const int r2 = row / g.Gw;
    const int gy = r2 % g.Gh, n = r2 / g.Gh;
    const int y = gy * g.stride - g.pad + ky, x = gx * g.stride - g.pad + kx;
    const bool in = y >= 0 && y < g.H && x >= 0 && x < g.W;
    const T* sp = src + n * g.sn + y * g.sy + x * g.sx;
    bf16* cp = cols + int64_t(row) * g.ld + t * g.C;
    for (int c = 0; c < g.C; ++c) cp[c] = __float2bfloat16_rn(in ? Vec<T>::load1(sp + c * g.sc) : 0.f);
    if (t == T_ - 1)
      for (int c = g.C * T_; c < g.ld; ++c) cols[int64_t(row) * g.ld + c] = __float2bfloat16_rn(0.f);
  }
}

template <typename TC>
__global__ void __launch_bounds__(256) col2im_pix_kernel(const TC* __restrict__ cols, bf16* __restrict__ dst, const ConvGeom g,
                                                         const float* __restrict__ bias, int act, const bf16* __restrict__ dact,
                                                         float dslope) {
  const int total = g.n_img * g.H * g.W;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int x = e % g.W;
    const int r = e / g.W;
    const int y = r % g.H, n = r / g.H;
    float acc[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[c] = (bias && c < g.C) ? bias[c] : 0.f;
    for (int ky = (y + g.pad) % g.stride; ky < g.kh; ky += g.stride) {
      const int gy = (y + g.pad - ky) / g.stride;
      if (y + g.pad - ky < 0) break;
      if (gy >= g.Gh) continue;
      for (int kx = (x + g.pad) % g.stride; kx < g.kw; kx += g.stride) {
        const int gx = (x + g.pad - kx) / g.stride;
        if (x + g.pad - kx < 0) break;
        if (gx >= g.Gw) continue;
        const TC* p = cols + (int64_t(n * g.Gh + gy) * g.Gw + gx) * g.ld + (ky * g.kw + kx) * g.C;
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (c < g.C) acc[c] += Vec<TC>::load1(p + c);
      }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (c >= g.C) break;
      float v = acc[c];
      if (act == MV_ACT_RELU) v = fmaxf(v, 0.f);
      else if (act == MV_ACT_LRELU02) v = fmaxf(v, 0.2f * v);
      else if (act == MV_ACT_SIGMOID) v = 1.f / (1.f + expf(-v));
      const int64_t o = n * g.sn + y * g.sy + x * g.sx + c * g.sc;
      if (dact) v *= __bfloat162float(dact[o]) > 0.f ? 1.f : dslope;
      dst[o] = __float2bfloat16_rn(v);
    }
  }
}

// Weight hand-over for the (tap, channel) column order, all layers of a network in one launch:
//   pack:    dst bf16 [N][ld], dst[n, t*C + c] = src fp32 [N][C][T]   (columns >= T*C zero)
//   unpack:  dst fp32 [N][C][T] += src fp32 [N][ld][t*C + c]
struct TcBatch {
  mv_pack_item it[MV_PACK_MAX_ITEMS];
};
__global__ void __launch_bounds__(256) pack_tc_kernel(const __grid_constant__ TcBatch b) {
  const mv_pack_item& w = b.it[blockIdx.y];
  const float* __restrict__ src = static_cast<const float*>(w.src);
  bf16* __restrict__ dst = static_cast<bf16*>(w.dst_fwd);
  const int ld = w.Cpad, TC_ = w.T * w.C;
  const int64_t total = int64_t(w.N) * ld;
  if (w.T == 1 && ld == w.C && (w.C & 7) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    // Linear weights: a plain cast, 8 elements per thread (two 16-byte loads, one 16-byte store)
    const int64_t nv = total >> 3;
    for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < nv; e += int64_t(gridDim.x) * blockDim.x) {
      const float4 a = *reinterpret_cast<const float4*>(src + e * 8), c = *reinterpret_cast<const float4*>(src + e * 8 + 4);
      *reinterpret_cast<uint4*>(dst + e * 8) = make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(c.x, c.y), pack_bf16(c.z, c.w));
    }
    return;
  }
  for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += int64_t(gridDim.x) * blockDim.x) {
    const int col = int(e % ld);
    const int64_t n = e / ld;
    float v = 0.f;
    if (col < TC_) {
      const int t = col / w.C, c = col - t * w.C;
      v = src[(n * w.C + c) * w.T + t];
    }
    dst[e] = __float2bfloat16_rn(v);
  }
}
__global__ void __launch_bounds__(256) unpack_tc_add_kernel(const __grid_constant__ TcBatch b) {
  const mv_pack_item& w = b.it[blockIdx.y];
  const float* __restrict__ src = static_cast<const float*>(w.src);
  float* __restrict__ dst = static_cast<float*>(w.dst_fwd);
  const int ld = w.Cpad;
  const int64_t total = int64_t(w.N) * w.C * w.T;
  for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += int64_t(gridDim.x) * blockDim.x) {
    const int t = int(e % w.T);
    const int64_t nc = e / w.T;
    const int c = int(nc % w.C);
    const int64_t n = nc / w.C;
    dst[e] += src[n * ld + t * w.C + c];
  }
}

// out[c] += sum over (n, y, x) of g[n, c, y, x] (bias gradient of a layer whose output is NCHW with few channels)
__global__ void __launch_bounds__(256) chan_sum_nchw_kernel(const bf16* __restrict__ g, int n_img, int C, int HW, float* __restrict__ out) {
  __shared__ float red[8];
  const int c = blockIdx.x;
  float acc = 0.f;
  for (int n = blockIdx.y; n < n_img; n += gridDim.y)
    for (int i = threadIdx.x; i < HW; i += blockDim.x) acc += __bfloat162float(g[(int64_t(n) * C + c) * HW + i]);
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[i];
    atomicAdd(out + c, s);
  }
}

int num_sms();

static int fill_geom(ConvGeom* g, const mv_conv_geom* a, const char* who) {
  MV_CHECK_ARG(a && a->n_img > 0 && a->H > 0 && a->W > 0 && a->C > 0 && a->kh > 0 && a->kw > 0 && a->stride > 0 && a->pad >= 0 &&
                   a->grid_h > 0 && a->grid_w > 0, "%s: bad geometry", who);
  MV_CHECK_ARG(a->ld >= a->C * a->kh * a->kw && a->ld % 4 == 0, "%s: patch-matrix pitch %d must be >= C*kh*kw and a multiple of 4", who, a->ld);
  g->n_img = a->n_img; g->H = a->H; g->W = a->W; g->C = a->C; g->kh = a->kh; g->kw = a->kw; g->stride = a->stride; g->pad = a->pad;
  g->Gh = a->grid_h; g->Gw = a->grid_w; g->ld = a->ld; g->tc = a->tc_order ? 1 : 0;
  if (a->nchw) { g->sn = int64_t(a->C) * a->H * a->W; g->sc = int64_t(a->H) * a->W; g->sy = a->W; g->sx = 1; }
  else { g->sn = int64_t(a->H) * a->W * a->C; g->sy = int64_t(a->W) * a->C; g->sx = a->C; g->sc = 1; }
  return MV_OK;
}

}  // namespace mv

using namespace mv;

extern "C" int mv_im2col(const void* src, int src_dtype, void* cols, const mv_conv_geom* geom, void* stream) {
  MV_CHECK_ARG(src && cols, "mv_im2col: null pointer");
  ConvGeom g;
  const int rc = fill_geom(&g, geom, "mv_im2col");
  if (rc != MV_OK) return rc;
  const int64_t total = int64_t(g.n_img) * g.Gh * g.Gw * g.ld;
  const int blocks = int(std::min<int64_t>((total + 255) / 256, int64_t(num_sms()) * 16));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (src_dtype == MV_BF16 && g.tc && !geom->nchw && g.C % 8 == 0 && g.ld == g.C * g.kh * g.kw &&
      reinterpret_cast<uintptr_t>(src) % 16 == 0 && reinterpret_cast<uintptr_t>(cols) % 16 == 0) {
    const int64_t tv = total / 8;
    const int vb = int(std::min<int64_t>((tv + 255) / 256, int64_t(num_sms()) * 16));
    im2col_vec_kernel<<<vb, 256, 0, st>>>(static_cast<const bf16*>(src), static_cast<bf16*>(cols), g);
  } else if (g.tc && g.C <= 4 && (src_dtype == MV_F32 || src_dtype == MV_BF16)) {
    const int64_t tp = int64_t(g.n_img) * g.Gh * g.Gw * g.kh * g.kw;
    const int pb = int(std::min<int64_t>((tp + 255) / 256, int64_t(num_sms()) * 16));
    if (src_dtype == MV_F32) im2col_pix_kernel<float><<<pb, 256, 0, st>>>(static_cast<const float*>(src), static_cast<bf16*>(cols), g);
    else im2col_pix_kernel<bf16><<<pb, 256, 0, st>>>(static_cast<const bf16*>(src), static_cast<bf16*>(cols), g);
  } else if (src_dtype == MV_F32) im2col_kernel<float><<<blocks, 256, 0, st>>>(static_cast<const float*>(src), static_cast<bf16*>(cols), g);
  else if (src_dtype == MV_BF16) im2col_kernel<bf16><<<blocks, 256, 0, st>>>(static_cast<const bf16*>(src), static_cast<bf16*>(cols), g);
  else {
    mv::set_error("mv_im2col: unsupported dtype %d", src_dtype);
    return MV_ERR_UNSUPPORTED;
  }
  MV_CHECK_LAUNCH("mv_im2col");
  return MV_OK;
}

extern "C" int mv_col2im(const void* cols, int cols_dtype, void* dst, const mv_conv_geom* geom, const float* bias, int act, const void* dact,
                         float dslope, void* stream) {
  MV_CHECK_ARG(cols && dst, "mv_col2im: null pointer");
  MV_CHECK_ARG(act >= MV_ACT_NONE && act <= MV_ACT_SIGMOID, "mv_col2im: bad activation %d", act);
  ConvGeom g;
  const int rc = fill_geom(&g, geom, "mv_col2im");
  if (rc != MV_OK) return rc;
  const int64_t total = int64_t(g.n_img) * g.H * g.W * g.C;
  const int blocks = int(std::min<int64_t>((total + 255) / 256, int64_t(num_sms()) * 16));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool vec = g.tc && !geom->nchw && g.C % 8 == 0 && g.ld % 8 == 0 && reinterpret_cast<uintptr_t>(cols) % 16 == 0 &&
                   reinterpret_cast<uintptr_t>(dst) % 16 == 0 && (!dact || reinterpret_cast<uintptr_t>(dact) % 16 == 0) &&
                   (cols_dtype == MV_F32 || cols_dtype == MV_BF16);
  if (vec) {
    const int vb = int(std::min<int64_t>((total / 8 + 255) / 256, int64_t(num_sms()) * 16));
    if (cols_dtype == MV_F32)
      col2im_vec_kernel<float><<<vb, 256, 0, st>>>(static_cast<const float*>(cols), static_cast<bf16*>(dst), g, bias, act,
                                                   static_cast<const bf16*>(dact), dslope);
    else
      col2im_vec_kernel<bf16><<<vb, 256, 0, st>>>(static_cast<const bf16*>(cols), static_cast<bf16*>(dst), g, bias, act,
                                                  static_cast<const bf16*>(dact), dslope);
  } else if (g.tc && g.C <= 4 && (cols_dtype == MV_F32 || cols_dtype == MV_BF16) && int64_t(g.n_img) * g.H * g.W < (int64_t(1) << 31)) {
    const int64_t npix = int64_t(g.n_img) * g.H * g.W;
    const int pb = int(std::min<int64_t>((npix + 255) / 256, int64_t(num_sms()) * 16));
    if (cols_dtype == MV_F32)
      col2im_pix_kernel<float><<<pb, 256, 0, st>>>(static_cast<const float*>(cols), static_cast<bf16*>(dst), g, bias, act,
                                                   static_cast<const bf16*>(dact), dslope);
    else
      col2im_pix_kernel<bf16><<<pb, 256, 0, st>>>(static_cast<const bf16*>(cols), static_cast<bf16*>(dst), g, bias, act,
                                                  static_cast<const bf16*>(dact), dslope);
  } else if (cols_dtype == MV_F32)
    col2im_kernel<float><<<blocks, 256, 0, st>>>(static_cast<const float*>(cols), static_cast<bf16*>(dst), g, bias, act,
                                                 static_cast<const bf16*>(dact), dslope, geom->nchw ? 0 : 1);
  else if (cols_dtype == MV_BF16)
    col2im_kernel<bf16><<<blocks, 256, 0, st>>>(static_cast<const bf16*>(cols), static_cast<bf16*>(dst), g, bias, act,
                                                static_cast<const bf16*>(dact), dslope, geom->nchw ? 0 : 1);
  else {
    mv::set_error("mv_col2im: unsupported dtype %d", cols_dtype);
    return MV_ERR_UNSUPPORTED;
  }
  MV_CHECK_LAUNCH("mv_col2im");
  return MV_OK;
}

extern "C" int mv_chan_sum_nchw(const void* g, int n_img, int C, int HW, float* out, void* stream) {
  MV_CHECK_ARG(g && out && n_img > 0 && C > 0 && HW > 0, "mv_chan_sum_nchw: bad arguments");
  const int yb = std::min(n_img, std::max(1, num_sms() * 4 / C));
  chan_sum_nchw_kernel<<<dim3(C, yb), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const bf16*>(g), n_img, C, HW, out);
  MV_CHECK_LAUNCH("mv_chan_sum_nchw");
  return MV_OK;
}

static int tc_batch(const mv_pack_item* items, int n_items, TcBatch* b, int64_t* biggest, const char* who) {
  MV_CHECK_ARG(items && n_items >= 1 && n_items <= MV_PACK_MAX_ITEMS, "%s: 1 <= n_items <= %d", who, MV_PACK_MAX_ITEMS);
  *biggest = 0;
  for (int i = 0; i < n_items; ++i) {
    const mv_pack_item& w = items[i];
    MV_CHECK_ARG(w.src && w.dst_fwd && w.N >= 1 && w.C >= 1 && w.T >= 1 && w.Cpad >= w.C * w.T, "%s: bad item %d", who, i);
    b->it[i] = w;
    const int64_t total = int64_t(w.N) * w.Cpad;
    *biggest = total > *biggest ? total : *biggest;
  }
  return MV_OK;
}

extern "C" int mv_pack_tc(const mv_pack_item* items, int n_items, void* stream) {
  TcBatch b{};
  int64_t biggest;
  const int rc = tc_batch(items, n_items, &b, &biggest, "mv_pack_tc");
  if (rc != MV_OK) return rc;
  const int gx = int(std::min<int64_t>(std::max<int64_t>((biggest + 1023) / 1024, 1), 1184));
  pack_tc_kernel<<<dim3(gx, n_items), 256, 0, static_cast<cudaStream_t>(stream)>>>(b);
  MV_CHECK_LAUNCH("mv_pack_tc");
  return MV_OK;
}

extern "C" int mv_unpack_tc_add(const mv_pack_item* items, int n_items, void* stream) {
  TcBatch b{};
  int64_t biggest;
  const int rc = tc_batch(items, n_items, &b, &biggest, "mv_unpack_tc_add");
  if (rc != MV_OK) return rc;
  const int gx = int(std::min<int64_t>(std::max<int64_t>((biggest + 1023) / 1024, 1), 1184));
  unpack_tc_add_kernel<<<dim3(gx, n_items), 256, 0, static_cast<cudaStream_t>(stream)>>>(b);
  MV_CHECK_LAUNCH("mv_unpack_tc_add");
  return MV_OK;
}
