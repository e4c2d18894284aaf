// HBM-bound helpers of the shared-halo NHWC layout (see include/multivae_b200.h): nearest-neighbour
// upsampling forward / backward (nn.Upsample(scale_factor=2), models/nn/mmnist.py:345), packing of the image-head
// gradient, and column sums (bias gradients).  One 16-byte vector (8 bf16 channels) per thread, coalesced.
#include <algorithm>

#include "common.cuh"

namespace mv {

using bf16 = __nv_bfloat16;

struct HaloGeom {
  int n_img, H, W, Wp, S, P;
};
static HaloGeom make_geom(int n_img, int H, int W) {
  HaloGeom g;
  g.n_img = n_img; g.H = H; g.W = W; g.Wp = W + 1; g.S = (H + 1) * (W + 1); g.P = n_img * g.S + g.Wp;
  return g;
}
// a / d for 0 <= a < 2^31, d > 0: float reciprocal estimate + exact correction (an integer division costs ~20 instructions,
// and these kernels do two or three of them per 16-byte vector)
__device__ __forceinline__ int fdiv(int a, int d) {
  int q = int(float(a) * (1.f / float(d)));
  int r = a - q * d;
  q += (r >= d) - (r < 0);
  r = a - q * d;
  q += (r >= d) - (r < 0);
  return q;
}
// row -> (img, y in 1..H, x in 0..W-1) or invalid
__device__ __forceinline__ bool decode_row(const HaloGeom& g, int row, int& img, int& y, int& x) {
  img = fdiv(row, g.S);
  const int r = row - img * g.S;
  y = fdiv(r, g.Wp);
  x = r - y * g.Wp;
  return row < g.P && img < g.n_img && y >= 1 && x < g.W;
}
// flat vector index -> (row, vector within the row); vec_per_row is a power of two for every layer of the networks
__device__ __forceinline__ void split_index(int64_t i, int vec_per_row, int& row, int& v) {
  if ((vec_per_row & (vec_per_row - 1)) == 0) {
    const int sh = 31 - __clz(vec_per_row);
    row = int(i >> sh);
    v = int(i) & (vec_per_row - 1);
  } else {
    row = int(i / vec_per_row);
    v = int(i - int64_t(row) * vec_per_row);
  }
}

__device__ __forceinline__ void unpack8f(const uint4& v, float* f) {
  f[0] = bf16lo(v.x); f[1] = bf16hi(v.x); f[2] = bf16lo(v.y); f[3] = bf16hi(v.y);
  f[4] = bf16lo(v.z); f[5] = bf16hi(v.z); f[6] = bf16lo(v.w); f[7] = bf16hi(v.w);
}
__device__ __forceinline__ uint4 pack8f(const float* f) {
  return make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
}

// out (2H x 2W) [Pout, C] <- in (H x W) [Pin, C]
__global__ void __launch_bounds__(256) upsample2x_fwd_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, HaloGeom gi,
                                                            HaloGeom go, int C) {
  const int vec_per_row = C >> 3;
  const int64_t total = int64_t(go.P) * vec_per_row;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    int row, v;
    split_index(i, vec_per_row, row, v);
    int img, y, x;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (decode_row(go, row, img, y, x)) {
      const int src = img * gi.S + (1 + ((y - 1) >> 1)) * gi.Wp + (x >> 1);
      val = ld_stream(in + int64_t(src) * C + v * 8);
    }
    st_stream(out + int64_t(row) * C + v * 8, val);
  }
}

// g_in (H x W) = sum of the 2x2 children of g_out (2H x 2W); optionally g_pre = alpha * g_in * lrelu'(act)
__global__ void __launch_bounds__(256) upsample2x_bwd_kernel(const bf16* __restrict__ g_out, const bf16* __restrict__ act,
                                                            bf16* __restrict__ g_in, bf16* __restrict__ g_pre, HaloGeom gi,
                                                            HaloGeom go, int C, float alpha, float slope) {
  const int vec_per_row = C >> 3;
  const int64_t total = int64_t(gi.P) * vec_per_row;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    int row, v;
    split_index(i, vec_per_row, row, v);
    int img, y, x;
    float s[8] = {0, 0, 0, 0, 0, 0, 0, 0}, pre[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (decode_row(gi, row, img, y, x)) {
      const int base = img * go.S + (1 + 2 * (y - 1)) * go.Wp + 2 * x;
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          float f[8];
          unpack8f(ld_stream(g_out + int64_t(base + dy * go.Wp + dx) * C + v * 8), f);
#pragma unroll
          for (int e = 0; e < 8; ++e) s[e] += f[e];
        }
      if (g_pre) {
        float a[8];
        unpack8f(ld_stream(act + int64_t(row) * C + v * 8), a);
#pragma unroll
        for (int e = 0; e < 8; ++e) pre[e] = alpha * s[e] * (a[e] > 0.f ? 1.f : slope);
      }
    }
    st_stream(g_in + int64_t(row) * C + v * 8, pack8f(s));
    if (g_pre) st_stream(g_pre + int64_t(row) * C + v * 8, pack8f(pre));
  }
}

// image-head gradient: g [n_img, ch, H, W] (dense NCHW bf16) * lrelu'(y) -> halo matrix [P, 16] (channels >= ch zero)
__global__ void __launch_bounds__(256) head_grad_pack_kernel(const bf16* __restrict__ g, const bf16* __restrict__ y_out,
                                                            bf16* __restrict__ out, HaloGeom gg, int ch, float slope) {
  for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < gg.P; row += gridDim.x * blockDim.x) {
    int img, y, x;
    float f[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) f[e] = 0.f;
    if (decode_row(gg, row, img, y, x)) {
      for (int c = 0; c < ch; ++c) {
        const int64_t idx = ((int64_t(img) * ch + c) * gg.H + (y - 1)) * gg.W + x;
        const float gv = __bfloat162float(g[idx]);
        f[c] = y_out ? gv * (__bfloat162float(y_out[idx]) > 0.f ? 1.f : slope) : gv;
      }
    }
    uint4* o = reinterpret_cast<uint4*>(out + int64_t(row) * 16);
    o[0] = pack8f(f);
    o[1] = pack8f(f + 8);
  }
}

// nn.AvgPool2d(3, stride=2, padding=1) (count_include_pad: divisor 9): in (H x W) -> out (H/2 x W/2).  The zero halo of
// the layout IS the padding, so the nine taps are read without bounds checks.
__global__ void __launch_bounds__(256) avgpool3s2_fwd_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, HaloGeom gi,
                                                            HaloGeom go, int C) {
  const int vec_per_row = C >> 3;
  const int64_t total = int64_t(go.P) * vec_per_row;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    int row, v;
    split_index(i, vec_per_row, row, v);
    int img, y, x;
    float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (decode_row(go, row, img, y, x)) {
      // centre of the window in the input: (2(y-1), 2x) in image coordinates = halo row 1 + 2(y-1)
      const int64_t centre = int64_t(img) * gi.S + int64_t(1 + 2 * (y - 1)) * gi.Wp + 2 * x;
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
          const int64_t src = centre + dy * gi.Wp + dx;
          if (src >= 0) {
            float f[8];
            unpack8f(ld_stream(in + src * C + v * 8), f);
#pragma unroll
            for (int e = 0; e < 8; ++e) s[e] += f[e];
          }
        }
#pragma unroll
      for (int e = 0; e < 8; ++e) s[e] *= (1.f / 9.f);
    }
    st_stream(out + int64_t(row) * C + v * 8, pack8f(s));
  }
}

// gradient of the pooling: g_in(Y, X) = 1/9 * sum of g_out(y, x) over the windows that contain (Y, X);
// optionally g_pre = alpha * g_in * lrelu'(act) (residual-branch gradient of the ResnetBlock below)
__global__ void __launch_bounds__(256) avgpool3s2_bwd_kernel(const bf16* __restrict__ g_out, const bf16* __restrict__ act,
                                                            bf16* __restrict__ g_in, bf16* __restrict__ g_pre, HaloGeom gi,
                                                            HaloGeom go, int C, float alpha, float slope) {
  const int vec_per_row = C >> 3;
  const int64_t total = int64_t(gi.P) * vec_per_row;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    int row, v;
    split_index(i, vec_per_row, row, v);
    int img, y, x;
    float s[8] = {0, 0, 0, 0, 0, 0, 0, 0}, pre[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (decode_row(gi, row, img, y, x)) {
      const int Y = y - 1;  // image coordinates
      // windows centred at 2*yo: contain Y iff |Y - 2 yo| <= 1
      const int yo0 = (Y + 1) >> 1, ny = (Y & 1) ? 2 : 1;          // Y odd: yo in {(Y-1)/2, (Y+1)/2}; even: {Y/2}
      const int xo0 = (x + 1) >> 1, nx = (x & 1) ? 2 : 1;
      for (int a = 0; a < ny; ++a) {
        const int yo = (Y & 1) ? ((Y - 1) >> 1) + a : yo0;
        if (yo >= go.H) continue;
        for (int b = 0; b < nx; ++b) {
          const int xo = (x & 1) ? ((x - 1) >> 1) + b : xo0;
          if (xo >= go.W) continue;
          float f[8];
          unpack8f(ld_stream(g_out + (int64_t(img) * go.S + int64_t(1 + yo) * go.Wp + xo) * C + v * 8), f);
#pragma unroll
          for (int e = 0; e < 8; ++e) s[e] += f[e];
        }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) s[e] *= (1.f / 9.f);
      if (g_pre) {
        float aa[8];
        unpack8f(ld_stream(act + int64_t(row) * C + v * 8), aa);
#pragma unroll
        for (int e = 0; e < 8; ++e) pre[e] = alpha * s[e] * (aa[e] > 0.f ? 1.f : slope);
      }
    }
    st_stream(g_in + int64_t(row) * C + v * 8, pack8f(s));
    if (g_pre) st_stream(g_pre + int64_t(row) * C + v * 8, pack8f(pre));
  }
}

// out = alpha * g * lrelu'(act)   (all [P, C] bf16; halo rows of g are zero, so they stay zero)
__global__ void __launch_bounds__(256) scale_dact_kernel(const bf16* __restrict__ g, const bf16* __restrict__ act, bf16* __restrict__ out,
                                                        int64_t nvec, float alpha, float slope) {
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec; i += int64_t(gridDim.x) * blockDim.x) {
    float f[8], a[8];
    unpack8f(ld_stream(g + i * 8), f);
    unpack8f(ld_stream(act + i * 8), a);
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = alpha * f[e] * (a[e] > 0.f ? 1.f : slope);
    st_stream(out + i * 8, pack8f(f));
  }
}

// out = leaky_relu(x)   ([P, C] bf16; halo rows are zero and stay zero): the pre-activation of the CUB ResNet blocks
__global__ void __launch_bounds__(256) lrelu_fwd_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, int64_t nvec, float slope) {
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec; i += int64_t(gridDim.x) * blockDim.x) {
    float f[8];
    unpack8f(ld_stream(x + i * 8), f);
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = f[e] > 0.f ? f[e] : slope * f[e];
    st_stream(out + i * 8, pack8f(f));
  }
}

// out[n] += sum_p G[p, n]   (N <= 256, N % 8 == 0)
__global__ void __launch_bounds__(256) colsum_kernel(const bf16* __restrict__ G, int64_t P, int ld, int N, float* __restrict__ out) {
  __shared__ float red[256 * 8];
  const int groups = N >> 3;                 // 16-byte column groups
  const int lanes = 256 / groups;            // row lanes per block
  const int cg = threadIdx.x % groups, rl = threadIdx.x / groups;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (rl < lanes) {
    for (int64_t row = int64_t(blockIdx.x) * lanes + rl; row < P; row += int64_t(gridDim.x) * lanes) {
      float f[8];
      unpack8f(ld_stream(G + row * ld + cg * 8), f);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += f[e];
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[threadIdx.x * 8 + e] = acc[e];
  __syncthreads();
  if (threadIdx.x < N) {
    const int g2 = threadIdx.x >> 3, e = threadIdx.x & 7;
    float s = 0.f;
    for (int l = 0; l < lanes; ++l) s += red[(l * groups + g2) * 8 + e];
    atomicAdd(out + threadIdx.x, s);
  }
}

int num_sms();

}  // namespace mv

using namespace mv;

extern "C" int mv_upsample2x_fwd(const void* in, void* out, int n_img, int H, int W, int C, void* stream) {
  MV_CHECK_ARG(in && out && n_img > 0 && H > 0 && W > 0 && C % 8 == 0, "mv_upsample2x_fwd: bad arguments");
  const HaloGeom gi = make_geom(n_img, H, W), go = make_geom(n_img, 2 * H, 2 * W);
  const int64_t total = int64_t(go.P) * (C / 8);
  const int blocks = int(std::min<int64_t>((total + 255) / 256, int64_t(num_sms()) * 16));
  upsample2x_fwd_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const bf16*>(in), static_cast<bf16*>(out),
                                                                               gi, go, C);
  MV_CHECK_LAUNCH("mv_upsample2x_fwd");
  return MV_OK;
}

extern "C" int mv_upsample2x_bwd(const void* g_out, const void* act, void* g_in, void* g_pre, int n_img, int H, int W, int C,
                                 float alpha, float slope, void* stream) {
  MV_CHECK_ARG(g_out && g_in && n_img > 0 && H > 0 && W > 0 && C % 8 == 0, "mv_upsample2x_bwd: bad arguments");
  MV_CHECK_ARG(!g_pre || act, "mv_upsample2x_bwd: g_pre needs the saved activation");
  const HaloGeom gi = make_geom(n_img, H, W), go = make_geom(n_img, 2 * H, 2 * W);
  const int64_t total = int64_t(gi.P) * (C / 8);
  const int blocks = int(std::min<int64_t>((total + 255) / 256, int64_t(num_sms()) * 16));
  upsample2x_bwd_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(g_out), static_cast<const bf16*>(act), static_cast<bf16*>(g_in), static_cast<bf16*>(g_pre), gi, go, C,
      alpha, slope);
  MV_CHECK_LAUNCH("mv_upsample2x_bwd");
  return MV_OK;
}

extern "C" int mv_head_grad_pack(const void* g, const void* y_out, void* out, int n_img, int H, int W, int ch, float slope,
                                 void* stream) {
  MV_CHECK_ARG(g && out && n_img > 0 && ch >= 1 && ch <= 16, "mv_head_grad_pack: bad arguments");
  const HaloGeom gg = make_geom(n_img, H, W);
  const int blocks = std::min((gg.P + 255) / 256, num_sms() * 16);
  head_grad_pack_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const bf16*>(g), static_cast<const bf16*>(y_out),
                                                                               static_cast<bf16*>(out), gg, ch, slope);
  MV_CHECK_LAUNCH("mv_head_grad_pack");
  return MV_OK;
}

extern "C" int mv_avgpool3s2_fwd(const void* in, void* out, int n_img, int H, int W, int C, void* stream) {
  MV_CHECK_ARG(in && out && n_img > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && C % 8 == 0, "mv_avgpool3s2_fwd: bad arguments");
  const HaloGeom gi = make_geom(n_img, H, W), go = make_geom(n_img, H / 2, W / 2);
  const int64_t total = int64_t(go.P) * (C / 8);
  const int blocks = int(std::min<int64_t>((total + 255) / 256, int64_t(num_sms()) * 16));
  avgpool3s2_fwd_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const bf16*>(in), static_cast<bf16*>(out),
                                                                               gi, go, C);
  MV_CHECK_LAUNCH("mv_avgpool3s2_fwd");
  return MV_OK;
}

extern "C" int mv_avgpool3s2_bwd(const void* g_out, const void* act, void* g_in, void* g_pre, int n_img, int H, int W, int C,
                                 float alpha, float slope, void* stream) {
  MV_CHECK_ARG(g_out && g_in && n_img > 0 && H % 2 == 0 && W % 2 == 0 && C % 8 == 0, "mv_avgpool3s2_bwd: bad arguments");
  MV_CHECK_ARG(!g_pre || act, "mv_avgpool3s2_bwd: g_pre needs the saved activation");
  const HaloGeom gi = make_geom(n_img, H, W), go = make_geom(n_img, H / 2, W / 2);
  const int64_t total = int64_t(gi.P) * (C / 8);
  const int blocks = int(std::min<int64_t>((total + 255) / 256, int64_t(num_sms()) * 16));
  avgpool3s2_bwd_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(g_out), static_cast<const bf16*>(act), static_cast<bf16*>(g_in), static_cast<bf16*>(g_pre), gi, go, C,
      alpha, slope);
  MV_CHECK_LAUNCH("mv_avgpool3s2_bwd");
  return MV_OK;
}

extern "C" int mv_scale_dact(const void* g, const void* act, void* out, int64_t P, int C, float alpha, float slope, void* stream) {
  MV_CHECK_ARG(g && act && out && P > 0 && C % 8 == 0, "mv_scale_dact: bad arguments");
  const int64_t nvec = P * (C / 8);
  const int blocks = int(std::min<int64_t>((nvec + 255) / 256, int64_t(num_sms()) * 16));
  scale_dact_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const bf16*>(g), static_cast<const bf16*>(act),
                                                                           static_cast<bf16*>(out), nvec, alpha, slope);
  MV_CHECK_LAUNCH("mv_scale_dact");
  return MV_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Index gathers of the fully connected layers that read / write the halo layout (nn/resnet_native.py FcHaloFn, FcFromHaloFn):
// the fp32 master weight is laid out in halo order as a bf16 GEMM operand in ONE launch (instead of cat + cast + index kernels),
// and the fp32 gradient in halo order is gathered back into the parameter's layout (optionally accumulating).
//   mode 0 (rows):     out[j, c] = src[idx[j], c]        idx[j] >= src_rows  ->  0
//   mode 1 (columns):  out[r, j] = src[r, idx[j]]        idx[j] >= src_cols  ->  0
// ---------------------------------------------------------------------------------------------------------------------
namespace mv {
template <typename TO, bool ACC>
__global__ void __launch_bounds__(256) gather2d_kernel(const float* __restrict__ src, int64_t src_rows, int64_t src_cols, int64_t src_ld,
                                                       const int64_t* __restrict__ idx, int64_t out_rows, int64_t out_cols, int mode,
                                                       TO* __restrict__ out, int64_t out_ld, int64_t out_cols_padded, float beta) {
  const int64_t total = out_rows * out_cols_padded;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = i / out_cols_padded, c = i - r * out_cols_padded;
    float v = 0.f;
    if (c < out_cols) {
      if (mode == 0) {
        const int64_t sr = idx[r];
        if (sr < src_rows) v = src[sr * src_ld + c];
      } else {
        const int64_t sc = idx[c];
        if (sc < src_cols) v = src[r * src_ld + sc];
      }
    }
    TO* o = out + r * out_ld + c;
    if constexpr (ACC) {
      if (c < out_cols) *o = beta * float(*o) + v;
    } else {
      *o = TO(v);
    }
  }
}
}  // namespace mv

extern "C" int mv_gather_cast(const float* src, int64_t src_rows, int64_t src_cols, int64_t src_ld, const int64_t* idx, int64_t out_rows,
                              int64_t out_cols, int mode, void* out_bf16, int64_t out_ld, void* stream) {
  MV_CHECK_ARG(src && idx && out_bf16 && out_rows > 0 && out_cols > 0 && out_ld >= out_cols && (mode == 0 || mode == 1),
               "mv_gather_cast: bad arguments");
  const int64_t total = out_rows * out_ld;   // padding columns [out_cols, out_ld) are written as zeros
  const int blocks = int(std::min<int64_t>((total + 255) / 256, int64_t(num_sms()) * 16));
  mv::gather2d_kernel<bf16, false><<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, src_rows, src_cols, src_ld, idx, out_rows, out_cols, mode, static_cast<bf16*>(out_bf16), out_ld, out_ld, 0.f);
  MV_CHECK_LAUNCH("mv_gather_cast");
  return MV_OK;
}

extern "C" int mv_gather_f32(const float* src, int64_t src_rows, int64_t src_cols, int64_t src_ld, const int64_t* idx, int64_t out_rows,
                             int64_t out_cols, int mode, float* out, int64_t out_ld, float beta, void* stream) {
  MV_CHECK_ARG(src && idx && out && out_rows > 0 && out_cols > 0 && out_ld >= out_cols && (mode == 0 || mode == 1),
               "mv_gather_f32: bad arguments");
  const int64_t total = out_rows * out_cols;
  const int blocks = int(std::min<int64_t>((total + 255) / 256, int64_t(num_sms()) * 16));
  if (beta == 0.f)
    mv::gather2d_kernel<float, false><<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, src_rows, src_cols, src_ld, idx, out_rows,
                                                                                            out_cols, mode, out, out_ld, out_cols, 0.f);
  else
    mv::gather2d_kernel<float, true><<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, src_rows, src_cols, src_ld, idx, out_rows,
                                                                                           out_cols, mode, out, out_ld, out_cols, beta);
  MV_CHECK_LAUNCH("mv_gather_f32");
  return MV_OK;
}

extern "C" int mv_lrelu_fwd(const void* x, void* out, int64_t P, int C, float slope, void* stream) {
  MV_CHECK_ARG(x && out && P > 0 && C % 8 == 0, "mv_lrelu_fwd: bad arguments");
  const int64_t nvec = P * (C / 8);
  const int blocks = int(std::min<int64_t>((nvec + 255) / 256, int64_t(num_sms()) * 16));
  lrelu_fwd_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const bf16*>(x), static_cast<bf16*>(out), nvec, slope);
  MV_CHECK_LAUNCH("mv_lrelu_fwd");
  return MV_OK;
}

extern "C" int mv_colsum(const void* G, int64_t P, int ld, int N, float* out, void* stream) {
  MV_CHECK_ARG(G && out && P > 0 && N % 8 == 0 && N >= 8 && N <= 256 && ld % 8 == 0, "mv_colsum: bad arguments (N=%d)", N);
  MV_CHECK_ARG(256 % (N / 8) == 0, "mv_colsum: N/8 must divide 256");
  const int lanes = 256 / (N / 8);
  const int blocks = int(std::min<int64_t>((P + lanes - 1) / lanes, int64_t(num_sms()) * 8));
  colsum_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const bf16*>(G), P, ld, N, out);
  MV_CHECK_LAUNCH("mv_colsum");
  return MV_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Weight packs of a whole network in ONE launch: fp32 Conv2d weights [N, C, kh, kw] -> bf16 tap-major K-major matrices for
// the forward pass [T * Npad, Cpad] and (taps flipped, roles swapped) for the data gradient [T * Cpad_d, Npad_d]
// (what multivae_b200/nn/halo.py pack_conv_weight / pack_conv_weight_dgrad build with ~3 ATen kernels per layer).
// ---------------------------------------------------------------------------------------------------------------------
namespace mv {
struct PackBatch {
  mv_pack_item it[MV_PACK_MAX_ITEMS];
};
__global__ void __launch_bounds__(256) pack_conv_weights_kernel(const __grid_constant__ PackBatch b) {
  const mv_pack_item& w = b.it[blockIdx.y];
  const float* __restrict__ src = static_cast<const float*>(w.src);
  bf16* __restrict__ fwd = static_cast<bf16*>(w.dst_fwd);
  bf16* __restrict__ dg = static_cast<bf16*>(w.dst_dgrad);
  const int T = w.T;
  const int64_t total = int64_t(T) * w.Npad * w.Cpad;
  for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += int64_t(gridDim.x) * blockDim.x) {
    const int c = int(e % w.Cpad);
    const int64_t tn = e / w.Cpad;
    const int n = int(tn % w.Npad), t = int(tn / w.Npad);
    const float v = (n < w.N && c < w.C) ? src[(int64_t(n) * w.C + c) * T + t] : 0.f;
    const bf16 h = __float2bfloat16_rn(v);
    if (fwd) fwd[e] = h;                                                        // [(t * Npad + n), c]
    if (dg) dg[(int64_t(T - 1 - t) * w.Cpad + c) * w.Npad + n] = h;             // [((T-1-t) * Cpad + c), n]
  }
}
}  // namespace mv

extern "C" int mv_pack_conv_weights(const mv_pack_item* items, int n_items, void* stream) {
  MV_CHECK_ARG(items && n_items >= 1 && n_items <= MV_PACK_MAX_ITEMS, "mv_pack_conv_weights: 1 <= n_items <= %d", MV_PACK_MAX_ITEMS);
  mv::PackBatch b{};
  int64_t biggest = 0;
  for (int i = 0; i < n_items; ++i) {
    const mv_pack_item& w = items[i];
    MV_CHECK_ARG(w.src && (w.dst_fwd || w.dst_dgrad), "mv_pack_conv_weights: null pointer in item %d", i);
    MV_CHECK_ARG(w.N >= 1 && w.C >= 1 && w.T >= 1 && w.Npad >= w.N && w.Cpad >= w.C, "mv_pack_conv_weights: bad sizes in item %d", i);
    b.it[i] = w;
    const int64_t total = int64_t(w.T) * w.Npad * w.Cpad;
    biggest = total > biggest ? total : biggest;
  }
  int gx = int((biggest + 256 * 4 - 1) / (256 * 4));
  gx = gx < 1 ? 1 : (gx > 1184 ? 1184 : gx);
  mv::pack_conv_weights_kernel<<<dim3(gx, n_items), 256, 0, static_cast<cudaStream_t>(stream)>>>(b);
  MV_CHECK_LAUNCH("mv_pack_conv_weights");
  return MV_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Gradient hand-over of a whole network in ONE launch: the weight-gradient kernels produce dW as [T][Npad][C] (or, for the
// role-swapped image convolution, [T][Cpad][N]) fp32 buffers; this adds them into the parameters' own .grad tensors in the
// torch Conv2d layout [N][C][T] (bias gradients are items with C = T = 1).  Replaces one permuted `grad += dW` ATen kernel
// per parameter of the autograd path.
// ---------------------------------------------------------------------------------------------------------------------
namespace mv {
struct UnpackBatch {
  mv_unpack_item it[MV_PACK_MAX_ITEMS];
};
__global__ void __launch_bounds__(256) unpack_wgrad_add_kernel(const __grid_constant__ UnpackBatch b) {
  const mv_unpack_item& w = b.it[blockIdx.y];
  const float* __restrict__ src = static_cast<const float*>(w.src);
  float* __restrict__ dst = static_cast<float*>(w.dst);
  const int64_t total = int64_t(w.N) * w.C * w.T;
  for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += int64_t(gridDim.x) * blockDim.x) {
    // e walks the destination [N][C][T] (coalesced read-modify-write of the gradient)
    const int t = int(e % w.T);
    const int64_t nc = e / w.T;
    const int c = int(nc % w.C), n = int(nc / w.C);
    const int64_t s = w.swapped ? (int64_t(t) * w.Cpad + c) * w.Npad + n     // [T][Cpad][Npad]
                                : (int64_t(t) * w.Npad + n) * w.Cpad + c;    // [T][Npad][Cpad]
    dst[e] += src[s];
  }
}
}  // namespace mv

extern "C" int mv_unpack_wgrad_add(const mv_unpack_item* items, int n_items, void* stream) {
  MV_CHECK_ARG(items && n_items >= 1 && n_items <= MV_PACK_MAX_ITEMS, "mv_unpack_wgrad_add: 1 <= n_items <= %d", MV_PACK_MAX_ITEMS);
  mv::UnpackBatch b{};
  int64_t biggest = 0;
  for (int i = 0; i < n_items; ++i) {
    const mv_unpack_item& w = items[i];
    MV_CHECK_ARG(w.src && w.dst, "mv_unpack_wgrad_add: null pointer in item %d", i);
    MV_CHECK_ARG(w.N >= 1 && w.C >= 1 && w.T >= 1 && w.Npad >= w.N && w.Cpad >= w.C, "mv_unpack_wgrad_add: bad sizes in item %d", i);
    b.it[i] = w;
    const int64_t total = int64_t(w.N) * w.C * w.T;
    biggest = total > biggest ? total : biggest;
  }
  int gx = int((biggest + 256 * 4 - 1) / (256 * 4));
  gx = gx < 1 ? 1 : (gx > 1184 ? 1184 : gx);
  mv::unpack_wgrad_add_kernel<<<dim3(gx, n_items), 256, 0, static_cast<cudaStream_t>(stream)>>>(b);
  MV_CHECK_LAUNCH("mv_unpack_wgrad_add");
  return MV_OK;
}
