// Tap-GEMM: the tensor-core contraction behind every convolution / linear layer of the decoders and
// encoders, forward and data-gradient (sm_100a: TMA -> shared memory -> tcgen05.mma -> TMEM -> epilogue).
//
//   out[p, n] = epilogue( sum_{t < T} sum_{c < Cin}  A[p + off_t, c] * Wt[t, n, c] )
//
// A is a row-major bf16 matrix of "flat pixels" x channels.  For a 3x3/stride-1/pad-1 convolution the
// activations live in a shared-halo NHWC layout (one zero row above every image and one zero column after
// every image row), so each filter tap is the same matrix shifted by a constant number of rows and the
// whole convolution is T = 9 shifted GEMMs accumulating into one TMEM tile.  The CTA stages the input
// window [p0 - halo, p0 + 128 + halo) ONCE per 64-channel chunk with a single TMA box and feeds the nine
// taps by offsetting the UMMA shared-memory descriptor by whole 128-byte rows (the hardware swizzle is a
// function of the absolute shared-memory address, verified by tests/cuda/umma_probe.cu), so the im2col
// expansion never exists anywhere and L2/HBM read amplification is (128 + 2 halo) / 128 instead of 9.
// T = 1 with off = 0 is a plain GEMM (1x1 convolutions, Linear layers).
//
// Warp roles (192 threads, persistent over output tiles of 128 rows x BN columns):
//   warp 0   TMA producer: input windows (ring of in_stages) and weight tiles (ring, or resident for the
//            whole kernel when all T * Cin/CK tiles fit)
//   warp 1   allocates TMEM, one lane issues tcgen05.mma and commits to the mbarriers
//   warps 2-5 epilogue: tcgen05.ld the fp32 accumulator (double-buffered in TMEM so the next tile's MMAs
//            overlap), bias / activation / activation-derivative / residual / halo mask, bf16 store
#include "common.cuh"
#include "tc.cuh"

namespace mv {

using bf16 = __nv_bfloat16;

constexpr int kMaxTaps = 9;
constexpr int kMaxInStages = 4;
constexpr int kMaxWStages = 20;
constexpr int kBM = 128;

struct TapGemmParams {
  int P, m_tiles, n_tiles, BN, N_total, n_kc, T;
  int tap_off[kMaxTaps];
  int halo_lo, R;
  int in_stages, w_stages, w_resident;
  uint32_t in_stage_bytes, w_stage_bytes;
  int acc_stride, tmem_cols;
  // epilogue
  const float* bias;
  int act;
  float alpha;
  const bf16* res;
  int res_ld;
  const bf16* dact1;
  int dact1_ld;
  float slope1;
  bf16* out;
  int out_ld;
  bf16* out2;
  int out2_ld, out2_pre;
  float alpha2;
  const bf16* dact2;
  int dact2_ld;
  float slope2;
  int img_stride, Wp, W, H, n_img;
  int out_mode, n_valid;
};

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case MV_ACT_RELU: return fmaxf(v, 0.f);
    case MV_ACT_LRELU02: return v > 0.f ? v : 0.2f * v;
    case MV_ACT_SIGMOID: return 1.f / (1.f + __expf(-v));
    default: return v;
  }
}

// 8 consecutive bf16 of a row (16-byte vector) -> floats
__device__ __forceinline__ void ld8(const bf16* p, float* f) {
  const uint4 v = *reinterpret_cast<const uint4*>(p);
  f[0] = bf16lo(v.x); f[1] = bf16hi(v.x); f[2] = bf16lo(v.y); f[3] = bf16hi(v.y);
  f[4] = bf16lo(v.z); f[5] = bf16hi(v.z); f[6] = bf16lo(v.w); f[7] = bf16hi(v.w);
}
__device__ __forceinline__ void st8(bf16* p, const float* f) {
  *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
}

template <int CK>
__global__ void __launch_bounds__(192, 1)
tapgemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const TapGemmParams p) {
  constexpr uint32_t ROWB = CK * 2;            // bytes per shared-memory row (one pixel, CK channels)
  constexpr int KSTEPS = CK / 16;              // UMMA K = 16 for bf16
  constexpr uint32_t SWZ = CK == 64 ? tc::SW_128 : tc::SW_32;
  constexpr uint32_t SBO = 8 * ROWB;           // byte distance between 8-row groups
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* in_base = smem;
  uint8_t* w_base = smem + size_t(p.in_stages) * p.in_stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(w_base + size_t(p.w_stages) * p.w_stage_bytes);
  uint64_t* in_full = bars;
  uint64_t* in_empty = in_full + kMaxInStages;
  uint64_t* w_full = in_empty + kMaxInStages;
  uint64_t* w_empty = w_full + kMaxWStages;
  uint64_t* tm_full = w_empty + kMaxWStages;
  uint64_t* tm_empty = tm_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tm_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.in_stages; ++i) { tc::mbar_init(&in_full[i], 1); tc::mbar_init(&in_empty[i], 1); }
    for (int i = 0; i < p.w_stages; ++i) { tc::mbar_init(&w_full[i], 1); tc::mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&tm_full[i], 1); tc::mbar_init(&tm_empty[i], 4); }
    tc::fence_barrier_init();
    tc::prefetch_tmap(&tmA);
    tc::prefetch_tmap(&tmW);
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, p.tmem_cols);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const int n_tiles_total = p.m_tiles * p.n_tiles;

  if (warp == 0) {
    if (lane == 0) {
      // ================= TMA producer =================
      int is = 0, iph = 0, ws = 0, wph = 0;
      if (p.w_resident) {
        for (int j = 0; j < p.T * p.n_kc; ++j) {
          const int kc = j / p.T, t = j % p.T;
          tc::mbar_expect_tx(&w_full[j], p.w_stage_bytes);
          tc::tma_load_2d(w_base + size_t(j) * p.w_stage_bytes, &tmW, &w_full[j], kc * CK, t * p.N_total);
        }
      }
      for (int tile = blockIdx.x; tile < n_tiles_total; tile += gridDim.x) {
        const int p0 = (tile / p.n_tiles) * kBM, n0 = (tile % p.n_tiles) * p.BN;
        for (int kc = 0; kc < p.n_kc; ++kc) {
          tc::mbar_wait(&in_empty[is], iph ^ 1);
          tc::mbar_expect_tx(&in_full[is], uint32_t(p.R) * ROWB);
          tc::tma_load_2d(in_base + size_t(is) * p.in_stage_bytes, &tmA, &in_full[is], kc * CK, p0 - p.halo_lo);
          if (++is == p.in_stages) { is = 0; iph ^= 1; }
          if (!p.w_resident) {
            for (int t = 0; t < p.T; ++t) {
              tc::mbar_wait(&w_empty[ws], wph ^ 1);
              tc::mbar_expect_tx(&w_full[ws], p.w_stage_bytes);
              tc::tma_load_2d(w_base + size_t(ws) * p.w_stage_bytes, &tmW, &w_full[ws], kc * CK, t * p.N_total + n0);
              if (++ws == p.w_stages) { ws = 0; wph ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ================= MMA issuer =================
      const uint32_t idesc = tc::idesc_bf16(kBM, p.BN, 0, 0);
      int is = 0, iph = 0, ws = 0, wph = 0, it = 0;
      for (int tile = blockIdx.x; tile < n_tiles_total; tile += gridDim.x, ++it) {
        const int acc = it & 1, acc_ph = (it >> 1) & 1;
        tc::mbar_wait(&tm_empty[acc], acc_ph ^ 1);
        tc::fence_after_sync();
        const uint32_t tmem_d = tmem_base + uint32_t(acc * p.acc_stride);
        for (int kc = 0; kc < p.n_kc; ++kc) {
          tc::mbar_wait(&in_full[is], iph);
          tc::fence_after_sync();
          const uint32_t a_base = tc::smem_u32(in_base + size_t(is) * p.in_stage_bytes);
          for (int t = 0; t < p.T; ++t) {
            uint32_t b_base;
            if (p.w_resident) {
              const int j = kc * p.T + t;
              if (it == 0) { tc::mbar_wait(&w_full[j], 0); tc::fence_after_sync(); }
              b_base = tc::smem_u32(w_base + size_t(j) * p.w_stage_bytes);
            } else {
              tc::mbar_wait(&w_full[ws], wph);
              tc::fence_after_sync();
              b_base = tc::smem_u32(w_base + size_t(ws) * p.w_stage_bytes);
            }
            const uint32_t a_tap = a_base + uint32_t(p.halo_lo + p.tap_off[t]) * ROWB;
#pragma unroll
            for (int k = 0; k < KSTEPS; ++k) {
              tc::umma_bf16(tmem_d, tc::smem_desc(a_tap + k * 32, 16, SBO, SWZ), tc::smem_desc(b_base + k * 32, 16, SBO, SWZ),
                            idesc, (kc | t | k) != 0);
            }
            if (!p.w_resident) {
              tc::umma_commit(&w_empty[ws]);
              if (++ws == p.w_stages) { ws = 0; wph ^= 1; }
            }
          }
          tc::umma_commit(&in_empty[is]);
          if (++is == p.in_stages) { is = 0; iph ^= 1; }
        }
        tc::umma_commit(&tm_full[acc]);
      }
    }
  } else {
    // ================= epilogue (warps 2..5; TMEM lane quarter = warp % 4) =================
    const int q = warp & 3;
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles_total; tile += gridDim.x, ++it) {
      const int p0 = (tile / p.n_tiles) * kBM, n0 = (tile % p.n_tiles) * p.BN;
      const int acc = it & 1, acc_ph = (it >> 1) & 1;
      const int row = p0 + q * 32 + lane;
      bool valid = row < p.P;
      int img = 0, y = 0, x = 0;
      if (p.img_stride > 0) {
        img = row / p.img_stride;
        const int r = row - img * p.img_stride;
        y = r / p.Wp;
        x = r - y * p.Wp;
        valid = valid && img < p.n_img && y >= 1 && x < p.W;
      }
      tc::mbar_wait(&tm_full[acc], acc_ph);
      tc::fence_after_sync();
      const uint32_t taddr = tmem_base + uint32_t(acc * p.acc_stride) + (uint32_t(q * 32) << 16);
      for (int c0 = 0; c0 < p.BN; c0 += 32) {
        uint32_t v[32];
        const int ncol = p.BN - c0 >= 32 ? 32 : 16;
        if (ncol == 32) tc::tmem_ld_32x32(taddr + c0, v);
        else tc::tmem_ld_32x16(taddr + c0, v);
        tc::tmem_ld_wait();
        if (row >= p.P) continue;
        const int n = n0 + c0;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (g * 8 >= ncol) break;
          float o[8], o2[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) { o[e] = 0.f; o2[e] = 0.f; }
          if (valid) {
            float yv[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              float a = __uint_as_float(v[g * 8 + e]);
              if (p.bias) a += __ldg(p.bias + n + g * 8 + e);
              yv[e] = apply_act(a, p.act);
            }
            if (p.dact1) {
              float d[8];
              ld8(p.dact1 + size_t(row) * p.dact1_ld + n + g * 8, d);
#pragma unroll
              for (int e = 0; e < 8; ++e) yv[e] *= d[e] > 0.f ? 1.f : p.slope1;
            }
            if (p.out2 && p.out2_pre) {
#pragma unroll
              for (int e = 0; e < 8; ++e) o2[e] = yv[e];
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = p.alpha * yv[e];
            if (p.res) {
              float r[8];
              ld8(p.res + size_t(row) * p.res_ld + n + g * 8, r);
#pragma unroll
              for (int e = 0; e < 8; ++e) o[e] += r[e];
            }
            if (p.out2 && !p.out2_pre) {
#pragma unroll
              for (int e = 0; e < 8; ++e) o2[e] = p.alpha2 * o[e];
              if (p.dact2) {
                float d[8];
                ld8(p.dact2 + size_t(row) * p.dact2_ld + n + g * 8, d);
#pragma unroll
                for (int e = 0; e < 8; ++e) o2[e] *= d[e] > 0.f ? 1.f : p.slope2;
              }
            }
          }
          if (p.out_mode == 0) {
            st8(p.out + size_t(row) * p.out_ld + n + g * 8, o);
            if (p.out2) st8(p.out2 + size_t(row) * p.out2_ld + n + g * 8, o2);
          } else if (valid) {  // NCHW scatter of the first n_valid channels (decoder image head)
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int ch = n + g * 8 + e;
              if (ch < p.n_valid)
                p.out[((size_t(img) * p.n_valid + ch) * p.H + (y - 1)) * p.W + x] = __float2bfloat16_rn(o[e]);
            }
          }
        }
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&tm_empty[acc]);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, p.tmem_cols);
}

static int g_num_sms = 0;
int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

constexpr size_t kSmemLimit = 232448;  // 227 KB opt-in maximum per CTA

}  // namespace mv

using namespace mv;

extern "C" int mv_tapgemm(const mv_tapgemm_args* a, void* stream) {
  MV_CHECK_ARG(a && a->A && a->Wt && a->out, "mv_tapgemm: null pointer");
  MV_CHECK_ARG(a->P > 0 && a->Cin > 0 && a->N_total > 0, "mv_tapgemm: bad sizes");
  MV_CHECK_ARG(a->T >= 1 && a->T <= kMaxTaps, "mv_tapgemm: 1 <= taps <= %d", kMaxTaps);
  const int CK = (a->Cin % 64 == 0) ? 64 : 16;
  MV_CHECK_ARG(a->Cin % CK == 0 && (CK == 64 || a->Cin == 16), "mv_tapgemm: Cin must be 16 or a multiple of 64, got %d", a->Cin);
  MV_CHECK_ARG(a->BN == 16 || a->BN == 32 || a->BN == 64 || a->BN == 128, "mv_tapgemm: BN must be 16/32/64/128");
  MV_CHECK_ARG(a->N_total % a->BN == 0, "mv_tapgemm: N_total %% BN != 0");
  MV_CHECK_ARG(a->a_ld % 8 == 0 && (a->out_ld % 8 == 0 || a->out_mode == 1), "mv_tapgemm: leading dimensions must be multiples of 8");
  TapGemmParams p{};
  p.P = int(a->P);
  p.m_tiles = int((a->P + kBM - 1) / kBM);
  p.n_tiles = a->N_total / a->BN;
  p.BN = a->BN;
  p.N_total = a->N_total;
  p.n_kc = a->Cin / CK;
  p.T = a->T;
  int lo = 0, hi = 0;
  for (int t = 0; t < a->T; ++t) {
    p.tap_off[t] = a->tap_off[t];
    lo = a->tap_off[t] < lo ? a->tap_off[t] : lo;
    hi = a->tap_off[t] > hi ? a->tap_off[t] : hi;
  }
  p.halo_lo = -lo;
  p.R = kBM - lo + hi;
  MV_CHECK_ARG(p.R <= 256, "mv_tapgemm: tap offsets span %d rows (> 256-row TMA box)", p.R);
  const uint32_t rowb = CK * 2;
  p.in_stage_bytes = (uint32_t(p.R) * rowb + 1023u) & ~1023u;
  p.w_stage_bytes = uint32_t(a->BN) * rowb;
  const size_t fixed = 1024 /*alignment slack*/ + (2 * kMaxInStages + 2 * kMaxWStages + 4) * 8 + 16;
  const int w_tiles = p.T * p.n_kc;
  // weights resident for the whole kernel if they fit next to >= 2 input stages
  p.w_resident = (p.n_tiles == 1 && w_tiles <= kMaxWStages &&
                  fixed + size_t(w_tiles) * p.w_stage_bytes + 2 * size_t(p.in_stage_bytes) <= kSmemLimit) ? 1 : 0;
  if (p.w_resident) {
    p.w_stages = w_tiles;
    size_t left = kSmemLimit - fixed - size_t(w_tiles) * p.w_stage_bytes;
    p.in_stages = int(left / p.in_stage_bytes);
  } else {
    p.in_stages = p.n_kc >= 2 ? 3 : 2;
    size_t left = kSmemLimit - fixed - size_t(p.in_stages) * p.in_stage_bytes;
    p.w_stages = int(left / p.w_stage_bytes);
    if (p.w_stages > 8) p.w_stages = 8;
    MV_CHECK_ARG(p.w_stages >= 2, "mv_tapgemm: not enough shared memory for the weight ring");
  }
  if (p.in_stages > kMaxInStages) p.in_stages = kMaxInStages;
  MV_CHECK_ARG(p.in_stages >= 1, "mv_tapgemm: not enough shared memory for one input stage");
  p.acc_stride = a->BN < 32 ? 32 : a->BN;
  p.tmem_cols = 2 * p.acc_stride;  // 64 / 128 / 256: powers of two
  p.bias = a->bias; p.act = a->act; p.alpha = a->alpha;
  p.res = static_cast<const bf16*>(a->res); p.res_ld = a->res_ld;
  p.dact1 = static_cast<const bf16*>(a->dact1); p.dact1_ld = a->dact1_ld; p.slope1 = a->slope1;
  p.out = static_cast<bf16*>(a->out); p.out_ld = a->out_ld;
  p.out2 = static_cast<bf16*>(a->out2); p.out2_ld = a->out2_ld; p.out2_pre = a->out2_pre; p.alpha2 = a->alpha2;
  p.dact2 = static_cast<const bf16*>(a->dact2); p.dact2_ld = a->dact2_ld; p.slope2 = a->slope2;
  p.img_stride = a->img_stride; p.Wp = a->Wp; p.W = a->W; p.H = a->H; p.n_img = a->n_img;
  p.out_mode = a->out_mode; p.n_valid = a->n_valid;
  MV_CHECK_ARG(a->out_mode == 0 || (a->img_stride > 0 && a->n_valid > 0), "mv_tapgemm: NCHW scatter needs the image geometry");

  CUtensorMap tmA, tmW;
  const CUtensorMapSwizzle sw = CK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B;
  if (!tc::make_tmap_2d_bf16(&tmA, a->A, uint64_t(a->a_rows), uint64_t(a->Cin), uint64_t(a->a_ld) * 2, uint32_t(p.R), CK, sw) ||
      !tc::make_tmap_2d_bf16(&tmW, a->Wt, uint64_t(a->T) * a->N_total, uint64_t(a->Cin), uint64_t(a->Cin) * 2, uint32_t(a->BN), CK, sw)) {
    mv::set_error("mv_tapgemm: cuTensorMapEncodeTiled failed (A %p rows %lld ld %d, W %p)", a->A, (long long)a->a_rows, a->a_ld, a->Wt);
    return MV_ERR_CUDA;
  }
  const size_t smem = fixed + size_t(p.in_stages) * p.in_stage_bytes + size_t(p.w_stages) * p.w_stage_bytes;
  const int tiles = p.m_tiles * p.n_tiles;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  static bool attr_set[2] = {false, false};
  if (CK == 64) {
    if (!attr_set[0]) { cudaFuncSetAttribute(tapgemm_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemLimit)); attr_set[0] = true; }
    tapgemm_kernel<64><<<grid, 192, smem, st>>>(tmA, tmW, p);
  } else {
    if (!attr_set[1]) { cudaFuncSetAttribute(tapgemm_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemLimit)); attr_set[1] = true; }
    tapgemm_kernel<16><<<grid, 192, smem, st>>>(tmA, tmW, p);
  }
  MV_CHECK_LAUNCH("mv_tapgemm");
  return MV_OK;
}
