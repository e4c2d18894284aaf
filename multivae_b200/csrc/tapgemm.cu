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
// Warp roles (384 threads, persistent over output tiles of 128 rows x BN columns):
//   warp 0    TMA producer: input windows (ring of in_stages) and weight tiles (ring, or resident for the
//             whole kernel when all T * Cin/CK tiles fit)
//   warp 1    one lane issues tcgen05.mma and commits to the mbarriers
//   warp 2    TMA producer of the side-input tiles (residual / activation-derivative source), own ring
//   warp 3    owns the TMEM allocation
//   warps 4-11 epilogue (two per TMEM lane quarter, half of the columns each): tcgen05.ld the fp32 accumulator
//             (double-buffered in TMEM so the next tile's MMAs overlap), bias / activation / activation-derivative /
//             residual / halo mask, bf16 pack into a swizzled staging tile, 32-row TMA stores
// 3x3 convolutions with 64 (or <= 16, NCHW) output channels are routed to the three-taps-per-MMA kernels of tapconv3.cu.
// 16-channel inputs with 64 outputs (one MMA per tap: the epilogue IS the kernel) run two resident CTAs per SM, and the
// sign-mask-only epilogue of the image head's data gradient has its own lean tile loop (profiles/r2_ncu_headd_2cta.md).
#include "common.cuh"
#include "tc.cuh"

namespace mv {

using bf16 = __nv_bfloat16;

constexpr int kMaxTaps = 9;
constexpr int kMaxInStages = 3;
constexpr int kMaxWStages = 20;
constexpr int kMaxSideStages = 4;
constexpr int kBM = 128;

struct TapGemmParams {
  int P, m_tiles, n_tiles, BN, N_total, n_kc, T;
  int tap_off[kMaxTaps];
  int halo_lo, R;
  int n_box, box_rows;      // the input window is loaded as n_box TMA boxes of box_rows rows (R > 256: two boxes)
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
  int use_tma_store;       // outputs leave through a swizzled shared-memory staging tile + TMA store
  uint32_t stage_out_bytes; // bytes of one output's staging area (BN/64 boxes of 128 rows x 128 B)
  uint32_t epi_flags;       // EF_* bits describing which epilogue terms are present
  int side_tma;             // the primary side input (res, else dact1, else dact2) arrives as TMA tiles in shared memory
  int side_kind;            // 0 res, 1 dact1, 2 dact2
  int n_out;                // staged outputs (1 or 2)
  int side_stages;          // ring depth of the side-input tiles (0: no TMA side input)
  const unsigned long long* dmask2;   // EF_DMASK2: sign bits of the second output's activation-derivative source, 64 per row
  int seq_boxes;            // BN = 128 with staged side inputs / two outputs: the epilogue handles the tile's two 64-column boxes one
                            // after the other through 16 KB buffers, which leaves shared memory for a deep weight ring
};


// 8 consecutive bf16 of a row (16-byte vector) <-> floats
__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
  f[0] = bf16lo(v.x); f[1] = bf16hi(v.x); f[2] = bf16lo(v.y); f[3] = bf16hi(v.y);
  f[4] = bf16lo(v.z); f[5] = bf16hi(v.z); f[6] = bf16lo(v.w); f[7] = bf16hi(v.w);
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  return make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
}
__device__ __forceinline__ uint4 ldg16(const bf16* p) {
  uint4 r;
  asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// Warp roles (256 threads): warp 0 input/weight TMA producer, warp 1 MMA issuer, warp 2 side-input TMA producer,
// warp 3 owns the TMEM allocation, warps 4-11 epilogue (warp % 4 = TMEM lane quarter; the two warps of a quarter split
// the tile's columns — the epilogue is a latency-bound dependent chain, so it wants warps, not unrolling).
constexpr int kEpiWarps = 8;
constexpr int kThreads = 384;

// Epilogue variants are compile-time flag sets (the runtime-flag version cost ~90 instructions per column).
enum : uint32_t {
  EF_BIAS = 1, EF_RES = 2, EF_DACT1 = 4, EF_OUT2_PRE = 8, EF_OUT2_POST = 16, EF_DACT2 = 32, EF_SIGMOID = 64,
  EF_NCHW = 128, EF_DMASK2 = 256, EF_DMASK1 = 512, EF_GENERIC = 0x80000000u
};
template <uint32_t F>
__device__ __forceinline__ bool has(const TapGemmParams& p, uint32_t bit) {
  return (F & EF_GENERIC) ? (p.epi_flags & bit) != 0 : (F & bit) != 0;
}

struct RowCtx {
  int row, rloc;  // global row, row within the tile
  bool valid;
  int img, y, x;
};

// a / d for 0 <= a < 2^31, d > 0 with a float reciprocal estimate and an exact correction (no integer division
// in the per-tile path)
__device__ __forceinline__ int fast_div(int a, int d, float inv_d) {
  int q = int(float(a) * inv_d);
  int r = a - q * d;
  q += (r >= d) - (r < 0);
  r = a - q * d;
  q += (r >= d) - (r < 0);
  return q;
}

// Side inputs (residual / activation-derivative sources).  The primary one arrives as TMA tiles in shared
// memory with the same swizzled layout as the output staging tile; others are read from global memory.
__device__ __forceinline__ uint32_t tile_off(int col, int rloc) {
  return uint32_t(col >> 6) * 16384u + uint32_t(rloc) * 128u + uint32_t((((col & 63) >> 3) ^ (rloc & 7)) << 4);
}
__device__ __forceinline__ uint4 side_vec(const TapGemmParams& p, const uint8_t* side_tile, int kind, int col, int n,
                                          const RowCtx& r) {
  if (side_tile && p.side_kind == kind) return *reinterpret_cast<const uint4*>(side_tile + tile_off(col, r.rloc));
  const bf16* base = kind == 0 ? p.res + size_t(r.row) * p.res_ld
                               : (kind == 1 ? p.dact1 + size_t(r.row) * p.dact1_ld : p.dact2 + size_t(r.row) * p.dact2_ld);
  return ldg16(base + n);
}

template <int CW, uint32_t F>
__device__ __forceinline__ void epi_apply(const TapGemmParams& p, const uint32_t* v, const uint8_t* side_tile, const float* s_bias,
                                          int cb /*column within the tile*/, int n /*global column*/, const RowCtx& r,
                                          float neg, uint8_t* st_out, uint8_t* st_out2, uint32_t mask_word) {
#pragma unroll
  for (int g = 0; g < CW / 8; ++g) {
    float o[8], o2[8];
    if (r.valid) {
      float yv[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) yv[e] = __uint_as_float(v[g * 8 + e]);
      if (has<F>(p, EF_BIAS)) {
        const float4 b0 = *reinterpret_cast<const float4*>(s_bias + cb + g * 8);
        const float4 b1 = *reinterpret_cast<const float4*>(s_bias + cb + g * 8 + 4);
        yv[0] += b0.x; yv[1] += b0.y; yv[2] += b0.z; yv[3] += b0.w;
        yv[4] += b1.x; yv[5] += b1.y; yv[6] += b1.z; yv[7] += b1.w;
      }
      // relu / leaky-relu as one max: slope `neg` is 0 / 0.2 (1 = no activation: skipped, warp-uniform)
      if (neg != 1.f) {
#pragma unroll
        for (int e = 0; e < 8; ++e) yv[e] = fmaxf(yv[e], neg * yv[e]);
      }
      if (has<F>(p, EF_SIGMOID)) {
#pragma unroll
        for (int e = 0; e < 8; ++e) yv[e] = 1.f / (1.f + __expf(-yv[e]));
      }
      if (has<F>(p, EF_DACT1)) {
        float d[8];
        unpack8(side_vec(p, side_tile, 1, cb + g * 8, n + g * 8, r), d);
#pragma unroll
        for (int e = 0; e < 8; ++e) yv[e] *= d[e] > 0.f ? 1.f : p.slope1;
      }
      if (F == EF_DMASK1) {
        // the only epilogue work of this variant (data gradient of the image head): alpha * y * lrelu'(bit) as ONE multiply
        const uint32_t m8 = (mask_word >> (g * 8)) & 0xffu;
        const float f1 = p.alpha, f0 = p.alpha * p.slope1;
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = yv[e] * (((m8 >> e) & 1u) ? f1 : f0);
      } else if (has<F>(p, EF_DMASK1)) {
        const uint32_t m8 = (mask_word >> (g * 8)) & 0xffu;
#pragma unroll
        for (int e = 0; e < 8; ++e) yv[e] *= ((m8 >> e) & 1u) ? 1.f : p.slope1;
      }
      if (has<F>(p, EF_OUT2_PRE)) {
#pragma unroll
        for (int e = 0; e < 8; ++e) o2[e] = yv[e];
      }
      if (has<F>(p, EF_RES)) {
        float rr[8];
        unpack8(side_vec(p, side_tile, 0, cb + g * 8, n + g * 8, r), rr);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = fmaf(p.alpha, yv[e], rr[e]);
      } else if (F == EF_DMASK1) {
        // (done above)
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = p.alpha * yv[e];
      }
      if (has<F>(p, EF_OUT2_POST)) {
#pragma unroll
        for (int e = 0; e < 8; ++e) o2[e] = p.alpha2 * o[e];
        if (has<F>(p, EF_DACT2)) {
          float d[8];
          unpack8(side_vec(p, side_tile, 2, cb + g * 8, n + g * 8, r), d);
#pragma unroll
          for (int e = 0; e < 8; ++e) o2[e] *= d[e] > 0.f ? 1.f : p.slope2;
        }
        if (has<F>(p, EF_DMASK2)) {
          const uint32_t m8 = (mask_word >> (g * 8)) & 0xffu;
#pragma unroll
          for (int e = 0; e < 8; ++e) o2[e] *= ((m8 >> e) & 1u) ? 1.f : p.slope2;
        }
      }
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) { o[e] = 0.f; o2[e] = 0.f; }
    }
    const int nn = n + g * 8;
    const bool two = has<F>(p, EF_OUT2_PRE) || has<F>(p, EF_OUT2_POST);
    if (has<F>(p, EF_NCHW)) {
      if (r.valid) {  // NCHW scatter of the first n_valid channels (decoder image head)
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int ch = nn + e;
          if (ch < p.n_valid)
            p.out[((size_t(r.img) * p.n_valid + ch) * p.H + (r.y - 1)) * p.W + r.x] = __float2bfloat16_rn(o[e]);
        }
      }
    } else if (st_out) {
      // staging tile: 64-column boxes of 128 rows x 128 B, SWIZZLE_128B (16-byte chunk j of row r at j ^ (r & 7))
      const uint32_t off = tile_off(cb + g * 8, r.rloc);
      *reinterpret_cast<uint4*>(st_out + off) = pack8(o);
      if (two) *reinterpret_cast<uint4*>(st_out2 + off) = pack8(o2);
    } else if (r.row < p.P) {
      *reinterpret_cast<uint4*>(p.out + size_t(r.row) * p.out_ld + nn) = pack8(o);
      if (two) *reinterpret_cast<uint4*>(p.out2 + size_t(r.row) * p.out2_ld + nn) = pack8(o2);
    }
  }
}

// The epilogue of one tile for one warp: its TMEM lane quarter (32 rows) and its half of the BN columns, in chunks of CW.
template <int BN, uint32_t F>
__device__ __forceinline__ void epi_tile(const TapGemmParams& p, const RowCtx& r, const float* s_bias, int n0, int half, uint32_t taddr,
                                         float neg, uint8_t* st_out, uint8_t* st_out2, const uint8_t* side_tile,
                                         unsigned long long mask_row) {
  constexpr int WC = BN >= 32 ? BN / 2 : BN;    // columns per warp (BN = 16: the second warp of a quarter idles)
  constexpr int CW = WC >= 32 ? 32 : 16;
  constexpr int NCH = WC / CW;
  if (BN < 32 && half != 0) return;
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    const int cb = half * WC + ch * CW;
    uint32_t v[32];
    if (CW == 32) tc::tmem_ld_32x32(taddr + cb, v);
    else tc::tmem_ld_32x16(taddr + cb, v);
    const uint32_t mask_word = uint32_t(mask_row >> ((n0 + cb) & 63));   // 32 (16) sign bits of this row's columns [n0 + cb, +CW)
    tc::tmem_ld_wait();
    epi_apply<CW, F>(p, v, side_tile, s_bias, cb, n0 + cb, r, neg, st_out, st_out2, mask_word);
  }
}

template <int CK, int BN, int TT>
__global__ void __launch_bounds__(kThreads, (CK == 16 && BN == 64) ? 2 : 1)
tapgemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
               const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmO2,
               const __grid_constant__ CUtensorMap tmS, const __grid_constant__ TapGemmParams p) {
  constexpr uint32_t ROWB = CK * 2;            // bytes per shared-memory row (one pixel, CK channels)
  constexpr int KSTEPS = CK / 16;              // UMMA K = 16 for bf16
  constexpr uint32_t SWZ = CK == 64 ? tc::SW_128 : tc::SW_32;
  constexpr uint32_t SBO = 8 * ROWB;           // byte distance between 8-row groups
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* in_base = smem;
  uint8_t* w_base = smem + size_t(p.in_stages) * p.in_stage_bytes;
  uint8_t* stg_base = w_base + ((size_t(p.w_stages) * p.w_stage_bytes + 1023) & ~size_t(1023));
  uint8_t* side_base = stg_base + (p.use_tma_store ? size_t(p.n_out) * p.stage_out_bytes : 0);
  uint64_t* bars = reinterpret_cast<uint64_t*>(side_base + size_t(p.side_stages) * p.stage_out_bytes);
  uint64_t* in_full = bars;
  uint64_t* in_empty = in_full + kMaxInStages;
  uint64_t* w_full = in_empty + kMaxInStages;
  uint64_t* w_empty = w_full + kMaxWStages;
  uint64_t* tm_full = w_empty + kMaxWStages;
  uint64_t* tm_empty = tm_full + 2;
  uint64_t* side_full = tm_empty + 2;
  uint64_t* side_empty = side_full + kMaxSideStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(side_empty + kMaxSideStages);
  float* s_bias_all = reinterpret_cast<float*>(tmem_slot + 4);   // one private copy of the tile's bias slice per epilogue warp
  const int T = TT > 0 ? TT : p.T;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.in_stages; ++i) { tc::mbar_init(&in_full[i], 1); tc::mbar_init(&in_empty[i], 1); }
    for (int i = 0; i < p.w_stages; ++i) { tc::mbar_init(&w_full[i], 1); tc::mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&tm_full[i], 1); tc::mbar_init(&tm_empty[i], kEpiWarps); }
    for (int i = 0; i < kMaxSideStages; ++i) { tc::mbar_init(&side_full[i], 1); tc::mbar_init(&side_empty[i], kEpiWarps); }
    tc::fence_barrier_init();
    tc::prefetch_tmap(&tmA);
    tc::prefetch_tmap(&tmW);
  }
  if (warp == 3) tc::tmem_alloc(tmem_slot, p.tmem_cols);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const int n_tiles_total = p.m_tiles * p.n_tiles;

  if (warp == 0) {
    // ================= TMA producer: input windows + weights =================
    // The whole warp runs the loop (warp-uniform control flow keeps addresses and coordinates in uniform
    // registers); one elected lane arms the barrier and issues the copy.
    int is = 0, iph = 0, ws = 0, wph = 0;
    if (p.w_resident) {
      for (int j = 0; j < T * p.n_kc; ++j) {
        const int kc = j / T, t = j % T;
        if (tc::elect_one()) {
          tc::mbar_expect_tx(&w_full[j], p.w_stage_bytes);
          tc::tma_load_2d(w_base + size_t(j) * p.w_stage_bytes, &tmW, &w_full[j], kc * CK, t * p.N_total);
        }
        __syncwarp();
      }
    }
    for (int tile = blockIdx.x; tile < n_tiles_total; tile += gridDim.x) {
      const int mt = p.n_tiles == 1 ? tile : tile / p.n_tiles;   // (no integer division in the common single-column-tile case)
      const int p0 = mt * kBM, n0 = (tile - mt * p.n_tiles) * BN;
      for (int kc = 0; kc < p.n_kc; ++kc) {
        tc::mbar_wait(&in_empty[is], iph ^ 1);
        if (tc::elect_one()) {
          // windows taller than the 256-row TMA box limit (64-pixel-wide images) arrive as two boxes
          tc::mbar_expect_tx(&in_full[is], uint32_t(p.n_box * p.box_rows) * ROWB);
          for (int bx = 0; bx < p.n_box; ++bx)
            tc::tma_load_2d(in_base + size_t(is) * p.in_stage_bytes + size_t(bx * p.box_rows) * ROWB, &tmA, &in_full[is], kc * CK,
                            p0 - p.halo_lo + bx * p.box_rows);
        }
        __syncwarp();
        if (++is == p.in_stages) { is = 0; iph ^= 1; }
        if (!p.w_resident) {
          for (int t = 0; t < T; ++t) {
            tc::mbar_wait(&w_empty[ws], wph ^ 1);
            if (tc::elect_one()) {
              tc::mbar_expect_tx(&w_full[ws], p.w_stage_bytes);
              tc::tma_load_2d(w_base + size_t(ws) * p.w_stage_bytes, &tmW, &w_full[ws], kc * CK, t * p.N_total + n0);
            }
            __syncwarp();
            if (++ws == p.w_stages) { ws = 0; wph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 2) {
    // ================= TMA producer: side-input tiles (their own ring, decoupled from the input prefetch) =====
    if (p.side_stages > 0) {
      int ss = 0, sph = 0;
      for (int tile = blockIdx.x; tile < n_tiles_total; tile += gridDim.x) {
        const int mt = p.n_tiles == 1 ? tile : tile / p.n_tiles;   // (no integer division in the common single-column-tile case)
      const int p0 = mt * kBM, n0 = (tile - mt * p.n_tiles) * BN;
        const int nsub = (BN == 128 && p.seq_boxes) ? 2 : 1, nbox = int(p.stage_out_bytes >> 14);
        for (int sub = 0; sub < nsub; ++sub) {
          tc::mbar_wait(&side_empty[ss], sph ^ 1);
          if (tc::elect_one()) {
            tc::mbar_expect_tx(&side_full[ss], p.stage_out_bytes);
            for (int b = 0; b < nbox; ++b)
              tc::tma_load_2d(side_base + size_t(ss) * p.stage_out_bytes + b * 16384, &tmS, &side_full[ss], n0 + (sub + b) * 64, p0);
          }
          __syncwarp();
          if (++ss == p.side_stages) { ss = 0; sph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    // Warp-uniform loop, one elected lane issues.  The 64-bit shared-memory descriptors differ only in
    // their 14-bit start-address field, so the loop carries 32-bit low words and one constant high word;
    // with the tap count known at compile time the per-tap row offsets sit in registers and the MMAs of a
    // K chunk are issued back to back.
    const uint32_t idesc = tc::idesc_bf16(kBM, BN, 0, 0);
    const uint32_t desc_hi = uint32_t(tc::smem_desc(0, 16, SBO, SWZ) >> 32);
    const uint32_t desc_lo_const = uint32_t(tc::smem_desc(0, 16, SBO, SWZ) & 0xffffffffu);
    constexpr int TU = TT > 0 ? TT : 1;
    uint32_t tap_lo[TU];   // (tap_off * ROWB) >> 4, added to the descriptor start-address field
#pragma unroll
    for (int t = 0; t < TU; ++t) tap_lo[t] = uint32_t(int(p.halo_lo + p.tap_off[t]) * int(ROWB)) >> 4;
    int is = 0, iph = 0, ws = 0, wph = 0, it = 0;
    for (int tile = blockIdx.x; tile < n_tiles_total; tile += gridDim.x, ++it) {
      const int acc = it & 1, acc_ph = (it >> 1) & 1;
      tc::mbar_wait(&tm_empty[acc], acc_ph ^ 1);
      tc::fence_after_sync();
      const uint32_t tmem_d = tmem_base + uint32_t(acc * p.acc_stride);
      for (int kc = 0; kc < p.n_kc; ++kc) {
        tc::mbar_wait(&in_full[is], iph);
        tc::fence_after_sync();
        const uint32_t a_lo0 = desc_lo_const | ((tc::smem_u32(in_base + size_t(is) * p.in_stage_bytes) & 0x3FFFFu) >> 4);
        if (p.w_resident) {
          if (it == 0) {
            for (int t = 0; t < T; ++t) tc::mbar_wait(&w_full[kc * T + t], 0);
            tc::fence_after_sync();
          }
          const uint32_t b_lo0 = desc_lo_const | ((tc::smem_u32(w_base + size_t(kc * T) * p.w_stage_bytes) & 0x3FFFFu) >> 4);
          const uint32_t b_step = p.w_stage_bytes >> 4;
          if (tc::elect_one()) {
            if (TT > 0) {
#pragma unroll
              for (int t = 0; t < TU; ++t) {
#pragma unroll
                for (int k = 0; k < KSTEPS; ++k) {
                  const uint64_t ad = (uint64_t(desc_hi) << 32) | (a_lo0 + tap_lo[t] + 2u * k);
                  const uint64_t bd = (uint64_t(desc_hi) << 32) | (b_lo0 + uint32_t(t) * b_step + 2u * k);
                  tc::umma_bf16(tmem_d, ad, bd, idesc, (kc | t | k) != 0);
                }
              }
            } else {
              for (int t = 0; t < T; ++t) {
                const uint32_t tl = uint32_t(int(p.halo_lo + p.tap_off[t]) * int(ROWB)) >> 4;
#pragma unroll
                for (int k = 0; k < KSTEPS; ++k) {
                  const uint64_t ad = (uint64_t(desc_hi) << 32) | (a_lo0 + tl + 2u * k);
                  const uint64_t bd = (uint64_t(desc_hi) << 32) | (b_lo0 + uint32_t(t) * b_step + 2u * k);
                  tc::umma_bf16(tmem_d, ad, bd, idesc, (kc | t | k) != 0);
                }
              }
            }
          }
          __syncwarp();
        } else {
          for (int t = 0; t < T; ++t) {
            tc::mbar_wait(&w_full[ws], wph);
            tc::fence_after_sync();
            const uint32_t b_lo = desc_lo_const | ((tc::smem_u32(w_base + size_t(ws) * p.w_stage_bytes) & 0x3FFFFu) >> 4);
            const uint32_t tl = TT > 0 ? 0u : (uint32_t(int(p.halo_lo + p.tap_off[t]) * int(ROWB)) >> 4);
            if (tc::elect_one()) {
              uint32_t a_lo = a_lo0 + tl;
              if (TT > 0) {
#pragma unroll
                for (int u = 0; u < TU; ++u) a_lo = (u == t) ? a_lo0 + tap_lo[u] : a_lo;
              }
#pragma unroll
              for (int k = 0; k < KSTEPS; ++k) {
                const uint64_t ad = (uint64_t(desc_hi) << 32) | (a_lo + 2u * k);
                const uint64_t bd = (uint64_t(desc_hi) << 32) | (b_lo + 2u * k);
                tc::umma_bf16(tmem_d, ad, bd, idesc, (kc | t | k) != 0);
              }
              tc::umma_commit(&w_empty[ws]);
            }
            __syncwarp();
            if (++ws == p.w_stages) { ws = 0; wph ^= 1; }
          }
        }
        if (tc::elect_one()) tc::umma_commit(&in_empty[is]);
        __syncwarp();
        if (++is == p.in_stages) { is = 0; iph ^= 1; }
      }
      if (tc::elect_one()) tc::umma_commit(&tm_full[acc]);
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ================= epilogue (warps 4..11) =================
    // A warp owns one TMEM lane quarter = 32 rows of the tile and half of its columns, with its own copy of the bias
    // slice.  Outputs go to a swizzled staging tile and leave as 32-row TMA stores: per warp when its columns are a
    // whole 64-column box (BN = 128), per pair of warps otherwise (named barrier of 64 threads) — never a CTA barrier.
    // The CTA uses (almost) all shared memory, so there is no L1: side inputs arrive as TMA tiles.
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    const bool seq = BN == 128 && p.seq_boxes != 0;           // the tile's two 64-column boxes one after the other (16 KB buffers)
    const bool pair_store = BN == 64 || seq;                  // the two warps of a quarter share one 64-column box
    const bool store_leader = lane == 0 && ((BN == 128 && !seq) || half == 0);
    float* s_bias = s_bias_all + (warp - 4) * (BN < 32 ? 32 : BN);
    const float neg = p.act == MV_ACT_LRELU02 ? 0.2f : (p.act == MV_ACT_RELU ? 0.f : 1.f);
    const float inv_S = p.img_stride > 0 ? 1.f / float(p.img_stride) : 0.f;
    const float inv_Wp = p.Wp > 0 ? 1.f / float(p.Wp) : 0.f;
    const bool two = (p.epi_flags & (EF_OUT2_PRE | EF_OUT2_POST)) != 0;
    uint8_t* st_out = p.use_tma_store ? stg_base : nullptr;
    uint8_t* st_out2 = p.use_tma_store ? stg_base + p.stage_out_bytes : nullptr;
    const int nsub = seq ? 2 : 1;
    int it = 0, ss = 0, sph = 0;
    bool stored = false;
    unsigned long long mask_next = 0ull;
    if (BN == 64 && p.epi_flags == EF_DMASK1 && p.use_tma_store && p.n_tiles == 1 && p.act == MV_ACT_NONE) {
      // ---- lean epilogue of the image head's data gradient: out = alpha * y * lrelu'(sign bit), halo rows -> 0 ----
      // The layer has K = 16, so its nine MMAs per tile cost ~460 cycles and the epilogue IS the kernel.  The generic tile loop
      // below spends ~500 warp-instructions per warp and tile on it (row decoding, variant dispatch, per-group validity
      // branches: ncu, profiles/r2_ncu_headd_2cta.md); this one keeps everything that does not change between tiles (staging
      // offsets, factors) in registers and folds the halo mask into the two factors.
      const int rloc = q * 32 + lane;
      const int cb = half * 32;
      uint32_t soff[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) soff[g] = tile_off(cb + g * 8, rloc);
      const float f1 = p.alpha, f0 = p.alpha * p.slope1;
      {
        const int row0 = int(blockIdx.x) * kBM + rloc;
        if (blockIdx.x < n_tiles_total && row0 < p.P) mask_next = p.dmask2[row0];
      }
      for (int tile = blockIdx.x; tile < n_tiles_total; tile += gridDim.x, ++it) {
        const int p0 = tile * kBM;
        const int acc = it & 1, acc_ph = (it >> 1) & 1;
        const int row = p0 + rloc;
        bool valid = row < p.P;
        if (p.img_stride > 0) {
          const int img = fast_div(row, p.img_stride, inv_S);
          const int rr = row - img * p.img_stride;
          const int y = fast_div(rr, p.Wp, inv_Wp);
          const int x = rr - y * p.Wp;
          valid = valid && img < p.n_img && y >= 1 && x < p.W;
        }
        const uint32_t mw = uint32_t(mask_next >> cb);   // the 32 sign bits of this row's columns [cb, cb + 32)
        {
          // the next tile's mask word is requested now: its L2 / HBM round trip overlaps this tile
          const int tn = tile + int(gridDim.x);
          const int rown = tn * kBM + rloc;
          if (tn < n_tiles_total && rown < p.P) mask_next = p.dmask2[rown];
        }
        const float a1 = valid ? f1 : 0.f, a0 = valid ? f0 : 0.f;
        tc::mbar_wait(&tm_full[acc], uint32_t(acc_ph));
        tc::fence_after_sync();
        uint32_t v[32];
        tc::tmem_ld_32x32(tmem_base + uint32_t(acc * p.acc_stride) + (uint32_t(q * 32) << 16) + uint32_t(cb), v);
        if (stored) {
          // the previous TMA store of this quarter's slab must have finished READING it before it is rewritten
          if (store_leader) tc::tma_store_wait_read<0>();
          asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");
        }
        tc::tmem_ld_wait();
        // accumulator drained: hand it back before the arithmetic
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&tm_empty[acc]);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = __uint_as_float(v[g * 8 + e]) * (((mw >> (g * 8 + e)) & 1u) ? a1 : a0);
          *reinterpret_cast<uint4*>(st_out + soff[g]) = pack8(o);
        }
        tc::fence_proxy_async();
        asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");
        if (store_leader) {
          tc::tma_store_2d(&tmO, st_out + q * 4096, 0, p0 + q * 32);
          tc::tma_store_commit();
        }
        stored = true;
      }
      if (store_leader) tc::tma_store_wait_all<0>();
    } else {
    for (int tile = blockIdx.x; tile < n_tiles_total; tile += gridDim.x, ++it) {
      const int mt = p.n_tiles == 1 ? tile : tile / p.n_tiles;   // (no integer division in the common single-column-tile case)
      const int p0 = mt * kBM, n0 = (tile - mt * p.n_tiles) * BN;
      const int acc = it & 1, acc_ph = (it >> 1) & 1;
      RowCtx r;
      r.rloc = q * 32 + lane;
      r.row = p0 + r.rloc;
      r.valid = r.row < p.P;
      r.img = 0; r.y = 0; r.x = 0;
      if (p.img_stride > 0) {
        r.img = fast_div(r.row, p.img_stride, inv_S);
        const int rr = r.row - r.img * p.img_stride;
        r.y = fast_div(rr, p.Wp, inv_Wp);
        r.x = rr - r.y * p.Wp;
        r.valid = r.valid && r.img < p.n_img && r.y >= 1 && r.x < p.W;
      }
      if ((p.epi_flags & EF_BIAS) && (it == 0 || p.n_tiles > 1)) {
        __syncwarp();
        for (int j = lane; j < BN; j += 32) s_bias[j] = p.bias[n0 + j];
        __syncwarp();
      }
      const uint32_t taddr = tmem_base + uint32_t(acc * p.acc_stride) + (uint32_t(q * 32) << 16);
      // sign-mask word of this row: requested ONE TILE AHEAD (first tile: here), so that the L2 / HBM round trip is over when the
      // tile's accumulator arrives (ncu: the epilogue used to wait ~15 % of its time on this load)
      unsigned long long mask_row = 0ull;
      if (p.epi_flags & (EF_DMASK2 | EF_DMASK1)) {
        if (it == 0 && r.row < p.P) mask_next = p.dmask2[r.row];
        mask_row = r.valid ? mask_next : 0ull;
        const int tn = tile + int(gridDim.x);
        if (tn < n_tiles_total) {
          const int mtn = p.n_tiles == 1 ? tn : tn / p.n_tiles;
          const int rown = mtn * kBM + r.rloc;
          if (rown < p.P) mask_next = p.dmask2[rown];
        }
      }
      tc::mbar_wait(&tm_full[acc], uint32_t(acc_ph));
      tc::fence_after_sync();
      for (int sub = 0; sub < nsub; ++sub) {
        const uint8_t* side_tile = p.side_stages > 0 ? side_base + size_t(ss) * p.stage_out_bytes : nullptr;
        if (side_tile) tc::mbar_wait(&side_full[ss], uint32_t(sph));
        if (p.use_tma_store && stored) {
          // the previous TMA stores of this slab must have finished READING it before it is rewritten
          if (store_leader) tc::tma_store_wait_read<0>();
          if (pair_store) asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");
          else __syncwarp();
        }
        {
#define MV_EPI_SWITCH(EPI)                                                              \
  switch (p.epi_flags) {                                                                \
    case 0u: EPI(0u); break;                                                            \
    case EF_BIAS: EPI(EF_BIAS); break;                                                  \
    case EF_BIAS | EF_RES | EF_OUT2_PRE: EPI(EF_BIAS | EF_RES | EF_OUT2_PRE); break;    \
    case EF_DACT1: EPI(EF_DACT1); break;                                                \
    case EF_RES: EPI(EF_RES); break;                                                    \
    case EF_OUT2_POST | EF_DACT2: EPI(EF_OUT2_POST | EF_DACT2); break;                  \
    case EF_OUT2_POST | EF_DMASK2: EPI(EF_OUT2_POST | EF_DMASK2); break;                \
    case EF_DMASK1: EPI(EF_DMASK1); break;                                              \
    case EF_BIAS | EF_NCHW: EPI(EF_BIAS | EF_NCHW); break;                              \
    default: EPI(EF_GENERIC); break;                                                    \
  }
#define MV_EPI(FLAGS) epi_tile<BN, (FLAGS)>(p, r, s_bias, n0, half, taddr, neg, st_out, st_out2, side_tile, mask_row)
#define MV_EPI64(FLAGS) \
  epi_tile<64, (FLAGS)>(p, r, s_bias + sub * 64, n0 + sub * 64, half, taddr + sub * 64, neg, st_out, st_out2, side_tile, mask_row)
          if constexpr (BN == 128) {
            if (seq) {
              MV_EPI_SWITCH(MV_EPI64)
            } else {
              MV_EPI_SWITCH(MV_EPI)
            }
          } else {
            MV_EPI_SWITCH(MV_EPI)
          }
#undef MV_EPI64
#undef MV_EPI
#undef MV_EPI_SWITCH
        }
        // accumulator drained (last box) and side tile consumed: hand them back before the stores go out
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) {
          if (sub == nsub - 1) tc::mbar_arrive(&tm_empty[acc]);
          if (side_tile) tc::mbar_arrive(&side_empty[ss]);
        }
        if (side_tile && ++ss == p.side_stages) { ss = 0; sph ^= 1; }
        if (p.use_tma_store) {
          tc::fence_proxy_async();
          if (pair_store) asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");
          else __syncwarp();
          if (store_leader) {
            const int b = (BN == 128 && !seq) ? half : 0;   // this leader's 64-column box of the staging area
            const int col = n0 + (b + sub) * 64;
            tc::tma_store_2d(&tmO, st_out + b * 16384 + q * 4096, col, p0 + q * 32);
            if (two) tc::tma_store_2d(&tmO2, st_out2 + b * 16384 + q * 4096, col, p0 + q * 32);
            tc::tma_store_commit();
          }
          stored = true;
        }
      }
    }
    if (p.use_tma_store && store_leader) tc::tma_store_wait_all<0>();
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 3) tc::tmem_dealloc(tmem_base, p.tmem_cols);
}

int conv3_try_launch(const mv_tapgemm_args* a, void* stream, bool* handled);
int head3_try_launch(const mv_tapgemm_args* a, void* stream, bool* handled);

// SM count of the CURRENT device (cached per device ordinal: one process may drive several GPUs)
int num_sms() {
  static int cache[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (cache[dev] == 0) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cache[dev] = n > 0 ? n : 148;
  }
  return cache[dev];
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
  {
    // 3x3 convolutions with 64 output channels: three-taps-per-MMA kernel (tapconv3.cu)
    bool handled = false;
    int rc = mv::conv3_try_launch(a, stream, &handled);
    if (rc != MV_OK || handled) return rc;
    rc = mv::head3_try_launch(a, stream, &handled);   // 64 -> <= 16 channel image head (NCHW output)
    if (rc != MV_OK || handled) return rc;
  }
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
  MV_CHECK_ARG(p.R <= 512, "mv_tapgemm: tap offsets span %d rows (> two 256-row TMA boxes)", p.R);
  p.n_box = p.R > 256 ? 2 : 1;
  p.box_rows = p.n_box == 1 ? p.R : (((p.R + 1) / 2 + 7) & ~7);   // whole 8-row swizzle groups per box
  const uint32_t rowb = CK * 2;
  p.in_stage_bytes = (uint32_t(p.n_box * p.box_rows) * rowb + 1023u) & ~1023u;
  p.w_stage_bytes = uint32_t(a->BN) * rowb;
  const size_t fixed = 1024 /*alignment slack*/ + (2 * kMaxInStages + 2 * kMaxWStages + 4 + 2 * kMaxSideStages) * 8 + 16 +
                       size_t(kEpiWarps) * (a->BN < 32 ? 32 : a->BN) * 4 /*bias copies*/;
  const int w_tiles = p.T * p.n_kc;
  p.use_tma_store = (a->out_mode == 0 && a->BN >= 64 && a->out_ld % 8 == 0 && (!a->out2 || a->out2_ld % 8 == 0)) ? 1 : 0;
  p.n_out = a->out2 ? 2 : 1;
  const void* side_ptr = a->res ? a->res : (a->dact1 ? a->dact1 : ((a->out2 && !a->out2_pre) ? a->dact2 : nullptr));
  const int side_ld = a->res ? a->res_ld : (a->dact1 ? a->dact1_ld : a->dact2_ld);
  p.side_kind = a->res ? 0 : (a->dact1 ? 1 : 2);
  p.side_tma = (side_ptr && a->BN >= 64 && side_ld % 8 == 0) ? 1 : 0;
  p.seq_boxes = (a->BN == 128 && p.use_tma_store && (p.side_tma || a->out2)) ? 1 : 0;
  p.stage_out_bytes = p.seq_boxes ? 16384u : uint32_t(a->BN / 64) * 16384u;
  const size_t out_stg = p.use_tma_store ? size_t(p.n_out) * p.stage_out_bytes : 0;
  auto round1k = [](size_t v) { return (v + 1023) & ~size_t(1023); };
  // shared-memory plan: try the deepest side ring first; the input ring keeps 3 stages whenever possible
  int side_try = p.side_tma ? kMaxSideStages : 0;
  for (;; --side_try) {
    const size_t stg = out_stg + size_t(side_try) * p.stage_out_bytes;
    const size_t budget = kSmemLimit - fixed - stg;
    // weights resident for the whole kernel if they fit next to the input ring
    const size_t w_res = round1k(size_t(w_tiles) * p.w_stage_bytes);
    const bool res_ok = p.n_tiles == 1 && w_tiles <= kMaxWStages && w_res + 2 * size_t(p.in_stage_bytes) <= budget;
    if (res_ok) {
      p.w_resident = 1;
      p.w_stages = w_tiles;
      p.in_stages = int((budget - w_res) / p.in_stage_bytes);
    } else {
      p.w_resident = 0;
      p.in_stages = p.n_kc >= 2 ? 3 : 2;
      const size_t in_b = size_t(p.in_stages) * p.in_stage_bytes;
      p.w_stages = budget > in_b + 1024 ? int((budget - in_b - 1024) / p.w_stage_bytes) : 0;
      if (p.w_stages > 8) p.w_stages = 8;
    }
    // a streaming weight ring wants depth (each 16 KB tile feeds only ~260 cycles of MMAs against ~1.5 k cycles of L2 latency)
    const bool deep_enough = p.w_resident ? p.in_stages >= 3 : p.w_stages >= 6;
    if (deep_enough || side_try <= (p.side_tma ? (p.seq_boxes ? 3 : 2) : 0)) break;
  }
  p.side_stages = side_try;
  MV_CHECK_ARG(p.w_resident || p.w_stages >= 2, "mv_tapgemm: not enough shared memory for the weight ring");
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

  p.epi_flags = (a->bias ? EF_BIAS : 0u) | (a->res ? EF_RES : 0u) | (a->dact1 ? EF_DACT1 : 0u) |
                ((a->out2 && a->out2_pre) ? EF_OUT2_PRE : 0u) | ((a->out2 && !a->out2_pre) ? EF_OUT2_POST : 0u) |
                ((a->out2 && !a->out2_pre && a->dact2) ? EF_DACT2 : 0u) | (a->act == MV_ACT_SIGMOID ? EF_SIGMOID : 0u) |
                (a->out_mode == 1 ? EF_NCHW : 0u) | ((a->out2 && !a->out2_pre && a->dmask2) ? EF_DMASK2 : 0u) |
                (a->dmask1 ? EF_DMASK1 : 0u);
  p.dmask2 = static_cast<const unsigned long long*>(a->dmask2 ? a->dmask2 : a->dmask1);   // one sign-mask word per row, either use
  MV_CHECK_ARG(!a->dmask2 || (a->N_total == 64 && !a->dact2), "mv_tapgemm: dmask2 needs N_total = 64 and no dact2");
  MV_CHECK_ARG(!a->dmask1 || (a->N_total == 64 && !a->dact1 && !a->dmask2), "mv_tapgemm: dmask1 needs N_total = 64, no dact1 and no dmask2");
  MV_CHECK_ARG(!a->A2, "mv_tapgemm: the fused 1x1 term (A2 / W2) is only available for 3x3 convolutions with 64 inputs and outputs in the halo layout");
  MV_CHECK_ARG(!a->res_mask, "mv_tapgemm: res_mask is only available for 3x3 convolutions with 64 outputs in the halo layout");
  MV_CHECK_ARG(!a->out2_mask, "mv_tapgemm: out2_mask is only available for 3x3 convolutions with 64 outputs in the halo layout");
  CUtensorMap tmA, tmW;
  const CUtensorMapSwizzle sw = CK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B;
  if (!tc::make_tmap_2d_bf16(&tmA, a->A, uint64_t(a->a_rows), uint64_t(a->Cin), uint64_t(a->a_ld) * 2, uint32_t(p.box_rows), CK, sw) ||
      !tc::make_tmap_2d_bf16(&tmW, a->Wt, uint64_t(a->T) * a->N_total, uint64_t(a->Cin), uint64_t(a->Cin) * 2, uint32_t(a->BN), CK, sw)) {
    mv::set_error("mv_tapgemm: cuTensorMapEncodeTiled failed (A %p rows %lld ld %d, W %p)", a->A, (long long)a->a_rows, a->a_ld, a->Wt);
    return MV_ERR_CUDA;
  }
  CUtensorMap tmO = tmA, tmO2 = tmA;  // placeholders when the direct-store path is used
  if (p.use_tma_store) {
    // each epilogue warp stores its own 32-row slab
    bool ok = tc::make_tmap_2d_bf16(&tmO, a->out, uint64_t(a->P), uint64_t(a->N_total), uint64_t(a->out_ld) * 2, 32, 64,
                                    CU_TENSOR_MAP_SWIZZLE_128B);
    if (ok && a->out2)
      ok = tc::make_tmap_2d_bf16(&tmO2, a->out2, uint64_t(a->P), uint64_t(a->N_total), uint64_t(a->out2_ld) * 2, 32, 64,
                                 CU_TENSOR_MAP_SWIZZLE_128B);
    if (!ok) {
      mv::set_error("mv_tapgemm: cuTensorMapEncodeTiled failed for the output (out %p ld %d)", a->out, a->out_ld);
      return MV_ERR_CUDA;
    }
  }
  CUtensorMap tmS = tmA;
  if (p.side_tma &&
      !tc::make_tmap_2d_bf16(&tmS, side_ptr, uint64_t(a->P), uint64_t(a->N_total), uint64_t(side_ld) * 2, 128, 64,
                             CU_TENSOR_MAP_SWIZZLE_128B)) {
    mv::set_error("mv_tapgemm: cuTensorMapEncodeTiled failed for the side input");
    return MV_ERR_CUDA;
  }
  const size_t smem = fixed + out_stg + size_t(p.side_stages) * p.stage_out_bytes + size_t(p.in_stages) * p.in_stage_bytes +
                      round1k(size_t(p.w_stages) * p.w_stage_bytes);
  MV_CHECK_ARG(smem <= kSmemLimit, "mv_tapgemm: shared-memory plan exceeds the limit (%zu bytes)", smem);
  const int tiles = p.m_tiles * p.n_tiles;
  // Cin = 16 with 64 outputs (data gradient of the image head, the encoders' image convolution): one MMA per tap and tile, so the
  // kernel is paced by the latency of its per-tile epilogue chain, not by the tensor pipe or by HBM.  These instantiations are
  // compiled for two resident CTAs per SM (80 registers; 128 of the 512 TMEM columns and ~56 KB of shared memory each).
  const int ctas_per_sm = (CK == 16 && a->BN == 64 && 2 * (smem + 1024) <= kSmemLimit) ? 2 : 1;
  const int grid = tiles < ctas_per_sm * num_sms() ? tiles : ctas_per_sm * num_sms();
  cudaStream_t st = static_cast<cudaStream_t>(stream);

#define MV_TG_LAUNCH(CK_, BN_, TT_)                                                                                      \
  do {                                                                                                                   \
    static bool attr_done = false;                                                                                       \
    if (!attr_done) {                                                                                                    \
      cudaFuncSetAttribute(tapgemm_kernel<CK_, BN_, TT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemLimit)); \
      attr_done = true;                                                                                                  \
    }                                                                                                                    \
    tapgemm_kernel<CK_, BN_, TT_><<<grid, kThreads, smem, st>>>(tmA, tmW, tmO, tmO2, tmS, p);                                 \
  } while (0)
#define MV_TG_T(CK_, BN_)                                  \
  do {                                                     \
    if (a->T == 9) MV_TG_LAUNCH(CK_, BN_, 9);              \
    else if (a->T == 1) MV_TG_LAUNCH(CK_, BN_, 1);         \
    else MV_TG_LAUNCH(CK_, BN_, 0);                        \
  } while (0)
  if (CK == 64) {
    if (a->BN == 128) MV_TG_T(64, 128);
    else if (a->BN == 64) MV_TG_T(64, 64);
    else if (a->BN == 32) MV_TG_T(64, 32);
    else MV_TG_T(64, 16);
  } else {
    if (a->BN == 128) MV_TG_T(16, 128);
    else if (a->BN == 64) MV_TG_T(16, 64);
    else if (a->BN == 32) MV_TG_T(16, 32);
    else MV_TG_T(16, 16);
  }
#undef MV_TG_T
#undef MV_TG_LAUNCH
  MV_CHECK_LAUNCH("mv_tapgemm");
  return MV_OK;
}
