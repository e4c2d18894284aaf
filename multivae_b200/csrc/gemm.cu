// General tensor-core GEMM for the fully connected layers (sm_100a: TMA -> shared memory -> tcgen05.mma -> TMEM -> epilogue):
//
//   C[m, n] (+)= epilogue( sum_k A(m, k) * B(n, k) )        bf16 operands, fp32 accumulate
//
// Each operand is either K-major (stored [rows = M or N][K], the reduction index contiguous) or MN-major (stored [K][M or N]),
// which is all a Linear layer needs without ever materialising a transpose:
//   forward        Y[b, n]  = X[b, :] . W[n, :]          A = X  (K-major),  B = W  (K-major)
//   data gradient  dX[b, k] = sum_n dY[b, n] W[n, k]     A = dY (K-major),  B = W  (MN-major: stored [n][k], reduction over rows)
//   weight grad.   dW[n, k] = sum_b dY[b, n] X[b, k]     A = dY (MN-major), B = X  (MN-major), fp32 atomics into the gradient
// (default_architectures.py:21-130,225-258 and every nn.Linear of mmnist.py / svhn.py; autograd's addmm backward).
//
// CTA tile 128 x 128, K blocks of 64 (one 128-byte swizzled row per operand row), ring of 5 stages, persistent over tiles
// (x split-K slices for the weight gradient), fp32 accumulator double-buffered in TMEM (2 x 128 columns).
//   warp 0      TMA producer (one 64 x 128 box per K-major operand, two 64 x 64 boxes per MN-major operand and K block)
//   warp 1      one elected lane issues tcgen05.mma (M = 128, N = 128, K = 16) and commits to the mbarriers
//   warp 2      owns the TMEM allocation
//   warps 4-11  epilogue: two per TMEM lane quarter, 64 columns each - bias, ReLU / LeakyReLU / Sigmoid, activation-derivative
//               mask of a saved activation, then either bf16 through a swizzled staging tile + TMA store (clipped at the matrix
//               edge by the tensor map) or fp32 stores / vector reductions straight to global memory
// Out-of-range rows, columns and reduction indices are zero-filled by TMA, so M, N, K need no padding (pitches: 16 bytes).
#include <algorithm>

#include "common.cuh"
#include "tc.cuh"

namespace mv {

using bf16 = __nv_bfloat16;

constexpr int kGBM = 128, kGBN = 128, kGBK = 64;
constexpr uint32_t kGOperandBytes = 16384;            // 128 x 64 bf16
constexpr uint32_t kGStageBytes = 2 * kGOperandBytes;
constexpr int kGMaxStages = 6;
constexpr int kGThreads = 384;
constexpr int kGEpiWarps = 8;

enum : int { GEPI_BF16 = 0, GEPI_F32 = 1, GEPI_F32_ADD = 2 };

struct GemmParams {
  int M, N, K;
  int m_tiles, n_tiles, k_blocks, splits, kb_per_split;
  int a_mn, b_mn, stages, epi;
  const float* bias;
  int act;
  float neg;              // max(v, neg * v): 1 none, 0 relu, 0.2 leaky relu
  float alpha;
  const bf16* dact;       // y *= dact[m, n] > 0 ? 1 : dslope
  int dact_ld;
  float dslope;
  void* out;
  int out_ld;
  int out_vec;            // fp32 output rows are 16-byte aligned (base and pitch): vector stores / reductions
};

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// OCC = 2: the variant for short reductions (<= 2 K blocks per tile, e.g. the decoders' first Linear layer: K = 64, 419 MB of
// output).  Such a tile costs ~260 cycles of MMAs, so the kernel is paced by the latency of the per-tile epilogue chain (TMEM load ->
// arithmetic -> staging -> TMA store): two resident CTAs per SM (2 stages + staging = 100 KB, 256 TMEM columns, <= 80 registers
// each) overlap two such chains.
template <int EPI, int OCC>
__global__ void __launch_bounds__(kGThreads, OCC)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmO,
            const __grid_constant__ GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stg = smem + size_t(p.stages) * kGStageBytes;                       // 32 KB staging tile (bf16 output only)
  uint64_t* bars = reinterpret_cast<uint64_t*>(stg + (EPI == GEPI_BF16 ? 32768 : 0));
  uint64_t* full = bars;
  uint64_t* empty = full + kGMaxStages;
  uint64_t* tm_full = empty + kGMaxStages;
  uint64_t* tm_empty = tm_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tm_empty + 2);
  float* s_bias_all = reinterpret_cast<float*>(tmem_slot + 4);                 // 64 floats per epilogue warp

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&tm_full[i], 1); tc::mbar_init(&tm_empty[i], kGEpiWarps); }
    tc::fence_barrier_init();
    tc::prefetch_tmap(&tmA);
    tc::prefetch_tmap(&tmB);
  }
  if (warp == 2) tc::tmem_alloc(tmem_slot, 256);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const int n_items = p.m_tiles * p.n_tiles * p.splits;

  if (warp == 0) {
    // ================= TMA producer =================
    int s = 0, ph = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int tile = item / p.splits, split = item - tile * p.splits;
      const int m0 = (tile % p.m_tiles) * kGBM, n0 = (tile / p.m_tiles) * kGBN;
      const int kb0 = split * p.kb_per_split;
      const int kb1 = min(p.k_blocks, kb0 + p.kb_per_split);
      for (int kb = kb0; kb < kb1; ++kb) {
        tc::mbar_wait(&empty[s], ph ^ 1);
        if (tc::elect_one()) {
          uint8_t* a = smem + size_t(s) * kGStageBytes;
          uint8_t* b = a + kGOperandBytes;
          tc::mbar_expect_tx(&full[s], kGStageBytes);
          const int k0 = kb * kGBK;
          if (p.a_mn) {
            tc::tma_load_2d(a, &tmA, &full[s], m0, k0);
            tc::tma_load_2d(a + 8192, &tmA, &full[s], m0 + 64, k0);
          } else {
            tc::tma_load_2d(a, &tmA, &full[s], k0, m0);
          }
          if (p.b_mn) {
            tc::tma_load_2d(b, &tmB, &full[s], n0, k0);
            tc::tma_load_2d(b + 8192, &tmB, &full[s], n0 + 64, k0);
          } else {
            tc::tma_load_2d(b, &tmB, &full[s], k0, n0);
          }
        }
        __syncwarp();
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    const uint32_t idesc = tc::idesc_bf16(kGBM, kGBN, uint32_t(p.a_mn), uint32_t(p.b_mn));
    // K-major: rows of 128 B, 8-row groups 1 KB apart, 16 reduction elements = 32 B along the row
    // MN-major: the 64 K-rows of a block are 128 B apart (8-row groups 1 KB), the two 64-element M/N blocks 8 KB apart,
    //           16 reduction elements = 16 rows = 2 KB
    const uint64_t adesc0 = p.a_mn ? tc::smem_desc(0, 8192, 1024, tc::SW_128) : tc::smem_desc(0, 16, 1024, tc::SW_128);
    const uint64_t bdesc0 = p.b_mn ? tc::smem_desc(0, 8192, 1024, tc::SW_128) : tc::smem_desc(0, 16, 1024, tc::SW_128);
    const uint32_t a_kstep = p.a_mn ? 2048u : 32u, b_kstep = p.b_mn ? 2048u : 32u;
    int s = 0, ph = 0, it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int tile = item / p.splits, split = item - tile * p.splits;
      const int kb0 = split * p.kb_per_split;
      const int kb1 = min(p.k_blocks, kb0 + p.kb_per_split);
      const int acc = it & 1, acc_ph = (it >> 1) & 1;
      tc::mbar_wait(&tm_empty[acc], acc_ph ^ 1);
      tc::fence_after_sync();
      const uint32_t tmem_d = tmem_base + uint32_t(acc * kGBN);
      for (int kb = kb0; kb < kb1; ++kb) {
        tc::mbar_wait(&full[s], ph);
        tc::fence_after_sync();
        const uint32_t a = tc::smem_u32(smem + size_t(s) * kGStageBytes), b = a + kGOperandBytes;
        if (tc::elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ad = adesc0 | uint64_t(((a + uint32_t(k) * a_kstep) & 0x3FFFFu) >> 4);
            const uint64_t bd = bdesc0 | uint64_t(((b + uint32_t(k) * b_kstep) & 0x3FFFFu) >> 4);
            tc::umma_bf16(tmem_d, ad, bd, idesc, (kb != kb0 || k != 0) ? 1u : 0u);
          }
          tc::umma_commit(&empty[s]);
        }
        __syncwarp();
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
      if (tc::elect_one()) tc::umma_commit(&tm_full[acc]);
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ================= epilogue =================
    const int q = warp & 3, half = (warp - 4) >> 2;
    float* s_bias = s_bias_all + (warp - 4) * 64;
    int it = 0;
    bool stored = false;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int tile = item / p.splits, split = item - tile * p.splits;
      const int m0 = (tile % p.m_tiles) * kGBM, n0 = (tile / p.m_tiles) * kGBN;
      const int kb0 = split * p.kb_per_split;
      const bool has_k = kb0 < p.k_blocks;
      const int acc = it & 1, acc_ph = (it >> 1) & 1;
      const int row = m0 + q * 32 + lane;
      const int nb = n0 + half * 64;                 // first global column of this warp
      if (p.bias) {
        __syncwarp();
        for (int j = lane; j < 64; j += 32) s_bias[j] = (nb + j < p.N && (EPI != GEPI_F32_ADD || split == 0)) ? p.bias[nb + j] : 0.f;
        __syncwarp();
      }
      tc::mbar_wait(&tm_full[acc], uint32_t(acc_ph));
      tc::fence_after_sync();
      if (nb >= p.N) {
        // this warp's 64 columns lie entirely outside the matrix (N <= 64 tiles): nothing to compute or store
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&tm_empty[acc]);
        continue;
      }
      if (EPI == GEPI_BF16 && stored) {
        if (lane == 0) tc::tma_store_wait_read<0>();   // this warp's previous store must have finished reading its slab
        __syncwarp();
      }
      const uint32_t taddr = tmem_base + uint32_t(acc * kGBN + half * 64) + (uint32_t(q * 32) << 16);
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        uint32_t v[32];
        tc::tmem_ld_32x32(taddr + ch * 32, v);
        tc::tmem_ld_wait();
        const int cb = ch * 32;                      // column within the warp's 64
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float y[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) y[e] = has_k ? __uint_as_float(v[g * 8 + e]) * p.alpha : 0.f;
          if (p.bias) {
#pragma unroll
            for (int e = 0; e < 8; ++e) y[e] += s_bias[cb + g * 8 + e];
          }
          if (EPI != GEPI_F32_ADD) {
            if (p.neg != 1.f) {   // relu / leaky-relu as one max (no activation: skipped, warp-uniform)
#pragma unroll
              for (int e = 0; e < 8; ++e) y[e] = fmaxf(y[e], p.neg * y[e]);
            }
            if (p.act == MV_ACT_SIGMOID) {
              // the result is rounded to bf16 (8 mantissa bits) or feeds a Bernoulli / sigmoid decoder: the fast exponential and
              // reciprocal (2 ulp of fp32) cost a fifth of the instructions of expf + IEEE division in this issue-bound epilogue
#pragma unroll
              for (int e = 0; e < 8; ++e) y[e] = __frcp_rn(1.f + __expf(-y[e]));
            }
            if (p.dact) {
              const int n = nb + cb + g * 8;
              if (row < p.M && n < p.N) {   // N % 8 == 0 is checked on the host when dact is given
                const uint4 dv = *reinterpret_cast<const uint4*>(p.dact + size_t(row) * p.dact_ld + n);
                float d[8] = {bf16lo(dv.x), bf16hi(dv.x), bf16lo(dv.y), bf16hi(dv.y), bf16lo(dv.z), bf16hi(dv.z), bf16lo(dv.w), bf16hi(dv.w)};
#pragma unroll
                for (int e = 0; e < 8; ++e) y[e] *= d[e] > 0.f ? 1.f : p.dslope;
              }
            }
          }
          if (EPI == GEPI_BF16) {
            // staging: box `half` (64 columns) of 128 rows x 128 B, SWIZZLE_128B: 16-byte chunk j of row r at j ^ (r & 7)
            const int rloc = q * 32 + lane, j = (cb >> 3) + g;
            *reinterpret_cast<uint4*>(stg + half * 16384 + rloc * 128 + ((j ^ (rloc & 7)) << 4)) =
                make_uint4(pack_bf16(y[0], y[1]), pack_bf16(y[2], y[3]), pack_bf16(y[4], y[5]), pack_bf16(y[6], y[7]));
          } else if (row < p.M) {
            float* dst = static_cast<float*>(p.out) + size_t(row) * p.out_ld + nb + cb + g * 8;
            const int n = nb + cb + g * 8;
            const bool vec = p.out_vec && n + 8 <= p.N;
            if (EPI == GEPI_F32) {
              if (vec) {
                *reinterpret_cast<float4*>(dst) = make_float4(y[0], y[1], y[2], y[3]);
                *reinterpret_cast<float4*>(dst + 4) = make_float4(y[4], y[5], y[6], y[7]);
              } else {
#pragma unroll
                for (int e = 0; e < 8; ++e)
                  if (n + e < p.N) dst[e] = y[e];
              }
            } else if (has_k || p.bias) {
              if (vec) {
                red_add_v4(dst, y[0], y[1], y[2], y[3]);
                red_add_v4(dst + 4, y[4], y[5], y[6], y[7]);
              } else {
#pragma unroll
                for (int e = 0; e < 8; ++e)
                  if (n + e < p.N) atomicAdd(dst + e, y[e]);
              }
            }
          }
        }
      }
      // accumulator drained: hand it back before the stores go out
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&tm_empty[acc]);
      if (EPI == GEPI_BF16) {
        tc::fence_proxy_async();
        __syncwarp();
        if (lane == 0 && nb < p.N && m0 + q * 32 < p.M) {
          tc::tma_store_2d(&tmO, stg + half * 16384 + q * 4096, nb, m0 + q * 32);
          tc::tma_store_commit();
        }
        stored = true;
      }
    }
    if (EPI == GEPI_BF16 && lane == 0) tc::tma_store_wait_all<0>();
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 2) tc::tmem_dealloc(tmem_base, 256);
}

// out[n] += sum_p G[p, n] for any N % 8 == 0: a block owns 256 columns (32 groups of 8) x a slice of the rows
__global__ void __launch_bounds__(256) colsum_wide_kernel(const bf16* __restrict__ G, int64_t P, int ld, int N, float* __restrict__ out,
                                                          int rows_per_block) {
  __shared__ float red[8][32][8];
  const int cg = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int n = (blockIdx.x * 32 + cg) * 8;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (n < N) {
    const int64_t r0 = int64_t(blockIdx.y) * rows_per_block;
    const int64_t r1 = r0 + rows_per_block < P ? r0 + rows_per_block : P;
    for (int64_t r = r0 + rl; r < r1; r += 8) {
      const uint4 v = ld_stream(G + r * ld + n);
      acc[0] += bf16lo(v.x); acc[1] += bf16hi(v.x); acc[2] += bf16lo(v.y); acc[3] += bf16hi(v.y);
      acc[4] += bf16lo(v.z); acc[5] += bf16hi(v.z); acc[6] += bf16lo(v.w); acc[7] += bf16hi(v.w);
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[rl][cg][e] = acc[e];
  __syncthreads();
  if (rl == 0 && n < N) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float s = 0.f;
#pragma unroll
      for (int l = 0; l < 8; ++l) s += red[l][cg][e];
      atomicAdd(out + n + e, s);
    }
  }
}

// y = g * act'(.) on [P, N] matrices: Sigmoid' from the saved OUTPUT (y (1 - y)), ReLU / LeakyReLU' from the sign of the saved
// output; g fp32 or bf16, result bf16 (the operand of the weight / data gradient GEMMs)
template <typename TG>
__global__ void __launch_bounds__(256) act_bwd_kernel(const TG* __restrict__ g, const bf16* __restrict__ y, bf16* __restrict__ out,
                                                      int64_t n_vec, int act, float slope) {
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n_vec; i += int64_t(gridDim.x) * blockDim.x) {
    float gv[8], yv[8];
    if (sizeof(TG) == 4) {
      const uint4 a = ld_stream(reinterpret_cast<const float*>(g) + i * 8), b = ld_stream(reinterpret_cast<const float*>(g) + i * 8 + 4);
      Vec<float>::unpack(a, gv);
      Vec<float>::unpack(b, gv + 4);
    } else {
      Vec<bf16>::unpack(ld_stream(reinterpret_cast<const bf16*>(g) + i * 8), gv);
    }
    Vec<bf16>::unpack(ld_stream(y + i * 8), yv);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (act == MV_ACT_SIGMOID) gv[e] *= yv[e] * (1.f - yv[e]);
      else if (act != MV_ACT_NONE) gv[e] *= yv[e] > 0.f ? 1.f : slope;
    }
    st_stream(out + i * 8, Vec<bf16>::pack(gv));
  }
}

int num_sms();
constexpr size_t kGSmemLimit = 232448;

}  // namespace mv

using namespace mv;

extern "C" int mv_gemm(const mv_gemm_args* a, void* stream) {
  MV_CHECK_ARG(a && a->A && a->B && a->out, "mv_gemm: null pointer");
  MV_CHECK_ARG(a->M > 0 && a->N > 0 && a->K > 0, "mv_gemm: bad sizes M=%lld N=%d K=%d", (long long)a->M, a->N, a->K);
  MV_CHECK_ARG(a->M < (int64_t(1) << 31), "mv_gemm: M too large");
  MV_CHECK_ARG(a->a_ld % 8 == 0 && a->b_ld % 8 == 0, "mv_gemm: operand pitches must be multiples of 8 elements (16 bytes), got %d / %d",
               a->a_ld, a->b_ld);
  MV_CHECK_ARG(a->out_kind >= GEPI_BF16 && a->out_kind <= GEPI_F32_ADD, "mv_gemm: bad out_kind %d", a->out_kind);
  MV_CHECK_ARG(a->out_kind != GEPI_BF16 || a->out_ld % 8 == 0, "mv_gemm: bf16 output pitch must be a multiple of 8 elements");
  MV_CHECK_ARG(!a->dact || (a->N % 8 == 0 && a->dact_ld % 8 == 0), "mv_gemm: dact needs N and its pitch to be multiples of 8");
  MV_CHECK_ARG(a->act >= MV_ACT_NONE && a->act <= MV_ACT_SIGMOID, "mv_gemm: bad activation %d", a->act);
  MV_CHECK_ARG(a->out_kind != GEPI_F32_ADD || (a->act == MV_ACT_NONE && !a->dact), "mv_gemm: the accumulating output takes no activation");
  GemmParams p{};
  p.M = int(a->M); p.N = a->N; p.K = a->K;
  p.m_tiles = (p.M + kGBM - 1) / kGBM;
  p.n_tiles = (p.N + kGBN - 1) / kGBN;
  p.k_blocks = (p.K + kGBK - 1) / kGBK;
  p.a_mn = a->a_mn ? 1 : 0; p.b_mn = a->b_mn ? 1 : 0;
  p.epi = a->out_kind;
  const int tiles = p.m_tiles * p.n_tiles;
  const int sms = num_sms();
  p.splits = 1;
  if (a->out_kind == GEPI_F32_ADD && tiles < sms) {   // split the reduction until the GPU is covered (>= 2 K blocks per slice)
    int s = (sms + tiles - 1) / tiles;
    if (s > p.k_blocks / 2) s = p.k_blocks / 2;
    p.splits = s < 1 ? 1 : s;
  }
  p.kb_per_split = (p.k_blocks + p.splits - 1) / p.splits;
  p.splits = (p.k_blocks + p.kb_per_split - 1) / p.kb_per_split;
  p.stages = a->out_kind == GEPI_BF16 ? 5 : 6;
  // short reductions with many tiles: two CTAs per SM with two stages each (see gemm_kernel)
  // (tried for the fp32 results of the transposed convolutions' patch-matrix products as well: no gain — those tiles are bound by
  //  their 64 KB of direct fp32 stores, cfg3 `convt1` 308 -> 344 us)
  const bool occ2 = a->out_kind == GEPI_BF16 && p.kb_per_split <= 2 && tiles >= 4 * sms;
  if (occ2) p.stages = 2;
  p.bias = a->bias; p.act = a->act;
  p.neg = a->act == MV_ACT_LRELU02 ? 0.2f : (a->act == MV_ACT_RELU ? 0.f : 1.f);
  p.alpha = a->alpha;
  p.dact = static_cast<const bf16*>(a->dact); p.dact_ld = a->dact_ld; p.dslope = a->dslope;
  p.out = a->out; p.out_ld = a->out_ld;
  p.out_vec = (a->out_ld % 4 == 0 && reinterpret_cast<uintptr_t>(a->out) % 16 == 0) ? 1 : 0;
  CUtensorMap tmA, tmB, tmO;
  // K-major operand [rows][K]: box 64 (K) x 128 rows; MN-major operand [K][rows]: box 64 (rows) x 64 (K)
  bool ok = p.a_mn ? tc::make_tmap_2d_bf16(&tmA, a->A, uint64_t(p.K), uint64_t(p.M), uint64_t(a->a_ld) * 2, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B)
                   : tc::make_tmap_2d_bf16(&tmA, a->A, uint64_t(p.M), uint64_t(p.K), uint64_t(a->a_ld) * 2, 128, 64, CU_TENSOR_MAP_SWIZZLE_128B);
  ok = ok && (p.b_mn ? tc::make_tmap_2d_bf16(&tmB, a->B, uint64_t(p.K), uint64_t(p.N), uint64_t(a->b_ld) * 2, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B)
                     : tc::make_tmap_2d_bf16(&tmB, a->B, uint64_t(p.N), uint64_t(p.K), uint64_t(a->b_ld) * 2, 128, 64, CU_TENSOR_MAP_SWIZZLE_128B));
  tmO = tmA;
  if (ok && a->out_kind == GEPI_BF16)
    ok = tc::make_tmap_2d_bf16(&tmO, a->out, uint64_t(p.M), uint64_t(p.N), uint64_t(a->out_ld) * 2, 32, 64, CU_TENSOR_MAP_SWIZZLE_128B);
  if (!ok) {
    mv::set_error("mv_gemm: cuTensorMapEncodeTiled failed (A %p ld %d, B %p ld %d, out %p ld %d; pointers must be 16-byte aligned)", a->A,
                  a->a_ld, a->B, a->b_ld, a->out, a->out_ld);
    return MV_ERR_CUDA;
  }
  const size_t smem = 1024 + size_t(p.stages) * kGStageBytes + (a->out_kind == GEPI_BF16 ? 32768 : 0) + (2 * kGMaxStages + 4) * 8 + 16 +
                      kGEpiWarps * 64 * 4;
  MV_CHECK_ARG(smem <= kGSmemLimit, "mv_gemm: shared-memory plan exceeds the limit");
  const int items = tiles * p.splits;
  const int ctas = occ2 ? 2 * sms : sms;
  const int grid = items < ctas ? items : ctas;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define MV_G_LAUNCH(E, O)                                                                                          \
  do {                                                                                                             \
    static bool attr_done = false;                                                                                 \
    if (!attr_done) {                                                                                              \
      cudaFuncSetAttribute(gemm_kernel<E, O>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kGSmemLimit));      \
      attr_done = true;                                                                                            \
    }                                                                                                              \
    gemm_kernel<E, O><<<grid, kGThreads, smem, st>>>(tmA, tmB, tmO, p);                                            \
  } while (0)
  if (a->out_kind == GEPI_BF16 && occ2) MV_G_LAUNCH(GEPI_BF16, 2);
  else if (a->out_kind == GEPI_BF16) MV_G_LAUNCH(GEPI_BF16, 1);
  else if (a->out_kind == GEPI_F32) MV_G_LAUNCH(GEPI_F32, 1);
  else MV_G_LAUNCH(GEPI_F32_ADD, 1);
#undef MV_G_LAUNCH
  MV_CHECK_LAUNCH("mv_gemm");
  return MV_OK;
}

extern "C" int mv_colsum_any(const void* G, int64_t P, int ld, int N, float* out, void* stream) {
  MV_CHECK_ARG(G && out && P > 0 && N >= 8 && N % 8 == 0 && ld % 8 == 0, "mv_colsum_any: bad arguments (N=%d, ld=%d)", N, ld);
  if (N <= 256 && 256 % (N / 8) == 0) return mv_colsum(G, P, ld, N, out, stream);   // narrow matrices: all 256 threads on rows
  const int col_blocks = (N / 8 + 31) / 32;
  int row_blocks = (num_sms() * 4 + col_blocks - 1) / col_blocks;
  const int64_t max_rb = (P + 63) / 64;
  if (row_blocks > max_rb) row_blocks = int(max_rb);
  if (row_blocks < 1) row_blocks = 1;
  const int rows_per_block = int((P + row_blocks - 1) / row_blocks);
  colsum_wide_kernel<<<dim3(col_blocks, row_blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const bf16*>(G), P, ld, N, out,
                                                                                                  rows_per_block);
  MV_CHECK_LAUNCH("mv_colsum_any");
  return MV_OK;
}

extern "C" int mv_act_bwd(const void* g, int g_dtype, const void* y, void* out, int64_t n, int act, float slope, void* stream) {
  MV_CHECK_ARG(g && y && out && n > 0 && n % 8 == 0, "mv_act_bwd: bad arguments (n must be a multiple of 8)");
  const int64_t nv = n / 8;
  const int blocks = int(std::min<int64_t>((nv + 255) / 256, int64_t(num_sms()) * 16));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (g_dtype == MV_F32)
    act_bwd_kernel<float><<<blocks, 256, 0, st>>>(static_cast<const float*>(g), static_cast<const bf16*>(y), static_cast<bf16*>(out), nv, act, slope);
  else if (g_dtype == MV_BF16)
    act_bwd_kernel<bf16><<<blocks, 256, 0, st>>>(static_cast<const bf16*>(g), static_cast<const bf16*>(y), static_cast<bf16*>(out), nv, act, slope);
  else {
    mv::set_error("mv_act_bwd: unsupported dtype %d", g_dtype);
    return MV_ERR_UNSUPPORTED;
  }
  MV_CHECK_LAUNCH("mv_act_bwd");
  return MV_OK;
}
