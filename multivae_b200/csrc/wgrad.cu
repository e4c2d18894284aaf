// Weight gradient of the tap-GEMM layers (3x3 / 1x1 convolutions and Linear) on tcgen05:
//
//   dW[t, n, c] += sum_p  G[p, n] * X[p + off_t, c]          (fp32 accumulate, bf16 operands)
//
// The reduction runs over flat pixels p, i.e. over the ROWS of both activation matrices, so both UMMA
// operands are "MN-major": the contiguous 128-byte row of 64 channels is the M (or N) direction and K
// walks down the rows.  Exactly the same TMA boxes as the forward kernel are used (X window of
// 128 + halo rows per 64-channel chunk, SWIZZLE_128B); a filter tap is again a row shift of the
// descriptor start address.  UMMA M is always 128 = two 64-channel blocks whose distance is the
// descriptor's leading-byte-offset: either the next 64-channel chunk buffer (Cin >= 128) or — for
// Cin = 64 — the SAME buffer shifted by the row distance to another tap, so two taps share one MMA.
// Every (tap pair | chunk pair) owns one fp32 accumulator of N columns in TMEM for the whole kernel;
// a CTA walks its share of the 128-pixel K tiles, then adds its partial dW into global memory with
// fp32 reductions.  Layers whose dW exceeds 512 TMEM columns are split into passes (grid.y); every pass
// walks ALL row tiles, so the passes are planned to be identical (same chunks staged, same number of
// accumulators): they then move through the rows in lock-step and each tile is read from HBM once
// (profiles/r2_ncu_wgrad_b1c0.md), and a pass stages only the 64-channel chunks its accumulators read.
#include <cstdlib>

#include "common.cuh"
#include "tc.cuh"

namespace mv {

using bf16 = __nv_bfloat16;

constexpr int kWgMaxAcc = 16;   // accumulators per pass (16 x 32 columns = 512)
constexpr int kWgMaxStages = 4;
constexpr int kWgThreads = 192;

struct WgAcc {
  int row_off;      // window row of block 0 (halo_lo + tap offset)
  int chunk;        // 64-channel chunk buffer of block 0
  uint32_t lbo;     // byte distance from block 0 to block 1 inside the stage
  int tap[2][8];    // output tap of (M block, N block); < 0: padding, not stored.  N blocks > 0 only exist in shift mode
  int cch[2];       // output channel chunk of each M block
};

struct WgradParams {
  int n_ktiles, kt_per_cta_stride;
  int n_chunks;                 // Cin / 64
  int N, Ncols;                 // G columns, TMEM column stride per accumulator
  int n_mma;                    // MMA N: N, or 2 N in pair mode (the N side is the G tile and the G tile shifted by one row)
  uint32_t g_lbo;               // byte distance between the 64-column blocks of the N side
  uint32_t g_rows;              // rows of the G box (128, or 136 in pair mode)
  int halo_lo, R;
  int n_box, box_rows;          // the X window is loaded as n_box TMA boxes of box_rows rows
  uint32_t x_chunk_bytes;       // bytes of one chunk buffer (R rows x 128 B, 1 KB aligned)
  uint32_t g_bytes, g_row_bytes, g_box_cols;
  uint32_t stage_bytes[8];      // per pass: a pass stages only the 64-channel chunks its accumulators read
  int stages[8];
  int chunk_lo[8], chunk_n[8];  // first chunk / number of chunks staged by the pass (WgAcc::chunk is relative to chunk_lo)
  int n_acc[8];                 // accumulators per pass
  int acc_begin[8];
  WgAcc acc[48];
  int Cin, N_total_out;         // dW is [T][N_total_out][Cin] fp32
  int nct, T, n_valid;          // nct: dW is [N][Cin][T] instead (the torch Conv2d layout); output columns >= n_valid are padding
  float* dW;
  float* db;                    // optional bias gradient: db[n] += sum_p G[p, n] (fused column sums of the G tiles)
  int tmem_cols;
};

__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmG, const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int pass = blockIdx.y;
  const int n_stages = p.stages[pass];
  const uint32_t stage_bytes = p.stage_bytes[pass];
  const int chunk_lo = p.chunk_lo[pass], chunk_n = p.chunk_n[pass];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + size_t(n_stages) * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kWgMaxStages;
  uint64_t* done = empty + kWgMaxStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  float* red = reinterpret_cast<float*>(tmem_slot + 4);   // 128 x 8 partial column sums
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool do_db = p.db != nullptr && pass == 0;   // the bias gradient is independent of the pass: pass 0 owns it
  const int nacc = p.n_acc[pass];
  const WgAcc* accs = p.acc + p.acc_begin[pass];

  if (threadIdx.x == 0) {
    for (int i = 0; i < n_stages; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], do_db ? 5 : 1); }
    tc::mbar_init(done, 1);
    tc::fence_barrier_init();
    tc::prefetch_tmap(&tmX);
    tc::prefetch_tmap(&tmG);
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, p.tmem_cols);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t g_off = uint32_t(chunk_n) * p.x_chunk_bytes;  // G tile follows the X chunk buffers in a stage

  if (warp == 0) {
    // ================= TMA producer =================
    int s = 0, ph = 0;
    for (int kt = blockIdx.x; kt < p.n_ktiles; kt += gridDim.x) {
      tc::mbar_wait(&empty[s], ph ^ 1);
      if (tc::elect_one()) {
        uint8_t* st = smem + size_t(s) * stage_bytes;
        tc::mbar_expect_tx(&full[s], uint32_t(chunk_n) * uint32_t(p.n_box * p.box_rows) * 128u + p.g_bytes);
        for (int c = 0; c < chunk_n; ++c)
          for (int bx = 0; bx < p.n_box; ++bx)   // windows taller than the 256-row TMA box limit arrive as two boxes
            tc::tma_load_2d(st + size_t(c) * p.x_chunk_bytes + size_t(bx * p.box_rows) * 128u, &tmX, &full[s], (chunk_lo + c) * 64,
                            kt * 128 - p.halo_lo + bx * p.box_rows);
        const int nbox = p.N / int(p.g_box_cols);
        for (int b = 0; b < nbox; ++b)
          tc::tma_load_2d(st + g_off + size_t(b) * p.g_rows * p.g_row_bytes, &tmG, &full[s], b * int(p.g_box_cols), kt * 128);
      }
      __syncwarp();
      if (++s == n_stages) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    const uint32_t idesc = tc::idesc_bf16(128, uint32_t(p.n_mma), 1, 1);
    const uint32_t g_sw = p.g_row_bytes == 128 ? tc::SW_128 : tc::SW_32;
    const uint32_t g_kstep = 16 * p.g_row_bytes;                       // 16 pixel rows per UMMA K step
    // G: blocks of 64 columns are separate boxes (128 rows x 128 B) -> LBO = 16 KB; 8-row groups -> SBO
    const uint64_t g_desc0 = tc::smem_desc(0, p.g_lbo, 8 * p.g_row_bytes, g_sw);
    int s = 0, ph = 0, it = 0;
    for (int kt = blockIdx.x; kt < p.n_ktiles; kt += gridDim.x, ++it) {
      tc::mbar_wait(&full[s], ph);
      tc::fence_after_sync();
      const uint32_t st = tc::smem_u32(smem + size_t(s) * stage_bytes);
      if (tc::elect_one()) {
        for (int a = 0; a < nacc; ++a) {
          const WgAcc& A = accs[a];
          const uint32_t x0 = st + uint32_t(A.chunk) * p.x_chunk_bytes + uint32_t(A.row_off) * 128u;
          const uint64_t xd0 = tc::smem_desc(0, A.lbo, 1024, tc::SW_128);
          const uint32_t tm = tmem_base + uint32_t(a * p.Ncols);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint64_t xd = xd0 | uint64_t(((x0 + uint32_t(k) * 2048u) & 0x3FFFFu) >> 4);
            const uint64_t gd = g_desc0 | uint64_t(((st + g_off + uint32_t(k) * g_kstep) & 0x3FFFFu) >> 4);
            tc::umma_bf16(tm, xd, gd, idesc, (it | k) != 0);
          }
        }
        tc::umma_commit(&empty[s]);
      }
      __syncwarp();
      if (++s == n_stages) { s = 0; ph ^= 1; }
    }
    if (tc::elect_one()) tc::umma_commit(done);
    __syncwarp();
  } else {
    // ================= bias gradient: column sums of the G tiles while the MMAs run =================
    if (do_db) {
      const int et = threadIdx.x - 64;                       // 0..127
      const int groups = p.N >> 3;                           // 16-byte column groups: 2 / 8 / 16
      const int cj = et % groups, rg = et / groups;          // this thread: column group cj, rows rg*rpt .. +rpt-1
      const int rpt = 128 / (128 / groups);                  // rows per thread = groups
      float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      int s = 0, ph = 0;
      for (int kt = blockIdx.x; kt < p.n_ktiles; kt += gridDim.x) {
        tc::mbar_wait(&full[s], ph);
        const uint8_t* gt = smem + size_t(s) * stage_bytes + g_off;
        for (int i = 0; i < rpt; ++i) {
          const int row = rg * rpt + i;
          uint32_t off;
          if (p.g_row_bytes == 128) off = uint32_t(cj >> 3) * (p.g_rows * 128u) + uint32_t(row) * 128u + uint32_t(((cj & 7) ^ (row & 7)) << 4);
          else off = uint32_t(row) * p.g_row_bytes + uint32_t((cj ^ ((row >> 2) & 1)) << 4);   // SWIZZLE_32B
          const uint4 v = *reinterpret_cast<const uint4*>(gt + off);
          acc[0] += bf16lo(v.x); acc[1] += bf16hi(v.x); acc[2] += bf16lo(v.y); acc[3] += bf16hi(v.y);
          acc[4] += bf16lo(v.z); acc[5] += bf16hi(v.z); acc[6] += bf16lo(v.w); acc[7] += bf16hi(v.w);
        }
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&empty[s]);
        if (++s == n_stages) { s = 0; ph ^= 1; }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) red[et * 8 + e] = acc[e];
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (et < p.N && et < p.n_valid) {
        const int g2 = et >> 3, e = et & 7;
        float sum = 0.f;
        for (int r2 = 0; r2 < 128 / groups; ++r2) sum += red[(r2 * groups + g2) * 8 + e];
        atomicAdd(p.db + et, sum);
      }
    }
    // ================= epilogue: partial dW -> global fp32 reductions =================
    const int q = warp & 3;
    tc::mbar_wait(done, 0);
    tc::fence_after_sync();
    const bool any = blockIdx.x < p.n_ktiles;   // a CTA without K tiles has nothing in TMEM
    const int m = q * 32 + lane;                // accumulator row = (block, channel)
    const int blk = m >> 6, ch = m & 63;
    for (int a = 0; a < nacc; ++a) {
      const WgAcc& A = accs[a];
      const int cc = A.cch[blk];
      for (int n0 = 0; n0 < p.n_mma; n0 += 16) {
        const int nb = n0 / p.N;                   // N block (shift mode: the copy of G shifted by nb rows)
        const int t = A.tap[blk][nb];
        uint32_t v[16];
        tc::tmem_ld_32x16(tmem_base + uint32_t(a * p.Ncols + n0) + (uint32_t(q * 32) << 16), v);
        tc::tmem_ld_wait();
        if (any && t >= 0) {
          const int nn = n0 - nb * p.N;
          if (p.nct) {
            float* dst = p.dW + (size_t(nn) * p.Cin + cc * 64 + ch) * p.T + t;
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (nn + j < p.n_valid) atomicAdd(dst + size_t(j) * p.Cin * p.T, __uint_as_float(v[j]));
          } else {
            float* dst = p.dW + (size_t(t) * p.N_total_out + nn) * p.Cin + cc * 64 + ch;
#pragma unroll
            for (int j = 0; j < 16; ++j) atomicAdd(dst + size_t(j) * p.Cin, __uint_as_float(v[j]));
          }
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, p.tmem_cols);
}

int num_sms();
constexpr size_t kWgSmemLimit = 232448;

}  // namespace mv

using namespace mv;

static int wgrad_launch(const void* X, int64_t x_rows, int x_ld, int Cin, const void* G, int64_t g_rows, int g_ld, int N,
                        int T, const int* tap_off, int64_t P, float* dW, int N_total, int n_offset, float* db, int nct, int n_valid,
                        void* stream) {
  MV_CHECK_ARG(N_total >= N && n_offset >= 0 && n_offset + N <= N_total, "mv_wgrad: bad column slice");
  MV_CHECK_ARG(X && G && dW && tap_off, "mv_wgrad: null pointer");
  MV_CHECK_ARG(Cin % 64 == 0 && Cin >= 64 && Cin <= 256, "mv_wgrad: Cin must be 64/128/192/256, got %d", Cin);
  MV_CHECK_ARG(N == 16 || N == 64 || N == 128, "mv_wgrad: N must be 16, 64 or 128, got %d", N);
  MV_CHECK_ARG(T >= 1 && T <= 9, "mv_wgrad: 1 <= taps <= 9");
  MV_CHECK_ARG(x_ld % 8 == 0 && g_ld % 8 == 0, "mv_wgrad: leading dimensions must be multiples of 8");
  WgradParams p{};
  int lo = 0, hi = 0;
  for (int t = 0; t < T; ++t) {
    lo = tap_off[t] < lo ? tap_off[t] : lo;
    hi = tap_off[t] > hi ? tap_off[t] : hi;
  }
  p.halo_lo = -lo;
  p.R = 128 - lo + hi;
  MV_CHECK_ARG(p.R <= 512, "mv_wgrad: tap offsets span %d rows", p.R);
  p.n_box = p.R > 256 ? 2 : 1;
  p.box_rows = p.n_box == 1 ? p.R : (((p.R + 1) / 2 + 7) & ~7);
  p.n_chunks = Cin / 64;
  p.N = N;
  p.Ncols = N < 32 ? 32 : N;
  p.Cin = Cin;
  p.N_total_out = N_total;
  p.nct = nct;
  p.T = T;
  p.n_valid = n_valid;
  p.dW = dW + size_t(n_offset) * Cin * (nct ? T : 1);
  p.db = db ? db + n_offset : nullptr;
  p.n_ktiles = int((P + 127) / 128);
  // +8 rows of slack: the padding block of an odd tap count may read a few rows past the window
  p.x_chunk_bytes = (uint32_t(p.n_box * p.box_rows + 8) * 128u + 1023u) & ~1023u;
  // Pair mode (Cin = 64, N = 64, a 3x3 tap pattern): the plain scheme issues N = 64 MMAs, which are bound by the operand
  // fetch (51 cycles against 32 of math, tests/cuda/mma_rate.cu).  Here the N side is the G tile AND the G tile shifted
  // by one row (N = 128, math-bound), the M side two X shifts per filter row:
  //   D[(xa, c), (gb, n)] = sum_p X[p + xa, c] G[p + gb, n] = dW at tap offset xa - gb      (the shifted sum covers the same
  //   rows over all tiles; rows outside the tensors read as zeros)
  // with xa in {base - 1, base + 1}, gb in {0, 1}: offsets base-1, base-2 (unused), base+1, base -> the three taps of a
  // filter row from 8 MMAs instead of 12, each twice as efficient.
  // Cin = 128, N = 64 (conv0 of the 128 -> 64 block at 14x14): the same idea with the two 64-channel chunks as the M side:
  //   D[(chunk, c), (gb, n)] = sum_p X[p + xa, chunk*64 + c] G[p + gb, n] = dW at tap offset xa - gb
  // xa = base gives the taps s = 1 and s = 0 of a filter row, xa = base + 1 the tap s = 2 (its second half repeats s = 1 and is not
  // stored): 2 MMAs of N = 128 (64.5 cycles each) per filter row and K step instead of 3 of N = 64 (51 each), and the six
  // accumulators split into two equal passes of 3 x 128 TMEM columns (before: 9 accumulators in passes of 5 and 4).
  int Wp3 = 0;
  bool pair = false, pair128 = false;
  if ((Cin == 64 || Cin == 128) && (N == 64 || (N == 16 && Cin == 64)) && T == 9 && (N_total == N || nct)) {
    Wp3 = tap_off[7] - tap_off[4];
    bool ok = Wp3 >= 2;
    for (int r = 0; r < 3 && ok; ++r)
      for (int s2 = 0; s2 < 3; ++s2)
        if (tap_off[3 * r + s2] != (r - 1) * Wp3 + (s2 - 1)) ok = false;
    pair = ok && Cin == 64;
    pair128 = ok && Cin == 128;
  }
  // N = 16 (image head): four row-shifted copies of the 16-column G tile (MMA N = 64; shifts 0..2 are the three taps of a filter
  // row, the fourth copy is padding).  Eight copies (N = 128) cost 64.5 cycles per MMA against 51 and made the kernel MMA-bound
  // (16 MMAs per 128-row tile) on a layer that moves 1.7 GB for 0.2 TFLOP.
  const int n_shift = pair128 ? 2 : (!pair ? 1 : (N == 64 ? 2 : 4));
  p.g_row_bytes = N >= 64 ? 128u : uint32_t(N) * 2u;
  p.g_box_cols = N >= 64 ? 64u : uint32_t(N);
  p.g_rows = (pair || pair128) ? 136u : 128u;
  p.n_mma = n_shift * N;
  p.g_lbo = (pair || pair128) ? p.g_row_bytes : 128u * p.g_row_bytes;
  p.Ncols = p.n_mma < 32 ? 32 : p.n_mma;
  p.g_bytes = p.g_rows * uint32_t(N) * 2u;
  // accumulator table
  int na = 0;
  if (pair128) {
    for (int r = 0; r < 3; ++r)
      for (int half = 0; half < 2; ++half) {
        WgAcc& A = p.acc[na++];
        for (int i = 0; i < 16; ++i) (&A.tap[0][0])[i] = -1;
        A.row_off = p.halo_lo + (r - 1) * Wp3 + half;   // xa = base | base + 1
        A.chunk = 0;
        A.lbo = p.x_chunk_bytes;                        // M block 1 = the second 64-channel chunk
        A.cch[0] = 0; A.cch[1] = 1;
        for (int b = 0; b < 2; ++b) {
          A.tap[b][0] = 3 * r + (half ? 2 : 1);         // gb = 0: offset xa
          A.tap[b][1] = half ? -1 : 3 * r + 0;          // gb = 1: offset xa - 1 (xa = base + 1: the centre tap again, not stored)
        }
      }
  } else if (pair && N == 64) {
    for (int r = 0; r < 3; ++r) {
      WgAcc& A = p.acc[na++];
      for (int i = 0; i < 16; ++i) (&A.tap[0][0])[i] = -1;
      const int base = (r - 1) * Wp3;
      A.row_off = p.halo_lo + base - 1;
      A.chunk = 0;
      A.lbo = 2u * 128u;                       // M block 1 = the same window two rows further (base + 1)
      A.tap[0][0] = 3 * r + 0;
      A.tap[1][0] = 3 * r + 2; A.tap[1][1] = 3 * r + 1;
      A.cch[0] = A.cch[1] = 0;
    }
  } else if (pair) {
    // N = 16 (image head): D[(r_blk, c), (i, n)] = sum_p X[p + (r-1)Wp + 1, c] G[p + i, n] = dW at tap offset (r-1)Wp + 1 - i:
    // shifts i = 0, 1, 2 are the taps s = 2, 1, 0 of filter row r; two filter rows per accumulator (M blocks Wp rows apart)
    for (int r0 = 0; r0 < 3; r0 += 2) {
      WgAcc& A = p.acc[na++];
      for (int i = 0; i < 16; ++i) (&A.tap[0][0])[i] = -1;
      A.row_off = p.halo_lo + (r0 - 1) * Wp3 + 1;
      A.chunk = 0;
      A.lbo = r0 + 1 < 3 ? uint32_t(Wp3) * 128u : 128u;   // M block 1 = filter row r0 + 1 (r0 = 2: padding, one row further)
      for (int i = 0; i < 3; ++i) {
        A.tap[0][i] = 3 * r0 + (2 - i);
        if (r0 + 1 < 3) A.tap[1][i] = 3 * (r0 + 1) + (2 - i);
      }
      A.cch[0] = A.cch[1] = 0;
    }
  } else if (p.n_chunks == 1) {
    // pair taps: block 1 is the same buffer shifted to another tap (taps sorted by offset so that LBO > 0)
    int order[9];
    for (int t = 0; t < T; ++t) order[t] = t;
    for (int i = 0; i < T; ++i)
      for (int j = i + 1; j < T; ++j)
        if (tap_off[order[j]] < tap_off[order[i]]) { int tmp = order[i]; order[i] = order[j]; order[j] = tmp; }
    for (int i = 0; i < T;) {
      WgAcc& A = p.acc[na++];
      A.row_off = p.halo_lo + tap_off[order[i]];
      A.chunk = 0;
      for (int q2 = 0; q2 < 16; ++q2) (&A.tap[0][0])[q2] = -1;
      A.tap[0][0] = order[i];
      A.cch[0] = A.cch[1] = 0;
      if (i + 1 < T && tap_off[order[i + 1]] > tap_off[order[i]]) {
        A.lbo = uint32_t(tap_off[order[i + 1]] - tap_off[order[i]]) * 128u;
        A.tap[1][0] = order[i + 1];
        i += 2;
      } else {
        A.lbo = 128u;  // padding block: one row further, never stored
        A.tap[1][0] = -1;
        i += 1;
      }
    }
  } else {
    MV_CHECK_ARG(p.n_chunks % 2 == 0, "mv_wgrad: Cin must be 64 or a multiple of 128");
    // chunk pair outermost: the accumulators of a pass then read as few chunks as possible and the pass stages only those
    // (256 -> 128 at 7x7: 18 accumulators in 5 passes; four of them now load 2 of the 4 chunks per 128-row tile)
    for (int c = 0; c < p.n_chunks; c += 2)
      for (int t = 0; t < T; ++t) {
        WgAcc& A = p.acc[na++];
        A.row_off = p.halo_lo + tap_off[t];
        A.chunk = c;
        A.lbo = p.x_chunk_bytes;
        for (int q2 = 0; q2 < 16; ++q2) (&A.tap[0][0])[q2] = -1;
        A.tap[0][0] = t; A.tap[1][0] = t;
        A.cch[0] = c; A.cch[1] = c + 1;
      }
  }
  const int per_pass_max = 512 / p.Ncols > kWgMaxAcc ? kWgMaxAcc : 512 / p.Ncols;
  int passes = (na + per_pass_max - 1) / per_pass_max;
  // Several chunk pairs and more accumulators than TMEM holds: every pass takes the taps of ONE chunk pair, the same number in
  // every pass (256 -> 128 at 7x7: 6 passes of 3 taps instead of 5 passes of 4 / 3 accumulators over all four chunks).  A pass
  // then stages 2 chunks + the G tile per 128-row tile (72 KB instead of 112 KB: the kernel is bound by the L2 -> shared-memory
  // feed, ncu: 41 B/clk/SM), and all passes walk the rows in lock-step, so every tile is read from HBM once (a 5-pass split with
  // unequal passes drifted apart and missed in L2: profiles/r2_ncu_wgrad_b1c0.md).
  const int n_pairs = p.n_chunks >= 2 ? p.n_chunks / 2 : 1;
  const int per_pair = na / n_pairs;
  const bool by_pair = n_pairs > 1 && na > per_pass_max && na == per_pair * n_pairs;
  if (by_pair) {
    const int ppp = (per_pair + per_pass_max - 1) / per_pass_max;   // passes per chunk pair
    passes = ppp * n_pairs;
    MV_CHECK_ARG(passes <= 8, "mv_wgrad: too many passes (%d)", passes);
    int i = 0;
    for (int pr = 0; pr < n_pairs; ++pr)
      for (int j = 0, b = pr * per_pair; j < ppp; ++j, ++i) {
        const int cnt = (pr * per_pair + per_pair - b + (ppp - j) - 1) / (ppp - j);
        p.acc_begin[i] = b;
        p.n_acc[i] = cnt;
        b += cnt;
      }
  } else {
    MV_CHECK_ARG(passes <= 8, "mv_wgrad: too many passes (%d)", passes);
    for (int i = 0, b = 0; i < passes; ++i) {
      const int cnt = (na - b + (passes - i) - 1) / (passes - i);  // balanced split
      p.acc_begin[i] = b;
      p.n_acc[i] = cnt;
      b += cnt;
    }
  }
  int max_acc = 0;
  for (int i = 0; i < passes; ++i) max_acc = p.n_acc[i] > max_acc ? p.n_acc[i] : max_acc;
  int cols = max_acc * p.Ncols;
  p.tmem_cols = cols <= 32 ? 32 : (cols <= 64 ? 64 : (cols <= 128 ? 128 : (cols <= 256 ? 256 : 512)));
  const size_t fixed = 2048 + (2 * kWgMaxStages + 1) * 8 + 16 + 128 * 8 * 4;
  size_t smem = 0;
  for (int i = 0; i < passes; ++i) {
    // the chunks this pass reads: [chunk_lo, chunk_lo + chunk_n); WgAcc::chunk becomes the slot inside the stage
    int c_lo = p.n_chunks, c_hi = 0;
    for (int a = 0; a < p.n_acc[i]; ++a) {
      const WgAcc& A = p.acc[p.acc_begin[i] + a];
      const int last = A.chunk + (A.lbo == p.x_chunk_bytes && p.n_chunks > 1 ? 1 : 0);   // block 1 in the next chunk buffer
      c_lo = A.chunk < c_lo ? A.chunk : c_lo;
      c_hi = last > c_hi ? last : c_hi;
    }
    p.chunk_lo[i] = c_lo;
    p.chunk_n[i] = c_hi - c_lo + 1;
    for (int a = 0; a < p.n_acc[i]; ++a) p.acc[p.acc_begin[i] + a].chunk -= c_lo;
    p.stage_bytes[i] = uint32_t(p.chunk_n[i]) * p.x_chunk_bytes + ((p.g_bytes + 1023u) & ~1023u);
    p.stages[i] = int((kWgSmemLimit - fixed) / p.stage_bytes[i]);
    if (p.stages[i] > kWgMaxStages) p.stages[i] = kWgMaxStages;
    MV_CHECK_ARG(p.stages[i] >= 1, "mv_wgrad: stage does not fit shared memory");
    const size_t need = fixed + size_t(p.stages[i]) * p.stage_bytes[i];
    smem = need > smem ? need : smem;
  }
  CUtensorMap tmX, tmG;
  const bool ok = tc::make_tmap_2d_bf16(&tmX, X, uint64_t(x_rows), uint64_t(Cin), uint64_t(x_ld) * 2, uint32_t(p.box_rows), 64,
                                        CU_TENSOR_MAP_SWIZZLE_128B) &&
                  tc::make_tmap_2d_bf16(&tmG, G, uint64_t(g_rows), uint64_t(N), uint64_t(g_ld) * 2, p.g_rows, p.g_box_cols,
                                        N >= 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B);
  if (!ok) {
    mv::set_error("mv_wgrad: cuTensorMapEncodeTiled failed");
    return MV_ERR_CUDA;
  }
  int gx = num_sms() / passes;
  if (gx > p.n_ktiles) gx = p.n_ktiles;
  if (gx < 1) gx = 1;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kWgSmemLimit));
    attr_done = true;
  }
  wgrad_kernel<<<dim3(gx, passes), kWgThreads, smem, static_cast<cudaStream_t>(stream)>>>(tmX, tmG, p);
  MV_CHECK_LAUNCH("mv_wgrad");
  return MV_OK;
}

extern "C" int mv_wgrad_slice(const void* X, int64_t x_rows, int x_ld, int Cin, const void* G, int64_t g_rows, int g_ld, int N,
                              int T, const int* tap_off, int64_t P, float* dW, int N_total, int n_offset, float* db,
                              void* stream) {
  return wgrad_launch(X, x_rows, x_ld, Cin, G, g_rows, g_ld, N, T, tap_off, P, dW, N_total, n_offset, db, 0, N, stream);
}

extern "C" int mv_wgrad(const void* X, int64_t x_rows, int x_ld, int Cin, const void* G, int64_t g_rows, int g_ld, int N,
                        int T, const int* tap_off, int64_t P, float* dW, float* db, void* stream) {
  return wgrad_launch(X, x_rows, x_ld, Cin, G, g_rows, g_ld, N, T, tap_off, P, dW, N, 0, db, 0, N, stream);
}

extern "C" int mv_wgrad_nct(const void* X, int64_t x_rows, int x_ld, int Cin, const void* G, int64_t g_rows, int g_ld, int N,
                            int T, const int* tap_off, int64_t P, float* dW, int n_offset, int n_valid, float* db,
                            void* stream) {
  MV_CHECK_ARG(n_valid >= 1 && n_valid <= N, "mv_wgrad_nct: 1 <= n_valid <= N");
  return wgrad_launch(X, x_rows, x_ld, Cin, G, g_rows, g_ld, N, T, tap_off, P, dW, n_offset + N, n_offset, db, 1, n_valid, stream);
}
