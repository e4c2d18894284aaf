// Probe of the tcgen05 shared-memory-descriptor conventions the conv kernels rely on (run on a B200):
//   T1  K-major SWIZZLE_128B operands whose start address is shifted by an arbitrary number of 128-byte rows
//   T2  MN-major SWIZZLE_128B operands (reduction over rows) with row shifts and an LBO that aliases a shifted window
//   T3  K-major SWIZZLE_32B (16-channel rows)      T4  MN-major SWIZZLE_32B B operand
// Prints max |err| against a CPU reference for each (test, shift, base_offset mode).
#include <cuda_bf16.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../multivae_b200/csrc/tc.cuh"

using namespace tc;

struct Params {
  int test, shift, bo_mode, d;  // d: row distance of the second M block (T2/T4)
};

__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                             float* __restrict__ out, Params prm) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;             // up to 192 rows x 128 B
  uint8_t* sB = smem + 32768;     // up to 128 rows x 128 B
  __shared__ uint64_t bar_full, bar_mma;
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5;
  const bool t3 = prm.test == 3, t4 = prm.test == 4;
  const uint32_t rowA = t3 ? 32u : 128u;           // bytes per row of A in smem
  const uint32_t rowB = (t3 || t4) ? 32u : 128u;   // bytes per row of B
  const uint32_t rowsA = 192, rowsB = (prm.test == 1) ? 64u : (t3 ? 16u : 128u);
  if (threadIdx.x == 0) {
    mbar_init(&bar_full, 1);
    mbar_init(&bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base, 64);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tm = tmem_base;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar_full, rowsA * rowA + rowsB * rowB);
    tma_load_2d(sA, &tmA, &bar_full, 0, 0);
    tma_load_2d(sB, &tmB, &bar_full, 0, 0);
    mbar_wait(&bar_full, 0);
    fence_after_sync();
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    auto bo = [&](uint32_t addr) { return prm.bo_mode ? ((addr >> 7) & 7u) : 0u; };
    if (prm.test == 1) {
      const uint32_t id = idesc_bf16(128, 64, 0, 0);
      for (int k = 0; k < 4; ++k) {
        uint32_t aa = a0 + prm.shift * 128 + k * 32, bb = b0 + k * 32;
        umma_bf16(tm, smem_desc(aa, 16, 1024, SW_128, bo(aa)), smem_desc(bb, 16, 1024, SW_128, 0), id, k > 0);
      }
    } else if (prm.test == 2) {
      const uint32_t id = idesc_bf16(128, 64, 1, 1);
      for (int k = 0; k < 8; ++k) {
        uint32_t aa = a0 + (prm.shift + k * 16) * 128, bb = b0 + k * 16 * 128;
        umma_bf16(tm, smem_desc(aa, prm.d * 128, 1024, SW_128, bo(aa)), smem_desc(bb, 8192, 1024, SW_128, 0), id, k > 0);
      }
    } else if (prm.test == 3) {
      const uint32_t id = idesc_bf16(128, 16, 0, 0);
      uint32_t aa = a0 + prm.shift * 32;
      umma_bf16(tm, smem_desc(aa, 16, 256, SW_32, prm.bo_mode ? ((aa >> 5) & 7u) : 0u), smem_desc(b0, 16, 256, SW_32, 0), id, 0);
    } else {
      const uint32_t id = idesc_bf16(128, 16, 1, 1);
      for (int k = 0; k < 8; ++k) {
        uint32_t aa = a0 + (prm.shift + k * 16) * 128, bb = b0 + k * 16 * 32;
        umma_bf16(tm, smem_desc(aa, prm.d * 128, 1024, SW_128, bo(aa)), smem_desc(bb, 4096, 256, SW_32, 0), id, k > 0);
      }
    }
    umma_commit(&bar_mma);
  }
  __syncwarp();
  mbar_wait(&bar_mma, 0);
  fence_after_sync();
  const int N = (t3 || t4) ? 16 : 64;
  const int row = warp * 32 + (threadIdx.x & 31);
  uint32_t v[32];
  for (int c0 = 0; c0 < N; c0 += 32) {
    if (N - c0 >= 32) tmem_ld_32x32(tm + (uint32_t(warp * 32) << 16) + c0, v);
    else tmem_ld_32x16(tm + (uint32_t(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 32 && c0 + j < N; ++j) out[row * N + c0 + j] = __uint_as_float(v[j]);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 64);
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main() {
  const int RA = 192;
  std::vector<float> A(RA * 64), Bm(128 * 64);
  srand(1);
  for (auto& x : A) x = bf((rand() % 2001 - 1000) / 1000.f);
  for (auto& x : Bm) x = bf((rand() % 2001 - 1000) / 1000.f);
  std::vector<__nv_bfloat16> hA(RA * 64), hB(128 * 64), hA16(RA * 16), hB16(128 * 16);
  for (int i = 0; i < RA * 64; ++i) hA[i] = __float2bfloat16(A[i]);
  for (int i = 0; i < 128 * 64; ++i) hB[i] = __float2bfloat16(Bm[i]);
  for (int r = 0; r < RA; ++r) for (int c = 0; c < 16; ++c) hA16[r * 16 + c] = hA[r * 64 + c];
  for (int r = 0; r < 128; ++r) for (int c = 0; c < 16; ++c) hB16[r * 16 + c] = hB[r * 64 + c];
  __nv_bfloat16 *dA, *dB, *dA16, *dB16;
  float* dO;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dA16, hA16.size() * 2); cudaMalloc(&dB16, hB16.size() * 2);
  cudaMalloc(&dO, 128 * 64 * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dA16, hA16.data(), hA16.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB16, hB16.data(), hB16.size() * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  int fails = 0;
  for (int test = 1; test <= 4; ++test) {
    CUtensorMap tA, tB;
    bool ok = true;
    if (test == 1) {
      ok &= make_tmap_2d_bf16(&tA, dA, RA, 64, 128, RA, 64, CU_TENSOR_MAP_SWIZZLE_128B);
      ok &= make_tmap_2d_bf16(&tB, dB, 64, 64, 128, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B);
    } else if (test == 2) {
      ok &= make_tmap_2d_bf16(&tA, dA, RA, 64, 128, RA, 64, CU_TENSOR_MAP_SWIZZLE_128B);
      ok &= make_tmap_2d_bf16(&tB, dB, 128, 64, 128, 128, 64, CU_TENSOR_MAP_SWIZZLE_128B);
    } else if (test == 3) {
      ok &= make_tmap_2d_bf16(&tA, dA16, RA, 16, 32, RA, 16, CU_TENSOR_MAP_SWIZZLE_32B);
      ok &= make_tmap_2d_bf16(&tB, dB16, 16, 16, 32, 16, 16, CU_TENSOR_MAP_SWIZZLE_32B);
    } else {
      ok &= make_tmap_2d_bf16(&tA, dA, RA, 64, 128, RA, 64, CU_TENSOR_MAP_SWIZZLE_128B);
      ok &= make_tmap_2d_bf16(&tB, dB16, 128, 16, 32, 128, 16, CU_TENSOR_MAP_SWIZZLE_32B);
    }
    if (!ok) { printf("tensor map creation failed (test %d)\n", test); return 2; }
    const int shifts[] = {0, 1, 3, 8, 13, 29};
    for (int bo_mode = 0; bo_mode < 2; ++bo_mode)
      for (int s : shifts) {
        Params prm{test, s, bo_mode, 31};
        const int N = (test >= 3) ? 16 : 64;
        cudaMemset(dO, 0, 128 * 64 * 4);
        probe<<<1, 128, 65536>>>(tA, tB, dO, prm);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("test %d shift %d bo %d: CUDA error %s\n", test, s, bo_mode, cudaGetErrorString(e)); return 3; }
        std::vector<float> O(128 * N);
        cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < N; ++n) {
            double ref = 0;
            if (test == 1) { for (int k = 0; k < 64; ++k) ref += A[(m + s) * 64 + k] * Bm[n * 64 + k]; }
            else if (test == 3) { for (int k = 0; k < 16; ++k) ref += A[(m + s) * 64 + k] * Bm[n * 64 + k]; }
            else { int j = m / 64, c = m % 64; for (int p = 0; p < 128; ++p) ref += A[(p + s + j * prm.d) * 64 + c] * Bm[p * 64 + n]; }
            maxerr = fmax(maxerr, fabs(ref - O[m * N + n]));
          }
        printf("test %d shift %2d base_offset_mode %d : max_err %.5f %s\n", test, s, bo_mode, maxerr, maxerr < 1e-2 ? "OK" : "MISMATCH");
        if (maxerr >= 1e-2) ++fails;
      }
  }
  printf("mismatching cases: %d\n", fails);
  return 0;
}
