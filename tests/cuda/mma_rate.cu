// tcgen05.mma issue-rate microbenchmark (run on a B200): cycles per MMA for SS-mode bf16 MMAs, M = 128, as a
// function of N, operand major-ness, accumulator dependence and how the A descriptor walks shared memory.
// One CTA per SM, garbage operands (shared memory is zero-filled), a long back-to-back MMA sequence.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mma_rate tests/cuda/mma_rate.cu -lcuda
#include <cstdio>
#include <cstdlib>

#include "../../multivae_b200/csrc/tc.cuh"

using namespace tc;

struct Cfg {
  int N, a_mn, b_mn, alt_acc, walk, iters;  // walk: 0 = same A address, 1 = conv-like (9 row shifts x 4 k-steps), 2 = MN k-walk
  int sw;                                   // K-major only: 0 = 128-byte rows SWIZZLE_128B, 1 = 32-byte rows SWIZZLE_32B (a 16-channel
                                            // layer: one K step per tap), 2 = two 16-byte column panels, no swizzle
};

__global__ void __launch_bounds__(128) rate_kernel(Cfg c, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) tmem_alloc(&tmem_base, 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  const uint32_t tm = tmem_base;
  if (threadIdx.x < 32) {
    const uint32_t id = idesc_bf16(128, c.N, c.a_mn, c.b_mn);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 64 * 1024);
    long long t0 = 0, t1 = 0;
    // all 36 descriptor pairs of one "tile" are built before the timed loop (the issue loop is then 2 moves + 1 MMA)
    uint32_t alo[36], blo[36];
    uint32_t ahi = 0, bhi = 0;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint32_t aa = a0, bb = b0;
        uint64_t ad, bd;
        if (c.a_mn) {
          aa = a0 + (c.walk ? uint32_t(t * 3 + k * 16) * 128u : 0u);
          ad = smem_desc(aa, 64 * 128, 1024, SW_128);
        } else {
          aa = a0 + (c.walk ? uint32_t(t * 29) * 128u + k * 32 : 0u);
          ad = smem_desc(aa, 16, 1024, SW_128);
          if (c.sw == 1) ad = smem_desc(a0 + (c.walk ? uint32_t(t * 29 + k * 3) * 32u : 0u), 16, 256, SW_32);
          if (c.sw == 2) ad = smem_desc(a0 + (c.walk ? uint32_t(t * 29 + k * 3) * 16u : 0u), 192 * 16, 128, SW_NONE);
        }
        if (c.b_mn) {
          bb = b0 + (c.walk ? uint32_t(k * 16) * 128u : 0u);
          bd = smem_desc(bb, 128 * 128, 1024, SW_128);
        } else {
          bb = b0 + (c.walk ? uint32_t(t) * uint32_t(c.N) * 128u + k * 32 : 0u);
          if (bb + c.N * 128 > b0 + 136 * 1024) bb = b0 + k * 32;
          bd = smem_desc(bb, 16, 1024, SW_128);
          if (c.sw == 1) bd = smem_desc(b0 + uint32_t(t) * uint32_t(c.N) * 32u, 16, 256, SW_32);
          if (c.sw == 2) bd = smem_desc(b0 + uint32_t(t) * uint32_t(c.N) * 32u, uint32_t(c.N) * 16u, 128, SW_NONE);
        }
        alo[t * 4 + k] = uint32_t(ad); ahi = uint32_t(ad >> 32);
        blo[t * 4 + k] = uint32_t(bd); bhi = uint32_t(bd >> 32);
      }
    }
    if (elect_one()) {
      t0 = clock64();
      for (int it = 0; it < c.iters; ++it) {
        const uint32_t acc = c.alt_acc ? uint32_t(it & 1) * 256u : 0u;
#pragma unroll
        for (int j = 0; j < 36; ++j)
          umma_bf16(tm + acc, (uint64_t(ahi) << 32) | alo[j], (uint64_t(bhi) << 32) | blo[j], id, 1);
      }
      umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    t1 = clock64();
    if (elect_one() && blockIdx.x == 0) out[0] = t1 - t0;
    (void)t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024 + 1024);
  const int iters = 200;
  struct { const char* name; Cfg c; } cases[] = {
      {"K-major  N=64  same-addr", {64, 0, 0, 0, 0, iters}},
      {"K-major  N=64  conv-walk", {64, 0, 0, 0, 1, iters}},
      {"K-major  N=64  conv-walk alt-acc", {64, 0, 0, 1, 1, iters}},
      {"K-major  N=128 conv-walk", {128, 0, 0, 0, 1, iters}},
      {"K-major  N=192 conv-walk", {192, 0, 0, 0, 1, iters}},
      {"K-major  N=256 conv-walk", {256, 0, 0, 0, 1, iters}},
      {"K-major  N=256 same-addr", {256, 0, 0, 0, 0, iters}},
      {"K-major  N=16  conv-walk", {16, 0, 0, 0, 1, iters}},
      {"K-major  N=32  conv-walk", {32, 0, 0, 0, 1, iters}},
      {"MN-major N=64  walk", {64, 1, 1, 0, 1, iters}},
      {"MN-major N=128 walk", {128, 1, 1, 0, 1, iters}},
      {"MN-major N=256 walk", {256, 1, 1, 0, 1, iters}},
      {"A MN / B K-major N=64", {64, 1, 0, 0, 1, iters}},
      {"A K / B MN-major N=64", {64, 0, 1, 0, 1, iters}},
      {"A K / B MN-major N=192", {192, 0, 1, 0, 1, iters}},
      {"K-major  N=64  32-byte rows SW_32", {64, 0, 0, 0, 1, iters, 1}},
      {"K-major  N=64  32-byte rows SW_32 alt-acc", {64, 0, 0, 1, 1, iters, 1}},
      {"K-major  N=64  16-byte panels no swizzle", {64, 0, 0, 0, 1, iters, 2}},
      {"K-major  N=128 32-byte rows SW_32", {128, 0, 0, 0, 1, iters, 1}},
      {"K-major  N=128 16-byte panels no swizzle", {128, 0, 0, 0, 1, iters, 2}},
  };
  for (auto& cs : cases) {
    for (int grid : {1, 148}) {
      cudaMemset(d, 0, 8);
      rate_kernel<<<grid, 128, 201 * 1024 + 1024>>>(cs.c, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long cyc = 0;
      cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
      printf("%-36s grid %3d: %8.1f cycles/MMA   (%s)\n", cs.name, grid, double(cyc) / (cs.c.iters * 36.0), cudaGetErrorString(e));
    }
  }
  return 0;
}
