// mma_rate.cu -- how many cycles one tcgen05.mma (M=128, bf16, K=16, operands in SWIZZLE_128B shared memory) takes as a
// function of N and of the A operand's start alignment inside the 1024-byte swizzle atom.  One CTA, one issuing thread,
// `iters` back-to-back MMAs on zeroed shared memory, timed with clock64 from first issue to commit completion.
// usage: mma_rate <N> <a_offset_bytes (multiple of 128)> [iters] [ctas]
#include "../../failure-aware-vision_b200/csrc/tc_ptx.cuh"
#include <cuda_runtime.h>
#include <cstdlib>
using namespace fav;

__global__ void __launch_bounds__(128, 1) k(int N, int a_off, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tslot;
  for (uint32_t i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0;
  const uint32_t bb = smem_u32(&bar);
  if (threadIdx.x == 0) { mbar_init(bb, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(smem_u32(&tslot), 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tslot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(N >> 3) << 17) | (uint32_t(128 >> 4) << 24);
    const uint64_t da = make_sw128_desc(base + a_off), db = make_sw128_desc(base + 64 * 1024);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) umma_f16(tmem + (i & 1) * 256, da + 2u * kk, db + 2u * kk, idesc, 1u);
    }
    umma_commit(bb);
    mbar_wait(bb, 0);
    out[blockIdx.x] = clock64() - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main(int argc, char** argv) {
  const int N = atoi(argv[1]), a_off = atoi(argv[2]), iters = argc > 3 ? atoi(argv[3]) : 2000, ctas = argc > 4 ? atoi(argv[4]) : 1;
  long long* d;
  cudaMalloc(&d, ctas * sizeof(long long));
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int rep = 0; rep < 2; ++rep) k<<<ctas, 128, 170 * 1024>>>(N, a_off, iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("failed: %s\n", cudaGetErrorString(e)); return 1; }
  long long h[256];
  cudaMemcpy(h, d, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
  const double cyc = double(h[0]) / (4.0 * iters);
  printf("N=%3d a_off=%5d ctas=%3d : %.1f cycles per MMA (M128 x N x K16) -> %.0f%% of 8192 FLOP/clk\n", N, a_off, ctas, cyc,
         100.0 * (2.0 * 128 * N * 16 / cyc) / 8192.0);
  return 0;
}
