// tma_probe.cu -- standalone check of what a tiled tensor map accepts (overlapping strides, narrow inner boxes).
// usage: tma_probe <swizzle 0..3> <rank> dims... strides_bytes(rank-1)... box... coords...
// Loads one box into shared memory, copies it out and compares every element against the address arithmetic.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe(const __grid_constant__ CUtensorMap tm, int rank, int c0, int c1, int c2, int c3, int c4, uint32_t bytes,
                      uint16_t* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  const uint32_t sb = (uint32_t)__cvta_generic_to_shared(smem), bb = (uint32_t)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bb));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bb), "r"(bytes) : "memory");
    if (rank == 5)
      asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                   ::"r"(sb), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(bb), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
    else if (rank == 4)
      asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                   ::"r"(sb), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(bb), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
    else if (rank == 3)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                   ::"r"(sb), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(bb), "r"(c0), "r"(c1), "r"(c2) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(sb), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(bb), "r"(c0), "r"(c1) : "memory");
  }
  uint32_t done = 0;
  long long t0 = clock64();
  while (!done) {
    asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p;}" : "=r"(done) : "r"(bb) : "memory");
    if (clock64() - t0 > 2000000000ll) { if (threadIdx.x == 0) printf("probe: timeout\n"); break; }
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < bytes / 2; i += blockDim.x) out[i] = reinterpret_cast<uint16_t*>(smem)[i];
}

int main(int argc, char** argv) {
  if (argc < 3) { printf("usage\n"); return 2; }
  int sw = atoi(argv[1]), rank = atoi(argv[2]);
  if (argc != 3 + rank + (rank - 1) + rank + rank) { printf("bad arg count\n"); return 2; }
  cuuint64_t dims[5] = {1, 1, 1, 1, 1}, strides[4] = {0, 0, 0, 0};
  cuuint32_t box[5] = {1, 1, 1, 1, 1}, estr[5] = {1, 1, 1, 1, 1};
  int co[5] = {0, 0, 0, 0, 0};
  int p = 3;
  for (int i = 0; i < rank; ++i) dims[i] = strtoull(argv[p++], 0, 10);
  for (int i = 0; i < rank - 1; ++i) strides[i] = strtoull(argv[p++], 0, 10);
  for (int i = 0; i < rank; ++i) box[i] = atoi(argv[p++]);
  for (int i = 0; i < rank; ++i) co[i] = atoi(argv[p++]);
  const size_t N = 8u << 20;                     // 16 MB of u16; value = element index mod 65521 (prime: no aliasing at 64 KB)
  std::vector<uint16_t> hx(N);
  for (size_t i = 0; i < N; ++i) hx[i] = uint16_t(i % 65521);
  uint16_t *dx, *dout;
  cudaMalloc(&dx, N * 2);
  cudaMemcpy(dx, hx.data(), N * 2, cudaMemcpyHostToDevice);
  size_t elems = 1;
  for (int i = 0; i < rank; ++i) elems *= box[i];
  cudaMalloc(&dout, elems * 2);
  cudaMemset(dout, 0xFF, elems * 2);
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  CUtensorMap tm;
  CUresult r = ((EncodeFn)fp)(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, dx, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              (CUtensorMapSwizzle)sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed: %d\n", int(r)); return 1; }
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  probe<<<1, 128, elems * 2 + 1024>>>(tm, rank, co[0], co[1], co[2], co[3], co[4], uint32_t(elems * 2), dout);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<uint16_t> ho(elems);
  cudaMemcpy(ho.data(), dout, elems * 2, cudaMemcpyDeviceToHost);
  // expected: dense box order, 16-byte chunk index XORed with (128-byte row index % span) for the swizzled modes
  size_t bad = 0, oob = 0;
  const int span = sw == 3 ? 8 : sw == 2 ? 4 : sw == 1 ? 2 : 1;
  for (size_t i = 0; i < elems; ++i) {
    size_t rem = i; long long byte = 0; bool in = true;
    for (int d = 0; d < rank; ++d) {
      const long long c = co[d] + (long long)(rem % box[d]); rem /= box[d];
      if (c < 0 || c >= (long long)dims[d]) in = false;
      byte += c * (d == 0 ? 2 : (long long)strides[d - 1]);
    }
    size_t off = i * 2;
    const size_t chunk = (off >> 4) & 7, row = (off >> 7);
    off = (off & ~size_t(0x70)) | (((chunk ^ (row % span)) & 7) << 4);
    if (span == 1) off = i * 2;
    const uint16_t want = in ? hx[byte / 2] : 0, got = ho[off / 2];
    if (!in) ++oob;
    if (want != got) { if (bad < 5) printf("  mismatch at box elem %zu: want %u got %u\n", i, want, got); ++bad; }
  }
  printf("ok: %zu elems, %zu out-of-bounds, %zu mismatches\n", elems, oob, bad);
  return bad ? 1 : 0;
}
