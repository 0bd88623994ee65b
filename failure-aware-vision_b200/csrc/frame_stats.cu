// frame_stats.cu -- f1: the reference's real per-frame pixel path as one fused integer kernel.
//
// Replaces (reference): the arithmetic of SignalAnalyzer.analyze_frame, platform/backend/signal_analyzer.py:62-105:
//   :62  gray = cv2.cvtColor(frame, COLOR_BGR2GRAY)          -> (B*3735 + G*19235 + R*9798 + 2^14) >> 15
//   :65  cv2.Laplacian(gray, CV_64F).var()                   -> 3x3 [[0,1,0],[1,-4,1],[0,1,0]], BORDER_REFLECT_101;
//                                                              we return sum and sum of squares (exact integers)
//   :70  np.mean(gray)                                       -> sum of gray
//   :77-78 cv2.absdiff(prev_gray, gray) mean                 -> sum |gray - prev|
//   :101 cv2.calcHist 256 bins                               -> hist[256]
// The host finishes var / mean / entropy in fp64 exactly as the reference does (gate.py).  One read of the
// frame (3 B/px) + previous gray (1 B/px) and one write of the new gray (1 B/px): HBM-bound, 5 B/px.
#include "common.cuh"

namespace fav {

constexpr int FS_ROWS = 4;

__device__ __forceinline__ int reflect101(int i, int n) {
  if (n == 1) return 0;
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  if (i < 0) i = -i;                         // n == 2: one more fold (rows -1 .. n+FS_ROWS of a 2-row frame)
  if (i >= n) i = 2 * (n - 1) - i;
  return min(max(i, 0), n - 1);
}

__global__ void __launch_bounds__(256) k_frame_stats(const uint8_t* __restrict__ frame, uint8_t* __restrict__ prev_gray,
                                                     int H, int W, int first_frame, unsigned long long* __restrict__ out) {
  extern __shared__ uint8_t s_gray[];                    // (FS_ROWS + 2) x W
  __shared__ unsigned s_hist[256];
  __shared__ long long s_red[4][8];
  const int y0 = blockIdx.x * FS_ROWS;
  s_hist[threadIdx.x] = 0;
  for (int i = threadIdx.x; i < (FS_ROWS + 2) * W; i += blockDim.x) {
    const int ry = i / W, x = i - ry * W;
    const int y = reflect101(y0 - 1 + ry, H);
    const uint8_t* p = frame + ((size_t)y * W + x) * 3;
    s_gray[i] = uint8_t((p[0] * 3735u + p[1] * 19235u + p[2] * 9798u + 16384u) >> 15);
  }
  __syncthreads();
  long long s_lap = 0, s_lap2 = 0, s_g = 0, s_d = 0;
  for (int i = threadIdx.x; i < FS_ROWS * W; i += blockDim.x) {
    const int ry = i / W, x = i - ry * W, y = y0 + ry;
    if (y >= H) break;
    const uint8_t* row = s_gray + (ry + 1) * W;
    const int c = row[x];
    const int lap = int(row[x - W]) + int(row[x + W]) + int(row[reflect101(x - 1, W)]) + int(row[reflect101(x + 1, W)]) - 4 * c;
    s_lap += lap; s_lap2 += (long long)lap * lap; s_g += c;
    const size_t gi = (size_t)y * W + x;
    if (!first_frame) s_d += abs(c - int(prev_gray[gi]));
    prev_gray[gi] = uint8_t(c);
    atomicAdd(&s_hist[c], 1u);
  }
  long long v[4] = {s_lap, s_lap2, s_g, s_d};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    if ((threadIdx.x & 31) == 0) s_red[k][threadIdx.x >> 5] = v[k];
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    long long t = 0;
    for (int w = 0; w < 8; ++w) t += s_red[threadIdx.x][w];
    atomicAdd(&out[threadIdx.x], (unsigned long long)t);     // two's complement: signed sums add correctly
  }
  const unsigned hv = s_hist[threadIdx.x];
  if (hv) atomicAdd(&out[4 + threadIdx.x], (unsigned long long)hv);
}

}  // namespace fav

using namespace fav;

extern "C" int fav_frame_stats(fav_handle h, const uint8_t* d_frame, uint8_t* d_prev_gray, int height, int width,
                               int first_frame, int64_t* d_out, void* stream) {
  FAV_REQUIRE(h && d_frame && d_prev_gray && d_out, "fav_frame_stats: null pointer");
  FAV_REQUIRE(height >= 2 && width >= 2 && width <= 8192, "fav_frame_stats: frame must be at least 2x2 and at most 8192 wide");
  FAV_DEVICE(h);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  FAV_CUDA_OK(cudaMemsetAsync(d_out, 0, (4 + 256) * sizeof(int64_t), st));
  const size_t smem = (size_t)(FS_ROWS + 2) * width;
  if (smem > 48 * 1024) FAV_CUDA_OK(cudaFuncSetAttribute(k_frame_stats, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  k_frame_stats<<<(height + FS_ROWS - 1) / FS_ROWS, 256, smem, st>>>(d_frame, d_prev_gray, height, width, first_frame,
                                                                    reinterpret_cast<unsigned long long*>(d_out));
  h->launches++;
  FAV_CUDA_OK(cudaGetLastError());
  return FAV_OK;
}
