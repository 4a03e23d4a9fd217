// Events -> voxel grid on the GPU: the on-disk -> tensor step in front of the hot path.
//   DSEC : VoxelGrid.convert, DSEC/dataset/representations.py:15-55  (trilinear x/y/t scatter, `put_` accumulate)
//   DDD17: generate_voxel_grid, datasets/data_util.py:54-126          (integer x/y, bilinear in time, +/- grids)
// One thread per event, up to 8 (resp. 2) float atomics into a grid that stays L2-resident; HBM-bound on
// the event stream (16-20 B/event).  Atomic accumulation order is not fixed: results agree with the
// reference to fp32 summation-order noise, not bit for bit.
#include "common.cuh"

namespace {

__global__ void voxel_dsec_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                  const float* __restrict__ pol, const float* __restrict__ t, long long n, int C, int H,
                                  int W, float* __restrict__ grid) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float t0 = t[0], t1 = t[n - 1];
  const float tn = (float)(C - 1) * (t[i] - t0) / (t1 - t0);          // representations.py:25-26
  const float xv = x[i], yv = y[i];
  const int x0 = (int)xv, y0 = (int)yv, tt0 = (int)tn;                 // .int() truncates toward zero (:28-30)
  const float value = 2.f * pol[i] - 1.f;                              // :32
#pragma unroll
  for (int dx = 0; dx < 2; ++dx)
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dt = 0; dt < 2; ++dt) {
        const int xl = x0 + dx, yl = y0 + dy, tl = tt0 + dt;
        if (xl < W && xl >= 0 && yl < H && yl >= 0 && tl >= 0 && tl < C) {   // :37
          const float w = value * (1.f - fabsf((float)xl - xv)) * (1.f - fabsf((float)yl - yv)) *
                          (1.f - fabsf((float)tl - tn));                     // :38
          atomicAdd(grid + ((size_t)tl * H + yl) * W + xl, w);               // :40-44
        }
      }
}

__global__ void voxel_ddd17_kernel(const double* __restrict__ ev, long long n, int C, int H, int W, int separate_pol,
                                   float* __restrict__ grid) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double first = ev[2], last = ev[(n - 1) * 4 + 2];
  double dT = last - first;                                            // data_util.py:71-76
  if (dT == 0.0) dT = 1.0;
  const int xs = (int)ev[i * 4 + 0], ys = (int)ev[i * 4 + 1];          // :79-80
  const double ts = (double)(C - 1) * (ev[i * 4 + 2] - first) / dT;    // :83
  const bool positive = ev[i * 4 + 3] == 1.0;                          // :85-86, :91 (0 -> -1)
  const int ti = (int)ts;                                              // :88
  const double dts = ts - (double)ti;
  const float vl = (float)(1.0 - dts), vr = (float)dts;                // :90-91 (|pol| = 1)
  if (!(xs < W && xs >= 0 && ys < H && ys >= 0 && ts >= 0.0 && ts < (double)C)) return;   // :95
  // separate_pol: channels [0,C) positive, [C,2C) negative; else positive - negative (:119-125)
  float* g = grid;
  float sign = 1.f;
  if (!positive) {
    if (separate_pol) g = grid + (size_t)C * H * W;
    else sign = -1.f;
  }
  if (ti < C) atomicAdd(g + ((size_t)ti * H + ys) * W + xs, sign * vl);             // :94-99
  if (ti + 1 < C) atomicAdd(g + ((size_t)(ti + 1) * H + ys) * W + xs, sign * vr);   // :101-104
}

}  // namespace

extern "C" int essb_voxel_grid_dsec(const float* x, const float* y, const float* pol, const float* t, int64_t n, int C,
                                    int H, int W, float* grid, void* stream) {
  ESSB_REQUIRE(x && y && pol && t && grid && n > 0 && C > 0 && H > 0 && W > 0, "essb_voxel_grid_dsec: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(grid, 0, sizeof(float) * (size_t)C * H * W, st) != cudaSuccess) {
    essb_set_error("essb_voxel_grid_dsec: memset failed");
    return ESSB_ERR_LAUNCH;
  }
  voxel_dsec_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, y, pol, t, n, C, H, W, grid);
  ESSB_LAUNCH_CHECK("essb_voxel_grid_dsec");
  return ESSB_OK;
}

extern "C" int essb_voxel_grid_ddd17(const double* events, int64_t n, int C, int H, int W, int separate_pol, float* grid,
                                     void* stream) {
  ESSB_REQUIRE(events && grid && n > 0 && C > 0 && H > 0 && W > 0, "essb_voxel_grid_ddd17: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t cells = (size_t)(separate_pol ? 2 : 1) * C * H * W;
  if (cudaMemsetAsync(grid, 0, sizeof(float) * cells, st) != cudaSuccess) {
    essb_set_error("essb_voxel_grid_ddd17: memset failed");
    return ESSB_ERR_LAUNCH;
  }
  voxel_ddd17_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(events, n, C, H, W, separate_pol, grid);
  ESSB_LAUNCH_CHECK("essb_voxel_grid_ddd17");
  return ESSB_OK;
}
