// Single-input-channel k x k convolution (k = 7) and its weight gradient: the ResNet-18 stem of StyleEncoderE2VID
// (models/style_networks.py:117-121: conv1 = Conv2d(1, 64, 7, stride 2, padding 3, bias=False) applied to the
// grayscale image / the E2VID reconstruction in the UDA step, training/ess_trainer.py:159-162,282).
//
// With Cin = 1 the layer is 49 MACs per output value on a 1-channel image: HBM-bound (it writes 64 channels
// per pixel), useless for tensor cores.  Mapping: a block owns an 8 x 32 output tile, the input patch it needs
// sits in shared memory; a warp walks one tile row, lane = output channel pair (co = lane, lane + 32), so the
// 49 weights (forward) or the 49 x 2 gradient accumulators (wgrad) of a lane live in registers, the image value
// of a tap is one broadcast shared-memory read, and every global access is a coalesced 128-byte row.
#include "common.cuh"

namespace {

constexpr int ST_THREADS = 256;
constexpr int ST_TH = 8, ST_TW = 32;            // output tile (rows = warps, columns = pixels walked by a warp)

template <int KSZ, int STRIDE>
struct StemGeom {
  static constexpr int PH = (ST_TH - 1) * STRIDE + KSZ;
  static constexpr int PW = (ST_TW - 1) * STRIDE + KSZ;
};

// input patch of tile (tx, ty) of sample n -> shared memory (zero outside the image = the conv's zero padding)
template <int KSZ, int STRIDE>
__device__ __forceinline__ void stem_load_patch(const float* __restrict__ x, int n, int H, int W, int ty, int tx, int pad,
                                                float* patch) {
  using G = StemGeom<KSZ, STRIDE>;
  const int iy0 = ty * ST_TH * STRIDE - pad, ix0 = tx * ST_TW * STRIDE - pad;
  for (int i = threadIdx.x; i < G::PH * G::PW; i += ST_THREADS) {
    const int r = i / G::PW, c = i - r * G::PW;
    const int iy = iy0 + r, ix = ix0 + c;
    patch[i] = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? x[((size_t)n * H + iy) * W + ix] : 0.f;
  }
}

// out[n, oy, ox, co] = sum_{ky,kx} x[n, oy*S - pad + ky, ox*S - pad + kx] * w[co][ky][kx]      (Cout = 64)
template <int KSZ, int STRIDE>
__global__ void __launch_bounds__(ST_THREADS) stem_conv_fwd_kernel(const float* __restrict__ x,
                                                                   const float* __restrict__ w, float* __restrict__ out,
                                                                   int H, int W, int OH, int OW, int pad, int tiles_x,
                                                                   int tiles_y, int n_tiles) {
  using G = StemGeom<KSZ, STRIDE>;
  constexpr int KK = KSZ * KSZ;
  __shared__ float patch[G::PH * G::PW];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float w0[KK], w1[KK];
#pragma unroll
  for (int t = 0; t < KK; ++t) {
    w0[t] = w[lane * KK + t];
    w1[t] = w[(lane + 32) * KK + t];
  }
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    int r = tile;
    const int tx = r % tiles_x; r /= tiles_x;
    const int ty = r % tiles_y;
    const int n = r / tiles_y;
    __syncthreads();
    stem_load_patch<KSZ, STRIDE>(x, n, H, W, ty, tx, pad, patch);
    __syncthreads();
    const int oy = ty * ST_TH + warp;
    if (oy >= OH) continue;
    const float* prow = patch + warp * STRIDE * G::PW;
    float* orow = out + (((size_t)n * OH + oy) * OW + (size_t)tx * ST_TW) * 64;
    const int npx = min(ST_TW, OW - tx * ST_TW);
    for (int px = 0; px < npx; px += 2) {             // two pixels per pass: independent FMA chains
      float a00 = 0.f, a01 = 0.f, a10 = 0.f, a11 = 0.f;
#pragma unroll
      for (int ky = 0; ky < KSZ; ++ky)
#pragma unroll
        for (int kx = 0; kx < KSZ; ++kx) {
          const float v0 = prow[ky * G::PW + px * STRIDE + kx];
          const float v1 = prow[ky * G::PW + (px + 1) * STRIDE + kx];   // inside the patch even for the odd tail
          a00 = fmaf(v0, w0[ky * KSZ + kx], a00);
          a01 = fmaf(v0, w1[ky * KSZ + kx], a01);
          a10 = fmaf(v1, w0[ky * KSZ + kx], a10);
          a11 = fmaf(v1, w1[ky * KSZ + kx], a11);
        }
      orow[(size_t)px * 64 + lane] = a00;
      orow[(size_t)px * 64 + lane + 32] = a01;
      if (px + 1 < npx) {
        orow[(size_t)(px + 1) * 64 + lane] = a10;
        orow[(size_t)(px + 1) * 64 + lane + 32] = a11;
      }
    }
  }
}

// part[block][t][co] = sum over the block's tiles of x[.. + tap t] * dy[n, oy, ox, co]
template <int KSZ, int STRIDE>
__global__ void __launch_bounds__(ST_THREADS) stem_conv_wgrad_kernel(const float* __restrict__ x,
                                                                     const float* __restrict__ dy,
                                                                     float* __restrict__ part, int H, int W, int OH,
                                                                     int OW, int pad, int tiles_x, int tiles_y,
                                                                     int n_tiles) {
  using G = StemGeom<KSZ, STRIDE>;
  constexpr int KK = KSZ * KSZ;
  __shared__ float patch[G::PH * G::PW];
  __shared__ float red[KK * 64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float a0[KK], a1[KK];
#pragma unroll
  for (int t = 0; t < KK; ++t) a0[t] = a1[t] = 0.f;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    int r = tile;
    const int tx = r % tiles_x; r /= tiles_x;
    const int ty = r % tiles_y;
    const int n = r / tiles_y;
    __syncthreads();
    stem_load_patch<KSZ, STRIDE>(x, n, H, W, ty, tx, pad, patch);
    __syncthreads();
    const int oy = ty * ST_TH + warp;
    if (oy >= OH) continue;
    const float* prow = patch + warp * STRIDE * G::PW;
    const float* grow = dy + (((size_t)n * OH + oy) * OW + (size_t)tx * ST_TW) * 64;
    const int npx = min(ST_TW, OW - tx * ST_TW);
    float g0 = grow[lane], g1 = grow[lane + 32];
    for (int px = 0; px < npx; ++px) {
      const float c0 = g0, c1 = g1;
      if (px + 1 < npx) {                              // prefetch the next pixel's gradient row
        g0 = grow[(size_t)(px + 1) * 64 + lane];
        g1 = grow[(size_t)(px + 1) * 64 + lane + 32];
      }
#pragma unroll
      for (int ky = 0; ky < KSZ; ++ky)
#pragma unroll
        for (int kx = 0; kx < KSZ; ++kx) {
          const float v = prow[ky * G::PW + px * STRIDE + kx];
          a0[ky * KSZ + kx] = fmaf(v, c0, a0[ky * KSZ + kx]);
          a1[ky * KSZ + kx] = fmaf(v, c1, a1[ky * KSZ + kx]);
        }
    }
  }
  // warps add their partials one after the other (fixed order => deterministic)
  for (int wv = 0; wv < ST_THREADS / 32; ++wv) {
    __syncthreads();
    if (warp == wv) {
#pragma unroll
      for (int t = 0; t < KK; ++t) {
        red[t * 64 + lane] = (wv == 0 ? 0.f : red[t * 64 + lane]) + a0[t];
        red[t * 64 + lane + 32] = (wv == 0 ? 0.f : red[t * 64 + lane + 32]) + a1[t];
      }
    }
  }
  __syncthreads();
  float* dst = part + (size_t)blockIdx.x * KK * 64;
  for (int i = threadIdx.x; i < KK * 64; i += ST_THREADS) dst[i] = red[i];
}

// dw[co][t] = sum_blocks part[b][t][co]: one warp per output, lanes stride over the blocks (double, fixed tree)
__global__ void stem_wgrad_reduce_kernel(const float* __restrict__ part, int nblocks, int KK, float* __restrict__ dw) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= KK * 64) return;
  const int co = i / KK, t = i - co * KK;
  double s = 0.0;
  for (int b = lane; b < nblocks; b += 32) s += (double)part[((size_t)b * KK + t) * 64 + co];
  s = warp_sum_d(s);
  if (lane == 0) dw[i] = (float)s;
}

int stem_blocks(int n_tiles) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int nb = sms * 2;
  return n_tiles < nb ? n_tiles : nb;
}

bool stem_supported(int Cout, int k, int stride) { return Cout == 64 && k == 7 && stride == 2; }

}  // namespace

extern "C" int essb_stem_conv_supported(int Cout, int k, int stride) { return stem_supported(Cout, k, stride) ? 1 : 0; }

extern "C" int essb_stem_conv_fwd(const float* x, const float* w, float* out, int N, int H, int W, int Cout, int k,
                                  int stride, int pad, void* stream) {
  ESSB_REQUIRE(x && w && out && N > 0 && H > 0 && W > 0, "essb_stem_conv_fwd: bad arguments");
  ESSB_REQUIRE(stem_supported(Cout, k, stride), "essb_stem_conv_fwd: only Cout=64, k=7, stride=2 is built (the ResNet-18 stem)");
  ESSB_REQUIRE(pad >= 0 && pad < k, "essb_stem_conv_fwd: bad padding");
  const int OH = (H + 2 * pad - k) / stride + 1, OW = (W + 2 * pad - k) / stride + 1;
  const int tiles_x = (OW + ST_TW - 1) / ST_TW, tiles_y = (OH + ST_TH - 1) / ST_TH;
  const int n_tiles = N * tiles_x * tiles_y;
  stem_conv_fwd_kernel<7, 2><<<stem_blocks(n_tiles), ST_THREADS, 0, (cudaStream_t)stream>>>(x, w, out, H, W, OH, OW, pad,
                                                                                          tiles_x, tiles_y, n_tiles);
  ESSB_LAUNCH_CHECK("essb_stem_conv_fwd");
  return ESSB_OK;
}

extern "C" int64_t essb_stem_conv_wgrad_workspace_bytes(int Cout, int k) {
  if (Cout != 64 || k != 7) return -1;
  return (int64_t)148 * 2 * k * k * 64 * (int64_t)sizeof(float);
}

extern "C" int essb_stem_conv_wgrad(const float* x, const float* dy, float* dw, int N, int H, int W, int Cout, int k,
                                    int stride, int pad, float* workspace, int64_t workspace_bytes, void* stream) {
  ESSB_REQUIRE(x && dy && dw && workspace && N > 0 && H > 0 && W > 0, "essb_stem_conv_wgrad: bad arguments");
  ESSB_REQUIRE(stem_supported(Cout, k, stride), "essb_stem_conv_wgrad: only Cout=64, k=7, stride=2 is built (the ResNet-18 stem)");
  ESSB_REQUIRE(pad >= 0 && pad < k, "essb_stem_conv_wgrad: bad padding");
  const int OH = (H + 2 * pad - k) / stride + 1, OW = (W + 2 * pad - k) / stride + 1;
  const int tiles_x = (OW + ST_TW - 1) / ST_TW, tiles_y = (OH + ST_TH - 1) / ST_TH;
  const int n_tiles = N * tiles_x * tiles_y;
  const int nb = stem_blocks(n_tiles);
  if (workspace_bytes < (int64_t)nb * k * k * 64 * (int64_t)sizeof(float)) {
    essb_set_error("essb_stem_conv_wgrad: workspace too small");
    return ESSB_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  stem_conv_wgrad_kernel<7, 2><<<nb, ST_THREADS, 0, st>>>(x, dy, workspace, H, W, OH, OW, pad, tiles_x, tiles_y, n_tiles);
  ESSB_LAUNCH_CHECK("essb_stem_conv_wgrad");
  const int total = k * k * 64;
  stem_wgrad_reduce_kernel<<<(total * 32 + 255) / 256, 256, 0, st>>>(workspace, nb, k * k, dw);
  ESSB_LAUNCH_CHECK("essb_stem_conv_wgrad reduce");
  return ESSB_OK;
}
