// 1x1 convolutions with few output channels: the SemSegE2VID classifier (decoder_scale_5 = Conv2d(32, K, 1),
// models/style_networks.py:34,88) forward / input gradient / weight gradient, and the E2VID prediction layer
// (conv1x1 + folded BN + sigmoid, e2vid/model/unet.py:65-67,179).
//
// These layers are HBM-bound (Cin <= 64, Cout <= 16: ~11 FLOP per byte), so they do not go through the
// implicit-GEMM kernels: one thread per pixel streams its channel vector with 128-bit loads, the weights sit in
// shared memory (broadcast reads), rows of the narrow side are staged through shared memory so global accesses
// stay coalesced.  The source may carry the InstanceNorm + ReLU of the preceding block (essb_src), applied on
// the fly exactly as the implicit-GEMM loaders do.
#include "common.cuh"

namespace {

constexpr int PW_THREADS = 256;
constexpr int PW_MAX_COUT = 16;

// out[p][k] = act( sum_c f(x[p][c]) * w[k][c] + bias[k] ),  f = optional (x - mean) * rstd, ReLU
template <int CIN>
__global__ void __launch_bounds__(PW_THREADS) pw_conv_fwd_kernel(essb_src s, const float* __restrict__ w,
                                                                 const float* __restrict__ bias,
                                                                 float* __restrict__ out, int ldo, long long rows,
                                                                 long long P, int Cout, int act) {
  __shared__ __align__(16) float w_s[PW_MAX_COUT * CIN];
  __shared__ float b_s[PW_MAX_COUT];
  __shared__ float o_s[PW_THREADS * PW_MAX_COUT];
  for (int i = threadIdx.x; i < Cout * CIN; i += PW_THREADS) w_s[i] = w[i];
  if (threadIdx.x < Cout) b_s[threadIdx.x] = bias ? bias[threadIdx.x] : 0.f;
  __syncthreads();
  const long long row0 = (long long)blockIdx.x * PW_THREADS;
  const long long row = row0 + threadIdx.x;
  if (row < rows) {
    float xv[CIN];
    const float* src = s.ptr + row * s.ld;
#pragma unroll
    for (int c = 0; c < CIN; c += 4) {
      const float4 v = *reinterpret_cast<const float4*>(src + c);
      xv[c] = v.x; xv[c + 1] = v.y; xv[c + 2] = v.z; xv[c + 3] = v.w;
    }
    if (s.mean) {
      const long long n = row / P;
      const float* mp = s.mean + n * CIN;
      const float* rp = s.rstd + n * CIN;
#pragma unroll
      for (int c = 0; c < CIN; c += 4) {
        const float4 m = *reinterpret_cast<const float4*>(mp + c);
        const float4 r = *reinterpret_cast<const float4*>(rp + c);
        xv[c] = (xv[c] - m.x) * r.x; xv[c + 1] = (xv[c + 1] - m.y) * r.y;
        xv[c + 2] = (xv[c + 2] - m.z) * r.z; xv[c + 3] = (xv[c + 3] - m.w) * r.w;
      }
    }
    if (s.relu) {
#pragma unroll
      for (int c = 0; c < CIN; ++c) xv[c] = fmaxf(xv[c], 0.f);
    }
    for (int k = 0; k < Cout; ++k) {
      float a0 = b_s[k], a1 = 0.f, a2 = 0.f, a3 = 0.f;
      const float* wk = w_s + k * CIN;
#pragma unroll
      for (int c = 0; c < CIN; c += 4) {
        const float4 ww = *reinterpret_cast<const float4*>(wk + c);
        a0 = fmaf(xv[c], ww.x, a0); a1 = fmaf(xv[c + 1], ww.y, a1);
        a2 = fmaf(xv[c + 2], ww.z, a2); a3 = fmaf(xv[c + 3], ww.w, a3);
      }
      float v = (a0 + a1) + (a2 + a3);
      if (act == ESSB_ACT_RELU) v = fmaxf(v, 0.f);
      else if (act == ESSB_ACT_SIGMOID) v = essb_sigmoid(v);
      o_s[threadIdx.x * Cout + k] = v;
    }
  }
  __syncthreads();
  // coalesced write-back of the block's [rows_here][Cout] slab
  const long long left = rows - row0;
  const int rows_here = left < PW_THREADS ? (int)left : PW_THREADS;
  if (ldo == Cout) {
    float* dst = out + row0 * Cout;
    for (int i = threadIdx.x; i < rows_here * Cout; i += PW_THREADS) dst[i] = o_s[i];
  } else {
    for (int i = threadIdx.x; i < rows_here * Cout; i += PW_THREADS) {
      const int r = i / Cout, k = i - r * Cout;
      out[(row0 + r) * ldo + k] = o_s[i];
    }
  }
}

// dx[p][c] = sum_k dy[p][k] * w[k][c]
template <int CIN>
__global__ void __launch_bounds__(PW_THREADS) pw_conv_dgrad_kernel(const float* __restrict__ dy, int ld_dy,
                                                                   const float* __restrict__ w,
                                                                   float* __restrict__ dx, int ld_dx, long long rows,
                                                                   int Cout) {
  __shared__ __align__(16) float w_s[PW_MAX_COUT * CIN];
  __shared__ float g_s[PW_THREADS * PW_MAX_COUT];
  for (int i = threadIdx.x; i < Cout * CIN; i += PW_THREADS) w_s[i] = w[i];
  const long long row0 = (long long)blockIdx.x * PW_THREADS;
  const long long left = rows - row0;
  const int rows_here = left < PW_THREADS ? (int)left : PW_THREADS;
  if (ld_dy == Cout) {
    const float* src = dy + row0 * Cout;
    for (int i = threadIdx.x; i < rows_here * Cout; i += PW_THREADS) g_s[i] = src[i];
  } else {
    for (int i = threadIdx.x; i < rows_here * Cout; i += PW_THREADS) {
      const int r = i / Cout, k = i - r * Cout;
      g_s[i] = dy[(row0 + r) * ld_dy + k];
    }
  }
  __syncthreads();
  if (threadIdx.x >= rows_here) return;
  float g[PW_MAX_COUT];
#pragma unroll
  for (int k = 0; k < PW_MAX_COUT; ++k) g[k] = k < Cout ? g_s[threadIdx.x * Cout + k] : 0.f;
  float* dst = dx + (row0 + threadIdx.x) * ld_dx;
#pragma unroll
  for (int c = 0; c < CIN; c += 4) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < PW_MAX_COUT; ++k) {
      if (k < Cout) {
        const float4 ww = *reinterpret_cast<const float4*>(w_s + k * CIN + c);
        a.x = fmaf(g[k], ww.x, a.x); a.y = fmaf(g[k], ww.y, a.y);
        a.z = fmaf(g[k], ww.z, a.z); a.w = fmaf(g[k], ww.w, a.w);
      }
    }
    *reinterpret_cast<float4*>(dst + c) = a;
  }
}

// part[block][k][c] = sum over the block's rows of dy[p][k] * f(x[p][c]);  part[block][Cout][k] = sum dy[p][k]
// lane = input channel (CPL channels per lane: lane, lane + 32), one pixel row per warp iteration.
template <int CPL>
__global__ void __launch_bounds__(PW_THREADS) pw_conv_wgrad_kernel(essb_src s, const float* __restrict__ dy, int ld_dy,
                                                                   long long rows, long long P, int Cout,
                                                                   long long rows_per_block, float* __restrict__ part) {
  constexpr int CIN = 32 * CPL;
  constexpr int GP = PW_MAX_COUT;                      // padded dy row in shared memory
  __shared__ __align__(16) float g_s[PW_THREADS * GP];
  __shared__ float red[PW_MAX_COUT + 1][CIN];          // cross-warp reduction ([PW_MAX_COUT] row = bias partials)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > rows) r1 = rows;
  float acc[CPL][PW_MAX_COUT];
  float accb = 0.f;
#pragma unroll
  for (int j = 0; j < CPL; ++j)
#pragma unroll
    for (int k = 0; k < PW_MAX_COUT; ++k) acc[j][k] = 0.f;

  for (long long t0 = r0; t0 < r1; t0 += PW_THREADS) {
    const long long left = r1 - t0;
    const int rows_here = left < PW_THREADS ? (int)left : PW_THREADS;
    __syncthreads();
    // stage the dy tile [rows_here][Cout] -> g_s[row][GP] (zero padded)
    for (int i = threadIdx.x; i < PW_THREADS * GP; i += PW_THREADS) {
      const int r = i / GP, k = i - r * GP;
      g_s[i] = (r < rows_here && k < Cout) ? dy[(t0 + r) * ld_dy + k] : 0.f;
    }
    __syncthreads();
    const int wr0 = warp * 32;
    if (wr0 >= rows_here) continue;
    const int wrows = min(32, rows_here - wr0);
    const long long p0 = t0 + wr0;
    long long n = p0 / P;
    long long next_n = (n + 1) * P;                     // first row of the next sample
    float mean[CPL], rstd[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      mean[j] = s.mean ? s.mean[n * CIN + lane + 32 * j] : 0.f;
      rstd[j] = s.mean ? s.rstd[n * CIN + lane + 32 * j] : 1.f;
    }
#pragma unroll 4
    for (int i = 0; i < wrows; ++i) {
      const long long p = p0 + i;
      if (p >= next_n) {                                // warp-uniform
        n = p / P;
        next_n = (n + 1) * P;
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
          mean[j] = s.mean ? s.mean[n * CIN + lane + 32 * j] : 0.f;
          rstd[j] = s.mean ? s.rstd[n * CIN + lane + 32 * j] : 1.f;
        }
      }
      float a[CPL];
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        float v = s.ptr[p * s.ld + lane + 32 * j];
        v = (v - mean[j]) * rstd[j];
        if (s.relu) v = fmaxf(v, 0.f);
        a[j] = v;
      }
      const float4* gr = reinterpret_cast<const float4*>(g_s + (wr0 + i) * GP);
#pragma unroll
      for (int k4 = 0; k4 < PW_MAX_COUT / 4; ++k4) {
        const float4 g = gr[k4];                        // broadcast read
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
          acc[j][k4 * 4 + 0] = fmaf(a[j], g.x, acc[j][k4 * 4 + 0]);
          acc[j][k4 * 4 + 1] = fmaf(a[j], g.y, acc[j][k4 * 4 + 1]);
          acc[j][k4 * 4 + 2] = fmaf(a[j], g.z, acc[j][k4 * 4 + 2]);
          acc[j][k4 * 4 + 3] = fmaf(a[j], g.w, acc[j][k4 * 4 + 3]);
        }
      }
      if (lane < PW_MAX_COUT) accb += g_s[(wr0 + i) * GP + lane];
    }
  }
  // warps add their partials into `red` one after the other (fixed order => deterministic)
  for (int wv = 0; wv < PW_THREADS / 32; ++wv) {
    __syncthreads();
    if (warp == wv) {
#pragma unroll
      for (int k = 0; k < PW_MAX_COUT; ++k)
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
          float* r = &red[k][lane + 32 * j];
          *r = (wv == 0 ? 0.f : *r) + acc[j][k];
        }
      if (lane < PW_MAX_COUT) {
        float* r = &red[PW_MAX_COUT][lane];
        *r = (wv == 0 ? 0.f : *r) + accb;
      }
    }
  }
  __syncthreads();
  float* dst = part + (size_t)blockIdx.x * (PW_MAX_COUT + 1) * CIN;
  for (int i = threadIdx.x; i < (PW_MAX_COUT + 1) * CIN; i += PW_THREADS) {
    const int k = i / CIN, c = i - k * CIN;
    dst[i] = (k < PW_MAX_COUT || c < PW_MAX_COUT) ? red[k][c] : 0.f;
  }
}

// dw[k][c] = sum_blocks part[b][k][c] (double, fixed order => deterministic); dbias[k] = sum_blocks part[b][16][k]
__global__ void pw_conv_wgrad_reduce_kernel(const float* __restrict__ part, int nblocks, int Cin, int Cout,
                                            float* __restrict__ dw, float* __restrict__ dbias) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int slab = (PW_MAX_COUT + 1) * Cin;
  if (i < Cout * Cin) {
    if (!dw) return;
    const int k = i / Cin, c = i - k * Cin;
    double t = 0.0;
    for (int b = 0; b < nblocks; ++b) t += (double)part[(size_t)b * slab + k * Cin + c];
    dw[i] = (float)t;
  } else if (i < Cout * Cin + Cout) {
    if (!dbias) return;
    const int k = i - Cout * Cin;
    double t = 0.0;
    for (int b = 0; b < nblocks; ++b) t += (double)part[(size_t)b * slab + PW_MAX_COUT * Cin + k];
    dbias[k] = (float)t;
  }
}

int pw_wgrad_blocks(long long rows) {
  long long nb = (rows + 2047) / 2048;
  if (nb > 148 * 8) nb = 148 * 8;
  if (nb < 1) nb = 1;
  return (int)nb;
}

bool pw_src_ok(const essb_src* s) {
  return s && s->ptr && (s->C == 32 || s->C == 64) && s->ld % 4 == 0 && s->ups == 0 && essb_aligned16(s->ptr) &&
         ((s->mean == nullptr) == (s->rstd == nullptr)) && (!s->mean || (essb_aligned16(s->mean) && essb_aligned16(s->rstd)));
}

}  // namespace

extern "C" int essb_pw_conv_fwd(const essb_src* src, const float* w, const float* bias, float* out, int ldo, int N,
                                int H, int W, int Cout, int act, void* stream) {
  ESSB_REQUIRE(pw_src_ok(src), "essb_pw_conv_fwd: source must be fp32 NHWC with 32 or 64 channels, ld %% 4 == 0, no upsampling");
  ESSB_REQUIRE(w && out && N > 0 && H > 0 && W > 0 && Cout >= 1 && Cout <= PW_MAX_COUT && ldo >= Cout,
               "essb_pw_conv_fwd: bad arguments (1 <= Cout <= %d)", PW_MAX_COUT);
  const long long P = (long long)H * W, rows = P * N;
  const unsigned blocks = (unsigned)((rows + PW_THREADS - 1) / PW_THREADS);
  cudaStream_t st = (cudaStream_t)stream;
  if (src->C == 32)
    pw_conv_fwd_kernel<32><<<blocks, PW_THREADS, 0, st>>>(*src, w, bias, out, ldo, rows, P, Cout, act);
  else
    pw_conv_fwd_kernel<64><<<blocks, PW_THREADS, 0, st>>>(*src, w, bias, out, ldo, rows, P, Cout, act);
  ESSB_LAUNCH_CHECK("essb_pw_conv_fwd");
  return ESSB_OK;
}

extern "C" int essb_pw_conv_dgrad(const float* dy, int ld_dy, const float* w, float* dx, int ld_dx, int64_t rows,
                                  int Cin, int Cout, void* stream) {
  ESSB_REQUIRE(dy && w && dx && rows > 0 && (Cin == 32 || Cin == 64) && Cout >= 1 && Cout <= PW_MAX_COUT &&
                   ld_dy >= Cout && ld_dx >= Cin && ld_dx % 4 == 0 && essb_aligned16(dx),
               "essb_pw_conv_dgrad: bad arguments (Cin 32 or 64, 1 <= Cout <= %d)", PW_MAX_COUT);
  const unsigned blocks = (unsigned)((rows + PW_THREADS - 1) / PW_THREADS);
  cudaStream_t st = (cudaStream_t)stream;
  if (Cin == 32)
    pw_conv_dgrad_kernel<32><<<blocks, PW_THREADS, 0, st>>>(dy, ld_dy, w, dx, ld_dx, rows, Cout);
  else
    pw_conv_dgrad_kernel<64><<<blocks, PW_THREADS, 0, st>>>(dy, ld_dy, w, dx, ld_dx, rows, Cout);
  ESSB_LAUNCH_CHECK("essb_pw_conv_dgrad");
  return ESSB_OK;
}

extern "C" int64_t essb_pw_conv_wgrad_workspace_bytes(int N, int H, int W, int Cin) {
  if (N <= 0 || H <= 0 || W <= 0 || (Cin != 32 && Cin != 64)) return -1;
  return (int64_t)pw_wgrad_blocks((long long)N * H * W) * (PW_MAX_COUT + 1) * Cin * (int64_t)sizeof(float);
}

extern "C" int essb_pw_conv_wgrad(const essb_src* src, const float* dy, int ld_dy, int N, int H, int W, int Cout,
                                  float* dw, float* dbias, float* workspace, int64_t workspace_bytes, void* stream) {
  ESSB_REQUIRE(pw_src_ok(src), "essb_pw_conv_wgrad: source must be fp32 NHWC with 32 or 64 channels, ld %% 4 == 0, no upsampling");
  ESSB_REQUIRE(dy && workspace && (dw || dbias) && N > 0 && H > 0 && W > 0 && Cout >= 1 && Cout <= PW_MAX_COUT &&
                   ld_dy >= Cout,
               "essb_pw_conv_wgrad: bad arguments (1 <= Cout <= %d)", PW_MAX_COUT);
  const long long P = (long long)H * W, rows = P * N;
  const int nb = pw_wgrad_blocks(rows);
  const int64_t need = essb_pw_conv_wgrad_workspace_bytes(N, H, W, src->C);
  if (workspace_bytes < need) {
    essb_set_error("essb_pw_conv_wgrad: workspace %lld < %lld bytes", (long long)workspace_bytes, (long long)need);
    return ESSB_ERR_WORKSPACE;
  }
  const long long rpb = (rows + nb - 1) / nb;
  cudaStream_t st = (cudaStream_t)stream;
  if (src->C == 32)
    pw_conv_wgrad_kernel<1><<<nb, PW_THREADS, 0, st>>>(*src, dy, ld_dy, rows, P, Cout, rpb, workspace);
  else
    pw_conv_wgrad_kernel<2><<<nb, PW_THREADS, 0, st>>>(*src, dy, ld_dy, rows, P, Cout, rpb, workspace);
  ESSB_LAUNCH_CHECK("essb_pw_conv_wgrad");
  const int total = Cout * src->C + Cout;
  pw_conv_wgrad_reduce_kernel<<<(total + 127) / 128, 128, 0, st>>>(workspace, nb, src->C, Cout, dw, dbias);
  ESSB_LAUNCH_CHECK("essb_pw_conv_wgrad reduce");
  return ESSB_OK;
}
