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

// 16-byte global -> shared copy that needs no register staging (all of a thread's copies are in flight at once)
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src)
               : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src)
               : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Pixels per block: the block's [TPIX][CIN] fp32 slab is staged through shared memory (row pitch CIN + 4 floats:
// 128-bit row reads by consecutive threads are bank-conflict free), so every global access is a fully
// coalesced 128-bit one even though a thread owns a whole pixel row.
template <int CIN>
struct PwTile {
  static constexpr int TPIX = CIN == 32 ? 256 : 128;
  static constexpr int PITCH = CIN + 4;
};

// out[p][k] = act( sum_c f(x[p][c]) * w[k][c] + bias[k] ),  f = optional (x - mean) * rstd, ReLU
template <int CIN>
__global__ void __launch_bounds__(PW_THREADS) pw_conv_fwd_kernel(essb_src s, const float* __restrict__ w,
                                                                 const float* __restrict__ bias,
                                                                 float* __restrict__ out, int ldo, long long rows,
                                                                 long long P, int Cout, int act) {
  constexpr int TPIX = PwTile<CIN>::TPIX, PITCH = PwTile<CIN>::PITCH, C4 = CIN / 4;
  __shared__ __align__(16) float w_s[PW_MAX_COUT * CIN];
  __shared__ float b_s[PW_MAX_COUT];
  __shared__ __align__(16) float x_s[TPIX * PITCH];     // input slab, later the [TPIX][Cout] output slab
  for (int i = threadIdx.x; i < Cout * CIN; i += PW_THREADS) w_s[i] = w[i];
  if (threadIdx.x < Cout) b_s[threadIdx.x] = bias ? bias[threadIdx.x] : 0.f;
  const long long row0 = (long long)blockIdx.x * TPIX;
  const long long left = rows - row0;
  const int rows_here = left < TPIX ? (int)left : TPIX;
  for (int i = threadIdx.x; i < rows_here * C4; i += PW_THREADS) {
    const int r = i / C4, c4 = i - r * C4;
    cp_async16(x_s + r * PITCH + c4 * 4, s.ptr + (row0 + r) * s.ld + c4 * 4);
  }
  cp_async_wait_all();
  __syncthreads();
  float ov[PW_MAX_COUT];
  const bool active = threadIdx.x < rows_here;
  if (active) {
    float xv[CIN];
#pragma unroll
    for (int c = 0; c < CIN; c += 4) {
      const float4 v = *reinterpret_cast<const float4*>(x_s + threadIdx.x * PITCH + c);
      xv[c] = v.x; xv[c + 1] = v.y; xv[c + 2] = v.z; xv[c + 3] = v.w;
    }
    if (s.mean) {
      const long long n = (row0 + threadIdx.x) / P;
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
#pragma unroll
    for (int k = 0; k < PW_MAX_COUT; ++k) {
      ov[k] = 0.f;
      if (k < Cout) {
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
        ov[k] = v;
      }
    }
  }
  __syncthreads();                                      // every row has been read: reuse x_s for the outputs
  if (active) {
#pragma unroll
    for (int k = 0; k < PW_MAX_COUT; ++k)
      if (k < Cout) x_s[threadIdx.x * Cout + k] = ov[k];
  }
  __syncthreads();
  // coalesced write-back of the block's [rows_here][Cout] slab
  if (ldo == Cout) {
    float* dst = out + row0 * Cout;
    for (int i = threadIdx.x; i < rows_here * Cout; i += PW_THREADS) dst[i] = x_s[i];
  } else {
    for (int i = threadIdx.x; i < rows_here * Cout; i += PW_THREADS) {
      const int r = i / Cout, k = i - r * Cout;
      out[(row0 + r) * ldo + k] = x_s[i];
    }
  }
}

// dx[p][c] = sum_k dy[p][k] * w[k][c]
template <int CIN>
__global__ void __launch_bounds__(PW_THREADS) pw_conv_dgrad_kernel(const float* __restrict__ dy, int ld_dy,
                                                                   const float* __restrict__ w,
                                                                   float* __restrict__ dx, int ld_dx, long long rows,
                                                                   int Cout) {
  constexpr int TPIX = PwTile<CIN>::TPIX, PITCH = PwTile<CIN>::PITCH, C4 = CIN / 4;
  __shared__ __align__(16) float w_s[PW_MAX_COUT * CIN];
  __shared__ __align__(16) float t_s[TPIX * PITCH];     // dy slab [TPIX][Cout], later the dx slab [TPIX][PITCH]
  for (int i = threadIdx.x; i < Cout * CIN; i += PW_THREADS) w_s[i] = w[i];
  const long long row0 = (long long)blockIdx.x * TPIX;
  const long long left = rows - row0;
  const int rows_here = left < TPIX ? (int)left : TPIX;
  if (ld_dy == Cout) {
    const float* src = dy + row0 * Cout;
    for (int i = threadIdx.x; i < rows_here * Cout; i += PW_THREADS) cp_async4(t_s + i, src + i);
  } else {
    for (int i = threadIdx.x; i < rows_here * Cout; i += PW_THREADS) {
      const int r = i / Cout, k = i - r * Cout;
      cp_async4(t_s + i, dy + (row0 + r) * ld_dy + k);
    }
  }
  cp_async_wait_all();
  __syncthreads();
  const bool active = threadIdx.x < rows_here;
  float g[PW_MAX_COUT];
#pragma unroll
  for (int k = 0; k < PW_MAX_COUT; ++k) g[k] = (active && k < Cout) ? t_s[threadIdx.x * Cout + k] : 0.f;
  __syncthreads();
  if (active) {
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
      *reinterpret_cast<float4*>(t_s + threadIdx.x * PITCH + c) = a;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < rows_here * C4; i += PW_THREADS) {
    const int r = i / C4, c4 = i - r * C4;
    *reinterpret_cast<float4*>(dx + (row0 + r) * ld_dx + c4 * 4) = *reinterpret_cast<const float4*>(t_s + r * PITCH + c4 * 4);
  }
}

// part[block][k][c] = sum over the block's rows of dy[p][k] * f(x[p][c]);  part[block][16][k] = sum dy[p][k].
// A warp instruction loads RPW = 128 / CIN consecutive pixel rows as float4 per lane (fully coalesced); lane =
// (row within the group, 4-channel group).  Each lane keeps 4 x 16 accumulators; row groups are folded with
// shuffles once at the end.
__device__ __forceinline__ void pw_cp_async16(float* dst_smem, const float* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void pw_cp_async4(float* dst_smem, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void pw_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void pw_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Tiles of TROWS pixel rows (activation rows [TROWS][CIN] and dy rows [TROWS][Cout]) are staged in shared memory with
// cp.async, double-buffered: the loads of tile t+1 are in flight while tile t is multiplied, at no register cost.  (With the
// loads issued from registers the kernel sat at 25 % occupancy -- 64 + 16 accumulators per lane -- waiting on global memory:
// ncu long-scoreboard 53 % of the stall samples, 1.2 TB/s; profiles/r02o_hbm_kernels.md.)
template <int CIN>
__global__ void __launch_bounds__(PW_THREADS) pw_conv_wgrad_kernel(essb_src s, const float* __restrict__ dy, int ld_dy,
                                                                   long long rows, long long P, int Cout,
                                                                   long long rows_per_block, float* __restrict__ part) {
  constexpr int C4 = CIN / 4;                          // lanes per pixel row
  constexpr int RPW = 32 / C4;                         // pixel rows per warp load (4 or 2)
  constexpr int GP = PW_MAX_COUT + 4;                  // dy row pitch in shared memory (conflict-free 128-bit reads)
  constexpr int TROWS = PW_THREADS;                    // rows per staged tile
  extern __shared__ __align__(16) float pw_dyn[];      // [2][TROWS * CIN] activation tiles, then [2][TROWS * GP] dy tiles
  float* const x_s = pw_dyn;
  float* const g_s = pw_dyn + 2 * TROWS * CIN;
  __shared__ float red[PW_MAX_COUT + 1][CIN];          // cross-warp reduction ([PW_MAX_COUT] row = bias partials)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cg = lane % C4, rg = lane / C4;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > rows) r1 = rows;
  float acc[4][PW_MAX_COUT];
  float accb[PW_MAX_COUT];
#pragma unroll
  for (int k = 0; k < PW_MAX_COUT; ++k) {
    accb[k] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j][k] = 0.f;
  }
  // the dy columns [Cout, 16) are never written by the copies: zero both stages once
  for (int i = threadIdx.x; i < 2 * TROWS * GP; i += PW_THREADS) g_s[i] = 0.f;
  __syncthreads();
  auto issue = [&](int st, long long t0) {
    const long long left = r1 - t0;
    const int rows_here = left < TROWS ? (int)left : TROWS;
    float* xs = x_s + st * (TROWS * CIN);
    float* gs = g_s + st * (TROWS * GP);
    for (int i = threadIdx.x; i < rows_here * C4; i += PW_THREADS) {
      const int r = i / C4, c4 = i - r * C4;
      pw_cp_async16(xs + r * CIN + c4 * 4, s.ptr + (t0 + r) * s.ld + c4 * 4);
    }
    for (int i = threadIdx.x; i < rows_here * Cout; i += PW_THREADS) {
      const int r = i / Cout, k = i - r * Cout;
      pw_cp_async4(gs + r * GP + k, dy + (t0 + r) * ld_dy + k);
    }
    pw_cp_async_commit();
  };
  int st = 0;
  if (r0 < r1) issue(0, r0);
  for (long long t0 = r0; t0 < r1; t0 += TROWS, st ^= 1) {
    const long long left = r1 - t0;
    const int rows_here = left < TROWS ? (int)left : TROWS;
    if (t0 + TROWS < r1) {
      issue(st ^ 1, t0 + TROWS);                       // stage st^1 was released by the barrier that ended the last iteration
      pw_cp_async_wait<1>();
    } else {
      pw_cp_async_wait<0>();
    }
    __syncthreads();
    const float* xs = x_s + st * (TROWS * CIN);
    const float* gsb = g_s + st * (TROWS * GP);
    const int wr0 = warp * 32;                          // this warp's 32 rows of the tile
    if (wr0 < rows_here) {
      const int wrows = min(32, rows_here - wr0);
      const long long p0 = t0 + wr0;
      const long long n_first = p0 / P, n_last = (p0 + wrows - 1) / P;
      float4 mean = make_float4(0.f, 0.f, 0.f, 0.f), rstd = make_float4(1.f, 1.f, 1.f, 1.f);
      if (s.mean) {
        mean = *reinterpret_cast<const float4*>(s.mean + n_first * CIN + cg * 4);
        rstd = *reinterpret_cast<const float4*>(s.rstd + n_first * CIN + cg * 4);
      }
#pragma unroll 4
      for (int i = 0; i < 32; i += RPW) {
        const int r = i + rg;
        if (r < wrows) {
          const long long p = p0 + r;
          float4 v = *reinterpret_cast<const float4*>(xs + (wr0 + r) * CIN + cg * 4);
          if (s.mean) {
            float4 m = mean, q = rstd;
            if (n_first != n_last && p / P != n_first) {    // the warp's rows straddle two samples (rare)
              m = *reinterpret_cast<const float4*>(s.mean + n_last * CIN + cg * 4);
              q = *reinterpret_cast<const float4*>(s.rstd + n_last * CIN + cg * 4);
            }
            v.x = (v.x - m.x) * q.x; v.y = (v.y - m.y) * q.y; v.z = (v.z - m.z) * q.z; v.w = (v.w - m.w) * q.w;
          }
          if (s.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
          const float4* gr = reinterpret_cast<const float4*>(gsb + (wr0 + r) * GP);
#pragma unroll
          for (int k4 = 0; k4 < PW_MAX_COUT / 4; ++k4) {
            const float4 g = gr[k4];
            const float gk[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int k = k4 * 4 + e;
              acc[0][k] = fmaf(v.x, gk[e], acc[0][k]);
              acc[1][k] = fmaf(v.y, gk[e], acc[1][k]);
              acc[2][k] = fmaf(v.z, gk[e], acc[2][k]);
              acc[3][k] = fmaf(v.w, gk[e], acc[3][k]);
              if (cg == 0) accb[k] += gk[e];
            }
          }
        }
      }
    }
    __syncthreads();                                    // every warp is done with stage st before it is refilled
  }
  // fold the row groups of the warp (lanes with equal cg), then the warps one after the other (deterministic)
#pragma unroll
  for (int k = 0; k < PW_MAX_COUT; ++k) {
#pragma unroll
    for (int o = C4; o < 32; o <<= 1) {
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j][k] += __shfl_xor_sync(0xffffffffu, acc[j][k], o);
      accb[k] += __shfl_xor_sync(0xffffffffu, accb[k], o);
    }
  }
  for (int wv = 0; wv < PW_THREADS / 32; ++wv) {
    __syncthreads();
    if (warp == wv && rg == 0) {
#pragma unroll
      for (int k = 0; k < PW_MAX_COUT; ++k) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float* r = &red[k][cg * 4 + j];
          *r = (wv == 0 ? 0.f : *r) + acc[j][k];
        }
        if (cg == 0) {
          float* r = &red[PW_MAX_COUT][k];
          *r = (wv == 0 ? 0.f : *r) + accb[k];
        }
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

// dw[k][c] = sum_blocks part[b][k][c], dbias[k] = sum_blocks part[b][16][k]: one warp per output, lanes stride over
// the blocks in double precision, fixed shuffle tree => deterministic
__global__ void pw_conv_wgrad_reduce_kernel(const float* __restrict__ part, int nblocks, int Cin, int Cout,
                                            float* __restrict__ dw, float* __restrict__ dbias) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int slab = (PW_MAX_COUT + 1) * Cin;
  if (i >= Cout * Cin + Cout) return;
  const bool is_w = i < Cout * Cin;
  if ((is_w && !dw) || (!is_w && !dbias)) return;
  const int off = is_w ? i : PW_MAX_COUT * Cin + (i - Cout * Cin);
  double t = 0.0;
  for (int b = lane; b < nblocks; b += 32) t += (double)part[(size_t)b * slab + off];
  t = warp_sum_d(t);
  if (lane == 0) {
    if (is_w) dw[i] = (float)t;
    else dbias[i - Cout * Cin] = (float)t;
  }
}

int pw_wgrad_blocks(long long rows) {
  long long nb = (rows + 2047) / 2048;
  if (nb > 148 * 2) nb = 148 * 2;      // two resident blocks per SM (104 KB of staging each): long tile loops keep the pipeline full
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
  const int tpix = src->C == 32 ? PwTile<32>::TPIX : PwTile<64>::TPIX;
  const unsigned blocks = (unsigned)((rows + tpix - 1) / tpix);
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
  const int tpix = Cin == 32 ? PwTile<32>::TPIX : PwTile<64>::TPIX;
  const unsigned blocks = (unsigned)((rows + tpix - 1) / tpix);
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
  ESSB_REQUIRE(P >= 32, "essb_pw_conv_wgrad: H*W must be >= 32 (a warp's 32 rows may straddle at most two samples)");
  const int nb = pw_wgrad_blocks(rows);
  const int64_t need = essb_pw_conv_wgrad_workspace_bytes(N, H, W, src->C);
  if (workspace_bytes < need) {
    essb_set_error("essb_pw_conv_wgrad: workspace %lld < %lld bytes", (long long)workspace_bytes, (long long)need);
    return ESSB_ERR_WORKSPACE;
  }
  const long long rpb = (rows + nb - 1) / nb;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)2 * PW_THREADS * (src->C + PW_MAX_COUT + 4) * sizeof(float);   // two stages of (x tile + dy tile)
  cudaError_t e;
  if (src->C == 32) {
    e = cudaFuncSetAttribute(pw_conv_wgrad_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) pw_conv_wgrad_kernel<32><<<nb, PW_THREADS, smem, st>>>(*src, dy, ld_dy, rows, P, Cout, rpb, workspace);
  } else {
    e = cudaFuncSetAttribute(pw_conv_wgrad_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) pw_conv_wgrad_kernel<64><<<nb, PW_THREADS, smem, st>>>(*src, dy, ld_dy, rows, P, Cout, rpb, workspace);
  }
  if (e != cudaSuccess) {
    essb_set_error("essb_pw_conv_wgrad: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    return ESSB_ERR_LAUNCH;
  }
  ESSB_LAUNCH_CHECK("essb_pw_conv_wgrad");
  const int total = Cout * src->C + Cout;
  pw_conv_wgrad_reduce_kernel<<<(total * 32 + 255) / 256, 256, 0, st>>>(workspace, nb, src->C, Cout, dw, dbias);
  ESSB_LAUNCH_CHECK("essb_pw_conv_wgrad reduce");
  return ESSB_OK;
}
