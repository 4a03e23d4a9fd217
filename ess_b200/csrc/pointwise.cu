// Memory-bound helpers around the convolutions: weight packing, instance-norm statistics and
// backward, event pre-processing, layout conversion.  All HBM-bound; coalesced float4 access,
// grids sized from the element count.
#include <cuda_bf16.h>

#include "common.cuh"

namespace {

constexpr int EW_THREADS = 256;

inline unsigned ew_blocks(long long items) {
  long long b = (items + EW_THREADS - 1) / EW_THREADS;
  if (b < 1) b = 1;
  return (unsigned)b;
}

// ------------------------------------------------------------------------------ weight packing
__global__ void pack_weight_kernel(const float* __restrict__ w, const float* __restrict__ scale,
                                   float* __restrict__ out, int Cout, int Cin, int T, int transposed_layout,
                                   int swap_io, int flip, int interleave) {
  const int Kin = swap_io ? Cout : Cin;
  const int Nout = swap_io ? Cin : Cout;
  const int NoutP = (Nout + 3) & ~3;
  const long long total = (long long)T * Kin * NoutP;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int np = (int)(idx % NoutP);
  const long long r = idx / NoutP;
  const int k = (int)(r % Kin);
  const int t = (int)(r / Kin);
  float v = 0.f;
  if (np < Nout) {
    int co, ci;
    if (swap_io) { co = k; ci = np; }
    else {
      ci = k;
      co = np;
      if (interleave > 1) {  // packed co' = (co % G)*g + co / G  with G = Cout / g   =>  invert
        const int G = Cout / interleave;
        co = (np % interleave) * G + np / interleave;
      }
    }
    const int ts = flip ? T - 1 - t : t;
    const size_t src = transposed_layout ? ((size_t)ci * Cout + co) * T + ts : ((size_t)co * Cin + ci) * T + ts;
    v = w[src];
    if (scale && !swap_io) v *= scale[co];
  }
  out[idx] = v;
}

// --------------------------------------------------------------------- partial-sum reductions
// One warp per (n, c): sum partial[n][b][c][0..1] over b in double, fixed order.
__global__ void in_finalize_kernel(const float* __restrict__ partial, int N, int tiles, int C, double inv_count,
                                   float eps, float* __restrict__ mean, float* __restrict__ rstd,
                                   float* __restrict__ totals) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= N * C) return;
  const int n = warp / C, c = warp - n * C;
  double s1 = 0.0, s2 = 0.0;
  for (int b = lane; b < tiles; b += 32) {
    const float* q = partial + (((size_t)n * tiles + b) * C + c) * 2;
    s1 += (double)q[0];
    s2 += (double)q[1];
  }
  s1 = warp_sum_d(s1);
  s2 = warp_sum_d(s2);
  if (lane == 0) {
    if (totals) {
      totals[((size_t)n * C + c) * 2 + 0] = (float)s1;
      totals[((size_t)n * C + c) * 2 + 1] = (float)s2;
    } else {
      const double m = s1 * inv_count;
      double var = s2 * inv_count - m * m;
      if (var < 0.0) var = 0.0;
      mean[(size_t)n * C + c] = (float)m;
      rstd[(size_t)n * C + c] = (float)(1.0 / sqrt(var + (double)eps));
    }
  }
}

// ------------------------------------------------------------------------ elementwise IN ops
__global__ void norm_act_add_kernel(const float* __restrict__ y, int ld_y, const float* __restrict__ mean,
                                    const float* __restrict__ rstd, int relu, const float* __restrict__ res,
                                    int ld_res, float* __restrict__ out, int ld_out, long long P, int C,
                                    long long total /* N*P*C/4 */) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int CQ = C >> 2;
  const int c = (int)(idx % CQ) * 4;
  const long long pix = idx / CQ;
  const int n = (int)(pix / P);
  float4 v = *reinterpret_cast<const float4*>(y + pix * ld_y + c);
  if (mean) {
    const float4 m = *reinterpret_cast<const float4*>(mean + (size_t)n * C + c);
    const float4 r = *reinterpret_cast<const float4*>(rstd + (size_t)n * C + c);
    v.x = (v.x - m.x) * r.x; v.y = (v.y - m.y) * r.y; v.z = (v.z - m.z) * r.z; v.w = (v.w - m.w) * r.w;
  }
  if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
  if (res) {
    const float4 q = *reinterpret_cast<const float4*>(res + pix * ld_res + c);
    v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
  }
  *reinterpret_cast<float4*>(out + pix * ld_out + c) = v;
}

// Pixels reduced by one block of the InstanceNorm statistics / backward kernels.  Small images get small blocks
// so that even a 55x80 layer launches >= 4 blocks per SM (ncu: 288 blocks of 256 pixels left the 1/8-scale
// layers at 23 % resident warps and 34 % of DRAM bandwidth); the per-block partials are reduced by one warp per
// (n, c) in a fixed order, so the result does not depend on the block size.
static inline int in_pix_per_block(long long P) { return P >= 65536 ? 256 : (P >= 16384 ? 128 : 64); }

// pass 1 of the IN(+ReLU)(+upsample) backward.  grid = (blocks_per_sample, N, channel groups of 128)
__global__ void __launch_bounds__(256) in_bwd_pass1_kernel(
    const float* __restrict__ dA, int ld_dA, int ups, const float* __restrict__ extra, int ld_extra,
    const float* __restrict__ y, int ld_y, const float* __restrict__ mean, const float* __restrict__ rstd,
    int relu, float* __restrict__ g, float* __restrict__ partial, int H, int W, int C, int ppb) {
  __shared__ float red[8][32][8];
  const int n = blockIdx.y, blk = blockIdx.x;
  const int CQ = C >> 2;
  const int cq_base = blockIdx.z * 32;
  const int cq_left = CQ - cq_base;
  // lanes per pixel: 32 when >=32 quads remain, else next power of two >= cq_left
  int lpp = 32;
  if (cq_left < 32) { lpp = 1; while (lpp < cq_left) lpp <<= 1; }
  const int ppw = 32 / lpp;  // pixels per warp iteration
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = lane % lpp, sub = lane / lpp;
  const bool cq_ok = q < cq_left;
  const int c = (cq_base + q) * 4;
  const long long P = (long long)H * W;
  float4 m4 = make_float4(0.f, 0.f, 0.f, 0.f), r4 = make_float4(1.f, 1.f, 1.f, 1.f);
  if (cq_ok && mean) {
    m4 = *reinterpret_cast<const float4*>(mean + (size_t)n * C + c);
    r4 = *reinterpret_cast<const float4*>(rstd + (size_t)n * C + c);
  }
  float sg[4] = {0.f, 0.f, 0.f, 0.f}, sgx[4] = {0.f, 0.f, 0.f, 0.f};
  const long long p_begin = (long long)blk * ppb;
  for (int i = warp * ppw + sub; i < ppb; i += 8 * ppw) {
    const long long pp = p_begin + i;
    if (pp >= P || !cq_ok) continue;
    const long long pix = (long long)n * P + pp;
    float4 a;
    if (ups) {
      const int py = (int)(pp / W), px = (int)(pp - (long long)py * W);
      const int W2 = W * 2;
      const long long base = ((long long)n * (2 * H) + 2 * py) * W2 + 2 * px;
      const float4 a0 = *reinterpret_cast<const float4*>(dA + base * ld_dA + c);
      const float4 a1 = *reinterpret_cast<const float4*>(dA + (base + 1) * ld_dA + c);
      const float4 a2 = *reinterpret_cast<const float4*>(dA + (base + W2) * ld_dA + c);
      const float4 a3 = *reinterpret_cast<const float4*>(dA + (base + W2 + 1) * ld_dA + c);
      a.x = (a0.x + a1.x) + (a2.x + a3.x); a.y = (a0.y + a1.y) + (a2.y + a3.y);
      a.z = (a0.z + a1.z) + (a2.z + a3.z); a.w = (a0.w + a1.w) + (a2.w + a3.w);
    } else {
      a = *reinterpret_cast<const float4*>(dA + pix * ld_dA + c);
    }
    if (extra) {
      const float4 e = *reinterpret_cast<const float4*>(extra + pix * ld_extra + c);
      a.x += e.x; a.y += e.y; a.z += e.z; a.w += e.w;
    }
    const float4 yv = *reinterpret_cast<const float4*>(y + pix * ld_y + c);
    const float xh[4] = {(yv.x - m4.x) * r4.x, (yv.y - m4.y) * r4.y, (yv.z - m4.z) * r4.z, (yv.w - m4.w) * r4.w};
    float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (relu && !(xh[e] > 0.f)) av[e] = 0.f;
      sg[e] += av[e];
      sgx[e] += av[e] * xh[e];
    }
    *reinterpret_cast<float4*>(g + pix * C + c) = make_float4(av[0], av[1], av[2], av[3]);
  }
  // reduce over the `sub` lanes that share a channel quad, then over the 8 warps
  for (int o = lpp; o < 32; o <<= 1) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      sg[e] += __shfl_xor_sync(0xffffffffu, sg[e], o);
      sgx[e] += __shfl_xor_sync(0xffffffffu, sgx[e], o);
    }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) { red[warp][lane][e] = sg[e]; red[warp][lane][4 + e] = sgx[e]; }
  __syncthreads();
  if (warp == 0 && lane < lpp && cq_ok) {
    float t[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float s = 0.f;
#pragma unroll
      for (int w8 = 0; w8 < 8; ++w8) s += red[w8][lane][e];
      t[e] = s;
    }
    float* dst = partial + (((size_t)n * gridDim.x + blk) * C + c) * 2;
#pragma unroll
    for (int e = 0; e < 4; ++e) { dst[e * 2] = t[e]; dst[e * 2 + 1] = t[4 + e]; }
  }
}

// per-block (sum, sumsq) of y over pixels: partial [N][blocks][C][2].  Same thread mapping as pass 1.
__global__ void __launch_bounds__(256) in_stats_kernel(const float* __restrict__ y, int ld_y,
                                                       float* __restrict__ partial, long long P, int C, int ppb) {
  __shared__ float red[8][32][8];
  const int n = blockIdx.y, blk = blockIdx.x;
  const int CQ = C >> 2;
  const int cq_base = blockIdx.z * 32;
  const int cq_left = CQ - cq_base;
  int lpp = 32;
  if (cq_left < 32) { lpp = 1; while (lpp < cq_left) lpp <<= 1; }
  const int ppw = 32 / lpp;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = lane % lpp, sub = lane / lpp;
  const bool cq_ok = q < cq_left;
  const int c = (cq_base + q) * 4;
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  const long long p_begin = (long long)blk * ppb;
  for (int i = warp * ppw + sub; i < ppb; i += 8 * ppw) {
    const long long pp = p_begin + i;
    if (pp >= P || !cq_ok) continue;
    const float4 v = *reinterpret_cast<const float4*>(y + ((long long)n * P + pp) * ld_y + c);
    s1[0] += v.x; s1[1] += v.y; s1[2] += v.z; s1[3] += v.w;
    s2[0] += v.x * v.x; s2[1] += v.y * v.y; s2[2] += v.z * v.z; s2[3] += v.w * v.w;
  }
  for (int o = lpp; o < 32; o <<= 1) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      s1[e] += __shfl_xor_sync(0xffffffffu, s1[e], o);
      s2[e] += __shfl_xor_sync(0xffffffffu, s2[e], o);
    }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) { red[warp][lane][e] = s1[e]; red[warp][lane][4 + e] = s2[e]; }
  __syncthreads();
  if (warp == 0 && lane < lpp && cq_ok) {
    float* dst = partial + (((size_t)n * gridDim.x + blk) * C + c) * 2;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float a = 0.f, b = 0.f;
#pragma unroll
      for (int w8 = 0; w8 < 8; ++w8) { a += red[w8][lane][e]; b += red[w8][lane][4 + e]; }
      dst[e * 2] = a;
      dst[e * 2 + 1] = b;
    }
  }
}

// I = index type: unsigned when the element count fits 32 bits (every shape of the decoder: three 64-bit divisions per
// thread made this memory-bound kernel instruction-bound), long long otherwise
template <typename I>
__global__ void in_bwd_pass2_kernel(const float* __restrict__ g, const float* __restrict__ y, int ld_y,
                                    const float* __restrict__ mean, const float* __restrict__ rstd,
                                    const float* __restrict__ gsum, float* __restrict__ dy, I P, int C,
                                    float inv_p, I total, __nv_bfloat16* __restrict__ hi,
                                    __nv_bfloat16* __restrict__ lo, int ld_planes, int c_write) {
  const I idx = (I)blockIdx.x * (I)blockDim.x + (I)threadIdx.x;
  if (idx >= total) return;
  const I CQ = (I)(c_write >> 2);
  const size_t pix = (size_t)(idx / CQ);
  const int c = (int)(idx - (I)pix * CQ) * 4;
  if (c >= C) {  // zero channel padding of the bf16 planes (K padding of the tensor-core operand)
    *reinterpret_cast<uint2*>(hi + pix * ld_planes + c) = make_uint2(0u, 0u);
    *reinterpret_cast<uint2*>(lo + pix * ld_planes + c) = make_uint2(0u, 0u);
    return;
  }
  const int n = (int)((I)pix / P);
  const float4 gv = *reinterpret_cast<const float4*>(g + pix * C + c);
  const float4 yv = *reinterpret_cast<const float4*>(y + pix * ld_y + c);
  const float4 m = *reinterpret_cast<const float4*>(mean + (size_t)n * C + c);
  const float4 r = *reinterpret_cast<const float4*>(rstd + (size_t)n * C + c);
  const float* gs = gsum + ((size_t)n * C + c) * 2;
  const float ga[4] = {gv.x, gv.y, gv.z, gv.w};
  const float ya[4] = {yv.x, yv.y, yv.z, yv.w};
  const float ma[4] = {m.x, m.y, m.z, m.w};
  const float ra[4] = {r.x, r.y, r.z, r.w};
  float o[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float xh = (ya[e] - ma[e]) * ra[e];
    o[e] = ra[e] * (ga[e] - gs[e * 2] * inv_p - xh * (gs[e * 2 + 1] * inv_p));
  }
  if (dy) *reinterpret_cast<float4*>(dy + pix * C + c) = make_float4(o[0], o[1], o[2], o[3]);
  if (hi) {  // the gradient goes straight into the tcgen05 operand format (dgrad / wgrad of the producing conv)
    __align__(8) __nv_bfloat16 h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      h[e] = __float2bfloat16_rn(o[e]);
      l[e] = __float2bfloat16_rn(o[e] - __bfloat162float(h[e]));
    }
    *reinterpret_cast<uint2*>(hi + pix * ld_planes + c) = *reinterpret_cast<const uint2*>(h);
    *reinterpret_cast<uint2*>(lo + pix * ld_planes + c) = *reinterpret_cast<const uint2*>(l);
  }
}

// out[n,y,x,c] (+)= sum of the 2x2 children of in
__global__ void upsample2_bwd_kernel(const float* __restrict__ in, int ld_in, float* __restrict__ out, int ld_out,
                                     int H, int W, int C, int accumulate, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int CQ = C >> 2;
  const int c = (int)(idx % CQ) * 4;
  const long long pix = idx / CQ;
  const long long P = (long long)H * W;
  const int n = (int)(pix / P);
  const long long pp = pix - (long long)n * P;
  const int py = (int)(pp / W), px = (int)(pp - (long long)py * W);
  const int W2 = 2 * W;
  const long long base = ((long long)n * (2 * H) + 2 * py) * W2 + 2 * px;
  const float4 a0 = *reinterpret_cast<const float4*>(in + base * ld_in + c);
  const float4 a1 = *reinterpret_cast<const float4*>(in + (base + 1) * ld_in + c);
  const float4 a2 = *reinterpret_cast<const float4*>(in + (base + W2) * ld_in + c);
  const float4 a3 = *reinterpret_cast<const float4*>(in + (base + W2 + 1) * ld_in + c);
  float4 v = make_float4((a0.x + a1.x) + (a2.x + a3.x), (a0.y + a1.y) + (a2.y + a3.y),
                         (a0.z + a1.z) + (a2.z + a3.z), (a0.w + a1.w) + (a2.w + a3.w));
  float* o = out + pix * ld_out + c;
  if (accumulate) {
    const float4 q = *reinterpret_cast<const float4*>(o);
    v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
  }
  *reinterpret_cast<float4*>(o) = v;
}

// ------------------------------------------------------------------------------- column sums
__global__ void colsum_stage1(const float* __restrict__ x, int ld, long long rows, int C, long long rows_per_block,
                              float* __restrict__ part) {
  __shared__ float red[8][32];
  const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > rows) r1 = rows;
  for (int c0 = 0; c0 < C; c0 += 32) {
    const int c = c0 + lane;
    float s = 0.f;
    if (c < C)
      for (long long r = r0 + wy; r < r1; r += 8) s += x[r * ld + c];
    red[wy][lane] = s;
    __syncthreads();
    if (wy == 0 && c < C) {
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) t += red[k][lane];
      part[(size_t)blockIdx.x * C + c] = t;
    }
    __syncthreads();
  }
}
__global__ void colsum_stage2(const float* __restrict__ part, int nblocks, int C, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s = 0.0;
  for (int b = 0; b < nblocks; ++b) s += (double)part[(size_t)b * C + c];
  out[c] = (float)s;
}

// ------------------------------------------------------------------------- event pre-processing
// stats[w][3] += (sum, sumsq, nnz) of window w; x is [B][T][count] with batch stride bstride.
// grid (chunks, T, B): one block reduces ES_CHUNK consecutive floats of the (b, w) slab -- no index division,
// 128-bit loads when the slab is 16 B aligned; <= 64 values per thread are summed in fp32, everything above in double.
constexpr int ES_THREADS = 256;
constexpr int ES_CHUNK = ES_THREADS * 4 * 16;
__global__ void __launch_bounds__(ES_THREADS) event_stats_kernel(const float* __restrict__ x, long long bstride,
                                                                 long long count, double* __restrict__ stats) {
  __shared__ double red[3][ES_THREADS / 32];
  const int w = blockIdx.y;
  const float* base = x + (long long)blockIdx.z * bstride + (long long)w * count;
  const long long i0 = (long long)blockIdx.x * ES_CHUNK;
  long long i1 = i0 + ES_CHUNK;
  if (i1 > count) i1 = count;
  float f0 = 0.f, f1 = 0.f;
  int nz = 0;
  if ((reinterpret_cast<uintptr_t>(base) & 15u) == 0 && ((i1 - i0) & 3) == 0) {
    const float4* b4 = reinterpret_cast<const float4*>(base + i0);
    const int n4 = (int)((i1 - i0) >> 2);
#pragma unroll 4
    for (int i = threadIdx.x; i < n4; i += ES_THREADS) {
      const float4 v = b4[i];
      f0 += (v.x + v.y) + (v.z + v.w);
      f1 += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
      nz += (v.x != 0.f) + (v.y != 0.f) + (v.z != 0.f) + (v.w != 0.f);
    }
  } else {
    for (long long i = i0 + threadIdx.x; i < i1; i += ES_THREADS) {
      const float v = base[i];
      f0 += v;
      f1 += v * v;
      nz += (v != 0.f);
    }
  }
  double s0 = warp_sum_d((double)f0), s1 = warp_sum_d((double)f1), s2 = warp_sum_d((double)nz);
  const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
  if (lane == 0) { red[0][wy] = s0; red[1][wy] = s1; red[2][wy] = s2; }
  __syncthreads();
  if (threadIdx.x < 3) {
    double t = 0.0;
    for (int k = 0; k < ES_THREADS / 32; ++k) t += red[threadIdx.x][k];
    atomicAdd(&stats[w * 3 + threadIdx.x], t);
  }
}

__device__ __forceinline__ int reflect_idx(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

__global__ void event_prepare_kernel(const float* __restrict__ x, long long bstride, const double* __restrict__ stats,
                                     int normalize, float* __restrict__ out, int ld_out, int C, int H, int W,
                                     int Hp, int Wp, int pad_top, int pad_left, int flip, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ox = (int)(idx % Wp);
  const long long r = idx / Wp;
  const int oy = (int)(r % Hp);
  const int n = (int)(r / Hp);
  int iy = reflect_idx(oy - pad_top, H), ix = reflect_idx(ox - pad_left, W);
  if (flip) { iy = H - 1 - iy; ix = W - 1 - ix; }   // torch.flip(events, dims=[2, 3]) precedes the padding
  float mean = 0.f, inv_std = 1.f;
  bool do_norm = false;
  if (normalize) {
    const double nnz = stats[2];
    if (nnz > 0.0) {
      // fp32 arithmetic on the reduced sums, as the reference does (inference_utils.py:104-107)
      const float fm = (float)stats[0] / (float)nnz;
      const float fs = sqrtf((float)stats[1] / (float)nnz - fm * fm);
      mean = fm;
      inv_std = fs;
      do_norm = true;
    }
  }
  const float* src = x + (long long)n * bstride + (long long)iy * W + ix;
  float* dst = out + idx * ld_out;
  for (int c = 0; c < ld_out; ++c) {
    float v = 0.f;
    if (c < C) {
      v = src[(long long)c * H * W];
      if (do_norm) v = (v != 0.f) ? (v - mean) / inv_std : 0.f;
    }
    dst[c] = v;
  }
}

// ------------------------------------------------------------------------------ layout kernels
// [N][C][P] -> [N][P][ld]   (32x32 smem transpose)
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out, int ld_out, int C,
                                    long long P) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const long long p0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j;
    const long long pp = p0 + tx;
    tile[j][tx] = (c < C && pp < P) ? in[((long long)n * C + c) * P + pp] : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const long long pp = p0 + j;
    const int c = c0 + tx;
    if (pp < P && c < C) out[((long long)n * P + pp) * ld_out + c] = tile[tx][j];
  }
}
__global__ void nhwc_to_nchw_kernel(const float* __restrict__ in, int ld_in, float* __restrict__ out, int C,
                                    long long P) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const long long p0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int j = ty; j < 32; j += 8) {
    const long long pp = p0 + j;
    const int c = c0 + tx;
    tile[j][tx] = (c < C && pp < P) ? in[((long long)n * P + pp) * ld_in + c] : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j;
    const long long pp = p0 + tx;
    if (c < C && pp < P) out[((long long)n * C + c) * P + pp] = tile[tx][j];
  }
}

// bilinear x2, align_corners=False (ATen area_pixel_compute_source_index semantics)
__global__ void bilinear_up2_kernel(const float* __restrict__ in, float* __restrict__ out, int H, int W, int C,
                                    long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx % C);
  long long r = idx / C;
  const int ox = (int)(r % (2 * W)); r /= (2 * W);
  const int oy = (int)(r % (2 * H));
  const int n = (int)(r / (2 * H));
  float sy = 0.5f * (oy + 0.5f) - 0.5f; if (sy < 0.f) sy = 0.f;
  float sx = 0.5f * (ox + 0.5f) - 0.5f; if (sx < 0.f) sx = 0.f;
  const int y0 = (int)sy, x0 = (int)sx;
  const int y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
  const float ly1 = sy - y0, ly0 = 1.f - ly1, lx1 = sx - x0, lx0 = 1.f - lx1;
  const float* b = in + (long long)n * H * W * C + c;
  const float v00 = b[((long long)y0 * W + x0) * C], v01 = b[((long long)y0 * W + x1) * C];
  const float v10 = b[((long long)y1 * W + x0) * C], v11 = b[((long long)y1 * W + x1) * C];
  out[idx] = ly0 * (lx0 * v00 + lx1 * v01) + ly1 * (lx0 * v10 + lx1 * v11);
}

// ------------------------------------------------------------------------------------ RAdam
__global__ void radam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, long long n, float beta1, float beta2, float step_lr, float eps,
                             float wd_lr, int rectified) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i];
  const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
  const float mi = beta1 * m[i] + (1.f - beta1) * gi;
  v[i] = vi;
  m[i] = mi;
  float pi = p[i];
  if (wd_lr != 0.f) pi += -wd_lr * pi;
  if (rectified) pi += -step_lr * (mi / (sqrtf(vi) + eps));
  else pi += -step_lr * mi;
  p[i] = pi;
}

// Multi-tensor form: ONE launch updates up to ESSB_RADAM_MAX tensors (the 34 decoder tensors of the supervised step).
// Block b works on a 4096-element chunk of the tensor whose block range [blk0[t], blk0[t+1]) contains b.
struct RadamMultiParams {
  essb_radam_multi d;
  int blk0[ESSB_RADAM_MAX + 1];
};
constexpr int RADAM_CHUNK = 4096;
__global__ void __launch_bounds__(256) radam_multi_kernel(const __grid_constant__ RadamMultiParams P) {
  __shared__ int s_t;
  if (threadIdx.x == 0) {
    int lo = 0, hi = P.d.count;              // largest t with blk0[t] <= blockIdx.x
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (P.blk0[mid] <= (int)blockIdx.x) lo = mid; else hi = mid;
    }
    s_t = lo;
  }
  __syncthreads();
  const int t = s_t;
  const long long n = P.d.n[t];
  const long long base = (long long)(blockIdx.x - P.blk0[t]) * RADAM_CHUNK;
  float* __restrict__ p = P.d.p[t];
  const float* __restrict__ g = P.d.g[t];
  float* __restrict__ m = P.d.m[t];
  float* __restrict__ v = P.d.v[t];
  const float beta1 = P.d.beta1, beta2 = P.d.beta2, step_lr = P.d.step_lr, eps = P.d.eps, wd_lr = P.d.wd_lr;
  const int rectified = P.d.rectified;
  auto upd = [&](float gi, float& mi, float& vi, float& pi) {
    vi = beta2 * vi + (1.f - beta2) * gi * gi;
    mi = beta1 * mi + (1.f - beta1) * gi;
    if (wd_lr != 0.f) pi += -wd_lr * pi;
    if (rectified) pi += -step_lr * (mi / (sqrtf(vi) + eps));
    else pi += -step_lr * mi;
  };
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15u) == 0 && (n & 3) == 0;
  if (vec) {
#pragma unroll
    for (int r = 0; r < RADAM_CHUNK / (256 * 4); ++r) {
      const long long i = base + ((long long)r * 256 + threadIdx.x) * 4;
      if (i >= n) break;
      const float4 g4 = *reinterpret_cast<const float4*>(g + i);
      float4 m4 = *reinterpret_cast<float4*>(m + i), v4 = *reinterpret_cast<float4*>(v + i);
      float4 p4 = *reinterpret_cast<float4*>(p + i);
      upd(g4.x, m4.x, v4.x, p4.x); upd(g4.y, m4.y, v4.y, p4.y); upd(g4.z, m4.z, v4.z, p4.z); upd(g4.w, m4.w, v4.w, p4.w);
      *reinterpret_cast<float4*>(m + i) = m4;
      *reinterpret_cast<float4*>(v + i) = v4;
      *reinterpret_cast<float4*>(p + i) = p4;
    }
  } else {
    for (int r = 0; r < RADAM_CHUNK / 256; ++r) {
      const long long i = base + (long long)r * 256 + threadIdx.x;
      if (i >= n) break;
      float mi = m[i], vi = v[i], pi = p[i];
      upd(g[i], mi, vi, pi);
      m[i] = mi; v[i] = vi; p[i] = pi;
    }
  }
}

}  // namespace

// =================================================================================== C ABI
extern "C" int essb_radam_multi_step(const essb_radam_multi* d, void* stream) {
  ESSB_REQUIRE(d && d->count > 0 && d->count <= ESSB_RADAM_MAX, "essb_radam_multi_step: count must be in [1, %d]", ESSB_RADAM_MAX);
  static thread_local RadamMultiParams P;
  P.d = *d;
  int blocks = 0;
  for (int t = 0; t < d->count; ++t) {
    ESSB_REQUIRE(d->p[t] && d->g[t] && d->m[t] && d->v[t] && d->n[t] > 0, "essb_radam_multi_step: bad tensor %d", t);
    P.blk0[t] = blocks;
    const long long nb = (d->n[t] + RADAM_CHUNK - 1) / RADAM_CHUNK;
    ESSB_REQUIRE(blocks + nb < (1ll << 30), "essb_radam_multi_step: too many elements");
    blocks += (int)nb;
  }
  P.blk0[d->count] = blocks;
  radam_multi_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(P);
  ESSB_LAUNCH_CHECK("essb_radam_multi_step");
  return ESSB_OK;
}

extern "C" int essb_radam_step(float* p, const float* g, float* m, float* v, int64_t n, float beta1, float beta2,
                               float step_lr, float eps, float wd_lr, int rectified, void* stream) {
  ESSB_REQUIRE(p && g && m && v && n > 0, "essb_radam_step: bad arguments");
  radam_kernel<<<ew_blocks(n), EW_THREADS, 0, (cudaStream_t)stream>>>(p, g, m, v, n, beta1, beta2, step_lr, eps, wd_lr,
                                                                     rectified);
  ESSB_LAUNCH_CHECK("essb_radam_step");
  return ESSB_OK;
}

extern "C" int essb_pack_weight(const float* w, const float* scale, float* out, int Cout, int Cin, int T,
                                int transposed_layout, int swap_io, int flip, int interleave, void* stream) {
  ESSB_REQUIRE(w && out && Cout > 0 && Cin > 0 && T > 0, "essb_pack_weight: bad arguments");
  ESSB_REQUIRE(interleave <= 1 || (Cout % interleave == 0 && !swap_io), "essb_pack_weight: bad interleave");
  const int Kin = swap_io ? Cout : Cin;
  const int NoutP = ((swap_io ? Cin : Cout) + 3) & ~3;
  const long long total = (long long)T * Kin * NoutP;
  pack_weight_kernel<<<ew_blocks(total), EW_THREADS, 0, (cudaStream_t)stream>>>(w, scale, out, Cout, Cin, T,
                                                                               transposed_layout, swap_io, flip,
                                                                               interleave);
  ESSB_LAUNCH_CHECK("essb_pack_weight");
  return ESSB_OK;
}

extern "C" int essb_in_finalize(const float* partial, int N, int tiles, int C, int64_t count, float eps, float* mean,
                                float* rstd, void* stream) {
  ESSB_REQUIRE(partial && mean && rstd && N > 0 && tiles > 0 && C > 0 && count > 0, "essb_in_finalize: bad arguments");
  const long long threads = (long long)N * C * 32;
  in_finalize_kernel<<<ew_blocks(threads), EW_THREADS, 0, (cudaStream_t)stream>>>(partial, N, tiles, C,
                                                                                 1.0 / (double)count, eps, mean, rstd,
                                                                                 nullptr);
  ESSB_LAUNCH_CHECK("essb_in_finalize");
  return ESSB_OK;
}

extern "C" int essb_partial_reduce(const float* partial, int N, int blocks, int C, float* totals, void* stream) {
  ESSB_REQUIRE(partial && totals && N > 0 && blocks > 0 && C > 0, "essb_partial_reduce: bad arguments");
  const long long threads = (long long)N * C * 32;
  in_finalize_kernel<<<ew_blocks(threads), EW_THREADS, 0, (cudaStream_t)stream>>>(partial, N, blocks, C, 0.0, 0.f,
                                                                                 nullptr, nullptr, totals);
  ESSB_LAUNCH_CHECK("essb_partial_reduce");
  return ESSB_OK;
}

static int check_vec4(const void* p, int ld, const char* who) {
  ESSB_REQUIRE(p == nullptr || (essb_aligned16(p) && ld % 4 == 0), "%s: tensors must be 16B aligned with ld %% 4 == 0", who);
  return ESSB_OK;
}

extern "C" int essb_norm_act_add(const float* y, int ld_y, const float* mean, const float* rstd, int relu,
                                 const float* res, int ld_res, float* out, int ld_out, int N, int64_t P, int C,
                                 void* stream) {
  ESSB_REQUIRE(y && out && N > 0 && P > 0 && C > 0 && C % 4 == 0, "essb_norm_act_add: bad arguments (C %% 4 == 0 required)");
  int rc;
  if ((rc = check_vec4(y, ld_y, "essb_norm_act_add")) || (rc = check_vec4(res, ld_res, "essb_norm_act_add")) ||
      (rc = check_vec4(out, ld_out, "essb_norm_act_add")) || (rc = check_vec4(mean, 4, "essb_norm_act_add")))
    return rc;
  const long long total = (long long)N * P * (C / 4);
  norm_act_add_kernel<<<ew_blocks(total), EW_THREADS, 0, (cudaStream_t)stream>>>(y, ld_y, mean, rstd, relu, res,
                                                                                ld_res, out, ld_out, P, C, total);
  ESSB_LAUNCH_CHECK("essb_norm_act_add");
  return ESSB_OK;
}

extern "C" int essb_in_stats(const float* y, int ld_y, float* partial, int N, int64_t P, int C, void* stream) {
  ESSB_REQUIRE(y && partial && N > 0 && P > 0 && C > 0 && C % 4 == 0, "essb_in_stats: bad arguments (C %% 4 == 0)");
  int rc;
  if ((rc = check_vec4(y, ld_y, "essb_in_stats"))) return rc;
  const int blocks = essb_in_bwd_blocks(P);
  dim3 grid(blocks, N, (C / 4 + 31) / 32);
  in_stats_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(y, ld_y, partial, P, C, in_pix_per_block(P));
  ESSB_LAUNCH_CHECK("essb_in_stats");
  return ESSB_OK;
}

extern "C" int essb_in_bwd_blocks(int64_t P) {
  const int ppb = in_pix_per_block(P);
  return (int)((P + ppb - 1) / ppb);
}

extern "C" int essb_in_bwd_pass1(const float* dA, int ld_dA, int ups, const float* extra, int ld_extra,
                                 const float* y, int ld_y, const float* mean, const float* rstd, int relu, float* g,
                                 float* partial, int N, int H, int W, int C, void* stream) {
  ESSB_REQUIRE(dA && y && g && partial && N > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0,
               "essb_in_bwd_pass1: bad arguments (C %% 4 == 0 required)");
  ESSB_REQUIRE((mean == nullptr) == (rstd == nullptr), "essb_in_bwd_pass1: mean/rstd must come together");
  int rc;
  if ((rc = check_vec4(dA, ld_dA, "essb_in_bwd_pass1")) || (rc = check_vec4(extra, ld_extra, "essb_in_bwd_pass1")) ||
      (rc = check_vec4(y, ld_y, "essb_in_bwd_pass1")) || (rc = check_vec4(g, 4, "essb_in_bwd_pass1")))
    return rc;
  const int blocks = essb_in_bwd_blocks((int64_t)H * W);
  dim3 grid(blocks, N, (C / 4 + 31) / 32);
  in_bwd_pass1_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dA, ld_dA, ups, extra, ld_extra, y, ld_y, mean, rstd,
                                                              relu, g, partial, H, W, C,
                                                              in_pix_per_block((long long)H * W));
  ESSB_LAUNCH_CHECK("essb_in_bwd_pass1");
  return ESSB_OK;
}

extern "C" int essb_in_bwd_pass2(const float* g, const float* y, int ld_y, const float* mean, const float* rstd,
                                 const float* gsum, float* dy, uint16_t* dy_hi, uint16_t* dy_lo, int ld_planes, int N,
                                 int64_t P, int C, void* stream) {
  ESSB_REQUIRE(g && y && mean && rstd && gsum && (dy || dy_hi) && N > 0 && P > 0 && C % 4 == 0,
               "essb_in_bwd_pass2: bad arguments");
  ESSB_REQUIRE(!dy_hi || (dy_lo && ld_planes >= C && ld_planes % 4 == 0 && (reinterpret_cast<uintptr_t>(dy_hi) & 7u) == 0 &&
                          (reinterpret_cast<uintptr_t>(dy_lo) & 7u) == 0),
               "essb_in_bwd_pass2: bad planes");
  int rc;
  if ((rc = check_vec4(y, ld_y, "essb_in_bwd_pass2")) || (rc = check_vec4(g, 4, "essb_in_bwd_pass2")) ||
      (dy && (rc = check_vec4(dy, 4, "essb_in_bwd_pass2"))))
    return rc;
  const int c_write = dy_hi ? ld_planes : C;   // planes are written across their whole pitch (zeros beyond C)
  const long long total = (long long)N * P * (c_write / 4);
  if (total < (1ll << 31))
    in_bwd_pass2_kernel<unsigned><<<ew_blocks(total), EW_THREADS, 0, (cudaStream_t)stream>>>(
        g, y, ld_y, mean, rstd, gsum, dy, (unsigned)P, C, 1.0f / (float)P, (unsigned)total, reinterpret_cast<__nv_bfloat16*>(dy_hi),
        reinterpret_cast<__nv_bfloat16*>(dy_lo), ld_planes, c_write);
  else
    in_bwd_pass2_kernel<long long><<<ew_blocks(total), EW_THREADS, 0, (cudaStream_t)stream>>>(
        g, y, ld_y, mean, rstd, gsum, dy, (long long)P, C, 1.0f / (float)P, total, reinterpret_cast<__nv_bfloat16*>(dy_hi),
        reinterpret_cast<__nv_bfloat16*>(dy_lo), ld_planes, c_write);
  ESSB_LAUNCH_CHECK("essb_in_bwd_pass2");
  return ESSB_OK;
}

extern "C" int essb_upsample2_bwd(const float* in, int ld_in, float* out, int ld_out, int N, int H, int W, int C,
                                  int accumulate, void* stream) {
  ESSB_REQUIRE(in && out && N > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0, "essb_upsample2_bwd: bad arguments");
  int rc;
  if ((rc = check_vec4(in, ld_in, "essb_upsample2_bwd")) || (rc = check_vec4(out, ld_out, "essb_upsample2_bwd")))
    return rc;
  const long long total = (long long)N * H * W * (C / 4);
  upsample2_bwd_kernel<<<ew_blocks(total), EW_THREADS, 0, (cudaStream_t)stream>>>(in, ld_in, out, ld_out, H, W, C,
                                                                                 accumulate, total);
  ESSB_LAUNCH_CHECK("essb_upsample2_bwd");
  return ESSB_OK;
}

extern "C" int essb_colsum(const float* x, int ld, int64_t rows, int C, float* out, float* workspace,
                           int64_t workspace_bytes, void* stream) {
  ESSB_REQUIRE(x && out && workspace && rows > 0 && C > 0, "essb_colsum: bad arguments");
  long long nb = (rows + 511) / 512;
  if (nb > 1024) nb = 1024;
  const long long fit = workspace_bytes / ((long long)C * (long long)sizeof(float));
  if (nb > fit) nb = fit;
  if (nb < 1) {
    essb_set_error("essb_colsum: workspace too small");
    return ESSB_ERR_WORKSPACE;
  }
  const long long rpb = (rows + nb - 1) / nb;
  nb = (rows + rpb - 1) / rpb;
  colsum_stage1<<<(unsigned)nb, 256, 0, (cudaStream_t)stream>>>(x, ld, rows, C, rpb, workspace);
  ESSB_LAUNCH_CHECK("essb_colsum stage1");
  colsum_stage2<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(workspace, (int)nb, C, out);
  ESSB_LAUNCH_CHECK("essb_colsum stage2");
  return ESSB_OK;
}

// x[b][c][y][x] = 0 for every listed pixel (hot-pixel removal, inference_utils.py:88-89); in place, as the reference
__global__ void zero_pixels_kernel(float* __restrict__ x, long long bstride, int B, int C, int H, int W,
                                   const int* __restrict__ xy, int n) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)n * B * C;
  if (idx >= total) return;
  const int i = (int)(idx % n);
  const long long r = idx / n;
  const int c = (int)(r % C), b = (int)(r / C);
  const int px = xy[2 * i], py = xy[2 * i + 1];
  if (px < 0 || px >= W || py < 0 || py >= H) return;
  x[(long long)b * bstride + ((long long)c * H + py) * W + px] = 0.f;
}

extern "C" int essb_zero_pixels(float* x, int64_t bstride, int B, int C, int H, int W, const int32_t* xy, int n,
                                void* stream) {
  ESSB_REQUIRE(x && B > 0 && C > 0 && H > 0 && W > 0 && n >= 0 && (n == 0 || xy), "essb_zero_pixels: bad arguments");
  if (n == 0) return ESSB_OK;
  const long long total = (long long)n * B * C;
  zero_pixels_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, bstride, B, C, H, W, xy, n);
  ESSB_LAUNCH_CHECK("essb_zero_pixels");
  return ESSB_OK;
}

extern "C" int essb_event_stats(const float* x, int64_t bstride, int B, int T, int64_t count, double* stats,
                                void* stream) {
  ESSB_REQUIRE(x && stats && B > 0 && T > 0 && count > 0, "essb_event_stats: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(stats, 0, sizeof(double) * 3 * T, st);
  if (e != cudaSuccess) {
    essb_set_error("essb_event_stats: memset failed: %s", cudaGetErrorString(e));
    return ESSB_ERR_LAUNCH;
  }
  ESSB_REQUIRE(B <= 65535 && T <= 65535, "essb_event_stats: B and T must fit a grid dimension");
  dim3 grid((unsigned)((count + ES_CHUNK - 1) / ES_CHUNK), T, B);
  event_stats_kernel<<<grid, ES_THREADS, 0, st>>>(x, bstride, count, stats);
  ESSB_LAUNCH_CHECK("essb_event_stats");
  return ESSB_OK;
}

extern "C" int essb_event_prepare(const float* x, int64_t bstride, const double* stats, int normalize, float* out,
                                  int ld_out, int B, int C, int H, int W, int Hp, int Wp, int pad_top, int pad_left,
                                  int flip, void* stream) {
  ESSB_REQUIRE(x && out && B > 0 && C > 0 && H > 0 && W > 0 && Hp >= H && Wp >= W && ld_out >= C,
               "essb_event_prepare: bad arguments");
  ESSB_REQUIRE(!normalize || stats, "essb_event_prepare: stats required when normalising");
  ESSB_REQUIRE(pad_top < H && Hp - H - pad_top < H && pad_left < W && Wp - W - pad_left < W,
               "essb_event_prepare: reflection padding must be smaller than the image");
  const long long total = (long long)B * Hp * Wp;
  event_prepare_kernel<<<ew_blocks(total), EW_THREADS, 0, (cudaStream_t)stream>>>(
      x, bstride, stats, normalize, out, ld_out, C, H, W, Hp, Wp, pad_top, pad_left, flip, total);
  ESSB_LAUNCH_CHECK("essb_event_prepare");
  return ESSB_OK;
}

extern "C" int essb_nchw_to_nhwc(const float* in, float* out, int ld_out, int N, int C, int64_t P, void* stream) {
  ESSB_REQUIRE(in && out && N > 0 && C > 0 && P > 0 && ld_out >= C && N <= 65535, "essb_nchw_to_nhwc: bad arguments");
  dim3 grid((unsigned)((P + 31) / 32), (C + 31) / 32, N), block(32, 8, 1);
  nchw_to_nhwc_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(in, out, ld_out, C, P);
  ESSB_LAUNCH_CHECK("essb_nchw_to_nhwc");
  return ESSB_OK;
}

extern "C" int essb_nhwc_to_nchw(const float* in, int ld_in, float* out, int N, int C, int64_t P, void* stream) {
  ESSB_REQUIRE(in && out && N > 0 && C > 0 && P > 0 && ld_in >= C && N <= 65535, "essb_nhwc_to_nchw: bad arguments");
  dim3 grid((unsigned)((P + 31) / 32), (C + 31) / 32, N), block(32, 8, 1);
  nhwc_to_nchw_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(in, ld_in, out, C, P);
  ESSB_LAUNCH_CHECK("essb_nhwc_to_nchw");
  return ESSB_OK;
}

extern "C" int essb_bilinear_up2(const float* in, float* out, int N, int H, int W, int C, void* stream) {
  ESSB_REQUIRE(in && out && N > 0 && H > 0 && W > 0 && C > 0, "essb_bilinear_up2: bad arguments");
  const long long total = (long long)N * 4 * H * W * C;
  bilinear_up2_kernel<<<ew_blocks(total), EW_THREADS, 0, (cudaStream_t)stream>>>(in, out, H, W, C, total);
  ESSB_LAUNCH_CHECK("essb_bilinear_up2");
  return ESSB_OK;
}
