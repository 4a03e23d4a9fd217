// fp32 CUDA-core implicit-GEMM gather-convolution (forward / input-gradient) and weight-gradient.
//
// This is the exact-fp32 path ("fp32" mode): every product is an fp32 FMA, so results match the
// reference's ATen/oneDNN fp32 convolutions to accumulation-order noise.  It covers every conv
// shape on the hot path (3x3 s1, 5x5 s1/s2, 4-phase transposed 5x5 s2, 1x1, dgrad of all of them)
// with one kernel; the tcgen05 kernel in conv_tc.cu accelerates the shapes that dominate the FLOPs.
//
// GEMM view: M = output pixels of one sample (tiles of 128), N = Cout (tiles of 64/128),
// K = taps x input channels (steps of 16 channels of one tap).  NHWC makes each K-step of the A
// operand a contiguous 64-byte run per pixel; the loader applies upsample / instance-norm / ReLU /
// concat on the fly (see essb_src in ess_b200.h).
#include <cuda_bf16.h>

#include "common.cuh"

namespace {

constexpr int BM = 128;       // pixels per CTA tile
constexpr int BK = 16;        // input channels per K-step
constexpr int AS_LD = BM + 4; // padded smem row (floats); 16B-aligned rows, 2-way store conflicts at most
constexpr int NTHREADS = 256;

struct ConvParams {
  essb_conv d;
  int cin_total;
  int coutp;
  int chunks0, chunks1;  // K-steps per tap for segment 0 / 1
  int vec0, vec1;        // float4 loads legal for the segment
  int vec_out;           // float4 stores legal
};

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// Load 4 consecutive channels [c, c+4) of one virtual input pixel, transformed; zero outside.
__device__ __forceinline__ float4 load_a4(const essb_src& s, int vec, int n, int iy, int ix, int H, int W,
                                          int c) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if ((unsigned)iy >= (unsigned)H || (unsigned)ix >= (unsigned)W || c >= s.C) return v;
  const int Hs = H >> s.ups, Ws = W >> s.ups;
  const size_t pix = ((size_t)n * Hs + (iy >> s.ups)) * Ws + (ix >> s.ups);
  const float* q = s.ptr + pix * (size_t)s.ld + c;
  if (vec) {
    v = ld4(q);
    if (s.mean) {
      const float4 m = ld4(s.mean + (size_t)n * s.C + c);
      const float4 r = ld4(s.rstd + (size_t)n * s.C + c);
      v.x = (v.x - m.x) * r.x; v.y = (v.y - m.y) * r.y; v.z = (v.z - m.z) * r.z; v.w = (v.w - m.w) * r.w;
    }
  } else {
    float t[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (c + e < s.C) {
        float x = q[e];
        if (s.mean) x = (x - s.mean[(size_t)n * s.C + c + e]) * s.rstd[(size_t)n * s.C + c + e];
        t[e] = x;
      }
    }
    v = make_float4(t[0], t[1], t[2], t[3]);
  }
  if (s.relu) {
    v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
  }
  return v;
}

template <int EPI, int TN>  // TN = output channels per thread (4 or 8); CTA N-tile = 16*TN
__global__ void __launch_bounds__(NTHREADS) conv_fp32_kernel(const __grid_constant__ ConvParams p) {
  constexpr int BN = 16 * TN;
  constexpr int NB4 = TN / 4;  // float4 groups per thread: couts n0 + g*64 + tx*4 + [0,4)
  __shared__ __align__(16) float As[2][BK][AS_LD];
  __shared__ __align__(16) float Bs[2][BK][BN];

  const essb_conv& d = p.d;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int tile = blockIdx.x, n = blockIdx.z;
  const int n0 = blockIdx.y * BN;
  const int npix = d.OH * d.OW;

  // A-load role: thread loads float4 #f of the 16-channel run of pixels m0 and m0+64.
  const int f = tid & 3;
  int l_oy[2], l_ox[2];
  bool l_ok[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int pp = tile * BM + (tid >> 2) + 64 * j;
    l_ok[j] = pp < npix;
    l_oy[j] = l_ok[j] ? pp / d.OW : 0;
    l_ox[j] = l_ok[j] ? pp - l_oy[j] * d.OW : 0;
  }
  // B-load role: row kb of the K-step, float4 columns.
  const int kb = tid >> 4;

  const int cpt = p.chunks0 + p.chunks1;
  const int iters = d.ntaps * cpt;

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float4 ra[2];
  float4 rb[NB4];

  auto gload = [&](int it) {
    const int tap = it / cpt;
    const int r = it - tap * cpt;
    const int seg = r >= p.chunks0 ? 1 : 0;
    const int c0 = (seg ? r - p.chunks0 : r) * BK;
    const essb_src& s = d.src[seg];
    const int vec = seg ? p.vec1 : p.vec0;
    const int dyv = d.dy[tap], dxv = d.dx[tap];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      ra[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (l_ok[j]) ra[j] = load_a4(s, vec, n, l_oy[j] * d.sy + dyv, l_ox[j] * d.sx + dxv, d.H, d.W, c0 + f * 4);
    }
    const int cseg = c0 + kb;
    const int cg = (seg ? d.src[0].C : 0) + cseg;
    const float* wrow = d.w + ((size_t)d.widx[tap] * p.cin_total + cg) * p.coutp;
#pragma unroll
    for (int g = 0; g < NB4; ++g) {
      const int co = n0 + g * 64 + tx * 4;
      rb[g] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (cseg < s.C && co < p.coutp) rb[g] = ld4(wrow + co);
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int m = (tid >> 2) + 64 * j;
      As[buf][f * 4 + 0][m] = ra[j].x;
      As[buf][f * 4 + 1][m] = ra[j].y;
      As[buf][f * 4 + 2][m] = ra[j].z;
      As[buf][f * 4 + 3][m] = ra[j].w;
    }
#pragma unroll
    for (int g = 0; g < NB4; ++g) *reinterpret_cast<float4*>(&Bs[buf][kb][g * 64 + tx * 4]) = rb[g];
  };

  gload(0);
  sstore(0);
  __syncthreads();
  for (int it = 0; it < iters; ++it) {
    const int buf = it & 1;
    if (it + 1 < iters) gload(it + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[TN];
#pragma unroll
      for (int g = 0; g < NB4; ++g) {
        const float4 bv = *reinterpret_cast<const float4*>(&Bs[buf][k][g * 64 + tx * 4]);
        b[g * 4 + 0] = bv.x; b[g * 4 + 1] = bv.y; b[g * 4 + 2] = bv.z; b[g * 4 + 3] = bv.w;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (it + 1 < iters) sstore(buf ^ 1);
    __syncthreads();
  }

  // ------------------------------------------------------------------ epilogues
  if constexpr (EPI == ESSB_EPI_LINEAR) {
    float s1[TN], s2[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int pp = tile * BM + ty * 8 + i;
      if (pp >= npix) continue;
      const int oy = pp / d.OW, ox = pp - oy * d.OW;
      const size_t opix = ((size_t)n * d.OHf + (oy * d.osy + d.ooy)) * d.OWf + (ox * d.osx + d.oox);
#pragma unroll
      for (int g = 0; g < NB4; ++g) {
        const int co = n0 + g * 64 + tx * 4;
        if (co >= d.Cout) continue;
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float x = acc[i][g * 4 + e];
          if (co + e < d.Cout) {
            if (d.bias) x += d.bias[co + e];
            if (d.res_pre) x += d.res_pre[opix * d.ld_res + co + e];
            if (d.act == ESSB_ACT_RELU) x = fmaxf(x, 0.f);
            else if (d.act == ESSB_ACT_SIGMOID) x = essb_sigmoid(x);
            if (d.res_post) x += d.res_post[opix * d.ld_res + co + e];
            s1[g * 4 + e] += x;
            s2[g * 4 + e] += x * x;
          }
          v[e] = x;
        }
        if (d.out_hi && co + 3 < d.Cout) {
          __nv_bfloat16 h[4], l[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            h[e] = __float2bfloat16_rn(v[e]);
            l[e] = __float2bfloat16_rn(v[e] - __bfloat162float(h[e]));
          }
          *reinterpret_cast<uint2*>(d.out_hi + opix * d.ld_planes + co) = *reinterpret_cast<uint2*>(h);
          *reinterpret_cast<uint2*>(d.out_lo + opix * d.ld_planes + co) = *reinterpret_cast<uint2*>(l);
        }
        float* o = d.out + opix * d.ldo + co;
        if (p.vec_out && co + 3 < d.Cout) {
          float4 w4 = make_float4(v[0], v[1], v[2], v[3]);
          if (d.accumulate) {
            const float4 old = ld4(o);
            w4.x += old.x; w4.y += old.y; w4.z += old.z; w4.w += old.w;
          }
          *reinterpret_cast<float4*>(o) = w4;
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (co + e < d.Cout) o[e] = d.accumulate ? o[e] + v[e] : v[e];
        }
      }
    }
    if (d.stats_partial) {
      // deterministic per-tile (sum, sumsq): reduce over the 16 ty-rows through shared memory
      float* red = &As[0][0][0];  // 16 x BN x 2 floats <= 2*16*132 floats
      __syncthreads();
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int col = (j >> 2) * 64 + tx * 4 + (j & 3);
        red[(ty * BN + col) * 2 + 0] = s1[j];
        red[(ty * BN + col) * 2 + 1] = s2[j];
      }
      __syncthreads();
      if (tid < BN) {
        const int co = n0 + tid;
        if (co < d.Cout) {
          float a1 = 0.f, a2 = 0.f;
#pragma unroll
          for (int r = 0; r < 16; ++r) { a1 += red[(r * BN + tid) * 2]; a2 += red[(r * BN + tid) * 2 + 1]; }
          float* dst = d.stats_partial + (((size_t)n * gridDim.x + tile) * d.Cout + co) * 2;
          dst[0] = a1;
          dst[1] = a2;
        }
      }
    }
  } else if constexpr (EPI == ESSB_EPI_LSTM) {
    const int hidden = d.Cout >> 2;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int pp = tile * BM + ty * 8 + i;
      if (pp >= npix) continue;
      const size_t pix = (size_t)n * npix + pp;
#pragma unroll
      for (int g = 0; g < NB4; ++g) {
        const int co = n0 + g * 64 + tx * 4;
        if (co >= d.Cout) continue;
        const int ch = co >> 2;
        float gi = acc[i][g * 4 + 0], gf = acc[i][g * 4 + 1], go = acc[i][g * 4 + 2], gc = acc[i][g * 4 + 3];
        if (d.bias) { gi += d.bias[co]; gf += d.bias[co + 1]; go += d.bias[co + 2]; gc += d.bias[co + 3]; }
        const float cprev = d.aux0 ? d.aux0[pix * hidden + ch] : 0.f;
        const float cell = essb_sigmoid(gf) * cprev + essb_sigmoid(gi) * tanhf(gc);
        d.out[pix * hidden + ch] = essb_sigmoid(go) * tanhf(cell);
        d.out2[pix * hidden + ch] = cell;
      }
    }
  } else if constexpr (EPI == ESSB_EPI_GRU_UR) {
    const int hidden = d.Cout >> 1;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int pp = tile * BM + ty * 8 + i;
      if (pp >= npix) continue;
      const size_t pix = (size_t)n * npix + pp;
#pragma unroll
      for (int g = 0; g < NB4; ++g) {
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          const int co = n0 + g * 64 + tx * 4 + h2 * 2;
          if (co >= d.Cout) continue;
          const int ch = co >> 1;
          float gu = acc[i][g * 4 + h2 * 2], gr = acc[i][g * 4 + h2 * 2 + 1];
          if (d.bias) { gu += d.bias[co]; gr += d.bias[co + 1]; }
          const float hp = d.aux0 ? d.aux0[pix * hidden + ch] : 0.f;
          d.out[pix * hidden + ch] = essb_sigmoid(gu);
          d.out2[pix * hidden + ch] = hp * essb_sigmoid(gr);
        }
      }
    }
  } else {  // ESSB_EPI_GRU_OUT
    const int hidden = d.Cout;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int pp = tile * BM + ty * 8 + i;
      if (pp >= npix) continue;
      const size_t pix = (size_t)n * npix + pp;
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int co = n0 + (j >> 2) * 64 + tx * 4 + (j & 3);
        if (co >= d.Cout) continue;
        float x = acc[i][j];
        if (d.bias) x += d.bias[co];
        const float hp = d.aux0 ? d.aux0[pix * hidden + co] : 0.f;
        const float u = d.aux1[pix * hidden + co];
        d.out[pix * hidden + co] = hp * (1.f - u) + tanhf(x) * u;
      }
    }
  }
}

template <int EPI>
int launch_conv(const ConvParams& p, cudaStream_t st) {
  const essb_conv& d = p.d;
  const int tiles = (d.OH * d.OW + BM - 1) / BM;
  if (d.Cout > 64) {
    dim3 grid(tiles, (d.Cout + 127) / 128, d.N);
    conv_fp32_kernel<EPI, 8><<<grid, NTHREADS, 0, st>>>(p);
  } else {
    dim3 grid(tiles, 1, d.N);
    conv_fp32_kernel<EPI, 4><<<grid, NTHREADS, 0, st>>>(p);
  }
  ESSB_LAUNCH_CHECK("essb_conv_fp32");
  return ESSB_OK;
}

bool src_vec_ok(const essb_src& s) {
  if (!s.ptr) return true;
  bool ok = (s.C % 4 == 0) && (s.ld % 4 == 0) && essb_aligned16(s.ptr);
  if (s.mean) ok = ok && essb_aligned16(s.mean) && essb_aligned16(s.rstd);
  return ok;
}

int check_src(const essb_src& s, int H, int W, const char* who) {
  if (!s.ptr) return ESSB_OK;
  ESSB_REQUIRE(s.C > 0 && s.ld >= s.C, "%s: bad segment C=%d ld=%d", who, s.C, s.ld);
  ESSB_REQUIRE(s.ups == 0 || s.ups == 1, "%s: ups must be 0 or 1", who);
  ESSB_REQUIRE(!s.ups || (H % 2 == 0 && W % 2 == 0), "%s: upsampled segment needs even H, W", who);
  ESSB_REQUIRE((s.mean == nullptr) == (s.rstd == nullptr), "%s: mean/rstd must come together", who);
  return ESSB_OK;
}

}  // namespace

extern "C" int essb_conv_tiles_per_sample(const essb_conv* d) { return (d->OH * d->OW + BM - 1) / BM; }

extern "C" int essb_conv_fp32(const essb_conv* dp, void* stream) {
  ESSB_REQUIRE(dp != nullptr, "essb_conv_fp32: null descriptor");
  ConvParams p;
  p.d = *dp;
  essb_conv& d = p.d;
  ESSB_REQUIRE(d.src[0].ptr && d.w && d.out, "essb_conv_fp32: null tensor");
  ESSB_REQUIRE(d.N > 0 && d.H > 0 && d.W > 0 && d.OH > 0 && d.OW > 0 && d.Cout > 0, "essb_conv_fp32: bad dims");
  ESSB_REQUIRE(d.ntaps > 0 && d.ntaps <= ESSB_MAX_TAPS, "essb_conv_fp32: ntaps=%d", d.ntaps);
  ESSB_REQUIRE(d.N <= 65535, "essb_conv_fp32: N too large");
  if (!d.src[1].ptr) d.src[1].C = 0;
  int rc;
  if ((rc = check_src(d.src[0], d.H, d.W, "essb_conv_fp32 src0")) != ESSB_OK) return rc;
  if ((rc = check_src(d.src[1], d.H, d.W, "essb_conv_fp32 src1")) != ESSB_OK) return rc;
  p.cin_total = d.src[0].C + d.src[1].C;
  p.coutp = (d.Cout + 3) & ~3;
  p.chunks0 = (d.src[0].C + BK - 1) / BK;
  p.chunks1 = (d.src[1].C + BK - 1) / BK;
  p.vec0 = src_vec_ok(d.src[0]);
  p.vec1 = src_vec_ok(d.src[1]);
  ESSB_REQUIRE(essb_aligned16(d.w), "essb_conv_fp32: weights must be 16B aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (d.epilogue) {
    case ESSB_EPI_LINEAR:
      ESSB_REQUIRE(d.ldo >= d.Cout, "essb_conv_fp32: ldo < Cout");
      ESSB_REQUIRE(!(d.res_pre || d.res_post) || d.ld_res >= d.Cout, "essb_conv_fp32: ld_res < Cout");
      p.vec_out = (d.ldo % 4 == 0) && essb_aligned16(d.out);
      ESSB_REQUIRE(!d.out_hi || (d.out_lo && d.Cout % 4 == 0 && d.ld_planes % 4 == 0 && !d.accumulate),
                   "essb_conv_fp32: bf16 planes need out_lo, Cout %% 4 == 0, ld_planes %% 4 == 0, no accumulate");
      return launch_conv<ESSB_EPI_LINEAR>(p, st);
    case ESSB_EPI_LSTM:
      ESSB_REQUIRE(d.Cout % 4 == 0 && d.out2, "essb_conv_fp32: LSTM needs Cout%%4==0 and out2");
      ESSB_REQUIRE(d.osy == 1 && d.osx == 1 && d.OHf == d.OH && d.OWf == d.OW, "essb_conv_fp32: LSTM output must be dense");
      return launch_conv<ESSB_EPI_LSTM>(p, st);
    case ESSB_EPI_GRU_UR:
      ESSB_REQUIRE(d.Cout % 4 == 0 && d.out2, "essb_conv_fp32: GRU_UR needs Cout%%4==0 and out2");
      return launch_conv<ESSB_EPI_GRU_UR>(p, st);
    case ESSB_EPI_GRU_OUT:
      ESSB_REQUIRE(d.aux1, "essb_conv_fp32: GRU_OUT needs the update gate in aux1");
      return launch_conv<ESSB_EPI_GRU_OUT>(p, st);
    default:
      ESSB_REQUIRE(false, "essb_conv_fp32: unknown epilogue %d", d.epilogue);
  }
  return ESSB_ERR_ARG;
}

// ================================================================================ weight gradient
namespace {

constexpr int WG_T = 64;   // ci tile = co tile
constexpr int WG_K = 16;   // pixels per K-step

struct WgradParams {
  essb_wgrad d;
  int cin_total;
  int ci_tiles, ci_tiles0, co_tiles, splits;  // ci tiles are per segment (never straddle the boundary)
  long long pix_total;      // N*OH*OW
  long long pix_per_split;  // multiple of WG_K
  int vec0, vec1, vec_dy;
};

// partial layout: [split][tap][ci][co] (co contiguous, Cout exact)
__global__ void __launch_bounds__(NTHREADS) wgrad_fp32_kernel(const __grid_constant__ WgradParams p) {
  __shared__ __align__(16) float As[2][WG_K][WG_T];
  __shared__ __align__(16) float Bs[2][WG_K][WG_T];
  const essb_wgrad& d = p.d;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  int b = blockIdx.x;
  const int cot = b % p.co_tiles; b /= p.co_tiles;
  const int cit = b % p.ci_tiles; b /= p.ci_tiles;
  const int tap = b;
  const int split = blockIdx.y;
  const int co0 = cot * WG_T;
  // ci tiles are enumerated per segment, so a tile lies entirely inside one segment
  const int seg = cit >= p.ci_tiles0 ? 1 : 0;
  const essb_src& s = d.src[seg];
  const int cseg0 = (seg ? cit - p.ci_tiles0 : cit) * WG_T;
  const int ci0 = (seg ? d.src[0].C : 0) + cseg0;
  const int vec = seg ? p.vec1 : p.vec0;
  const int dyv = d.dy[tap], dxv = d.dx[tap];
  const int npix = d.OH * d.OW;

  const long long k_begin = (long long)split * p.pix_per_split;
  long long k_end = k_begin + p.pix_per_split;
  if (k_end > p.pix_total) k_end = p.pix_total;
  const int iters = k_begin < k_end ? (int)((k_end - k_begin + WG_K - 1) / WG_K) : 0;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int lk = tid >> 4;  // pixel within K-step
  const int lf = tid & 15;  // float4 within the 64-channel row
  float4 ra, rb;
  auto gload = [&](int it) {
    const long long gp = k_begin + (long long)it * WG_K + lk;
    ra = make_float4(0.f, 0.f, 0.f, 0.f);
    rb = ra;
    if (gp < k_end) {
      const int n = (int)(gp / npix);
      const int pp = (int)(gp - (long long)n * npix);
      const int oy = pp / d.OW, ox = pp - oy * d.OW;
      ra = load_a4(s, vec, n, oy * d.sy + dyv, ox * d.sx + dxv, d.H, d.W, cseg0 + lf * 4);
      const int co = co0 + lf * 4;
      const float* q = d.dy_ptr + (size_t)gp * d.ld_dy + co;
      if (p.vec_dy && co + 3 < d.Cout) rb = ld4(q);
      else {
        float t[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int e = 0; e < 4; ++e) if (co + e < d.Cout) t[e] = q[e];
        rb = make_float4(t[0], t[1], t[2], t[3]);
      }
    }
  };
  auto sstore = [&](int buf) {
    *reinterpret_cast<float4*>(&As[buf][lk][lf * 4]) = ra;
    *reinterpret_cast<float4*>(&Bs[buf][lk][lf * 4]) = rb;
  };
  if (iters > 0) {
    gload(0);
    sstore(0);
  }
  __syncthreads();
  for (int it = 0; it < iters; ++it) {
    const int buf = it & 1;
    if (it + 1 < iters) gload(it + 1);
#pragma unroll
    for (int k = 0; k < WG_K; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float a[4] = {av.x, av.y, av.z, av.w};
      const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    if (it + 1 < iters) sstore(buf ^ 1);
    __syncthreads();
  }
  float* part = d.workspace + ((size_t)split * d.ntaps + tap) * (size_t)p.cin_total * d.Cout;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ci = ci0 + ty * 4 + i;
    if (cseg0 + ty * 4 + i >= s.C) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = co0 + tx * 4 + j;
      if (co < d.Cout) part[(size_t)ci * d.Cout + co] = acc[i][j];
    }
  }
}

// dw[co][ci][tap] = sum_split partial[split][tap][ci][co]   (fixed order => deterministic)
__global__ void wgrad_reduce_kernel(const float* __restrict__ part, float* __restrict__ dw, int splits, int ntaps,
                                    int cin, int cout) {
  const long long total = (long long)ntaps * cin * cout;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  // idx enumerates (tap, ci, co) with co fastest => coalesced reads
  const int co = (int)(idx % cout);
  const long long r = idx / cout;
  const int ci = (int)(r % cin);
  const int tap = (int)(r / cin);
  float s = 0.f;
  for (int sp = 0; sp < splits; ++sp) s += part[(size_t)sp * total + idx];
  dw[((size_t)co * cin + ci) * ntaps + tap] = s;
}

struct WgradPlan {
  int ci_tiles, ci_tiles0, co_tiles, splits;
  long long pix_per_split;
};

int wgrad_plan(const essb_wgrad& d, WgradPlan* pl) {
  const int c0 = d.src[0].C, c1 = d.src[1].ptr ? d.src[1].C : 0;
  const int t0 = (c0 + WG_T - 1) / WG_T, t1 = (c1 + WG_T - 1) / WG_T;
  pl->ci_tiles = t0 + t1;
  pl->ci_tiles0 = t0;
  pl->co_tiles = (d.Cout + WG_T - 1) / WG_T;
  const long long pix = (long long)d.N * d.OH * d.OW;
  const long long base = (long long)d.ntaps * pl->ci_tiles * pl->co_tiles;
  long long splits = (148LL * 4 + base - 1) / base;  // ~4 CTAs per SM overall
  const long long max_splits = (pix + 255) / 256;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  if (splits > 1024) splits = 1024;
  long long pps = (pix + splits - 1) / splits;
  pps = (pps + WG_K - 1) / WG_K * WG_K;
  pl->splits = (int)((pix + pps - 1) / pps);
  pl->pix_per_split = pps;
  return 0;
}

}  // namespace

extern "C" int64_t essb_wgrad_workspace_bytes(const essb_wgrad* d) {
  WgradPlan pl;
  if (wgrad_plan(*d, &pl) != 0) return -1;
  const int cin = d->src[0].C + (d->src[1].ptr ? d->src[1].C : 0);
  return (int64_t)pl.splits * d->ntaps * cin * d->Cout * (int64_t)sizeof(float);
}

extern "C" int essb_colsum(const float* x, int ld, int64_t rows, int C, float* out, float* workspace,
                           int64_t workspace_bytes, void* stream);

extern "C" int essb_wgrad_fp32(const essb_wgrad* dp, void* stream) {
  ESSB_REQUIRE(dp != nullptr, "essb_wgrad_fp32: null descriptor");
  WgradParams p;
  p.d = *dp;
  essb_wgrad& d = p.d;
  ESSB_REQUIRE(d.src[0].ptr && d.dy_ptr && d.dw && d.workspace, "essb_wgrad_fp32: null tensor");
  ESSB_REQUIRE(d.ntaps > 0 && d.ntaps <= ESSB_MAX_TAPS, "essb_wgrad_fp32: ntaps=%d", d.ntaps);
  if (!d.src[1].ptr) d.src[1].C = 0;
  int rc;
  if ((rc = check_src(d.src[0], d.H, d.W, "essb_wgrad_fp32 src0")) != ESSB_OK) return rc;
  if ((rc = check_src(d.src[1], d.H, d.W, "essb_wgrad_fp32 src1")) != ESSB_OK) return rc;
  WgradPlan pl;
  ESSB_REQUIRE(wgrad_plan(d, &pl) == 0, "essb_wgrad_fp32: bad plan");
  p.cin_total = d.src[0].C + d.src[1].C;
  p.ci_tiles = pl.ci_tiles;
  p.ci_tiles0 = pl.ci_tiles0;
  p.co_tiles = pl.co_tiles;
  p.splits = pl.splits;
  p.pix_total = (long long)d.N * d.OH * d.OW;
  p.pix_per_split = pl.pix_per_split;
  p.vec0 = src_vec_ok(d.src[0]);
  p.vec1 = src_vec_ok(d.src[1]);
  p.vec_dy = (d.ld_dy % 4 == 0) && essb_aligned16(d.dy_ptr);
  const int64_t need = (int64_t)pl.splits * d.ntaps * p.cin_total * d.Cout * (int64_t)sizeof(float);
  if (d.workspace_bytes < need) {
    essb_set_error("essb_wgrad_fp32: workspace %lld < %lld bytes", (long long)d.workspace_bytes, (long long)need);
    return ESSB_ERR_WORKSPACE;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid(d.ntaps * pl.ci_tiles * pl.co_tiles, pl.splits, 1);
  wgrad_fp32_kernel<<<grid, NTHREADS, 0, st>>>(p);
  ESSB_LAUNCH_CHECK("essb_wgrad_fp32");
  const long long total = (long long)d.ntaps * p.cin_total * d.Cout;
  wgrad_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d.workspace, d.dw, pl.splits, d.ntaps,
                                                                        p.cin_total, d.Cout);
  ESSB_LAUNCH_CHECK("essb_wgrad_reduce");
  if (d.dbias) {
    // reuse the (now consumed) split-K workspace for the column-sum partials
    return essb_colsum(d.dy_ptr, d.ld_dy, p.pix_total, d.Cout, d.dbias, d.workspace, d.workspace_bytes, stream);
  }
  return ESSB_OK;
}
