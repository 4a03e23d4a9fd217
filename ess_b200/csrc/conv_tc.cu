// tcgen05 / TMA implicit-GEMM convolution for sm_100a ("bf16x3" fp32-parity mode and "bf16" mode).
//
// GEMM view of a gather-convolution on NHWC activations:
//   M = 128 output pixels (a BH x BW spatial patch of one sample), N = BN <= 256 output channels,
//   K = taps x input channels, consumed 64 channels of one tap per pipeline stage.
// A operand: the activation is kept in HBM as two bf16 planes (hi = bf16(x), lo = bf16(x - hi)); for
//   tap (dy,dx) the A tile is the SAME 4-D TMA box shifted by (dy,dx) -- TMA zero-fills outside the
//   image, which is exactly the convolution's zero padding, so no im2col buffer ever exists.
//   Stride-2 convolutions address the input through parity-plane views (one tensor map per parity).
// B operand: packed K-major weights [Cout][taps*Cin] as bf16 hi/lo planes, one 2-D TMA box per stage.
// MMA: tcgen05.mma.cta_group::1.kind::f16, M=128 x N=BN x K=16, fp32 accumulators in TMEM.  In the
//   3-pass mode each K-step issues hi*hi + lo*hi + hi*lo (drops only lo*lo ~ 2^-16 relative), which
//   restores fp32-level accuracy (logits 1.1e-4 vs 4.7e-2 for single-pass bf16).
//   Narrow tiles (2*BN <= 256) fuse the two B planes: [B_hi; B_lo] are adjacent K-major tiles, so A_hi x [B_hi; B_lo]
//   is ONE MMA of N = 2*BN and the epilogue adds the upper BN accumulator columns back (2 MMAs / 2 A reads per K-step).
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer,
//   warps 2..9 = epilogue (TMEM lane quarter = warp % 4, two warps per quarter split the columns).  Two TMEM accumulator buffers let the
//   epilogue of tile i overlap the main loop of tile i+1.
// Scheduling: persistent CTAs (one per SM; two per SM for the halo variant when its pipeline fits half an SM) pull
//   tiles from a global atomic counter; the producer thread publishes each tile to the other warps through an
//   mbarrier-guarded ring in shared memory (TcUnit / tc_sched_*).  Optional split-K of the tail wave (off by default).
// Epilogues: LINEAR (bias / residual / ReLU / sigmoid / strided placement / bf16 planes out) and
//   LSTM (sigma/tanh gates + cell/hidden update in registers: the 4C-channel `gates` tensor of
//   e2vid/model/submodules.py:213 never reaches HBM).
#include <cuda.h>
#include <cuda_bf16.h>

#include <stdlib.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int TC_THREADS = 320;               // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int TC_M = 128;
constexpr int TC_KCH = 64;                      // channels per stage (128 B of bf16 = one swizzle row)
constexpr int A_TILE_BYTES = TC_M * TC_KCH * 2; // 16 KB
constexpr int MAX_VIEWS = 8;
constexpr int MAX_STAGES = 8;

struct TcParams {
  CUtensorMap tmA_hi[MAX_VIEWS];
  CUtensorMap tmA_lo[MAX_VIEWS];
  CUtensorMap tmB_hi;
  CUtensorMap tmB_lo;
  // schedule
  int n_items, n_tiles, tiles_x, tiles_y;
  int bw_log2;             // BW = 1 << bw_log2, BH = 128 >> bw_log2
  int BN;                  // N tile (multiple of 16, <= 256)
  int stages, passes;
  int stage_bytes, tx_bytes;  // bytes per pipeline stage / per-stage TMA transaction bytes
  int off_alo, off_bhi, off_blo;  // tile offsets inside a stage
  // K loop
  int ntaps, nseg;
  int seg_chunks[2];       // 64-channel chunks per segment
  int seg_view0[2];        // first tensor-map index of the segment
  int seg_koff[2];         // channel offset of the segment inside one tap's K range
  int k_per_tap;           // total (padded) channels per tap in the packed weights
  int8_t dy[ESSB_MAX_TAPS], dx[ESSB_MAX_TAPS], view[ESSB_MAX_TAPS], widx[ESSB_MAX_TAPS];
  // epilogue
  int N, OH, OW, Cout;
  int OHf, OWf, osy, ooy, osx, oox;
  int ldo, ld_res, ld_planes;
  int act;
  const float* bias;
  const float* res_pre;
  const float* res_post;
  const float* aux0;
  const float* aux1;
  float* out;
  float* out2;
  __nv_bfloat16* out_hi;
  __nv_bfloat16* out_lo;
  // halo-reuse mode (conv_tc_halo_kernel): one A tile (output patch + halo) per (segment, chunk, view),
  // shared by every tap that reads that view; separate A and B pipelines
  int halo_w, halo_h, hx0, hy0;          // halo box extents (pixels) and the smallest tap offsets
  int a_stages, b_stages, a_stage_bytes, b_stage_bytes, a_lo_off, b_lo_off;
  int b_taps_per_stage, b_tap_bytes;     // a B stage holds up to G consecutive taps of one group behind ONE barrier
  int base_offset_mode;                  // EXPERIMENT flags (env ESSB_TC_DEBUG, default 0; tools/halo_probe.py): 1 = epilogue skips math + stores,
                                         // 2 = halo kernel issues no MMAs, 4 = epilogue also skips the TMEM loads.  Results are garbage when set.
  int wide;                              // TC_WIDE_* bits (256-bit epilogue accesses)
  int n_groups;                          // tap groups = views actually used; taps are sorted by group
  int8_t grp_view[MAX_VIEWS], grp_first[MAX_VIEWS], grp_count[MAX_VIEWS];
  // dynamic tile scheduler + split-K tail (see TcUnit)
  int tmem_cols, tmem_buf_stride;        // TMEM columns allocated per CTA (power of two) / column offset of accumulator buffer 1
  int fuse_b;                            // bf16x3 with 2*BN <= 256: A_hi x [B_hi; B_lo] is ONE MMA of N = 2*BN (see below)
  int pair;                              // conv_tc_pair_kernel: CTA pairs (cta_group::2), B maps carry BN/2-row boxes
  int b_split;                           // wide halo tiles (BN = 256): a B stage is ONE plane of one tap (hi and lo planes behind separate barriers)
  float acc_scale;                       // the epilogue multiplies the raw accumulator by this (f16f8 mode: 2^-(w8+14))
  int planes_fmt;                        // format of out_hi/out_lo: 0 = bf16 hi/lo, 2 = hf8 (tc_ptx.cuh)
  int phase_cout;                        // > 0: merged sub-pixel phases (essb_conv_tc.phase_cout): column block j of width phase_cout is phase (j >> 1, j & 1)
  int row_period, rows_valid;            // row-stacked batch (essb_conv_tc): rows with oy % row_period >= rows_valid are not stored
  int* sched;                            // [0] next unit, [1] CTAs done; zero before the launch, reset by the last CTA
  int n_units, n_whole, split;           // units [0, n_whole) are whole tiles; the rest are 1/split K-slices of the tail tiles
  float* splitk_ws;                      // [tail tile][part][32-col chunk][128 rows][32] fp32 partial accumulators
  int* splitk_cnt;                       // [tail tile][8 epilogue warps] arrival counters (zero; the last arriver resets)
};

// Work unit handed out by the scheduler.  The persistent CTAs pull units from a global counter (warp 0 fetches,
// an mbarrier-guarded ring in shared memory broadcasts the unit to the MMA and epilogue warps), so a CTA that
// runs slow simply takes fewer tiles.  The tiles of the last, partially filled wave are cut into `split`
// K-slices: each slice leaves its fp32 partial accumulator in splitk_ws, and the epilogue warp that arrives
// last on the tile's counter sums the slices in slice order (deterministic) and runs the fused epilogue.
struct TcUnit {
  int item, k0, k1, part, slot;          // slot < 0: whole tile
};
__device__ __forceinline__ TcUnit tc_decode_unit(const TcParams& p, int u, int k_iters) {
  TcUnit t;
  if (u < p.n_whole) {
    t.item = u; t.k0 = 0; t.k1 = k_iters; t.part = 0; t.slot = -1;
  } else {
    const int v = u - p.n_whole;
    t.slot = v / p.split;
    t.part = v - t.slot * p.split;
    t.item = p.n_whole + t.slot;
    t.k0 = (t.part * k_iters) / p.split;
    t.k1 = ((t.part + 1) * k_iters) / p.split;
  }
  return t;
}
constexpr int SCHED_DEPTH = 4;
// producer side: fetch the next unit (or -1) and publish it in ring slot local % SCHED_DEPTH
__device__ __forceinline__ int tc_sched_fetch(const TcParams& p, int local, int* ring, uint64_t* s_full, uint64_t* s_empty) {
  const int slot = local % SCHED_DEPTH;
  mbar_wait(&s_empty[slot], ((uint32_t)(local / SCHED_DEPTH) & 1u) ^ 1u);
  int u = atomicAdd(p.sched, 1);
  if (u >= p.n_units) u = -1;
  ring[slot] = u;
  mbar_arrive(&s_full[slot]);
  return u;
}
// consumer side (whole warp): read the unit published for iteration `local`
__device__ __forceinline__ int tc_sched_read(int local, int lane, const int* ring, uint64_t* s_full, uint64_t* s_empty) {
  const int slot = local % SCHED_DEPTH;
  mbar_wait(&s_full[slot], (uint32_t)(local / SCHED_DEPTH) & 1u);
  const int u = *reinterpret_cast<const volatile int*>(&ring[slot]);
  __syncwarp();
  if (lane == 0) mbar_arrive(&s_empty[slot]);
  return u;
}
// last CTA out resets the scheduler counters for the next launch that uses this slot
__device__ __forceinline__ void tc_sched_finish(const TcParams& p) {
  const int old = atomicAdd(p.sched + 1, 1);
  if (old == (int)gridDim.x - 1) {
    p.sched[0] = 0;
    p.sched[1] = 0;
  }
}

// TcParams::wide bits: which epilogue streams are 32 B-aligned and may use 256-bit accesses
constexpr int TC_WIDE_OUT = 1, TC_WIDE_PLANES = 2, TC_WIDE_RES = 4, TC_WIDE_AUX = 8;
constexpr int TC_BIAS_SMEM_FLOATS = 1024;  // the bias vector is staged in shared memory once per CTA

// v[0..15] += 16 consecutive floats of a residual stream
__device__ __forceinline__ void tc_add_res(const TcParams& p, const float* src, float (&v)[2][8]) {
  if (p.wide & TC_WIDE_RES) {
    float a[8];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      ld_global_256(src + h * 8, a);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[h][e] += a[e];
    }
  } else {
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const float4 a = *reinterpret_cast<const float4*>(src + h * 4);
      float* vv = &v[h >> 1][(h & 1) * 4];
      vv[0] += a.x; vv[1] += a.y; vv[2] += a.z; vv[3] += a.w;
    }
  }
}

// 8 consecutive channels (co a multiple of 8) of output pixel `opix` -> the operand planes of the next tcgen05 launch
__device__ __forceinline__ void tc_store_planes8(const TcParams& p, size_t opix, int co, const float (&v)[8]) {
  if (p.planes_fmt == 2) {
    uint32_t h[4];
    uint16_t a[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) split_hf8x2(v[2 * e], v[2 * e + 1], h[e], a[e], l[e]);
    *reinterpret_cast<uint4*>(p.out_hi + opix * p.ld_planes + co) = make_uint4(h[0], h[1], h[2], h[3]);
    uint8_t* lo = reinterpret_cast<uint8_t*>(p.out_lo) + opix * (size_t)p.ld_planes * 2 + hf8_lo_off(co);
    *reinterpret_cast<uint2*>(lo) = make_uint2((uint32_t)a[0] | ((uint32_t)a[1] << 16), (uint32_t)a[2] | ((uint32_t)a[3] << 16));
    *reinterpret_cast<uint2*>(lo + 64) = make_uint2((uint32_t)l[0] | ((uint32_t)l[1] << 16), (uint32_t)l[2] | ((uint32_t)l[3] << 16));
  } else {
    bf16x8 hh, hl;
#pragma unroll
    for (int e = 0; e < 8; ++e) split_bf16(v[e], hh.v[e], hl.v[e]);
    *reinterpret_cast<bf16x8*>(p.out_hi + opix * p.ld_planes + co) = hh;
    *reinterpret_cast<bf16x8*>(p.out_lo + opix * p.ld_planes + co) = hl;
  }
}

// Fused epilogue of one 32-column chunk of an accumulator row (r = the fp32 accumulators of row `pix`, columns
// n0 + c0 .. + 31): bias / gates / activation / residuals -> global stores.  cpj = previous cell state (LSTM).
template <int EPI>
__device__ __forceinline__ void tc_epilogue_chunk(const TcParams& p, const uint32_t (&r)[32], int n0, int c0, size_t pix,
                                                  int n, int oy, int ox, const float (&cpj)[8],
                                                  const float* __restrict__ s_bias) {
  if constexpr (EPI == ESSB_EPI_LSTM) {
    // columns co = 4*ch + {in, remember, out, cell}; 32 columns = 8 channels
    const int hidden = p.Cout >> 2;
    const int ch0 = (n0 + c0) >> 2;
    float hv[8], cv[8];
    const float sc = p.acc_scale;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int co = n0 + c0 + e * 4;
      float gi = __uint_as_float(r[e * 4 + 0]) * sc, gf = __uint_as_float(r[e * 4 + 1]) * sc;
      float go = __uint_as_float(r[e * 4 + 2]) * sc, gc = __uint_as_float(r[e * 4 + 3]) * sc;
      if (s_bias) {
        const float4 b4 = *reinterpret_cast<const float4*>(s_bias + co);
        gi += b4.x; gf += b4.y; go += b4.z; gc += b4.w;
      }
      const float cell = sigmoid_fast(gf) * cpj[e] + sigmoid_fast(gi) * tanh_fast(gc);
      cv[e] = cell;
      hv[e] = sigmoid_fast(go) * tanh_fast(cell);
    }
    float* ho = p.out + pix * hidden + ch0;
    float* co_ = p.out2 + pix * hidden + ch0;
    if (p.wide & TC_WIDE_OUT) {
      st_global_256(ho, hv);
      st_global_256(co_, cv);
    } else {
      *reinterpret_cast<float4*>(ho) = make_float4(hv[0], hv[1], hv[2], hv[3]);
      *reinterpret_cast<float4*>(ho + 4) = make_float4(hv[4], hv[5], hv[6], hv[7]);
      *reinterpret_cast<float4*>(co_) = make_float4(cv[0], cv[1], cv[2], cv[3]);
      *reinterpret_cast<float4*>(co_ + 4) = make_float4(cv[4], cv[5], cv[6], cv[7]);
    }
    if (p.out_hi) tc_store_planes8(p, pix, ch0, hv);
  } else if constexpr (EPI == ESSB_EPI_GRU_UR) {
    // ConvGRU update / reset gates (submodules.py:267-268): columns co = 2*ch + {update, reset};
    // 32 columns = 16 channels.  Writes update (fp32) and prev_state*reset as bf16 planes (the A
    // operand of the out-gate convolution).
    const int hidden = p.Cout >> 1;
    const int ch0 = (n0 + c0) >> 1;
#pragma unroll
    const float sc = p.acc_scale;
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      float uv[8], hr[8];
      float hp[8];
      if (p.aux0) {
        const float4 a0 = *reinterpret_cast<const float4*>(p.aux0 + pix * hidden + ch0 + g * 8);
        const float4 a1 = *reinterpret_cast<const float4*>(p.aux0 + pix * hidden + ch0 + g * 8 + 4);
        hp[0] = a0.x; hp[1] = a0.y; hp[2] = a0.z; hp[3] = a0.w; hp[4] = a1.x; hp[5] = a1.y; hp[6] = a1.z; hp[7] = a1.w;
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) hp[e] = 0.f;
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int col = g * 16 + e * 2;
        float gu = __uint_as_float(r[col]) * sc, gr = __uint_as_float(r[col + 1]) * sc;
        if (s_bias) { gu += s_bias[n0 + c0 + col]; gr += s_bias[n0 + c0 + col + 1]; }
        uv[e] = sigmoid_fast(gu);
        hr[e] = hp[e] * sigmoid_fast(gr);
      }
      float* uo = p.out + pix * hidden + ch0 + g * 8;
      *reinterpret_cast<float4*>(uo) = make_float4(uv[0], uv[1], uv[2], uv[3]);
      *reinterpret_cast<float4*>(uo + 4) = make_float4(uv[4], uv[5], uv[6], uv[7]);
      tc_store_planes8(p, pix, ch0 + g * 8, hr);
    }
  } else if constexpr (EPI == ESSB_EPI_GRU_OUT) {
    // ConvGRU out gate + blend (submodules.py:269-271): h' = h*(1-u) + tanh(acc + b)*u
    const int hidden = p.Cout;
    const int ch0 = n0 + c0;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int ch = ch0 + g * 8;
      float hv[8], hp[8], uu[8];
      const float4 u0 = *reinterpret_cast<const float4*>(p.aux1 + pix * hidden + ch);
      const float4 u1 = *reinterpret_cast<const float4*>(p.aux1 + pix * hidden + ch + 4);
      uu[0] = u0.x; uu[1] = u0.y; uu[2] = u0.z; uu[3] = u0.w; uu[4] = u1.x; uu[5] = u1.y; uu[6] = u1.z; uu[7] = u1.w;
      if (p.aux0) {
        const float4 a0 = *reinterpret_cast<const float4*>(p.aux0 + pix * hidden + ch);
        const float4 a1 = *reinterpret_cast<const float4*>(p.aux0 + pix * hidden + ch + 4);
        hp[0] = a0.x; hp[1] = a0.y; hp[2] = a0.z; hp[3] = a0.w; hp[4] = a1.x; hp[5] = a1.y; hp[6] = a1.z; hp[7] = a1.w;
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) hp[e] = 0.f;
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float x = __uint_as_float(r[g * 8 + e]) * p.acc_scale;
        if (s_bias) x += s_bias[ch + e];
        hv[e] = hp[e] * (1.f - uu[e]) + tanh_fast(x) * uu[e];
      }
      float* ho = p.out + pix * hidden + ch;
      *reinterpret_cast<float4*>(ho) = make_float4(hv[0], hv[1], hv[2], hv[3]);
      *reinterpret_cast<float4*>(ho + 4) = make_float4(hv[4], hv[5], hv[6], hv[7]);
      if (p.out_hi) tc_store_planes8(p, pix, ch, hv);
    }
  } else {
    // merged sub-pixel phases: this 32-column chunk belongs to phase ph = column / phase_cout, whose outputs live at
    // pixel (oy * osy + (ph >> 1), ox * osx + (ph & 1)) and channel column % phase_cout (phase_cout is a multiple of 32)
    int ccol = n0 + c0, ppy = p.ooy, ppx = p.oox;
    if (p.phase_cout) {
      const int ph = ccol / p.phase_cout;
      ccol -= ph * p.phase_cout;
      ppy = ph >> 1;
      ppx = ph & 1;
    }
    const size_t opix = ((size_t)n * p.OHf + (oy * p.osy + ppy)) * p.OWf + (ox * p.osx + ppx);
    const bool do_st = !(p.base_offset_mode & 8);   // experiment flag: compute but do not store
    uint32_t q8[8], q8l[8];   // hf8 planes: the chunk's 32 e4m3 a8 / a8l bytes, stored as ONE 32 B sector each after the loop
#pragma unroll
    for (int g = 0; g < 2; ++g) {  // 16 channels per group: 2 x 32 B of fp32, 32 B per bf16 plane
      const int co = ccol + g * 16;          // output channel (bias is indexed by the GEMM column: cb)
      const int cb = n0 + c0 + g * 16;
      float v[2][8];
#pragma unroll
      for (int e = 0; e < 16; ++e) v[e >> 3][e & 7] = __uint_as_float(r[g * 16 + e]) * p.acc_scale;
      if (s_bias) {
#pragma unroll
        for (int e4 = 0; e4 < 4; ++e4) {
          const float4 b4 = *reinterpret_cast<const float4*>(s_bias + cb + e4 * 4);
          float* vv = &v[e4 >> 1][(e4 & 1) * 4];
          vv[0] += b4.x; vv[1] += b4.y; vv[2] += b4.z; vv[3] += b4.w;
        }
      }
      if (p.res_pre) tc_add_res(p, p.res_pre + opix * p.ld_res + co, v);
      if (p.act == ESSB_ACT_RELU) {
#pragma unroll
        for (int e = 0; e < 16; ++e) v[e >> 3][e & 7] = fmaxf(v[e >> 3][e & 7], 0.f);
      } else if (p.act == ESSB_ACT_SIGMOID) {
#pragma unroll
        for (int e = 0; e < 16; ++e) v[e >> 3][e & 7] = essb_sigmoid(v[e >> 3][e & 7]);
      }
      if (p.res_post) tc_add_res(p, p.res_post + opix * p.ld_res + co, v);
      if (p.out && do_st) {
        float* o = p.out + opix * p.ldo + co;
        if (p.wide & TC_WIDE_OUT) {
          st_global_256(o, v[0]);
          st_global_256(o + 8, v[1]);
        } else {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            *reinterpret_cast<float4*>(o + h * 8) = make_float4(v[h][0], v[h][1], v[h][2], v[h][3]);
            *reinterpret_cast<float4*>(o + h * 8 + 4) = make_float4(v[h][4], v[h][5], v[h][6], v[h][7]);
          }
        }
      }
      if (p.out_hi && p.planes_fmt == 2) {
        uint32_t h[8];
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
          uint16_t a0, l0, a1, l1;
          split_hf8x2(v[e >> 2][(e & 3) * 2], v[e >> 2][(e & 3) * 2 + 1], h[e], a0, l0);
          split_hf8x2(v[(e + 1) >> 2][((e + 1) & 3) * 2], v[(e + 1) >> 2][((e + 1) & 3) * 2 + 1], h[e + 1], a1, l1);
          q8[g * 4 + (e >> 1)] = (uint32_t)a0 | ((uint32_t)a1 << 16);
          q8l[g * 4 + (e >> 1)] = (uint32_t)l0 | ((uint32_t)l1 << 16);
        }
        __nv_bfloat16* oh = p.out_hi + opix * p.ld_planes + co;
        if (!do_st) {
          if (h[0] == 0x12345678u && q8[0] == 0x9abcdef0u) p.out_hi[0] = __float2bfloat16(1.f);   // keep the math alive
        } else if (p.wide & TC_WIDE_PLANES) {
          st_global_256(oh, h);
        } else {
          *reinterpret_cast<uint4*>(oh) = make_uint4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<uint4*>(oh + 8) = make_uint4(h[4], h[5], h[6], h[7]);
        }
      } else if (p.out_hi) {
        uint32_t ph[8], pl[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          __nv_bfloat16 h0, l0, h1, l1;
          split_bf16(v[e >> 2][(e & 3) * 2], h0, l0);
          split_bf16(v[e >> 2][(e & 3) * 2 + 1], h1, l1);
          ph[e] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
          pl[e] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
        }
        __nv_bfloat16* oh = p.out_hi + opix * p.ld_planes + co;
        __nv_bfloat16* ol = p.out_lo + opix * p.ld_planes + co;
        if (!do_st) {
          if (ph[0] == 0x12345678u && pl[3] == 0x9abcdef0u) p.out_hi[0] = __float2bfloat16(1.f);   // keep the math alive
        } else if (p.wide & TC_WIDE_PLANES) {
          st_global_256(oh, ph);
          st_global_256(ol, pl);
        } else {
          *reinterpret_cast<uint4*>(oh) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
          *reinterpret_cast<uint4*>(oh + 8) = make_uint4(ph[4], ph[5], ph[6], ph[7]);
          *reinterpret_cast<uint4*>(ol) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
          *reinterpret_cast<uint4*>(ol + 8) = make_uint4(pl[4], pl[5], pl[6], pl[7]);
        }
      }
    }
    if (p.out_hi && p.planes_fmt == 2 && do_st) {   // n0 + c0 is a multiple of 32: both 32 B pieces lie inside one 64-channel chunk row
      uint8_t* lo = reinterpret_cast<uint8_t*>(p.out_lo) + opix * (size_t)p.ld_planes * 2 + hf8_lo_off(ccol);
      if (p.wide & TC_WIDE_PLANES) {
        st_global_256(lo, q8);
        st_global_256(lo + 64, q8l);
      } else {
        *reinterpret_cast<uint4*>(lo) = make_uint4(q8[0], q8[1], q8[2], q8[3]);
        *reinterpret_cast<uint4*>(lo + 16) = make_uint4(q8[4], q8[5], q8[6], q8[7]);
        *reinterpret_cast<uint4*>(lo + 64) = make_uint4(q8l[0], q8l[1], q8l[2], q8l[3]);
        *reinterpret_cast<uint4*>(lo + 80) = make_uint4(q8l[4], q8l[5], q8l[6], q8l[7]);
      }
    }
  }
}

// 256-bit L2-only load (split-K partials written by another SM: must not be served from this SM's L1)
__device__ __forceinline__ void ld_global_cg_256(const float* ptr, float (&v)[8]) {
  asm volatile("ld.global.cg.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(ptr)
               : "memory");
}

// LSTM: previous cell state of this thread's row for column chunk c0 (8 channels)
__device__ __forceinline__ void tc_load_cprev(const TcParams& p, bool valid, size_t pix, int n0, int c0, float (&cpj)[8]) {
  const int hidden = p.Cout >> 2;
#pragma unroll
  for (int e = 0; e < 8; ++e) cpj[e] = 0.f;
  if (valid && p.aux0 && c0 < p.BN) {
    const float* src = p.aux0 + pix * hidden + ((n0 + c0) >> 2);
    if (p.wide & TC_WIDE_AUX) {
      ld_global_256(src, cpj);
    } else {
      const float4 c0v = *reinterpret_cast<const float4*>(src);
      const float4 c1v = *reinterpret_cast<const float4*>(src + 4);
      cpj[0] = c0v.x; cpj[1] = c0v.y; cpj[2] = c0v.z; cpj[3] = c0v.w;
      cpj[4] = c1v.x; cpj[5] = c1v.y; cpj[6] = c1v.z; cpj[7] = c1v.w;
    }
  }
}

// One work unit: wait for the MMAs, TMEM -> registers -> fused epilogue -> global stores, then hand the TMEM
// buffer back.  Shared by the classic and the halo-reuse kernels.  q = TMEM lane quarter (warp % 4), half =
// which of the two warps of that quarter (interleaved 32-column chunks).  A split-K unit (un.slot >= 0) parks
// its partial accumulator in splitk_ws instead; the warp that arrives last on the tile's counter sums all
// slices in slice order and runs the epilogue.
template <int EPI>
__device__ __forceinline__ void tc_epilogue_item(const TcParams& p, uint32_t tmem_base, uint64_t* tfull_bar,
                                                 uint64_t* tempty_bar, uint32_t (&tph)[2], int local, const TcUnit& un,
                                                 int q, int half, int lane, const float* __restrict__ s_bias,
                                                 const uint32_t* tempty_cluster = nullptr) {
  const int BW = 1 << p.bw_log2, BH = TC_M >> p.bw_log2;
  const int buf = local & 1;
  const int item = un.item;
  const int nt = item % p.n_tiles;
  int mt = item / p.n_tiles;
  const int txi = mt % p.tiles_x; mt /= p.tiles_x;
  const int tyi = mt % p.tiles_y;
  const int n = mt / p.tiles_y;
  const int m = q * 32 + lane;
  const int oy = tyi * BH + (m >> p.bw_log2), ox = txi * BW + (m & (BW - 1));
  const bool valid = oy < p.OH && ox < p.OW && (p.row_period == 0 || (oy % p.row_period) < p.rows_valid);
  const int n0 = nt * p.BN;
  const size_t pix = ((size_t)n * p.OH + oy) * p.OW + ox;
  const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * p.tmem_buf_stride);
  float cp[4][8];
  if (un.slot < 0) {
    // LSTM: the previous cell state does not depend on the MMA -> fetch it BEFORE waiting for the
    // accumulator so its latency hides under the main loop of this tile
    if constexpr (EPI == ESSB_EPI_LSTM) {
#pragma unroll
      for (int j = 0; j < 4; ++j) tc_load_cprev(p, valid, pix, n0, half * 32 + j * 64, cp[j]);
    }
    mbar_wait(&tfull_bar[buf], tph[buf]);
    tph[buf] ^= 1;
    tc_fence_after();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c0 = half * 32 + j * 64;
      if (c0 >= p.BN) break;  // warp-uniform
      uint32_t r[32];
      __syncwarp();
      if (p.base_offset_mode & 4) continue;
      tmem_ld32(t_addr + (uint32_t)c0, r);
      if (p.fuse_b) {  // columns [BN, 2*BN) hold the A_hi x B_lo products of the same outputs
        uint32_t r2[32];
        tmem_ld32(t_addr + (uint32_t)(p.BN + c0), r2);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) + __uint_as_float(r2[e]));
      } else {
        tmem_ld_wait();
      }
      if (valid && !(p.base_offset_mode & 1)) tc_epilogue_chunk<EPI>(p, r, n0, c0, pix, n, oy, ox, cp[j], s_bias);
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      if (tempty_cluster) mbar_arrive_cluster(tempty_cluster[buf]);   // CTA pair: the accumulator barrier lives in the leader CTA
      else mbar_arrive(&tempty_bar[buf]);
    }
    return;
  }
  // ---- split-K slice
  mbar_wait(&tfull_bar[buf], tph[buf]);
  tph[buf] ^= 1;
  tc_fence_after();
  const int nchunks = p.BN >> 5;
  float* ws_tile = p.splitk_ws + (size_t)un.slot * p.split * nchunks * (TC_M * 32);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c0 = half * 32 + j * 64;
    if (c0 >= p.BN) break;
    uint32_t r[32];
    __syncwarp();
    tmem_ld32(t_addr + (uint32_t)c0, r);
    tmem_ld_wait();
    float* dst = ws_tile + ((size_t)(un.part * nchunks + (c0 >> 5)) * TC_M + m) * 32;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const uint32_t v8[8] = {r[e * 8], r[e * 8 + 1], r[e * 8 + 2], r[e * 8 + 3],
                              r[e * 8 + 4], r[e * 8 + 5], r[e * 8 + 6], r[e * 8 + 7]};
      st_global_256(dst + e * 8, v8);
    }
  }
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(&tempty_bar[buf]);
  if (half * 32 >= p.BN) return;                       // this warp owns no columns of a narrow tile
  __threadfence();                                     // partials visible device-wide before the arrival
  __syncwarp();
  int last = 0;
  if (lane == 0) {
    int* cnt = p.splitk_cnt + un.slot * 8 + half * 4 + q;
    last = atomicAdd(cnt, 1) == p.split - 1;
    if (last) *cnt = 0;                                // ready for the next launch
  }
  last = __shfl_sync(0xffffffffu, last, 0);
  if (!last) return;
  __threadfence();
  if constexpr (EPI == ESSB_EPI_LSTM) {
#pragma unroll
    for (int j = 0; j < 4; ++j) tc_load_cprev(p, valid, pix, n0, half * 32 + j * 64, cp[j]);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c0 = half * 32 + j * 64;
    if (c0 >= p.BN) break;
    float acc[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) acc[e] = 0.f;
    for (int part = 0; part < p.split; ++part) {
      const float* src = ws_tile + ((size_t)(part * nchunks + (c0 >> 5)) * TC_M + m) * 32;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float v[8];
        ld_global_cg_256(src + e * 8, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[e * 8 + i] += v[i];
      }
    }
    uint32_t r[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) r[e] = __float_as_uint(acc[e]);
    if (valid) tc_epilogue_chunk<EPI>(p, r, n0, c0, pix, n, oy, ox, cp[j], s_bias);
  }
}

// ------------------------------------------------------------------------------------ the kernel
template <int EPI>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const __grid_constant__ TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* stage_base = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * p.stage_bytes);
  uint64_t* full_bar = bars;                    // [stages]
  uint64_t* empty_bar = bars + MAX_STAGES;      // [stages]
  uint64_t* tfull_bar = bars + 2 * MAX_STAGES;  // [2]
  uint64_t* tempty_bar = tfull_bar + 2;         // [2]
  uint64_t* sfull_bar = tempty_bar + 2;         // [SCHED_DEPTH] scheduler ring: unit published
  uint64_t* sempty_bar = sfull_bar + SCHED_DEPTH;  // [SCHED_DEPTH] unit read by the MMA warp + 8 epilogue warps
  int* sched_ring = reinterpret_cast<int*>(sempty_bar + SCHED_DEPTH);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sched_ring + SCHED_DEPTH);

  // the shuffle tells ptxas the warp index is warp-uniform, so the role branches below are uniform branches
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int cpt = p.seg_chunks[0] + (p.nseg > 1 ? p.seg_chunks[1] : 0);
  const int k_iters = p.ntaps * cpt;
  const int BW = 1 << p.bw_log2, BH = TC_M >> p.bw_log2;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], 8);  // one arrive per epilogue warp
    }
    for (int i = 0; i < SCHED_DEPTH; ++i) {
      mbar_init(&sfull_bar[i], 1);
      mbar_init(&sempty_bar[i], 9);
    }
    fence_barrier_init();
    tma_prefetch_desc(&p.tmB_hi);
    if (p.passes != 1) tma_prefetch_desc(&p.tmB_lo);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // bias vector -> shared memory (read by every epilogue warp for every tile)
  float* s_bias_buf = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512);
  if (p.bias)
    for (int i = threadIdx.x; i < p.Cout; i += blockDim.x) s_bias_buf[i] = p.bias[i];
  const float* s_bias = p.bias ? s_bias_buf : nullptr;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ============================================================ TMA producer (one lane)
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      const uint32_t tx_bytes = (uint32_t)p.tx_bytes;
      for (int local = 0;; ++local) {
        const int u = tc_sched_fetch(p, local, sched_ring, sfull_bar, sempty_bar);
        if (u < 0) break;
        const TcUnit un = tc_decode_unit(p, u, k_iters);
        const int nt = un.item % p.n_tiles;
        int mt = un.item / p.n_tiles;
        const int txi = mt % p.tiles_x; mt /= p.tiles_x;
        const int tyi = mt % p.tiles_y;
        const int n = mt / p.tiles_y;
        const int x0 = txi * BW, y0 = tyi * BH;
        for (int it = un.k0; it < un.k1; ++it) {
          const int tap = it / cpt;
          const int r = it - tap * cpt;
          const int seg = (r >= p.seg_chunks[0]) ? 1 : 0;
          const int ch = (seg ? r - p.seg_chunks[0] : r) * TC_KCH;
          const int v = p.seg_view0[seg] + p.view[tap];
          const int kcoord = p.widx[tap] * p.k_per_tap + p.seg_koff[seg] + ch;
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* st = stage_base + (size_t)s * p.stage_bytes;
          mbar_expect_tx(&full_bar[s], tx_bytes);
          const int cx = x0 + p.dx[tap], cy = y0 + p.dy[tap];
          tma_load_4d(st, &p.tmA_hi[v], &full_bar[s], ch, cx, cy, n);
          tma_load_2d(st + p.off_bhi, &p.tmB_hi, &full_bar[s], kcoord, nt * p.BN);
          if (p.passes != 1) {
            tma_load_4d(st + p.off_alo, &p.tmA_lo[v], &full_bar[s], ch, cx, cy, n);
            tma_load_2d(st + p.off_blo, &p.tmB_lo, &full_bar[s], kcoord, nt * p.BN);
          }
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
      }
      tc_sched_finish(p);
    }
  } else if (warp == 1) {
    // ============================================================ MMA issuer (whole warp, convergent:
    // umma_bf16 / umma_commit elect the issuing lane themselves, everything else stays in uniform registers)
    {
      // kind::f16 operand format: bf16 (1) in the bf16 / bf16x3 modes, fp16 (0) in the f16f8 mode
      const uint32_t idesc = p.passes == 2 ? make_idesc_fmt(TC_M, p.BN, 0u, 0u) : make_idesc(TC_M, p.BN);
      const uint32_t idesc_f8 = make_idesc_fmt(TC_M, p.BN, 0u, 0u);   // kind::f8f6f4: e4m3 x e4m3
      int s = 0;
      uint32_t ph = 0;
      uint32_t tph[2] = {0, 0};
      for (int local = 0;; ++local) {
        const int u = tc_sched_read(local, lane, sched_ring, sfull_bar, sempty_bar);
        if (u < 0) break;
        const TcUnit un = tc_decode_unit(p, u, k_iters);
        const int buf = local & 1;
        mbar_wait(&tempty_bar[buf], tph[buf] ^ 1);  // epilogue has drained this accumulator
        tph[buf] ^= 1;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * 256);
        for (int it = un.k0; it < un.k1; ++it) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(stage_base + (size_t)s * p.stage_bytes);
          if (p.passes == 2) {
            // f16f8: fp16 main product + ONE K = 128 e4m3 product over the [a8 | a8l] x [w8l | w8] pair rows (both cross
            // terms): 8 MMAs per 64-channel stage instead of 12, same operand bytes
            const UDesc a_hi = make_smem_desc(sa), a_lo = make_smem_desc(sa + p.off_alo);
            const UDesc b_hi = make_smem_desc(sa + p.off_bhi);
            const UDesc b_lo = make_smem_desc(sa + p.off_blo);
#pragma unroll
            for (int k = 0; k < TC_KCH / 16; ++k) {
              const uint32_t ko = (uint32_t)(k * 2);
              umma_f8(d_tmem, a_lo + ko, b_lo + ko, idesc_f8, ((it - un.k0) | k) != 0);
              umma_bf16(d_tmem, a_hi + ko, b_hi + ko, idesc, 1u);
            }
          } else if (p.passes == 3) {
            const UDesc a_hi = make_smem_desc(sa), a_lo = make_smem_desc(sa + p.off_alo);
            const UDesc b_hi = make_smem_desc(sa + p.off_bhi);
            const UDesc b_lo = make_smem_desc(sa + p.off_blo);
#pragma unroll
            for (int k = 0; k < TC_KCH / 16; ++k) {
              const uint32_t ko = (uint32_t)(k * 2);  // +32 bytes (>>4) inside the 128 B swizzle row
              umma_bf16(d_tmem, a_lo + ko, b_hi + ko, idesc, ((it - un.k0) | k) != 0);
              umma_bf16(d_tmem, a_hi + ko, b_lo + ko, idesc, 1u);
              umma_bf16(d_tmem, a_hi + ko, b_hi + ko, idesc, 1u);
            }
          } else {
            const UDesc a_hi = make_smem_desc(sa), b_hi = make_smem_desc(sa + p.off_bhi);
#pragma unroll
            for (int k = 0; k < TC_KCH / 16; ++k) {
              const uint32_t ko = (uint32_t)(k * 2);
              umma_bf16(d_tmem, a_hi + ko, b_hi + ko, idesc, ((it - un.k0) | k) != 0);
            }
          }
          umma_commit(&empty_bar[s]);  // frees the smem stage when these MMAs retire
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
        umma_commit(&tfull_bar[buf]);  // accumulator complete -> epilogue
      }
    }
  } else {
    // ============================================================ epilogue warps (2..9)
    // TMEM lane quarter = warp % 4 (hardware rule); the two warps that share a quarter split the
    // accumulator columns in interleaved 32-column chunks (half 0: 0, 64, ..; half 1: 32, 96, ..).
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    uint32_t tph[2] = {0, 0};
    for (int local = 0;; ++local) {
      const int u = tc_sched_read(local, lane, sched_ring, sfull_bar, sempty_bar);
      if (u < 0) break;
      const TcUnit un = tc_decode_unit(p, u, k_iters);
      tc_epilogue_item<EPI>(p, tmem_base, tfull_bar, tempty_bar, tph, local, un, q, half, lane, s_bias);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------ halo-reuse kernel
// Same GEMM as conv_tc_kernel, different operand traffic: for narrow layers (N <= 128) the A operand dominates
// the TMA request stream because every tap re-fetches an almost identical 128-pixel box.  Here the output patch
// is 16 rows x 8 pixels and ONE box of (16 + dy-range) x (8 + dx-range) pixels is loaded per (segment, 64-channel
// chunk, view); tap (dy, dx) is then just a different START ADDRESS into that tile (whole 128-byte rows, so the
// 128 B swizzle phase stays address-consistent) with the stride between 8-row groups = one halo row.
// A and B have their own pipelines (A: 1-2 halo tiles, B: one [N x 64] tile per tap).
constexpr int HALO_THREADS = 384;  // warp 0 A-TMA, 1 MMA, 2 B-TMA, 3 idle, 4..11 epilogue

// K-major SWIZZLE_128B descriptor whose 8-row groups are `sbo_bytes` apart (= one halo-tile pixel row).
// The swizzle XOR is taken from the shared-memory address bits, so a start address on any whole 128 B row
// of the 1024 B-aligned halo tile needs no base-offset field (measured: setting it breaks the results).
__device__ __forceinline__ UDesc make_smem_desc_sbo(uint32_t saddr, uint32_t sbo_bytes) {
  return make_udesc(saddr, 1u, sbo_bytes);
}

template <int EPI, int OCC>
__global__ void __launch_bounds__(HALO_THREADS, OCC) conv_tc_halo_kernel(const __grid_constant__ TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_base = smem;
  uint8_t* b_base = smem + (size_t)p.a_stages * p.a_stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_base + (size_t)p.b_stages * p.b_stage_bytes);
  uint64_t* a_full = bars;
  uint64_t* a_empty = bars + MAX_STAGES;
  uint64_t* b_full = bars + 2 * MAX_STAGES;
  uint64_t* b_empty = bars + 3 * MAX_STAGES;
  uint64_t* tfull_bar = bars + 4 * MAX_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* sfull_bar = tempty_bar + 2;            // scheduler ring (see tc_sched_fetch / tc_sched_read)
  uint64_t* sempty_bar = sfull_bar + SCHED_DEPTH;  // readers: B producer, MMA warp, 8 epilogue warps
  int* sched_ring = reinterpret_cast<int*>(sempty_bar + SCHED_DEPTH);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sched_ring + SCHED_DEPTH);
  // the shuffle tells ptxas the warp index is warp-uniform, so the role branches below are uniform branches
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  constexpr int BW = 8, BH = 16;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.a_stages; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < p.b_stages; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], 8); }
    for (int i = 0; i < SCHED_DEPTH; ++i) { mbar_init(&sfull_bar[i], 1); mbar_init(&sempty_bar[i], 10); }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // bias vector -> shared memory (read by every epilogue warp for every tile)
  float* s_bias_buf = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512);
  if (p.bias)
    for (int i = threadIdx.x; i < p.Cout; i += blockDim.x) s_bias_buf[i] = p.bias[i];
  const float* s_bias = p.bias ? s_bias_buf : nullptr;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ============================================================ A (halo tile) producer
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      const uint32_t tx = (uint32_t)(p.passes != 1 ? 2 : 1) * (uint32_t)(p.halo_w * p.halo_h * 128);
      for (int local = 0;; ++local) {
        const int item = tc_sched_fetch(p, local, sched_ring, sfull_bar, sempty_bar);   // whole tiles only
        if (item < 0) break;
        int mt = item / p.n_tiles;
        const int txi = mt % p.tiles_x; mt /= p.tiles_x;
        const int tyi = mt % p.tiles_y;
        const int n = mt / p.tiles_y;
        const int cx = txi * BW + p.hx0, cy = tyi * BH + p.hy0;
        for (int seg = 0; seg < p.nseg; ++seg)
          for (int c = 0; c < p.seg_chunks[seg]; ++c)
            for (int g = 0; g < p.n_groups; ++g) {
              const int v = p.seg_view0[seg] + p.grp_view[g];
              mbar_wait(&a_empty[s], ph ^ 1);
              uint8_t* st = a_base + (size_t)s * p.a_stage_bytes;
              mbar_expect_tx(&a_full[s], tx);
              tma_load_4d(st, &p.tmA_hi[v], &a_full[s], c * TC_KCH, cx, cy, n);
              if (p.passes != 1) tma_load_4d(st + p.a_lo_off, &p.tmA_lo[v], &a_full[s], c * TC_KCH, cx, cy, n);
              if (++s == p.a_stages) { s = 0; ph ^= 1; }
            }
      }
      tc_sched_finish(p);
    }
  } else if (warp == 2) {
    // ============================================================ B (weights) producer
    {
      int s = 0;
      uint32_t ph = 0;
      const uint32_t tx = (uint32_t)(p.passes != 1 ? 2 : 1) * (uint32_t)(p.BN * TC_KCH * 2);
      for (int local = 0;; ++local) {
        const int item = tc_sched_read(local, lane, sched_ring, sfull_bar, sempty_bar);
        if (item < 0) break;
        if (lane != 0) continue;
        const int nt = item % p.n_tiles;
        for (int seg = 0; seg < p.nseg; ++seg)
          for (int c = 0; c < p.seg_chunks[seg]; ++c)
            for (int g = 0; g < p.n_groups; ++g) {
              const int t_end = p.grp_first[g] + p.grp_count[g];
              for (int t0 = p.grp_first[g]; t0 < t_end; t0 += p.b_taps_per_stage) {
                if (p.b_split) {   // one plane of one tap per stage
                  const int kcoord = p.widx[t0] * p.k_per_tap + p.seg_koff[seg] + c * TC_KCH;
                  for (int pl = 0; pl < (p.passes != 1 ? 2 : 1); ++pl) {
                    mbar_wait(&b_empty[s], ph ^ 1);
                    mbar_expect_tx(&b_full[s], (uint32_t)(p.BN * TC_KCH * 2));
                    tma_load_2d(b_base + (size_t)s * p.b_stage_bytes, pl ? &p.tmB_lo : &p.tmB_hi, &b_full[s], kcoord, nt * p.BN);
                    if (++s == p.b_stages) { s = 0; ph ^= 1; }
                  }
                  continue;
                }
                const int nt_taps = min(p.b_taps_per_stage, t_end - t0);
                mbar_wait(&b_empty[s], ph ^ 1);
                uint8_t* st = b_base + (size_t)s * p.b_stage_bytes;
                mbar_expect_tx(&b_full[s], tx * (uint32_t)nt_taps);
                for (int j = 0; j < nt_taps; ++j) {
                  const int kcoord = p.widx[t0 + j] * p.k_per_tap + p.seg_koff[seg] + c * TC_KCH;
                  tma_load_2d(st + j * p.b_tap_bytes, &p.tmB_hi, &b_full[s], kcoord, nt * p.BN);
                  if (p.passes != 1)
                    tma_load_2d(st + j * p.b_tap_bytes + p.b_lo_off, &p.tmB_lo, &b_full[s], kcoord, nt * p.BN);
                }
                if (++s == p.b_stages) { s = 0; ph ^= 1; }
              }
            }
      }
    }
  } else if (warp == 1) {
    // ============================================================ MMA issuer (whole warp, convergent)
    {
      const uint32_t idesc = p.passes == 2 ? make_idesc_fmt(TC_M, p.BN, 0u, 0u) : make_idesc(TC_M, p.BN);
      const uint32_t idesc_f8 = make_idesc_fmt(TC_M, p.BN, 0u, 0u);
      const uint32_t idesc2 = make_idesc(TC_M, 2 * p.BN);
      const uint32_t sbo = (uint32_t)p.halo_w * 128u;
      int sa = 0, sb = 0;
      uint32_t pha = 0, phb = 0;
      uint32_t tph[2] = {0, 0};
      for (int local = 0;; ++local) {
        const int item = tc_sched_read(local, lane, sched_ring, sfull_bar, sempty_bar);
        if (item < 0) break;
        const int buf = local & 1;
        mbar_wait(&tempty_bar[buf], tph[buf] ^ 1);
        tph[buf] ^= 1;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * p.tmem_buf_stride);
        uint32_t acc = 0;
        for (int seg = 0; seg < p.nseg; ++seg)
          for (int c = 0; c < p.seg_chunks[seg]; ++c)
            for (int g = 0; g < p.n_groups; ++g) {
              mbar_wait(&a_full[sa], pha);
              tc_fence_after();
              const uint32_t a_addr = smem_u32(a_base + (size_t)sa * p.a_stage_bytes);
              const int t_end = p.grp_first[g] + p.grp_count[g];
              for (int t0 = p.grp_first[g]; t0 < t_end; t0 += p.b_taps_per_stage) {
               if (p.b_split) {
                 // wide tile: the hi-plane MMAs of the tap run from one B stage, the lo-plane MMAs from the next one
                 const uint32_t a_off = (uint32_t)((p.dy[t0] - p.hy0) * p.halo_w + (p.dx[t0] - p.hx0)) * 128u;
                 const UDesc a_hi = make_smem_desc_sbo(a_addr + a_off, sbo);
                 const UDesc a_lo = make_smem_desc_sbo(a_addr + p.a_lo_off + a_off, sbo);
                 mbar_wait(&b_full[sb], phb);
                 tc_fence_after();
                 {
                   const UDesc b_hi = make_smem_desc(smem_u32(b_base + (size_t)sb * p.b_stage_bytes));
#pragma unroll
                   for (int k = 0; k < TC_KCH / 16; ++k) {
                     const uint32_t ko = (uint32_t)(k * 2);
                     if (p.passes == 3) {
                       umma_bf16(d_tmem, a_lo + ko, b_hi + ko, idesc, acc);
                       umma_bf16(d_tmem, a_hi + ko, b_hi + ko, idesc, 1u);
                     } else {
                       umma_bf16(d_tmem, a_hi + ko, b_hi + ko, idesc, acc);
                     }
                     acc = 1u;
                   }
                 }
                 umma_commit(&b_empty[sb]);
                 if (++sb == p.b_stages) { sb = 0; phb ^= 1; }
                 if (p.passes != 1) {
                   mbar_wait(&b_full[sb], phb);
                   tc_fence_after();
                   const UDesc b_lo = make_smem_desc(smem_u32(b_base + (size_t)sb * p.b_stage_bytes));
#pragma unroll
                   for (int k = 0; k < TC_KCH / 16; ++k) {
                     const uint32_t ko = (uint32_t)(k * 2);
                     if (p.passes == 2) umma_f8(d_tmem, a_lo + ko, b_lo + ko, idesc_f8, 1u);
                     else umma_bf16(d_tmem, a_hi + ko, b_lo + ko, idesc, 1u);
                   }
                   umma_commit(&b_empty[sb]);
                   if (++sb == p.b_stages) { sb = 0; phb ^= 1; }
                 }
                 continue;
               }
               const int nt_taps = min(p.b_taps_per_stage, t_end - t0);
               mbar_wait(&b_full[sb], phb);
               tc_fence_after();
               for (int j = 0; j < nt_taps && !(p.base_offset_mode & 2); ++j) {
                const int t = t0 + j;
                const uint32_t b_addr = smem_u32(b_base + (size_t)sb * p.b_stage_bytes) + (uint32_t)(j * p.b_tap_bytes);
                const uint32_t a_off = (uint32_t)((p.dy[t] - p.hy0) * p.halo_w + (p.dx[t] - p.hx0)) * 128u;
                const UDesc a_hi = make_smem_desc_sbo(a_addr + a_off, sbo);
                const UDesc b_hi = make_smem_desc(b_addr);
                if (p.passes == 2) {
                  const UDesc a_lo = make_smem_desc_sbo(a_addr + p.a_lo_off + a_off, sbo);
                  const UDesc b_lo = make_smem_desc(b_addr + p.b_lo_off);
#pragma unroll
                  for (int k = 0; k < TC_KCH / 16; ++k) {
                    const uint32_t ko = (uint32_t)(k * 2);
                    umma_f8(d_tmem, a_lo + ko, b_lo + ko, idesc_f8, acc);
                    umma_bf16(d_tmem, a_hi + ko, b_hi + ko, idesc, 1u);
                    acc = 1u;
                  }
                } else if (p.passes == 3 && p.fuse_b) {
                  // B_hi and B_lo tiles are adjacent K-major tiles, i.e. ONE tile of 2*BN rows: A_hi x [B_hi; B_lo]
                  // is a single MMA of N = 2*BN whose upper BN accumulator columns collect the hi*lo products
                  // (added back by the epilogue).  Two MMAs and two A reads per K-step instead of three: the
                  // narrow layers are bound by the shared-memory operand feed, not by tensor math.
                  const UDesc a_lo = make_smem_desc_sbo(a_addr + p.a_lo_off + a_off, sbo);
#pragma unroll
                  for (int k = 0; k < TC_KCH / 16; ++k) {
                    const uint32_t ko = (uint32_t)(k * 2);
                    umma_bf16(d_tmem, a_hi + ko, b_hi + ko, idesc2, acc);
                    umma_bf16(d_tmem, a_lo + ko, b_hi + ko, idesc, 1u);
                    acc = 1u;
                  }
                } else if (p.passes == 3) {
                  const UDesc a_lo = make_smem_desc_sbo(a_addr + p.a_lo_off + a_off, sbo);
                  const UDesc b_lo = make_smem_desc(b_addr + p.b_lo_off);
#pragma unroll
                  for (int k = 0; k < TC_KCH / 16; ++k) {
                    const uint32_t ko = (uint32_t)(k * 2);
                    umma_bf16(d_tmem, a_lo + ko, b_hi + ko, idesc, acc);
                    umma_bf16(d_tmem, a_hi + ko, b_lo + ko, idesc, 1u);
                    umma_bf16(d_tmem, a_hi + ko, b_hi + ko, idesc, 1u);
                    acc = 1u;
                  }
                } else {
#pragma unroll
                  for (int k = 0; k < TC_KCH / 16; ++k) {
                    umma_bf16(d_tmem, a_hi + (uint32_t)(k * 2), b_hi + (uint32_t)(k * 2), idesc, acc);
                    acc = 1u;
                  }
                }
               }
               umma_commit(&b_empty[sb]);
               if (++sb == p.b_stages) { sb = 0; phb ^= 1; }
              }
              umma_commit(&a_empty[sa]);  // every tap of this view has been issued: the halo tile can be refilled
              if (++sa == p.a_stages) { sa = 0; pha ^= 1; }
            }
        umma_commit(&tfull_bar[buf]);
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    uint32_t tph[2] = {0, 0};
    for (int local = 0;; ++local) {
      const int item = tc_sched_read(local, lane, sched_ring, sfull_bar, sempty_bar);
      if (item < 0) break;
      const TcUnit un = {item, 0, 0, 0, -1};
      tc_epilogue_item<EPI>(p, tmem_base, tfull_bar, tempty_bar, tph, local, un, q, half, lane, s_bias);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}


// ------------------------------------------------------------------------------ CTA-pair kernel (cta_group::2)
// The N = 256 layers (ConvLSTM cells, deepest encoder conv, 256-wide decoder convs) with the halo-reuse operand scheme on
// a CTA PAIR: the two SMs of a TPC compute two horizontally adjacent 16 x 8 output patches (M = 256) with ONE
// tcgen05.mma.cta_group::2 per K-step.  Each CTA stages its own A halo tile but only HALF of the weight rows (128 of the
// 256 N rows, 16 KB per plane and tap instead of 32 KB): the tensor cores read the other half from the peer's shared
// memory.  Why: the single-CTA N = 256 kernels are bound by the SM's shared-memory port, which serves both the UMMA operand
// reads (A 4 KB + B 8 KB per 128-cycle MMA = 96 B/clk) and the TMA fills (67-94 B/clk at full tensor rate in the f16f8
// mode; ncu: tensor pipe 62-68 %).  The pair halves the B bytes on both paths (64 B/clk operand reads, ~36 B/clk fills).
// Protocol (as CUTLASS's 2-SM UMMA pipelines): TMA loads of both CTAs complete on the LEADER's full barrier (the leader
// arms it with the bytes of both CTAs); the leader issues every MMA and multicasts its commits to the empty / accumulator-
// full barriers of both CTAs; the epilogue warps of both CTAs arrive on the leader's accumulator-empty barrier.
// Static schedule (pair p takes tile pairs p, p + #pairs, ...): the dynamic scheduler bought nothing on these layers.
template <int EPI>
__global__ void __launch_bounds__(HALO_THREADS, 1) conv_tc_pair_kernel(const __grid_constant__ TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_base = smem;
  uint8_t* b_base = smem + (size_t)p.a_stages * p.a_stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_base + (size_t)p.b_stages * p.b_stage_bytes);
  uint64_t* a_full = bars;
  uint64_t* a_empty = bars + MAX_STAGES;
  uint64_t* b_full = bars + 2 * MAX_STAGES;
  uint64_t* b_empty = bars + 3 * MAX_STAGES;
  uint64_t* tfull_bar = bars + 4 * MAX_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  constexpr int BW = 8, BH = 16;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.a_stages; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < p.b_stages; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], 16); }   // 8 epilogue warps x 2 CTAs
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  float* s_bias_buf = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512);
  if (p.bias)
    for (int i = threadIdx.x; i < p.Cout; i += blockDim.x) s_bias_buf[i] = p.bias[i];
  const float* s_bias = p.bias ? s_bias_buf : nullptr;
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // barriers of both CTAs initialised before any remote arrive / multicast commit / 2-SM TMA
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int pair = (int)(blockIdx.x >> 1), npairs = (int)(gridDim.x >> 1);
  const int tx2n = (p.tiles_x + 1) >> 1;
  const int n_pair_items = p.N * p.tiles_y * tx2n * p.n_tiles;
  const int planes = p.passes != 1 ? 2 : 1;

  if (warp == 0) {
    // ============================================================ A (halo tile) producer, both CTAs
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      const uint32_t tx = (uint32_t)planes * (uint32_t)(p.halo_w * p.halo_h * 128);
      for (int t = pair; t < n_pair_items; t += npairs) {
        int mt = t / p.n_tiles;
        const int txi = 2 * (mt % tx2n) + (int)rank; mt /= tx2n;
        const int tyi = mt % p.tiles_y;
        const int n = mt / p.tiles_y;
        const int cx = txi * BW + p.hx0, cy = tyi * BH + p.hy0;
        for (int seg = 0; seg < p.nseg; ++seg)
          for (int c = 0; c < p.seg_chunks[seg]; ++c)
            for (int g = 0; g < p.n_groups; ++g) {
              const int v = p.seg_view0[seg] + p.grp_view[g];
              mbar_wait(&a_empty[s], ph ^ 1);
              uint8_t* st = a_base + (size_t)s * p.a_stage_bytes;
              if (leader) mbar_expect_tx(&a_full[s], 2u * tx);       // the halo tiles of both CTAs
              const uint32_t bar = mapa_u32(smem_u32(&a_full[s]), 0u);
              tma_load_4d_2sm(st, &p.tmA_hi[v], bar, c * TC_KCH, cx, cy, n);
              if (planes == 2) tma_load_4d_2sm(st + p.a_lo_off, &p.tmA_lo[v], bar, c * TC_KCH, cx, cy, n);
              if (++s == p.a_stages) { s = 0; ph ^= 1; }
            }
      }
    }
  } else if (warp == 2) {
    // ============================================================ B (weights) producer, both CTAs: this CTA's half of N
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      const int half_n = p.BN >> 1;
      const uint32_t tx = (uint32_t)(half_n * TC_KCH * 2);
      for (int t = pair; t < n_pair_items; t += npairs) {
        const int nt = t % p.n_tiles;
        const int row0 = nt * p.BN + (int)rank * half_n;
        for (int seg = 0; seg < p.nseg; ++seg)
          for (int c = 0; c < p.seg_chunks[seg]; ++c)
            for (int g = 0; g < p.n_groups; ++g) {
              const int t_end = p.grp_first[g] + p.grp_count[g];
              for (int t0 = p.grp_first[g]; t0 < t_end; t0 += p.b_taps_per_stage) {
                if (!p.b_split) {   // narrow tiles: a stage holds b_taps_per_stage taps, each [hi | lo]
                  const int nt_taps = min(p.b_taps_per_stage, t_end - t0);
                  mbar_wait(&b_empty[s], ph ^ 1);
                  if (leader) mbar_expect_tx(&b_full[s], 2u * tx * (uint32_t)(planes * nt_taps));
                  const uint32_t bar = mapa_u32(smem_u32(&b_full[s]), 0u);
                  for (int j = 0; j < nt_taps; ++j) {
                    const int kc = p.widx[t0 + j] * p.k_per_tap + p.seg_koff[seg] + c * TC_KCH;
                    uint8_t* st = b_base + (size_t)s * p.b_stage_bytes + (size_t)j * p.b_tap_bytes;
                    tma_load_2d_2sm(st, &p.tmB_hi, bar, kc, row0);
                    if (planes == 2) tma_load_2d_2sm(st + p.b_lo_off, &p.tmB_lo, bar, kc, row0);
                  }
                  if (++s == p.b_stages) { s = 0; ph ^= 1; }
                  continue;
                }
                const int kcoord = p.widx[t0] * p.k_per_tap + p.seg_koff[seg] + c * TC_KCH;
                for (int pl = 0; pl < planes; ++pl) {
                  mbar_wait(&b_empty[s], ph ^ 1);
                  if (leader) mbar_expect_tx(&b_full[s], 2u * tx);
                  const uint32_t bar = mapa_u32(smem_u32(&b_full[s]), 0u);
                  tma_load_2d_2sm(b_base + (size_t)s * p.b_stage_bytes, pl ? &p.tmB_lo : &p.tmB_hi, bar, kcoord, row0);
                  if (++s == p.b_stages) { s = 0; ph ^= 1; }
                }
              }
            }
      }
    }
  } else if (warp == 1) {
    // ============================================================ MMA issuer: leader CTA only (whole warp, convergent)
    if (leader) {
      const uint32_t idesc = p.passes == 2 ? make_idesc_fmt(2 * TC_M, p.BN, 0u, 0u) : make_idesc(2 * TC_M, p.BN);
      const uint32_t idesc_f8 = make_idesc_fmt(2 * TC_M, p.BN, 0u, 0u);
      const uint32_t sbo = (uint32_t)p.halo_w * 128u;
      int sa = 0, sb = 0;
      uint32_t pha = 0, phb = 0;
      uint32_t tph[2] = {0, 0};
      int local = 0;
      for (int t = pair; t < n_pair_items; t += npairs, ++local) {
        const int buf = local & 1;
        mbar_wait(&tempty_bar[buf], tph[buf] ^ 1);
        tph[buf] ^= 1;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * p.tmem_buf_stride);
        uint32_t acc = 0;
        for (int seg = 0; seg < p.nseg; ++seg)
          for (int c = 0; c < p.seg_chunks[seg]; ++c)
            for (int g = 0; g < p.n_groups; ++g) {
              mbar_wait(&a_full[sa], pha);
              tc_fence_after();
              const uint32_t a_addr = smem_u32(a_base + (size_t)sa * p.a_stage_bytes);
              const int t_end = p.grp_first[g] + p.grp_count[g];
              for (int t0 = p.grp_first[g]; t0 < t_end; t0 += p.b_taps_per_stage) {
                if (!p.b_split) {
                  const int nt_taps = min(p.b_taps_per_stage, t_end - t0);
                  mbar_wait(&b_full[sb], phb);
                  tc_fence_after();
                  for (int j = 0; j < nt_taps; ++j) {
                    const int t = t0 + j;
                    const uint32_t a_off = (uint32_t)((p.dy[t] - p.hy0) * p.halo_w + (p.dx[t] - p.hx0)) * 128u;
                    const UDesc a_hi = make_smem_desc_sbo(a_addr + a_off, sbo);
                    const UDesc a_lo = make_smem_desc_sbo(a_addr + p.a_lo_off + a_off, sbo);
                    const uint32_t b_addr = smem_u32(b_base + (size_t)sb * p.b_stage_bytes) + (uint32_t)(j * p.b_tap_bytes);
                    const UDesc b_hi = make_smem_desc(b_addr), b_lo = make_smem_desc(b_addr + p.b_lo_off);
#pragma unroll
                    for (int k = 0; k < TC_KCH / 16; ++k) {
                      const uint32_t ko = (uint32_t)(k * 2);
                      if (p.passes == 2) {
                        umma2_f8(d_tmem, a_lo + ko, b_lo + ko, idesc_f8, acc);
                        umma2_bf16(d_tmem, a_hi + ko, b_hi + ko, idesc, 1u);
                      } else if (p.passes == 3) {
                        umma2_bf16(d_tmem, a_lo + ko, b_hi + ko, idesc, acc);
                        umma2_bf16(d_tmem, a_hi + ko, b_lo + ko, idesc, 1u);
                        umma2_bf16(d_tmem, a_hi + ko, b_hi + ko, idesc, 1u);
                      } else {
                        umma2_bf16(d_tmem, a_hi + ko, b_hi + ko, idesc, acc);
                      }
                      acc = 1u;
                    }
                  }
                  umma2_commit_mc(&b_empty[sb]);
                  if (++sb == p.b_stages) { sb = 0; phb ^= 1; }
                  continue;
                }
                const uint32_t a_off = (uint32_t)((p.dy[t0] - p.hy0) * p.halo_w + (p.dx[t0] - p.hx0)) * 128u;
                const UDesc a_hi = make_smem_desc_sbo(a_addr + a_off, sbo);
                const UDesc a_lo = make_smem_desc_sbo(a_addr + p.a_lo_off + a_off, sbo);
                mbar_wait(&b_full[sb], phb);
                tc_fence_after();
                {
                  const UDesc b_hi = make_smem_desc(smem_u32(b_base + (size_t)sb * p.b_stage_bytes));
#pragma unroll
                  for (int k = 0; k < TC_KCH / 16; ++k) {
                    const uint32_t ko = (uint32_t)(k * 2);
                    if (p.passes == 3) {
                      umma2_bf16(d_tmem, a_lo + ko, b_hi + ko, idesc, acc);
                      umma2_bf16(d_tmem, a_hi + ko, b_hi + ko, idesc, 1u);
                    } else {
                      umma2_bf16(d_tmem, a_hi + ko, b_hi + ko, idesc, acc);
                    }
                    acc = 1u;
                  }
                }
                umma2_commit_mc(&b_empty[sb]);
                if (++sb == p.b_stages) { sb = 0; phb ^= 1; }
                if (p.passes != 1) {
                  mbar_wait(&b_full[sb], phb);
                  tc_fence_after();
                  const UDesc b_lo = make_smem_desc(smem_u32(b_base + (size_t)sb * p.b_stage_bytes));
#pragma unroll
                  for (int k = 0; k < TC_KCH / 16; ++k) {
                    const uint32_t ko = (uint32_t)(k * 2);
                    if (p.passes == 2) umma2_f8(d_tmem, a_lo + ko, b_lo + ko, idesc_f8, 1u);
                    else umma2_bf16(d_tmem, a_hi + ko, b_lo + ko, idesc, 1u);
                  }
                  umma2_commit_mc(&b_empty[sb]);
                  if (++sb == p.b_stages) { sb = 0; phb ^= 1; }
                }
              }
              umma2_commit_mc(&a_empty[sa]);
              if (++sa == p.a_stages) { sa = 0; pha ^= 1; }
            }
        umma2_commit_mc(&tfull_bar[buf]);
      }
    }
  } else if (warp >= 4) {
    // ============================================================ epilogue warps, both CTAs (each drains its own TMEM)
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    uint32_t tph[2] = {0, 0};
    const uint32_t tempty_cluster[2] = {mapa_u32(smem_u32(&tempty_bar[0]), 0u), mapa_u32(smem_u32(&tempty_bar[1]), 0u)};
    int local = 0;
    for (int t = pair; t < n_pair_items; t += npairs, ++local) {
      const int nt = t % p.n_tiles;
      int mt = t / p.n_tiles;
      const int txi = 2 * (mt % tx2n) + (int)rank; mt /= tx2n;
      if (txi >= p.tiles_x) {   // odd tile count per row: this CTA's half of the last pair lies outside the image
        const int buf = local & 1;
        mbar_wait(&tfull_bar[buf], tph[buf]);
        tph[buf] ^= 1;
        tc_fence_after();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(tempty_cluster[buf]);
        continue;
      }
      const TcUnit un = {(mt * p.tiles_x + txi) * p.n_tiles + nt, 0, 0, 0, -1};
      tc_epilogue_item<EPI>(p, tmem_base, tfull_bar, tempty_bar, tph, local, un, q, half, lane, s_bias, tempty_cluster);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // the peer may still multicast into / arrive on this CTA's barriers until both are done
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// --------------------------------------------------------------------------- fp32 -> bf16 hi/lo
__global__ void split_bf16_kernel(essb_src s, int N, int H, int W, __nv_bfloat16* __restrict__ hi,
                                  __nv_bfloat16* __restrict__ lo, int ld_out, int c_off, int c_write, long long total,
                                  int fmt) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int CQ = c_write >> 2;
  const int c = (int)(idx % CQ) * 4;
  const long long pix = idx / CQ;
  uint8_t* lo8 = reinterpret_cast<uint8_t*>(lo) + pix * (long long)ld_out * 2 + hf8_lo_off(c_off + c);   // fmt 2 only
  if (c >= s.C) {  // zero channel padding [C, c_write) (K padding of the tensor-core operand)
    *reinterpret_cast<uint2*>(hi + pix * ld_out + c_off + c) = make_uint2(0u, 0u);
    if (fmt == 2) {
      *reinterpret_cast<uint32_t*>(lo8) = 0u;
      *reinterpret_cast<uint32_t*>(lo8 + 64) = 0u;
    } else {
      *reinterpret_cast<uint2*>(lo + pix * ld_out + c_off + c) = make_uint2(0u, 0u);
    }
    return;
  }
  const long long P = (long long)H * W;
  const int n = (int)(pix / P);
  const long long pp = pix - (long long)n * P;
  const int y = (int)(pp / W), x = (int)(pp - (long long)y * W);
  const int Hs = H >> s.ups, Ws = W >> s.ups;
  const size_t sp = ((size_t)n * Hs + (y >> s.ups)) * Ws + (x >> s.ups);
  float4 v = *reinterpret_cast<const float4*>(s.ptr + sp * s.ld + c);
  if (s.mean) {
    const float4 m = *reinterpret_cast<const float4*>(s.mean + (size_t)n * s.C + c);
    const float4 r = *reinterpret_cast<const float4*>(s.rstd + (size_t)n * s.C + c);
    v.x = (v.x - m.x) * r.x; v.y = (v.y - m.y) * r.y; v.z = (v.z - m.z) * r.z; v.w = (v.w - m.w) * r.w;
  }
  if (s.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
  if (fmt == 2) {
    uint32_t h2[2];
    uint16_t a[2], l2[2];
    split_hf8x2(v.x, v.y, h2[0], a[0], l2[0]);
    split_hf8x2(v.z, v.w, h2[1], a[1], l2[1]);
    *reinterpret_cast<uint2*>(hi + pix * ld_out + c_off + c) = make_uint2(h2[0], h2[1]);
    *reinterpret_cast<uint32_t*>(lo8) = (uint32_t)a[0] | ((uint32_t)a[1] << 16);
    *reinterpret_cast<uint32_t*>(lo8 + 64) = (uint32_t)l2[0] | ((uint32_t)l2[1] << 16);
    return;
  }
  __nv_bfloat16 h[4], l[4];
  split_bf16(v.x, h[0], l[0]); split_bf16(v.y, h[1], l[1]); split_bf16(v.z, h[2], l[2]); split_bf16(v.w, h[3], l[3]);
  __nv_bfloat16* ho = hi + pix * ld_out + c_off + c;
  __nv_bfloat16* lo_ = lo + pix * ld_out + c_off + c;
  *reinterpret_cast<uint2*>(ho) = *reinterpret_cast<uint2*>(h);
  *reinterpret_cast<uint2*>(lo_) = *reinterpret_cast<uint2*>(l);
}

// Fast path of split_bf16_kernel: 8 channels per thread (two 128-bit loads, 128-bit plane stores), 32-bit index arithmetic,
// the upsampling coordinates only when the source IS upsampled.  Same values as the general kernel.
template <int UPS, int FMT>
__global__ void __launch_bounds__(256) split_planes8_kernel(essb_src s, int H, int W, __nv_bfloat16* __restrict__ hi,
                                                            __nv_bfloat16* __restrict__ lo, int ld_out, int c_off, int cq8,
                                                            unsigned total) {
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const unsigned pix = idx / (unsigned)cq8;
  const int c = (int)(idx - pix * (unsigned)cq8) * 8;
  const unsigned P = (unsigned)(H * W);
  __nv_bfloat16* ho = hi + (size_t)pix * ld_out + c_off + c;
  if (c >= s.C) {  // zero channel padding [C, c_write)
    *reinterpret_cast<uint4*>(ho) = make_uint4(0u, 0u, 0u, 0u);
    if (FMT == 2) {
      uint8_t* l8 = reinterpret_cast<uint8_t*>(lo) + (size_t)pix * ld_out * 2 + hf8_lo_off(c_off + c);
      *reinterpret_cast<uint2*>(l8) = make_uint2(0u, 0u);
      *reinterpret_cast<uint2*>(l8 + 64) = make_uint2(0u, 0u);
    } else {
      *reinterpret_cast<uint4*>(lo + (size_t)pix * ld_out + c_off + c) = make_uint4(0u, 0u, 0u, 0u);
    }
    return;
  }
  const unsigned n = (s.mean || UPS) ? pix / P : 0u;
  size_t sp = pix;
  if (UPS) {
    const unsigned pp = pix - n * P;
    const unsigned y = pp / (unsigned)W, x = pp - y * (unsigned)W;
    sp = ((size_t)n * (H >> 1) + (y >> 1)) * (W >> 1) + (x >> 1);
  }
  const float* src = s.ptr + sp * s.ld + c;
  float v[8];
  {
    const float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  if (s.mean) {
    const float* mp = s.mean + (size_t)n * s.C + c;
    const float* rp = s.rstd + (size_t)n * s.C + c;
    const float4 m0 = *reinterpret_cast<const float4*>(mp), m1 = *reinterpret_cast<const float4*>(mp + 4);
    const float4 q0 = *reinterpret_cast<const float4*>(rp), q1 = *reinterpret_cast<const float4*>(rp + 4);
    const float m[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
    const float q[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = (v[e] - m[e]) * q[e];
  }
  if (s.relu) {
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], 0.f);
  }
  if (FMT == 2) {
    uint32_t h2[4];
    uint16_t a[4], l2[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) split_hf8x2(v[2 * e], v[2 * e + 1], h2[e], a[e], l2[e]);
    *reinterpret_cast<uint4*>(ho) = make_uint4(h2[0], h2[1], h2[2], h2[3]);
    uint8_t* l8 = reinterpret_cast<uint8_t*>(lo) + (size_t)pix * ld_out * 2 + hf8_lo_off(c_off + c);
    *reinterpret_cast<uint2*>(l8) = make_uint2((uint32_t)a[0] | ((uint32_t)a[1] << 16), (uint32_t)a[2] | ((uint32_t)a[3] << 16));
    *reinterpret_cast<uint2*>(l8 + 64) = make_uint2((uint32_t)l2[0] | ((uint32_t)l2[1] << 16), (uint32_t)l2[2] | ((uint32_t)l2[3] << 16));
  } else {
    bf16x8 hh, hl;
#pragma unroll
    for (int e = 0; e < 8; ++e) split_bf16(v[e], hh.v[e], hl.v[e]);
    *reinterpret_cast<bf16x8*>(ho) = hh;
    *reinterpret_cast<bf16x8*>(lo + (size_t)pix * ld_out + c_off + c) = hl;
  }
}

// Event pre-processing straight into the head convolution's operand format: normalise (as
// event_prepare_kernel in pointwise.cu), reflect-pad to (Hp, Wp), NCHW -> pixel-major with `cpad`
// channels, split to bf16 hi/lo, and store at offset (off_y, off_x) inside a zero-bordered buffer
// [B][Hb][Wb][cpad].  The zero border is the 5x5 convolution's padding, so the head conv can read
// 8-pixel x cpad-channel row windows as ONE contiguous K-chunk through an overlapping TMA view.
__device__ __forceinline__ int reflect_idx_tc(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}
template <int CPAD>
__global__ void event_prepare_planes_kernel(const float* __restrict__ x, long long bstride,
                                            const double* __restrict__ stats, int normalize,
                                            __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int C, int H,
                                            int W, int Hp, int Wp, int pad_top, int pad_left, int Hb, int Wb, int off_y,
                                            int off_x, int flip) {
  // grid (ceil(Wp / blockDim.x), ceil(Hp / 2), B): no index division; a thread converts the pixels (oy, ox) and
  // (oy + 1, ox) and issues the loads of both before any arithmetic (the kernel is latency-bound otherwise)
  const int ox = blockIdx.x * blockDim.x + threadIdx.x;
  if (ox >= Wp) return;
  const int oy0 = blockIdx.y * 2, n = blockIdx.z;
  float mean = 0.f, inv_std = 1.f;
  bool do_norm = false;
  if (normalize) {
    const double nnz = stats[2];
    if (nnz > 0.0) {
      const float fm = (float)stats[0] / (float)nnz;
      inv_std = 1.f / sqrtf((float)stats[1] / (float)nnz - fm * fm);
      mean = fm;
      do_norm = true;
    }
  }
  int ix = reflect_idx_tc(ox - pad_left, W);
  if (flip) ix = W - 1 - ix;                            // torch.flip(events, dims=[2, 3]) precedes the padding
  float v[2][CPAD];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int oy = min(oy0 + r, Hp - 1);
    int iy = reflect_idx_tc(oy - pad_top, H);
    if (flip) iy = H - 1 - iy;
    const float* src = x + (long long)n * bstride + (long long)iy * W + ix;
#pragma unroll
    for (int c = 0; c < CPAD; ++c) v[r][c] = (c < C) ? src[(long long)c * H * W] : 0.f;
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int oy = oy0 + r;
    if (oy >= Hp) break;
    __align__(16) __nv_bfloat16 h[CPAD], l[CPAD];
#pragma unroll
    for (int c = 0; c < CPAD; ++c) {
      float t = v[r][c];
      if (do_norm) t = (t != 0.f) ? (t - mean) * inv_std : 0.f;
      split_bf16(t, h[c], l[c]);
    }
    const size_t dst = (((size_t)n * Hb + oy + off_y) * Wb + ox + off_x) * CPAD;
#pragma unroll
    for (int c = 0; c < CPAD; c += 8) {
      *reinterpret_cast<uint4*>(hi + dst + c) = *reinterpret_cast<const uint4*>(h + c);
      *reinterpret_cast<uint4*>(lo + dst + c) = *reinterpret_cast<const uint4*>(l + c);
    }
  }
}

// packed K-major weights: out[np][t*KinP + k], np < NoutP (zero rows beyond Nout)
__global__ void pack_weight_tc_kernel(const float* __restrict__ w, const float* __restrict__ scale,
                                      __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int Cout, int Cin,
                                      int T, int transposed_layout, int swap_io, int flip, int interleave, int KinP,
                                      int NoutP, int fmt, int w8) {
  const int Kin = swap_io ? Cout : Cin;
  const int Nout = swap_io ? Cin : Cout;
  const long long total = (long long)NoutP * T * KinP;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int k = (int)(idx % KinP);
  const long long r = idx / KinP;
  const int t = (int)(r % T);
  const int np = (int)(r / T);
  float v = 0.f;
  if (np < Nout && k < Kin) {
    int co, ci;
    if (swap_io) { co = k; ci = np; }
    else {
      ci = k;
      co = np;
      if (interleave > 1) {
        const int G = Cout / interleave;
        co = (np % interleave) * G + np / interleave;
      }
    }
    const int ts = flip ? T - 1 - t : t;
    const size_t src = transposed_layout ? ((size_t)ci * Cout + co) * T + ts : ((size_t)co * Cin + ci) * T + ts;
    v = w[src];
    if (scale && !swap_io) v *= scale[co];
  }
  if (fmt == 2) {   // hf8 weights: fp16(W * 2^(w8+8)) and the [w8l | w8] pair row (see tc_ptx.cuh)
    const float sh = exp2f((float)(w8 + 8));
    const __half wh = __float2half_rn(hf8_sat_f16(v * sh));
    reinterpret_cast<__half*>(hi)[idx] = wh;
    const float r = v - __half2float(wh) / sh;
    const long long kk = (long long)t * KinP + k;
    uint8_t* row = reinterpret_cast<uint8_t*>(lo) + (long long)np * T * KinP * 2 + ((kk >> 6) << 7) + (kk & 63);
    row[0] = (uint8_t)__nv_cvt_float_to_fp8(r * exp2f((float)(w8 + 11)), __NV_SATFINITE, __NV_E4M3);
    row[64] = (uint8_t)__nv_cvt_float_to_fp8(v * exp2f((float)w8), __NV_SATFINITE, __NV_E4M3);
    return;
  }
  __nv_bfloat16 h, l;
  split_bf16(v, h, l);
  hi[idx] = h;
  lo[idx] = l;
}

// ------------------------------------------------------------------------ tensor-map encoding
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(sym);
  }
  return fn;
}

int encode_a_map(CUtensorMap* tm, const void* base, const essb_tc_view& v, int N, int BW, int BH) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) {
    essb_set_error("conv_tc: cuTensorMapEncodeTiled unavailable");
    return ESSB_ERR_DRIVER;
  }
  cuuint64_t dims[4] = {(cuuint64_t)v.C, (cuuint64_t)v.W, (cuuint64_t)v.H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)v.stride_x * 2, (cuuint64_t)v.stride_y * 2, (cuuint64_t)v.stride_n * 2};
  cuuint32_t box[4] = {(cuuint32_t)TC_KCH, (cuuint32_t)BW, (cuuint32_t)BH, 1u};
  cuuint32_t es[4] = {1u, 1u, 1u, 1u};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    essb_set_error("conv_tc: cuTensorMapEncodeTiled(A) failed with %d (C=%d W=%d H=%d N=%d strides %lld %lld %lld)",
                   (int)r, v.C, v.W, v.H, N, (long long)v.stride_x, (long long)v.stride_y, (long long)v.stride_n);
    return ESSB_ERR_DRIVER;
  }
  return ESSB_OK;
}

int encode_b_map(CUtensorMap* tm, const void* base, long long ktot, int rows, int BN) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) {
    essb_set_error("conv_tc: cuTensorMapEncodeTiled unavailable");
    return ESSB_ERR_DRIVER;
  }
  cuuint64_t dims[2] = {(cuuint64_t)ktot, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ktot * 2};
  cuuint32_t box[2] = {(cuuint32_t)TC_KCH, (cuuint32_t)BN};
  cuuint32_t es[2] = {1u, 1u};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    essb_set_error("conv_tc: cuTensorMapEncodeTiled(B) failed with %d (ktot=%lld rows=%d BN=%d)", (int)r, ktot, rows, BN);
    return ESSB_ERR_DRIVER;
  }
  return ESSB_OK;
}

int g_num_sms = 0;
int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

}  // namespace

extern "C" int essb_split_bf16(const essb_src* src, int N, int H, int W, uint16_t* hi, uint16_t* lo, int ld_out,
                               int c_off, int c_pad, void* stream) {
  return essb_split_planes(src, N, H, W, hi, lo, ld_out, c_off, c_pad, 0, stream);
}

extern "C" int essb_split_planes(const essb_src* src, int N, int H, int W, uint16_t* hi, uint16_t* lo, int ld_out,
                                 int c_off, int c_pad, int fmt, void* stream) {
  ESSB_REQUIRE(src && src->ptr && hi && lo && N > 0 && H > 0 && W > 0, "essb_split_bf16: bad arguments");
  ESSB_REQUIRE(fmt == 0 || (fmt == 2 && ld_out % 64 == 0), "essb_split_planes: fmt must be 0 (bf16 hi/lo) or 2 (hf8, ld_out %% 64 == 0)");
  ESSB_REQUIRE(src->C % 4 == 0 && src->ld % 4 == 0 && essb_aligned16(src->ptr) && ld_out % 4 == 0 && c_off % 4 == 0,
               "essb_split_bf16: C, ld, ld_out, c_off must be multiples of 4 and pointers 16B aligned");
  ESSB_REQUIRE((src->mean == nullptr) == (src->rstd == nullptr), "essb_split_bf16: mean/rstd must come together");
  const int c_write = c_pad > src->C ? c_pad : src->C;
  ESSB_REQUIRE(c_write % 4 == 0 && c_off + c_write <= ld_out, "essb_split_bf16: c_off + max(C, c_pad) exceeds ld_out");
  // fast path: 8 channels per thread, 32-bit indices (every shape of the decoder / encoder paths)
  {
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    const long long total8 = (long long)N * H * W * (c_write / 8);
    if (src->C % 8 == 0 && c_write % 8 == 0 && c_off % 8 == 0 && ld_out % 8 == 0 && src->ld % 4 == 0 && al16(hi) && al16(lo) &&
        (src->ups == 0 || (src->ups == 1 && H % 2 == 0 && W % 2 == 0)) && total8 < (1ll << 31) &&
        (long long)N * H * W < (1ll << 31) && (!src->mean || (al16(src->mean) && al16(src->rstd)))) {
      const unsigned blocks = (unsigned)((total8 + 255) / 256);
      __nv_bfloat16* h = reinterpret_cast<__nv_bfloat16*>(hi);
      __nv_bfloat16* l = reinterpret_cast<__nv_bfloat16*>(lo);
      cudaStream_t st = (cudaStream_t)stream;
      const int cq8 = c_write / 8;
      if (src->ups == 0 && fmt == 0) split_planes8_kernel<0, 0><<<blocks, 256, 0, st>>>(*src, H, W, h, l, ld_out, c_off, cq8, (unsigned)total8);
      else if (src->ups == 0) split_planes8_kernel<0, 2><<<blocks, 256, 0, st>>>(*src, H, W, h, l, ld_out, c_off, cq8, (unsigned)total8);
      else if (fmt == 0) split_planes8_kernel<1, 0><<<blocks, 256, 0, st>>>(*src, H, W, h, l, ld_out, c_off, cq8, (unsigned)total8);
      else split_planes8_kernel<1, 2><<<blocks, 256, 0, st>>>(*src, H, W, h, l, ld_out, c_off, cq8, (unsigned)total8);
      ESSB_LAUNCH_CHECK("essb_split_planes (8 channels per thread)");
      return ESSB_OK;
    }
  }
  const long long total = (long long)N * H * W * (c_write / 4);
  split_bf16_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      *src, N, H, W, reinterpret_cast<__nv_bfloat16*>(hi), reinterpret_cast<__nv_bfloat16*>(lo), ld_out, c_off, c_write,
      total, fmt);
  ESSB_LAUNCH_CHECK("essb_split_bf16");
  return ESSB_OK;
}

extern "C" int essb_event_prepare_planes(const float* x, int64_t bstride, const double* stats, int normalize,
                                         uint16_t* hi, uint16_t* lo, int cpad, int B, int C, int H, int W, int Hp, int Wp,
                                         int pad_top, int pad_left, int Hb, int Wb, int off_y, int off_x, int flip,
                                         void* stream) {
  ESSB_REQUIRE(x && hi && lo && B > 0 && C > 0 && H > 0 && W > 0 && Hp >= H && Wp >= W, "essb_event_prepare_planes: bad arguments");
  ESSB_REQUIRE((cpad == 8 || cpad == 16) && C <= cpad, "essb_event_prepare_planes: cpad must be 8 or 16 and >= C");
  ESSB_REQUIRE(!normalize || stats, "essb_event_prepare_planes: stats required when normalising");
  ESSB_REQUIRE(off_y >= 0 && off_x >= 0 && Hp + off_y <= Hb && Wp + off_x <= Wb, "essb_event_prepare_planes: image does not fit the padded buffer");
  ESSB_REQUIRE(pad_top < H && Hp - H - pad_top < H && pad_left < W && Wp - W - pad_left < W,
               "essb_event_prepare_planes: reflection padding must be smaller than the image");
  ESSB_REQUIRE(essb_aligned16(hi) && essb_aligned16(lo), "essb_event_prepare_planes: planes must be 16B aligned");
  ESSB_REQUIRE(B <= 65535 && Hp <= 65535, "essb_event_prepare_planes: B and Hp must fit a grid dimension");
  const int threads = Wp >= 512 ? 128 : 64;
  const dim3 grid((unsigned)((Wp + threads - 1) / threads), (unsigned)((Hp + 1) / 2), (unsigned)B);
  __nv_bfloat16* h = reinterpret_cast<__nv_bfloat16*>(hi);
  __nv_bfloat16* l = reinterpret_cast<__nv_bfloat16*>(lo);
  if (cpad == 8)
    event_prepare_planes_kernel<8><<<grid, threads, 0, (cudaStream_t)stream>>>(x, bstride, stats, normalize, h, l, C, H, W, Hp,
                                                                           Wp, pad_top, pad_left, Hb, Wb, off_y, off_x, flip);
  else
    event_prepare_planes_kernel<16><<<grid, threads, 0, (cudaStream_t)stream>>>(x, bstride, stats, normalize, h, l, C, H, W, Hp,
                                                                            Wp, pad_top, pad_left, Hb, Wb, off_y, off_x, flip);
  ESSB_LAUNCH_CHECK("essb_event_prepare_planes");
  return ESSB_OK;
}

extern "C" int essb_pack_weight_tc(const float* w, const float* scale, uint16_t* hi, uint16_t* lo, int Cout, int Cin,
                                   int T, int transposed_layout, int swap_io, int flip, int interleave, int KinP,
                                   int NoutP, void* stream) {
  return essb_pack_weight_tc_fmt(w, scale, hi, lo, Cout, Cin, T, transposed_layout, swap_io, flip, interleave, KinP, NoutP,
                                 0, 0, stream);
}

extern "C" int essb_pack_weight_tc_fmt(const float* w, const float* scale, uint16_t* hi, uint16_t* lo, int Cout, int Cin,
                                       int T, int transposed_layout, int swap_io, int flip, int interleave, int KinP,
                                       int NoutP, int fmt, int w8, void* stream) {
  ESSB_REQUIRE(w && hi && lo && Cout > 0 && Cin > 0 && T > 0, "essb_pack_weight_tc: bad arguments");
  ESSB_REQUIRE(fmt == 0 || fmt == 2, "essb_pack_weight_tc_fmt: fmt must be 0 (bf16 hi/lo) or 2 (hf8)");
  ESSB_REQUIRE(w8 >= -20 && w8 <= 40, "essb_pack_weight_tc_fmt: w8 (log2 of the e4m3 weight scale) out of range");
  const int Kin = swap_io ? Cout : Cin, Nout = swap_io ? Cin : Cout;
  ESSB_REQUIRE(KinP >= Kin && KinP % TC_KCH == 0 && NoutP >= Nout, "essb_pack_weight_tc: bad padding");
  ESSB_REQUIRE(interleave <= 1 || (Cout % interleave == 0 && !swap_io), "essb_pack_weight_tc: bad interleave");
  const long long total = (long long)NoutP * T * KinP;
  pack_weight_tc_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      w, scale, reinterpret_cast<__nv_bfloat16*>(hi), reinterpret_cast<__nv_bfloat16*>(lo), Cout, Cin, T,
      transposed_layout, swap_io, flip, interleave, KinP, NoutP, fmt, w8);
  ESSB_LAUNCH_CHECK("essb_pack_weight_tc");
  return ESSB_OK;
}

extern "C" int essb_conv_tc_run(const essb_conv_tc* d, void* stream) {
  ESSB_REQUIRE(d != nullptr, "essb_conv_tc_run: null descriptor");
  ESSB_REQUIRE(d->n_views >= 1 && d->n_views <= MAX_VIEWS, "essb_conv_tc_run: n_views=%d", d->n_views);
  ESSB_REQUIRE(d->nseg == 1 || d->nseg == 2, "essb_conv_tc_run: nseg=%d", d->nseg);
  ESSB_REQUIRE(d->ntaps >= 1 && d->ntaps <= ESSB_MAX_TAPS, "essb_conv_tc_run: ntaps=%d", d->ntaps);
  ESSB_REQUIRE(d->passes >= 1 && d->passes <= 3, "essb_conv_tc_run: passes must be 1 (bf16), 2 (f16f8) or 3 (bf16x3)");
  ESSB_REQUIRE(d->N > 0 && d->OH > 0 && d->OW > 0 && d->Cout > 0, "essb_conv_tc_run: bad dims");
  ESSB_REQUIRE(d->w_hi && (d->passes == 1 || d->w_lo), "essb_conv_tc_run: null weights");
  ESSB_REQUIRE(d->bw_log2 >= 0 && d->bw_log2 <= 7, "essb_conv_tc_run: bw_log2=%d", d->bw_log2);
  ESSB_REQUIRE(d->k_per_tap % TC_KCH == 0, "essb_conv_tc_run: k_per_tap must be a multiple of 64");
  ESSB_REQUIRE(d->planes_fmt == 0 || d->planes_fmt == 2, "essb_conv_tc_run: planes_fmt must be 0 (bf16 hi/lo) or 2 (hf8)");
  ESSB_REQUIRE(d->planes_fmt == 0 || !d->out_hi || d->ld_planes % 64 == 0,
               "essb_conv_tc_run: hf8 output planes need ld_planes %% 64 == 0 (got %d)", d->ld_planes);
  const int Ngemm = d->Cout;
  int BN = Ngemm < 256 ? Ngemm : 256;
  ESSB_REQUIRE(BN % 32 == 0 && Ngemm % BN == 0, "essb_conv_tc_run: Cout=%d must be 32/64/128 or a multiple of 256", Ngemm);
  ESSB_REQUIRE(d->w_rows >= Ngemm, "essb_conv_tc_run: packed weight rows %d < Cout %d", d->w_rows, Ngemm);
  if (d->epilogue == ESSB_EPI_LSTM) {
    ESSB_REQUIRE(d->out && d->out2 && Ngemm % 4 == 0, "essb_conv_tc_run: LSTM needs out/out2");
    ESSB_REQUIRE(!d->out_hi || (d->out_lo && d->ld_planes % 8 == 0), "essb_conv_tc_run: bad planes");
  } else if (d->epilogue == ESSB_EPI_GRU_UR) {
    ESSB_REQUIRE(d->out && d->out_hi && d->out_lo && d->ld_planes % 8 == 0, "essb_conv_tc_run: GRU_UR needs out (update) and the bf16 planes of prev_state*reset");
  } else if (d->epilogue == ESSB_EPI_GRU_OUT) {
    ESSB_REQUIRE(d->out && d->aux1, "essb_conv_tc_run: GRU_OUT needs out and the update gate in aux1");
    ESSB_REQUIRE(!d->out_hi || (d->out_lo && d->ld_planes % 8 == 0), "essb_conv_tc_run: bad planes");
  } else {
    ESSB_REQUIRE(d->epilogue == ESSB_EPI_LINEAR, "essb_conv_tc_run: unknown epilogue %d", d->epilogue);
    ESSB_REQUIRE(d->out || d->out_hi, "essb_conv_tc_run: no output");
    ESSB_REQUIRE(!d->out || (d->ldo % 4 == 0 && essb_aligned16(d->out)), "essb_conv_tc_run: out must be 16B aligned, ldo %% 4 == 0");
    ESSB_REQUIRE(!d->out_hi || (d->out_lo && d->ld_planes % 8 == 0), "essb_conv_tc_run: bad planes");
    ESSB_REQUIRE(!(d->res_pre || d->res_post) || d->ld_res % 4 == 0, "essb_conv_tc_run: ld_res %% 4 != 0");
  }

  static thread_local TcParams p;  // 2.5 KB; filled per call, copied into the launch
  // ---- halo-reuse mode?  Narrow layers (N <= 128) with several taps: one halo tile per view feeds all its taps.
  //      ESSB_TC_HALO=0 disables it (A/B comparison).  Measured on B200: the 128 B swizzle of both TMA writes and
  //      UMMA reads is a pure function of the shared-memory ADDRESS bits, so a descriptor may start at any
  //      128 B-aligned row of a TMA-written tile with base_offset = 0 (setting base_offset from the address
  //      gives wrong results -- tried, tests/test_gpu_tc.py fails).
  static const int halo_env = [] { const char* e = getenv("ESSB_TC_HALO"); return e ? atoi(e) : 1; }();
  const int baseoff_env = [] { const char* e = getenv("ESSB_TC_DEBUG"); return e ? atoi(e) : 0; }();   // experiments only
  // ESSB_TC_HALO256=0 keeps the N = 256 layers (ConvLSTM cells, deepest encoder conv) on the classic kernel.  Measured
  // (ncu, profiles/r02c_window.txt): in the f16f8 mode the classic N = 256 kernel is bound by the SM's L2 -> shared-memory
  // fill (58.6 B/clk/SM of ~64: every tap re-fetches its 32 KB A box next to the 64 KB of weights), tensor pipe 62 %; the
  // halo tile cuts the A traffic 9x -> 1.4x (1728 -> 1244 KB per 128 x 256 tile of a 64-channel ConvLSTM).
  // (read per call, not cached: tests/test_gpu_tc.py flips it to keep both kernels covered)
  const int halo256_env = [] { const char* e = getenv("ESSB_TC_HALO256"); return e ? atoi(e) : 1; }();
  bool halo = halo_env != 0 && d->ntaps >= 2 && (BN <= 128 || (halo256_env != 0 && BN == 256));
  bool wide_halo = halo && BN == 256;
  if (wide_halo) {
    // The halo patch is fixed at 16 rows x 8 pixels and CTA pairs take tile pairs in whole rounds; the classic kernel picks
    // its patch shape per layer.  Stay classic when that saves whole scheduling rounds: 256-channel decoder convs at 55 x 80,
    // B = 8: 160 tile pairs on 74 CTA pairs = 3 rounds, classic 8 x 16 patches = 280 tiles on 148 CTAs = 2 rounds
    // (measured, profiles/r02i_bench_dense_l2classic.json: 0.161 -> 0.140 ms per decoder conv).  A classic tile costs ~1.1x a
    // pair tile (ConvLSTM cells: 0.54 vs 0.52 ms at equal rounds).  ESSB_TC_HALO256_WASTE (max ratio of padded pixel
    // counts, default off) is the older A/B switch.
    const double waste_env = [] { const char* e = getenv("ESSB_TC_HALO256_WASTE"); return e ? atof(e) : 0.0; }();
    const int cbw = 1 << d->bw_log2, cbh = TC_M >> d->bw_log2;
    const long long n_nt = Ngemm / BN;
    const long long t_halo = (long long)((d->OW + 7) / 8) * ((d->OH + 15) / 16);
    const long long t_classic = (long long)((d->OW + cbw - 1) / cbw) * ((d->OH + cbh - 1) / cbh);
    const long long pair_items = (long long)d->N * ((d->OH + 15) / 16) * (((d->OW + 7) / 8 + 1) / 2) * n_nt;
    const long long npairs = num_sms() / 2, nsm = num_sms();
    const long long rounds_pair = (pair_items + npairs - 1) / npairs;
    const long long rounds_classic = ((long long)d->N * t_classic * n_nt + nsm - 1) / nsm;
    const int pair_mode = [] { const char* e = getenv("ESSB_TC_PAIR"); return e ? atoi(e) : 1; }();   // 0 off, 1 by rounds, 2 always
    const bool pair_wanted = pair_mode != 0;
    bool classic = false;
    if (waste_env > 0.0) classic = (double)t_halo > waste_env * (double)t_classic;
    else if (pair_mode >= 2) classic = false;
    else if (pair_wanted) classic = 10 * rounds_pair > 11 * rounds_classic;
    else classic = (double)t_halo > 2.0 * (double)t_classic;
    if (classic) { halo = false; wide_halo = false; }
  }
  // shared-memory bias staging area: as small as the layer allows (it competes with pipeline stages at 2 CTAs/SM)
  const int bias_floats = d->bias ? ((d->Cout + 255) / 256) * 256 : 0;
  // CTA pairs (cta_group::2) for the wide halo tiles; ESSB_TC_PAIR=0 keeps them on single CTAs (read per call, see above)
  const int pair_env = [] { const char* e = getenv("ESSB_TC_PAIR"); return e ? atoi(e) : 1; }();
  bool pair = wide_halo && pair_env != 0;
  // CTA pairs for N = 128 tiles of the f16f8 mode (second encoder conv): OFF by default (ESSB_TC_PAIR128=1 enables, 2 = at any
  // size).  The idea: a single CTA's UMMA operand reads of an M128 x N128 tile (A 4 KB + B 4 KB per 64-cycle MMA) fill the
  // 128 B/clk shared-memory port (ncu: tensor pipe 49 % = tc_wavefronts_mem_shared 49 %), a pair reads and stages only half
  // of the weight rows per SM.  Measured (profiles/r02s_bench_*.json): SLOWER -- encoder conv 1 0.121 -> 0.147 ms; one
  // resident CTA per SM with a static schedule hides less latency than the two co-resident CTAs of the halo kernel.
  const int pair128_env = [] { const char* e = getenv("ESSB_TC_PAIR128"); return e ? atoi(e) : 0; }();
  const bool pair128 = halo && BN == 128 && d->passes == 2 && pair_env != 0 && pair128_env != 0 && d->ntaps >= 4 &&
                       (pair128_env >= 2 ||     // 2 = always (tests); 1 = only when every CTA pair gets a tile pair
                        (long long)d->N * ((d->OH + 15) / 16) * (((d->OW + 7) / 8 + 1) / 2) >= num_sms() / 2);
  if (pair128) pair = true;
  int halo_occ = 1;
  bool halo_fuse = false;
  int hx0 = 0, hx1 = 0, hy0 = 0, hy1 = 0;
  if (halo) {
    hx0 = hx1 = d->dx[0]; hy0 = hy1 = d->dy[0];
    for (int t = 1; t < d->ntaps; ++t) {
      hx0 = d->dx[t] < hx0 ? d->dx[t] : hx0; hx1 = d->dx[t] > hx1 ? d->dx[t] : hx1;
      hy0 = d->dy[t] < hy0 ? d->dy[t] : hy0; hy1 = d->dy[t] > hy1 ? d->dy[t] : hy1;
    }
    p.halo_w = 8 + (hx1 - hx0);
    p.halo_h = 16 + (hy1 - hy0);
    const int planes = d->passes != 1 ? 2 : 1;
    const int a_plane = (p.halo_w * p.halo_h * 128 + 1023) & ~1023;
    p.a_lo_off = a_plane;
    p.a_stage_bytes = planes * a_plane;
    const int b_plane = (pair ? BN / 2 : BN) * TC_KCH * 2;   // a CTA of a pair stages half of the N rows
    p.b_lo_off = b_plane;
    p.b_tap_bytes = planes * b_plane;
    // Two CTAs per SM (ESSB_TC_OCC=2, default) when the pipeline fits half an SM's shared memory: the narrow-N
    // layers are latency-bound (42 % tensor-pipe utilisation for the first encoder at one CTA/SM, ncu), so a
    // second resident CTA with its own TMEM accumulators fills the stalls; each CTA then allocates only the
    // TMEM columns it needs (2 x BN) instead of all 512.
    static const int occ_env = [] { const char* e = getenv("ESSB_TC_OCC"); return e ? atoi(e) : 2; }();
    const int tail_bytes = 1024 + 512 + bias_floats * (int)sizeof(float);
    int budget = 200 * 1024;
    halo_occ = 1;
    static const int fuse_env = [] { const char* e = getenv("ESSB_TC_FUSEB"); return e ? atoi(e) : 1; }();
    halo_fuse = fuse_env != 0 && d->passes == 3 && 2 * BN <= 256;
    // TMEM per CTA: two accumulator buffers of BN (or 2*BN when fused) columns; two CTAs/SM need <= 256 each
    if (occ_env >= 2 && 2 * BN * (halo_fuse ? 2 : 1) <= 256) {
      const int half = 113 * 1024 - tail_bytes;   // 2 x (113 KB + 1 KB reserved per CTA) = the SM's 228 KB
      if (p.a_stage_bytes + 2 * p.b_tap_bytes <= half) { budget = half; halo_occ = 2; }
    }
    if (wide_halo || pair128) budget = 227 * 1024 - tail_bytes;   // the whole SM: 2 halo tiles + 4 single-plane weight stages
    p.a_stages = (2 * p.a_stage_bytes + 3 * p.b_tap_bytes <= budget) ? 2 : 1;
    // Narrow tiles (N <= 64) are bound by the MMA thread's per-stage cost (barrier wait + fence + commit, ~200
    // cycles) rather than by tensor work (12 MMAs x N/2 cycles per tap): put G taps behind one barrier.
    static const int g_env = [] { const char* e = getenv("ESSB_TC_HALO_G"); return e ? atoi(e) : 0; }();
    int G = 1;
    if (BN <= 64) {
      G = (budget - p.a_stages * p.a_stage_bytes) / (2 * p.b_tap_bytes);
      if (g_env > 0) G = g_env;
      if (G > 8) G = 8;
      if (G > d->ntaps) G = d->ntaps;
      if (G < 1) G = 1;
    }
    p.b_taps_per_stage = G;
    p.b_stage_bytes = G * p.b_tap_bytes;
    p.b_split = 0;
    if (wide_halo) {   // B stage = one plane of one tap (32 KB): finer-grained than [hi | lo] pairs, same bytes in flight
      p.b_split = 1;
      p.b_taps_per_stage = 1;
      p.b_stage_bytes = b_plane;
      p.a_stages = (2 * p.a_stage_bytes + 4 * b_plane <= budget) ? 2 : 1;
    }
    if (pair128) {     // B stage = G taps x [hi | lo] x 64 rows (16 KB per tap): >= 1024 MMA cycles per barrier round trip
      halo_occ = 1;
      halo_fuse = false;
      p.a_stages = 2;
      G = 2;
      p.b_taps_per_stage = G;
      p.b_stage_bytes = G * p.b_tap_bytes;
    }
    int nb = (budget - p.a_stages * p.a_stage_bytes) / p.b_stage_bytes;
    if (nb > MAX_STAGES) nb = MAX_STAGES;
    p.b_stages = nb;
    if (nb < 2 || p.halo_w > 64 || p.halo_h > 64) halo = false;
    if (wide_halo && (nb < 3 || p.a_stages < 2)) halo = false;
    if (pair128 && nb < 3) halo = false;
  }
  pair = pair && halo;
  p.pair = pair ? 1 : 0;
  const int bw_log2 = halo ? 3 : d->bw_log2;
  const int BW = 1 << bw_log2, BH = TC_M >> bw_log2;
  const int boxW = halo ? p.halo_w : BW, boxH = halo ? p.halo_h : BH;
  int rc;
  for (int v = 0; v < d->n_views; ++v) {
    const essb_tc_view& vw = d->views[v];
    ESSB_REQUIRE(vw.hi && (d->passes == 1 || vw.lo), "essb_conv_tc_run: view %d has null planes", v);
    ESSB_REQUIRE(vw.C % TC_KCH == 0, "essb_conv_tc_run: view %d channels %d not a multiple of 64", v, vw.C);
    ESSB_REQUIRE(vw.stride_x % 8 == 0 && vw.stride_y % 8 == 0 && vw.stride_n % 8 == 0 && essb_aligned16(vw.hi) &&
                     essb_aligned16(vw.lo),
                 "essb_conv_tc_run: view %d strides must be multiples of 8 elements and bases 16B aligned", v);
    if ((rc = encode_a_map(&p.tmA_hi[v], vw.hi, vw, d->N, boxW, boxH)) != ESSB_OK) return rc;
    if (d->passes != 1 && (rc = encode_a_map(&p.tmA_lo[v], vw.lo, vw, d->N, boxW, boxH)) != ESSB_OK) return rc;
  }
  const long long ktot = (long long)d->n_w_taps * d->k_per_tap;
  const int b_box_rows = pair ? BN / 2 : BN;
  if ((rc = encode_b_map(&p.tmB_hi, d->w_hi, ktot, d->w_rows, b_box_rows)) != ESSB_OK) return rc;
  if (d->passes != 1 && (rc = encode_b_map(&p.tmB_lo, d->w_lo, ktot, d->w_rows, b_box_rows)) != ESSB_OK) return rc;

  p.tiles_x = (d->OW + BW - 1) / BW;
  p.tiles_y = (d->OH + BH - 1) / BH;
  p.n_tiles = Ngemm / BN;
  p.n_items = d->N * p.tiles_x * p.tiles_y * p.n_tiles;
  p.bw_log2 = bw_log2;
  p.BN = BN;
  p.passes = d->passes;
  const int b_tile_bytes = BN * TC_KCH * 2;
  if (d->passes != 1) {  // stage = [A_hi | A_lo | B_hi | B_lo]
    p.off_alo = A_TILE_BYTES;
    p.off_bhi = 2 * A_TILE_BYTES;
    p.off_blo = 2 * A_TILE_BYTES + b_tile_bytes;
    p.stage_bytes = 2 * (A_TILE_BYTES + b_tile_bytes);
  } else {               // stage = [A_hi | B_hi]
    p.off_alo = 0;
    p.off_bhi = A_TILE_BYTES;
    p.off_blo = 0;
    p.stage_bytes = A_TILE_BYTES + b_tile_bytes;
  }
  p.tx_bytes = p.stage_bytes;
  const int smem_budget = 200 * 1024;
  int stages = smem_budget / p.stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  ESSB_REQUIRE(halo || stages >= 2, "essb_conv_tc_run: tile does not fit two pipeline stages");
  p.stages = stages;
  p.ntaps = d->ntaps;
  p.nseg = d->nseg;
  for (int s = 0; s < 2; ++s) {
    p.seg_chunks[s] = s < d->nseg ? d->seg_C[s] / TC_KCH : 0;
    p.seg_view0[s] = d->seg_view0[s];
    p.seg_koff[s] = d->seg_koff[s];
    ESSB_REQUIRE(s >= d->nseg || (d->seg_C[s] > 0 && d->seg_C[s] % TC_KCH == 0), "essb_conv_tc_run: segment %d channels %d", s, d->seg_C[s]);
  }
  p.k_per_tap = d->k_per_tap;
  for (int t = 0; t < d->ntaps; ++t) {
    ESSB_REQUIRE(d->widx[t] >= 0 && d->widx[t] < d->n_w_taps, "essb_conv_tc_run: widx out of range");
    ESSB_REQUIRE(d->view[t] >= 0, "essb_conv_tc_run: negative view index");
    for (int s = 0; s < d->nseg; ++s)
      ESSB_REQUIRE(d->seg_view0[s] + d->view[t] < d->n_views, "essb_conv_tc_run: view index out of range");
  }
  if (!halo) {
    for (int t = 0; t < d->ntaps; ++t) {
      p.dy[t] = d->dy[t]; p.dx[t] = d->dx[t]; p.view[t] = d->view[t]; p.widx[t] = d->widx[t];
    }
  } else {  // taps sorted by the view they read: one tap group per used view
    p.hx0 = hx0; p.hy0 = hy0;
    int ng = 0, pos = 0;
    bool used[MAX_VIEWS] = {false};
    for (int t0 = 0; t0 < d->ntaps; ++t0) {
      const int v = d->view[t0];
      if (v >= MAX_VIEWS || used[v]) continue;
      used[v] = true;
      p.grp_view[ng] = (int8_t)v;
      p.grp_first[ng] = (int8_t)pos;
      for (int t = t0; t < d->ntaps; ++t)
        if (d->view[t] == v) {
          p.dy[pos] = d->dy[t]; p.dx[pos] = d->dx[t]; p.view[pos] = d->view[t]; p.widx[pos] = d->widx[t];
          ++pos;
        }
      p.grp_count[ng] = (int8_t)(pos - p.grp_first[ng]);
      ++ng;
    }
    p.n_groups = ng;
  }
  p.N = d->N; p.OH = d->OH; p.OW = d->OW; p.Cout = d->Cout;
  p.OHf = d->OHf; p.OWf = d->OWf; p.osy = d->osy; p.ooy = d->ooy; p.osx = d->osx; p.oox = d->oox;
  p.ldo = d->ldo; p.ld_res = d->ld_res; p.ld_planes = d->ld_planes; p.act = d->act;
  p.bias = d->bias; p.res_pre = d->res_pre; p.res_post = d->res_post; p.aux0 = d->aux0; p.aux1 = d->aux1;
  p.out = d->out; p.out2 = d->out2;
  p.out_hi = reinterpret_cast<__nv_bfloat16*>(d->out_hi);
  p.out_lo = reinterpret_cast<__nv_bfloat16*>(d->out_lo);
  ESSB_REQUIRE(!d->bias || d->Cout <= TC_BIAS_SMEM_FLOATS, "essb_conv_tc_run: Cout=%d > %d with a bias", d->Cout,
               TC_BIAS_SMEM_FLOATS);
  {
    auto al32 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 31u) == 0; };
    int wide = 0;
    if (d->epilogue == ESSB_EPI_LINEAR) {
      if (!d->out || (al32(d->out) && d->ldo % 8 == 0)) wide |= TC_WIDE_OUT;
      if (!d->out_hi || (al32(d->out_hi) && al32(d->out_lo) && d->ld_planes % 16 == 0)) wide |= TC_WIDE_PLANES;
      if (al32(d->res_pre) && al32(d->res_post) && d->ld_res % 8 == 0) wide |= TC_WIDE_RES;
    } else if (d->epilogue == ESSB_EPI_LSTM) {
      const int hidden = d->Cout / 4;
      if (al32(d->out) && al32(d->out2) && hidden % 8 == 0) wide |= TC_WIDE_OUT;
      if (al32(d->aux0) && hidden % 8 == 0) wide |= TC_WIDE_AUX;
    }
    p.wide = wide;
  }

  const size_t tail = 1024 /*align slack*/ + 512 /*barriers*/ + (size_t)bias_floats * sizeof(float);
  size_t smem_bytes = (size_t)stages * p.stage_bytes + tail;
  if (halo) smem_bytes = (size_t)p.a_stages * p.a_stage_bytes + (size_t)p.b_stages * p.b_stage_bytes + tail;
  p.tmem_cols = 512;
  p.tmem_buf_stride = 256;
  p.fuse_b = (halo && halo_fuse) ? 1 : 0;
  p.acc_scale = d->acc_scale != 0.f ? d->acc_scale : 1.f;
  p.planes_fmt = d->planes_fmt;
  p.base_offset_mode = baseoff_env;
  ESSB_REQUIRE(d->row_period == 0 || (d->row_period > 0 && d->rows_valid > 0 && d->rows_valid <= d->row_period),
               "essb_conv_tc_run: bad row_period / rows_valid (%d / %d)", d->row_period, d->rows_valid);
  ESSB_REQUIRE(d->phase_cout == 0 || (d->epilogue == ESSB_EPI_LINEAR && d->phase_cout % 32 == 0 && d->Cout == 4 * d->phase_cout),
               "essb_conv_tc_run: phase_cout must be a multiple of 32 with Cout == 4 * phase_cout (LINEAR epilogue)");
  p.phase_cout = d->phase_cout;
  p.row_period = d->row_period;
  p.rows_valid = d->rows_valid;
  if (halo && halo_occ == 2) {
    const int stride = BN * (halo_fuse ? 2 : 1);
    int cols = 32;
    while (cols < 2 * stride) cols <<= 1;
    p.tmem_cols = cols;
    p.tmem_buf_stride = stride;
    if (smem_bytes < 80 * 1024) smem_bytes = 80 * 1024;  // at most two CTAs per SM (2 x tmem_cols <= 512)
  } else if (smem_bytes < 120 * 1024) {
    smem_bytes = 120 * 1024;  // one CTA per SM: each CTA allocates all 512 TMEM columns
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int max_ctas = num_sms() * ((halo && halo_occ == 2) ? 2 : 1);
  int grid = p.n_items < max_ctas ? p.n_items : max_ctas;
  // ---- scheduler: whole tiles first, then the tiles of the partially filled last wave cut into K-slices
  ESSB_REQUIRE(d->sched != nullptr, "essb_conv_tc_run: sched (2 zeroed int32) is required");
  p.sched = d->sched;
  p.n_whole = p.n_items;
  p.split = 1;
  p.splitk_ws = d->splitk_ws;
  p.splitk_cnt = d->splitk_cnt;
  if (!halo && d->splitk_ws && d->splitk_cnt) {
    const int k_iters = d->ntaps * (p.seg_chunks[0] + p.seg_chunks[1]);
    const int rounds = p.n_items / grid, rem = p.n_items % grid;
    if (rounds >= 1 && rem > 0 && rem * 8 <= 148 * 8) {
      const long long tile_bytes = (long long)TC_M * BN * (long long)sizeof(float);
      int best = 1;
      double best_cost = 1.0 - 0.08;   // a split must shorten the tail wave by at least 8 %
      for (int S = 2; S <= 8 && S * 4 <= k_iters; ++S) {
        if ((long long)rem * S * tile_bytes > d->splitk_ws_bytes) break;
        const double cost = (double)((rem * S + grid - 1) / grid) / S;
        if (cost < best_cost - 1e-9) { best_cost = cost; best = S; }
      }
      if (best > 1) {
        p.n_whole = p.n_items - rem;
        p.split = best;
      }
    }
  }
  p.n_units = p.n_whole + (p.n_items - p.n_whole) * p.split;
  cudaError_t e;
  if (pair) {
    const int n_pair_items = d->N * p.tiles_y * ((p.tiles_x + 1) / 2) * p.n_tiles;
    int pairs = num_sms() / 2;
    if (pairs > n_pair_items) pairs = n_pair_items;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * pairs));
    cfg.blockDim = dim3(HALO_THREADS);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
#define ESSB_LAUNCH_PAIR(EPI)                                                                                       \
  e = cudaFuncSetAttribute(conv_tc_pair_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes); \
  if (e == cudaSuccess) e = cudaLaunchKernelEx(&cfg, conv_tc_pair_kernel<EPI>, p);
    if (d->epilogue == ESSB_EPI_LSTM) { ESSB_LAUNCH_PAIR(ESSB_EPI_LSTM) }
    else if (d->epilogue == ESSB_EPI_GRU_UR) { ESSB_LAUNCH_PAIR(ESSB_EPI_GRU_UR) }
    else if (d->epilogue == ESSB_EPI_GRU_OUT) { ESSB_LAUNCH_PAIR(ESSB_EPI_GRU_OUT) }
    else { ESSB_LAUNCH_PAIR(ESSB_EPI_LINEAR) }
#undef ESSB_LAUNCH_PAIR
    if (e != cudaSuccess) {
      essb_set_error("essb_conv_tc_run: CTA-pair launch failed: %s", cudaGetErrorString(e));
      return ESSB_ERR_LAUNCH;
    }
    ESSB_LAUNCH_CHECK("essb_conv_tc_run (pair)");
    return ESSB_OK;
  }
  if (halo) {
#define ESSB_LAUNCH_HALO(EPI)                                                                                            \
  if (halo_occ == 2) {                                                                                                   \
    e = cudaFuncSetAttribute(conv_tc_halo_kernel<EPI, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes); \
    if (e == cudaSuccess) conv_tc_halo_kernel<EPI, 2><<<grid, HALO_THREADS, smem_bytes, st>>>(p);                        \
  } else {                                                                                                               \
    e = cudaFuncSetAttribute(conv_tc_halo_kernel<EPI, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes); \
    if (e == cudaSuccess) conv_tc_halo_kernel<EPI, 1><<<grid, HALO_THREADS, smem_bytes, st>>>(p);                        \
  }
    if (d->epilogue == ESSB_EPI_LSTM) { ESSB_LAUNCH_HALO(ESSB_EPI_LSTM) }
    else if (d->epilogue == ESSB_EPI_GRU_UR) { ESSB_LAUNCH_HALO(ESSB_EPI_GRU_UR) }
    else if (d->epilogue == ESSB_EPI_GRU_OUT) { ESSB_LAUNCH_HALO(ESSB_EPI_GRU_OUT) }
    else { ESSB_LAUNCH_HALO(ESSB_EPI_LINEAR) }
#undef ESSB_LAUNCH_HALO
    if (e != cudaSuccess) {
      essb_set_error("essb_conv_tc_run: cudaFuncSetAttribute (halo) failed: %s", cudaGetErrorString(e));
      return ESSB_ERR_LAUNCH;
    }
    ESSB_LAUNCH_CHECK("essb_conv_tc_run (halo)");
    return ESSB_OK;
  }
  if (d->epilogue == ESSB_EPI_LSTM) {
    e = cudaFuncSetAttribute(conv_tc_kernel<ESSB_EPI_LSTM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e == cudaSuccess) conv_tc_kernel<ESSB_EPI_LSTM><<<grid, TC_THREADS, smem_bytes, st>>>(p);
  } else if (d->epilogue == ESSB_EPI_GRU_UR) {
    e = cudaFuncSetAttribute(conv_tc_kernel<ESSB_EPI_GRU_UR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e == cudaSuccess) conv_tc_kernel<ESSB_EPI_GRU_UR><<<grid, TC_THREADS, smem_bytes, st>>>(p);
  } else if (d->epilogue == ESSB_EPI_GRU_OUT) {
    e = cudaFuncSetAttribute(conv_tc_kernel<ESSB_EPI_GRU_OUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e == cudaSuccess) conv_tc_kernel<ESSB_EPI_GRU_OUT><<<grid, TC_THREADS, smem_bytes, st>>>(p);
  } else {
    e = cudaFuncSetAttribute(conv_tc_kernel<ESSB_EPI_LINEAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e == cudaSuccess) conv_tc_kernel<ESSB_EPI_LINEAR><<<grid, TC_THREADS, smem_bytes, st>>>(p);
  }
  if (e != cudaSuccess) {
    essb_set_error("essb_conv_tc_run: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    return ESSB_ERR_LAUNCH;
  }
  ESSB_LAUNCH_CHECK("essb_conv_tc_run");
  return ESSB_OK;
}
