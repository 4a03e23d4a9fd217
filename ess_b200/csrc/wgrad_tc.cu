// tcgen05 / TMA weight-gradient of a stride-1 gather-convolution (autograd of nn.Conv2d w.r.t. weight):
//     dW[co][ci][t] = sum_{n,y,x} A[n, y+dy_t, x+dx_t, ci] * dY[n, y, x, co]
// GEMM per tap: D[ci (M=128), co (N<=256)] += A_t^T[ci][pixel] * dY[pixel][co], K = pixels.
// Both operands are "MN-major" for this GEMM (channels contiguous, pixels strided), which is exactly
// how a 4-D TMA box [64 pixels][64 channels] of the NHWC bf16 planes lands in shared memory with
// SWIZZLE_128B: rows of 128 B (64 channels) stacked along K.  The UMMA descriptors are therefore
// MN-major / SWIZZLE_128B: SBO = 1024 B (8-pixel K atom), LBO = 8192 B (next 64-channel block).
// The tap shift of A is, as in conv_tc.cu, just a shifted TMA box with hardware zero fill.
// Work item = (split-K range of 64-pixel patches, tap, 128-wide ci tile); partial results go to a
// workspace [split][tap][ci][co] and are reduced in fixed order (deterministic).
// bf16x3: D += A_lo*G_hi + A_hi*G_lo + A_hi*G_hi, fp32 accumulation in TMEM (for N <= 128 the G planes are fused:
// A_hi x [G_hi | G_lo] is one MMA of 2N columns).  Cin <= 64: the two 64-row halves of the M tile are two TAPS.
// Stride-2 convolutions read A through its four parity planes.  wgrad_tc_halo_kernel (below) is the all-taps
// variant for the high-resolution Cin = 64 / 128, Cout <= 64 layers.
#include <cuda.h>
#include <stdlib.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int WG_THREADS = 320;
constexpr int WG_PIX = 64;                 // pixels (K) per pipeline stage
constexpr int WG_BOX_BYTES = WG_PIX * 128; // one [64 px][64 ch] bf16 box = 8 KB
constexpr int WG_MAX_STAGES = 8;

struct WgTcParams {
  CUtensorMap tmA_hi[4], tmA_lo[4], tmG_hi, tmG_lo;   // A: one map per input parity plane (stride 2) or map 0 only
  int n_items, m_tiles, splits, ntaps;
  int patches_total, patches_per_split, tiles_x, tiles_y;
  int bw_log2;           // patch = BW x (64/BW) pixels
  int nb;                // 64-channel boxes of dY (N = 64*nb)
  int stages, passes, stage_bytes;
  int off_alo, off_ghi, off_glo;
  int fuse_g;            // bf16x3 and 2*BN <= 256: A_hi x [G_hi | G_lo] is one MMA of N = 2*BN (upper half added by the epilogue)
  int pair;              // Cin <= 64: the two 64-row halves of the M tile are two TAPS (tap 2*i, 2*i+1) of the same channels
  int cin, cout;
  float* partial;
  int8_t dy[ESSB_MAX_TAPS], dx[ESSB_MAX_TAPS], view[ESSB_MAX_TAPS];   // per tap: shift inside its A view, view index
};

// MN-major SWIZZLE_128B descriptor: [0,14) start>>4, [16,30) LBO>>4 (stride between 64-element blocks
// along M/N), [32,46) SBO>>4 (stride between 8-row groups along K), version 1, layout SWIZZLE_128B.
__device__ __forceinline__ UDesc make_smem_desc_mn(uint32_t saddr) {
  return make_udesc(saddr, (uint32_t)(WG_BOX_BYTES >> 4), 1024u);
}
__device__ __forceinline__ uint32_t make_idesc_mn(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_tc_kernel(const __grid_constant__ WgTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * p.stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + WG_MAX_STAGES;
  uint64_t* tfull_bar = bars + 2 * WG_MAX_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  // the shuffle tells ptxas the warp index is warp-uniform, so the role branches below are uniform branches
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int BW = 1 << p.bw_log2, BH = WG_PIX >> p.bw_log2;
  const int BN = 64 * p.nb;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int tap_items = p.pair ? (p.ntaps + 1) / 2 : p.ntaps;   // tap (or tap-pair) items per split
  const int per_split = tap_items * p.m_tiles;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const int split = item / per_split;
        const int r = item - split * per_split;
        const int tap = r / p.m_tiles, mt = r - tap * p.m_tiles;
        const int p0 = split * p.patches_per_split;
        const int p1 = min(p0 + p.patches_per_split, p.patches_total);
        for (int pp = p0; pp < p1; ++pp) {
          int q = pp;
          const int txi = q % p.tiles_x; q /= p.tiles_x;
          const int tyi = q % p.tiles_y;
          const int n = q / p.tiles_y;
          const int x0 = txi * BW, y0 = tyi * BH;
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* st = smem + (size_t)s * p.stage_bytes;
          mbar_expect_tx(&full_bar[s], (uint32_t)p.stage_bytes);
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            // pair mode: `tap` counts tap pairs; half j of the M tile is tap 2*tap + j (the odd tap out re-reads
            // the last tap, its rows are dropped by the epilogue)
            const int tj = p.pair ? min(2 * tap + j, p.ntaps - 1) : tap;
            const int cj = p.pair ? 0 : mt * 128 + j * 64;
            const int ax = x0 + p.dx[tj], ay = y0 + p.dy[tj];
            tma_load_4d(st + j * WG_BOX_BYTES, &p.tmA_hi[p.view[tj]], &full_bar[s], cj, ax, ay, n);
            if (p.passes == 3)
              tma_load_4d(st + p.off_alo + j * WG_BOX_BYTES, &p.tmA_lo[p.view[tj]], &full_bar[s], cj, ax, ay, n);
          }
          for (int j = 0; j < p.nb; ++j) {
            tma_load_4d(st + p.off_ghi + j * WG_BOX_BYTES, &p.tmG_hi, &full_bar[s], j * 64, x0, y0, n);
            if (p.passes == 3)
              tma_load_4d(st + p.off_glo + j * WG_BOX_BYTES, &p.tmG_lo, &full_bar[s], j * 64, x0, y0, n);
          }
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    {  // whole warp, convergent: umma_bf16 / umma_commit elect the issuing lane themselves
      const uint32_t idesc = make_idesc_mn(128, BN);
      const uint32_t idesc2 = make_idesc_mn(128, 2 * BN);
      int s = 0;
      uint32_t ph = 0;
      uint32_t tph[2] = {0, 0};
      int local = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++local) {
        const int split = item / per_split;
        const int p0 = split * p.patches_per_split;
        const int p1 = min(p0 + p.patches_per_split, p.patches_total);
        const int buf = local & 1;
        mbar_wait(&tempty_bar[buf], tph[buf] ^ 1);
        tph[buf] ^= 1;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * 256);
        for (int pp = p0; pp < p1; ++pp) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)s * p.stage_bytes);
          const UDesc a_hi = make_smem_desc_mn(sa), g_hi = make_smem_desc_mn(sa + p.off_ghi);
          const UDesc a_lo = make_smem_desc_mn(sa + p.off_alo), g_lo = make_smem_desc_mn(sa + p.off_glo);
#pragma unroll
          for (int k = 0; k < WG_PIX / 16; ++k) {
            const uint32_t ko = (uint32_t)(k * (16 * 128 >> 4));  // 16 pixel rows of 128 B
            const uint32_t first = (pp != p0 || k != 0) ? 1u : 0u;
            if (p.passes == 3 && p.fuse_g) {
              // G_lo's 64-channel boxes follow G_hi's in the stage, i.e. [G_hi | G_lo] is one MN-major operand of
              // 2*BN columns: two MMAs and two A reads per K-step instead of three (same trick as conv_tc.cu)
              umma_bf16(d_tmem, a_hi + ko, g_hi + ko, idesc2, first);
              umma_bf16(d_tmem, a_lo + ko, g_hi + ko, idesc, 1u);
            } else if (p.passes == 3) {
              umma_bf16(d_tmem, a_lo + ko, g_hi + ko, idesc, first);
              umma_bf16(d_tmem, a_hi + ko, g_lo + ko, idesc, 1u);
              umma_bf16(d_tmem, a_hi + ko, g_hi + ko, idesc, 1u);
            } else {
              umma_bf16(d_tmem, a_hi + ko, g_hi + ko, idesc, first);
            }
          }
          umma_commit(&empty_bar[s]);
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
        umma_commit(&tfull_bar[buf]);
      }
    }
  } else {
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    uint32_t tph[2] = {0, 0};
    int local = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++local) {
      const int split = item / per_split;
      const int r = item - split * per_split;
      const int tap = r / p.m_tiles, mt = r - tap * p.m_tiles;
      const int buf = local & 1;
      const int row = q * 32 + lane;
      const int ci = p.pair ? (row & 63) : mt * 128 + row;
      const int otap = p.pair ? 2 * tap + (row >> 6) : tap;
      mbar_wait(&tfull_bar[buf], tph[buf]);
      tph[buf] ^= 1;
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 256);
      float* dst = p.partial + (((size_t)split * p.ntaps + min(otap, p.ntaps - 1)) * p.cin + ci) * p.cout;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c0 = half * 32 + j * 64;
        if (c0 >= BN) break;
        uint32_t rr[32];
        __syncwarp();
        tmem_ld32(t_addr + (uint32_t)c0, rr);
        if (p.fuse_g) {
          uint32_t r2[32];
          tmem_ld32(t_addr + (uint32_t)(BN + c0), r2);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) rr[e] = __float_as_uint(__uint_as_float(rr[e]) + __uint_as_float(r2[e]));
        } else {
          tmem_ld_wait();
        }
        if (ci < p.cin && otap < p.ntaps && c0 < p.cout) {   // cout is a multiple of 32
#pragma unroll
          for (int e = 0; e < 32; e += 4)
            *reinterpret_cast<float4*>(dst + c0 + e) =
                make_float4(__uint_as_float(rr[e]), __uint_as_float(rr[e + 1]), __uint_as_float(rr[e + 2]),
                            __uint_as_float(rr[e + 3]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------- all-taps (halo) variant
// For Cin = 64, Cout <= 64 (the high-resolution decoder layers) the kernel above is bound by operand traffic: every
// tap pair re-loads a shifted A box and the same dY box (48 KB per 384 MMA cycles).  Here a CTA owns a contiguous
// range of 8x8-pixel patches and, per patch, loads ONE A tile (patch + halo, e.g. 10x10 pixels) and ONE dY tile and
// issues the MMAs of ALL taps from them: tap (dy, dx) is a different start row inside the halo tile (rows are whole
// 128-byte pixels, so the address-based 128 B swizzle stays consistent, as in conv_tc_halo_kernel), an 8-pixel
// K atom is one patch row, and the stride between K atoms (SBO) is one halo row.  The two 64-row halves of the
// M = 128 tile are two taps (LBO = distance between their start rows).  All (ntaps + 1) / 2 accumulators stay in
// TMEM for the CTA's whole pixel range; one epilogue per CTA writes the split-K partials.
struct WgHaloParams {
  CUtensorMap tmA_hi, tmA_lo, tmG_hi, tmG_lo;
  int splits, patches_total, patches_per_split, tiles_x, tiles_y;
  int ntaps, npairs, passes, stages, stage_bytes, a_plane_bytes, a_tx_bytes;
  int a_blocks;          // 64-channel blocks of A per stage: 1 = Cin 64 (M halves = two taps), 2 = Cin 128 (M halves = the
                         // two channel blocks of ONE tap; npairs then counts taps and tap0 is the first tap of this launch)
  int tap0;
  int halo_w, hx0, hy0;
  int cin, cout;
  float* partial;
  int off[ESSB_MAX_TAPS];   // byte offset of tap t's first pixel row inside the halo tile
};

__device__ __forceinline__ UDesc make_smem_desc_mn_halo(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return make_udesc(saddr, lbo_bytes >> 4, sbo_bytes);
}

__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_tc_halo_kernel(const __grid_constant__ WgHaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * p.stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + WG_MAX_STAGES;
  uint64_t* tfull_bar = bars + 2 * WG_MAX_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 1);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int off_alo = p.a_blocks * p.a_plane_bytes, off_ghi = 2 * p.a_blocks * p.a_plane_bytes;
  const int off_glo = off_ghi + WG_BOX_BYTES;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int split = blockIdx.x;
  const int p0 = split * p.patches_per_split;
  const int p1 = min(p0 + p.patches_per_split, p.patches_total);

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      const uint32_t tx = (uint32_t)(p.passes == 3 ? 2 : 1) * (uint32_t)(p.a_blocks * p.a_tx_bytes + WG_BOX_BYTES);
      for (int pp = p0; pp < p1; ++pp) {
        int q = pp;
        const int txi = q % p.tiles_x; q /= p.tiles_x;
        const int tyi = q % p.tiles_y;
        const int n = q / p.tiles_y;
        const int x0 = txi * 8, y0 = tyi * 8;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* st = smem + (size_t)s * p.stage_bytes;
        mbar_expect_tx(&full_bar[s], tx);
        for (int b = 0; b < p.a_blocks; ++b)
          tma_load_4d(st + b * p.a_plane_bytes, &p.tmA_hi, &full_bar[s], b * 64, x0 + p.hx0, y0 + p.hy0, n);
        tma_load_4d(st + off_ghi, &p.tmG_hi, &full_bar[s], 0, x0, y0, n);
        if (p.passes == 3) {
          for (int b = 0; b < p.a_blocks; ++b)
            tma_load_4d(st + off_alo + b * p.a_plane_bytes, &p.tmA_lo, &full_bar[s], b * 64, x0 + p.hx0, y0 + p.hy0, n);
          tma_load_4d(st + off_glo, &p.tmG_lo, &full_bar[s], 0, x0, y0, n);
        }
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_mn(128, 64);
    const uint32_t sbo = (uint32_t)p.halo_w * 128u;          // next 8-pixel K atom = next patch row = one halo row
    const uint32_t kstep = (2u * sbo) >> 4;                   // a K step of 16 pixels = two patch rows
    int s = 0;
    uint32_t ph = 0;
    for (int pp = p0; pp < p1; ++pp) {
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      const uint32_t sa = smem_u32(smem + (size_t)s * p.stage_bytes);
      const UDesc g_hi = make_smem_desc_mn(sa + off_ghi), g_lo = make_smem_desc_mn(sa + off_glo);
      for (int tp = 0; tp < p.npairs; ++tp) {
        // pair mode: M halves = taps 2tp, 2tp+1 (LBO = distance of their start rows); full mode: M halves = the two
        // channel blocks of tap tap0 + tp (LBO = distance of the two block tiles)
        const int t0 = p.a_blocks == 1 ? 2 * tp : p.tap0 + tp;
        const int t1 = p.a_blocks == 1 ? min(2 * tp + 1, p.ntaps - 1) : t0;
        const uint32_t lbo = p.a_blocks == 1 ? (uint32_t)(p.off[t1] - p.off[t0]) : (uint32_t)p.a_plane_bytes;
        const UDesc a_hi = make_smem_desc_mn_halo(sa + (uint32_t)p.off[t0], lbo, sbo);
        const UDesc a_lo = make_smem_desc_mn_halo(sa + off_alo + (uint32_t)p.off[t0], lbo, sbo);
        const uint32_t d_tmem = tmem_base + (uint32_t)(tp * 64);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t ka = (uint32_t)k * kstep;
          const uint32_t kg = (uint32_t)(k * (16 * 128 >> 4));
          const uint32_t first = (pp != p0 || k != 0) ? 1u : 0u;
          if (p.passes == 3) {
            umma_bf16(d_tmem, a_lo + ka, g_hi + kg, idesc, first);
            umma_bf16(d_tmem, a_hi + ka, g_lo + kg, idesc, 1u);
            umma_bf16(d_tmem, a_hi + ka, g_hi + kg, idesc, 1u);
          } else {
            umma_bf16(d_tmem, a_hi + ka, g_hi + kg, idesc, first);
          }
        }
      }
      umma_commit(&empty_bar[s]);
      if (++s == p.stages) { s = 0; ph ^= 1; }
    }
    umma_commit(tfull_bar);
  } else {
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
    if (p1 > p0) {
      for (int tp = 0; tp < p.npairs; ++tp) {
        const int tap = p.a_blocks == 1 ? 2 * tp + (row >> 6) : p.tap0 + tp;
        const int ci = p.a_blocks == 1 ? (row & 63) : row;
        uint32_t rr[32];
        __syncwarp();
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(tp * 64 + half * 32), rr);
        tmem_ld_wait();
        const int c0 = half * 32;
        if (tap < p.ntaps && ci < p.cin && c0 < p.cout) {
          float* dst = p.partial + (((size_t)split * p.ntaps + tap) * p.cin + ci) * p.cout + c0;
#pragma unroll
          for (int e = 0; e < 32; e += 4)
            *reinterpret_cast<float4*>(dst + e) = make_float4(__uint_as_float(rr[e]), __uint_as_float(rr[e + 1]),
                                                              __uint_as_float(rr[e + 2]), __uint_as_float(rr[e + 3]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// dw[co][ci][tap] = sum_split partial[split][tap][ci][co]
__global__ void wgrad_tc_reduce_kernel(const float* __restrict__ part, float* __restrict__ dw, int splits, int ntaps,
                                       int cin, int cout) {
  const long long total = (long long)ntaps * cin * cout;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int co = (int)(idx % cout);
  const long long r = idx / cout;
  const int ci = (int)(r % cin);
  const int tap = (int)(r / cin);
  float s = 0.f;
  for (int sp = 0; sp < splits; ++sp) s += part[(size_t)sp * total + idx];
  dw[((size_t)co * cin + ci) * ntaps + tap] = s;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled wg_get_encode() {
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

int wg_encode_view(CUtensorMap* tm, const void* base, int C, long long sx, long long sy, long long sn, int N, int H,
                   int W, int BW, int BH) {
  PFN_encodeTiled enc = wg_get_encode();
  if (!enc) {
    essb_set_error("wgrad_tc: cuTensorMapEncodeTiled unavailable");
    return ESSB_ERR_DRIVER;
  }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)sx * 2, (cuuint64_t)sy * 2, (cuuint64_t)sn * 2};
  cuuint32_t box[4] = {64u, (cuuint32_t)BW, (cuuint32_t)BH, 1u};
  cuuint32_t es[4] = {1u, 1u, 1u, 1u};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    essb_set_error("wgrad_tc: cuTensorMapEncodeTiled (view) failed with %d", (int)r);
    return ESSB_ERR_DRIVER;
  }
  return ESSB_OK;
}

int wg_encode(CUtensorMap* tm, const void* base, int C, int ld, int N, int H, int W, int BW, int BH) {
  PFN_encodeTiled enc = wg_get_encode();
  if (!enc) {
    essb_set_error("wgrad_tc: cuTensorMapEncodeTiled unavailable");
    return ESSB_ERR_DRIVER;
  }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2};
  cuuint32_t box[4] = {64u, (cuuint32_t)BW, (cuuint32_t)BH, 1u};
  cuuint32_t es[4] = {1u, 1u, 1u, 1u};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    essb_set_error("wgrad_tc: cuTensorMapEncodeTiled failed with %d", (int)r);
    return ESSB_ERR_DRIVER;
  }
  return ESSB_OK;
}

struct WgPlan {
  int m_tiles, splits, patches_total, patches_per_split, tiles_x, tiles_y, bw_log2, nb, pair;
};

int wg_plan(const essb_wgrad_tc& d, WgPlan* pl) {
  if (d.Cin <= 0 || d.Cout <= 0 || d.Cout % 32 != 0 || d.Cout > 256 || d.g_ld % 64 != 0 || d.g_ld < d.Cout) return -1;
  pl->m_tiles = (d.Cin + 127) / 128;
  pl->pair = (d.Cin <= 64 && d.ntaps >= 2) ? 1 : 0;
  pl->nb = (d.Cout + 63) / 64;
  int best = 4;
  double best_waste = 1e30;
  for (int b = 2; b <= 6; ++b) {
    const int bw = 1 << b, bh = WG_PIX >> b;
    const double waste = (double)((d.W + bw - 1) / bw * bw) * ((d.H + bh - 1) / bh * bh) / ((double)d.W * d.H);
    if (waste < best_waste - 1e-9) { best_waste = waste; best = b; }
  }
  pl->bw_log2 = best;
  const int BW = 1 << best, BH = WG_PIX >> best;
  pl->tiles_x = (d.W + BW - 1) / BW;
  pl->tiles_y = (d.H + BH - 1) / BH;
  pl->patches_total = d.N * pl->tiles_x * pl->tiles_y;
  const int per_split = (pl->pair ? (d.ntaps + 1) / 2 : d.ntaps) * pl->m_tiles;
  // two full waves of work items on the 148 SMs, never a third, mostly empty one: items = splits * per_split must
  // not exceed 2 * 148 (rounding UP here gave 306 items = 3 waves at 69 % occupancy for the 256 -> 256 layers)
  int splits = (2 * 148) / per_split;
  const int max_splits = (pl->patches_total + 3) / 4;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  pl->patches_per_split = (pl->patches_total + splits - 1) / splits;
  pl->splits = (pl->patches_total + pl->patches_per_split - 1) / pl->patches_per_split;
  return 0;
}

}  // namespace

// all-taps variant: which layers, and how the patches are split over the CTAs
static bool wg_halo_ok(const essb_wgrad_tc& d) {
  static const int env = [] { const char* e = getenv("ESSB_WGRAD_HALO"); return e ? atoi(e) : 1; }();
  static const int env128 = [] { const char* e = getenv("ESSB_WGRAD_HALO_128"); return e ? atoi(e) : 1; }();
  const bool cin_ok = d.Cin == 64 || (d.Cin == 128 && env128 != 0);
  if (!env || d.a_stride == 2 || !cin_ok || d.Cout > 64 || d.Cout % 32 != 0 || d.g_ld != 64 || d.ntaps < 2 ||
      d.ntaps > (d.Cin == 64 ? 10 : 16))
    return false;
  int x0 = d.dx[0], x1 = d.dx[0], y0 = d.dy[0], y1 = d.dy[0];
  for (int t = 1; t < d.ntaps; ++t) {
    x0 = d.dx[t] < x0 ? d.dx[t] : x0; x1 = d.dx[t] > x1 ? d.dx[t] : x1;
    y0 = d.dy[t] < y0 ? d.dy[t] : y0; y1 = d.dy[t] > y1 ? d.dy[t] : y1;
    // tap pairs (2i, 2i+1) need non-decreasing start rows inside the halo tile (LBO is unsigned)
    if ((t & 1) && (d.dy[t] < d.dy[t - 1] || (d.dy[t] == d.dy[t - 1] && d.dx[t] < d.dx[t - 1]))) return false;
  }
  return (x1 - x0) <= 8 && (y1 - y0) <= 8;
}
static void wg_halo_split(const essb_wgrad_tc& d, int* splits, int* per_split, int* total, int* tiles_x, int* tiles_y) {
  *tiles_x = (d.W + 7) / 8;
  *tiles_y = (d.H + 7) / 8;
  *total = d.N * *tiles_x * *tiles_y;
  int sp = *total < 148 ? *total : 148;
  *per_split = (*total + sp - 1) / sp;
  *splits = (*total + *per_split - 1) / *per_split;
}

extern "C" int64_t essb_wgrad_tc_workspace_bytes(const essb_wgrad_tc* d) {
  WgPlan pl;
  if (!d || wg_plan(*d, &pl) != 0) return -1;
  int64_t splits = pl.splits;
  if (wg_halo_ok(*d)) {
    int sp, per, total, tx, ty;
    wg_halo_split(*d, &sp, &per, &total, &tx, &ty);
    if (sp > splits) splits = sp;
  }
  return splits * d->ntaps * d->Cin * d->Cout * (int64_t)sizeof(float);
}

extern "C" int essb_wgrad_tc_run(const essb_wgrad_tc* d, void* stream) {
  ESSB_REQUIRE(d != nullptr, "essb_wgrad_tc_run: null descriptor");
  ESSB_REQUIRE(d->a_hi && d->g_hi && d->dw && d->workspace, "essb_wgrad_tc_run: null tensor");
  ESSB_REQUIRE(d->passes == 1 || (d->passes == 3 && d->a_lo && d->g_lo), "essb_wgrad_tc_run: passes must be 1 or 3 (with lo planes)");
  ESSB_REQUIRE(d->ntaps >= 1 && d->ntaps <= ESSB_MAX_TAPS, "essb_wgrad_tc_run: ntaps=%d", d->ntaps);
  ESSB_REQUIRE(d->a_ld % 8 == 0 && d->g_ld % 8 == 0 && d->a_ld >= d->Cin, "essb_wgrad_tc_run: plane pitches must be multiples of 8");
  ESSB_REQUIRE(essb_aligned16(d->a_hi) && essb_aligned16(d->a_lo) && essb_aligned16(d->g_hi) && essb_aligned16(d->g_lo) &&
                   essb_aligned16(d->workspace),
               "essb_wgrad_tc_run: pointers must be 16B aligned");
  WgPlan pl;
  ESSB_REQUIRE(wg_plan(*d, &pl) == 0, "essb_wgrad_tc_run: unsupported shape (Cout %% 32 == 0, Cout <= 256, g_ld %% 64 == 0 required)");
  const int64_t need = (int64_t)pl.splits * d->ntaps * d->Cin * d->Cout * (int64_t)sizeof(float);
  if (d->workspace_bytes < need) {
    essb_set_error("essb_wgrad_tc_run: workspace %lld < %lld bytes", (long long)d->workspace_bytes, (long long)need);
    return ESSB_ERR_WORKSPACE;
  }
  if (wg_halo_ok(*d)) {
    static thread_local WgHaloParams h;
    int hx0 = d->dx[0], hx1 = d->dx[0], hy0 = d->dy[0], hy1 = d->dy[0];
    for (int t = 1; t < d->ntaps; ++t) {
      hx0 = d->dx[t] < hx0 ? d->dx[t] : hx0; hx1 = d->dx[t] > hx1 ? d->dx[t] : hx1;
      hy0 = d->dy[t] < hy0 ? d->dy[t] : hy0; hy1 = d->dy[t] > hy1 ? d->dy[t] : hy1;
    }
    const int halo_w = 8 + (hx1 - hx0), halo_h = 8 + (hy1 - hy0);
    int rc2;
    if ((rc2 = wg_encode(&h.tmA_hi, d->a_hi, d->Cin, d->a_ld, d->N, d->H, d->W, halo_w, halo_h)) != ESSB_OK) return rc2;
    if ((rc2 = wg_encode(&h.tmG_hi, d->g_hi, d->g_ld, d->g_ld, d->N, d->H, d->W, 8, 8)) != ESSB_OK) return rc2;
    if (d->passes == 3) {
      if ((rc2 = wg_encode(&h.tmA_lo, d->a_lo, d->Cin, d->a_ld, d->N, d->H, d->W, halo_w, halo_h)) != ESSB_OK) return rc2;
      if ((rc2 = wg_encode(&h.tmG_lo, d->g_lo, d->g_ld, d->g_ld, d->N, d->H, d->W, 8, 8)) != ESSB_OK) return rc2;
    }
    wg_halo_split(*d, &h.splits, &h.patches_per_split, &h.patches_total, &h.tiles_x, &h.tiles_y);
    h.a_blocks = d->Cin == 64 ? 1 : 2;
    h.ntaps = d->ntaps; h.passes = d->passes;
    h.halo_w = halo_w; h.hx0 = hx0; h.hy0 = hy0;
    h.a_tx_bytes = halo_w * halo_h * 128;
    h.a_plane_bytes = (h.a_tx_bytes + 1023) & ~1023;
    // stage = [A_hi blocks | A_lo blocks | G_hi | G_lo] (lo halves unused in 1-pass mode)
    h.stage_bytes = 2 * h.a_blocks * h.a_plane_bytes + 2 * WG_BOX_BYTES;
    int stages = (216 * 1024) / h.stage_bytes;
    if (stages > WG_MAX_STAGES) stages = WG_MAX_STAGES;
    ESSB_REQUIRE(stages >= 2, "essb_wgrad_tc_run: halo tile does not fit two stages");
    h.stages = stages;
    h.cin = d->Cin; h.cout = d->Cout; h.partial = d->workspace;
    for (int t = 0; t < d->ntaps; ++t) h.off[t] = ((d->dy[t] - hy0) * halo_w + (d->dx[t] - hx0)) * 128;
    const int64_t need_h = (int64_t)h.splits * d->ntaps * d->Cin * d->Cout * (int64_t)sizeof(float);
    if (d->workspace_bytes < need_h) {
      essb_set_error("essb_wgrad_tc_run: workspace %lld < %lld bytes", (long long)d->workspace_bytes, (long long)need_h);
      return ESSB_ERR_WORKSPACE;
    }
    size_t smem_h = (size_t)stages * h.stage_bytes + 1024 + 256;
    if (smem_h < 120 * 1024) smem_h = 120 * 1024;
    cudaStream_t sth = (cudaStream_t)stream;
    cudaError_t eh = cudaFuncSetAttribute(wgrad_tc_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_h);
    if (eh != cudaSuccess) {
      essb_set_error("essb_wgrad_tc_run: cudaFuncSetAttribute (halo) failed: %s", cudaGetErrorString(eh));
      return ESSB_ERR_LAUNCH;
    }
    if (h.a_blocks == 1) {      // Cin = 64: all taps in one launch, two taps per accumulator
      h.tap0 = 0;
      h.npairs = (d->ntaps + 1) / 2;
      wgrad_tc_halo_kernel<<<h.splits, WG_THREADS, smem_h, sth>>>(h);
      ESSB_LAUNCH_CHECK("essb_wgrad_tc_run (halo)");
    } else {                    // Cin = 128: one tap per accumulator, 8 accumulators (512 TMEM columns) per launch
      const int ngroups = (d->ntaps + 7) / 8, gsize = (d->ntaps + ngroups - 1) / ngroups;   // 9 taps -> 5 + 4
      for (int t0 = 0; t0 < d->ntaps; t0 += gsize) {
        h.tap0 = t0;
        h.npairs = d->ntaps - t0 < gsize ? d->ntaps - t0 : gsize;
        wgrad_tc_halo_kernel<<<h.splits, WG_THREADS, smem_h, sth>>>(h);
        ESSB_LAUNCH_CHECK("essb_wgrad_tc_run (halo)");
      }
    }
    const long long total_h = (long long)d->ntaps * d->Cin * d->Cout;
    wgrad_tc_reduce_kernel<<<(unsigned)((total_h + 255) / 256), 256, 0, sth>>>(d->workspace, d->dw, h.splits, d->ntaps,
                                                                             d->Cin, d->Cout);
    ESSB_LAUNCH_CHECK("essb_wgrad_tc_reduce");
    return ESSB_OK;
  }
  static thread_local WgTcParams p;
  const int BW = 1 << pl.bw_log2, BH = WG_PIX >> pl.bw_log2;
  int rc;
  const int astride = d->a_stride == 2 ? 2 : 1;
  ESSB_REQUIRE(d->a_stride == 0 || d->a_stride == 1 || d->a_stride == 2, "essb_wgrad_tc_run: a_stride must be 1 or 2");
  if (astride == 1) {
    if ((rc = wg_encode(&p.tmA_hi[0], d->a_hi, d->Cin, d->a_ld, d->N, d->H, d->W, BW, BH)) != ESSB_OK) return rc;
    if (d->passes == 3 && (rc = wg_encode(&p.tmA_lo[0], d->a_lo, d->Cin, d->a_ld, d->N, d->H, d->W, BW, BH)) != ESSB_OK)
      return rc;
  } else {
    // stride-2 convolution: the input [N, 2H, 2W, a_ld] is read through its four parity planes (view = 2*py + px;
    // base shifted by (py*2W + px)*a_ld, pixel strides doubled), exactly like the forward kernel's parity views
    const long long ld = d->a_ld, Wi = 2LL * d->W, Hi = 2LL * d->H;
    for (int v = 0; v < 4; ++v) {
      const long long off = ((long long)(v >> 1) * Wi + (v & 1)) * ld;
      if ((rc = wg_encode_view(&p.tmA_hi[v], d->a_hi + off, d->Cin, 2 * ld, 2 * Wi * ld, Hi * Wi * ld, d->N, d->H, d->W, BW,
                               BH)) != ESSB_OK)
        return rc;
      if (d->passes == 3 && (rc = wg_encode_view(&p.tmA_lo[v], d->a_lo + off, d->Cin, 2 * ld, 2 * Wi * ld, Hi * Wi * ld,
                                                 d->N, d->H, d->W, BW, BH)) != ESSB_OK)
        return rc;
    }
  }
  if ((rc = wg_encode(&p.tmG_hi, d->g_hi, d->g_ld, d->g_ld, d->N, d->H, d->W, BW, BH)) != ESSB_OK) return rc;
  if (d->passes == 3 && (rc = wg_encode(&p.tmG_lo, d->g_lo, d->g_ld, d->g_ld, d->N, d->H, d->W, BW, BH)) != ESSB_OK) return rc;
  p.m_tiles = pl.m_tiles; p.splits = pl.splits; p.ntaps = d->ntaps;
  static const int fuse_env = [] { const char* e = getenv("ESSB_TC_FUSEB"); return e ? atoi(e) : 1; }();
  p.fuse_g = (fuse_env != 0 && d->passes == 3 && 2 * 64 * pl.nb <= 256) ? 1 : 0;
  p.pair = pl.pair;
  p.n_items = pl.splits * (pl.pair ? (d->ntaps + 1) / 2 : d->ntaps) * pl.m_tiles;
  p.patches_total = pl.patches_total; p.patches_per_split = pl.patches_per_split;
  p.tiles_x = pl.tiles_x; p.tiles_y = pl.tiles_y; p.bw_log2 = pl.bw_log2; p.nb = pl.nb;
  p.passes = d->passes;
  const int a_bytes = 2 * WG_BOX_BYTES, g_bytes = pl.nb * WG_BOX_BYTES;
  if (d->passes == 3) {
    p.off_alo = a_bytes; p.off_ghi = 2 * a_bytes; p.off_glo = 2 * a_bytes + g_bytes;
    p.stage_bytes = 2 * (a_bytes + g_bytes);
  } else {
    p.off_alo = 0; p.off_ghi = a_bytes; p.off_glo = 0;
    p.stage_bytes = a_bytes + g_bytes;
  }
  int stages = (200 * 1024) / p.stage_bytes;
  if (stages > WG_MAX_STAGES) stages = WG_MAX_STAGES;
  ESSB_REQUIRE(stages >= 2, "essb_wgrad_tc_run: tile does not fit two stages");
  p.stages = stages;
  p.cin = d->Cin; p.cout = d->Cout; p.partial = d->workspace;
  for (int t = 0; t < d->ntaps; ++t) {
    if (astride == 1) {
      p.dy[t] = d->dy[t]; p.dx[t] = d->dx[t]; p.view[t] = 0;
    } else {  // input offset o = 2a + parity (floor division): view = parity plane, shift a inside it
      const int oy = d->dy[t], ox = d->dx[t];
      const int py = oy & 1, px = ox & 1;
      p.dy[t] = (int8_t)((oy - py) / 2); p.dx[t] = (int8_t)((ox - px) / 2); p.view[t] = (int8_t)(py * 2 + px);
    }
  }
  size_t smem_bytes = (size_t)stages * p.stage_bytes + 1024 + 256;
  if (smem_bytes < 120 * 1024) smem_bytes = 120 * 1024;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
  if (e != cudaSuccess) {
    essb_set_error("essb_wgrad_tc_run: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    return ESSB_ERR_LAUNCH;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = p.n_items < sms ? p.n_items : sms;
  wgrad_tc_kernel<<<grid, WG_THREADS, smem_bytes, st>>>(p);
  ESSB_LAUNCH_CHECK("essb_wgrad_tc_run");
  const long long total = (long long)d->ntaps * d->Cin * d->Cout;
  wgrad_tc_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d->workspace, d->dw, pl.splits, d->ntaps, d->Cin,
                                                                          d->Cout);
  ESSB_LAUNCH_CHECK("essb_wgrad_tc_reduce");
  return ESSB_OK;
}
