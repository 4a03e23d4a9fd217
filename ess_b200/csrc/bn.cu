// Train-mode BatchNorm2d forward/backward helpers for the UDA image encoder (StyleEncoderE2VID =
// ResNet-18 stem + layer1-3, models/style_networks.py:110-145; torchvision BasicBlock) and the UDA
// consistency losses (symJSDivLoss, utils/loss_functions.py:27-37; torch.nn.L1Loss).
// Statistics reuse essb_in_stats / essb_in_finalize / essb_partial_reduce with N = 1 (per-channel
// reductions over all N*H*W rows).  All kernels are HBM-bound, float4 vectorised.
#include "common.cuh"

namespace {

constexpr int BN_ROWS_PER_BLOCK = 256;

// out = relu?(x * a[c] + b[c] + res)
__global__ void affine_act_kernel(const float* __restrict__ x, int ld_x, const float* __restrict__ a,
                                  const float* __restrict__ b, const float* __restrict__ res, int ld_res, int relu,
                                  float* __restrict__ out, int ld_out, int C, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int CQ = C >> 2;
  const int c = (int)(idx % CQ) * 4;
  const long long row = idx / CQ;
  float4 v = *reinterpret_cast<const float4*>(x + row * ld_x + c);
  const float4 av = *reinterpret_cast<const float4*>(a + c);
  const float4 bv = *reinterpret_cast<const float4*>(b + c);
  v.x = v.x * av.x + bv.x; v.y = v.y * av.y + bv.y; v.z = v.z * av.z + bv.z; v.w = v.w * av.w + bv.w;
  if (res) {
    const float4 r = *reinterpret_cast<const float4*>(res + row * ld_res + c);
    v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
  }
  if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
  *reinterpret_cast<float4*>(out + row * ld_out + c) = v;
}

// g = dout * (mask > 0);  per-block partial sums of g and g * xhat  -> partial [blocks][C][2]
__global__ void __launch_bounds__(256) bn_bwd_pass1_kernel(const float* __restrict__ dout, int ld_d,
                                                           const float* __restrict__ mask, int ld_m,
                                                           const float* __restrict__ x, int ld_x,
                                                           const float* __restrict__ mean, const float* __restrict__ rstd,
                                                           float* __restrict__ g, float* __restrict__ partial,
                                                           long long rows, int C) {
  __shared__ float red[8][32][8];
  const int blk = blockIdx.x;
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
  float4 m4 = make_float4(0.f, 0.f, 0.f, 0.f), r4 = make_float4(1.f, 1.f, 1.f, 1.f);
  if (cq_ok) {
    m4 = *reinterpret_cast<const float4*>(mean + c);
    r4 = *reinterpret_cast<const float4*>(rstd + c);
  }
  float sg[4] = {0.f, 0.f, 0.f, 0.f}, sgx[4] = {0.f, 0.f, 0.f, 0.f};
  const long long r_begin = (long long)blk * BN_ROWS_PER_BLOCK;
  for (int i = warp * ppw + sub; i < BN_ROWS_PER_BLOCK; i += 8 * ppw) {
    const long long row = r_begin + i;
    if (row >= rows || !cq_ok) continue;
    const float4 d = *reinterpret_cast<const float4*>(dout + row * ld_d + c);
    float av[4] = {d.x, d.y, d.z, d.w};
    if (mask) {
      const float4 mk = *reinterpret_cast<const float4*>(mask + row * ld_m + c);
      if (!(mk.x > 0.f)) av[0] = 0.f;
      if (!(mk.y > 0.f)) av[1] = 0.f;
      if (!(mk.z > 0.f)) av[2] = 0.f;
      if (!(mk.w > 0.f)) av[3] = 0.f;
    }
    const float4 xv = *reinterpret_cast<const float4*>(x + row * ld_x + c);
    const float xh[4] = {(xv.x - m4.x) * r4.x, (xv.y - m4.y) * r4.y, (xv.z - m4.z) * r4.z, (xv.w - m4.w) * r4.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      sg[e] += av[e];
      sgx[e] += av[e] * xh[e];
    }
    *reinterpret_cast<float4*>(g + row * C + c) = make_float4(av[0], av[1], av[2], av[3]);
  }
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
    float* dst = partial + ((size_t)blk * C + c) * 2;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int w8 = 0; w8 < 8; ++w8) { s0 += red[w8][lane][e]; s1 += red[w8][lane][4 + e]; }
      dst[e * 2] = s0;
      dst[e * 2 + 1] = s1;
    }
  }
}

// dx = gamma * rstd * (g - sum(g)/M - xhat * sum(g*xhat)/M)
__global__ void bn_bwd_pass2_kernel(const float* __restrict__ g, const float* __restrict__ x, int ld_x,
                                    const float* __restrict__ mean, const float* __restrict__ rstd,
                                    const float* __restrict__ gamma, const float* __restrict__ totals, float inv_m,
                                    float* __restrict__ dx, int C, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int CQ = C >> 2;
  const int c = (int)(idx % CQ) * 4;
  const long long row = idx / CQ;
  const float4 gv = *reinterpret_cast<const float4*>(g + row * C + c);
  const float4 xv = *reinterpret_cast<const float4*>(x + row * ld_x + c);
  const float4 m = *reinterpret_cast<const float4*>(mean + c);
  const float4 r = *reinterpret_cast<const float4*>(rstd + c);
  const float4 ga = *reinterpret_cast<const float4*>(gamma + c);
  const float* ts = totals + (size_t)c * 2;
  const float gvv[4] = {gv.x, gv.y, gv.z, gv.w}, xa[4] = {xv.x, xv.y, xv.z, xv.w};
  const float ma[4] = {m.x, m.y, m.z, m.w}, ra[4] = {r.x, r.y, r.z, r.w}, gm[4] = {ga.x, ga.y, ga.z, ga.w};
  float o[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float xh = (xa[e] - ma[e]) * ra[e];
    o[e] = gm[e] * ra[e] * (gvv[e] - ts[e * 2] * inv_m - xh * (ts[e * 2 + 1] * inv_m));
  }
  *reinterpret_cast<float4*>(dx + row * C + c) = make_float4(o[0], o[1], o[2], o[3]);
}

// ------------------------------------------------------------------------------ UDA losses
// sums[0] += sum |a - b|
__global__ void l1_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n,
                              double* __restrict__ sums) {
  __shared__ double red[8];
  double s = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    s += (double)fabsf(a[i] - b[i]);
  s = warp_sum_d(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(sums, t);
  }
}
// da = gscale/n * sign(a - b)
__global__ void l1_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n,
                              const float* __restrict__ gscale, float* __restrict__ da) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float d = a[i] - b[i];
  const float s = gscale[0] / (float)n;
  da[i] = d > 0.f ? s : (d < 0.f ? -s : 0.f);
}

// symmetric JS-style divergence of utils/loss_functions.py:27-37 on pixel-major logits [rows][K]:
//   L = 0.5*mean_elem( t*(log t - log p) ) + 0.5*mean_elem( p*(log p - log t) ),
//   p = clamp(softmax(predict), 1e-10), t = clamp(softmax(target), 1e-10); mean over rows*K elements.
template <int KMAX>
__device__ __forceinline__ void softmax_k(const float* __restrict__ q, int K, float (&p)[KMAX]) {
  float mx = -INFINITY;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) { p[k] = (k < K) ? q[k] : -INFINITY; mx = fmaxf(mx, p[k]); }
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) { p[k] = (k < K) ? expf(p[k] - mx) : 0.f; sum += p[k]; }
  const float inv = 1.f / sum;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) p[k] *= inv;
}

template <int KMAX>
__global__ void __launch_bounds__(256) jsdiv_kernel(const float* __restrict__ predict, int ld_p,
                                                    const float* __restrict__ target, int ld_t, long long rows, int K,
                                                    double* __restrict__ sums, const float* __restrict__ gscale,
                                                    float* __restrict__ dpredict, int ld_d) {
  __shared__ double red[8];
  double acc = 0.0;
  const float inv_m = 1.f / ((float)rows * (float)K);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += (long long)gridDim.x * blockDim.x) {
    float q[KMAX], t[KMAX];
    softmax_k<KMAX>(predict + i * ld_p, K, q);
    softmax_k<KMAX>(target + i * ld_t, K, t);
    float gk[KMAX];
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      gk[k] = 0.f;
      if (k < K) {
        const float pc = fmaxf(q[k], 1e-10f), tc = fmaxf(t[k], 1e-10f);
        const float lp = logf(pc), lt = logf(tc);
        acc += (double)(0.5f * (tc * (lt - lp) + pc * (lp - lt)));
        if (dpredict && q[k] > 1e-10f) gk[k] = 0.5f * inv_m * (-tc / pc + (lp - lt + 1.f));
        dot += gk[k] * q[k];
      }
    }
    if (dpredict) {
      const float gs = gscale ? gscale[0] : 1.f;
#pragma unroll
      for (int k = 0; k < KMAX; ++k)
        if (k < K) dpredict[i * ld_d + k] = gs * q[k] * (gk[k] - dot);
    }
  }
  if (sums) {
    acc = warp_sum_d(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tt = 0.0;
      for (int w = 0; w < 8; ++w) tt += red[w];
      atomicAdd(sums, tt);
    }
  }
}

// Train-mode BatchNorm bookkeeping in one launch (nn.BatchNorm2d forward, training=True): folded scale/shift of the
// apply pass, and the running-statistics update with the UNBIASED batch variance (momentum form), num_batches_tracked += 1.
__global__ void bn_train_finalize_kernel(const float* __restrict__ mean, const float* __restrict__ rstd,
                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                         float* __restrict__ a, float* __restrict__ b, float* __restrict__ running_mean,
                                         float* __restrict__ running_var, long long* __restrict__ num_batches, int C,
                                         float momentum, float eps, float unbias) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && num_batches) *num_batches += 1;
  if (c >= C) return;
  const float m = mean[c], r = rstd[c];
  const float g = gamma ? gamma[c] : 1.f;
  const float av = g * r;
  a[c] = av;
  b[c] = (beta ? beta[c] : 0.f) - m * av;
  if (running_mean) {
    const float var_b = 1.f / (r * r) - eps;
    running_mean[c] = running_mean[c] * (1.f - momentum) + m * momentum;
    running_var[c] = running_var[c] * (1.f - momentum) + var_b * unbias * momentum;
  }
}

inline unsigned grid_for(long long n, int per_thread) {
  long long b = (n + 256LL * per_thread - 1) / (256LL * per_thread);
  if (b > 148 * 8) b = 148 * 8;
  if (b < 1) b = 1;
  return (unsigned)b;
}

int vec_ok(const void* p, int ld, const char* who) {
  ESSB_REQUIRE(p == nullptr || (essb_aligned16(p) && ld % 4 == 0), "%s: tensors must be 16B aligned with ld %% 4 == 0", who);
  return ESSB_OK;
}

}  // namespace

extern "C" int essb_affine_act(const float* x, int ld_x, const float* a, const float* b, const float* res, int ld_res,
                               int relu, float* out, int ld_out, int64_t rows, int C, void* stream) {
  ESSB_REQUIRE(x && a && b && out && rows > 0 && C > 0 && C % 4 == 0, "essb_affine_act: bad arguments (C %% 4 == 0)");
  int rc;
  if ((rc = vec_ok(x, ld_x, "essb_affine_act")) || (rc = vec_ok(res, ld_res, "essb_affine_act")) ||
      (rc = vec_ok(out, ld_out, "essb_affine_act")) || (rc = vec_ok(a, 4, "essb_affine_act")) ||
      (rc = vec_ok(b, 4, "essb_affine_act")))
    return rc;
  const long long total = rows * (C / 4);
  affine_act_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, ld_x, a, b, res, ld_res, relu,
                                                                                     out, ld_out, C, total);
  ESSB_LAUNCH_CHECK("essb_affine_act");
  return ESSB_OK;
}

extern "C" int essb_bn_train_finalize(const float* mean, const float* rstd, const float* gamma, const float* beta,
                                      float* a, float* b, float* running_mean, float* running_var,
                                      int64_t* num_batches_tracked, int C, int64_t rows, float momentum, float eps,
                                      void* stream) {
  ESSB_REQUIRE(mean && rstd && a && b && C > 0 && rows > 0, "essb_bn_train_finalize: bad arguments");
  ESSB_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "essb_bn_train_finalize: running_mean / running_var must come together");
  const float unbias = (float)((double)rows / (double)(rows > 1 ? rows - 1 : 1));
  bn_train_finalize_kernel<<<(unsigned)((C + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      mean, rstd, gamma, beta, a, b, running_mean, running_var, reinterpret_cast<long long*>(num_batches_tracked), C,
      momentum, eps, unbias);
  ESSB_LAUNCH_CHECK("essb_bn_train_finalize");
  return ESSB_OK;
}

extern "C" int essb_bn_bwd_blocks(int64_t rows) { return (int)((rows + BN_ROWS_PER_BLOCK - 1) / BN_ROWS_PER_BLOCK); }

extern "C" int essb_bn_bwd_pass1(const float* dout, int ld_d, const float* mask, int ld_m, const float* x, int ld_x,
                                 const float* mean, const float* rstd, float* g, float* partial, int64_t rows, int C,
                                 void* stream) {
  ESSB_REQUIRE(dout && x && mean && rstd && g && partial && rows > 0 && C > 0 && C % 4 == 0,
               "essb_bn_bwd_pass1: bad arguments (C %% 4 == 0)");
  int rc;
  if ((rc = vec_ok(dout, ld_d, "essb_bn_bwd_pass1")) || (rc = vec_ok(mask, ld_m, "essb_bn_bwd_pass1")) ||
      (rc = vec_ok(x, ld_x, "essb_bn_bwd_pass1")) || (rc = vec_ok(g, 4, "essb_bn_bwd_pass1")))
    return rc;
  dim3 grid(essb_bn_bwd_blocks(rows), 1, (C / 4 + 31) / 32);
  bn_bwd_pass1_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dout, ld_d, mask, ld_m, x, ld_x, mean, rstd, g, partial,
                                                              rows, C);
  ESSB_LAUNCH_CHECK("essb_bn_bwd_pass1");
  return ESSB_OK;
}

extern "C" int essb_bn_bwd_pass2(const float* g, const float* x, int ld_x, const float* mean, const float* rstd,
                                 const float* gamma, const float* totals, float* dx, int64_t rows, int C,
                                 void* stream) {
  ESSB_REQUIRE(g && x && mean && rstd && gamma && totals && dx && rows > 0 && C % 4 == 0, "essb_bn_bwd_pass2: bad arguments");
  int rc;
  if ((rc = vec_ok(x, ld_x, "essb_bn_bwd_pass2")) || (rc = vec_ok(g, 4, "essb_bn_bwd_pass2")) ||
      (rc = vec_ok(dx, 4, "essb_bn_bwd_pass2")))
    return rc;
  const long long total = rows * (C / 4);
  bn_bwd_pass2_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      g, x, ld_x, mean, rstd, gamma, totals, 1.0f / (float)rows, dx, C, total);
  ESSB_LAUNCH_CHECK("essb_bn_bwd_pass2");
  return ESSB_OK;
}

extern "C" int essb_l1_fwd(const float* a, const float* b, int64_t n, double* sums, void* stream) {
  ESSB_REQUIRE(a && b && sums && n > 0, "essb_l1_fwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(sums, 0, sizeof(double), st) != cudaSuccess) {
    essb_set_error("essb_l1_fwd: memset failed");
    return ESSB_ERR_LAUNCH;
  }
  l1_fwd_kernel<<<grid_for(n, 8), 256, 0, st>>>(a, b, n, sums);
  ESSB_LAUNCH_CHECK("essb_l1_fwd");
  return ESSB_OK;
}

extern "C" int essb_l1_bwd(const float* a, const float* b, int64_t n, const float* gscale, float* da, void* stream) {
  ESSB_REQUIRE(a && b && gscale && da && n > 0, "essb_l1_bwd: bad arguments");
  l1_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a, b, n, gscale, da);
  ESSB_LAUNCH_CHECK("essb_l1_bwd");
  return ESSB_OK;
}

extern "C" int essb_jsdiv(const float* predict, int ld_p, const float* target, int ld_t, int64_t rows, int K,
                          double* sums, const float* gscale, float* dpredict, int ld_d, void* stream) {
  ESSB_REQUIRE(predict && target && rows > 0 && K > 0 && K <= 32 && ld_p >= K && ld_t >= K && (sums || dpredict),
               "essb_jsdiv: bad arguments (K <= 32)");
  cudaStream_t st = (cudaStream_t)stream;
  if (sums && cudaMemsetAsync(sums, 0, sizeof(double), st) != cudaSuccess) {
    essb_set_error("essb_jsdiv: memset failed");
    return ESSB_ERR_LAUNCH;
  }
  const unsigned grid = grid_for(rows, 4);
  if (K <= 8) jsdiv_kernel<8><<<grid, 256, 0, st>>>(predict, ld_p, target, ld_t, rows, K, sums, gscale, dpredict, ld_d);
  else if (K <= 16) jsdiv_kernel<16><<<grid, 256, 0, st>>>(predict, ld_p, target, ld_t, rows, K, sums, gscale, dpredict, ld_d);
  else jsdiv_kernel<32><<<grid, 256, 0, st>>>(predict, ld_p, target, ld_t, rows, K, sums, gscale, dpredict, ld_d);
  ESSB_LAUNCH_CHECK("essb_jsdiv");
  return ESSB_OK;
}
