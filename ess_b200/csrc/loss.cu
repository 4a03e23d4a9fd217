// Task loss (Dice + cross-entropy) forward / backward and the confusion matrix.
// Reference: utils/loss_functions.py:6-24 (TaskLoss), :63-135 (BinaryDiceLoss, DiceLoss),
// torch.nn.CrossEntropyLoss(ignore_index) (:15); evaluation/metrics.py:4-24.
//
// HBM-bound: the forward reads K logits + one label per pixel once; per-pixel softmax lives in
// registers; per-class partial sums are reduced warp -> block -> one double atomic per block.
// The reference needs ~60 launches and materialises a one-hot tensor of the logits' size.
#include "common.cuh"

namespace {

constexpr int LOSS_THREADS = 256;

template <int KMAX>
__device__ __forceinline__ void load_softmax(const float* __restrict__ q, int K, float (&p)[KMAX], float& lse) {
  float mx = -INFINITY;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    p[k] = (k < K) ? q[k] : -INFINITY;
    mx = fmaxf(mx, p[k]);
  }
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    p[k] = (k < K) ? expf(p[k] - mx) : 0.f;
    sum += p[k];
  }
  const float inv = 1.f / sum;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) p[k] *= inv;
  lse = mx + logf(sum);
}

template <int KMAX>
__global__ void __launch_bounds__(LOSS_THREADS) task_loss_fwd_kernel(const float* __restrict__ logits, int ld,
                                                                     const int64_t* __restrict__ target,
                                                                     long long npix, int K, long long ignore_index,
                                                                     double* __restrict__ sums) {
  float ce = 0.f, cnt = 0.f;
  float I[KMAX], S[KMAX], T[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) { I[k] = 0.f; S[k] = 0.f; T[k] = 0.f; }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix;
       i += (long long)gridDim.x * blockDim.x) {
    const long long t64 = target[i];
    if (t64 == ignore_index) continue;
    const int t = (t64 >= 0 && t64 < K) ? (int)t64 : -1;      // 32-bit compares in the unrolled class loop
    float p[KMAX], lse;
    const float* q = logits + i * ld;
    load_softmax<KMAX>(q, K, p, lse);
    cnt += 1.f;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      if (k < K) {
        S[k] += p[k] * p[k];
        if (t == k) { I[k] += p[k]; T[k] += 1.f; ce += lse - q[k]; }
      }
    }
  }
  // block reduction: 2 + 3K values
  __shared__ double red[LOSS_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  auto block_add = [&](float v, int slot) {
    double d = warp_sum_d((double)v);
    if (lane == 0) red[warp] = d;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < LOSS_THREADS / 32; ++w) t += red[w];
      if (t != 0.0) atomicAdd(&sums[slot], t);
    }
    __syncthreads();
  };
  block_add(ce, 0);
  block_add(cnt, 1);
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    if (k < K) {
      block_add(I[k], 2 + k);
      block_add(S[k], 2 + K + k);
      block_add(T[k], 2 + 2 * K + k);
    }
  }
}

__global__ void task_loss_finish_kernel(const double* __restrict__ sums, int K, long long ignore_index, int use_dice,
                                        int use_ce, float* __restrict__ loss) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double total = 0.0;
  if (use_dice) {
    double dice = 0.0;
    for (int k = 0; k < K; ++k) {
      if ((long long)k == ignore_index) continue;  // loss_functions.py:128
      const double num = 2.0 * sums[2 + k] + 1.0;
      const double den = sums[2 + K + k] + sums[2 + 2 * K + k] + 1.0;
      dice += 1.0 - num / den;
    }
    total += dice / (double)K;  // :135 divides by target.shape[1] = K
  }
  if (use_ce) total += sums[0] / sums[1];  // mean over non-ignored pixels (NaN if there are none, as torch)
  loss[0] = (float)total;
}

template <int KMAX>
__global__ void __launch_bounds__(LOSS_THREADS) task_loss_bwd_kernel(
    const float* __restrict__ logits, int ld, const int64_t* __restrict__ target, long long npix, int K,
    long long ignore_index, const double* __restrict__ sums, int use_dice, int use_ce,
    const float* __restrict__ gscale, float* __restrict__ dlogits, int ld_d) {
  // per-class Dice coefficients: dL/dp_k = a_k * p_k - b_k * t_k
  __shared__ float sa[KMAX], sb[KMAX];
  __shared__ float s_ce;
  if (threadIdx.x < KMAX) {
    const int k = threadIdx.x;
    float a = 0.f, b = 0.f;
    if (k < K && use_dice && (long long)k != ignore_index) {
      const double num = 2.0 * sums[2 + k] + 1.0;
      const double den = sums[2 + K + k] + sums[2 + 2 * K + k] + 1.0;
      a = (float)(2.0 * num / (den * den) / (double)K);
      b = (float)(2.0 / den / (double)K);
    }
    sa[k] = a;
    sb[k] = b;
  }
  if (threadIdx.x == 0) s_ce = use_ce ? (float)(1.0 / sums[1]) : 0.f;
  __syncthreads();
  const float gs = gscale ? gscale[0] : 1.f;
  const float ce_w = s_ce;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix;
       i += (long long)gridDim.x * blockDim.x) {
    const long long t64 = target[i];
    float* o = dlogits + i * ld_d;
    const int t = (t64 >= 0 && t64 < K) ? (int)t64 : -1;      // 32-bit compares in the unrolled class loops
    if (t64 == ignore_index) {
#pragma unroll
      for (int k = 0; k < KMAX; ++k)
        if (k < K) o[k] = 0.f;
      continue;
    }
    float p[KMAX], lse;
    load_softmax<KMAX>(logits + i * ld, K, p, lse);
    float gk[KMAX];
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      gk[k] = sa[k] * p[k] - ((t == k) ? sb[k] : 0.f);
      dot += gk[k] * p[k];
    }
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      if (k < K) {
        const float d_dice = p[k] * (gk[k] - dot);
        const float d_ce = ce_w * (p[k] - ((t == k) ? 1.f : 0.f));
        o[k] = gs * (d_dice + d_ce);
      }
    }
  }
}

template <int KMAX>
__global__ void confusion_kernel(const float* __restrict__ logits, int ld, const int64_t* __restrict__ target,
                                 long long npix, int K, long long ignore_index, unsigned long long* __restrict__ conf) {
  __shared__ unsigned int hist[KMAX * KMAX];
  for (int i = threadIdx.x; i < K * K; i += blockDim.x) hist[i] = 0u;
  __syncthreads();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix;
       i += (long long)gridDim.x * blockDim.x) {
    const long long t = target[i];
    if (t == ignore_index || t < 0 || t >= K) continue;
    const float* q = logits + i * ld;
    int best = 0;
    float bv = q[0];
    for (int k = 1; k < K; ++k) {
      const float v = q[k];
      if (v > bv) { bv = v; best = k; }  // first maximum wins, as torch.argmax
    }
    atomicAdd(&hist[(int)t * K + best], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * K; i += blockDim.x)
    if (hist[i]) atomicAdd(&conf[i], (unsigned long long)hist[i]);
}

__global__ void confusion_labels_kernel(const int64_t* __restrict__ pred, const int64_t* __restrict__ target,
                                        long long npix, int K, long long ignore_index,
                                        unsigned long long* __restrict__ conf) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix;
       i += (long long)gridDim.x * blockDim.x) {
    const long long t = target[i], y = pred[i];
    if (t == ignore_index || t < 0 || t >= K || y < 0 || y >= K) continue;
    atomicAdd(&conf[t * K + y], 1ull);
  }
}

inline unsigned loss_grid(long long npix) {
  long long b = (npix + LOSS_THREADS * 8 - 1) / (LOSS_THREADS * 8);
  if (b > 148 * 8) b = 148 * 8;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace

// the class loops are fully unrolled over KM >= K and these kernels are instruction-bound (profiles/r02o_hbm_kernels.md):
// KM close to K matters (K = 11 -> 12 instead of 16: a quarter fewer exp / compare / accumulate instructions)
#define ESSB_DISPATCH_K(K, CALL)                      \
  do {                                                \
    if ((K) <= 4) { constexpr int KM = 4; CALL; }     \
    else if ((K) <= 6) { constexpr int KM = 6; CALL; } \
    else if ((K) <= 8) { constexpr int KM = 8; CALL; } \
    else if ((K) <= 12) { constexpr int KM = 12; CALL; } \
    else if ((K) <= 16) { constexpr int KM = 16; CALL; } \
    else if ((K) <= 20) { constexpr int KM = 20; CALL; } \
    else { constexpr int KM = 32; CALL; }             \
  } while (0)

extern "C" int essb_task_loss_fwd(const float* logits, int ld, const int64_t* target, int64_t npix, int K,
                                  int64_t ignore_index, double* sums, void* stream) {
  ESSB_REQUIRE(logits && target && sums && npix > 0 && K > 0 && K <= 32 && ld >= K,
               "essb_task_loss_fwd: bad arguments (K <= 32)");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(double) * (2 + 3 * K), st);
  if (e != cudaSuccess) {
    essb_set_error("essb_task_loss_fwd: memset failed: %s", cudaGetErrorString(e));
    return ESSB_ERR_LAUNCH;
  }
  ESSB_DISPATCH_K(K, (task_loss_fwd_kernel<KM><<<loss_grid(npix), LOSS_THREADS, 0, st>>>(
                         logits, ld, target, npix, K, ignore_index, sums)));
  ESSB_LAUNCH_CHECK("essb_task_loss_fwd");
  return ESSB_OK;
}

extern "C" int essb_task_loss_finish(const double* sums, int K, int64_t ignore_index, int use_dice, int use_ce,
                                     float* loss, void* stream) {
  ESSB_REQUIRE(sums && loss && K > 0 && K <= 32, "essb_task_loss_finish: bad arguments");
  task_loss_finish_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(sums, K, ignore_index, use_dice, use_ce, loss);
  ESSB_LAUNCH_CHECK("essb_task_loss_finish");
  return ESSB_OK;
}

extern "C" int essb_task_loss_bwd(const float* logits, int ld, const int64_t* target, int64_t npix, int K,
                                  int64_t ignore_index, const double* sums, int use_dice, int use_ce,
                                  const float* gscale, float* dlogits, int ld_d, void* stream) {
  ESSB_REQUIRE(logits && target && sums && dlogits && npix > 0 && K > 0 && K <= 32 && ld >= K && ld_d >= K,
               "essb_task_loss_bwd: bad arguments (K <= 32)");
  cudaStream_t st = (cudaStream_t)stream;
  ESSB_DISPATCH_K(K, (task_loss_bwd_kernel<KM><<<loss_grid(npix), LOSS_THREADS, 0, st>>>(
                         logits, ld, target, npix, K, ignore_index, sums, use_dice, use_ce, gscale, dlogits, ld_d)));
  ESSB_LAUNCH_CHECK("essb_task_loss_bwd");
  return ESSB_OK;
}

extern "C" int essb_confusion(const float* logits, int ld, const int64_t* target, int64_t npix, int K,
                              int64_t ignore_index, int64_t* conf, void* stream) {
  ESSB_REQUIRE(logits && target && conf && npix > 0 && K > 0 && K <= 32 && ld >= K, "essb_confusion: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  ESSB_DISPATCH_K(K, (confusion_kernel<KM><<<loss_grid(npix), LOSS_THREADS, 0, st>>>(
                         logits, ld, target, npix, K, ignore_index, reinterpret_cast<unsigned long long*>(conf))));
  ESSB_LAUNCH_CHECK("essb_confusion");
  return ESSB_OK;
}

extern "C" int essb_confusion_labels(const int64_t* pred, const int64_t* target, int64_t npix, int K,
                                     int64_t ignore_index, int64_t* conf, void* stream) {
  ESSB_REQUIRE(pred && target && conf && npix > 0 && K > 0, "essb_confusion_labels: bad arguments");
  confusion_labels_kernel<<<loss_grid(npix), LOSS_THREADS, 0, (cudaStream_t)stream>>>(
      pred, target, npix, K, ignore_index, reinterpret_cast<unsigned long long*>(conf));
  ESSB_LAUNCH_CHECK("essb_confusion_labels");
  return ESSB_OK;
}
