/* ess_b200 -- C ABI of the B200-native (sm_100a) kernels behind the ESS hot path.
 *
 * The reference (uzh-rpg/ess) has no FFI layer: every op below replaces an ATen/cuDNN library call
 * reached from a reference nn.Module.  Each entry point cites the reference call site it replaces
 * (paths relative to the reference root).  Conventions (all entry points):
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer owned by the caller
 *     (PyTorch's caching allocator); the library never allocates, frees or retains device memory;
 *   - asynchronous on the given `stream` (a cudaStream_t passed as void*), never synchronises;
 *   - returns 0 on success, a negative essb_status otherwise; essb_last_error() gives a
 *     thread-local message.  No exceptions cross the boundary; no exit();
 *   - activations are fp32 NHWC ("pixel-major"): element (n, y, x, c) at ((n*H + y)*W + x)*ld + c
 *     where `ld` (floats per pixel) may exceed C so that channel slices of wider tensors can be
 *     addressed without copies.
 */
#ifndef ESS_B200_H
#define ESS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ESSB_VERSION 100
#define ESSB_MAX_TAPS 49   /* up to a 7x7 filter (ResNet stem of StyleEncoderE2VID) */

typedef enum essb_status {
  ESSB_OK = 0,
  ESSB_ERR_ARG = -1,      /* bad shape / alignment / null pointer */
  ESSB_ERR_ARCH = -2,     /* device is not sm_100 */
  ESSB_ERR_LAUNCH = -3,   /* cudaGetLastError() after launch */
  ESSB_ERR_DRIVER = -4,   /* driver entry point (tensor map encode) unavailable */
  ESSB_ERR_WORKSPACE = -5 /* workspace too small */
} essb_status;

/* ---- library info ----------------------------------------------------------------------- */
int essb_version(void);
const char* essb_build_arch(void);        /* "sm_100a" */
const char* essb_last_error(void);
int essb_device_check(void);              /* ESSB_OK iff current device is compute capability 10.x */

/* ---- implicit-GEMM convolution --------------------------------------------------------- */
/* One input channel segment.  A conv reads the channel-concatenation of up to two segments
 * (replaces torch.cat in ConvLSTM.forward, e2vid/model/submodules.py:212, and skip_concat in
 * SemSegE2VID.forward, models/style_networks.py:78,82) and can apply, while loading,
 *   nearest x2 upsampling   (f.interpolate(scale_factor=2,'nearest'), style_networks.py:77,81,85)
 *   instance normalisation  (nn.InstanceNorm2d(affine=False), style_networks.py:163,180,183)
 *   ReLU                    (style_networks.py:164,181)
 * so none of those intermediate tensors is ever written to HBM. */
typedef struct essb_src {
  const float* ptr;   /* NHWC fp32, source dims (H>>ups, W>>ups); NULL => segment absent */
  const float* mean;  /* [N, C] or NULL */
  const float* rstd;  /* [N, C] or NULL */
  int32_t ld;         /* floats per pixel in memory */
  int32_t C;          /* channels in this segment */
  int32_t ups;        /* 0 or 1 */
  int32_t relu;       /* apply max(.,0) after normalisation */
} essb_src;

typedef enum essb_epilogue {
  ESSB_EPI_LINEAR = 0, /* out = post( act(acc + bias + res_pre) ) + res_post                      */
  ESSB_EPI_LSTM = 1,   /* ConvLSTM gates, e2vid/model/submodules.py:216-228 (Cout = 4*hidden,
                          packed gate-interleaved: co = 4*ch + {in, remember, out, cell})          */
  ESSB_EPI_GRU_UR = 2, /* ConvGRU update/reset, submodules.py:267-268 (co = 2*ch + {update,reset});
                          writes update -> out, prev_state*reset -> out2                            */
  ESSB_EPI_GRU_OUT = 3 /* ConvGRU out gate + blend, submodules.py:269-271                          */
} essb_epilogue;

typedef enum essb_act { ESSB_ACT_NONE = 0, ESSB_ACT_RELU = 1, ESSB_ACT_SIGMOID = 2 } essb_act;

/* Generic gather-convolution launch:
 *   acc[n,oy,ox,co] = sum_t sum_ci  A[n, oy*sy + dy[t], ox*sx + dx[t], ci] * w[widx[t]][ci][co]
 * with A the (upsampled / normalised / ReLU'd / concatenated) virtual input of dims H x W, zero
 * outside.  The result for (oy,ox) in [0,OH)x[0,OW) is stored at pixel (oy*osy+ooy, ox*osx+oox) of
 * an OHf x OWf output image -- strided output implements the four sub-pixel phases of
 * ConvTranspose2d(k5,s2,p2,op1) (e2vid/model/submodules.py:39-40).
 * Replaces nn.Conv2d / nn.ConvTranspose2d forward AND their autograd input-gradient
 * (dgrad = the same gather with flipped taps and transposed weights). */
typedef struct essb_conv {
  essb_src src[2];
  const float* w;        /* packed weights [n_w_taps][Cin_total][CoutP], CoutP = Cout rounded up to 4 */
  const float* bias;     /* [Cout] or NULL */
  const float* res_pre;  /* LINEAR: added before the activation (ResidualBlock, submodules.py:170) or NULL */
  const float* res_post; /* LINEAR: added after the activation (skip_sum, e2vid/model/unet.py:175,179) or NULL */
  const float* aux0;     /* LSTM: prev cell [N,OH,OW,hidden] or NULL(=0). GRU_*: prev state or NULL(=0) */
  const float* aux1;     /* GRU_OUT: update gate [N,OH,OW,hidden] */
  float* out;            /* LINEAR: result. LSTM: hidden. GRU_UR: update. GRU_OUT: new state */
  float* out2;           /* LSTM: cell. GRU_UR: prev_state*reset. else NULL */
  float* stats_partial;  /* LINEAR only: per-tile (sum, sumsq) partials [N][tiles][Cout][2] or NULL */
  uint16_t* out_hi;      /* LINEAR only, optional: bf16 hi/lo planes of the result (operand format of */
  uint16_t* out_lo;      /*   the tensor-core path), pixel pitch ld_planes; needs Cout % 4 == 0      */
  int32_t N, H, W;       /* virtual input dims */
  int32_t OH, OW;        /* output grid of this launch */
  int32_t Cout;          /* GEMM N (for LSTM = 4*hidden, GRU_UR = 2*hidden) */
  int32_t sy, sx;        /* input step per output pixel */
  int32_t OHf, OWf, osy, ooy, osx, oox; /* output image dims and phase placement */
  int32_t ldo;           /* floats per output pixel (out, out2, res_*, aux* all share it for LINEAR;
                            for LSTM/GRU all state tensors are dense [.., hidden]) */
  int32_t ld_res;        /* floats per pixel of res_pre / res_post */
  int32_t ld_planes;     /* elements per pixel of out_hi / out_lo */
  int32_t accumulate;    /* LINEAR: out += result (used to sum gradient contributions) */
  int32_t epilogue;      /* essb_epilogue */
  int32_t act;           /* essb_act (LINEAR) */
  int32_t ntaps;
  int8_t dy[ESSB_MAX_TAPS];
  int8_t dx[ESSB_MAX_TAPS];
  int8_t widx[ESSB_MAX_TAPS];
} essb_conv;

/* fp32 CUDA-core path (exact fp32 FMA accumulation; parity reference mode "fp32"). */
int essb_conv_fp32(const essb_conv* d, void* stream);
int essb_conv_tiles_per_sample(const essb_conv* d); /* tiles dimension of stats_partial */

/* Weight gradient of the same gather-convolution (autograd of nn.Conv2d w.r.t. weight/bias):
 *   dW[co][ci][t] = sum_{n,oy,ox} A[n, oy*sy+dy[t], ox*sx+dx[t], ci] * dY[n,oy,ox,co]
 * written in the reference parameter layout [Cout, Cin, KH*KW] (tap index = position in dy/dx);
 * dbias[co] = sum dY.  `workspace` holds split-K partials. */
typedef struct essb_wgrad {
  essb_src src[2];
  const float* dy_ptr;   /* [N, OH, OW, ld_dy] */
  float* dw;             /* [Cout, Cin_total, ntaps] */
  float* dbias;          /* [Cout] or NULL */
  float* workspace;
  int64_t workspace_bytes;
  int32_t N, H, W, OH, OW, Cout, ld_dy, sy, sx;
  int32_t ntaps;
  int8_t dy[ESSB_MAX_TAPS];
  int8_t dx[ESSB_MAX_TAPS];
} essb_wgrad;
int64_t essb_wgrad_workspace_bytes(const essb_wgrad* d);
int essb_wgrad_fp32(const essb_wgrad* d, void* stream);

/* Weight packing: reference layout -> packed [T][Cin][CoutP].
 *   transposed_layout = 0: w is nn.Conv2d weight  [Cout, Cin, T]
 *   transposed_layout = 1: w is nn.ConvTranspose2d weight [Cin, Cout, T]
 *   swap_io: pack for dgrad (roles of Cin/Cout exchanged); flip: reverse tap order (dgrad)
 *   scale[Cout] (or NULL): eval-mode BatchNorm fold, gamma/sqrt(var+eps) (submodules.py:19-20,26-27)
 *   interleave g > 1: packed output channel co' = (co % (Cout/g))*g + co / (Cout/g)
 *                     (gate-interleaving for ESSB_EPI_LSTM / GRU_UR) */
int essb_pack_weight(const float* w, const float* scale, float* out, int Cout, int Cin, int T,
                     int transposed_layout, int swap_io, int flip, int interleave, void* stream);

/* ---- instance-norm helpers (nn.InstanceNorm2d fwd/bwd, style_networks.py:163,180,183) ---- */
/* partial [N][tiles][C][2] -> mean, rstd [N][C]; biased variance, eps inside the sqrt. */
int essb_in_finalize(const float* partial, int N, int tiles, int C, int64_t count, float eps,
                     float* mean, float* rstd, void* stream);
/* out = act((y - mean)*rstd) + res   (materialises an IN(+ReLU)(+residual) output, e.g. the
 * INSResBlock output `out += residual`, style_networks.py:190-193).  mean==NULL => identity norm. */
int essb_norm_act_add(const float* y, int ld_y, const float* mean, const float* rstd, int relu,
                      const float* res, int ld_res, float* out, int ld_out, int N, int64_t P, int C,
                      void* stream);
/* Backward through  A = relu?(IN(y))  (optionally followed by nearest x2 upsampling), two passes.
 * pass 1: g = [relu mask] * (sum over the 2^ups x 2^ups children of dA) (+ extra);  writes g and
 *         per-block partial sums of g and g*xhat  -> partial [N][blocks][C][2]
 * pass 2: dy = rstd * (g - mean_p(g) - xhat * mean_p(g*xhat))                                    */
int essb_in_bwd_blocks(int64_t P);
/* per-block (sum, sumsq) partials of y [N][P][ld_y] -> partial [N][essb_in_bwd_blocks(P)][C][2]; used
 * on the tensor-core path, whose conv epilogue does not emit the statistics itself. */
int essb_in_stats(const float* y, int ld_y, float* partial, int N, int64_t P, int C, void* stream);
int essb_in_bwd_pass1(const float* dA, int ld_dA, int ups, const float* extra, int ld_extra,
                      const float* y, int ld_y, const float* mean, const float* rstd, int relu,
                      float* g, float* partial, int N, int H, int W, int C, void* stream);
/* dy (fp32, pitch C) and/or its bf16 hi/lo planes (pitch ld_planes >= C, channels [C, ld_planes) zeroed:
 * the operand format of the producing conv's tcgen05 dgrad / wgrad) -- either output may be NULL. */
int essb_in_bwd_pass2(const float* g, const float* y, int ld_y, const float* mean, const float* rstd,
                      const float* gsum /* [N][C][2] totals */, float* dy, uint16_t* dy_hi, uint16_t* dy_lo,
                      int ld_planes, int N, int64_t P, int C, void* stream);
/* totals[N][C][2] = sum over blocks of partial (fixed order => deterministic) */
int essb_partial_reduce(const float* partial, int N, int blocks, int C, float* totals, void* stream);
/* out[c] = sum over rows of x[rows][ld] (bias gradients) */
int essb_colsum(const float* x, int ld, int64_t rows, int C, float* out, float* workspace,
                int64_t workspace_bytes, void* stream);
/* g_child -> parent:  out[n,y,x,c] = sum over the 2x2 children of in (backward of nearest x2) */
int essb_upsample2_bwd(const float* in, int ld_in, float* out, int ld_out, int N, int H, int W, int C,
                       int accumulate, void* stream);

/* ---- 1x1 convolutions with few output channels (HBM-bound; pw_conv.cu) ---------------------------
 * Replaces nn.Conv2d(32, K, 1) of SemSegE2VID.decoder_scale_5 (models/style_networks.py:34,88: forward, and
 * its autograd input / weight / bias gradients) and the E2VID prediction layer conv1x1 + BatchNorm(eval) +
 * sigmoid (e2vid/model/unet.py:65-67,179; BN folded into w / bias by the caller).
 * src: fp32 NHWC with C = 32 or 64 channels, no upsampling; its optional (mean, rstd, relu) apply the
 * preceding InstanceNorm + ReLU on the fly.  w is the reference layout [Cout][Cin] (fp32), 1 <= Cout <= 16.
 *   fwd  : out[p][k] = act(sum_c f(x[p][c]) * w[k][c] + bias[k])          (out pitch ldo >= Cout)
 *   dgrad: dx[p][c]  = sum_k dy[p][k] * w[k][c]                           (gradient w.r.t. f(x))
 *   wgrad: dw[k][c]  = sum_p dy[p][k] * f(x[p][c]),  dbias[k] = sum_p dy[p][k]   (either may be NULL);
 *          two-stage fixed-order reduction (deterministic) through `workspace`. */
int essb_pw_conv_fwd(const essb_src* src, const float* w, const float* bias, float* out, int ldo, int N,
                     int H, int W, int Cout, int act, void* stream);
int essb_pw_conv_dgrad(const float* dy, int ld_dy, const float* w, float* dx, int ld_dx, int64_t rows,
                       int Cin, int Cout, void* stream);
int64_t essb_pw_conv_wgrad_workspace_bytes(int N, int H, int W, int Cin);
int essb_pw_conv_wgrad(const essb_src* src, const float* dy, int ld_dy, int N, int H, int W, int Cout,
                       float* dw, float* dbias, float* workspace, int64_t workspace_bytes, void* stream);

/* ---- single-input-channel stem convolution (stem_conv.cu) ------------------------------------------
 * Replaces resnet18.conv1 = nn.Conv2d(1, 64, 7, stride 2, padding 3, bias=False) of StyleEncoderE2VID
 * (models/style_networks.py:117-121) forward and its autograd weight gradient (UDA step,
 * training/ess_trainer.py:159-162,282).  x: [N][H][W] fp32 (one channel), w / dw: [Cout][k*k] (reference
 * layout), out / dy: pixel-major [N][OH][OW][Cout].  Only (Cout, k, stride) = (64, 7, 2) is built
 * (essb_stem_conv_supported); HBM-bound, deterministic two-stage wgrad reduction through `workspace`. */
int essb_stem_conv_supported(int Cout, int k, int stride);
int essb_stem_conv_fwd(const float* x, const float* w, float* out, int N, int H, int W, int Cout, int k,
                       int stride, int pad, void* stream);
int64_t essb_stem_conv_wgrad_workspace_bytes(int Cout, int k);
int essb_stem_conv_wgrad(const float* x, const float* dy, float* dw, int N, int H, int W, int Cout, int k,
                         int stride, int pad, float* workspace, int64_t workspace_bytes, void* stream);

/* ---- event pre-processing (e2vid/utils/inference_utils.py:84-109, 311-338) --------------- */
/* stats[w][3] = (sum x, sum x^2, count of non-zeros) of window w for ALL T windows in one launch;
 * x is [B][T][count] with batch stride `bstride` floats (count = C*H*W of one window).  One launch
 * before the unroll removes the reference's per-window blocking `if num_nonzeros > 0` D2H sync. */
/* Hot-pixel removal (EventPreprocessor.__call__, inference_utils.py:88-89): x[:, :, y, x] = 0 for the n listed
 * (x, y) pairs, in place like the reference; runs before essb_event_stats so the statistics exclude them. */
int essb_zero_pixels(float* x, int64_t bstride, int B, int C, int H, int W, const int32_t* xy, int n,
                     void* stream);
int essb_event_stats(const float* x, int64_t bstride, int B, int T, int64_t count, double* stats,
                     void* stream);
/* Normalise non-zeros to mean 0 / std 1 using stats (no-op scale if nnz == 0), reflect-pad to
 * (Hp, Wp) with (pad_top, pad_left), convert NCHW -> NHWC with channel pitch ld_out (extra
 * channels zero-filled).  x is a [B, C, H, W] slice with batch stride `bstride` floats. */
int essb_event_prepare(const float* x, int64_t bstride, const double* stats, int normalize,
                       float* out, int ld_out, int B, int C, int H, int W, int Hp, int Wp,
                       int pad_top, int pad_left, int flip /* torch.flip(dims=[2,3]) before padding (inference_utils.py:92-93) */, void* stream);

/* ---- layout plumbing ---------------------------------------------------------------------- */
int essb_nchw_to_nhwc(const float* in, float* out, int ld_out, int N, int C, int64_t P, void* stream);
int essb_nhwc_to_nchw(const float* in, int ld_in, float* out, int N, int C, int64_t P, void* stream);
/* bilinear x2 upsampling, align_corners=False (UpsampleConvLayer, e2vid/model/submodules.py:84) */
int essb_bilinear_up2(const float* in, float* out, int N, int H, int W, int C, void* stream);

/* ---- task loss: Dice + cross-entropy (utils/loss_functions.py:6-24, 63-135) ---------------- */
/* logits NHWC [N*P][ld] fp32, target int64 [N*P].  sums layout (double):
 *   [0] = sum of -log softmax[target] over valid pixels, [1] = valid pixel count,
 *   [2 + k] = I_k = sum p_k t_k, [2 + K + k] = S_k = sum p_k^2, [2 + 2K + k] = T_k = sum t_k.
 * essb_task_loss_fwd zeroes and fills `sums`; _finish writes loss[0] = dice*use_dice + ce*use_ce.
 * Between the two a data-parallel caller all-reduces `sums` (global-batch semantics). */
int essb_task_loss_fwd(const float* logits, int ld, const int64_t* target, int64_t npix, int K,
                       int64_t ignore_index, double* sums, void* stream);
int essb_task_loss_finish(const double* sums, int K, int64_t ignore_index, int use_dice, int use_ce,
                          float* loss, void* stream);
/* dlogits = gscale[0] * dLoss/dlogits  (gscale: device scalar, the upstream gradient). */
int essb_task_loss_bwd(const float* logits, int ld, const int64_t* target, int64_t npix, int K,
                       int64_t ignore_index, const double* sums, int use_dice, int use_ce,
                       const float* gscale, float* dlogits, int ld_d, void* stream);
/* argmax over channels + confusion matrix conf[y][y_hat] (evaluation/metrics.py:4-24), int64. */
int essb_confusion(const float* logits, int ld, const int64_t* target, int64_t npix, int K,
                   int64_t ignore_index, int64_t* conf, void* stream);
/* same from already-computed label predictions (MetricsSemseg.update_batch, metrics.py:50-56) */
int essb_confusion_labels(const int64_t* pred, const int64_t* target, int64_t npix, int K,
                          int64_t ignore_index, int64_t* conf, void* stream);

/* ---- train-mode BatchNorm + UDA consistency losses (UDA companion path) ----
 * StyleEncoderE2VID = ResNet-18 stem + layer1-3 with BatchNorm in TRAIN mode (models/style_networks.py:
 * 110-145, training/ess_trainer.py:159-162).  Per-channel batch statistics reuse essb_in_stats /
 * essb_in_finalize / essb_partial_reduce with N = 1 over rows = N*H*W. */
/* out = relu?(x * a[c] + b[c] + res)   (BN apply with folded gamma/beta, BasicBlock residual add) */
int essb_affine_act(const float* x, int ld_x, const float* a, const float* b, const float* res, int ld_res,
                    int relu, float* out, int ld_out, int64_t rows, int C, void* stream);
/* Train-mode nn.BatchNorm2d bookkeeping in one launch (torchvision resnet18 layers of StyleEncoderE2VID,
 * models/style_networks.py:117-121 in .train()): a = gamma*rstd, b = beta - mean*a for essb_affine_act, and (when
 * running_mean != NULL) running_* = (1-momentum)*running_* + momentum*{mean, unbiased batch variance},
 * *num_batches_tracked += 1 (NULL = skip).  rows = N*H*W. */
int essb_bn_train_finalize(const float* mean, const float* rstd, const float* gamma, const float* beta, float* a,
                           float* b, float* running_mean, float* running_var, int64_t* num_batches_tracked, int C,
                           int64_t rows, float momentum, float eps, void* stream);
int essb_bn_bwd_blocks(int64_t rows);
/* g = dout * (mask > 0) (mask = the block's post-ReLU output, or NULL); partial [blocks][C][2] = sums of
 * g and g*xhat per block.  dbeta = sum g, dgamma = sum g*xhat after essb_partial_reduce. */
int essb_bn_bwd_pass1(const float* dout, int ld_d, const float* mask, int ld_m, const float* x, int ld_x,
                      const float* mean, const float* rstd, float* g, float* partial, int64_t rows, int C,
                      void* stream);
/* dx = gamma * rstd * (g - sum(g)/M - xhat * sum(g*xhat)/M), totals = [C][2] */
int essb_bn_bwd_pass2(const float* g, const float* x, int ld_x, const float* mean, const float* rstd,
                      const float* gamma, const float* totals, float* dx, int64_t rows, int C, void* stream);
/* torch.nn.L1Loss (mean): sums[0] = sum |a-b|;  da = gscale[0]/n * sign(a-b) */
int essb_l1_fwd(const float* a, const float* b, int64_t n, double* sums, void* stream);
int essb_l1_bwd(const float* a, const float* b, int64_t n, const float* gscale, float* da, void* stream);
/* symJSDivLoss (utils/loss_functions.py:27-37) on pixel-major logits [rows][K]: sums[0] = rows*K * loss
 * (when sums != NULL); dpredict = gscale[0] * dLoss/dpredict (when dpredict != NULL; target gets none). */
int essb_jsdiv(const float* predict, int ld_p, const float* target, int ld_t, int64_t rows, int K,
               double* sums, const float* gscale, float* dpredict, int ld_d, void* stream);

/* ---- events -> voxel grid (the tensor-building step in front of the path) ------ */
/* VoxelGrid.convert without normalisation (DSEC/dataset/representations.py:15-44): x, y, pol, t are float
 * device arrays of n events (time-sorted); grid [C][H][W] is zeroed and filled by trilinear scatter. */
int essb_voxel_grid_dsec(const float* x, const float* y, const float* pol, const float* t, int64_t n, int C,
                         int H, int W, float* grid, void* stream);
/* generate_voxel_grid (datasets/data_util.py:54-126): events = n rows of double [x, y, t, polarity in {0,1}];
 * grid [2C][H][W] (separate_pol: positive then negative) or [C][H][W] (positive - negative). */
int essb_voxel_grid_ddd17(const double* events, int64_t n, int C, int H, int W, int separate_pol, float* grid,
                          void* stream);

/* ---- optimizer (utils/radam.py:15-80, one fused elementwise update per tensor) ------------------ */
/* v = b2*v + (1-b2)*g^2; m = b1*m + (1-b1)*g; p -= wd_lr*p; p -= step_lr * (rectified ? m/(sqrt(v)+eps) : m)
 * with step_lr = step_size*lr and wd_lr = weight_decay*lr computed on the host exactly as radam.py:53-75. */
int essb_radam_step(float* p, const float* g, float* m, float* v, int64_t n, float beta1, float beta2,
                    float step_lr, float eps, float wd_lr, int rectified, void* stream);
/* Multi-tensor form of the same update (the reference loops over `group['params']` in Python, utils/radam.py:25-78):
 * up to ESSB_RADAM_MAX tensors that share the hyper-parameters and the step count are updated by ONE launch. */
#define ESSB_RADAM_MAX 48
typedef struct essb_radam_multi {
  float* p[ESSB_RADAM_MAX];
  const float* g[ESSB_RADAM_MAX];
  float* m[ESSB_RADAM_MAX];
  float* v[ESSB_RADAM_MAX];
  int64_t n[ESSB_RADAM_MAX];
  int32_t count;
  float beta1, beta2, step_lr, eps, wd_lr;
  int32_t rectified;
} essb_radam_multi;
int essb_radam_multi_step(const essb_radam_multi* d, void* stream);

/* ---- tcgen05 / TMA tensor-core path (sm_100a) --------------------------------------------- */
/* Split an fp32 tensor into bf16 hi/lo planes: hi = bf16(x), lo = bf16(x - hi)  (x ~= hi + lo to
 * 2^-17 relative).  Optionally applies the same normalise/ReLU/upsample transform as essb_src. */
int essb_split_bf16(const essb_src* src, int N, int H, int W, uint16_t* hi, uint16_t* lo, int ld_out,
                    int c_off, int c_pad /* channels [C, c_pad) are written as zeros (K padding); 0 = none */,
                    void* stream);
/* Same with the output format selectable: fmt 0 = bf16 hi/lo planes; fmt 2 = "hf8": hi = fp16(x * 2^6), and per 64-channel
 * chunk of the lo plane 128 bytes [e4m3(x * 2^3) x 64 | e4m3((x - hi/2^6) * 2^14) x 64] (ld_out % 64 == 0).  The tcgen05
 * kernels consume hf8 operands with passes == 2 (fp16 main product + e4m3 cross terms). */
int essb_split_planes(const essb_src* src, int N, int H, int W, uint16_t* hi, uint16_t* lo, int ld_out,
                      int c_off, int c_pad, int fmt, void* stream);
/* Event pre-processing (same arithmetic as essb_event_prepare) written directly in the head
 * convolution's tensor-core operand format: bf16 hi/lo planes with `cpad` (8 or 16) channels per
 * pixel, stored at (off_y, off_x) inside a caller-zeroed bordered buffer [B][Hb][Wb][cpad].  With a
 * border of 2 rows/columns before and >= 2 rows / 6 columns after, the head conv5x5 (C=5, unet.py:131)
 * reads, per kernel row, the 8-pixel window starting at x-2 as ONE contiguous 8*cpad-element K-chunk
 * through an overlapping-stride TMA view (element stride of x = cpad). */
int essb_event_prepare_planes(const float* x, int64_t bstride, const double* stats, int normalize,
                              uint16_t* hi, uint16_t* lo, int cpad, int B, int C, int H, int W, int Hp,
                              int Wp, int pad_top, int pad_left, int Hb, int Wb, int off_y, int off_x,
                              int flip, void* stream);
/* Pack conv weights for the tensor-core path: K-major [NoutP][T*KinP] bf16 hi and lo planes
 * (same transposed_layout / swap_io / flip / interleave / scale semantics as essb_pack_weight;
 * KinP = input channels padded to a multiple of 64, rows beyond Nout are zero). */
int essb_pack_weight_tc(const float* w, const float* scale, uint16_t* hi, uint16_t* lo, int Cout,
                        int Cin, int T, int transposed_layout, int swap_io, int flip, int interleave,
                        int KinP, int NoutP, void* stream);
/* fmt 2: hf8 weights for passes == 2: hi = fp16(W * 2^(w8+8)), lo = per 64-k chunk [e4m3((W - hi/2^(w8+8)) * 2^(w8+11)) x 64 |
 * e4m3(W * 2^w8) x 64]; w8 = floor(log2(128 / max|W|)) (after the BN fold) is chosen by the caller, who passes
 * acc_scale = 2^-(w8+14) to essb_conv_tc_run.  fmt 0 = essb_pack_weight_tc. */
int essb_pack_weight_tc_fmt(const float* w, const float* scale, uint16_t* hi, uint16_t* lo, int Cout,
                            int Cin, int T, int transposed_layout, int swap_io, int flip, int interleave,
                            int KinP, int NoutP, int fmt, int w8, void* stream);

/* One TMA view of an A operand: bf16 hi/lo NHWC planes addressed as view[n][y][x][c] with element
 * strides (stride_n, stride_y, stride_x, 1).  A dense NHWC tensor is one view; a stride-2 conv
 * reads its input through four parity views (base shifted by (py*W+px)*ld, strides doubled). */
typedef struct essb_tc_view {
  const uint16_t* hi;
  const uint16_t* lo;        /* may be NULL when passes == 1 */
  int64_t stride_x, stride_y, stride_n;  /* in elements; multiples of 8 */
  int32_t C, W, H;           /* view extent: channels (multiple of 64), width, height */
  int32_t reserved;
} essb_tc_view;

/* tcgen05 gather-convolution (same GEMM semantics as essb_conv; operands pre-split to bf16 hi/lo):
 *   acc[n,oy,ox,co] = sum_t sum_seg sum_c  view[seg_view0[seg] + view[t]][n, oy+dy[t], ox+dx[t], c]
 *                                          * W[co][widx[t]*k_per_tap + seg_koff[seg] + c]
 * Epilogues: ESSB_EPI_LINEAR (bias, res_pre, act, res_post, strided placement, fp32 and/or bf16
 * hi/lo outputs), ESSB_EPI_LSTM (as essb_conv; additionally writes the hidden state as bf16 planes so
 * the next window's MMA can consume it without a conversion pass), ESSB_EPI_GRU_UR (update -> out,
 * prev_state*reset -> out_hi/out_lo planes) and ESSB_EPI_GRU_OUT (new state -> out (+ planes)). */
typedef struct essb_conv_tc {
  essb_tc_view views[8];
  const uint16_t* w_hi;  /* [w_rows][n_w_taps * k_per_tap] */
  const uint16_t* w_lo;
  const float* bias;
  const float* res_pre;
  const float* res_post;
  const float* aux0;     /* LSTM: previous cell or NULL; GRU_UR / GRU_OUT: previous state or NULL */
  const float* aux1;     /* GRU_OUT: update gate */
  float* out;            /* LINEAR: fp32 result or NULL; LSTM: hidden; GRU_UR: update; GRU_OUT: new state */
  float* out2;           /* LSTM: cell */
  uint16_t* out_hi;      /* optional bf16 planes of the result (LSTM: of the hidden state) */
  uint16_t* out_lo;
  /* Dynamic tile scheduler (required): 2 int32 that are ZERO when the launch starts; the kernel leaves them
   * zero again (last CTA resets), so a slot can be reused by the next launch on the same stream.  Launches
   * that may overlap (different streams) need different slots. */
  int32_t* sched;
  /* Split-K of the last, partially filled wave of tiles (optional, NULL = off): fp32 partial-accumulator
   * workspace and >= 148*8 zero-initialised int32 arrival counters (left zero by the kernel).  The library
   * splits only as far as splitk_ws_bytes allows; the reduction order is fixed (deterministic). */
  float* splitk_ws;
  int32_t* splitk_cnt;
  int64_t splitk_ws_bytes;
  int32_t n_views, nseg;
  int32_t seg_C[2];      /* channels consumed per segment (multiples of 64) */
  int32_t seg_view0[2];
  int32_t seg_koff[2];
  int32_t k_per_tap, n_w_taps, w_rows;
  int32_t N, OH, OW, Cout;
  int32_t OHf, OWf, osy, ooy, osx, oox;
  int32_t ldo, ld_res, ld_planes;
  int32_t epilogue, act;
  int32_t passes;        /* 3 = bf16x3 split (fp32-parity mode), 1 = single bf16 pass, 2 = f16f8: operands in the hf8
                            format (fp16 main product + e4m3 cross terms = 2 tensor-pass equivalents, fp32-parity) */
  int32_t bw_log2;       /* spatial tile: BW = 1<<bw_log2 columns x 128/BW rows */
  int32_t ntaps;
  int8_t dy[ESSB_MAX_TAPS];
  int8_t dx[ESSB_MAX_TAPS];
  int8_t view[ESSB_MAX_TAPS];
  int8_t widx[ESSB_MAX_TAPS];
  float acc_scale;       /* the epilogue multiplies the accumulator by this first (0 = 1.0); passes == 2: the
                            2^-(w8+14) that belongs to the packed weights */
  int32_t planes_fmt;    /* format of out_hi / out_lo: 0 = bf16 hi/lo planes, 2 = hf8 (ld_planes % 64 == 0) */
  /* Row-stacked batches (0 = off): the caller presents a batch of images [N][row_period][W] (rows_valid image rows
   * followed by row_period - rows_valid ZERO rows, which double as the convolution's zero padding between
   * neighbours) as ONE tall image (N = 1, OH = N * row_period).  Output rows with oy % row_period >= rows_valid are
   * computed but never stored, so the zero rows stay zero.  Removes the per-image rounding of the 16-row output
   * patches (55-row maps: 8 x 4 patches per column -> 28). */
  int32_t row_period, rows_valid;
  /* Merged sub-pixel phases (0 = off; LINEAR epilogue): the four output phases of a stride-2 transposed convolution
   * (ConvTranspose2d(k5, s2, p2, output_padding 1), e2vid/model/submodules.py:39-40) as ONE launch of Cout = 4 * phase_cout
   * GEMM columns over the union of their input offsets (weights of the taps a phase does not use are zero): column block
   * j = co / phase_cout is phase (py, px) = (j >> 1, j & 1) and is stored at pixel (oy*osy + py, ox*osx + px), channel
   * co % phase_cout (ooy / oox are ignored); bias has 4 * phase_cout entries.  phase_cout must be a multiple of 32. */
  int32_t phase_cout;
} essb_conv_tc;
int essb_conv_tc_run(const essb_conv_tc* d, void* stream);

/* tcgen05 weight gradient of a stride-1 gather-convolution (autograd of nn.Conv2d w.r.t. weight,
 * same definition as essb_wgrad_fp32): operands are the bf16 hi/lo planes of the (transformed,
 * concatenated) conv input A [N,H,W,a_ld] and of dY [N,H,W,g_ld] (g_ld a multiple of 64 >= Cout,
 * zero padded).  dw is written in the reference layout [Cout][Cin][ntaps]. */
typedef struct essb_wgrad_tc {
  const uint16_t* a_hi;
  const uint16_t* a_lo;
  const uint16_t* g_hi;
  const uint16_t* g_lo;
  float* dw;
  float* workspace;
  int64_t workspace_bytes;
  int32_t a_ld, Cin, g_ld, Cout;
  int32_t N, H, W;
  int32_t passes;
  int32_t ntaps;
  int32_t a_stride;      /* 0/1: stride-1 conv, A is [N,H,W,a_ld]; 2: stride-2 conv, A is [N,2H,2W,a_ld] (read through
                            its parity planes) and dy/dx are INPUT-pixel offsets (ky - pad) of dY's pixel (2y, 2x) */
  int8_t dy[ESSB_MAX_TAPS];
  int8_t dx[ESSB_MAX_TAPS];
} essb_wgrad_tc;
int64_t essb_wgrad_tc_workspace_bytes(const essb_wgrad_tc* d);
int essb_wgrad_tc_run(const essb_wgrad_tc* d, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ESS_B200_H */
