/*
 * b200fno — C ABI of the B200-native FNO forward / rollout engine.
 *
 * Drop-in boundary for the RealPDEBench FNO hot path.
 * The reference has no FFI for this path (it is stock PyTorch); each entry
 * point below names the reference code it replaces.  The host-side mirror of
 * the reference's plugin interface (realpdebench.model.fno / load_model /
 * eval.py rollout) lives in realpdebench_b200/ and binds these symbols with
 * ctypes; INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions
 *   - every data pointer is CALLER-OWNED DEVICE memory (fp32 unless stated);
 *     the library owns only the small constant tables inside the plan;
 *   - all work is enqueued on the caller's stream (a cudaStream_t passed as
 *     void*); no hidden synchronisation, no allocation after plan creation:
 *     every entry point that takes a stream is CUDA-graph capturable;
 *   - functions return 0 on success or a negative B200FNO_E* code; the message
 *     is available from b200fno_last_error() (thread-local).  No C++ exception
 *     crosses the ABI;
 *   - a plan is not thread-safe: one plan per (device, stream).
 */
#ifndef B200FNO_H
#define B200FNO_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200FNO_ABI_VERSION 1

#define B200FNO_OK 0
#define B200FNO_EINVAL (-1)   /* bad descriptor / argument */
#define B200FNO_ECUDA (-2)    /* CUDA runtime error (message has the cudaError string) */
#define B200FNO_ENODEV (-3)   /* no CUDA device / not an sm_100 part */
#define B200FNO_ESTATE (-4)   /* weights not packed, workspace not bound, ... */

/* impl selector for b200fno_plan_set_impl */
#define B200FNO_IMPL_AUTO 0   /* tensor-core path where the shape allows, else SIMT */
#define B200FNO_IMPL_SIMT 1   /* fp32 FFMA kernels (every shape) */
#define B200FNO_IMPL_TC 2     /* tcgen05 3xTF32 kernels (width 64 only); error otherwise */

/* compute mode for b200fno_plan_set_compute */
#define B200FNO_COMPUTE_F32 0   /* fp32 semantics (3xTF32 on the tensor cores): 1e-5 parity with the fp32 reference */
#define B200FNO_COMPUTE_BF16 1  /* torch.autocast(bfloat16) semantics of the reference: the operands of every
                                 * Linear / Conv (fc0, the 1x1 convolutions, fc1, fc2) are cast to bf16 and multiplied
                                 * in ONE tensor-core pass with fp32 accumulation; the FFT stages and the per-mode
                                 * mixing stay fp32, as do BatchNorm and the tensors in HBM.  1e-2 parity. */

typedef struct b200fno_plan b200fno_plan_t;

/*
 * Model + problem description.  Mirrors the constructor arguments of
 * FNO3d(modes1, modes2, modes3, n_layers, width, shape_in, shape_out)
 * (reference realpdebench/model/fno.py:67-103, called from
 * model/load_model.py:12-22) plus the batch size the workspace is sized for.
 *
 * ndim = 3: FNO3d as shipped; FFT axes (T,H,W), lift features c_in + 3 grid
 *           coordinates, projection features c_out * (t_out / t_in).
 * ndim = 2: FNO-2D of SURVEY.md 8(c): frames folded into channels; FFT axes
 *           (H,W); lift features t_in*c_in + 2; projection t_out*c_out;
 *           modes1 is ignored (YAML modes2,modes3 -> 2-D modes).
 */
typedef struct b200fno_desc {
  int32_t abi_version; /* B200FNO_ABI_VERSION */
  int32_t ndim;        /* 2 | 3 */
  int32_t max_batch;   /* workspace is sized for this many samples */
  int32_t t_in, t_out; /* frames in / out (shape_in[0], shape_out[0]) */
  int32_t h, w;        /* spatial grid (shape_in[1], shape_in[2]) */
  int32_t c_in, c_out; /* physical channels (shape_in[3], shape_out[3]) */
  int32_t width;       /* hidden channels */
  int32_t n_layers;
  int32_t modes1, modes2, modes3;
  int32_t padding;     /* fno.py:87 -> 6 */
  int32_t proj_hidden; /* fno.py:102 -> 128 */
  float bn_eps;        /* nn.BatchNorm3d default 1e-5 */
} b200fno_desc_t;

/*
 * Device pointers to the parameters in the REFERENCE state_dict layout
 * (checkpoint row).  Arrays are host arrays of device
 * pointers with one entry per layer (spec_w: n_layers * ncorner, corner-major
 * inside a layer: weights1..4, fno.py:31-38; ncorner = 4 for ndim 3, 2 for 2).
 * Spectral weights are complex64 [Ci][Co][m1][m2][m3] = interleaved (re,im) floats.
 */
typedef struct b200fno_weights {
  const float* fc0_w;             /* [width][lift_features]  nn.Linear weight, fno.py:89 */
  const float* fc0_b;             /* [width] */
  const float* const* spec_w;     /* complex64, see above */
  const float* const* conv_w;     /* [width][width] (1x1x1 kernel squeezed), fno.py:99 */
  const float* const* conv_b;     /* [width] */
  const float* const* bn_weight;  /* [width] fno.py:100 */
  const float* const* bn_bias;
  const float* const* bn_mean;    /* running_mean */
  const float* const* bn_var;     /* running_var */
  const float* fc1_w;             /* [proj_hidden][width] fno.py:102 */
  const float* fc1_b;
  const float* fc2_w;             /* [proj_features][proj_hidden] fno.py:103 */
  const float* fc2_b;
} b200fno_weights_t;

/* Thread-local message of the last failing call on this thread. */
const char* b200fno_last_error(void);
int b200fno_abi_version(void);

/* ---- plan life cycle ---------------------------------------------------- */
/* Replaces FNO3d.__init__'s shape bookkeeping (fno.py:78-87): validates the
 * descriptor, builds the truncated-DFT tables on the current device. */
int b200fno_plan_create(const b200fno_desc_t* desc, b200fno_plan_t** out);
int b200fno_plan_destroy(b200fno_plan_t* plan);
int b200fno_plan_set_impl(b200fno_plan_t* plan, int impl);
/* Which implementation the layer kernels resolve to (B200FNO_IMPL_SIMT|TC). */
int b200fno_plan_get_impl(const b200fno_plan_t* plan);
/* Arithmetic of the Linear / Conv products (B200FNO_COMPUTE_*).  May be changed at any time; a change invalidates the
 * packed weights (call b200fno_pack_weights again).  The training entry points honour it too: both operands of every
 * Linear / Conv GEMM of the forward AND the backward pass are rounded to bf16 (autocast's backward), fp32 accumulation;
 * the arithmetic of the mode, not its speed - the training kernels are FFMA kernels either way. */
int b200fno_plan_set_compute(b200fno_plan_t* plan, int compute);
int b200fno_plan_get_compute(const b200fno_plan_t* plan);

/* Bytes the caller must provide (activation ping-pong + spectral scratch). */
size_t b200fno_plan_workspace_bytes(const b200fno_plan_t* plan);
/* Bytes of the packed (engine-layout) copy of the parameters. */
size_t b200fno_plan_packed_bytes(const b200fno_plan_t* plan);
int b200fno_plan_bind(b200fno_plan_t* plan, void* workspace, size_t workspace_bytes,
                      void* packed, size_t packed_bytes);

/* Re-pack parameters from the reference layout into `packed` (device kernels
 * on `stream`): transposes, zero-pads channels to a multiple of 4, folds
 * conv bias + eval-mode BatchNorm (fno.py:115-117) into one per-channel
 * affine, resolves overlapping spectral corners the way the successive
 * assignments fno.py:53-60 do (later corner wins). */
int b200fno_pack_weights(b200fno_plan_t* plan, const b200fno_weights_t* w, void* stream);

/* ---- the hot path ------------------------------------------------------- */
/* FNO3d.forward (fno.py:105-129), eval-mode BatchNorm.
 *   x [batch][t_in][h][w][c_in]  ->  y [batch][t_out][h][w][c_out]  */
int b200fno_forward(b200fno_plan_t* plan, int32_t batch, const float* x, float* y, void* stream);

/* The autoregressive loop eval.py:313-321 for one batch, with the
 * de-normalise / concat-parameters / re-normalise glue (eval.py:315-318,
 * data_normalizer.py:50-62) folded into the projection epilogue as the
 * per-channel affine  p' = p * affine_a[c] + affine_b[c]  (SURVEY F6).
 *   x0    [batch][t_in][h][w][c_in]   normalised input (output of preprocess)
 *   pred  [batch][n_steps*t_out][h][w][c_out]  = torch.cat(preds[1:],1)[..., :c_out]
 *   state [2][batch][t_in][h][w][c_in] scratch for the fed-back inputs when c_in > c_out: parameter channels
 *         c_out..c_in-1 are carried over from x0 (eval.py:317).  May be NULL when n_steps == 1 or c_in == c_out:
 *         without parameter channels step i + 1 reads its input straight from the prediction slice step i wrote
 *         (the same tensor in eval.py:315-319), no state buffer and no second store exist.
 * n_steps > 1 requires t_out == t_in. */
int b200fno_rollout(b200fno_plan_t* plan, int32_t batch, const float* x0, const float* affine_a,
                    const float* affine_b, int32_t n_steps, float* state, float* pred, void* stream);

/* SpectralConv3d.forward (fno.py:45-64) / its 2-D analogue as a stand-alone
 * operator in the REFERENCE tensor layout (channels first):
 *   x [batch][ci][t][h][w] -> y [batch][co][t][h][w]       (ndim 3)
 *   x [batch][ci][h][w]    -> y [batch][co][h][w]          (ndim 2)
 * weights: ncorner device pointers, complex64 [ci][co][m1][m2][m3].
 * workspace: b200fno_spectral_workspace_bytes(...) bytes of device scratch.
 * The constant DFT tables of a geometry are built on the first call (one allocation + synchronous upload) and cached
 * per device; every later call of that geometry only enqueues kernels on `stream` (no allocation, no synchronisation,
 * capturable).  Width 64 runs the forward transforms on the tensor cores.  b200fno_spectral_cache_clear() frees the
 * cached tables (synchronise the streams that used them first). */
void b200fno_spectral_cache_clear(void);
size_t b200fno_spectral_workspace_bytes(int32_t ndim, int32_t batch, int32_t ci, int32_t co, int32_t t,
                                        int32_t h, int32_t w, int32_t m1, int32_t m2, int32_t m3);
int b200fno_spectral_conv(int32_t ndim, int32_t batch, int32_t ci, int32_t co, int32_t t, int32_t h,
                          int32_t w, int32_t m1, int32_t m2, int32_t m3, const float* const* weights,
                          const float* x, float* y, void* workspace, size_t workspace_bytes, void* stream);

/* ---- training path ------------------------------------------------------- */
/* Gradient outputs in the REFERENCE parameter layout (what torch's optimiser and
 * DDP see): same fields as b200fno_weights_t without the BatchNorm buffers.
 * Every non-NULL tensor is OVERWRITTEN with dL/dparam (not accumulated).
 * spec_w gradients are complex64 [ci][co][m1][m2][m3] interleaved (d/dRe, d/dIm)
 * = torch's .grad of a complex parameter viewed with view_as_real.  fc0_b is
 * only written together with fc0_w. */
typedef struct b200fno_grads {
  float* fc0_w;
  float* fc0_b;
  float* const* spec_w;    /* n_layers * ncorner */
  float* const* conv_w;    /* n_layers */
  float* const* conv_b;
  float* const* bn_weight;
  float* const* bn_bias;
  float* fc1_w;
  float* fc1_b;
  float* fc2_w;
  float* fc2_b;
  float* x;                /* optional: gradient w.r.t. the input field, layout of x [B][t_in][h][w][c_in]; NULL = not wanted */
} b200fno_grads_t;

/* Device bytes the training path needs on top of the plan workspace: the saved
 * layer inputs / pre-BatchNorm sums of one forward (2*n_layers+1 activations),
 * two gradient activations and the projection / weight-gradient scratch. */
size_t b200fno_train_workspace_bytes(const b200fno_plan_t* plan);
int b200fno_train_bind(b200fno_plan_t* plan, void* workspace, size_t workspace_bytes);

/* FNO3d.forward in .train() mode (fno.py:105-129 as called by train.py:328):
 * BatchNorm uses the batch statistics of the PADDED tensor (fno.py:111,117).
 * bn_running_mean / bn_running_var: n_layers device pointers updated in place
 * like nn.BatchNorm3d (running = (1-momentum)*running + momentum*batch stat,
 * unbiased variance); either array may be NULL (track_running_stats off).
 * Keeps the activations needed by b200fno_train_backward in the training
 * workspace (one forward outstanding per plan). */
int b200fno_train_forward(b200fno_plan_t* plan, int32_t batch, const float* x, float* y,
                          float* const* bn_running_mean, float* const* bn_running_var, float momentum,
                          void* stream);

/* Backward of the last b200fno_train_forward (autograd of fno.py:105-129 as
 * driven by loss.backward(), train.py:329): dy = dL/dy [batch][t_out][h][w][c_out]
 * -> parameter gradients.  The gradient w.r.t. the input x is not produced
 * (train.py never needs it).
 * grads_ready: NULL, or n_layers+1 cudaEvent_t handles (as void*, entries may be
 * NULL) recorded on `stream` as soon as a group of gradients is final - entry
 * n_layers: fc1/fc2; entry l: spec_w, conv_*, bn_* of layer l (recorded from the
 * last layer to the first); fc0 is final when the call's work completes.  A
 * data-parallel caller waits on them from a second stream to overlap the
 * gradient all-reduce with the rest of the backward pass. */
int b200fno_train_backward(b200fno_plan_t* plan, int32_t batch, const float* x, const float* dy,
                           const b200fno_grads_t* grads, void* const* grads_ready, void* stream);

/* One Adam update of a flat fp32 tensor in a single pass (torch.optim.Adam as built at train.py:290: betas
 * (0.9, 0.999), eps 1e-8, no weight decay, no amsgrad; `step` counts from 1).  Complex parameters are passed as
 * their real view, 2n floats.  param, exp_avg, exp_avg_sq are updated in place. */
int b200fno_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                      float beta1, float beta2, float eps, int64_t step, void* stream);

/* ---- evaluation metrics --------------------------------------------------- */
/* eval_metrics (reference realpdebench/utils/metrics.py:24-131) for ONE chunk of `b` samples (the reference's
 * batch_size loop, :41, stays with the caller): pred, target [b][t][h][w][channels] device fp32, the first `c` channels
 * are evaluated.  out13 (device, 13 floats) = rmse, mae, rel_l2_error, r2, ke_error, f_error, low_f_error, mid_f_error,
 * high_f_error, rel_low_f_error, rel_mid_f_error, rel_high_f_error, freq_error (order of :126-131).  The radial-bin
 * spectra only use wavenumbers below min(t,h,w)/2 per axis (:75-81), computed as truncated DFTs on the device. */
size_t b200fno_metrics_workspace_bytes(int32_t b, int32_t t, int32_t h, int32_t w, int32_t channels, int32_t c);
int b200fno_eval_metrics(const float* pred, const float* target, int32_t b, int32_t t, int32_t h, int32_t w,
                         int32_t channels, int32_t c, void* workspace, size_t workspace_bytes, float* out13,
                         void* stream);

/* ---- introspection used by bench.py / tests ------------------------------ */
/* Kernels launched by this library on this thread since the last reset. */
/* Diagnostics.  With B200FNO_DEBUG_FINITE set in the environment, b200fno_train_backward scans the output of every
 * kernel it launches for non-finite values (on the device, in stream order, no host synchronisation).  Returns the id
 * of the first stage of the last backward that produced one (100 * (layer + 1) + index inside the layer loop; 1..9
 * projection backward; 9000+ lift backward), 0 if all were finite, < 0 if the facility is off.  Synchronises. */
int b200fno_debug_first_nonfinite(b200fno_plan_t* plan);
int64_t b200fno_launch_count(void);
void b200fno_launch_count_reset(void);
/* Per-stage device timing: when enabled, every stage launch of forward/rollout is
 * bracketed by CUDA events on the caller's stream (not graph-capturable while on).
 * collect() waits for the recorded events and returns, per stage, the summed
 * milliseconds and the number of launches since enable()/the last collect().
 * Stage order: lift, fwdW, fwdH, fwdT, modes, invT, invH, layer, proj. */
#define B200FNO_NUM_STAGES 9
int b200fno_timing_enable(b200fno_plan_t* plan, int on);
int b200fno_timing_collect(b200fno_plan_t* plan, double* ms /*[9]*/, int64_t* count /*[9]*/);
/* Tensor-core primitive self-test: one CTA computes D[128][N] = A * B^T with tcgen05.mma kind::tf32.
 *   mode_a: 0 A[128][K] staged by threads (manual 128B swizzle), 1 same via TMA, 2 A given as [K][128]
 *           (MN-major) via TMA, 3 A[128][K] placed in TMEM (tcgen05.st) -- the .ts MMA form
 *   mode_b: 0 B[N][K] manual, 1 B[N][K] via TMA, 2 B given as [K][N] (MN-major) via TMA
 *   out_tma: 0 D written with st.global, 1 through a swizzled staging tile + TMA store
 * N in {32,64,128}, K in {32,64,96,128}; all pointers device fp32. */
int b200fno_selftest_umma(int32_t mode_a, int32_t mode_b, int32_t out_tma, int32_t N, int32_t K, const float* A,
                          const float* B, float* D, void* stream);
/* MMA issue-rate probe: `iters` back-to-back tcgen05.mma (M=128, N, K=8, tf32) from one CTA, cycling over
 * `nacc` accumulators; writes the elapsed SM cycles to out_cycles_dev[0] (device int64).
 * a_in_tmem: 1 = .ts form, 0 = .ss form; accumulate: 0 overwrites D (no read-modify-write). */
int b200fno_selftest_mma_rate(int32_t N, int32_t iters, int32_t a_in_tmem, int32_t nacc, int32_t accumulate,
                              long long* out_cycles_dev, void* stream);
/* Host copy of truncated-DFT table `which` (0 fwdW, 1 fwdH, 2 fwdT, 3 invT, 4 invH,
 * 5 invW) for a transformed grid (t,h,w) -- the values the kernels multiply by.
 * Needs no device.  Writes at most `cap` floats to `out`, the row pitch to *ld and
 * the kept T / H frequency indices to freqs_t / freqs_h (each sized >= t / h;
 * may be NULL).  Returns the table length in floats or a negative error code. */
int64_t b200fno_host_table(int32_t ndim, int32_t t, int32_t h, int32_t w, int32_t m1, int32_t m2, int32_t m3,
                           int32_t which, float* out, int64_t cap, int32_t* ld, int32_t* freqs_t, int32_t* freqs_h);
/* The same for one W-mode slice [kw0, kw0 + m3) - the tables a plan builds per slice when modes3 in (32, 64] runs as two
 * slices on the tensor-core kernels (the W tables carry the frequency offset, the H / T tables are unchanged). */
int64_t b200fno_host_table_slice(int32_t ndim, int32_t t, int32_t h, int32_t w, int32_t m1, int32_t m2, int32_t m3,
                                 int32_t kw0, int32_t which, float* out, int64_t cap, int32_t* ld, int32_t* freqs_t,
                                 int32_t* freqs_h);
/* Which kernel family stage `stage` (order of b200fno_timing_collect) resolves to after b200fno_plan_bind:
 * 1 = tcgen05 tensor-core kernel, 0 = fp32 FFMA kernel, negative = error.  Lets tests and bench.py state
 * which path produced a number. */
int b200fno_plan_stage_impl(const b200fno_plan_t* plan, int32_t stage);
/* Algorithmic HBM bytes of one forward at `batch` (SURVEY.md 8d formula). */
double b200fno_algorithmic_bytes(const b200fno_plan_t* plan, int32_t batch);

#ifdef __cplusplus
}
#endif
#endif /* B200FNO_H */
