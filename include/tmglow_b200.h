/*
 * tmglow_b200.h -- C ABI of libtmglow_b200.so: the B200 (sm_100a) implementation of the
 * TM-Glow conditional-flow hot path of zabaras/deep-turbulence.
 *
 * The reference has no FFI: its "operator API" for this path is the TMGlow nn.Module
 * (tmglow/nn/tmGlow.py:305-509).  Each entry point below names the reference interface it
 * replaces (file:line relative to /root/reference/tmglow).  The Python host side
 * (deep-turbulence_b200/tmglow_b200) binds these with ctypes and keeps the reference's module
 * API; INTEGRATION.md shows the binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success or a negative tmg_status; tmg_last_error() gives text
 *   - all pointers are DEVICE pointers to fp32 unless stated otherwise; the library never
 *     allocates on the hot path: the caller passes a workspace of tmg_workspace_bytes()
 *   - `stream` is a cudaStream_t passed as void* (no CUDA headers needed to bind); the call is
 *     asynchronous and stream-ordered; buffers are borrowed until the work completes
 *   - user-facing fields (x, y, z, eps) are NCHW like the reference; LSTM states are
 *     channels-last ([B,H,W,rec], i.e. torch.channels_last memory of a [B,rec,H,W] tensor)
 *   - there is NO CPU fallback: without a CUDA device every compute call fails with TMG_ERR_CUDA
 */
#ifndef TMGLOW_B200_H
#define TMGLOW_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TMG_MAX_LEVELS 6

typedef enum tmg_status {
  TMG_OK = 0,
  TMG_ERR_BAD_CONFIG = -1,   /* reference: constructor asserts */
  TMG_ERR_BAD_SHAPE = -2,    /* reference: flowUtils.py:112,136 / tmGlow.py:249,286 asserts */
  TMG_ERR_NULL = -3,
  TMG_ERR_WORKSPACE = -4,    /* workspace too small */
  TMG_ERR_CUDA = -5,         /* CUDA runtime error (text in tmg_last_error) */
  TMG_ERR_UNSUPPORTED = -6,  /* shape/feature outside what the kernels implement */
  TMG_ERR_NOT_READY = -7     /* tmg_model_refresh() has not been called */
} tmg_status;

/* Constructor arguments of TMGlow (nn/tmGlow.py:336-338) that shape the path. */
typedef struct tmg_config {
  int32_t in_features;
  int32_t out_features;
  int32_t n_levels;                     /* len(enc_blocks) == len(glow_blocks) */
  int32_t enc_blocks[TMG_MAX_LEVELS];
  int32_t glow_blocks[TMG_MAX_LEVELS];
  int32_t cond_features;
  int32_t cglow_upscale;
  int32_t growth_rate;
  int32_t init_features;
  int32_t rec_features;
} tmg_config;

typedef struct tmg_model tmg_model;     /* opaque; owns packed weights on ONE device */

/* flags for the whole-path calls */
#define TMG_FLAG_BN_TRAIN   1u          /* encoder BatchNorm uses batch statistics and updates the
                                           running stats in the parameter buffer (nn.Module.train()) */
#define TMG_FLAG_SHARED_X   2u          /* x holds ONE low-fidelity input [1,nic,h,w] shared by all B samples
                                           (uncertainty quantification: many stochastic samples of one input,
                                           trainFlowParallel.py:345-358): the encoder runs once and the flow steps
                                           reuse the conditioning maps.  Same results as x.expand(B,...) */

/* ---- library ------------------------------------------------------------------------- */
int         tmg_version(void);
const char* tmg_last_error(void);                 /* thread-local text of the last failure */
int         tmg_device_count(void);               /* 0 when no CUDA device is visible */

/* ---- model handle: replaces TMGlow.__init__ (nn/tmGlow.py:336-376) --------------------- */
int  tmg_model_create(const tmg_config* cfg, tmg_model** out);   /* on the current CUDA device */
void tmg_model_destroy(tmg_model* m);

/* Which kernels run the heavy 3x3 convolutions (ConvLSTM gate/out convs, Conv2dZeros):
 *   TMG_PREC_FP32   exact fp32 FMA on CUDA cores (default)
 *   TMG_PREC_TF32X3 tcgen05 tensor cores, 3xTF32 operand split: fp32-grade accuracy
 *   TMG_PREC_TF32   tcgen05 tensor cores, single-pass TF32 (what stock PyTorch/cuDNN does on
 *                   GPU by default); looser, separately stated tolerance */
#define TMG_PREC_FP32   0
#define TMG_PREC_TF32X3 1
#define TMG_PREC_TF32   2
/*   TMG_PREC_F16X3  tcgen05, fp16 operands split into hi + lo halves with power-of-two weight scaling
 *                   (three MMAs, fp32 accumulation): fp32-grade accuracy at half the tensor work and half the
 *                   shared memory of the 3xTF32 split; persistent fused flow-step kernel (flow_step_f16.cu)
 *   TMG_PREC_F16    tcgen05, single-pass fp16 operands (11-bit mantissa like TF32); looser tolerance */
#define TMG_PREC_F16X3  3
#define TMG_PREC_F16    4
int tmg_model_set_precision(tmg_model* m, int mode);
int tmg_model_get_precision(const tmg_model* m);

/* Sticky overflow flag of the fp16-operand modes (TMG_PREC_F16X3 / TMG_PREC_F16): activations are staged as fp16 hi/lo halves
 * and clamped to +-6e4; a clamp that changed a value (never with well-conditioned weights; the reference would carry the value
 * in fp32) raises the flag.  Returns 1 when raised (text in tmg_last_error), 0 otherwise, < 0 on error; `clear` resets it.
 * Synchronises `stream`. */
int tmg_model_overflow(tmg_model* m, int clear, void* stream);

/* Parameter table: the flat fp32 parameter buffer holds every floating-point state_dict entry
 * (parameters AND buffers, reference names) back to back in the order reported here. */
int64_t     tmg_model_param_entries(const tmg_model* m);
const char* tmg_model_param_name(const tmg_model* m, int64_t i);
int64_t     tmg_model_param_offset(const tmg_model* m, int64_t i);   /* in floats */
int64_t     tmg_model_param_numel(const tmg_model* m, int64_t i);
int         tmg_model_param_shape(const tmg_model* m, int64_t i, int64_t dims[4]);   /* returns ndim */
int64_t     tmg_model_param_total(const tmg_model* m);               /* floats in the flat buffer */

/* Re-derive the kernel-side weights (tap-major conv weights, W and W^-1 of every
 * InvertibleConv1x1LU -- glowConv.py:151-174 --, ActNorm/1x1 log-det sums, Conv2dZeros gains,
 * folded eval-mode BatchNorm) from the flat parameter buffer.  Call after every parameter
 * change.  `params` stays borrowed: BN-train calls write the running statistics back into it. */
int tmg_model_refresh(tmg_model* m, float* params, void* stream);

/* Copies of derived weights for tests (glowConv.py:151-174). dst: [C*C] device floats. */
int tmg_model_get_conv1x1(const tmg_model* m, int level, int step /*1-based*/, int inverse,
                          float* dst, void* stream);

size_t tmg_workspace_bytes(const tmg_model* m, int B, int h, int w);   /* (h,w) = LF input size */

/* ---- whole-path operators -------------------------------------------------------------- */
/* TMGlow.reconstruct / TMGlow.sample (nn/tmGlow.py:417-467) + LSTMCFlowDecoder.reverse
 * (:269-303).  eps[0..n_levels-1] = split noise [B,C_l/2,H_l,W_l] NCHW, eps[n_levels] = top
 * noise [B,Cz,H_L,W_L] NCHW (TMGlow.sample = this call with eps drawn by the caller in the
 * reference's RNG order).  h_in/c_in: n_levels pointers or NULL (zero states, convLSTM.py:66-68).
 * Outputs: y [B,out_features,H,W] NCHW, log_det [B] (excludes the top prior, like the
 * reference), h_out/c_out channels-last. */
int tmg_reconstruct(tmg_model* m, int B, int h, int w, const float* x,
                    const float* const* h_in, const float* const* c_in,
                    const float* const* eps,
                    float* y, float* log_det, float* const* h_out, float* const* c_out,
                    void* workspace, size_t workspace_bytes, uint32_t flags, void* stream);

/* Training forward: tmg_reconstruct that also records, per flow step, the step input and (fused fp16 step kernel) the
 * coupling-network intermediates relu(d1), relu(d2), h into `tape` (tmg_tape_bytes) for tmg_reconstruct_backward.  Runs in
 * the model's precision mode. */
size_t tmg_tape_bytes(const tmg_model* m, int B, int h, int w);
int tmg_reconstruct_train(tmg_model* m, int B, int h, int w, const float* x,
                          const float* const* h_in, const float* const* c_in, const float* const* eps,
                          float* y, float* log_det, float* const* h_out, float* const* c_out,
                          void* tape, size_t tape_bytes, void* workspace, size_t workspace_bytes, uint32_t flags,
                          void* stream);

/* Backward of TMGlow.sample / reconstruct (what loss.backward() does through the reference's autograd graph,
 * nn/trainFlowParallel.py:259-277): given g_y [B,out,H,W], g_log_det [B] and the gradients w.r.t. the returned LSTM
 * states (channels-last, entries may be NULL), ACCUMULATES the gradient of every flow (decoder) parameter into `grads`
 * (flat, laid out like the parameter buffer) and returns the gradients w.r.t. the incoming states g_h_in / g_c_in
 * (written where a state was passed in).  Coupling-network intermediates come from the tape (LSTM steps: recomputed);
 * in the f16 modes the weight and data gradients of the flow convolutions run on tcgen05 (fp16 hi/lo split, fp32-grade),
 * in the fp32 mode on the exact-fp32 CUDA-core kernels; the encoder adjoints are exact fp32.  Deterministic.  The gradients
 * of the LU-parameterised 1x1 convolutions are left in accumulated form: call tmg_backward_finalize once per optimizer step. */
size_t tmg_reconstruct_backward_workspace_bytes(const tmg_model* m, int B, int h, int w);
int tmg_reconstruct_backward(tmg_model* m, int B, int h, int w, const float* x,
                             const float* const* h_in, const float* const* c_in, const float* const* eps,
                             const void* tape, const float* g_y, const float* g_log_det,
                             const float* const* g_h_out, const float* const* g_c_out,
                             float* const* g_h_in, float* const* g_c_in, float* grads,
                             void* workspace, size_t workspace_bytes, uint32_t flags, void* stream);

/* tmg_reconstruct_backward leaves the gradients of the LU-parameterised 1x1 convolutions (l, u, log_s, glowConv.py:151-174)
 * and the log-det terms of log_s / ActNorm weights in an accumulated, linear form (dW in the gradient slot of the buffer
 * `p`, H_l*W_l*sum(g_log_det) in the slot of `sign_s`): call this ONCE after the last backward call of an optimizer step
 * (before reading `grads`) to turn them into the parameter gradients for all flow steps in one launch; the slots are
 * cleared.  tmg_flow_step_backward is not deferred. */
int tmg_backward_finalize(tmg_model* m, float* grads, void* stream);

/* One BPTT block of the reference trainer in ONE call (the `tback` sample() calls of nn/trainFlowParallel.py:248-277 and the
 * loss.backward() through all of them, :277).  Inside a flow level only the LSTM step couples the time steps, so the block is
 * evaluated level by level: the level's LSTM step T times in sequence on batch B, its plain steps / split prior / squeeze
 * ONCE on batch T*B.  Same results as T chained tmg_reconstruct_train calls (per-sample arithmetic is unchanged), about a
 * fifth of the launches, ten times the work per launch.  All tensors are TIME-MAJOR: x [T,B,nic,h,w], eps[l] [T,B,...],
 * y [T,B,out,H,W], log_det [T,B]; h_in / c_in: states before the first time step (NULL = zeros), h_out / c_out: states after
 * the last one (channels-last).  The encoder runs per time step (BatchNorm batch statistics per sample() call).
 * Backward: g_y [T,B,out,H,W], g_log_det [T,B], g_h_out / g_c_out w.r.t. the final states (entries may be NULL); ACCUMULATES
 * into `grads` like tmg_reconstruct_backward (call tmg_backward_finalize once per optimizer step) and returns g_h_in / g_c_in. */
size_t tmg_bptt_tape_bytes(const tmg_model* m, int T, int B, int h, int w);
size_t tmg_bptt_workspace_bytes(const tmg_model* m, int T, int B, int h, int w);
int tmg_bptt_forward(tmg_model* m, int T, int B, int h, int w, const float* x, const float* const* h_in,
                     const float* const* c_in, const float* const* eps, float* y, float* log_det, float* const* h_out,
                     float* const* c_out, void* tape, size_t tape_bytes, void* workspace, size_t workspace_bytes,
                     uint32_t flags, void* stream);
int tmg_bptt_backward(tmg_model* m, int T, int B, int h, int w, const float* x, const float* const* h_in,
                      const float* const* c_in, const float* const* eps, const void* tape, const float* g_y,
                      const float* g_log_det, const float* const* g_h_out, const float* const* g_c_out,
                      float* const* g_h_in, float* const* g_c_in, float* grads, void* workspace, size_t workspace_bytes,
                      uint32_t flags, void* stream);

/* With TMG_BWD_GRAPH=1 in the environment tmg_reconstruct_backward replays a CUDA graph of its ~1 400 launches when it is
 * called again with the same buffers (same pointers, shapes, precision): a key is run eagerly the first time it is seen,
 * captured the second time and replayed from then on.  Off by default: it pays only for callers with stable buffer
 * addresses.  Counters for measurement: graphs held, replays, eager runs. */
int tmg_backward_graph_stats(const tmg_model* m, int64_t* graphs, int64_t* replays, int64_t* eager);

/* TMGlow.forward (nn/tmGlow.py:378-414) + LSTMCFlowDecoder.forward (:231-267).
 * Outputs: z [B,Cz,H_L,W_L] NCHW, logp [B] (= log prior + sum of log-dets), states, and when
 * eps_out != NULL the n_levels+1 noise tensors (return_eps=True). */
int tmg_forward(tmg_model* m, int B, int h, int w, const float* x, const float* y,
                const float* const* h_in, const float* const* c_in,
                float* z, float* logp, float* const* h_out, float* const* c_out,
                float* const* eps_out,
                void* workspace, size_t workspace_bytes, uint32_t flags, void* stream);

/* Encoder.forward (nn/tmGlow.py:104-129): c_out[l] [B,cond,H_l,W_l] and z_out [B,2*Cz,H_L,W_L],
 * both written NCHW for inspection/tests. */
int tmg_encoder_forward(tmg_model* m, int B, int h, int w, const float* x,
                        float* const* c_out, float* z_out,
                        void* workspace, size_t workspace_bytes, uint32_t flags, void* stream);

/* ---- single operators (parity tests, and building blocks for callers) ------------------ */
/* CheckerSqueeze.forward / .reverse (flowUtils.py:99-145), NCHW in and out, bit-exact. */
int tmg_squeeze_forward(const float* x, float* y, int B, int C, int H, int W, void* stream);
int tmg_squeeze_reverse(const float* y, float* x, int B, int C4, int H2, int W2, void* stream);

/* One flow step of block `level` (1-based `step`), NCHW in/out:
 * UnNormedAffineCouplingBlock / AffineCouplingBlock / LSTMCouplingBlock .forward/.reverse
 * (flowLSTMBlock.py:53-86,116-146,180-218).  x,out: [B,C_l,Hl,Wl]; cond: [B,cond,Hl,Wl];
 * logdet: [B]; states channels-last, used only by the LSTM step (may be NULL). */
int tmg_flow_step(tmg_model* m, int level, int step, int reverse, int B, int Hl, int Wl,
                  const float* x, const float* cond, const float* h_in, const float* c_in,
                  float* out, float* logdet, float* h_out, float* c_out,
                  void* workspace, size_t workspace_bytes, void* stream);

/* Split.forward / Split.reverse (flowUtils.py:292-335), NCHW.
 * forward: z [B,C,Hl,Wl] -> z1 [B,C/2,..], logp [B], eps [B,C/2,..] (eps may be NULL)
 * reverse: z1, eps -> z [B,C,..], logp [B] */
int tmg_split_forward(tmg_model* m, int level, int B, int Hl, int Wl, const float* z,
                      float* z1, float* logp, float* eps,
                      void* workspace, size_t workspace_bytes, void* stream);
int tmg_split_reverse(tmg_model* m, int level, int B, int Hl, int Wl, const float* z1,
                      const float* eps, float* z, float* logp,
                      void* workspace, size_t workspace_bytes, void* stream);

/* nn.Conv2d(Cin, Cout, 3, stride=1, padding=1) on NHWC tensors with OIHW weights (+ optional input
 * ReLU, replicate padding as in Conv2dZeros flowUtils.py:246, bias, activation 0 none / 1 ReLU /
 * 2 hardtanh(-2, ln5)); `mode` is a TMG_PREC_* value.  Exposes the convolution kernels behind every
 * conv of the path (convLSTM.py:44,129; denseBlock.py:136; flowUtils.py:229) for isolated tests. */
size_t tmg_conv3x3_workspace_bytes(int Cin, int Cout);
int tmg_conv3x3(int mode, const float* x_nhwc, int B, int H, int W, int Cin, const float* w_oihw,
                const float* bias, int Cout, int relu_in, int pad_replicate, int act, float* out_nhwc,
                void* workspace, size_t workspace_bytes, void* stream);

/* Backward of the same convolution (autograd of F.conv2d as used by every conv of the path; first building block of
 * training through TMGlow.sample, nn/trainFlowParallel.py:259-277): data gradient gx [B,H,W,Cin] (gated by the input
 * ReLU when relu_in), weight gradient gw (OIHW) and bias gradient gbias [Cout]; any of the three may be NULL.
 * Exact fp32, deterministic (fixed-order reductions, no atomics). */
size_t tmg_conv3x3_backward_workspace_bytes(int B, int H, int W, int Cin, int Cout);
int tmg_conv3x3_backward(const float* x_nhwc, int B, int H, int W, int Cin, const float* w_oihw, int Cout,
                         int relu_in, int pad_replicate, const float* gout_nhwc, float* gx_nhwc, float* gw_oihw,
                         float* gbias, void* workspace, size_t workspace_bytes, void* stream);

/* Weight and bias gradient of the same convolution on the tensor cores (tcgen05 GEMM per filter tap with the pixels as the
 * contraction dimension, fp16 hi/lo operand split, fp32 accumulate): the kernel the training path uses in the f16x3 mode.
 * fp32-grade (hi*hi + lo*hi + hi*lo), deterministic.  gbias may be NULL. */
size_t tmg_conv3x3_wgrad_tc_workspace_bytes(int B, int H, int W, int Cin, int Cout);
int tmg_conv3x3_wgrad_tc(const float* x_nhwc, int B, int H, int W, int Cin, int Cout, int relu_in, int pad_replicate,
                         const float* gout_nhwc, float* gw_oihw, float* gbias, void* workspace, size_t workspace_bytes,
                         void* stream);

/* Backward of one REVERSE flow step (the direction training runs: TMGlow.sample -> loss.backward(),
 * nn/trainFlowParallel.py:259-277): UnNormedAffineCouplingBlock / AffineCouplingBlock / LSTMCouplingBlock .reverse
 * (flowLSTMBlock.py:71-86,132-146,200-218).  Given the gradients w.r.t. the step output g_out [B,C_l,Hl,Wl] (NCHW),
 * the per-sample log-det g_logdet [B] and, for the LSTM step, the returned states g_h_out / g_c_out (channels-last,
 * may be NULL), returns g_x, g_cond (NCHW), g_h_in / g_c_in (channels-last, may be NULL) and ACCUMULATES the gradient
 * of every parameter of the step into `grads`, a flat fp32 buffer laid out like the parameter buffer
 * (tmg_model_param_offset).  The forward is recomputed with the exact-fp32 kernels; deterministic. */
size_t tmg_flow_step_backward_workspace_bytes(tmg_model* m, int level, int B, int Hl, int Wl);
int tmg_flow_step_backward(tmg_model* m, int level, int step, int B, int Hl, int Wl, const float* x, const float* cond,
                           const float* h_in, const float* c_in, const float* g_out, const float* g_logdet,
                           const float* g_h_out, const float* g_c_out, float* g_x, float* g_cond, float* g_h_in,
                           float* g_c_in, float* grads, void* workspace, size_t workspace_bytes, void* stream);

/* TMGLowLoss.forward (nn/trainFlowParallel.py:121-153) with calcVPres / calcVDiv (:155-177), the residuals of
 * PhysConstrainedLES (pc/physicsConstrained.py:43-94) and the 3x3 filters of pc/grad1Filter.py, pc/grad2Filter.py, FUSED
 * WITH ITS BACKWARD: loss = beta*(vPres + vDiv + vL1 + vRMS) + mean(logp)/(ln2 * 3HW).
 *   y_pred, target [B,T,3,H,W] (channels u, v, p); logp: n_logp floats (only the mean is used; the reference passes [B,T]);
 *   target_rms [B,3,H,W]; out_mu, out_std [3] (model.out_mu / out_std, trainFlowParallel.py:119-120); dx, dy, beta =
 *   args.dx, args.dy, args.beta.
 * Outputs (device): loss [1]; terms [5] = vPres, vDiv, vL1, vRMS, neg_entropy (may be NULL); g_y = d loss / d y_pred
 * [B,T,3,H,W] and g_logp [n_logp] (each may be NULL).  Deterministic; H, W >= 3. */
size_t tmg_tmglow_loss_workspace_bytes(int B, int T, int H, int W);
int tmg_tmglow_loss(const float* y_pred, const float* logp, const float* target, const float* target_rms,
                    const float* out_mu, const float* out_std, int B, int T, int H, int W, int n_logp, double dx, double dy,
                    double beta, float* loss, float* terms, float* g_y, float* g_logp, void* workspace,
                    size_t workspace_bytes, void* stream);

/* Optimizer step of the reference trainer on the flat parameter buffer: torch.nn.utils.clip_grad_norm_ followed by
 * torch.optim.Adam(..., weight_decay, amsgrad).step() (nn/trainFlowParallel.py:290-291, main.py:78) as two launches with no host
 * synchronisation.  `mask` (0/1 per entry, may be NULL): entries with 0 -- the buffers that share the flat layout: permutations,
 * masks, BatchNorm running statistics -- are left untouched.  hyper (device, 8 floats): lr, beta1, beta2, eps, weight_decay,
 * max_norm (<= 0: no clipping), step counter (incremented by the call), amsgrad flag; out (device, 2 floats): gradient norm
 * before clipping, clip coefficient.  max_exp_avg_sq may be NULL when amsgrad is off.  Deterministic. */
size_t tmg_adam_workspace_bytes(void);
int tmg_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, float* max_exp_avg_sq, const float* mask,
                  int64_t n, float* hyper, float* out, void* workspace, size_t workspace_bytes, void* stream);

/* layout helpers for the LSTM states at the API boundary */
int tmg_nchw_to_nhwc(const float* src, float* dst, int B, int C, int H, int W, void* stream);
int tmg_nhwc_to_nchw(const float* src, float* dst, int B, int C, int H, int W, void* stream);

/* Optional per-kernel-class timing with CUDA events (used by bench.py for the roofline figures;
 * off by default).  tmg_profile_enable(1) starts recording on the calling thread;
 * tmg_profile_query synchronises the recorded events and returns, for class `tag`
 * (0..tmg_profile_classes()-1), the summed device time [ms], launch count, algorithmic FLOPs and
 * algorithmic HBM bytes; tmg_profile_enable(0) stops and clears. */
int         tmg_profile_enable(int on);
int         tmg_profile_classes(void);
const char* tmg_profile_class_name(int tag);
int         tmg_profile_query(int tag, double* ms, int64_t* launches, double* flops, double* bytes);

/* number of kernels this library launched (process-wide, all threads) since the last reset
 * (bench.py's "gpu_launches") */
int64_t tmg_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif /* TMGLOW_B200_H */
