// Shared declarations of libtmglow_b200 (sm_100a).  Internal layout: NHWC fp32 ("pixels x channels").
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/tmglow_b200.h"

namespace tmg {

constexpr float kLog2Pi = 1.8378770664093453f;   // GaussianDiag.Log2PI   flowUtils.py:155
constexpr float kLog5 = 1.6094379124341003f;     // clamp / hardtanh max  flowUtils.py:163,270
constexpr float kLog4 = 1.3862943611198906f;     // Conv2dZeros clamp max flowUtils.py:247
constexpr int kPixTile = 128;                    // pixels per CTA of the pointwise kernels
constexpr int kMaxC = 64;                        // widest flow state the pointwise kernels take

void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

#define TMG_CUDA_OK(expr)                                                              \
  do {                                                                                 \
    cudaError_t e__ = (expr);                                                          \
    if (e__ != cudaSuccess) {                                                          \
      tmg::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return TMG_ERR_CUDA;                                                             \
    }                                                                                  \
  } while (0)

#define TMG_LAUNCH_CHECK()                                                             \
  do {                                                                                 \
    ++tmg::g_launches;                                                                 \
    cudaError_t e__ = cudaGetLastError();                                              \
    if (e__ != cudaSuccess) {                                                          \
      tmg::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
      return TMG_ERR_CUDA;                                                             \
    }                                                                                  \
  } while (0)

#define TMG_TRY(expr)                 \
  do {                                \
    int s__ = (expr);                 \
    if (s__ != TMG_OK) return s__;    \
  } while (0)

// ---- optional per-kernel-class timing (bench.py's roofline figures): CUDA events around launches
enum ProfTag {
  PROF_CONV_GATE = 0,    // ConvLSTM gate conv            (convLSTM.py:44)
  PROF_CONV_OUT = 1,     // LSTM_out_conv                 (convLSTM.py:129)
  PROF_CONV_ZERO = 2,    // Conv2dZeros of coupling nets  (flowUtils.py:229)
  PROF_CONV_DENSE1 = 3,  // Cout=1 dense layers           (denseBlock.py:136)
  PROF_CONV_SPLIT = 4,   // split prior conv
  PROF_CONV_ENC = 5,     // encoder convs
  PROF_POINTWISE = 6,    // fused coupling + 1x1 + ActNorm + log-det
  PROF_LSTM_PW = 7,
  PROF_GAUSS = 8,
  PROF_PERMUTE = 9,
  PROF_MISC = 10,
  PROF_STEP_FUSED = 11,  // coupling net + coupling + 1x1 + ActNorm in one launch (coupling_tc.cu)
  PROF_LEVEL_RES = 12,   // all plain steps of a level in one launch, state resident on the SM (flow_level_f16.cu)
  PROF_NTAGS = 13
};
struct ProfScope {
  cudaStream_t st;
  int idx;
  ProfScope(cudaStream_t st, int tag, double flops, double bytes);
  ~ProfScope();
};

inline int cdiv(int a, int b) { return (a + b - 1) / b; }
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device) instead of on every launch: the call costs
// about as much CPU time as the launch itself, and a training step is ~15 000 launches
cudaError_t ensure_dyn_smem(const void* kernel, int bytes);
template <class K> inline cudaError_t ensure_dyn_smem_k(K kernel, int bytes) { return ensure_dyn_smem(reinterpret_cast<const void*>(kernel), bytes); }
#define TMG_SMEM_ATTR(...) TMG_CUDA_OK(tmg::ensure_dyn_smem_k(__VA_ARGS__))     // variadic: template commas in the kernel name
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ------------------------------------------------------------------ generic 3x3 convolution
// One virtual input = channel-concatenation of up to 3 NHWC sources (torch.cat(...,1) of the
// reference is never materialised).
struct ConvSrc {
  const float* p;
  int cstride;   // floats between pixels
  int coff;      // first channel inside the pixel
  int nch;       // channels taken
  int relu;      // apply ReLU while staging
  int bshared;   // 1: the tensor has batch 1 and is shared by every sample (one LF input, many samples)
};

struct ConvArgs {
  ConvSrc src[3];
  int nsrc;
  const float* bn_scale;   // optional per-(concatenated)-channel affine applied before ReLU
  const float* bn_shift;
  const float* w;          // tap-major weights [9][cin_w][cout_w]
  int cin_w, cout_w;
  const float* bias;       // [cout] or null
  const float* gain;       // scalar multiplier (exp(clamp(scale))) or null
  int act;                 // 0 none, 1 ReLU, 2 hardtanh(-2, ln5)
  float* out;
  int out_cstride, out_coff, cout;
  int B, Hin, Win, Hout, Wout, stride;
  int pad_replicate;       // 0: zero padding, 1: replicate padding (Conv2dZeros)
  // backward use (data gradient = the same kernel on flipped/transposed weights): gate the result with the sign of
  // the forward input (ReLU) and accumulate into the destination
  const float* mask;       // same layout as out (out_cstride, out_coff): result kept where mask > 0, or null
  int accum;               // 1: out += result
};
int launch_conv3x3(const ConvArgs& a, cudaStream_t st);

// Tensor-core (tcgen05, TF32 / 3xTF32) variant: stride 1 only, see conv3x3_tc.cu
struct TcConvArgs {
  ConvSrc src[3];
  int nsrc;
  int cin;                 // real input channels (K is walked in chunks of 16)
  const float* wpk;        // packed weights [chunk][tap][hi|lo][plane][npad][4]
  int npad;                // MMA N: multiple of 16, <= 256
  const float* bias;
  const float* gain;
  int act;
  float* out;
  int out_cstride, out_coff, cout;
  int B, H, W;
  int pad_replicate;
  int split3;              // 1: 3xTF32 (fp32-grade), 0: single-pass TF32
  // fused ConvLSTM cell epilogue (gate conv): lstm_R = rec_features, else 0
  int lstm_R;
  const float* c_prev;
  float* h_out;
  float* c_out;
};
int launch_conv3x3_tc(const TcConvArgs& a, cudaStream_t st);
int launch_pack_tc(const float* w_oihw, float* dst, int O, int I, int npad, cudaStream_t st);
// floats of one conv's packed tensor-core weights: [chunk][tap][hi|lo][plane(4)][npad][4]
__host__ __device__ inline size_t tc_packed_floats(int cin, int npad) {
  return (size_t)((cin + 15) / 16) * 9 * 2 * 4 * npad * 4;
}
inline int tc_npad(int cout) { return (cout + 15) / 16 * 16; }

// one destination of a routed (data-gradient) epilogue: columns [base, base + nch) -> p[pixel * cstride + coff + (n - base)]
struct ConvDst {
  float* p;
  const float* mask;       // same layout as p: result kept where mask > 0 (ReLU of the forward input), or null
  int cstride, coff, nch;
  int accum;               // 1: p += result
};

// fp16-operand persistent conv (conv3x3_f16.cu): stride 1, sources start on 8-channel plane boundaries
struct ConvF16Args {
  ConvSrc src[3];
  int nsrc;
  const void* wpk;         // fp16 [kstep][tap][hl][2 planes][npad][8]
  const float* inv_scale;  // 1 / (power-of-two weight scale)
  int npad;                // MMA N: multiple of 16, <= 256
  const float* bias;
  const float* gain;
  int act;
  float* out;
  int out_cstride, out_coff, cout;
  int B, H, W;
  int pad_replicate;
  int x3;
  int lstm_R;              // fused ConvLSTM cell epilogue when > 0 (multiple of 32, 4*R == npad)
  const float* c_prev;
  float* h_out;
  float* c_out;
  // backward use (data gradient = this kernel on transposed / tap-flipped weights, JOB_CONV_F16_T):
  const float* in_scale;   // device [2]: power-of-two scale applied to the source while staging (gradients are far below
                           // the fp16 normal range) and its inverse, applied in the epilogue; null = none
  int ndst;                // > 0: the output columns are routed to up to 3 destinations (the sources of the forward
  ConvDst dst[3];          // conv), each optionally gated by the sign of the forward input and accumulated
  unsigned* overflow;      // sticky flag (or null): set when a staged operand left the fp16 range (+-6e4) and was clamped
  const float* addend;     // optional per-pixel term added before bias / activation: [H*W][addend_stride], the SAME for every
  int addend_stride;       // sample (conditioning contribution hoisted out of the convolution when all samples share one input)
};
int launch_conv3x3_f16(const ConvF16Args& a, cudaStream_t st);
// scale[0] = 2^k with max|g| * 2^k in [2^10, 2^11) (1 when g == 0), scale[1] = 2^-k; scratch: 1024 floats
int launch_absmax_colsum(const float* g, int64_t npix, int cstride, int coff, int nch, float* scale, unsigned* sync, float* part,
                         float* gbias, int accum, cudaStream_t st);
// sync: two zero-initialised device words owned by the caller (self-resetting) -> one launch; null -> two launches via scratch
int launch_absmax_scale(const float* g, int64_t npix, int cstride, int coff, int nch, float* scale, float* scratch, cudaStream_t st,
                        unsigned* sync = nullptr);
bool convf16_supported(const ConvF16Args& a);
int convf16_ksteps(const int* nch, int nsrc);
size_t convf16_packed_floats(const int* nch, int nsrc, int npad);
// lstm_gate_f16.cu: ConvLSTM gate conv + cell update as a two-pass N-split implicit GEMM (R = 64).  Same argument block;
// wpk = JOB_GATE2P packing, addend (optional) = plane-transposed table [4R / 4][H*W] float4 in pass order WITH the bias.
bool lstm_gate_f16_supported(const ConvF16Args& a);
int launch_lstm_gate_f16(const ConvF16Args& a, cudaStream_t st);
size_t lstm_gate_packed_floats(const int* nch, int nsrc);
int launch_gate_addend_transpose(const float* src, const float* bias, float* dst, int HW, int ncol, int R, cudaStream_t st);

// Fused flow step on tensor cores (coupling_tc.cu): coupling net + coupling + 1x1 + ActNorm + log-det
struct CouplingArgs {
  ConvSrc src[2];          // the coupling-net input t as 1-2 NHWC sources; each starts on a 4-channel plane
  int nsrc;
  const float* w1;         // dense layer 1, packed [hi|lo][plane][16 (taps)][4]
  const float* w2;         // dense layer 2 (one more plane: d1)
  const float* w3;         // Conv2dZeros, packed [tap][hi|lo][plane][npad][4]
  int npad;
  const float* bias3;
  const float* gain3;
  int C;
  const float* y_in;       // [B,HW,C] flow state read by the epilogue
  float* y_out;            // [B,HW,C] (must differ from y_in: other CTAs read y_in's halo)
  const float* wmat;       // CxC mix or null
  const float* nw;         // ActNorm or null
  const float* nb;
  int reverse;
  float* ld_part;
  int ld_stride;
  int B, H, W;
  int split3;
};
int launch_coupling_tc(const CouplingArgs& a, cudaStream_t st);
int coupling_tc_tiles(int H, int W);
// plane bookkeeping shared by host packing and the kernel: sources start on plane boundaries
inline int cpl_planes(int nch0, int nch1) { return (nch0 + 3) / 4 + (nch1 + 3) / 4; }

// Fused flow step, fp16-operand generation (flow_step_f16.cu): persistent + pipelined, K layout [src0 | d1 d2 | src1]
struct Step2Args {
  ConvSrc src[2];          // coupling-net input t (src[1] = the conditioning map, skipped when hoisted)
  int nsrc;
  int hoist;               // 1: src[1]'s contribution comes from the per-step tables dc / hc (same for every sample)
  const float* dc;         // [HW][dc_stride]: cond part of d1 (+0) and d2 (+1), gathered (pre-ReLU partial sums)
  int dc_stride;
  const float* hc;         // [HW][hc_stride]: cond part of the Conv2dZeros output (before bias and gain)
  int hc_stride;
  int hoist_bstride;       // 0: one table shared by all samples; 1: tables per sample ([B][HW][stride])
  const void* wE;          // fp16 [hl][planes][32][8]   taps of dense layers 1 (cols 0-8) and 2 (cols 16-24)
  const void* wZ;          // fp16 [tap][hl][planes][NP][8]
  const float* wmisc;      // [0..8] taps of layer 2's d1 input row, [9..11] inverse weight scales (w1, w2, w3)
  const float* bias3;
  const float* gain3;
  int C;
  const float* y_in;
  float* y_out;
  const float* wmat;
  const float* nw;
  const float* nb;
  int reverse;
  float* ld_part;
  int ld_stride;
  int B, H, W;
  int x3;                  // 1: hi/lo operand split (fp32-grade), 0: single-pass fp16
  long long* prof;         // developer profiling (TMG_STEP2_PROF=1): per-CTA cycle counters per role, else null
  float* d_emit;           // training forward: relu(d1), relu(d2) [B,HW,2] and ...
  float* h_emit;           // ... h [B,HW,C] written for the backward pass (both or neither)
  unsigned* overflow;      // sticky flag (or null): set when a staged operand left the fp16 range (+-6e4) and was clamped
};
int launch_step2(const Step2Args& a, cudaStream_t st);
bool step2_supported(const Step2Args& a);
int step2_ld_slots(int H, int W);   // log-det partial slots per sample: 8 (epilogue warps) per 16x16 tile
void step2_klayout(int nch0, int nch1, int& KSy, int& KS1, int& kd);
size_t step2_wE_floats(int nch0, int nch1);
size_t step2_wZ_floats(int nch0, int nch1, int C);

// Level-resident fused flow steps (flow_level_f16.cu): ONE launch runs every plain step of a level in the reverse
// (sampling) direction with the flow state resident on the SM -- fp32 state in registers, fp16 hi/lo operand planes of the
// coupling-net input in shared memory --, so the state makes one HBM round trip per level instead of one per step.
// Requires the hoisted conditioning tables (dc / hc).  Offsets are in floats from the parameter / packed bases.
struct LevelStep {
  int64_t wE, wZ, misc, gain, W;   // packed buffer: fp16 weights of flow_step_f16 (same packing), misc[12], gain, 1x1 mix W [C][C]
  int64_t wEc, wZc;                // packed buffer: compact / tap-paired weights (JOB_STEP2C) or -1
  int64_t bias, nw, nb;            // parameter buffer: Conv2dZeros bias, ActNorm weight / bias (-1: step without ActNorm)
  int dc_off, hc_off;              // this step's index inside the (transposed) hoisted tables
};
struct LevelArgs {
  const LevelStep* steps;          // device, in execution order (steps n-1 .. 1 of the block, reversed)
  int nsteps;
  const float* params;
  const float* packed;
  const float* y_in;               // [B,HW,C]
  float* y_out;                    // [B,HW,C] (may alias y_in: every sample is read and written by one CTA)
  const float* dc;                 // hoisted tables in the plane-transposed layout (launch_hoist_transpose):
  const float* hc;                 //   dc [Bx][nsteps_tab][HW] float2, hc [Bx][nsteps_tab][C/4][HW] float4
  int nsteps_tab;
  int hoist_bstride;               // 0: tables shared by all samples; 1: per sample
  float* ld_part;                  // [B][ld_stride]: sum over the steps and pixels of the coupling log-det -> entry 0
  int ld_stride;
  int B, H, W, C;
  int x3;
  int nch1;                        // conditioning channels (needed for the strides of the packed weights)
  int compact;                     // 1: every step of the table carries the compact / tap-paired weights (wEc, wZc)
  unsigned* overflow;              // sticky flag: set when an activation exceeded the fp16 operand range (+-6e4)
  long long* prof;                 // developer profiling (TMG_LV_PROF=n): per-CTA cycle counters per role and phase, else null
};
// wmx: device [nsteps] offsets (floats, packed buffer) of each step's mix W as fp16 hi/lo MMA operand (JOB_1X1, dst[3]) or
// null -- the 1x1 mix then runs on the CUDA cores.  (Passed beside LevelArgs, not inside: the kernel argument block must keep
// its layout -- the register allocation of the C = 12 kernel, at its 80-register cap, changed with the size of LevelArgs /
// LevelStep: spills 88 -> 182 bytes, +9 % time.)
bool level_resident_supported(const LevelArgs& a, bool mix_mma = false);
// dc_all [Bx][HW][dstride] (cols 2s, 2s+1) / hc_all [Bx][HW][hstride] (cols s*C + n) -> the transposed layouts above
int launch_hoist_transpose(const float* dc_all, int dstride, const float* hc_all, int hstride, float* dcT, float* hcT,
                           int Bx, int HW, int nsteps, int C, cudaStream_t st);
// unsq (wide levels, or null): the result goes un-squeezed into the next level's state [B, 2H, 2W, unsq_cstride] (channels
// [0, C/4)) instead of y_out -- CheckerSqueeze.reverse fused into the kernel's final store
int launch_level_resident(const LevelArgs& a, const int64_t* wmx, float* unsq, int unsq_cstride, cudaStream_t st);
// floats of the fp16 hi/lo mix operand of one step: [4: 1/scale, pad][hl 2][K planes, even][NP rows][8 halves]
inline int64_t level_mix_floats(int C) { const int PM = (C / 8 + 1) / 2 * 2, NP = (C + 15) / 16 * 16; return 4 + (int64_t)2 * PM * NP * 8 / 2; }

// ------------------------------------------------------------------ pointwise flow step
struct PointArgs {
  float* y;            // [B, HW, C] in/out
  const float* hr;     // coupling-net output [B, HW, C] or null (no coupling)
  const float* wmat;   // [C][C] row-major or null
  const float* nw;     // ActNorm weight [C] or null
  const float* nb;     // ActNorm bias   [C]
  int reverse;         // 0: coupling_fwd -> norm_fwd -> W ; 1: coupling_rev -> W -> norm_rev
  int B, HW, C;
  float* ld_part;      // per-CTA partial sums of the coupling log-det: [B][ld_stride] (+ cta)
  int ld_stride;
};
int launch_flow_pointwise(const PointArgs& a, cudaStream_t st);

struct LstmArgs {
  const float* gates;  // [B,HW,4R]  order i,f,o,g
  const float* c_prev; // [B,HW,R] or null (zeros)
  float* h_out;        // [B,HW,R]
  float* c_out;
  int64_t n;           // B*HW*R
  int R;
};
int launch_lstm_pointwise(const LstmArgs& a, cudaStream_t st);

struct GaussArgs {
  const float* prm;    // NHWC [B,HW,prm_cstride]: mean at ch j, log-std at ch n+j
  int prm_cstride;
  int prm_bshared;     // 1: prm has batch 1, shared by every sample
  float* val;          // NHWC value tensor, element (pixel, val_coff + j)
  int val_cstride, val_coff;
  const float* eps_in; // NCHW [B,n,HW]   (reverse: val = mean + exp(logsd)*eps)
  float* eps_out;      // NCHW [B,n,HW] or null (forward: eps = (val-mean)/exp(logsd))
  float* val_nchw;     // optional NCHW copy of val (user-facing z)
  int reverse;
  int B, HW, n;
  float* ld_part;
  int ld_stride;
};
int launch_gaussian(const GaussArgs& a, cudaStream_t st);

enum PermuteMode {
  PERM_NCHW_TO_NHWC = 0,
  PERM_NHWC_TO_NCHW = 1,
  PERM_SQUEEZE_NCHW_TO_NHWC = 2,    // src [B,c,2H,2W] NCHW -> dst [B,H,W,4c]
  PERM_SQUEEZE_NHWC_TO_NHWC = 3,    // src [B,2H,2W,c]      -> dst [B,H,W,4c]
  PERM_UNSQUEEZE_NHWC_TO_NHWC = 4,  // src [B,H,W,4c]       -> dst [B,2H,2W,c]
  PERM_UNSQUEEZE_NHWC_TO_NCHW = 5,  // src [B,H,W,4c]       -> dst [B,c,2H,2W] NCHW
  PERM_SQUEEZE_NCHW_TO_NCHW = 6,    // reference layout on both sides (flowUtils.py:99-121)
  PERM_UNSQUEEZE_NCHW_TO_NCHW = 7   // (flowUtils.py:124-145)
};
struct PermArgs {
  const float* src;
  float* dst;
  int mode;
  int B, C, H, W;       // C,H,W of the *un-squeezed* tensor for squeeze modes, plain dims otherwise
  int src_cstride, src_coff, dst_cstride, dst_coff;   // used for NHWC sides
};
int launch_permute(const PermArgs& a, cudaStream_t st);

struct UpsampleArgs {
  const float* src;    // NHWC [B,h,w,C]
  float* dst;          // NHWC [B,h*f,w*f,C]
  int B, h, w, C, f;
};
int launch_upsample(const UpsampleArgs& a, cudaStream_t st);

struct BnStatArgs {
  const float* x;      // NHWC [N, cstride]
  int cstride, c0, n;  // channels [c0, c0+n)
  int64_t N;           // B*h*w
  float* mean;         // [>= c0+n]
  float* var;          // biased
};
int launch_bn_stats(const BnStatArgs& a, cudaStream_t st);

struct BnFoldArgs {
  const float* mean;
  const float* var;
  const float* w;
  const float* b;
  float* run_mean;
  float* run_var;
  float* scale;
  float* shift;
  int n;
  int64_t N;
  float eps, momentum;
};
int launch_bn_fold_train(const BnFoldArgs& a, cudaStream_t st);

struct LogdetArgs {
  const float* ld_part;    // [B][ld_stride]
  int ld_stride;
  const float* step_const; // per flow step: sum log|w| - sum log_s
  int n_levels;
  int step_begin[TMG_MAX_LEVELS + 1];
  int hw[TMG_MAX_LEVELS];
  float* out;              // [B]
  int B;
};
int launch_logdet_reduce(const LogdetArgs& a, cudaStream_t st);

// ------------------------------------------------------------------ backward building blocks (backward.cu)
// Weight gradient of a 3x3 stride-1 convolution: gw[o][c][tap] = sum_{b,p} g[b,p,o] * xin[b,p+off(tap),c] with xin the
// forward input (virtual concat, ReLU / padding applied while staging).  Deterministic: partial sums per pixel
// split, reduced in a fixed order.
struct WgradArgs {
  ConvSrc src[3];
  int nsrc;
  int cin;                 // channels of the concatenation
  const float* g;          // output gradient [B,HW,g_cstride], channels [g_coff, g_coff + cout)
  int g_cstride, g_coff, cout;
  int B, H, W;             // OUTPUT size of the convolution (= input size when stride is 1)
  int stride;              // 0/1: stride 1, 2: stride 2 (input size Hin x Win)
  int Hin, Win;
  const float* bn_scale;   // optional per-(concatenated)-channel affine applied before the ReLU (folded BatchNorm)
  const float* bn_shift;
  int pad_replicate;
  float* gw;               // OIHW [cout][cin][3][3], accumulated into (+=) when accum
  float* gbias;            // [cout] column sums of g (+= when accum) or null
  int accum;
  float* scratch;          // >= wgrad_scratch_floats(...)
};
size_t wgrad_scratch_floats(int cout, int cin, int B, int H, int W);
int launch_wgrad(const WgradArgs& a, cudaStream_t st);
// exact-fp32 streaming weight gradient of a Cout = 1, zero-padded, bias-free convolution (dense layers)
bool wgrad_cout1_supported(const WgradArgs& a);
size_t wgrad_cout1_scratch_floats(int cin, int B, int H, int W);
int launch_wgrad_cout1(const WgradArgs& a, cudaStream_t st);
// tensor-core weight gradient (wgrad_f16.cu): stride 1, no folded BatchNorm; gscale = launch_absmax_scale of a.g;
// a.scratch >= wgrad_f16_scratch_floats
bool wgrad_f16_supported(const WgradArgs& a);
size_t wgrad_f16_scratch_floats(int cout, const int* nch, int nsrc, int B, int H, int W);
int launch_wgrad_f16(const WgradArgs& a, const float* gscale, cudaStream_t st);
float* wgrad_f16_bias_partials(const WgradArgs& a);
// tap-major transposed + flipped weights for the data gradient: wt[tap][o][c (padded to 4)] = w[o][c][8 - tap]
int launch_pack_dgrad(const float* w_oihw, float* wt, int O, int I, cudaStream_t st);
// replicate padding: adds the gradient of the out-of-image ring to the clamped border pixels
struct RingArgs {
  const float* g; int g_cstride, g_coff, cout;
  const float* w_oihw; int cin_total;                 // all input channels of the concatenation, routed to dst[]
  int ndst; ConvDst dst[3];                           // mask = forward input of the source (ReLU gate) or null; p += ring
  int B, H, W;
};
int launch_dgrad_ring(const RingArgs& a, cudaStream_t st);
// exact-fp32 data gradient of a Cout = 1, zero-padded 3x3 convolution, routed to the forward sources (ReLU mask, accumulate)
int launch_dgrad_cout1(const float* g, int g_cstride, int g_coff, const float* w_oihw, int cin, const ConvDst* dst, int ndst,
                       int B, int H, int W, cudaStream_t st);

// Pointwise part of one REVERSE flow step (z -> y, the direction training runs: TMGlow.sample), backward:
//   forward:  a = 2*softsign(h_odd), v = [y1, y2*exp(-a) - h_even], u = W v, out = (u - nb)/nw,
//             h = (z + b3)*gain with z the raw Conv2dZeros output, ld += sum(a)
struct StepBwdArgs {
  const float* y_in;       // [B,HW,C] step input
  const float* h;          // [B,HW,C] coupling-net output (after bias and gain)
  const float* g_out;      // [B,HW,C] gradient w.r.t. the step output
  const float* g_ld;       // [B] gradient w.r.t. the per-sample log-det
  const float* wmat;       // [C][C] W (or null: no mix)
  const float* nw;         // ActNorm weight / bias or null
  const float* nb;
  const float* gain;       // scalar exp(clamp(scale))
  float* g_y;              // [B,HW,C] gradient w.r.t. y_in (direct part: the coupling-net part is added by the conv dgrads)
  float* g_z;              // [B,HW,C] gradient w.r.t. the raw Conv2dZeros output
  float* gu;               // [B,HW,C] gradient w.r.t. u   (1x1 weight gradient = sum gu (x) v)
  float* v;                // [B,HW,C] mixed vector v
  float* part;             // per-CTA partial sums [nblocks][3C+2]: -sum gu | -sum gu*out | sum g_h*h/gain | sum gz | max|gz|
  int B, HW, C;
};
int launch_step_bwd(const StepBwdArgs& a, cudaStream_t st);
int step_bwd_blocks(int B, int HW);
// sum over CTAs (fixed order) of column-partials: out[i] (+)= scale * sum_k part[k][stride] ...
int launch_reduce_cols(const float* part, int nrows, int row_stride, int col0, int ncols, float* out, int accum, cudaStream_t st);
// gw[o][k] (+)= sum_px gu[px][o] * v[px][k]   (C x C), deterministic
int launch_outer_wgrad(const float* gu, const float* v, int64_t npix, int C, float* gw, float* scratch, cudaStream_t st, int accum = 0);
size_t outer_wgrad_scratch_floats(int64_t npix, int C);
// Chain rule of W = P (l*lm + I)(u*um + diag(sign*exp(log_s)) + 0.01 I)  (glowConv.py:151-160) and of the log-det
// constants: g_l, g_u, g_log_s (+=), g_norm_w += HW/w * sum_b g_ld
struct LuBwdArgs {
  const float* dW;         // [C][C]
  const float *l, *u, *log_s, *p, *sign_s, *lmask, *umask, *eye;
  float *g_l, *g_u, *g_log_s;
  const float* nw; float* g_nw;      // ActNorm weight and its gradient (or null)
  const float* g_ld; int B; float hw;
  int C;
};
int launch_lu_bwd(const LuBwdArgs& a, cudaStream_t st);
// deferred LU backward (once per optimizer step): per flow step the offsets (floats, same in the parameter and gradient
// buffers) of l, u, log_s, p, sign_s, l_mask, u_mask, eye and of the ActNorm weight (-1: none)
struct LuTabEntry { int C; int64_t off[8]; int64_t norm_w; };
int launch_lu_bwd_batched(const LuTabEntry* tab_dev, int n, int cmax, const float* params, float* grads, cudaStream_t st);
// g_bias: bias gradient of the Conv2dZeros conv (column sums of gz, accumulated) or null; gz_scale: [2] power-of-two scale
// of gz and its inverse (or null)
int launch_step_param_grads(const float* part, int nrows, int C, float* g_nb, float* g_nw, const float* scale_param, float* g_scale,
                            const float* g_ld, int B, float hw, float* gld_stash, float* g_bias, float* gz_scale, cudaStream_t st);
int launch_scale_grad(const float* s_gain, const float* scale_param, float* g_scale, cudaStream_t st);
// ConvLSTM cell backward (convLSTM.py:76-83): gates = pre-activations [B,HW,4R] (i,f,o,g), returns g_gates and g_c_prev
struct LstmBwdArgs {
  const float* gates; const float* c_prev;      // c_prev may be null (zeros)
  const float* g_h; const float* g_c;           // gradients w.r.t. h', c' (g_c may be null)
  float* g_gates; float* g_c_prev;              // g_c_prev may be null
  int64_t n; int R;
};
int launch_lstm_bwd(const LstmBwdArgs& a, cudaStream_t st);
// Diagonal Gaussian, reverse direction (value = mean + exp(logsd)*eps [, logp = sum -0.5(log2pi + 2 logsd + eps^2)]):
// gradient w.r.t. the prior parameters.  Split prior: prm = hardtanh((conv+b)*gain, -2, ln5) (flowUtils.py:270-274),
// the result is the gradient w.r.t. the raw conv output (gain applied) and the gain partial sums; top prior
// (tmGlow.py:460-463): prm = encoder output, clamp(-10, ln5) on the log-std half, no log-prob term.
struct GaussBwdArgs {
  const float* prm; int prm_cstride;     // NHWC: mean at j, log-std at n + j
  const float* g_val; int gv_cstride, gv_coff;    // gradient w.r.t. the value (NHWC)
  const float* eps;                       // NCHW [B,n,HW]
  const float* g_ld;                      // [B] or null (no log-prob term)
  const float* gain;                      // scalar or null
  int hardtanh;                           // 1: prm went through hardtanh(-2, ln5)
  float* g_prm; int gp_cstride;           // out: gradient w.r.t. the 2n prior channels (before gain when gain != null)
  float* part;                            // per-CTA partial sums of g_prm * prm / gain (gain gradient) or null
  int B, HW, n;
};
int launch_gauss_bwd(const GaussBwdArgs& a, cudaStream_t st);
// adjoint of the bilinear upsampling (align_corners=True): gsrc[b,y,x,c] = sum of the dst gradients that read it
int launch_upsample_bwd(const float* gdst, float* gsrc, int B, int h, int w, int C, int f, cudaStream_t st);
// data gradient of a stride-2 3x3 convolution (pad 1): gx[b,q,c] (+)= sum_{tap,o} w[o][c][tap] g[b,(q-off)/2,o], gated by mask > 0
struct S2DgradArgs {
  const float* g; int g_cstride, g_coff, cout; int Hout, Wout;
  const float* w_oihw; int cin;
  const float* mask; float* gx; int gx_cstride, gx_coff; int accum;
  int B, Hin, Win;
};
int launch_dgrad_s2(const S2DgradArgs& a, cudaStream_t st);
// BatchNorm2d + ReLU backward (denseBlock.py:49-50): ga = gradient w.r.t. relu(bn(x)); batch statistics (train) or
// running statistics (eval); accumulates g_x into gx and g_gamma / g_beta into the flat gradient buffer
struct BnBwdArgs {
  const float* x; int x_cstride;          // channels [0, n)
  const float* ga; int ga_cstride;
  const float* mean; const float* var;    // per channel (batch or running)
  const float* gamma; const float* beta;
  float* gx; int gx_cstride;              // +=
  float* g_gamma; float* g_beta;          // +=
  float* sums;                            // scratch [2n]
  int n; int64_t N; float eps; int train;
};
int launch_bn_relu_bwd(const BnBwdArgs& a, cudaStream_t st);
int gauss_bwd_blocks(int B, int HW);

// ------------------------------------------------------------------ weight packing jobs
enum PackJobType { JOB_CONVW = 0, JOB_1X1 = 1, JOB_GAIN = 2, JOB_BN = 3, JOB_CONVW_TC = 4, JOB_CPL_W12 = 5, JOB_CPL_W3 = 6,
                   JOB_STEP2 = 7, JOB_HOIST = 8, JOB_CONV_F16 = 9, JOB_CONV_F16_T = 10, JOB_STEP2C = 11, JOB_SLICE = 12, JOB_GATE2P = 13 };
struct PackJob {
  int type;
  int a, b;              // JOB_CONVW: O, I ; JOB_1X1: C ; JOB_BN: n
  int64_t src[9];        // offsets (floats) into the flat parameter buffer, -1 = absent
                         // JOB_1X1: l,u,log_s,p,sign_s,l_mask,u_mask,eye,norm.weight
  int64_t dst[4];        // offsets (floats) into the packed buffer (JOB_1X1: W, W^-1, step constant, fp16 hi/lo W for the MMA mix or -1)
  int opad;              // JOB_CONVW: padded O ; JOB_CONVW_TC: npad
  int part, nparts;      // JOB_CONVW_TC: this CTA packs elements [part*chunk, (part+1)*chunk)
  int nch0, nch1, nd;    // JOB_CPL_*: channels of source 0 / source 1 / d channels (concat order: src0, src1, d)
  int nplanes;           // JOB_CPL_*: K planes in the packed tensor
};
int launch_pack(const PackJob* jobs_dev, int njobs, const float* params, float* packed, int cmax,
                cudaStream_t st);

}  // namespace tmg
