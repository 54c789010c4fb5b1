// Optimizer step of the reference trainer as two launches and no host round trip:
//   torch.nn.utils.clip_grad_norm_(model.parameters(), max_grad_norm); optimizer.step()     (nn/trainFlowParallel.py:290-291)
//   optimizer = torch.optim.Adam(model.parameters(), lr, weight_decay=1e-8, amsgrad=True)    (main.py:78)
// on the FLAT parameter / gradient buffers (every parameter and buffer of the model back to back, tmg_model_param_*).
// Round 1 ran torch.optim.Adam plus `float(g.norm())` on the host: a device synchronisation in every optimizer step.
//   launch 1: per-block sums of squares of the gradient (fixed order -> bit-reproducible); block 0 advances the step counter
//   launch 2: every block adds the partial sums in the same fixed order -> ||g||, clip coefficient min(1, max_norm / (||g|| + 1e-6))
//             (exactly clip_grad_norm_), then weight decay on the TRAINABLE entries only (mask: the flat buffer also holds
//             permutations, masks and BatchNorm running statistics, which must not move), Adam / AMSGrad update in the order of
//             operations of torch.optim.Adam's single-tensor path.
// Hyper-parameters and the step counter live in device memory (`hyper`): a CUDA graph of the training step replays unchanged
// while the learning-rate schedule writes a new lr between replays.
#include "common.cuh"

namespace tmg {

constexpr int kAdamThreads = 256;
constexpr int kAdamMaxBlocks = 1024;

// hyper: [0] lr, [1] beta1, [2] beta2, [3] eps, [4] weight_decay, [5] max_norm (<= 0: no clipping), [6] step (as float, exact
// up to 2^24), [7] amsgrad flag; out: [0] gradient norm before clipping, [1] clip coefficient
__global__ void __launch_bounds__(kAdamThreads)
grad_sumsq_kernel(const float* __restrict__ g, int64_t n, float* __restrict__ part, float* __restrict__ hyper) {
  __shared__ double red[kAdamThreads / 32];
  double s = 0.0;
  const int64_t per = (n + gridDim.x - 1) / gridDim.x;
  const int64_t lo = (int64_t)blockIdx.x * per, hi = min(n, lo + per);
  for (int64_t i = lo + threadIdx.x; i < hi; i += kAdamThreads) { const double v = (double)g[i]; s += v * v; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kAdamThreads / 32; ++w) t += red[w];
    reinterpret_cast<double*>(part)[blockIdx.x] = t;
    if (blockIdx.x == 0) hyper[6] += 1.f;          // optimizer.step(): state['step'] += 1
  }
}

__global__ void __launch_bounds__(kAdamThreads)
adam_update_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m1, float* __restrict__ m2,
                   float* __restrict__ m2max, const float* __restrict__ mask, int64_t n, const float* __restrict__ part,
                   int nparts, const float* __restrict__ hyper, float* __restrict__ out) {
  __shared__ float s_coef;
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < nparts; ++i) t += reinterpret_cast<const double*>(part)[i];
    const float norm = (float)sqrt(t);
    const float mx = hyper[5];
    float coef = 1.f;
    if (mx > 0.f) { const float c = mx / (norm + 1e-6f); coef = c < 1.f ? c : 1.f; }     // clip_grad_norm_: clamp(max_norm / (norm + 1e-6), max = 1)
    s_coef = coef;
    if (blockIdx.x == 0) { out[0] = norm; out[1] = coef; }
  }
  __syncthreads();
  const float coef = s_coef;
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], wd = hyper[4], step = hyper[6];
  const bool ams = hyper[7] != 0.f;
  const float bc1 = 1.f - powf(b1, step), bc2 = 1.f - powf(b2, step);
  const float step_size = lr / bc1, bc2_sqrt = sqrtf(bc2);
  for (int64_t i = (int64_t)blockIdx.x * kAdamThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kAdamThreads) {
    const float mk = mask ? mask[i] : 1.f;
    if (mk == 0.f) continue;                                       // buffers inside the flat parameter: untouched
    float gi = g[i] * coef;
    const float pi = p[i];
    if (wd != 0.f) gi = fmaf(wd, pi, gi);                          // grad = grad.add(param, alpha=weight_decay)
    float a = m1[i];
    a = a + (gi - a) * (1.f - b1);                                 // exp_avg.lerp_(grad, 1 - beta1)
    const float v = m2[i] * b2 + (1.f - b2) * gi * gi;             // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    m1[i] = a; m2[i] = v;
    float vv = v;
    if (ams) { vv = fmaxf(m2max[i], v); m2max[i] = vv; }           // torch.maximum(max_exp_avg_sqs, exp_avg_sq)
    const float denom = sqrtf(vv) / bc2_sqrt + eps;
    p[i] = pi - step_size * (a / denom);                           // param.addcdiv_(exp_avg, denom, value=-step_size)
  }
}

}  // namespace tmg

using namespace tmg;

extern "C" {

size_t tmg_adam_workspace_bytes(void) { return (size_t)kAdamMaxBlocks * sizeof(double) + 64; }

int tmg_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, float* max_exp_avg_sq, const float* mask,
                  int64_t n, float* hyper, float* out, void* workspace, size_t workspace_bytes, void* stream) {
  if (!params || !grads || !exp_avg || !exp_avg_sq || !hyper || !out || !workspace) { set_error("null argument"); return TMG_ERR_NULL; }
  if (n <= 0) { set_error("empty parameter buffer"); return TMG_ERR_BAD_SHAPE; }
  if (workspace_bytes < tmg_adam_workspace_bytes()) { set_error("workspace too small"); return TMG_ERR_WORKSPACE; }
  if (((uintptr_t)workspace & 7) != 0) { set_error("workspace must be 8-byte aligned"); return TMG_ERR_BAD_SHAPE; }
  cudaStream_t st = (cudaStream_t)stream;
  const int nb = (int)std::min<int64_t>(kAdamMaxBlocks, (n + 4095) / 4096);
  float* part = (float*)workspace;
  grad_sumsq_kernel<<<nb, kAdamThreads, 0, st>>>(grads, n, part, hyper);
  TMG_LAUNCH_CHECK();
  const int nu = (int)std::min<int64_t>(4 * 148, (n + kAdamThreads - 1) / kAdamThreads);
  adam_update_kernel<<<nu, kAdamThreads, 0, st>>>(params, grads, exp_avg, exp_avg_sq, max_exp_avg_sq, mask, n, part, nb, hyper, out);
  TMG_LAUNCH_CHECK();
  return TMG_OK;
}

}  // extern "C"
