// Backward building blocks of the flow path (exact fp32 on CUDA cores): first step towards training through
// TMGlow.sample (reverse KL, nn/trainFlowParallel.py:204-311).  The adjoints restated here:
//   conv3x3 weight gradient        (every nn.Conv2d of the path; autograd of F.conv2d)
//   conv3x3 data gradient          = the forward kernel (conv3x3_ffma.cu) on flipped / transposed weights,
//                                    plus the border terms of replicate padding (Conv2dZeros, flowUtils.py:246)
// Deterministic: no atomics; partial sums are reduced in a fixed order.
#include "common.cuh"

namespace tmg {

// ------------------------------------------------------------------ dgrad weights: wt[tap][o][c] = w[o][c][8 - tap]
__global__ void pack_dgrad_kernel(const float* __restrict__ w, float* __restrict__ wt, int O, int I, int Ip) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 9 * O * Ip) return;
  const int c = i % Ip; int t = i / Ip; const int o = t % O, tap = t / O;
  wt[i] = c < I ? w[((size_t)o * I + c) * 9 + (8 - tap)] : 0.f;
}
int launch_pack_dgrad(const float* w_oihw, float* wt, int O, int I, cudaStream_t st) {
  const int Ip = (I + 3) / 4 * 4, n = 9 * O * Ip;
  pack_dgrad_kernel<<<cdiv(n, 256), 256, 0, st>>>(w_oihw, wt, O, I, Ip);
  TMG_LAUNCH_CHECK();
  return TMG_OK;
}

// ------------------------------------------------------------------ weight gradient
constexpr int kWgOB = 16, kWgCB = 8, kWgThreads = 128, kWgMaxPx = 512;

__global__ void __launch_bounds__(kWgThreads)
wgrad_kernel(WgradArgs a, int TR, int nsplit, float* part, float* part_b) {
  extern __shared__ __align__(16) float smem[];
  const int tw = a.W + 2;
  float* g_s = smem;                                   // [TR*W][16]
  float* x_s = smem + (size_t)TR * a.W * kWgOB;        // [(TR+2)*(W+2)][8]
  const int tid = threadIdx.x, o_l = tid & 15, c_l = tid >> 4;
  const int c0 = blockIdx.x * kWgCB, o0 = blockIdx.y * kWgOB, split = blockIdx.z;
  const int rtiles = (a.H + TR - 1) / TR, units = a.B * rtiles;
  // source of this thread's staging channels is resolved per element (scalar path, like the forward kernel)
  float acc[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) acc[t] = 0.f;
  float accb = 0.f;
  for (int u = split; u < units; u += nsplit) {
    const int b = u / rtiles, r0 = (u - b * rtiles) * TR;
    const int rows = min(TR, a.H - r0);
    __syncthreads();
    for (int i = tid; i < rows * a.W * kWgOB; i += kWgThreads) {
      const int o = i & 15, p = i >> 4;
      const int r = r0 + p / a.W, cc = p % a.W;
      g_s[i] = (o0 + o < a.cout) ? __ldg(a.g + ((size_t)(b * a.H + r) * a.W + cc) * a.g_cstride + a.g_coff + o0 + o) : 0.f;
    }
    for (int i = tid; i < (rows + 2) * tw * kWgCB; i += kWgThreads) {
      const int c = i & 7, q = i >> 3;
      int r = r0 - 1 + q / tw, cc = q % tw - 1;
      bool inb = r >= 0 && r < a.H && cc >= 0 && cc < a.W;
      if (a.pad_replicate) { r = min(max(r, 0), a.H - 1); cc = min(max(cc, 0), a.W - 1); inb = true; }
      float v = 0.f;
      int ch = c0 + c;
      if (inb && ch < a.cin) {
        const ConvSrc* s = &a.src[0];
        if (ch >= s->nch && a.nsrc > 1) { ch -= s->nch; s = &a.src[1];
          if (ch >= s->nch && a.nsrc > 2) { ch -= s->nch; s = &a.src[2]; } }
        if (ch < s->nch && s->p) {
          v = __ldg(s->p + ((size_t)((s->bshared ? 0 : b) * a.H + r) * a.W + cc) * s->cstride + s->coff + ch);
          if (s->relu) v = fmaxf(v, 0.f);
        }
      }
      x_s[i] = v;
    }
    __syncthreads();
    for (int pr = 0; pr < rows; ++pr) {
      const float* gp = g_s + (size_t)pr * a.W * kWgOB + o_l;
      const float* xp = x_s + (size_t)pr * tw * kWgCB + c_l;
      for (int pc = 0; pc < a.W; ++pc) {
        const float gv = gp[pc * kWgOB];
        accb += gv;
#pragma unroll
        for (int t = 0; t < 9; ++t) acc[t] = fmaf(gv, xp[((t / 3) * tw + pc + t % 3) * kWgCB], acc[t]);
      }
    }
  }
  if (o0 + o_l < a.cout && c0 + c_l < a.cin) {
    float* pp = part + (((size_t)split * a.cout + o0 + o_l) * a.cin + c0 + c_l) * 9;
#pragma unroll
    for (int t = 0; t < 9; ++t) pp[t] = acc[t];
  }
  if (part_b && blockIdx.x == 0 && c_l == 0 && o0 + o_l < a.cout) part_b[(size_t)split * a.cout + o0 + o_l] = accb;
}

// out[i] (+)= sum_s part[s][i], fixed order
__global__ void reduce_partials_kernel(const float* __restrict__ part, float* __restrict__ out, int n, int nsplit, int accum) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int k = 0; k < nsplit; ++k) s += part[(size_t)k * n + i];
  out[i] = accum ? out[i] + s : s;
}

static void wgrad_plan(int cout, int cin, int B, int H, int W, int& TR, int& nsplit) {
  TR = std::max(1, std::min(H, kWgMaxPx / std::max(W, 1)));
  const int units = B * cdiv(H, TR);
  const int blocks = cdiv(cin, kWgCB) * cdiv(cout, kWgOB);
  nsplit = std::max(1, std::min(units, cdiv(592, blocks)));
}
size_t wgrad_scratch_floats(int cout, int cin, int B, int H, int W) {
  int TR, ns;
  wgrad_plan(cout, cin, B, H, W, TR, ns);
  return (size_t)ns * ((size_t)cout * cin * 9 + cout) + 64;
}
int launch_wgrad(const WgradArgs& a, cudaStream_t st) {
  if (a.B <= 0 || a.cout <= 0 || a.cin <= 0) return TMG_OK;
  if (a.W > kWgMaxPx) { set_error("wgrad: width %d not supported", a.W); return TMG_ERR_UNSUPPORTED; }
  int TR, ns;
  wgrad_plan(a.cout, a.cin, a.B, a.H, a.W, TR, ns);
  const size_t smem = ((size_t)TR * a.W * kWgOB + (size_t)(TR + 2) * (a.W + 2) * kWgCB) * sizeof(float);
  if (smem > 96 * 1024) { set_error("wgrad: tile needs %zu B of shared memory", smem); return TMG_ERR_UNSUPPORTED; }
  TMG_CUDA_OK(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  float* part = a.scratch;
  float* part_b = a.gbias ? a.scratch + (size_t)ns * a.cout * a.cin * 9 : nullptr;
  dim3 grid(cdiv(a.cin, kWgCB), cdiv(a.cout, kWgOB), ns);
  wgrad_kernel<<<grid, kWgThreads, smem, st>>>(a, TR, ns, part, part_b);
  TMG_LAUNCH_CHECK();
  const int n = a.cout * a.cin * 9;
  reduce_partials_kernel<<<cdiv(n, 256), 256, 0, st>>>(part, a.gw, n, ns, a.accum);
  TMG_LAUNCH_CHECK();
  if (a.gbias) {
    reduce_partials_kernel<<<cdiv(a.cout, 256), 256, 0, st>>>(part_b, a.gbias, a.cout, ns, a.accum);
    TMG_LAUNCH_CHECK();
  }
  return TMG_OK;
}

// ------------------------------------------------------------------ replicate padding: border terms of the data gradient
// Forward reads xin[clamp(p + off)].  The zero-padding data gradient covers every in-image position; what is missing
// is the gradient that belongs to the out-of-image ring, which the clamp routes to the border pixels.  One thread per
// (sample, border pixel, channel) sums its ring positions in a fixed order (no atomics).
__global__ void dgrad_ring_kernel(RingArgs a, int nborder) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.B * nborder * a.nch) return;
  const int c = i % a.nch; int t = i / a.nch; const int k = t % nborder, b = t / nborder;
  // border pixel k: top row, bottom row, then the left / right columns without the corners
  int y, x;
  if (k < a.W) { y = 0; x = k; }
  else if (k < 2 * a.W) { y = a.H - 1; x = k - a.W; }
  else { const int r = k - 2 * a.W; y = 1 + (r >> 1); x = (r & 1) ? a.W - 1 : 0; }
  if (a.H == 1 && k >= a.W) return;                       // single row: the top row already covers it
  if (a.W == 1 && k >= 2 * a.W && ((k - 2 * a.W) & 1)) return;
  float sum = 0.f;
  // ring positions (yr, xr) outside the image with clamp(yr, xr) == (y, x)
  for (int dy = -1; dy <= 1; ++dy) {
    for (int dx = -1; dx <= 1; ++dx) {
      const int yr = y + dy, xr = x + dx;
      if (yr >= 0 && yr < a.H && xr >= 0 && xr < a.W) continue;                 // inside: covered by the main pass
      if (min(max(yr, 0), a.H - 1) != y || min(max(xr, 0), a.W - 1) != x) continue;
      // gxp[yr,xr][c] = sum_tap sum_o w[o][c][tap] * g[(yr,xr) - off(tap)][o]
      for (int tap = 0; tap < 9; ++tap) {
        const int py = yr - (tap / 3 - 1), px = xr - (tap % 3 - 1);
        if (py < 0 || py >= a.H || px < 0 || px >= a.W) continue;
        const float* gp = a.g + ((size_t)(b * a.H + py) * a.W + px) * a.g_cstride + a.g_coff;
        for (int o = 0; o < a.cout; ++o) sum = fmaf(__ldg(a.w_oihw + ((size_t)o * a.cin_total + a.c0 + c) * 9 + tap), __ldg(gp + o), sum);
      }
    }
  }
  const size_t pix = (size_t)(b * a.H + y) * a.W + x;
  if (a.mask && !(__ldg(a.mask + pix * a.gx_cstride + a.gx_coff + c) > 0.f)) return;
  a.gx[pix * a.gx_cstride + a.gx_coff + c] += sum;
}
int launch_dgrad_ring(const RingArgs& a, cudaStream_t st) {
  const int nborder = 2 * a.W + 2 * std::max(a.H - 2, 0);
  const int n = a.B * nborder * a.nch;
  if (n <= 0) return TMG_OK;
  dgrad_ring_kernel<<<cdiv(n, 128), 128, 0, st>>>(a, nborder);
  TMG_LAUNCH_CHECK();
  return TMG_OK;
}

}  // namespace tmg
