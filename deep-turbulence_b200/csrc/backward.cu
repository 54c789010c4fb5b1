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
constexpr int kWgOB = 32, kWgCB = 8, kWgThreads = 128, kWgMaxPx = 256;   // thread: 2 output channels x 1 input channel x 9 taps

__global__ void __launch_bounds__(kWgThreads)
wgrad_kernel(WgradArgs a, int TR, int nsplit, float* part, float* part_b) {
  extern __shared__ __align__(16) float smem[];
  const int S = a.stride == 2 ? 2 : 1;
  const int Hin = S == 2 ? a.Hin : a.H, Win = S == 2 ? a.Win : a.W;
  const int tw = Win + 2;
  float* g_s = smem;                                   // [TR*W][16]
  float* x_s = smem + (size_t)TR * a.W * kWgOB;        // [S*(TR-1)+3][Win+2][8]
  const int tid = threadIdx.x, o_l = tid & 15, c_l = tid >> 4;
  const int c0 = blockIdx.x * kWgCB, o0 = blockIdx.y * kWgOB, split = blockIdx.z;
  const int rtiles = (a.H + TR - 1) / TR, units = a.B * rtiles;
  float acc[9], acc2[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) { acc[t] = 0.f; acc2[t] = 0.f; }
  float accb = 0.f, accb2 = 0.f;
  const bool two = o0 + 16 < a.cout;                  // warp-uniform: this CTA has a second half of output channels
  for (int u = split; u < units; u += nsplit) {
    const int b = u / rtiles, r0 = (u - b * rtiles) * TR;
    const int rows = min(TR, a.H - r0);
    const int xrows = S * (rows - 1) + 3;
    __syncthreads();
    for (int i = tid; i < rows * a.W * kWgOB; i += kWgThreads) {
      const int o = i & (kWgOB - 1), p = i / kWgOB;
      const int r = r0 + p / a.W, cc = p % a.W;
      g_s[i] = (o0 + o < a.cout) ? __ldg(a.g + ((size_t)(b * a.H + r) * a.W + cc) * a.g_cstride + a.g_coff + o0 + o) : 0.f;
    }
    for (int i = tid; i < xrows * tw * kWgCB; i += kWgThreads) {
      const int c = i & 7, q = i >> 3;
      int r = S * r0 - 1 + q / tw, cc = q % tw - 1;
      bool inb = r >= 0 && r < Hin && cc >= 0 && cc < Win;
      if (a.pad_replicate) { r = min(max(r, 0), Hin - 1); cc = min(max(cc, 0), Win - 1); inb = true; }
      float v = 0.f;
      int ch = c0 + c;
      const int chc = ch;
      if (inb && ch < a.cin) {
        const ConvSrc* s = &a.src[0];
        if (ch >= s->nch && a.nsrc > 1) { ch -= s->nch; s = &a.src[1];
          if (ch >= s->nch && a.nsrc > 2) { ch -= s->nch; s = &a.src[2]; } }
        if (ch < s->nch && s->p) {
          v = __ldg(s->p + ((size_t)((s->bshared ? 0 : b) * Hin + r) * Win + cc) * s->cstride + s->coff + ch);
          if (a.bn_scale) v = fmaf(v, __ldg(a.bn_scale + chc), __ldg(a.bn_shift + chc));
          if (s->relu) v = fmaxf(v, 0.f);
        }
      }
      x_s[i] = v;
    }
    __syncthreads();
    for (int pr = 0; pr < rows; ++pr) {
      const float* gp = g_s + (size_t)pr * a.W * kWgOB + o_l;
      const float* xp = x_s + (size_t)(S * pr) * tw * kWgCB + c_l;
      if (two) {
        for (int pc = 0; pc < a.W; ++pc) {
          const float gv = gp[pc * kWgOB], gv2 = gp[pc * kWgOB + 16];
          accb += gv; accb2 += gv2;
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            const float xv = xp[((t / 3) * tw + S * pc + t % 3) * kWgCB];
            acc[t] = fmaf(gv, xv, acc[t]);
            acc2[t] = fmaf(gv2, xv, acc2[t]);
          }
        }
      } else {
        for (int pc = 0; pc < a.W; ++pc) {
          const float gv = gp[pc * kWgOB];
          accb += gv;
#pragma unroll
          for (int t = 0; t < 9; ++t) acc[t] = fmaf(gv, xp[((t / 3) * tw + S * pc + t % 3) * kWgCB], acc[t]);
        }
      }
    }
  }
  if (c0 + c_l < a.cin) {
    if (o0 + o_l < a.cout) {
      float* pp = part + (((size_t)split * a.cout + o0 + o_l) * a.cin + c0 + c_l) * 9;
#pragma unroll
      for (int t = 0; t < 9; ++t) pp[t] = acc[t];
    }
    if (o0 + o_l + 16 < a.cout) {
      float* pp = part + (((size_t)split * a.cout + o0 + o_l + 16) * a.cin + c0 + c_l) * 9;
#pragma unroll
      for (int t = 0; t < 9; ++t) pp[t] = acc2[t];
    }
  }
  if (part_b && blockIdx.x == 0 && c_l == 0) {
    if (o0 + o_l < a.cout) part_b[(size_t)split * a.cout + o0 + o_l] = accb;
    if (o0 + o_l + 16 < a.cout) part_b[(size_t)split * a.cout + o0 + o_l + 16] = accb2;
  }
}

// out[i] (+)= sum_s part[s][i], fixed order
__global__ void reduce_partials_par_kernel(const float* __restrict__ part, float* __restrict__ out, int n, int nsplit, int accum);
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
  const int S = a.stride == 2 ? 2 : 1, Win = S == 2 ? a.Win : a.W;
  const size_t smem = ((size_t)TR * a.W * kWgOB + (size_t)(S * (TR - 1) + 3) * (Win + 2) * kWgCB) * sizeof(float);
  if (smem > 96 * 1024) { set_error("wgrad: tile needs %zu B of shared memory", smem); return TMG_ERR_UNSUPPORTED; }
  TMG_SMEM_ATTR(wgrad_kernel, 96 * 1024);
  float* part = a.scratch;
  float* part_b = a.gbias ? a.scratch + (size_t)ns * a.cout * a.cin * 9 : nullptr;
  dim3 grid(cdiv(a.cin, kWgCB), cdiv(a.cout, kWgOB), ns);
  wgrad_kernel<<<grid, kWgThreads, smem, st>>>(a, TR, ns, part, part_b);
  TMG_LAUNCH_CHECK();
  const int n = a.cout * a.cin * 9;
  reduce_partials_par_kernel<<<cdiv(n, 16), 256, 0, st>>>(part, a.gw, n, ns, a.accum);
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
  if (i >= a.B * nborder * a.cin_total) return;
  const int c = i % a.cin_total; int t = i / a.cin_total; const int k = t % nborder, b = t / nborder;
  // destination of concatenated input channel c
  int d = 0, cl = c;
  while (d < a.ndst - 1 && cl >= a.dst[d].nch) { cl -= a.dst[d].nch; ++d; }
  const ConvDst& ds = a.dst[d];
  if (ds.p == nullptr || cl >= ds.nch) return;
  // border pixel k: top row, bottom row, then the left / right columns without the corners
  int y, x;
  if (k < a.W) { y = 0; x = k; }
  else if (k < 2 * a.W) { y = a.H - 1; x = k - a.W; }
  else { const int r = k - 2 * a.W; y = 1 + (r >> 1); x = (r & 1) ? a.W - 1 : 0; }
  if (a.H == 1 && k >= a.W) return;                       // single row: the top row already covers it
  if (a.W == 1 && k >= 2 * a.W && ((k - 2 * a.W) & 1)) return;
  const size_t pix = (size_t)(b * a.H + y) * a.W + x;
  const size_t o_dst = pix * ds.cstride + ds.coff + cl;
  if (ds.mask && !(__ldg(ds.mask + o_dst) > 0.f)) return;
  float sum = 0.f;
  // ring positions (yr, xr) outside the image with clamp(yr, xr) == (y, x): rows {y, -1 if y == 0, H if y == H-1} x columns
  // likewise, minus (y, x) itself -- enumerated directly (at most 8, typically 1 or 3), and for each only the output pixels
  // p = (yr, xr) - off(tap) that lie inside the image (the first version walked 9 x 9 candidates with bounds checks: 2 000
  // warp instructions at 14 active lanes, ncu r01h_ring)
  int yc[3], xc[3], ny = 0, nx = 0;
  yc[ny++] = y; if (y == 0) yc[ny++] = -1; if (y == a.H - 1) yc[ny++] = a.H;
  xc[nx++] = x; if (x == 0) xc[nx++] = -1; if (x == a.W - 1) xc[nx++] = a.W;
  for (int iy = 0; iy < ny; ++iy) {
    for (int ix = 0; ix < nx; ++ix) {
      if (iy == 0 && ix == 0) continue;                                        // (y, x) itself: covered by the main pass
      const int yr = yc[iy], xr = xc[ix];
      // gxp[yr,xr][c] = sum_tap sum_o w[o][c][tap] * g[(yr,xr) - off(tap)][o],  tap = (yr - py + 1, xr - px + 1)
      for (int py = max(0, yr - 1); py <= min(a.H - 1, yr + 1); ++py) {
        for (int px = max(0, xr - 1); px <= min(a.W - 1, xr + 1); ++px) {
          const int tap = (yr - py + 1) * 3 + (xr - px + 1);
          const float* gp = a.g + ((size_t)(b * a.H + py) * a.W + px) * a.g_cstride + a.g_coff;
          const float* wp = a.w_oihw + (size_t)c * 9 + tap;
          for (int o = 0; o < a.cout; ++o) sum = fmaf(__ldg(wp + (size_t)o * a.cin_total * 9), __ldg(gp + o), sum);
        }
      }
    }
  }
  ds.p[o_dst] += sum;
}
int launch_dgrad_ring(const RingArgs& a, cudaStream_t st) {
  const int nborder = 2 * a.W + 2 * std::max(a.H - 2, 0);
  const int n = a.B * nborder * a.cin_total;
  if (n <= 0) return TMG_OK;
  dgrad_ring_kernel<<<cdiv(n, 128), 128, 0, st>>>(a, nborder);
  TMG_LAUNCH_CHECK();
  return TMG_OK;
}

// ------------------------------------------------------------------ data gradient of a Cout = 1 convolution (dense layers)
// gx[p][c] = sum_tap g[p - off(tap)] * w[0][c][tap]: nine FMAs per element -- a streaming kernel, exact fp32; on the tensor
// cores this is a K = 1 GEMM padded to 16 and costs a full pipeline trip.  Zero padding only (denseBlock.py:136).  The
// channels are routed to up to three destinations (the sources of the forward concatenation), gated by the ReLU of the
// forward input, optionally accumulated.
struct DgradC1Args {
  const float* g; int g_cstride, g_coff;
  const float* w;                  // OIHW with O = 1: [cin][9]
  int cin;
  int ndst; ConvDst dst[3];
  int B, H, W;
};
__global__ void __launch_bounds__(256)
dgrad_cout1_kernel(DgradC1Args a, int npix) {
  // a thread owns one pixel and every eighth channel: the nine taps of g are loaded once, the weights come from shared
  // memory, consecutive lanes write consecutive channels
  __shared__ float s_w[9 * 256];
  for (int i = threadIdx.x; i < a.cin * 9; i += 256) s_w[i] = __ldg(a.w + i);
  __syncthreads();
  const int pix = blockIdx.x * 32 + (threadIdx.x >> 3), l = threadIdx.x & 7;
  if (pix >= npix) return;
  const int x = pix % a.W; const int t = pix / a.W; const int y = t % a.H; const int b = t / a.H;
  float gt[9];
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    // forward: out[q] += w[tap] * in[q + off(tap)]  =>  gx[p] += w[tap] * g[p - off(tap)]
    const int py = y - (tap / 3 - 1), px = x - (tap % 3 - 1);
    gt[tap] = (py >= 0 && py < a.H && px >= 0 && px < a.W)
                  ? __ldg(a.g + ((size_t)(b * a.H + py) * a.W + px) * a.g_cstride + a.g_coff) : 0.f;
  }
  const int nb1 = a.dst[0].nch, nb2 = nb1 + (a.ndst > 1 ? a.dst[1].nch : 0);
  for (int c = l; c < a.cin; c += 8) {
    const int d = c < nb1 ? 0 : (c < nb2 || a.ndst < 3 ? 1 : 2);
    const ConvDst& ds = a.dst[d];
    const int cl = c - (d == 0 ? 0 : (d == 1 ? nb1 : nb2));
    if (ds.p == nullptr || cl >= ds.nch) continue;
    const size_t o = (size_t)pix * ds.cstride + ds.coff + cl;
    float acc = 0.f;
    if (!ds.mask || __ldg(ds.mask + o) > 0.f) {
      const float* wc = s_w + c * 9;
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) acc = fmaf(wc[tap], gt[tap], acc);
    }
    ds.p[o] = ds.accum ? ds.p[o] + acc : acc;
  }
}
int launch_dgrad_cout1(const float* g, int g_cstride, int g_coff, const float* w_oihw, int cin, const ConvDst* dst, int ndst,
                       int B, int H, int W, cudaStream_t st) {
  DgradC1Args a{};
  a.g = g; a.g_cstride = g_cstride; a.g_coff = g_coff; a.w = w_oihw; a.cin = cin;
  a.ndst = ndst;
  for (int d = 0; d < ndst && d < 3; ++d) a.dst[d] = dst[d];
  a.B = B; a.H = H; a.W = W;
  const int64_t npix = (int64_t)B * H * W;
  if (npix <= 0 || cin <= 0) return TMG_OK;
  if (cin > 256 || npix > 0x7fffffff) { set_error("dgrad_cout1: %d input channels / %lld pixels", cin, (long long)npix); return TMG_ERR_UNSUPPORTED; }
  dgrad_cout1_kernel<<<(unsigned)((npix + 31) / 32), 256, 0, st>>>(a, (int)npix);
  TMG_LAUNCH_CHECK();
  return TMG_OK;
}

// ------------------------------------------------------------------ weight gradient of a Cout = 1 convolution (dense layers)
// dW[c][tap] = sum_q relu?(x[q][c]) * g[q - off(tap)]: each input element is loaded once and meets the nine output-gradient
// values around it (g has ONE channel: the loads are warp-uniform).  Exact fp32; a block sums a pixel range in a fixed
// order, reduce_partials_kernel adds the blocks in order (deterministic).  Zero padding, no bias (denseBlock.py:136).
constexpr int kWc1Lanes = 4, kWc1CW = 64, kWc1Rows = 4;      // 256 threads = 4 pixel lanes x 64 channel slots; 4 rows per block (measured: 8 -> 14.9 us, 4 -> 10.3 us, 2 -> 9.8 us per launch but a costlier reduce)
__global__ void __launch_bounds__(kWc1Lanes * kWc1CW)
wgrad_cout1_kernel(WgradArgs a, int chunks, float* __restrict__ part) {
  extern __shared__ float s_g[];                       // [(rows + 2)][W + 2] output gradient of this row chunk, zero halo
  __shared__ float s_acc[kWc1Lanes][kWc1CW][9];
  const int c = threadIdx.x % kWc1CW, lane = threadIdx.x / kWc1CW;
  const int b = blockIdx.x / chunks, r0 = (blockIdx.x - b * chunks) * kWc1Rows;
  const int rows = min(kWc1Rows, a.H - r0), gw = a.W + 2;
  for (int i = threadIdx.x; i < (rows + 2) * gw; i += kWc1Lanes * kWc1CW) {
    const int ry = i / gw, rx = i - ry * gw;
    const int y = r0 - 1 + ry, x = rx - 1;
    s_g[i] = (y >= 0 && y < a.H && x >= 0 && x < a.W) ? __ldg(a.g + ((size_t)(b * a.H + y) * a.W + x) * a.g_cstride + a.g_coff) : 0.f;
  }
  float acc[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) acc[t] = 0.f;
  // source of concatenated channel c
  const float* sp = nullptr; int scs = 0, srelu = 0;
  if (c < a.cin) {
    int cc = c, si = 0;
    while (si < a.nsrc - 1 && cc >= a.src[si].nch) { cc -= a.src[si].nch; ++si; }
    if (cc < a.src[si].nch && a.src[si].p) {
      sp = a.src[si].p + a.src[si].coff + cc + (size_t)(a.src[si].bshared ? 0 : b) * a.H * a.W * a.src[si].cstride;
      scs = a.src[si].cstride; srelu = a.src[si].relu;
    }
  }
  __syncthreads();
  if (sp) {
    for (int q = lane; q < rows * a.W; q += kWc1Lanes) {
      const int ly = q / a.W, lx = q - ly * a.W;
      float xv = __ldg(sp + (size_t)((r0 + ly) * a.W + lx) * scs);
      if (srelu) xv = fmaxf(xv, 0.f);
      // forward: out[p] += w[tap] * in[p + off(tap)]  =>  dW[tap] += g[p] * in[q], p = q - off(tap)
      const float* gp = s_g + (ly + 1) * gw + (lx + 1);
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) acc[tap] = fmaf(xv, gp[-(tap / 3 - 1) * gw - (tap % 3 - 1)], acc[tap]);
    }
  }
#pragma unroll
  for (int t = 0; t < 9; ++t) s_acc[lane][c][t] = acc[t];
  __syncthreads();
  for (int i = threadIdx.x; i < a.cin * 9; i += kWc1Lanes * kWc1CW) {
    const int cc = i / 9, t = i - cc * 9;
    part[(size_t)blockIdx.x * a.cin * 9 + i] = (s_acc[0][cc][t] + s_acc[1][cc][t]) + (s_acc[2][cc][t] + s_acc[3][cc][t]);
  }
}
// out[i] (+)= sum over shares of part[s][i]: 32 outputs per block, 8 threads per output walk the shares 8 apart, then a
// fixed-order sum of the 8 partials (deterministic)
__global__ void __launch_bounds__(256)
reduce_partials_par_kernel(const float* __restrict__ part, float* __restrict__ out, int n, int nsplit, int accum) {
  // 16 outputs x 16 share-lanes per block, two independent chains per thread
  __shared__ float sm[16][16];
  const int x = threadIdx.x & 15, y = threadIdx.x >> 4;
  const int i = blockIdx.x * 16 + x;
  float s0 = 0.f, s1 = 0.f;
  if (i < n) {
    int k = y;
    for (; k + 16 < nsplit; k += 32) { s0 += part[(size_t)k * n + i]; s1 += part[(size_t)(k + 16) * n + i]; }
    if (k < nsplit) s0 += part[(size_t)k * n + i];
  }
  sm[y][x] = s0 + s1;
  __syncthreads();
  if (y == 0 && i < n) {
    float t = sm[0][x];
#pragma unroll
    for (int q = 1; q < 16; ++q) t += sm[q][x];
    out[i] = accum ? out[i] + t : t;
  }
}
bool wgrad_cout1_supported(const WgradArgs& a) {
  return a.cout == 1 && a.cin <= kWc1CW && a.stride != 2 && !a.bn_scale && !a.pad_replicate && !a.gbias &&
         (size_t)(kWc1Rows + 2) * (a.W + 2) * sizeof(float) <= 32 * 1024 && (int64_t)a.B * a.H * a.W < 0x7fffffff;
}
size_t wgrad_cout1_scratch_floats(int cin, int B, int H, int W) { (void)W; return (size_t)B * cdiv(H, kWc1Rows) * cin * 9 + 64; }
int launch_wgrad_cout1(const WgradArgs& a, cudaStream_t st) {
  if (a.B <= 0 || a.H <= 0 || a.W <= 0) return TMG_OK;
  const int chunks = cdiv(a.H, kWc1Rows), nb = a.B * chunks;
  const size_t smem = (size_t)(kWc1Rows + 2) * (a.W + 2) * sizeof(float);
  wgrad_cout1_kernel<<<nb, kWc1Lanes * kWc1CW, smem, st>>>(a, chunks, a.scratch);
  TMG_LAUNCH_CHECK();
  const int n = a.cin * 9;
  reduce_partials_par_kernel<<<cdiv(n, 16), 256, 0, st>>>(a.scratch, a.gw, n, nb, a.accum);
  TMG_LAUNCH_CHECK();
  return TMG_OK;
}

// ------------------------------------------------------------------ pointwise step backward
constexpr int kSbThreads = 128;

template <int C>
__global__ void __launch_bounds__(kSbThreads)
step_bwd_kernel(StepBwdArgs a) {
  __shared__ __align__(16) float s_w[C * C];
  __shared__ float s_nw[C], s_nb[C];
  __shared__ float s_red[kSbThreads / 32][3 * C + 2];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int b = blockIdx.y;
  const int p = blockIdx.x * kSbThreads + tid;
  for (int i = tid; i < C * C; i += kSbThreads) s_w[i] = a.wmat ? __ldg(a.wmat + i) : ((i / C) == (i % C) ? 1.f : 0.f);
  for (int i = tid; i < C; i += kSbThreads) { s_nw[i] = a.nw ? __ldg(a.nw + i) : 1.f; s_nb[i] = a.nw ? __ldg(a.nb + i) : 0.f; }
  __syncthreads();
  const float gain = __ldg(a.gain), gld = __ldg(a.g_ld + b);
  float acc_nb[C], acc_nw[C], acc_gain = 0.f;
  float gz[C];                  // gradient w.r.t. the raw Conv2dZeros output: its column sums are the bias gradient
#pragma unroll
  for (int i = 0; i < C; ++i) { acc_nb[i] = 0.f; acc_nw[i] = 0.f; gz[i] = 0.f; }
  if (p < a.HW) {
    const size_t pix = (size_t)b * a.HW + p;
    float y[C], h[C], v[C], gu[C];
    const float4* y4 = reinterpret_cast<const float4*>(a.y_in + pix * C);
    const float4* h4 = reinterpret_cast<const float4*>(a.h + pix * C);
    const float4* g4 = reinterpret_cast<const float4*>(a.g_out + pix * C);
#pragma unroll
    for (int i = 0; i < C / 4; ++i) {
      float4 t = __ldg(y4 + i); y[4 * i] = t.x; y[4 * i + 1] = t.y; y[4 * i + 2] = t.z; y[4 * i + 3] = t.w;
      t = __ldg(h4 + i); h[4 * i] = t.x; h[4 * i + 1] = t.y; h[4 * i + 2] = t.z; h[4 * i + 3] = t.w;
      t = __ldg(g4 + i); gu[4 * i] = t.x; gu[4 * i + 1] = t.y; gu[4 * i + 2] = t.z; gu[4 * i + 3] = t.w;
    }
    float ea[C / 2], dadr[C / 2];
#pragma unroll
    for (int j = 0; j < C / 2; ++j) {
      const float raw = h[2 * j + 1], den = 1.f + fabsf(raw);
      const float la = 2.f * (raw / den);
      ea[j] = expf(-la);
      dadr[j] = 2.f / (den * den);                 // d(2*softsign)/d raw
      v[j] = y[j];
      v[C / 2 + j] = fmaf(y[C / 2 + j], ea[j], -h[2 * j]);
    }
    // u = W v, out = (u - nb)/nw; g_u = g_out / nw
#pragma unroll
    for (int c = 0; c < C; ++c) gu[c] = gu[c] / s_nw[c];
    if (a.nw) {
#pragma unroll 1
      for (int c = 0; c < C; ++c) {
        // four independent partial sums and 16-byte weight loads: this loop is a chain of dependent FMAs otherwise
        float u0 = 0.f, u1 = 0.f, u2 = 0.f, u3 = 0.f;
        const float4* w4 = reinterpret_cast<const float4*>(s_w + c * C);
#pragma unroll
        for (int k = 0; k < C / 4; ++k) {
          const float4 w = w4[k];
          u0 = fmaf(w.x, v[4 * k], u0); u1 = fmaf(w.y, v[4 * k + 1], u1);
          u2 = fmaf(w.z, v[4 * k + 2], u2); u3 = fmaf(w.w, v[4 * k + 3], u3);
        }
        const float u = (u0 + u1) + (u2 + u3);
        const float out = (u - s_nb[c]) / s_nw[c];
        acc_nb[c] = -gu[c];
        acc_nw[c] = -gu[c] * out;
      }
    }
    float gv[C];
#pragma unroll
    for (int k = 0; k < C; ++k) gv[k] = 0.f;
#pragma unroll 1
    for (int c = 0; c < C; ++c) {
      const float g = gu[c];
      const float4* w4 = reinterpret_cast<const float4*>(s_w + c * C);
#pragma unroll
      for (int k = 0; k < C / 4; ++k) {
        const float4 w = w4[k];
        gv[4 * k] = fmaf(w.x, g, gv[4 * k]); gv[4 * k + 1] = fmaf(w.y, g, gv[4 * k + 1]);
        gv[4 * k + 2] = fmaf(w.z, g, gv[4 * k + 2]); gv[4 * k + 3] = fmaf(w.w, g, gv[4 * k + 3]);
      }
    }
    float4* gu4 = reinterpret_cast<float4*>(a.gu + pix * C);
    float4* v4 = reinterpret_cast<float4*>(a.v + pix * C);
#pragma unroll
    for (int i = 0; i < C / 4; ++i) {
      gu4[i] = make_float4(gu[4 * i], gu[4 * i + 1], gu[4 * i + 2], gu[4 * i + 3]);
      v4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    }
    float gy[C];
#pragma unroll
    for (int j = 0; j < C / 2; ++j) {
      const float g2 = gv[C / 2 + j];
      gy[j] = gv[j];
      gy[C / 2 + j] = g2 * ea[j];
      const float g_shift = -g2;
      const float g_a = fmaf(-g2 * y[C / 2 + j], ea[j], gld);
      const float g_raw = g_a * dadr[j];
      acc_gain += (g_shift * h[2 * j] + g_raw * h[2 * j + 1]) / gain;
      gz[2 * j] = g_shift * gain;
      gz[2 * j + 1] = g_raw * gain;
    }
    float4* gy4 = reinterpret_cast<float4*>(a.g_y + pix * C);
    float4* gz4 = reinterpret_cast<float4*>(a.g_z + pix * C);
#pragma unroll
    for (int i = 0; i < C / 4; ++i) {
      gy4[i] = make_float4(gy[4 * i], gy[4 * i + 1], gy[4 * i + 2], gy[4 * i + 3]);
      gz4[i] = make_float4(gz[4 * i], gz[4 * i + 1], gz[4 * i + 2], gz[4 * i + 3]);
    }
  }
  // CTA partial sums in a fixed order: warp shuffle tree, then the four warps in order
#pragma unroll
  for (int i = 0; i < 2 * C + 1; ++i) {
    float x = i < C ? acc_nb[i < C ? i : 0] : (i < 2 * C ? acc_nw[i < 2 * C ? (i >= C ? i - C : 0) : 0] : acc_gain);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) s_red[wid][i] = x;
  }
  // columns [2C+1, 3C+1): sum of gz (bias gradient of the Conv2dZeros conv); column 3C+1: max |gz| (the power-of-two scale
  // the tensor-core weight / data gradients stage gz with) -- saves a separate pass over gz
  float gmax = 0.f;
#pragma unroll
  for (int i = 0; i < C; ++i) {
    float x = gz[i];
    gmax = fmaxf(gmax, fabsf(x));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) s_red[wid][2 * C + 1 + i] = x;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) gmax = fmaxf(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
  if (lane == 0) s_red[wid][3 * C + 1] = gmax;
  __syncthreads();
  float* pp = a.part + ((size_t)b * gridDim.x + blockIdx.x) * (3 * C + 2);
  for (int i = tid; i < 3 * C + 1; i += kSbThreads) pp[i] = (s_red[0][i] + s_red[1][i]) + (s_red[2][i] + s_red[3][i]);
  if (tid == 0) pp[3 * C + 1] = fmaxf(fmaxf(s_red[0][3 * C + 1], s_red[1][3 * C + 1]), fmaxf(s_red[2][3 * C + 1], s_red[3][3 * C + 1]));
}

int step_bwd_blocks(int B, int HW) { return B * cdiv(HW, kSbThreads); }

int launch_step_bwd(const StepBwdArgs& a, cudaStream_t st) {
  dim3 grid(cdiv(a.HW, kSbThreads), a.B);
  switch (a.C) {
#define TMG_CASE(CC) case CC: step_bwd_kernel<CC><<<grid, kSbThreads, 0, st>>>(a); break;
    TMG_CASE(4) TMG_CASE(8) TMG_CASE(12) TMG_CASE(16) TMG_CASE(24) TMG_CASE(32) TMG_CASE(48)
#undef TMG_CASE
    default:
      set_error("step backward: %d channels not supported", a.C);
      return TMG_ERR_UNSUPPORTED;
  }
  TMG_LAUNCH_CHECK();
  return TMG_OK;
}

// out[c] (+)= sum over rows (fixed order) of part[row*row_stride + col0 + c]
// out[c] (+)= sum over rows of part[r][col0 + c]: one block per column, rows strided over the threads, fixed-order tree
__device__ __forceinline__ double block_col_sum(const float* __restrict__ part, int nrows, int row_stride, int col) {
  __shared__ double sm[128];
  double s = 0.0;
  for (int r = threadIdx.x; r < nrows; r += 128) s += (double)part[(size_t)r * row_stride + col];
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int o = 64; o; o >>= 1) {
    if ((int)threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  return sm[0];
}
__global__ void __launch_bounds__(128)
reduce_cols_kernel(const float* __restrict__ part, int nrows, int row_stride, int col0, int ncols,
                   float* __restrict__ out, int accum) {
  const int c = blockIdx.x;
  const double s = block_col_sum(part, nrows, row_stride, col0 + c);
  if (threadIdx.x == 0) out[c] = accum ? out[c] + (float)s : (float)s;
}
int launch_reduce_cols(const float* part, int nrows, int row_stride, int col0, int ncols, float* out, int accum, cudaStream_t st) {
  if (ncols <= 0) return TMG_OK;
  reduce_cols_kernel<<<ncols, 128, 0, st>>>(part, nrows, row_stride, col0, ncols, out, accum);
  TMG_LAUNCH_CHECK();
  return TMG_OK;
}
// The per-CTA partial sums of step_bwd_kernel [nrows][2C+1] -> ActNorm bias / weight gradients (columns [0,C), [C,2C),
// accumulated, skipped when null) and the Conv2dZeros log-scale gradient (column 2C: d/dscale of exp(clamp(scale,-4,ln4)),
// flowUtils.py:247), in ONE launch.
__global__ void __launch_bounds__(128)
step_param_grads_kernel(const float* __restrict__ part, int nrows, int C, float* __restrict__ g_nb, float* __restrict__ g_nw,
                        const float* __restrict__ scale_param, float* __restrict__ g_scale, const float* __restrict__ g_ld, int B,
                        float hw, float* __restrict__ gld_stash, float* __restrict__ g_bias, float* __restrict__ gz_scale) {
  const int c = blockIdx.x, stride = 3 * C + 2;
  if (c == 3 * C + 2) {       // deferred LU backward: hw * sum_b g_ld[b] accumulated over the time steps of the block
    if (gld_stash && threadIdx.x < 32) {
      double s = 0.0;
      for (int b = threadIdx.x; b < B; b += 32) s += (double)g_ld[b];
#pragma unroll
      for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (threadIdx.x == 0) *gld_stash += hw * (float)s;
    }
    return;
  }
  if (c == 3 * C + 1) {       // max |gz| over the CTAs -> power-of-two scale (same rule as launch_absmax_scale)
    if (!gz_scale) return;
    __shared__ float sm[128];
    float m = 0.f;
    for (int r = threadIdx.x; r < nrows; r += 128) m = fmaxf(m, part[(size_t)r * stride + c]);
    sm[threadIdx.x] = m;
    __syncthreads();
    for (int o = 64; o; o >>= 1) {
      if ((int)threadIdx.x < o) sm[threadIdx.x] = fmaxf(sm[threadIdx.x], sm[threadIdx.x + o]);
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      const float mx = sm[0];
      int ex = 0;
      float sc = 1.f;
      if (mx > 0.f && mx < 3.0e38f) { frexpf(mx, &ex); sc = ldexpf(1.f, min(max(11 - ex, -100), 100)); }
      gz_scale[0] = sc; gz_scale[1] = 1.f / sc;
    }
    return;
  }
  if (c < 2 * C && g_nb == nullptr) return;
  if (c > 2 * C && g_bias == nullptr) return;
  const double s = block_col_sum(part, nrows, stride, c);
  if (threadIdx.x != 0) return;
  if (c < C) g_nb[c] += (float)s;
  else if (c < 2 * C) g_nw[c - C] += (float)s;
  else if (c == 2 * C) {
    const float sp = *scale_param;
    if (sp > -4.f && sp < kLog4) *g_scale += (float)s * expf(sp);
  } else g_bias[c - 2 * C - 1] += (float)s;
}
int launch_step_param_grads(const float* part, int nrows, int C, float* g_nb, float* g_nw, const float* scale_param, float* g_scale,
                            const float* g_ld, int B, float hw, float* gld_stash, float* g_bias, float* gz_scale, cudaStream_t st) {
  step_param_grads_kernel<<<3 * C + 3, 128, 0, st>>>(part, nrows, C, g_nb, g_nw, scale_param, g_scale, g_ld, B, hw, gld_stash,
                                                     g_bias, gz_scale);
  TMG_LAUNCH_CHECK();
  return TMG_OK;
}

// ------------------------------------------------------------------ 1x1 weight gradient: gw[o][k] = sum_px gu[px][o] v[px][k]
constexpr int kOwThreads = 256, kOwTile = 32;
__global__ void __launch_bounds__(kOwThreads)
outer_wgrad_kernel(const float* __restrict__ gu, const float* __restrict__ v, int64_t npix, int C, int nsplit, float* part) {
  extern __shared__ float sm[];
  float* g_s = sm;                 // [tile][C]
  float* v_s = sm + kOwTile * C;
  const int tid = threadIdx.x, split = blockIdx.x;
  const int npairs = C * C;
  float acc[9];                    // C <= 48: at most 9 (o,k) pairs per thread
#pragma unroll
  for (int i = 0; i < 9; ++i) acc[i] = 0.f;
  const int64_t ntiles = (npix + kOwTile - 1) / kOwTile;
  for (int64_t t = split; t < ntiles; t += nsplit) {
    const int64_t p0 = t * kOwTile;
    const int np = (int)min((int64_t)kOwTile, npix - p0);
    __syncthreads();
    for (int i = tid; i < np * C; i += kOwThreads) { g_s[i] = __ldg(gu + p0 * C + i); v_s[i] = __ldg(v + p0 * C + i); }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      const int pr = tid + i * kOwThreads;
      if (pr < npairs) {
        const int o = pr / C, k = pr - o * C;
        float s = acc[i];
        for (int q = 0; q < np; ++q) s = fmaf(g_s[q * C + o], v_s[q * C + k], s);
        acc[i] = s;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const int pr = tid + i * kOwThreads;
    if (pr < npairs) part[(size_t)split * npairs + pr] = acc[i];
  }
}
static int outer_splits(int64_t npix) { return (int)std::max<int64_t>(1, std::min<int64_t>(296, (npix + kOwTile - 1) / kOwTile)); }
size_t outer_wgrad_scratch_floats(int64_t npix, int C) { return (size_t)outer_splits(npix) * C * C + 64; }
int launch_outer_wgrad(const float* gu, const float* v, int64_t npix, int C, float* gw, float* scratch, cudaStream_t st, int accum) {
  if (C * C > 9 * kOwThreads) { set_error("1x1 weight gradient: C=%d not supported", C); return TMG_ERR_UNSUPPORTED; }
  const int ns = outer_splits(npix);
  outer_wgrad_kernel<<<ns, kOwThreads, 2 * kOwTile * C * sizeof(float), st>>>(gu, v, npix, C, ns, scratch);
  TMG_LAUNCH_CHECK();
  reduce_partials_par_kernel<<<cdiv(C * C, 16), 256, 0, st>>>(scratch, gw, C * C, ns, accum);
  TMG_LAUNCH_CHECK();
  return TMG_OK;
}

// ------------------------------------------------------------------ LU parameterisation + log-det constants, backward
__global__ void __launch_bounds__(256)
lu_bwd_kernel(LuBwdArgs a) {
  extern __shared__ float sm[];
  const int C = a.C, tid = threadIdx.x;
  float* L = sm; float* U = L + C * C; float* A = U + C * C;
  __shared__ float s_gld;
  for (int i = tid; i < C * C; i += blockDim.x) {
    const int r = i / C, c = i % C;
    L[i] = a.l[i] * a.lmask[i] + a.eye[i];
    U[i] = a.u[i] * a.umask[i] + (r == c ? expf(a.log_s[r]) * a.sign_s[r] : 0.f) + 0.01f * a.eye[i];
  }
  if (tid < 32) {      // sum of the per-sample log-det gradients: one warp, fixed order
    double s = 0.0;
    for (int b = tid; b < a.B; b += 32) s += (double)a.g_ld[b];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (tid == 0) s_gld = (float)s;
  }
  __syncthreads();
  // A = P^T dW
  for (int i = tid; i < C * C; i += blockDim.x) {
    const int r = i / C, c = i % C;
    float s = 0.f;
    for (int q = 0; q < C; ++q) s = fmaf(a.p[q * C + r], a.dW[q * C + c], s);
    A[i] = s;
  }
  __syncthreads();
  for (int i = tid; i < C * C; i += blockDim.x) {
    const int r = i / C, c = i % C;
    float dl = 0.f, du = 0.f;
    for (int j = 0; j < C; ++j) dl = fmaf(A[r * C + j], U[c * C + j], dl);        // dL = A U^T
    for (int q = 0; q < C; ++q) du = fmaf(L[q * C + r], A[q * C + c], du);        // dU = L^T A
    a.g_l[i] += dl * a.lmask[i];
    a.g_u[i] += du * a.umask[i];
    if (r == c) a.g_log_s[r] += du * a.sign_s[r] * expf(a.log_s[r]) - a.hw * s_gld;
  }
  if (a.nw) for (int c = tid; c < C; c += blockDim.x) a.g_nw[c] += a.hw * s_gld / a.nw[c];
}
__global__ void lstm_bwd_kernel(LstmBwdArgs a) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  const int64_t pix = i / a.R;
  const int r = (int)(i - pix * a.R);
  const float* g = a.gates + pix * 4 * a.R;
  const float gi = 1.f / (1.f + expf(-g[r])), gf = 1.f / (1.f + expf(-g[a.R + r]));
  const float go = 1.f / (1.f + expf(-g[2 * a.R + r])), gg = tanhf(g[3 * a.R + r]);
  const float c = a.c_prev ? a.c_prev[i] : 0.f;
  const float cn = gf * c + gi * gg;
  const float tc = tanhf(cn);
  const float gh = a.g_h[i];
  const float gct = (a.g_c ? a.g_c[i] : 0.f) + gh * go * (1.f - tc * tc);
  float* o = a.g_gates + pix * 4 * a.R;
  o[r] = gct * gg * gi * (1.f - gi);
  o[a.R + r] = gct * c * gf * (1.f - gf);
  o[2 * a.R + r] = gh * tc * go * (1.f - go);
  o[3 * a.R + r] = gct * gi * (1.f - gg * gg);
  if (a.g_c_prev) a.g_c_prev[i] = gct * gf;
}
int launch_lstm_bwd(const LstmBwdArgs& a, cudaStream_t st) {
  lstm_bwd_kernel<<<(unsigned)((a.n + 255) / 256), 256, 0, st>>>(a);
  TMG_LAUNCH_CHECK();
  return TMG_OK;
}

__global__ void __launch_bounds__(128)
gauss_bwd_kernel(GaussBwdArgs a) {
  __shared__ float s_red[4];
  const int b = blockIdx.y, p = blockIdx.x * 128 + threadIdx.x;
  float acc = 0.f;
  if (p < a.HW) {
    const size_t pix = (size_t)b * a.HW + p;
    const float* pr = a.prm + pix * a.prm_cstride;
    const float* gv = a.g_val + pix * a.gv_cstride + a.gv_coff;
    float* gp = a.g_prm + pix * a.gp_cstride;
    const float gain = a.gain ? __ldg(a.gain) : 1.f;
    const float gld = a.g_ld ? __ldg(a.g_ld + b) : 0.f;
    for (int j = 0; j < a.n; ++j) {
      const float mu = pr[j], lsr = pr[a.n + j];
      const float ls = fminf(fmaxf(lsr, -10.f), kLog5);
      const float e = __ldg(a.eps + ((size_t)b * a.n + j) * a.HW + p);
      const float g = gv[j];
      float g_mu = g;
      float g_ls = g * expf(ls) * e - gld;                  // d logp / d ls = -1 per element
      if (!(lsr > -10.f && lsr < kLog5)) g_ls = 0.f;        // clamp_ (flowUtils.py:163)
      if (a.hardtanh) {
        if (!(mu > -2.f && mu < kLog5)) g_mu = 0.f;
        if (!(lsr > -2.f && lsr < kLog5)) g_ls = 0.f;
      }
      acc += (g_mu * mu + g_ls * lsr) / gain;
      gp[j] = g_mu * gain;
      gp[a.n + j] = g_ls * gain;
    }
  }
  if (a.part) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) a.part[(size_t)b * gridDim.x + blockIdx.x] = (s_red[0] + s_red[1]) + (s_red[2] + s_red[3]);
  }
}
int gauss_bwd_blocks(int B, int HW) { return B * cdiv(HW, 128); }
int launch_gauss_bwd(const GaussBwdArgs& a, cudaStream_t st) {
  dim3 grid(cdiv(a.HW, 128), a.B);
  gauss_bwd_kernel<<<grid, 128, 0, st>>>(a);
  TMG_LAUNCH_CHECK();
  return TMG_OK;
}

// g_scale += S_gain * d gain/d scale, gain = exp(clamp(scale, -4, ln 4))  (flowUtils.py:247)
__global__ void scale_grad_kernel(const float* s_gain, const float* scale_param, float* g_scale) {
  const float sp = *scale_param;
  if (sp > -4.f && sp < kLog4) *g_scale += *s_gain * expf(sp);
}
int launch_scale_grad(const float* s_gain, const float* scale_param, float* g_scale, cudaStream_t st) {
  scale_grad_kernel<<<1, 1, 0, st>>>(s_gain, scale_param, g_scale);
  TMG_LAUNCH_CHECK();
  return TMG_OK;
}
// Deferred form, once per optimizer step for ALL flow steps (one CTA each): the gradients of l, u, log_s (and the ActNorm
// log-det term) are linear in dW and in hw * sum(g_ld), which the per-time-step backward accumulates in the gradient slots
// of the (non-trainable) buffers `p` and `sign_s`; this kernel consumes and clears those slots.
__global__ void __launch_bounds__(256)
lu_bwd_batched_kernel(const LuTabEntry* __restrict__ tab, const float* __restrict__ P, float* __restrict__ G) {
  extern __shared__ float sm[];
  const LuTabEntry e = tab[blockIdx.x];
  const int C = e.C, tid = threadIdx.x;
  float* L = sm; float* U = L + C * C; float* A = U + C * C; float* DW = A + C * C;
  const float* l = P + e.off[0]; const float* u = P + e.off[1]; const float* log_s = P + e.off[2]; const float* pm = P + e.off[3];
  const float* sign_s = P + e.off[4]; const float* lmask = P + e.off[5]; const float* umask = P + e.off[6]; const float* eye = P + e.off[7];
  float* dWs = G + e.off[3];
  __shared__ float s_hg;
  if (tid == 0) s_hg = G[e.off[4]];
  for (int i = tid; i < C * C; i += blockDim.x) {
    const int r = i / C, c = i % C;
    L[i] = l[i] * lmask[i] + eye[i];
    U[i] = u[i] * umask[i] + (r == c ? expf(log_s[r]) * sign_s[r] : 0.f) + 0.01f * eye[i];
    DW[i] = dWs[i];
  }
  __syncthreads();
  for (int i = tid; i < C * C; i += blockDim.x) dWs[i] = 0.f;           // the stash slots go back to zero gradients
  if (tid == 0) G[e.off[4]] = 0.f;
  for (int i = tid; i < C * C; i += blockDim.x) {                       // A = P^T dW
    const int r = i / C, c = i % C;
    float s = 0.f;
    for (int q = 0; q < C; ++q) s = fmaf(pm[q * C + r], DW[q * C + c], s);
    A[i] = s;
  }
  __syncthreads();
  const float hg = s_hg;
  for (int i = tid; i < C * C; i += blockDim.x) {
    const int r = i / C, c = i % C;
    float dl = 0.f, du = 0.f;
    for (int j = 0; j < C; ++j) dl = fmaf(A[r * C + j], U[c * C + j], dl);        // dL = A U^T
    for (int q = 0; q < C; ++q) du = fmaf(L[q * C + r], A[q * C + c], du);        // dU = L^T A
    G[e.off[0] + i] += dl * lmask[i];
    G[e.off[1] + i] += du * umask[i];
    if (r == c) G[e.off[2] + r] += du * sign_s[r] * expf(log_s[r]) - hg;
  }
  if (e.norm_w >= 0) for (int c = tid; c < C; c += blockDim.x) G[e.norm_w + c] += hg / P[e.norm_w + c];
}
int launch_lu_bwd_batched(const LuTabEntry* tab_dev, int n, int cmax, const float* params, float* grads, cudaStream_t st) {
  if (n <= 0) return TMG_OK;
  const size_t smem = (size_t)4 * cmax * cmax * sizeof(float);
  TMG_SMEM_ATTR(lu_bwd_batched_kernel, (int)smem);
  lu_bwd_batched_kernel<<<n, 256, smem, st>>>(tab_dev, params, grads);
  TMG_LAUNCH_CHECK();
  return TMG_OK;
}
int launch_lu_bwd(const LuBwdArgs& a, cudaStream_t st) {
  lu_bwd_kernel<<<1, 256, 3 * a.C * a.C * sizeof(float), st>>>(a);
  TMG_LAUNCH_CHECK();
  return TMG_OK;
}

// ------------------------------------------------------------------ encoder adjoints
__global__ void upsample_bwd_kernel(const float* __restrict__ gdst, float* __restrict__ gsrc, int B, int h, int w, int C, int f,
                                    float sy, float sx, int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C); int64_t t = i / C; const int x = (int)(t % w); t /= w; const int y = (int)(t % h); const int b = (int)(t / h);
  const int H = h * f, W = w * f;
  const int Y0 = max(0, (y - 1) * f - f), Y1 = min(H - 1, (y + 1) * f + f);
  const int X0 = max(0, (x - 1) * f - f), X1 = min(W - 1, (x + 1) * f + f);
  float sum = 0.f;
  for (int Y = Y0; Y <= Y1; ++Y) {
    const float fy = sy * Y; const int y0 = (int)fy; const int y1 = min(y0 + 1, h - 1); const float ly = fy - y0;
    float wy = 0.f;
    if (y0 == y) wy += 1.f - ly;
    if (y1 == y) wy += ly;
    if (wy == 0.f) continue;
    for (int X = X0; X <= X1; ++X) {
      const float fx = sx * X; const int x0 = (int)fx; const int x1 = min(x0 + 1, w - 1); const float lx = fx - x0;
      float wx = 0.f;
      if (x0 == x) wx += 1.f - lx;
      if (x1 == x) wx += lx;
      if (wx == 0.f) continue;
      sum = fmaf(wy * wx, __ldg(gdst + (((size_t)b * H + Y) * W + X) * C + c), sum);
    }
  }
  gsrc[i] = sum;
}
int launch_upsample_bwd(const float* gdst, float* gsrc, int B, int h, int w, int C, int f, cudaStream_t st) {
  const int64_t total = (int64_t)B * h * w * C;
  if (total == 0) return TMG_OK;
  const int H = h * f, W = w * f;
  const float sy = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.f, sx = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.f;
  upsample_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(gdst, gsrc, B, h, w, C, f, sy, sx, total);
  TMG_LAUNCH_CHECK();
  return TMG_OK;
}

__global__ void dgrad_s2_kernel(S2DgradArgs a, int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % a.cin); int64_t t = i / a.cin; const int x = (int)(t % a.Win); t /= a.Win; const int y = (int)(t % a.Hin); const int b = (int)(t / a.Hin);
  const size_t pix = ((size_t)b * a.Hin + y) * a.Win + x;
  float* dst = a.gx + pix * a.gx_cstride + a.gx_coff + c;
  float sum = 0.f;
  if (!a.mask || __ldg(a.mask + pix * a.gx_cstride + a.gx_coff + c) > 0.f) {
    for (int tap = 0; tap < 9; ++tap) {
      const int yy = y - (tap / 3 - 1), xx = x - (tap % 3 - 1);       // = 2 * p
      if ((yy & 1) || (xx & 1) || yy < 0 || xx < 0) continue;
      const int py = yy >> 1, px = xx >> 1;
      if (py >= a.Hout || px >= a.Wout) continue;
      const float* gp = a.g + (((size_t)b * a.Hout + py) * a.Wout + px) * a.g_cstride + a.g_coff;
      for (int o = 0; o < a.cout; ++o) sum = fmaf(__ldg(a.w_oihw + ((size_t)o * a.cin + c) * 9 + tap), __ldg(gp + o), sum);
    }
  }
  *dst = a.accum ? *dst + sum : sum;
}
int launch_dgrad_s2(const S2DgradArgs& a, cudaStream_t st) {
  const int64_t total = (int64_t)a.B * a.Hin * a.Win * a.cin;
  if (total == 0) return TMG_OK;
  dgrad_s2_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(a, total);
  TMG_LAUNCH_CHECK();
  return TMG_OK;
}

// per-channel sums S1 = sum g_m, S2 = sum g_m * xhat with g_m = ga gated by relu(bn(x)) > 0; one CTA per channel
__global__ void __launch_bounds__(256)
bn_bwd_sums_kernel(BnBwdArgs a) {
  __shared__ double s_red[2][8];
  const int c = blockIdx.x;
  const float mean = a.mean[c], rstd = rsqrtf(a.var[c] + a.eps), gm = a.gamma[c], bt = a.beta[c];
  double s1 = 0.0, s2 = 0.0;
  for (int64_t i = threadIdx.x; i < a.N; i += 256) {
    const float xh = (a.x[i * a.x_cstride + c] - mean) * rstd;
    const float g = (fmaf(gm, xh, bt) > 0.f) ? a.ga[i * a.ga_cstride + c] : 0.f;
    s1 += g; s2 += (double)g * xh;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
  if ((threadIdx.x & 31) == 0) { s_red[0][threadIdx.x >> 5] = s1; s_red[1][threadIdx.x >> 5] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t1 = 0.0, t2 = 0.0;
    for (int k = 0; k < 8; ++k) { t1 += s_red[0][k]; t2 += s_red[1][k]; }
    a.sums[c] = (float)t1; a.sums[a.n + c] = (float)t2;
    a.g_beta[c] += (float)t1; a.g_gamma[c] += (float)t2;
  }
}
__global__ void bn_bwd_apply_kernel(BnBwdArgs a, int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % a.n); const int64_t px = i / a.n;
  const float mean = a.mean[c], rstd = rsqrtf(a.var[c] + a.eps), gm = a.gamma[c], bt = a.beta[c];
  const float xh = (a.x[px * a.x_cstride + c] - mean) * rstd;
  float g = (fmaf(gm, xh, bt) > 0.f) ? a.ga[px * a.ga_cstride + c] : 0.f;
  if (a.train) g = g - a.sums[c] / (float)a.N - xh * (a.sums[a.n + c] / (float)a.N);
  a.gx[px * a.gx_cstride + c] += gm * rstd * g;
}
int launch_bn_relu_bwd(const BnBwdArgs& a, cudaStream_t st) {
  if (a.n <= 0) return TMG_OK;
  bn_bwd_sums_kernel<<<a.n, 256, 0, st>>>(a);
  TMG_LAUNCH_CHECK();
  const int64_t total = a.N * a.n;
  bn_bwd_apply_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(a, total);
  TMG_LAUNCH_CHECK();
  return TMG_OK;
}

}  // namespace tmg
