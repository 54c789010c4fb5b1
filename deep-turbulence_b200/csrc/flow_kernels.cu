// Fused streaming kernels of the flow stack (everything that is not a 3x3 convolution).
// Reference operators restated here (file:line relative to /root/reference/tmglow):
//   coupling scale/shift        nn/modules/flowAffine.py:76-81,102-107
//   InvertibleConv1x1LU apply   nn/modules/glowConv.py:193-194,219-220
//   ActNorm                     nn/modules/actNorm.py:66-67,82-83
//   ConvLSTM gates              nn/modules/convLSTM.py:76-83
//   GaussianDiag                nn/modules/flowUtils.py:163-209
//   CheckerSqueeze              nn/modules/flowUtils.py:99-145
//   UpsamplingLinear            nn/modules/misc.py:34
//   BatchNorm2d (train)         nn/modules/denseBlock.py:49
#include "common.cuh"

namespace tmg {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum over the CTA (blockDim.x multiple of 32, <= 1024); result valid in thread 0.
__device__ __forceinline__ float block_sum(float v, float* s_red) {
  v = warp_sum(v);
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) s_red[wid] = v;
  __syncthreads();
  float r = 0.f;
  if (wid == 0) {
    int nw = (blockDim.x + 31) >> 5;
    r = lane < nw ? s_red[lane] : 0.f;
    r = warp_sum(r);
  }
  return r;
}

// ------------------------------------------------------------------ one flow step, pointwise part
// One thread owns one pixel (all C channels in registers): coupling update of the second half,
// CxC channel mix and ActNorm in one read + one write of the flow state; the coupling log-det
// (sum of the log-scales) is reduced with warp shuffles into one partial per CTA.
template <int C>
__global__ void __launch_bounds__(kPixTile)
flow_pointwise_kernel(PointArgs a) {
  __shared__ __align__(16) float s_w[C * C];
  __shared__ float s_nw[C], s_nb[C];
  __shared__ float s_red[kPixTile / 32];
  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int p = blockIdx.x * kPixTile + tid;
  if (a.wmat) for (int i = tid; i < C * C; i += kPixTile) s_w[i] = __ldg(a.wmat + i);
  if (a.nw) for (int i = tid; i < C; i += kPixTile) { s_nw[i] = __ldg(a.nw + i); s_nb[i] = __ldg(a.nb + i); }
  __syncthreads();

  float ldsum = 0.f;
  if (p < a.HW) {
    float v[C];
    float* yp = a.y + ((size_t)b * a.HW + p) * C;
    const float4* y4 = reinterpret_cast<const float4*>(yp);
#pragma unroll
    for (int i = 0; i < C / 4; ++i) {
      float4 t = y4[i];
      v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
    if (a.hr) {
      const float4* h4 = reinterpret_cast<const float4*>(a.hr + ((size_t)b * a.HW + p) * C);
#pragma unroll
      for (int i = 0; i < C / 4; ++i) {
        float4 t = __ldg(h4 + i);                       // (shift, raw, shift, raw): channels 0::2 / 1::2
        float a0 = 2.f * (t.y / (1.f + fabsf(t.y)));    // 2*softsign
        float a1 = 2.f * (t.w / (1.f + fabsf(t.w)));
        ldsum += a0 + a1;
        float s0 = expf(a0), s1 = expf(a1);
        int j = C / 2 + 2 * i;
        if (a.reverse) { v[j] = v[j] / s0 - t.x; v[j + 1] = v[j + 1] / s1 - t.z; }
        else           { v[j] = (v[j] + t.x) * s0; v[j + 1] = (v[j + 1] + t.z) * s1; }
      }
    }
    if (!a.reverse && a.nw) {
#pragma unroll
      for (int i = 0; i < C; ++i) v[i] = fmaf(s_nw[i], v[i], s_nb[i]);
    }
    if (a.wmat) {
      float o[C];
#pragma unroll
      for (int r = 0; r < C; ++r) {
        float s = 0.f;
        const float4* w4 = reinterpret_cast<const float4*>(s_w + r * C);
#pragma unroll
        for (int i = 0; i < C / 4; ++i) {
          float4 w = w4[i];
          s = fmaf(w.x, v[4 * i], s); s = fmaf(w.y, v[4 * i + 1], s);
          s = fmaf(w.z, v[4 * i + 2], s); s = fmaf(w.w, v[4 * i + 3], s);
        }
        o[r] = s;
      }
#pragma unroll
      for (int i = 0; i < C; ++i) v[i] = o[i];
    }
    if (a.reverse && a.nw) {
#pragma unroll
      for (int i = 0; i < C; ++i) v[i] = (v[i] - s_nb[i]) / s_nw[i];
    }
    float4* o4 = reinterpret_cast<float4*>(yp);
#pragma unroll
    for (int i = 0; i < C / 4; ++i) o4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  }
  if (a.hr && a.ld_part) {
    float r = block_sum(ldsum, s_red);
    if (tid == 0) a.ld_part[(size_t)b * a.ld_stride + blockIdx.x] = r;
  }
}

int launch_flow_pointwise(const PointArgs& a, cudaStream_t st) {
  dim3 grid(cdiv(a.HW, kPixTile), a.B);
  switch (a.C) {
#define TMG_CASE(CC) case CC: flow_pointwise_kernel<CC><<<grid, kPixTile, 0, st>>>(a); break;
    TMG_CASE(4) TMG_CASE(8) TMG_CASE(12) TMG_CASE(16) TMG_CASE(24) TMG_CASE(32) TMG_CASE(48) TMG_CASE(64)
#undef TMG_CASE
    default:
      set_error("flow step: %d channels not supported (4,8,12,16,24,32,48,64)", a.C);
      return TMG_ERR_UNSUPPORTED;
  }
  TMG_LAUNCH_CHECK();
  return TMG_OK;
}

// ------------------------------------------------------------------ ConvLSTM cell update
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void lstm_pointwise_kernel(LstmArgs a) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  int64_t pix = i / a.R;
  int r = (int)(i - pix * a.R);
  const float* g = a.gates + pix * 4 * a.R;
  float gi = sigmoidf_(g[r]);
  float gf = sigmoidf_(g[a.R + r]);
  float go = sigmoidf_(g[2 * a.R + r]);
  float gg = tanhf(g[3 * a.R + r]);
  float c = a.c_prev ? a.c_prev[i] : 0.f;
  float cn = gf * c + gi * gg;
  a.c_out[i] = cn;
  a.h_out[i] = go * tanhf(cn);
}

int launch_lstm_pointwise(const LstmArgs& a, cudaStream_t st) {
  int thr = 256;
  lstm_pointwise_kernel<<<(unsigned)((a.n + thr - 1) / thr), thr, 0, st>>>(a);
  TMG_LAUNCH_CHECK();
  return TMG_OK;
}

// ------------------------------------------------------------------ diagonal Gaussian (split / top prior)
__global__ void __launch_bounds__(kPixTile)
gaussian_kernel(GaussArgs a) {
  __shared__ float s_red[kPixTile / 32];
  const int b = blockIdx.y;
  const int p = blockIdx.x * kPixTile + threadIdx.x;
  float lsum = 0.f;
  if (p < a.HW) {
    const float* pr = a.prm + ((size_t)(a.prm_bshared ? 0 : b) * a.HW + p) * a.prm_cstride;
    float* vp = a.val + ((size_t)b * a.HW + p) * a.val_cstride + a.val_coff;
    // one exp per element (exp(ls) and its reciprocal serve the sample, the noise and the density); channel pairs as 8-byte
    // accesses when every offset is even (the flow's n = C / 2 always is)
    auto one = [&](int j, float mu, float ls, float zin, float& zout) {
      ls = fminf(fmaxf(ls, -10.f), kLog5);                    // GaussianDiag.__init__ clamp
      const size_t ei = ((size_t)b * a.n + j) * a.HW + p;
      const float e = expf(ls), ie = 1.f / e;
      float z;
      if (a.reverse) {
        z = fmaf(e, a.eps_in[ei], mu);
      } else {
        z = zin;
        if (a.eps_out) a.eps_out[ei] = (z - mu) * ie;
      }
      zout = z;
      if (a.val_nchw) a.val_nchw[ei] = z;
      const float d = (z - mu) * ie;
      lsum += -0.5f * (kLog2Pi + ls * 2.f + d * d);
    };
    const bool pair = ((a.n | a.prm_cstride | a.val_cstride | a.val_coff) & 1) == 0 &&
                      (reinterpret_cast<uintptr_t>(a.prm) & 7) == 0 && (reinterpret_cast<uintptr_t>(a.val) & 7) == 0;
    if (pair) {
      for (int j = 0; j < a.n; j += 2) {
        const float2 mu = __ldg(reinterpret_cast<const float2*>(pr + j)), ls = __ldg(reinterpret_cast<const float2*>(pr + a.n + j));
        float2 zin = make_float2(0.f, 0.f), z;
        if (!a.reverse) zin = *reinterpret_cast<const float2*>(vp + j);
        one(j, mu.x, ls.x, zin.x, z.x);
        one(j + 1, mu.y, ls.y, zin.y, z.y);
        if (a.reverse) *reinterpret_cast<float2*>(vp + j) = z;
      }
    } else {
      for (int j = 0; j < a.n; ++j) {
        float z;
        one(j, pr[j], pr[a.n + j], a.reverse ? 0.f : vp[j], z);
        if (a.reverse) vp[j] = z;
      }
    }
  }
  float r = block_sum(lsum, s_red);
  if (threadIdx.x == 0 && a.ld_part) a.ld_part[(size_t)b * a.ld_stride + blockIdx.x] = r;
}

int launch_gaussian(const GaussArgs& a, cudaStream_t st) {
  dim3 grid(cdiv(a.HW, kPixTile), a.B);
  gaussian_kernel<<<grid, kPixTile, 0, st>>>(a);
  TMG_LAUNCH_CHECK();
  return TMG_OK;
}

// ------------------------------------------------------------------ layout / checkerboard permutations
__device__ __forceinline__ void checker(int k, int& dr, int& dc) {
  // (0,0),(1,0),(1,1),(0,1)   flowUtils.py:117-120
  dr = (k == 1 || k == 2);
  dc = (k >= 2);
}

__global__ void permute_kernel(PermArgs a, int64_t total) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int C = a.C, H = a.H, W = a.W;
  switch (a.mode) {
    case PERM_NCHW_TO_NHWC: {      // i over dst [B,H,W,C]
      int c = i % C; int64_t t = i / C; int x = t % W; t /= W; int y = t % H; int b = t / H;
      a.dst[((int64_t)(b * H + y) * W + x) * a.dst_cstride + a.dst_coff + c] =
          a.src[((int64_t)(b * C + c) * H + y) * W + x];
    } break;
    case PERM_NHWC_TO_NCHW: {      // i over dst [B,C,H,W]
      int x = i % W; int64_t t = i / W; int y = t % H; t /= H; int c = t % C; int b = t / C;
      a.dst[i] = a.src[((int64_t)(b * H + y) * W + x) * a.src_cstride + a.src_coff + c];
    } break;
    case PERM_SQUEEZE_NCHW_TO_NHWC:
    case PERM_SQUEEZE_NHWC_TO_NHWC: {   // C,H,W = un-squeezed dims; i over dst [B,H/2,W/2,4C]
      int H2 = H / 2, W2 = W / 2, C4 = 4 * C;
      int kc = i % C4; int64_t t = i / C4; int x = t % W2; t /= W2; int y = t % H2; int b = t / H2;
      int k = kc / C, c = kc - k * C, dr, dc;
      checker(k, dr, dc);
      int sy = 2 * y + dr, sx = 2 * x + dc;
      float v = (a.mode == PERM_SQUEEZE_NCHW_TO_NHWC)
                    ? a.src[((int64_t)(b * C + c) * H + sy) * W + sx]
                    : a.src[((int64_t)(b * H + sy) * W + sx) * a.src_cstride + a.src_coff + c];
      a.dst[((int64_t)(b * H2 + y) * W2 + x) * a.dst_cstride + a.dst_coff + kc] = v;
    } break;
    case PERM_UNSQUEEZE_NHWC_TO_NHWC: { // i over dst [B,H,W,C] (un-squeezed); src [B,H/2,W/2,4C]
      int c = i % C; int64_t t = i / C; int x = t % W; t /= W; int y = t % H; int b = t / H;
      int dr = y & 1, dc = x & 1;
      int k = dr ? (dc ? 2 : 1) : (dc ? 3 : 0);
      a.dst[((int64_t)(b * H + y) * W + x) * a.dst_cstride + a.dst_coff + c] =
          a.src[((int64_t)(b * (H / 2) + (y >> 1)) * (W / 2) + (x >> 1)) * a.src_cstride + a.src_coff + k * C + c];
    } break;
    case PERM_UNSQUEEZE_NHWC_TO_NCHW: { // i over dst [B,C,H,W]
      int x = i % W; int64_t t = i / W; int y = t % H; t /= H; int c = t % C; int b = t / C;
      int dr = y & 1, dc = x & 1;
      int k = dr ? (dc ? 2 : 1) : (dc ? 3 : 0);
      a.dst[i] = a.src[((int64_t)(b * (H / 2) + (y >> 1)) * (W / 2) + (x >> 1)) * a.src_cstride + a.src_coff + k * C + c];
    } break;
    case PERM_SQUEEZE_NCHW_TO_NCHW: {   // i over dst [B,4C,H/2,W/2]
      int H2 = H / 2, W2 = W / 2;
      int x = i % W2; int64_t t = i / W2; int y = t % H2; t /= H2; int kc = t % (4 * C); int b = t / (4 * C);
      int k = kc / C, c = kc - k * C, dr, dc;
      checker(k, dr, dc);
      a.dst[i] = a.src[((int64_t)(b * C + c) * H + 2 * y + dr) * W + 2 * x + dc];
    } break;
    case PERM_UNSQUEEZE_NCHW_TO_NCHW: { // i over dst [B,C,H,W]; src [B,4C,H/2,W/2]
      int x = i % W; int64_t t = i / W; int y = t % H; t /= H; int c = t % C; int b = t / C;
      int dr = y & 1, dc = x & 1;
      int k = dr ? (dc ? 2 : 1) : (dc ? 3 : 0);
      a.dst[i] = a.src[((int64_t)(b * 4 * C + k * C + c) * (H / 2) + (y >> 1)) * (W / 2) + (x >> 1)];
    } break;
  }
}

// CheckerSqueeze.reverse into the user-facing NCHW field (flowUtils.py:124-145): one thread per squeezed pixel reads
// its 4C contiguous channels (float4) and writes, per channel and output row, the two adjacent columns as one float2:
// a warp reads 32*4C*4 contiguous bytes and writes 256 contiguous bytes per (channel, row).
template <int C>
__global__ void __launch_bounds__(256)
unsqueeze_nchw_kernel(PermArgs a, int64_t npix) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  const int H2 = a.H / 2, W2 = a.W / 2;
  const int x = (int)(i % W2); int64_t t = i / W2; const int y = (int)(t % H2); const int b = (int)(t / H2);
  float v[4 * C];
  const float4* s4 = reinterpret_cast<const float4*>(a.src + i * a.src_cstride + a.src_coff);
#pragma unroll
  for (int q = 0; q < C; ++q) { const float4 f = __ldg(s4 + q); v[4 * q] = f.x; v[4 * q + 1] = f.y; v[4 * q + 2] = f.z; v[4 * q + 3] = f.w; }
  // channel groups k = 0..3 hold offsets (dr,dc) = (0,0),(1,0),(1,1),(0,1)
#pragma unroll
  for (int c = 0; c < C; ++c) {
    float* d = a.dst + (((int64_t)b * C + c) * a.H + 2 * y) * a.W + 2 * x;
    *reinterpret_cast<float2*>(d) = make_float2(v[c], v[3 * C + c]);                 // row 2y  : (0,0), (0,1)
    *reinterpret_cast<float2*>(d + a.W) = make_float2(v[C + c], v[2 * C + c]);       // row 2y+1: (1,0), (1,1)
  }
}

// CheckerSqueeze.reverse between two flow levels (NHWC -> NHWC, into the first C channels of the wider state): one thread
// per SOURCE pixel reads its 4C contiguous channels and writes the four destination pixels (C contiguous floats each) with
// 8-byte accesses -- the generic kernel above spends its time on 64-bit index arithmetic per element.
__global__ void __launch_bounds__(256)
unsqueeze_nhwc_kernel(PermArgs a, int64_t npix2) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix2) return;
  const int H2 = a.H / 2, W2 = a.W / 2, C = a.C;
  const int x = (int)(i % W2); const int64_t t = i / W2; const int y = (int)(t % H2); const int b = (int)(t / H2);
  const float2* s2 = reinterpret_cast<const float2*>(a.src + i * a.src_cstride + a.src_coff);
  const int h = C >> 1;                                   // float2 per channel group
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int dr = (k == 1 || k == 2), dc = (k >= 2);    // (0,0),(1,0),(1,1),(0,1)   flowUtils.py:117-120
    float2* d2 = reinterpret_cast<float2*>(a.dst + (((int64_t)b * a.H + 2 * y + dr) * a.W + 2 * x + dc) * a.dst_cstride + a.dst_coff);
    for (int q = 0; q < h; ++q) d2[q] = __ldg(s2 + k * h + q);
  }
}

int launch_permute(const PermArgs& a, cudaStream_t st) {
  int64_t total = (int64_t)a.B * a.C * a.H * a.W;
  if (total == 0) return TMG_OK;
  if (a.mode == PERM_UNSQUEEZE_NHWC_TO_NHWC && (a.C & 1) == 0 && (a.src_cstride & 1) == 0 && (a.src_coff & 1) == 0 &&
      (a.dst_cstride & 1) == 0 && (a.dst_coff & 1) == 0 && (a.H & 1) == 0 && (a.W & 1) == 0 &&
      (reinterpret_cast<uintptr_t>(a.src) & 7) == 0 && (reinterpret_cast<uintptr_t>(a.dst) & 7) == 0) {
    const int64_t npix2 = total / (4 * a.C);
    unsqueeze_nhwc_kernel<<<(unsigned)((npix2 + 255) / 256), 256, 0, st>>>(a, npix2);
    TMG_LAUNCH_CHECK();
    return TMG_OK;
  }
  if (a.mode == PERM_UNSQUEEZE_NHWC_TO_NCHW && a.src_cstride % 4 == 0 && a.src_coff % 4 == 0 && a.W % 2 == 0 &&
      (reinterpret_cast<uintptr_t>(a.src) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.dst) & 7) == 0 &&
      (a.C == 3 || a.C == 1 || a.C == 2 || a.C == 4)) {
    const int64_t npix = total / (4 * a.C);
    const unsigned grid = (unsigned)((npix + 255) / 256);
    switch (a.C) {
      case 1: unsqueeze_nchw_kernel<1><<<grid, 256, 0, st>>>(a, npix); break;
      case 2: unsqueeze_nchw_kernel<2><<<grid, 256, 0, st>>>(a, npix); break;
      case 3: unsqueeze_nchw_kernel<3><<<grid, 256, 0, st>>>(a, npix); break;
      default: unsqueeze_nchw_kernel<4><<<grid, 256, 0, st>>>(a, npix); break;
    }
    TMG_LAUNCH_CHECK();
    return TMG_OK;
  }
  int thr = 256;
  permute_kernel<<<(unsigned)((total + thr - 1) / thr), thr, 0, st>>>(a, total);
  TMG_LAUNCH_CHECK();
  return TMG_OK;
}

// ------------------------------------------------------------------ bilinear upsample, align_corners=True
__global__ void upsample_kernel(UpsampleArgs a, int64_t total, float sy, float sx) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int H = a.h * a.f, W = a.w * a.f;
  int c = i % a.C; int64_t t = i / a.C; int X = t % W; t /= W; int Y = t % H; int b = t / H;
  float fy = sy * Y, fx = sx * X;
  int y0 = (int)fy, x0 = (int)fx;
  int y1 = min(y0 + 1, a.h - 1), x1 = min(x0 + 1, a.w - 1);
  float ly = fy - y0, lx = fx - x0;
  const float* s = a.src + (size_t)b * a.h * a.w * a.C + c;
  float v00 = s[((size_t)y0 * a.w + x0) * a.C], v01 = s[((size_t)y0 * a.w + x1) * a.C];
  float v10 = s[((size_t)y1 * a.w + x0) * a.C], v11 = s[((size_t)y1 * a.w + x1) * a.C];
  a.dst[i] = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
}

int launch_upsample(const UpsampleArgs& a, cudaStream_t st) {
  int64_t total = (int64_t)a.B * a.h * a.f * a.w * a.f * a.C;
  if (total == 0) return TMG_OK;
  int H = a.h * a.f, W = a.w * a.f;
  float sy = H > 1 ? (float)(a.h - 1) / (float)(H - 1) : 0.f;
  float sx = W > 1 ? (float)(a.w - 1) / (float)(W - 1) : 0.f;
  int thr = 256;
  upsample_kernel<<<(unsigned)((total + thr - 1) / thr), thr, 0, st>>>(a, total, sy, sx);
  TMG_LAUNCH_CHECK();
  return TMG_OK;
}

// ------------------------------------------------------------------ BatchNorm batch statistics (train mode)
__global__ void __launch_bounds__(256)
bn_stats_kernel(BnStatArgs a) {
  __shared__ double s_red[8];
  __shared__ float s_mean;
  const int c = a.c0 + blockIdx.x;
  const float* x = a.x + c;
  auto block_sum_d = [&](double v) -> double {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x == 0) for (int i = 0; i < 8; ++i) r += s_red[i];
    __syncthreads();
    return r;
  };
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < a.N; i += 256) s += x[i * a.cstride];
  double tot = block_sum_d(s);
  if (threadIdx.x == 0) s_mean = (float)(tot / (double)a.N);
  __syncthreads();
  float m = s_mean;
  double q = 0.0;
  for (int64_t i = threadIdx.x; i < a.N; i += 256) { float d = x[i * a.cstride] - m; q += (double)d * d; }
  double qt = block_sum_d(q);
  if (threadIdx.x == 0) { a.mean[c] = m; a.var[c] = (float)(qt / (double)a.N); }
}

int launch_bn_stats(const BnStatArgs& a, cudaStream_t st) {
  if (a.n <= 0) return TMG_OK;
  bn_stats_kernel<<<a.n, 256, 0, st>>>(a);
  TMG_LAUNCH_CHECK();
  return TMG_OK;
}

__global__ void bn_fold_train_kernel(BnFoldArgs a) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.n) return;
  float m = a.mean[c], v = a.var[c];
  float sc = a.w[c] * rsqrtf(v + a.eps);
  a.scale[c] = sc;
  a.shift[c] = a.b[c] - m * sc;
  float unb = a.N > 1 ? v * ((float)a.N / (float)(a.N - 1)) : v;
  a.run_mean[c] = (1.f - a.momentum) * a.run_mean[c] + a.momentum * m;
  a.run_var[c] = (1.f - a.momentum) * a.run_var[c] + a.momentum * unb;
}

int launch_bn_fold_train(const BnFoldArgs& a, cudaStream_t st) {
  bn_fold_train_kernel<<<cdiv(a.n, 64), 64, 0, st>>>(a);
  TMG_LAUNCH_CHECK();
  return TMG_OK;
}

// ------------------------------------------------------------------ per-sample log-det assembly
__global__ void logdet_reduce_kernel(LogdetArgs a) {
  // one warp per sample; every lane sums a fixed strided subset in fp64, then a fixed butterfly: the order of
  // additions depends only on (ld_stride), so results are bit-reproducible run to run
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= a.B) return;
  double s = 0.0;
  const float* p = a.ld_part + (size_t)b * a.ld_stride;
  for (int i = lane; i < a.ld_stride; i += 32) s += (double)p[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) {
    for (int l = 0; l < a.n_levels; ++l) {
      double cs = 0.0;
      for (int k = a.step_begin[l]; k < a.step_begin[l + 1]; ++k) cs += a.step_const[k];
      s += cs * (double)a.hw[l];
    }
    a.out[b] = (float)s;
  }
}

int launch_logdet_reduce(const LogdetArgs& a, cudaStream_t st) {
  logdet_reduce_kernel<<<cdiv(a.B * 32, 128), 128, 0, st>>>(a);
  TMG_LAUNCH_CHECK();
  return TMG_OK;
}

}  // namespace tmg
