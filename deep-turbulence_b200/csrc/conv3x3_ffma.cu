// Generic fp32 3x3 convolution on CUDA cores (exact-fp32 path of every conv on the hot path:
// Encoder convs tmGlow.py:88,94,148,154,180; dense layers denseBlock.py:51,136; ConvLSTM gate and
// output convs convLSTM.py:44,129; Conv2dZeros flowUtils.py:229).
//
// Implicit GEMM, M = 128*PX consecutive output pixels of one sample, N = NB output channels,
// K = 9 taps x Cin walked in chunks of 16 channels.  The input halo tile is staged channel-planar
// in shared memory (ReLU / folded BatchNorm / padding applied while staging, so the reference's
// torch.cat + relu + pad tensors never exist in HBM); weights are staged tap-major and read as
// warp-uniform float4 broadcasts.
#include "common.cuh"

namespace tmg {

constexpr int kChunk = 16;     // input channels per K chunk
constexpr int kThreads = 128;

template <int NB, int PX>
__global__ void __launch_bounds__(kThreads)
conv3x3_kernel(ConvArgs a, int plane, int rows_in_max) {
  extern __shared__ __align__(16) float smem[];
  float* s_in = smem;                       // [kChunk][plane]
  float* s_w = smem + kChunk * plane;       // [9][kChunk][NB]

  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int n0 = blockIdx.z * NB;
  const int T = kThreads * PX;
  const int HWo = a.Hout * a.Wout;
  const int p0 = blockIdx.x * T;
  const int p_last = min(p0 + T, HWo) - 1;
  const int r0 = p0 / a.Wout;
  const int r1 = p_last / a.Wout;
  const int tw = a.Win + 2;                          // tile width  (input cols -1 .. Win)
  const int th = (r1 - r0) * a.stride + 3;           // tile height (input rows r0*s-1 .. r1*s+1)
  const int row_in0 = r0 * a.stride - 1;
  const int tile_px = th * tw;

  int base[PX];
  bool valid[PX];
#pragma unroll
  for (int j = 0; j < PX; ++j) {
    int p = p0 + tid + j * kThreads;
    valid[j] = p < HWo;
    int pp = valid[j] ? p : p0;
    int ro = pp / a.Wout, co = pp - ro * a.Wout;
    base[j] = (ro - r0) * a.stride * tw + co * a.stride;
  }

  float acc[PX][NB];
#pragma unroll
  for (int j = 0; j < PX; ++j)
#pragma unroll
    for (int n = 0; n < NB; ++n) acc[j][n] = 0.f;

  const int cin = a.cin_w;
  for (int c0 = 0; c0 < cin; c0 += kChunk) {
    __syncthreads();
    // ---- stage the input halo tile for channels [c0, c0+16)
    for (int idx = tid; idx < tile_px * kChunk; idx += kThreads) {
      int ch = idx & (kChunk - 1);
      int pix = idx >> 4;
      int ri = pix / tw, ci = pix - ri * tw;
      int gr = row_in0 + ri, gc = ci - 1;
      int c = c0 + ch;
      float v = 0.f;
      bool inb = (gr >= 0) && (gr < a.Hin) && (gc >= 0) && (gc < a.Win);
      if (a.pad_replicate) {
        gr = min(max(gr, 0), a.Hin - 1);
        gc = min(max(gc, 0), a.Win - 1);
        inb = true;
      }
      if (inb && c < cin) {
        int cc = c;
        const ConvSrc* s = &a.src[0];
        if (cc >= s->nch && a.nsrc > 1) { cc -= s->nch; s = &a.src[1];
          if (cc >= s->nch && a.nsrc > 2) { cc -= s->nch; s = &a.src[2]; } }
        if (cc < s->nch) {
          v = __ldg(s->p + ((size_t)((s->bshared ? 0 : b) * a.Hin + gr) * a.Win + gc) * s->cstride + s->coff + cc);
          if (a.bn_scale) v = fmaf(v, __ldg(a.bn_scale + c), __ldg(a.bn_shift + c));
          if (s->relu) v = fmaxf(v, 0.f);
        }
      }
      s_in[ch * plane + pix] = v;
    }
    // ---- stage the weights of this chunk: s_w[tap][ch][n]
    for (int idx = tid; idx < 9 * kChunk * NB; idx += kThreads) {
      int n = idx % NB;
      int rest = idx / NB;
      int ch = rest & (kChunk - 1);
      int tap = rest >> 4;
      int c = c0 + ch, o = n0 + n;
      float v = 0.f;
      if (c < cin && o < a.cout_w) v = __ldg(a.w + ((size_t)tap * cin + c) * a.cout_w + o);
      s_w[idx] = v;
    }
    __syncthreads();

    const int chmax = min(kChunk, cin - c0);
    for (int ch = 0; ch < chmax; ++ch) {
      const float* sp = s_in + ch * plane;
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int toff = (tap / 3) * tw + (tap % 3);
        float av[PX];
#pragma unroll
        for (int j = 0; j < PX; ++j) av[j] = sp[base[j] + toff];
        const float4* wp = reinterpret_cast<const float4*>(s_w + (tap * kChunk + ch) * NB);
#pragma unroll
        for (int n4 = 0; n4 < NB / 4; ++n4) {
          float4 w4 = wp[n4];
#pragma unroll
          for (int j = 0; j < PX; ++j) {
            acc[j][n4 * 4 + 0] = fmaf(av[j], w4.x, acc[j][n4 * 4 + 0]);
            acc[j][n4 * 4 + 1] = fmaf(av[j], w4.y, acc[j][n4 * 4 + 1]);
            acc[j][n4 * 4 + 2] = fmaf(av[j], w4.z, acc[j][n4 * 4 + 2]);
            acc[j][n4 * 4 + 3] = fmaf(av[j], w4.w, acc[j][n4 * 4 + 3]);
          }
        }
      }
    }
  }

  // ---- epilogue: bias, gain, activation, store into the (possibly wider) NHWC destination
  const float gain = a.gain ? __ldg(a.gain) : 1.f;
#pragma unroll
  for (int j = 0; j < PX; ++j) {
    if (!valid[j]) continue;
    int p = p0 + tid + j * kThreads;
    float* op = a.out + ((size_t)b * HWo + p) * a.out_cstride + a.out_coff + n0;
    const float* mp = a.mask ? a.mask + ((size_t)b * HWo + p) * a.out_cstride + a.out_coff + n0 : nullptr;
#pragma unroll
    for (int n = 0; n < NB; ++n) {
      if (n0 + n < a.cout) {
        float v = acc[j][n];
        if (a.bias) v += __ldg(a.bias + n0 + n);
        if (a.gain) v *= gain;
        if (a.act == 1) v = fmaxf(v, 0.f);
        else if (a.act == 2) v = fminf(fmaxf(v, -2.f), kLog5);
        if (mp && !(__ldg(mp + n) > 0.f)) v = 0.f;
        op[n] = a.accum ? op[n] + v : v;
      }
    }
  }
}

template <int NB, int PX>
static int launch_t(const ConvArgs& a, cudaStream_t st) {
  const int T = kThreads * PX;
  const int HWo = a.Hout * a.Wout;
  int rows_out = min(a.Hout, (T + a.Wout - 2) / a.Wout + 1);
  int rows_in = (rows_out - 1) * a.stride + 3;
  int tile_px = rows_in * (a.Win + 2);
  int plane = tile_px;
  while ((plane & 31) != 2) ++plane;         // plane stride == 2 (mod 32): conflict-free staging
  size_t smem = (size_t)(kChunk * plane + 9 * kChunk * NB) * sizeof(float);
  if (smem > 200 * 1024) {
    set_error("conv3x3: input width %d needs %zu B of shared memory", a.Win, smem);
    return TMG_ERR_UNSUPPORTED;
  }
  static bool attr_set = false;      // per instantiation
  if (!attr_set) {
    TMG_SMEM_ATTR(conv3x3_kernel<NB, PX>, 200 * 1024);
    attr_set = true;
  }
  dim3 grid(cdiv(HWo, T), a.B, cdiv(a.cout, NB));
  conv3x3_kernel<NB, PX><<<grid, kThreads, smem, st>>>(a, plane, rows_in);
  TMG_LAUNCH_CHECK();
  return TMG_OK;
}

int launch_conv3x3(const ConvArgs& a, cudaStream_t st) {
  if (a.B <= 0 || a.Hout <= 0 || a.Wout <= 0) return TMG_OK;
  if (a.cout_w % 4 != 0) { set_error("conv3x3: packed cout %d not a multiple of 4", a.cout_w); return TMG_ERR_BAD_SHAPE; }
  // small problems (the encoder and the hoisted conditioning tables at batch 1): a grid of a few CTAs is latency
  // bound, so split the output channels four at a time over blockIdx.z and use 128-pixel tiles
  if ((long long)cdiv(a.Hout * a.Wout, 256) * a.B * cdiv(a.cout, 16) < 256) return launch_t<4, 1>(a, st);
  if (a.cout <= 4) return launch_t<4, 2>(a, st);
  if (a.cout <= 16) return launch_t<16, 2>(a, st);
  if (a.cout <= 32) return launch_t<32, 2>(a, st);
  return launch_t<64, 1>(a, st);
}

}  // namespace tmg
