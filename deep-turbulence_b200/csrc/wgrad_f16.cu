// Weight gradient of a 3x3 convolution (stride 1) on the 5th-generation tensor cores (tcgen05, sm_100a): the adjoint
// every nn.Conv2d of the flow steps needs when the reference trains through TMGlow.sample()
// (nn/trainFlowParallel.py:259-277, autograd of F.conv2d in convLSTM.py:44,129, denseBlock.py:136, flowUtils.py:229).
//
//   dW[o][c][tap] = sum over pixels p of  g[p][o] * x[p + off(tap)][c]
//
// is one GEMM per tap with the PIXELS as the contraction dimension: D_tap[M = o][N = c] += A[K = p][M] * B_tap[K = p][N].
// Both operands live in HBM as NHWC, i.e. with the GEMM's M / N index contiguous and K strided: the MN-major operand
// form of tcgen05 (instruction-descriptor bits 15/16).  Staged as 8-channel planes with 16 B per pixel -- the layout the
// forward kernel (conv3x3_f16.cu) already uses -- a core matrix is 8 pixels x 8 channels, a filter tap is a start-address
// shift of the x tile (halo staged once, zero or replicate padding applied while staging), and one MMA contracts one row
// of 16 pixels.  fp16 hi/lo operand split (hi*hi + lo*hi + hi*lo, fp32 accumulate in TMEM): fp32-grade.  The output
// gradient is far below the fp16 normal range, so it is scaled by a power of two (launch_absmax_scale) while staging.
//
// Work split: the accumulators of all 9 taps do not fit the 512 TMEM columns, so a CTA owns one (M tile of 128 output
// channels, group of taps) pair and a share of the pixel tiles (4 rows x 16 columns of one sample); partial sums per
// share are reduced in a fixed order by wgrad_f16_reduce_kernel (deterministic, no atomics).
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace tmg {

constexpr int kWfRows = 4;                         // pixel tile: 4 rows x 16 columns
constexpr int kWfPosA = kWfRows * 16;              // 64 positions of the g tile
constexpr int kWfRP = 18;                          // x tile pitch (16 + 2)
constexpr int kWfPosB = (kWfRows + 2) * kWfRP;     // 108 positions
constexpr int kWfPosBA = 112;
constexpr uint32_t kWfPLA = kWfPosA * 16;          // bytes of one 8-channel plane of the g tile
constexpr uint32_t kWfPLB = kWfPosBA * 16;
constexpr int kWfProd = 8;                         // producer warps (1..8); warp 0 issues the MMAs
constexpr int kWfThreads = (1 + kWfProd) * 32;
constexpr int kWfNPT = kWfProd * 32;
constexpr int kWfBatch = 5;                        // staged items (8 channels of one position) per thread and load batch

struct WgF16Geom {
  int NPl;               // x planes (even): MMA N = 8 * NPl
  int nplB;              // staged (real) x planes
  int plane0[3], nplanes[3];
  int mtiles, ntg, tps;  // M tiles of 128 output channels, tap groups, taps per group
  int nsplit;            // pixel shares per (M tile, tap group)
  int tiles_x, tiles_y, ntiles;
  int Mrows;             // rows of a partial block (cout rounded up to 32)
  uint32_t hlA, hlB, stageBytes, oB, total;
};

__device__ __forceinline__ uint32_t wf_idesc(int n) {     // D = F32, A = B = F16, both MN-major, M = 128
  return (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void wf_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

__device__ __forceinline__ void wf_store_hl(uint8_t* dst, uint32_t hl, const float* v, float scale, bool relu) {
  uint32_t ph[4], pl[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float y0 = fminf(v[2 * e] * scale, 60000.f), y1 = fminf(v[2 * e + 1] * scale, 60000.f);
    y0 = fmaxf(y0, relu ? 0.f : -60000.f); y1 = fmaxf(y1, relu ? 0.f : -60000.f);
    const __half2 h2 = __floats2half2_rn(y0, y1);
    const float2 hf = __half22float2(h2);
    const __half2 l2 = __floats2half2_rn(y0 - hf.x, y1 - hf.y);
    ph[e] = *reinterpret_cast<const uint32_t*>(&h2);
    pl[e] = *reinterpret_cast<const uint32_t*>(&l2);
  }
  *reinterpret_cast<uint4*>(dst) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
  *reinterpret_cast<uint4*>(dst + hl) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
}

__device__ __forceinline__ void wf_load8(const float* ptr, int nv, float* v) {
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] = 0.f;
  if (nv == 8 && (reinterpret_cast<uintptr_t>(ptr) & 15) == 0) {
    const float4 t0 = __ldg(reinterpret_cast<const float4*>(ptr)), t1 = __ldg(reinterpret_cast<const float4*>(ptr) + 1);
    v[0] = t0.x; v[1] = t0.y; v[2] = t0.z; v[3] = t0.w; v[4] = t1.x; v[5] = t1.y; v[6] = t1.z; v[7] = t1.w;
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) if (e < nv) v[e] = __ldg(ptr + e);
  }
}

__global__ void __launch_bounds__(kWfThreads, 1)
wgrad_f16_kernel(WgradArgs a, WgF16Geom g, const float* __restrict__ gscale, float* __restrict__ part) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bars[5];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint64_t* full = bars;           // [2] producers (kWfNPT)
  uint64_t* freeb = bars + 2;      // [2] one commit
  uint64_t* acc_full = bars + 4;   // one commit

  const int combo = blockIdx.x / g.nsplit, split = blockIdx.x - combo * g.nsplit;
  const int mt = combo / g.ntg, tg = combo - mt * g.ntg;
  const int N = 8 * g.NPl;
  const int HW = a.H * a.W;
  const int nmy = split < g.ntiles ? (g.ntiles - 1 - split) / g.nsplit + 1 : 0;
  const int o0 = mt * 128;
  const int nplA = min(16, (a.cout - o0 + 7) / 8);      // real planes of the g tile

  if (tid == 0) {
    mbar_init(full, kWfNPT); mbar_init(full + 1, kWfNPT);
    mbar_init(freeb, 1); mbar_init(freeb + 1, 1);
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  {   // zero both stages once: planes / positions that are never staged stay zero
    uint4* z4 = reinterpret_cast<uint4*>(smem);
    const int n4 = (int)(2 * g.stageBytes / 16);
    for (int i = tid; i < n4; i += kWfThreads) z4[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    // =========================================================== MMA issue
    if (elect_one()) {
      const uint32_t idesc = wf_idesc(N);
      const uint64_t hlA16 = g.hlA >> 4, hlB16 = g.hlB >> 4;
      // MN-major, no swizzle: LBO = stride between groups of 8 pixels (K), SBO = stride between 8-channel planes (M / N)
      const uint32_t lboA = 128u, sboA = kWfPLA, lboB = 128u, sboB = kWfPLB;
      for (int k = 0; k < nmy; ++k) {
        const int sidx = k & 1;
        mbar_wait(full + sidx, (uint32_t)((k >> 1) & 1));
        tc_fence_after();
        const uint8_t* sA = smem + (size_t)sidx * g.stageBytes;
        const uint64_t a0 = make_desc(smem_u32(sA), lboA, sboA);
        const uint64_t b0 = make_desc(smem_u32(sA + g.oB), lboB, sboB);
        for (int rr = 0; rr < kWfRows; ++rr) {
          const uint64_t ad = a0 + (uint64_t)(rr * 16);
          for (int tl = 0; tl < g.tps; ++tl) {
            const int tap = tg * g.tps + tl;
            if (tap > 8) break;
            const int dr = tap / 3, dc = tap - 3 * dr;
            const uint64_t bd = b0 + (uint64_t)((rr + dr) * kWfRP + dc);
            const uint32_t td = tmem_base + (uint32_t)(tl * N);
            const uint32_t acc = (k > 0 || rr > 0) ? 1u : 0u;
            wf_mma(td, ad, bd, idesc, acc);
            wf_mma(td, ad + hlA16, bd, idesc, 1u);
            wf_mma(td, ad, bd + hlB16, idesc, 1u);
          }
        }
        mma_commit(freeb + sidx);
      }
      mma_commit(acc_full);
    }
    __syncwarp();
  } else {
    // =========================================================== producers: g tile (A) and x tile with halo (B)
    const int ptid = tid - 32;
    const float gs = __ldg(gscale);
    const int itemsA = nplA * kWfPosA, itemsB = g.nplB * kWfPosB, items = itemsA + itemsB;
    for (int k = 0; k < nmy; ++k) {
      const int sidx = k & 1, use = k >> 1;
      if (use >= 1) mbar_wait(freeb + sidx, (uint32_t)((use - 1) & 1));
      const int t = split + k * g.nsplit;
      const int tiles_img = g.tiles_x * g.tiles_y;
      const int b = t / tiles_img, ti = t - b * tiles_img;
      const int ty = ti / g.tiles_x, tx = ti - ty * g.tiles_x;
      const int r0 = ty * kWfRows, c0 = tx * 16;
      uint8_t* sA = smem + (size_t)sidx * g.stageBytes;
      uint8_t* sB = sA + g.oB;
      // batches of kWfBatch items per thread, all loads of a batch before its first conversion: the operands of a large batch
      // come from DRAM, and a tile staged in batches of 2 waited five round trips (measured on the forward convolutions:
      // conv3x3_f16.cu, lstm_gate_f16.cu)
      for (int it0 = ptid; it0 < items; it0 += kWfBatch * kWfNPT) {
        float v[kWfBatch][8];
        uint8_t* dst[kWfBatch];
        uint32_t hl[kWfBatch];
        float sc[kWfBatch];
        bool relu[kWfBatch];
#pragma unroll
        for (int q = 0; q < kWfBatch; ++q) {
          const int it = it0 + q * kWfNPT;
          dst[q] = nullptr; hl[q] = 0; sc[q] = 1.f; relu[q] = false;
#pragma unroll
          for (int e = 0; e < 8; ++e) v[q][e] = 0.f;
          if (it >= items) continue;
          if (it < itemsA) {
            const int pl = it / kWfPosA, pos = it - pl * kWfPosA;
            const int r = r0 + (pos >> 4), c = c0 + (pos & 15);
            dst[q] = sA + (size_t)pl * kWfPLA + (size_t)pos * 16; hl[q] = g.hlA; sc[q] = gs;
            const int ch = o0 + pl * 8;
            if (r < a.H && c < a.W)
              wf_load8(a.g + ((size_t)b * HW + (size_t)r * a.W + c) * a.g_cstride + a.g_coff + ch, min(8, a.cout - ch), v[q]);
          } else {
            const int itb = it - itemsA;
            const int pl = itb / kWfPosB, pos = itb - pl * kWfPosB;
            const int rr = pos / kWfRP, rc = pos - rr * kWfRP;
            int r = r0 - 1 + rr, c = c0 - 1 + rc;
            bool inb = r >= 0 && r < a.H && c >= 0 && c < a.W;
            if (a.pad_replicate) { r = min(max(r, 0), a.H - 1); c = min(max(c, 0), a.W - 1); inb = true; }
            dst[q] = sB + (size_t)pl * kWfPLB + (size_t)pos * 16; hl[q] = g.hlB;
            int si = 0;
            if (a.nsrc > 1 && pl >= g.plane0[1]) si = 1;
            if (a.nsrc > 2 && pl >= g.plane0[2]) si = 2;
            const ConvSrc& s = a.src[si];
            const int ch = (pl - g.plane0[si]) * 8;
            const int nv = min(8, s.nch - ch);
            relu[q] = s.relu != 0;
            if (inb && nv > 0 && s.p != nullptr)      // null source = zeros (LSTM step without incoming states)
              wf_load8(s.p + ((s.bshared ? 0 : (size_t)b * HW) + (size_t)r * a.W + c) * s.cstride + s.coff + ch, nv, v[q]);
          }
        }
#pragma unroll
        for (int q = 0; q < kWfBatch; ++q)
          if (dst[q]) wf_store_hl(dst[q], hl[q], v[q], sc[q], relu[q]);
      }
      fence_proxy_async();
      mbar_arrive(full + sidx);
    }
    // =========================================================== epilogue: TMEM -> partial sums of this share
    if (nmy > 0) {
      mbar_wait(acc_full, 0u);
      tc_fence_after();
    }
    const int lg = warp & 3, half = (warp - 1) >> 2;        // TMEM lane group of this warp; column-chunk parity
    const int row = lg * 32 + lane;                          // output channel inside the M tile
    const bool warp_live = o0 + lg * 32 < a.cout;
    if (warp_live) {
      for (int tl = 0; tl < g.tps; ++tl) {
        const int tap = tg * g.tps + tl;
        if (tap > 8) break;
        for (int n0 = half * 16; n0 < N; n0 += 32) {
          float v[16];
          if (nmy > 0) tmem_ld16(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(tl * N + n0), v);
          else {
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = 0.f;
          }
          if (o0 + row < a.cout) {
            float* pp = part + ((size_t)(split * 9 + tap) * N + n0) * g.Mrows + o0 + row;
#pragma unroll
            for (int e = 0; e < 16; ++e) pp[(size_t)e * g.Mrows] = v[e];
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// gw[o][c][tap] (+)= inv_scale * sum over shares; column of channel c follows the plane layout.  A block reduces 32
// outputs: 8 threads per output walk the shares 8 apart, then a fixed-order sum of the 8 partials (deterministic).
__global__ void __launch_bounds__(256)
wgrad_f16_reduce_kernel(const float* __restrict__ part, const float* __restrict__ gscale, WgF16Geom g, int O, int I, int nch0,
                        int nch1, float* __restrict__ gw, int accum) {
  __shared__ float sm[8][32];
  const int x = threadIdx.x & 31, y = threadIdx.x >> 5;
  const int idx = blockIdx.x * 32 + x;
  const bool live = idx < O * I * 9;
  float s = 0.f;
  int o = 0, c = 0, tap = 0;
  if (live) {
    o = idx % O; const int rest = idx / O; c = rest % I; tap = rest / I;
    int n;
    if (c < nch0) n = g.plane0[0] * 8 + c;
    else if (c < nch0 + nch1) n = g.plane0[1] * 8 + (c - nch0);
    else n = g.plane0[2] * 8 + (c - nch0 - nch1);
    const int N = 8 * g.NPl;
    const float* p = part + ((size_t)tap * N + n) * g.Mrows + o;
    const size_t sstride = (size_t)9 * N * g.Mrows;
    for (int k = y; k < g.nsplit; k += 8) s += p[(size_t)k * sstride];
  }
  sm[y][x] = s;
  __syncthreads();
  if (y == 0 && live) {
    float t = sm[0][x];
#pragma unroll
    for (int q = 1; q < 8; ++q) t += sm[q][x];
    t *= __ldg(gscale + 1);
    float* d = gw + ((size_t)o * I + c) * 9 + tap;
    *d = accum ? *d + t : t;
  }
}

// column sums of g (bias gradient): per-block partial sums [blocks][cout]
__global__ void __launch_bounds__(256)
colsum_part_kernel(const float* __restrict__ gsrc, int64_t npix, int cstride, int coff, int cout, int cw, float* __restrict__ part) {
  __shared__ float s[256];
  const int c = threadIdx.x % cw, rowi = threadIdx.x / cw, rows = 256 / cw;
  float acc = 0.f;
  if (c < cout)
    for (int64_t p = (int64_t)blockIdx.x * rows + rowi; p < npix; p += (int64_t)gridDim.x * rows) acc += __ldg(gsrc + p * cstride + coff + c);
  s[threadIdx.x] = acc;
  __syncthreads();
  if (rowi == 0 && c < cout) {
    for (int r = 1; r < rows; ++r) acc += s[r * cw + c];
    part[(size_t)blockIdx.x * cout + c] = acc;
  }
}

static int wf_sm_count() {
  int dev = 0, nsm = 148;
  cudaGetDevice(&dev);
  static int cached[64] = {0};
  if (dev >= 0 && dev < 64) {
    if (!cached[dev]) cudaDeviceGetAttribute(&cached[dev], cudaDevAttrMultiProcessorCount, dev);
    if (cached[dev] > 0) nsm = cached[dev];
  }
  return nsm;
}

static bool wf_geom(const WgradArgs& a, WgF16Geom& g) {
  if (a.stride == 2 || a.bn_scale || a.nsrc < 1 || a.nsrc > 3 || a.cout < 1 || a.B < 1) return false;
  int pl = 0;
  for (int i = 0; i < 3; ++i) {
    g.plane0[i] = pl;
    g.nplanes[i] = i < a.nsrc ? (a.src[i].nch + 7) / 8 : 0;
    pl += g.nplanes[i];
  }
  g.nplB = pl;
  g.NPl = (pl + 1) / 2 * 2;
  const int N = 8 * g.NPl;
  if (N < 16 || N > 256) return false;
  g.mtiles = cdiv(a.cout, 128);
  const int tmax = std::min(9, 512 / N);
  g.ntg = cdiv(9, tmax);
  g.tps = cdiv(9, g.ntg);
  g.tiles_x = cdiv(a.W, 16); g.tiles_y = cdiv(a.H, kWfRows);
  g.ntiles = g.tiles_x * g.tiles_y * a.B;
  const int combos = g.mtiles * g.ntg;
  g.nsplit = std::max(1, std::min(g.ntiles, wf_sm_count() / combos));
  g.Mrows = (a.cout + 31) / 32 * 32;
  g.hlA = 16 * kWfPLA; g.hlB = (uint32_t)g.NPl * kWfPLB;
  g.oB = 2 * g.hlA;
  g.stageBytes = g.oB + 2 * g.hlB;
  g.total = 2 * g.stageBytes;
  return g.total <= 220 * 1024;
}

bool wgrad_f16_supported(const WgradArgs& a) {
  WgF16Geom g{};
  return wf_geom(a, g);
}

size_t wgrad_f16_scratch_floats(int cout, const int* nch, int nsrc, int B, int H, int W) {
  WgradArgs a{};
  a.nsrc = nsrc; a.cout = cout; a.B = B; a.H = H; a.W = W;
  for (int i = 0; i < nsrc && i < 3; ++i) a.src[i].nch = nch[i];
  WgF16Geom g{};
  if (!wf_geom(a, g)) return 0;
  return (size_t)g.nsplit * 9 * 8 * g.NPl * g.Mrows + (size_t)std::max(wf_sm_count(), 444) * cout + 64;
}
// where the bias column-sum partials ([blocks <= 444][cout]) live inside a.scratch
float* wgrad_f16_bias_partials(const WgradArgs& a) {
  WgF16Geom g{};
  if (!wf_geom(a, g)) return nullptr;
  return a.scratch + (size_t)g.nsplit * 9 * 8 * g.NPl * g.Mrows;
}

// a.scratch: wgrad_f16_scratch_floats; gscale: [2] from launch_absmax_scale on a.g
int launch_wgrad_f16(const WgradArgs& a, const float* gscale, cudaStream_t st) {
  WgF16Geom g{};
  if (!wf_geom(a, g)) { set_error("tensor-core weight gradient: unsupported shape"); return TMG_ERR_UNSUPPORTED; }
  float* part = a.scratch;
  TMG_SMEM_ATTR(wgrad_f16_kernel, 220 * 1024);
  const int grid = g.mtiles * g.ntg * g.nsplit;
  wgrad_f16_kernel<<<grid, kWfThreads, g.total, st>>>(a, g, gscale, part);
  TMG_LAUNCH_CHECK();
  const int n = a.cout * a.cin * 9;
  wgrad_f16_reduce_kernel<<<cdiv(n, 32), 256, 0, st>>>(part, gscale, g, a.cout, a.cin, a.src[0].nch, a.nsrc > 1 ? a.src[1].nch : 0,
                                                        a.gw, a.accum);
  TMG_LAUNCH_CHECK();
  if (a.gbias) {
    if (a.cout > 256) { set_error("tensor-core weight gradient: bias of %d channels", a.cout); return TMG_ERR_UNSUPPORTED; }
    float* pb = part + (size_t)g.nsplit * 9 * 8 * g.NPl * g.Mrows;
    int cw = 1;
    while (cw < a.cout) cw <<= 1;
    const int64_t npix = (int64_t)a.B * a.H * a.W;
    const int nb = (int)std::max<int64_t>(1, std::min<int64_t>(wf_sm_count(), npix / 64));
    colsum_part_kernel<<<nb, 256, 0, st>>>(a.g, npix, a.g_cstride, a.g_coff, a.cout, cw, pb);
    TMG_LAUNCH_CHECK();
    TMG_TRY(launch_reduce_cols(pb, nb, a.cout, 0, a.cout, a.gbias, a.accum, st));
  }
  return TMG_OK;
}

}  // namespace tmg
