// Fused flow step, second generation (sm_100a): persistent, software-pipelined, fp16-operand tcgen05.
//
// One launch = one flow step over all samples:
//   t = cat(x1, cond) -> d1 = conv3x3(relu(t)) -> d2 = conv3x3(relu(cat(t,d1))) -> h = Conv2dZeros(relu(cat(t,d1,d2)))
//   (flowAffine.py:49-55,73; denseBlock.py:135-150; flowUtils.py:238-247), then the coupling update, the invertible
//   1x1 convolution, ActNorm and the per-sample log-det partial (flowAffine.py:76-81,102-107; glowConv.py:193,219;
//   actNorm.py:66,82).
//
// What changed against coupling_tc.cu (measured on B200, tools/ubench_mma.cu: a tcgen05.mma with N <= 64 costs
// ~45 cycles whatever N is, and a single issuing thread that does address arithmetic between MMAs runs at
// 100-200 cycles per MMA):
//   * operands are fp16 (kind::f16, K = 16 per MMA).  "x3" mode splits both operands into hi + lo halves
//     (a = a_hi + a_lo, three MMAs: hi*hi + lo*hi + hi*lo); with weights pre-scaled by a power of two the
//     result is fp32-grade (2^-22 relative per operand) at HALF the MMA count and HALF the shared memory of
//     the 3xTF32 split.  Single-pass fp16 (11-bit mantissa like TF32) is the fast mode.
//   * persistent CTAs (one per SM) loop over 16x16-pixel tiles; the step's weights stay in shared memory.
//   * three roles run concurrently on different tiles: producer warps stage tile k+2, the MMA warp issues
//     E(k+1) then Z(k), the epilogue warps gather d1/d2 of tile k+1 then finish tile k.
//   * E = one un-shifted GEMM with the 9 taps of BOTH Cout=1 dense layers in N (cols 0-8: layer 1, 16-24: layer 2);
//     the d1 -> d2 dependency (one input channel, 9 taps) is 9 FMAs per position on CUDA cores.
//   * Z = Conv2dZeros as a tap-shifted implicit GEMM whose M tile is 16 rows x 8 columns of output pixels:
//     the shared-memory descriptor's stride between 8-row groups (SBO) is one padded image row, so no
//     accumulator row is wasted on halo columns.
//   * the A operand lives in a ring of K-step slots (16 channels of one tile each).  A tile's chunks are staged
//     conditioning first, the x1+d chunk last: E and the d-independent part of Z consume (and release) the
//     conditioning chunks while the gather of d1/d2 is still running, so the un-hoisted path (distinct LF inputs,
//     LSTM tail, forward likelihood) pipelines across tiles with 3-4 slots instead of whole-tile double buffers.
//   * K layout [x1 | d1 d2 | cond]: when all samples share one conditioning input (one LF snapshot, many
//     stochastic samples) the cond K-steps are dropped and their contribution, which is the same for every
//     sample, is read from a per-step table computed once per call ("hoisted", model.cu).
#include <cuda_fp16.h>
#include <cstdlib>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace tmg {

constexpr int kRP = 22;                  // padded tile pitch: 16 + 2*3
constexpr int kNPOS = 484;               // 22 x 22 staged positions
constexpr int kNPOSA = 512;              // allocated positions (4 E tiles of 128)
constexpr uint32_t kPLB = kNPOSA * 16;   // bytes of one 8-channel plane
constexpr int kEpiThreads = 256;

struct Step2Geom {
  int KS, KSy, PLtot, kd, nbuf, pipelined;   // nbuf: slots of the K-step ring (one slot = 16 channels of one tile)
  int tiles_x, tiles_y, ntiles, nhl;
  int step_b, step_t; uint32_t inv_tx; int ngroups;
  uint32_t hlA, bufA, scratch;               // hlA: hi->lo distance inside a ring slot, bufA: slot size
  uint32_t oA, oDsc, oWE, oWZ, oWm, oBar, total;
  uint32_t wE_hl, wZ_tap, wZ_hl;         // shared-memory strides of the weight copies
  uint32_t gE_hl, gZ_tap, gZ_hl;         // strides of the packed (global) weights
};

__device__ __forceinline__ uint32_t idesc_f16(int n) {   // D = F32, A = B = F16, K-major, M = 128
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// fp32 -> (hi, lo) fp16 pair; |v| is clamped to the fp16 range (flow states beyond 6e4 are not meaningful)
__device__ __forceinline__ void split_h(float v, __half& hi, __half& lo) {
  v = fminf(fmaxf(v, -60000.f), 60000.f);
  hi = __float2half_rn(v);
  lo = __float2half_rn(v - __half2float(hi));
}
__device__ __forceinline__ uint32_t pack_h2(__half a, __half b) {
  return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}

// developer profiling: cycles between marks, accumulated per role (written by one thread per role at the end)
#define PROF_DECL long long pt_ = 0, pacc_[8] = {0, 0, 0, 0, 0, 0, 0, 0}; const bool prof_on_ = a.prof != nullptr; if (prof_on_) pt_ = clock64();
#define PROF_MARK(i) if (prof_on_) { const long long n_ = clock64(); pacc_[i] += n_ - pt_; pt_ = n_; }
#define PROF_FLUSH(base) if (prof_on_) { for (int i_ = 0; i_ < 8; ++i_) a.prof[(size_t)blockIdx.x * 32 + (base) + i_] = pacc_[i_]; }

// Tile bookkeeping without per-tile integer divisions: tile t = blockIdx.x + k * gridDim.x  ->  (sample, tile in image)
struct TileIt {
  int b, timg;
  __device__ __forceinline__ void init(int t, int tiles_img) { b = t / tiles_img; timg = t - b * tiles_img; }
  __device__ __forceinline__ void advance(const Step2Geom& g, int tiles_img) {
    timg += g.step_t; b += g.step_b;
    if (timg >= tiles_img) { timg -= tiles_img; ++b; }
  }
  __device__ __forceinline__ void origin(const Step2Geom& g, int& r0, int& c0) const {
    const int ty = (int)(((uint32_t)timg * g.inv_tx) >> 16);
    r0 = ty * 16; c0 = (timg - ty * g.tiles_x) * 16;
  }
};

// NG epilogue groups of 256 threads (8 warps); warps 8*NG .. 8*NG+2 issue the MMAs (E, Z of M tile 0, Z of M tile 1:
// three issuers on three schedulers -- a lone issuing warp that shares its scheduler with busy epilogue warps runs
// at ~100 cycles per MMA instead of 40, tools/ubench_mma.cu); the next 4 warps are producers.
// EMIT (training forward): the coupling-network intermediates relu(d1), relu(d2) and h = Conv2dZeros output (after bias and
// gain) are also written to HBM (a.d_emit [B,HW,2], a.h_emit [B,HW,C]): the backward pass reads them from the tape
// instead of recomputing three convolutions per step.  A separate instantiation: the inference kernels are unchanged.
template <int C, bool X3, int NG, bool EMIT>
__global__ void __launch_bounds__(NG * 256 + 96 + (NG == 2 ? 192 : 128), 1)
flow_step_f16_kernel(Step2Args a, Step2Geom g) {
  constexpr int NP = (C + 15) / 16 * 16;
  // TMEM: E accumulators (4 M tiles x 32 columns) in two slots [0,256), released as soon as the gather read them;
  // Z accumulators (2 M tiles x NP columns) in NZ slots from column 256 on, released after the finishing epilogue.
  // Four Z slots (C <= 32) let each epilogue group gather its NEXT tile while the Z MMAs of its current one run.
  constexpr int ZW = 2 * NP;
  constexpr int NZ = (4 * ZW <= 256) ? 4 : 2;
  constexpr int kProdThreads = NG == 2 ? 192 : 128;       // producer warps: 6 with two epilogue groups, else 4
  constexpr int kThreads = NG * 256 + 96 + kProdThreads;
  constexpr int kMmaWarp = 8 * NG;        // E issuer (+ TMEM owner); kMmaWarp+1, +2: Z issuers
  constexpr int kProdWarp = 8 * NG + 3;
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int HW = a.H * a.W;
  const int tiles_img = g.tiles_x * g.tiles_y;

  uint8_t* A = smem + g.oA;
  uint8_t* WE = smem + g.oWE;
  uint8_t* WZ = smem + g.oWZ;
  float* Wm = reinterpret_cast<float*>(smem + g.oWm);       // C*C mix | nw | nb | bias3 | w2d[9] | inv scales[3] | 1/nw
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + g.oBar);
  uint64_t* a_full = bars;          // [4] ring slots: producers
  uint64_t* a_free = bars + 4;      // [4] ring slots: 3 commits (E issuer + 2 Z issuers)
  uint64_t* d_ready = bars + 8;     // [4] tiles in flight: epilogue group (256)
  uint64_t* e_full = bars + 12;     // [2] commit
  uint64_t* e_free = bars + 14;     // [2] epilogue group (256)
  uint64_t* z_full = bars + 16;     // [4] commit
  uint64_t* z_free = bars + 20;     // [4] epilogue group (256)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

  float* s_nw = Wm + C * C;
  float* s_nb = s_nw + C;
  float* s_b3 = s_nb + C;
  float* s_w2d = s_b3 + C;
  float* s_inv = s_w2d + 9;
  float* s_rnw = s_inv + 3;                                 // 1 / ActNorm weight

  // ------------------------------------------------------------------ one-time setup
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) { mbar_init(a_full + i, kProdThreads); mbar_init(a_free + i, 3); mbar_init(d_ready + i, kEpiThreads); }
    for (int i = 0; i < 2; ++i) { mbar_init(e_full + i, 1); mbar_init(e_free + i, kEpiThreads); }
    for (int i = 0; i < 4; ++i) { mbar_init(z_full + i, 2); mbar_init(z_free + i, kEpiThreads); }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc(tmem_slot, 512);
  {
    // zero every A buffer once: tail positions, padding channels and the d slots start as finite zeros
    uint4* z4 = reinterpret_cast<uint4*>(A);
    const int n4 = (int)((size_t)g.nbuf * g.bufA / 16);
    for (int i = tid; i < n4; i += kThreads) z4[i] = make_uint4(0u, 0u, 0u, 0u);
    // weights: packed global [hl][PLtot][..] -> shared [hl][2*KS][..] (only the K-steps this launch multiplies)
    const int pl = 2 * g.KS;
    const uint4* gE = reinterpret_cast<const uint4*>(a.wE);
    for (int i = tid; i < g.nhl * pl * 32; i += kThreads) {
      const int r = i % (pl * 32), hl = i / (pl * 32);
      reinterpret_cast<uint4*>(WE + hl * g.wE_hl)[r] = __ldg(gE + (size_t)hl * (g.gE_hl / 16) + r);
    }
    const uint4* gZ = reinterpret_cast<const uint4*>(a.wZ);
    const int per = pl * NP;
    for (int i = tid; i < 9 * g.nhl * per; i += kThreads) {
      const int r = i % per; int t = i / per; const int hl = t % g.nhl, tap = t / g.nhl;
      reinterpret_cast<uint4*>(WZ + tap * g.wZ_tap + hl * g.wZ_hl)[r] =
          __ldg(gZ + (size_t)tap * (g.gZ_tap / 16) + (size_t)hl * (g.gZ_hl / 16) + r);
    }
    for (int i = tid; i < C * C; i += kThreads) Wm[i] = a.wmat ? __ldg(a.wmat + i) : 0.f;
    for (int i = tid; i < C; i += kThreads) {
      s_nw[i] = a.nw ? __ldg(a.nw + i) : 1.f;
      s_nb[i] = a.nw ? __ldg(a.nb + i) : 0.f;
      s_b3[i] = __ldg(a.bias3 + i);
      s_rnw[i] = a.nw ? 1.f / __ldg(a.nw + i) : 1.f;
    }
    if (tid < 12) s_w2d[tid] = __ldg(a.wmisc + tid);        // w2d[9] then the three inverse scales
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int nmy = ((int)blockIdx.x < g.ntiles) ? (g.ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp < kMmaWarp) {
    // =========================================================== epilogue groups (256 threads each)
    const int grp = warp >> 3;
    const int etid = tid & 255, wg = etid >> 7, el = etid & 127;
    const int bar_id = 1 + grp;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const float gain = a.gain3 ? __ldg(a.gain3) : 1.f;
    const float inv1 = s_inv[0], inv2 = s_inv[1], inv3 = s_inv[2];
    float* Ds1 = reinterpret_cast<float*>(smem + g.oDsc + (size_t)grp * g.scratch) + 32 * 9;   // [rows -32..543][9]
    float* Ds2 = Ds1 + (kNPOSA + 64) * 9;
    float* D1 = Ds2 + (kNPOSA + 32) * 9;                                                      // [512] relu(d1)
    PROF_DECL

    // ---- G(k): E accumulators -> d1, d2 written into the A buffer of tile k.
    // The tap partials are stored PRE-SHIFTED: the partial of tap t computed at position q goes to row q - off(t),
    // so row p holds the nine terms of the 3x3 sum at p; out-of-image q store zeros (zero padding of the dense
    // layers), so the sums need no masks.
    auto gather = [&](int k, const TileIt& it) {
      const int s = k & 1, u = (k * g.KS + g.KS - 1) % g.nbuf;      // u: ring slot of this tile's x1+d chunk
      int r0, c0;
      it.origin(g, r0, c0);
      // positions of this thread in the d1 / d2 passes and the hoisted conditioning terms (prefetched)
      int pd[2], pcl[2];
      bool ok1[2], in1[2], ok2[2];
      float dcv1[2] = {0.f, 0.f}, dcv2[2] = {0.f, 0.f};
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int p = etid + j * kEpiThreads;
        const int rr = p / kRP, rc = p - rr * kRP;
        const int ir = r0 - 3 + rr, ic = c0 - 3 + rc;
        pd[j] = p;
        ok1[j] = p < kNPOS && rr >= 1 && rr <= 20 && rc >= 1 && rc <= 20;
        in1[j] = ir >= 0 && ir < a.H && ic >= 0 && ic < a.W;
        ok2[j] = p < kNPOS && rr >= 2 && rr <= 19 && rc >= 2 && rc <= 19;
        const int irc = min(max(ir, 0), a.H - 1), icc = min(max(ic, 0), a.W - 1);
        pcl[j] = (irc - (r0 - 3)) * kRP + (icc - (c0 - 3));
        ok2[j] = ok2[j] && pcl[j] >= 2 * kRP + 2 && pcl[j] < 20 * kRP;
        if (a.hoist) {
          const size_t hb = a.hoist_bstride ? (size_t)it.b * HW : 0;
          if (ok1[j] && in1[j]) dcv1[j] = __ldg(a.dc + (hb + (size_t)(ir * a.W + ic)) * a.dc_stride);
          if (ok2[j]) dcv2[j] = __ldg(a.dc + (hb + (size_t)(irc * a.W + icc)) * a.dc_stride + 1);
        }
      }
      PROF_MARK(0)
      mbar_wait(e_full + s, (uint32_t)((k >> 1) & 1));
      PROF_MARK(1)
      tc_fence_after();
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int mt = 2 * wg + j, q = mt * 128 + el;
        float v[32];
        tmem_ld32(tmem_base + lane_base + (uint32_t)(s * 128 + mt * 32), v);
        const int rr = q / kRP, rc = q - rr * kRP;
        const int ir = r0 - 3 + rr, ic = c0 - 3 + rc;
        const bool in = q < kNPOS && ir >= 0 && ir < a.H && ic >= 0 && ic < a.W;
        const float m1 = in ? 1.f : 0.f, m2 = m1;
        float* d1p = Ds1 + q * 9;
        float* d2p = Ds2 + q * 9;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const int off = (t / 3 - 1) * kRP + (t % 3 - 1);
          d1p[t - off * 9] = v[t] * m1;
          d2p[t - off * 9] = v[16 + t] * m2;
        }
      }
      tc_fence_before();
      mbar_arrive(e_free + s);
      named_bar_sync(bar_id, kEpiThreads);
      // d1 = relu(sum) on the halo-2 region (zero outside the image)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        if (ok1[j]) {
          float val = 0.f;
          if (in1[j]) {
            const float* e = Ds1 + pd[j] * 9;
            const float s0 = e[0] + e[1], s1 = e[2] + e[3], s2 = e[4] + e[5], s3 = e[6] + e[7];
            val = fmaxf(fmaf(((s0 + s1) + (s2 + s3)) + e[8], inv1, dcv1[j]), 0.f);
          }
          D1[pd[j]] = val;
        }
      }
      named_bar_sync(bar_id, kEpiThreads);
      // d2 on the halo-1 region, evaluated at the replicate-clamped pixel (Conv2dZeros pads by replication);
      // the (d1, d2) pair goes into K slots kd, kd+1 of the A buffer
      uint8_t* dslot = A + (size_t)u * g.bufA + kPLB + (size_t)(g.kd & 7) * 2;      // second plane of the chunk, channels 6,7
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        if (ok2[j]) {
          const int pc = pcl[j];
          const float* e = Ds2 + pc * 9;
          float sa = 0.f, sb = 0.f, da = dcv2[j], db = 0.f;          // sa/sb: scaled tensor-core partials, da/db: fp32 d1 taps
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            const int off = (t / 3 - 1) * kRP + (t % 3 - 1);
            if (t & 1) { sb += e[t]; db = fmaf(s_w2d[t], D1[pc + off], db); }
            else { sa += e[t]; da = fmaf(s_w2d[t], D1[pc + off], da); }
          }
          const float d1 = D1[pc], d2 = fmaxf(fmaf(sa + sb, inv2, da + db), 0.f);
          if constexpr (EMIT) {
            const int rr = pd[j] / kRP, rc = pd[j] - rr * kRP;
            const int ir = r0 - 3 + rr, ic = c0 - 3 + rc;
            if (rr >= 3 && rr <= 18 && rc >= 3 && rc <= 18 && ir < a.H && ic < a.W)      // tile interior, inside the image
              *reinterpret_cast<float2*>(a.d_emit + ((size_t)it.b * HW + (size_t)ir * a.W + ic) * 2) = make_float2(d1, d2);
          }
          __half h1, l1, h2, l2;
          split_h(d1, h1, l1); split_h(d2, h2, l2);
          *reinterpret_cast<uint32_t*>(dslot + (size_t)pd[j] * 16) = pack_h2(h1, h2);
          if (X3) *reinterpret_cast<uint32_t*>(dslot + g.hlA + (size_t)pd[j] * 16) = pack_h2(l1, l2);
        }
      }
      fence_proxy_async();
      mbar_arrive(d_ready + (k & 3));
      PROF_MARK(2)
    };

    // ---- F(k): h -> coupling, 1x1 mix, ActNorm, store, log-det partial
    auto finish = [&](int k, const TileIt& it) {
      const int s = k & (NZ - 1);
      const uint32_t zpar = (uint32_t)((k / NZ) & 1);
      int r0, c0;
      it.origin(g, r0, c0);
      const int b = it.b;
      const int ir = r0 + (el >> 3), ic = c0 + 8 * wg + (el & 7);
      const bool valid = ir < a.H && ic < a.W;
      const size_t pix = valid ? (size_t)b * HW + (size_t)ir * a.W + ic : 0;
      float v[C];
      constexpr int NHC = C <= 24 ? C / 4 : 1;
      float4 hcv[NHC];
      const float4* hcp = (a.hoist && valid) ? reinterpret_cast<const float4*>(a.hc + ((a.hoist_bstride ? (size_t)b * HW : 0) + (size_t)(ir * a.W + ic)) * a.hc_stride) : nullptr;
      if (valid) {
        const float4* y4 = reinterpret_cast<const float4*>(a.y_in + pix * C);
#pragma unroll
        for (int q = 0; q < C / 4; ++q) { float4 t = __ldg(y4 + q); v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w; }
      }
      if (C <= 24) {
#pragma unroll
        for (int q = 0; q < NHC; ++q) hcv[q] = hcp ? __ldg(hcp + q) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      PROF_MARK(3)
      mbar_wait(z_full + s, zpar);
      PROF_MARK(4)
      tc_fence_after();
      float ldsum = 0.f;
      const uint32_t trow = tmem_base + lane_base + (uint32_t)(256 + s * ZW + wg * NP);
#pragma unroll
      for (int n0 = 0; n0 < NP; n0 += 16) {
        float h[16];
        tmem_ld16(trow + n0, h);
        if (valid) {
#pragma unroll
          for (int q = 0; q < 16; q += 4) {
            if (n0 + q < C) {
              float4 hc4;
              if (C <= 24) hc4 = hcv[(n0 + q) / 4 < NHC ? (n0 + q) / 4 : 0];
              else hc4 = hcp ? __ldg(hcp + (n0 + q) / 4) : make_float4(0.f, 0.f, 0.f, 0.f);
              const float hv[4] = {fmaf(h[q], inv3, hc4.x), fmaf(h[q + 1], inv3, hc4.y), fmaf(h[q + 2], inv3, hc4.z),
                                   fmaf(h[q + 3], inv3, hc4.w)};
#pragma unroll
              for (int e = 0; e < 4; e += 2) {
                const float shift = (hv[e] + s_b3[n0 + q + e]) * gain;            // h[:, 0::2]
                const float raw = (hv[e + 1] + s_b3[n0 + q + e + 1]) * gain;      // h[:, 1::2]
                if constexpr (EMIT) *reinterpret_cast<float2*>(a.h_emit + pix * C + n0 + q + e) = make_float2(shift, raw);
                const float la = 2.f * __fdividef(raw, 1.f + fabsf(raw));      // 2*softsign; rcp.approx, 2 ulp
                ldsum += la;
                const int j = C / 2 + (n0 + q + e) / 2;
                v[j] = a.reverse ? fmaf(v[j], __expf(-la), -shift) : (v[j] + shift) * __expf(la);   // |la| <= 2: ex2.approx, 2 ulp
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(z_free + s);
      // Wide levels on short images (C >= 32, H <= 8: the lower half of every 16-row M tile is outside the image):
      // the C x C mix of a pixel is split between its own thread (output rows [0, C/2)) and the idle thread 64 lanes
      // up (rows [C/2, C)); the post-coupling state goes through shared memory (the gather scratch is free here).
      const bool split_mix = C >= 32 && a.H <= 8 && a.wmat != nullptr;
      if (split_mix) {
        constexpr int VP = C + 4;
        float* Vx = Ds1 - 32 * 9;
        if (valid) {
          if (!a.reverse && a.nw) {
#pragma unroll
            for (int q = 0; q < C; ++q) v[q] = fmaf(s_nw[q], v[q], s_nb[q]);
          }
          float4* vd = reinterpret_cast<float4*>(Vx + (wg * 64 + (el & 63)) * VP);
#pragma unroll
          for (int q = 0; q < C / 4; ++q) vd[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
        named_bar_sync(bar_id, kEpiThreads);
        const int px = el & 63, part = el >> 6;
        const int ir2 = r0 + (px >> 3), ic2 = c0 + 8 * wg + (px & 7);
        if (ir2 < a.H && ic2 < a.W) {
          const float4* vs = reinterpret_cast<const float4*>(Vx + (wg * 64 + px) * VP);
#pragma unroll
          for (int q = 0; q < C / 4; ++q) { const float4 t = vs[q]; v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w; }
          float* yo = a.y_out + ((size_t)b * HW + (size_t)ir2 * a.W + ic2) * C;
#pragma unroll 1
          for (int r4 = part * (C / 2); r4 < (part + 1) * (C / 2); r4 += 4) {
            float o[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float sum = 0.f;
              const float4* w4 = reinterpret_cast<const float4*>(Wm + (r4 + q) * C);
#pragma unroll
              for (int kk = 0; kk < C / 4; ++kk) {
                const float4 w = w4[kk];
                sum = fmaf(w.x, v[4 * kk], sum); sum = fmaf(w.y, v[4 * kk + 1], sum);
                sum = fmaf(w.z, v[4 * kk + 2], sum); sum = fmaf(w.w, v[4 * kk + 3], sum);
              }
              if (a.reverse && a.nw) sum = (sum - s_nb[r4 + q]) * s_rnw[r4 + q];
              o[q] = sum;
            }
            *reinterpret_cast<float4*>(yo + r4) = make_float4(o[0], o[1], o[2], o[3]);
          }
        }
        named_bar_sync(bar_id, kEpiThreads);       // the scratch goes back to the next gather
      } else if (valid) {
        if (!a.reverse && a.nw) {
#pragma unroll
          for (int q = 0; q < C; ++q) v[q] = fmaf(s_nw[q], v[q], s_nb[q]);
        }
        float* yo = a.y_out + pix * C;
        if (a.wmat) {
#pragma unroll 1
          for (int r4 = 0; r4 < C; r4 += 4) {
            float o[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float sum = 0.f;
              const float4* w4 = reinterpret_cast<const float4*>(Wm + (r4 + q) * C);
#pragma unroll
              for (int kk = 0; kk < C / 4; ++kk) {
                const float4 w = w4[kk];
                sum = fmaf(w.x, v[4 * kk], sum); sum = fmaf(w.y, v[4 * kk + 1], sum);
                sum = fmaf(w.z, v[4 * kk + 2], sum); sum = fmaf(w.w, v[4 * kk + 3], sum);
              }
              if (a.reverse && a.nw) sum = (sum - s_nb[r4 + q]) * s_rnw[r4 + q];
              o[q] = sum;
            }
            *reinterpret_cast<float4*>(yo + r4) = make_float4(o[0], o[1], o[2], o[3]);
          }
        } else {
#pragma unroll
          for (int q = 0; q < C; q += 4) {
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) o[e] = (a.reverse && a.nw) ? (v[q + e] - s_nb[q + e]) * s_rnw[q + e] : v[q + e];
            *reinterpret_cast<float4*>(yo + q) = make_float4(o[0], o[1], o[2], o[3]);
          }
        }
      }
      if (a.ld_part) {          // one partial per epilogue warp: slot = tile * 8 + warp (summed in fixed order later)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ldsum += __shfl_xor_sync(0xffffffffu, ldsum, o);
        if (lane == 0) a.ld_part[(size_t)b * a.ld_stride + it.timg * 8 + (warp & 7)] = ldsum;
      }
      PROF_MARK(5)
    };

    if (NG == 2) {
      TileIt itf, itg;
      itf.init(blockIdx.x + grp * gridDim.x, tiles_img); itg = itf;
      if (NZ == 4 && g.pipelined) {          // gather of this group's next tile overlaps the Z MMAs of the current one
        if (grp < nmy) gather(grp, itg);
        for (int k = grp; k < nmy; k += 2) {
          itg.advance(g, tiles_img); itg.advance(g, tiles_img);
          if (k + 2 < nmy) gather(k + 2, itg);
          finish(k, itf);
          itf.advance(g, tiles_img); itf.advance(g, tiles_img);
        }
      } else {
        for (int k = grp; k < nmy; k += 2) {
          gather(k, itf);
          finish(k, itf);
          itf.advance(g, tiles_img); itf.advance(g, tiles_img);
        }
      }
    } else if (g.pipelined) {
      TileIt itf, itg;
      itf.init(blockIdx.x, tiles_img); itg = itf;
      if (nmy > 0) gather(0, itg);
      for (int k = 0; k < nmy; ++k) {
        itg.advance(g, tiles_img);
        if (k + 1 < nmy) gather(k + 1, itg);
        finish(k, itf);
        itf.advance(g, tiles_img);
      }
    } else {
      TileIt it;
      it.init(blockIdx.x, tiles_img);
      for (int k = 0; k < nmy; ++k) { gather(k, it); finish(k, it); it.advance(g, tiles_img); }
    }
    if (etid == 0 && grp == 0) { PROF_FLUSH(0) }
  } else if (warp < kProdWarp) {
    // =========================================================== MMA issue: one elected lane per issuer warp
    const int role = warp - kMmaWarp;         // 0: E, 1: Z of M tile 0 (columns 0-7), 2: Z of M tile 1 (columns 8-15)
    if (elect_one()) {
      const uint64_t hlA16 = g.hlA >> 4;
      const int nsrc1 = g.KS - g.KSy;            // conditioning chunks staged (and consumed) before the x1+d chunks
      PROF_DECL
      if (role == 0) {
        const uint32_t idE = idesc_f16(32);
        const uint64_t bE0 = make_desc(smem_u32(WE), 512, 128);
        const uint64_t wEhl16 = g.wE_hl >> 4;
        int c = 0;                                // running chunk counter
        for (int k = 0; k < nmy; ++k) {
          const int s = k & 1;
          PROF_MARK(0)
          if (k >= 2) mbar_wait(e_free + s, (uint32_t)(((k >> 1) - 1) & 1));
          PROF_MARK(2)
          const uint32_t tE = tmem_base + (uint32_t)(s * 128);
          for (int j = 0; j < g.KS; ++j, ++c) {
            const int u = c % g.nbuf;
            const int wk = j < nsrc1 ? g.KSy + j : j - nsrc1;        // K-step of the packed weights
            mbar_wait(a_full + u, (uint32_t)((c / g.nbuf) & 1));
            tc_fence_after();
            const uint64_t aE0 = make_desc(smem_u32(A + (size_t)u * g.bufA), kPLB, 128);
            const uint64_t bd = bE0 + (uint64_t)wk * (uint64_t)(2 * 512 >> 4);
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) {
              const uint64_t ad = aE0 + (uint64_t)(mt * 128);          // 128 positions x 16 B, in 16-byte units
              mma_f16(tE + (uint32_t)(mt * 32), ad, bd, idE, j > 0 ? 1u : 0u);
              if (X3) {
                mma_f16(tE + (uint32_t)(mt * 32), ad + hlA16, bd, idE, 1u);
                mma_f16(tE + (uint32_t)(mt * 32), ad, bd + wEhl16, idE, 1u);
              }
            }
            mma_commit(a_free + u);
          }
          mma_commit(e_full + s);
          PROF_MARK(3)
        }
        PROF_FLUSH(8)
      } else {
        const int mt = role - 1;
        const uint32_t idZ = idesc_f16(NP);
        const uint64_t bZ0 = make_desc(smem_u32(WZ), (uint32_t)NP * 16u, 128);
        const uint64_t wZhl16 = g.wZ_hl >> 4, wZtap16 = g.wZ_tap >> 4;
        int c = 0;
        for (int k = 0; k < nmy; ++k) {
          const int s = k & (NZ - 1);
          PROF_MARK(0)
          if (k >= NZ) mbar_wait(z_free + s, (uint32_t)(((k / NZ) - 1) & 1));
          PROF_MARK(5)
          const uint32_t tZ = tmem_base + (uint32_t)(256 + s * ZW + mt * NP);
          for (int j = 0; j < g.KS; ++j, ++c) {
            const int u = c % g.nbuf;
            const int wk = j < nsrc1 ? g.KSy + j : j - nsrc1;
            mbar_wait(a_full + u, (uint32_t)((c / g.nbuf) & 1));
            if (j == g.KS - 1) {                   // the chunk that carries d1, d2: wait for the gather of this tile
              PROF_MARK(6)
              mbar_wait(d_ready + (k & 3), (uint32_t)((k >> 2) & 1));
              PROF_MARK(4)
            }
            tc_fence_after();
            const uint64_t aZ0 = make_desc(smem_u32(A + (size_t)u * g.bufA), kPLB, kRP * 16) + (uint64_t)(8 * mt);
            const uint64_t bk = bZ0 + (uint64_t)wk * (uint64_t)(2 * NP);   // 2 planes x NP rows x 16 B, in 16-byte units
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const uint64_t ad = aZ0 + (uint64_t)((3 + tap / 3 - 1) * kRP + 3 + tap % 3 - 1);
              const uint64_t bd = bk + (uint64_t)tap * wZtap16;
              mma_f16(tZ, ad, bd, idZ, (tap > 0 || j > 0) ? 1u : 0u);
              if (X3) {
                mma_f16(tZ, ad + hlA16, bd, idZ, 1u);
                mma_f16(tZ, ad, bd + wZhl16, idZ, 1u);
              }
            }
            mma_commit(a_free + u);
          }
          mma_commit(z_full + s);
          PROF_MARK(6)
        }
        if (role == 1) { PROF_FLUSH(24) }
      }
    }
  } else {
    // =========================================================== producers (128 threads): stage relu(t) as fp16 hi/lo
    const int ptid = tid - kProdWarp * 32;
    bool ovf = false;
    const int npl0 = 2 * g.KSy;
    const int nsrc1 = g.KS - g.KSy;
    TileIt it;
    it.init(blockIdx.x, tiles_img);
    PROF_DECL
    int c = 0;                                    // running chunk counter
    for (int k = 0; k < nmy; ++k) {
      const int b = it.b;
      int r0, c0;
      it.origin(g, r0, c0);
      it.advance(g, tiles_img);
      for (int j = 0; j < g.KS; ++j, ++c) {
        const int u = c % g.nbuf, use = c / g.nbuf;
        const int wk = j < nsrc1 ? g.KSy + j : j - nsrc1;            // K-step in the K layout [src0 + d | src1]
        PROF_MARK(0)
        if (use >= 1) mbar_wait(a_free + u, (uint32_t)((use - 1) & 1));
        PROF_MARK(1)
        uint8_t* Ab = A + (size_t)u * g.bufA;
        // 2 planes x 484 positions; batches of up to 4 items per thread, loads first
        for (int it0 = ptid; it0 < 2 * kNPOS; it0 += kProdThreads * 4) {
          float v[4][8];
          int pos[4], lpl[4];
          bool relu[4], hasd[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int itx = it0 + q * kProdThreads;
            pos[q] = -1; lpl[q] = 0; relu[q] = true; hasd[q] = false;
#pragma unroll
            for (int e = 0; e < 8; ++e) v[q][e] = 0.f;
            if (itx < 2 * kNPOS) {
              const int lp = itx >= kNPOS ? 1 : 0, p = itx - lp * kNPOS;
              const int plane = 2 * wk + lp;                          // plane in the K layout
              const bool s1 = plane >= npl0;
              const ConvSrc& sc = a.src[s1 ? 1 : 0];
              const int ch = (s1 ? plane - npl0 : plane) * 8;
              const int nv = min(8, sc.nch - ch);                     // <= 0: pure padding (or the d-slot plane)
              pos[q] = p; lpl[q] = lp; relu[q] = sc.relu != 0;
              hasd[q] = !s1 && (g.kd >> 3) == plane;
              if (nv <= 0 && g.KS == 1) pos[q] = -1;      // one chunk type only: its padding plane was zeroed once and stays zero
              if (nv > 0) {
                const int rr = p / kRP, rc = p - rr * kRP;
                const int r = min(max(r0 - 3 + rr, 0), a.H - 1), cc = min(max(c0 - 3 + rc, 0), a.W - 1);
                const size_t pixi = (sc.bshared ? 0 : (size_t)b * HW) + (size_t)r * a.W + cc;
                const float* ptr = sc.p + pixi * sc.cstride + sc.coff + ch;
                if ((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (nv & 1) == 0) {
                  if (nv >= 4) { const float4 t0 = __ldg(reinterpret_cast<const float4*>(ptr)); v[q][0] = t0.x; v[q][1] = t0.y; v[q][2] = t0.z; v[q][3] = t0.w; }
                  else { const float2 t0 = __ldg(reinterpret_cast<const float2*>(ptr)); v[q][0] = t0.x; v[q][1] = t0.y; }
                  if (nv == 8) { const float4 t1 = __ldg(reinterpret_cast<const float4*>(ptr) + 1); v[q][4] = t1.x; v[q][5] = t1.y; v[q][6] = t1.z; v[q][7] = t1.w; }
                  else if (nv == 6) { const float2 t1 = __ldg(reinterpret_cast<const float2*>(ptr) + 2); v[q][4] = t1.x; v[q][5] = t1.y; }
                } else {
#pragma unroll
                  for (int e = 0; e < 8; ++e) if (e < nv) v[q][e] = __ldg(ptr + e);
                }
              }
            }
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (pos[q] < 0) continue;
            uint32_t ph[4], pl[4];
            float umax = 0.f;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float u0 = v[q][2 * e], u1 = v[q][2 * e + 1];
              umax = fmaxf(umax, relu[q] ? fmaxf(u0, u1) : fmaxf(fabsf(u0), fabsf(u1)));      // overflow detection: one compare per unit
              float y0 = fminf(u0, 60000.f), y1 = fminf(u1, 60000.f);
              y0 = fmaxf(y0, relu[q] ? 0.f : -60000.f); y1 = fmaxf(y1, relu[q] ? 0.f : -60000.f);
              const __half2 h2 = __floats2half2_rn(y0, y1);
              const float2 hf = __half22float2(h2);
              const __half2 l2 = __floats2half2_rn(y0 - hf.x, y1 - hf.y);
              ph[e] = *reinterpret_cast<const uint32_t*>(&h2);
              pl[e] = *reinterpret_cast<const uint32_t*>(&l2);
            }
            ovf = ovf || umax > 60000.f;
            // the d slots (last 32-bit word of their plane) are owned by the epilogue warps
            uint8_t* dst = Ab + (size_t)lpl[q] * kPLB + (size_t)pos[q] * 16;
            if (hasd[q]) {
              uint32_t* d32 = reinterpret_cast<uint32_t*>(dst);
              d32[0] = ph[0]; d32[1] = ph[1]; d32[2] = ph[2];
              if (X3) {
                uint32_t* l32 = reinterpret_cast<uint32_t*>(dst + g.hlA);
                l32[0] = pl[0]; l32[1] = pl[1]; l32[2] = pl[2];
              }
            } else {
              *reinterpret_cast<uint4*>(dst) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
              if (X3) *reinterpret_cast<uint4*>(dst + g.hlA) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
            }
          }
        }
        fence_proxy_async();
        mbar_arrive(a_full + u);
        PROF_MARK(2)
      }
    }
    if (ptid == 0) { PROF_FLUSH(16) }
    if (ovf && a.overflow) atomicOr(a.overflow, 1u);       // an operand left the fp16 range and was clamped: sticky flag
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ------------------------------------------------------------------ host side
void step2_klayout(int nch0, int nch1, int& KSy, int& KS1, int& kd) {
  KSy = (nch0 + 2 + 15) / 16;
  KS1 = (nch1 + 15) / 16;
  kd = KSy * 16 - 2;
}
size_t step2_wE_floats(int nch0, int nch1) {
  int KSy, KS1, kd; step2_klayout(nch0, nch1, KSy, KS1, kd);
  return (size_t)2 * 2 * (KSy + KS1) * 32 * 16 / 4;             // [hl][planes][32][16 B]
}
size_t step2_wZ_floats(int nch0, int nch1, int C) {
  int KSy, KS1, kd; step2_klayout(nch0, nch1, KSy, KS1, kd);
  const int NP = (C + 15) / 16 * 16;
  return (size_t)9 * 2 * 2 * (KSy + KS1) * NP * 16 / 4;         // [tap][hl][planes][NP][16 B]
}

static int sm_count() {
  int dev = 0, nsm = 148;
  cudaGetDevice(&dev);
  static int cached_sm[64] = {0};
  if (dev >= 0 && dev < 64) {
    if (!cached_sm[dev]) cudaDeviceGetAttribute(&cached_sm[dev], cudaDevAttrMultiProcessorCount, dev);
    if (cached_sm[dev] > 0) nsm = cached_sm[dev];
  }
  return nsm;
}

static bool make_geom2(const Step2Args& a, Step2Geom& g, int grid) {
  int KSy, KS1, kd;
  const int nch0 = a.src[0].nch, nch1 = a.nsrc > 1 ? a.src[1].nch : 0;
  step2_klayout(nch0, nch1, KSy, KS1, kd);
  const int NP = (a.C + 15) / 16 * 16;
  g.KSy = KSy; g.kd = kd; g.PLtot = 2 * (KSy + KS1);
  g.KS = (a.hoist || a.nsrc < 2) ? KSy : KSy + KS1;
  g.nhl = a.x3 ? 2 : 1;
  g.hlA = 2 * kPLB;                 // one ring slot = one K-step: [hi | lo][2 planes][512 positions][16 B]
  g.bufA = g.hlA * g.nhl;
  g.tiles_x = cdiv(a.W, 16); g.tiles_y = cdiv(a.H, 16);
  const int tiles_img = g.tiles_x * g.tiles_y;
  g.ntiles = tiles_img * a.B;
  if (grid <= 0) grid = std::min(g.ntiles, sm_count());
  g.step_b = grid / tiles_img; g.step_t = grid % tiles_img;
  g.inv_tx = (uint32_t)((65536 + g.tiles_x - 1) / g.tiles_x);      // exact floor(t / tiles_x) for t < 4096
  if (tiles_img >= 4096) return false;
  g.gE_hl = (uint32_t)g.PLtot * 512; g.gZ_hl = (uint32_t)g.PLtot * NP * 16; g.gZ_tap = 2 * g.gZ_hl;
  g.wE_hl = (uint32_t)(2 * g.KS) * 512; g.wZ_hl = (uint32_t)(2 * g.KS) * NP * 16; g.wZ_tap = g.nhl * g.wZ_hl;
  g.scratch = (uint32_t)(2 * (kNPOSA + 64) * 9 + kNPOSA) * 4;
  // preference: two epilogue groups (narrow levels only: register budget) with the deepest K-step ring that fits
  // (at most 4 slots; at least KS so that a whole tile can be resident)
  // one epilogue group by default: what this kernel runs today (LSTM tails with 3 K-steps of per-sample input, the training
  // forward) is bound by its 4 staging warps, and a second group of 8 epilogue warps takes issue slots from them (LSTM tails of
  // a S = 4096 call 3.04 -> 2.80 ms, training unchanged); TMG_STEP2_NG2=1 restores round 1's two groups at C <= 24 (A/B runs)
  static const bool ng2 = [] { const char* e = getenv("TMG_STEP2_NG2"); return e && e[0] == '1'; }();
  for (int ng = ((a.C <= 24 && ng2) ? 2 : 1); ng >= 1; --ng) {
    for (int nbuf = 4; nbuf >= std::max(g.KS, 2); --nbuf) {
      uint32_t off = 0;
      auto take = [&](uint32_t n) { uint32_t o = off; off += (n + 127) / 128 * 128; return o; };
      g.nbuf = nbuf; g.ngroups = ng;
      g.oA = take((uint32_t)nbuf * g.bufA);
      g.oDsc = take((uint32_t)ng * g.scratch);
      g.oWE = take(g.nhl * g.wE_hl);
      g.oWZ = take(9u * g.wZ_tap);
      g.oWm = take((uint32_t)(a.C * a.C + 4 * a.C + 9 + 3) * 4);
      g.oBar = take(24 * 8 + 16);
      g.total = off;
      // pipelining across tiles needs room for the next tile's chunks while this tile's x1+d chunk is held
      if (g.total <= 227 * 1024) { g.pipelined = nbuf >= g.KS + 1; return true; }
    }
  }
  return false;
}

int step2_ld_slots(int H, int W) { return 8 * cdiv(H, 16) * cdiv(W, 16); }

bool step2_supported(const Step2Args& a) {
  Step2Geom g{};
  return a.C % 4 == 0 && a.C <= kMaxC && make_geom2(a, g, 148);
}

int launch_step2(const Step2Args& a_in, cudaStream_t st) {
  if (a_in.B <= 0) return TMG_OK;
  Step2Args a = a_in;
  static const bool prof_env = getenv("TMG_STEP2_PROF") != nullptr;
  static long long* prof_buf = nullptr;
  static int prof_left = 0;
  if (prof_env && !prof_buf) { cudaMalloc(&prof_buf, 148 * 32 * sizeof(long long)); prof_left = atoi(getenv("TMG_STEP2_PROF")); }
  if (prof_env && prof_left > 0) { cudaMemsetAsync(prof_buf, 0, 148 * 32 * sizeof(long long), st); a.prof = prof_buf; }
  Step2Geom g{};
  const int tiles = cdiv(a.W, 16) * cdiv(a.H, 16) * a.B;
  const int grid = std::min(tiles, sm_count());
  if (a.C % 4 || a.C > kMaxC || !make_geom2(a, g, grid)) {
    set_error("fused fp16 flow step: unsupported shape (C=%d, %dx%d)", a.C, a.H, a.W);
    return TMG_ERR_UNSUPPORTED;
  }
  const bool emit = a.d_emit != nullptr && a.h_emit != nullptr;
#define TMG_S2E(CC, XX, GG, EE)                                                                                                      \
  {                                                                                                                                  \
    TMG_SMEM_ATTR(flow_step_f16_kernel<CC, XX, GG, EE>, 227 * 1024); \
    flow_step_f16_kernel<CC, XX, GG, EE><<<grid, GG * 256 + 96 + (GG == 2 ? 192 : 128), g.total, st>>>(a, g);                        \
  }
#define TMG_S2(CC, XX, GG) { if (emit) TMG_S2E(CC, XX, GG, true) else TMG_S2E(CC, XX, GG, false) }
#define TMG_S2N(CC)                                                                              \
  case CC:                                                                                       \
    if (g.ngroups == 2) { if (a.x3) TMG_S2(CC, true, 2) else TMG_S2(CC, false, 2) }              \
    else { if (a.x3) TMG_S2(CC, true, 1) else TMG_S2(CC, false, 1) }                             \
    break;
#define TMG_S2W(CC) case CC: if (a.x3) TMG_S2(CC, true, 1) else TMG_S2(CC, false, 1) break;
  switch (a.C) {
    TMG_S2N(4) TMG_S2N(8) TMG_S2N(12) TMG_S2N(16) TMG_S2N(24) TMG_S2W(32) TMG_S2W(48) TMG_S2W(64)
    default:
      set_error("fused fp16 flow step: %d channels not supported", a.C);
      return TMG_ERR_UNSUPPORTED;
  }
#undef TMG_S2W
#undef TMG_S2N
#undef TMG_S2
#undef TMG_S2E
  TMG_LAUNCH_CHECK();
  if (a.prof) {      // developer profiling: average cycles per role and phase over the CTAs of this launch
    --prof_left;
    static long long h[148 * 32];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, prof_buf, sizeof(h), cudaMemcpyDeviceToHost);
    double s[32] = {0};
    for (int b = 0; b < grid; ++b) for (int i = 0; i < 32; ++i) s[i] += (double)h[b * 32 + i] / grid;
    const double tiles = (double)g.ntiles / grid;
    fprintf(stderr, "[step2 prof] C=%d x3=%d hoist=%d KS=%d nbuf=%d ng=%d tiles/CTA=%.1f | cycles per tile: "
            "EPI(g0) pre %.0f wait_e %.0f gather %.0f pre_f %.0f wait_z %.0f finish %.0f | "
            "MMA gap %.0f wait_a %.0f wait_efree %.0f issueE %.0f wait_d %.0f wait_zfree %.0f issueZ %.0f | PROD gap %.0f wait_free %.0f stage %.0f\n",
            a.C, a.x3, a.hoist, g.KS, g.nbuf, g.ngroups, tiles,
            s[0] / tiles * g.ngroups, s[1] / tiles * g.ngroups, s[2] / tiles * g.ngroups, s[3] / tiles * g.ngroups, s[4] / tiles * g.ngroups, s[5] / tiles * g.ngroups,
            (s[8] + s[24]) / tiles, s[9] / tiles, s[10] / tiles, s[11] / tiles, s[28] / tiles, s[29] / tiles, s[30] / tiles,
            s[16] / tiles, s[17] / tiles, s[18] / tiles);
  }
  return TMG_OK;
}

}  // namespace tmg
