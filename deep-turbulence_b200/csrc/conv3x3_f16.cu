// 3x3 convolution (stride 1) as a persistent fp16-operand tcgen05 implicit GEMM (sm_100a), second generation of
// conv3x3_tc.cu.  Used for the two heavy convolutions of the LSTM flow step
//   ConvLSTM gate conv   nn/modules/convLSTM.py:44,74   (N = 4*rec = 256, ConvLSTM cell update fused in the epilogue)
//   LSTM_out_conv        nn/modules/convLSTM.py:129     (N = C/2+cond, bias + ReLU)
// and for any other stride-1 conv of the path (bias / gain / activation epilogue).
//
//   * CTA tile = 16x16 output pixels = two MMA M tiles of 16 rows x 8 columns: the A descriptor's stride between
//     8-row groups (SBO) is one padded row of the staged tile, so the 1-pixel halo costs (18*18)/256 = 1.27x
//     staging instead of the 2.05x of the linear-index trick, and no accumulator row is wasted.
//   * operands are fp16 (kind::f16, K = 16 per MMA); "x3" splits both operands into hi + lo halves
//     (hi*hi + lo*hi + hi*lo, fp32 accumulate, weights pre-scaled by a power of two): fp32-grade results at half
//     the tensor work of the 3xTF32 split.
//   * persistent CTAs, warp-specialised: 4 producer warps stage activation K-steps (16 channels) into a ring,
//     one warp streams packed weights per (K-step, tap) with cp.async.bulk, two warps issue the MMAs (one per
//     M tile, elect.sync), eight warps run the epilogue (exp2/rcp based sigmoid/tanh for the LSTM cell).
//   * every source of the virtual concatenation starts on an 8-channel plane boundary (weights packed to match),
//     so staging is vector loads + cvt, no per-channel source selection.
#include <cuda_fp16.h>

#include <type_traits>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace tmg {

constexpr int kCvRP = 18;                 // staged tile pitch: 16 + 2
constexpr int kCvNPOS = 324;
constexpr int kCvNPOSA = 332;               // plane pitch 5312 B = 16 banks mod 32: the two planes of a K-step do not collide
constexpr uint32_t kCvPLB = kCvNPOSA * 16;        // bytes of one 8-channel plane of a K-step
// warps 0-7 epilogue, 8-9 MMA, 10 weights, 11.. producers (NPW warps: 4 when the MMAs dominate (wide N), 8 when the
// activation staging does (narrow N))
constexpr int kCvNA = 3;                  // activation K-step ring
constexpr int kCvMaxNB = 8;               // weight stage ring (upper bound)

struct ConvF16Geom {
  int KS;                 // K-steps of 16 channels
  int nhl, nb, tps;       // weight ring depth; taps per weight stage (1, 3 or 9)
  int tiles_x, tiles_y, ntiles, step_b, step_t;
  uint32_t inv_tx;
  uint32_t hlA, bufA;     // bytes: hi->lo distance inside an A K-step buffer, buffer size
  uint32_t stageB, hlB;   // bytes of one weight stage (hi [+ lo]); hi->lo distance
  uint32_t gstageB;       // bytes between packed (global) stages (always hi + lo)
  uint32_t oA, oB, oMisc, oBar, total;
  int plane0[3];          // first 8-channel plane of each source
  int nplanes[3];
};

#ifdef TMG_GT_PROFILE
// developer build (-DTMG_GT_PROFILE, TMG_CV_PROF=n at run time): cycles per role spent in each wait
__device__ long long cv_prof[148 * 16];
#define CVP_T0() const long long cvp_t0 = clock64()
#define CVP_ADD(slot) cvp_acc[slot] += clock64() - cvp_t0
#define CVP_DECL() long long cvp_acc[4] = {0, 0, 0, 0}; const long long cvp_start = clock64()
#else
#define CVP_T0()
#define CVP_ADD(slot)
#define CVP_DECL()
#endif

struct CvTileIt {
  int b, timg;
  __device__ __forceinline__ void init(int t, int tiles_img) { b = t / tiles_img; timg = t - b * tiles_img; }
  __device__ __forceinline__ void advance(const ConvF16Geom& g, int tiles_img) {
    timg += g.step_t; b += g.step_b;
    if (timg >= tiles_img) { timg -= tiles_img; ++b; }
  }
  __device__ __forceinline__ void origin(const ConvF16Geom& g, int& r0, int& c0) const {
    const int ty = (int)(((uint32_t)timg * g.inv_tx) >> 16);
    r0 = ty * 16; c0 = (timg - ty * g.tiles_x) * 16;
  }
};

template <bool X3, int NPW>
__global__ void __launch_bounds__((11 + NPW) * 32, 1)
conv3x3_f16_kernel(ConvF16Args a, ConvF16Geom g) {
  constexpr int kCvThreads = (11 + NPW) * 32;
  // Producer warps work in groups of 4 (128 threads): group gi stages the K-steps with (running index) % groups == gi, so
  // with 8 warps two K-steps are in flight at once -- at narrow N the kernel is bound by the DRAM round trip of a K-step,
  // not by the tensor pipe.
  constexpr int kNPG = NPW / 4;           // producer groups
  constexpr int kNPT = 128;               // threads staging one K-step
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int HW = a.H * a.W;
  const int tiles_img = g.tiles_x * g.tiles_y;
  const int NP = a.npad;

  uint8_t* As = smem + g.oA;                 // kCvNA K-step buffers [hl][2 planes][NPOSA][16 B]
  uint8_t* Bs = smem + g.oB;                 // nb weight stages [hl][2 planes][NP][16 B]
  float* s_bias = reinterpret_cast<float*>(smem + g.oMisc);      // [NP]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + g.oBar);
  uint64_t* a_full = bars;                   // [3]  producers (128)
  uint64_t* a_free = bars + 3;               // [3]  2 commits
  uint64_t* b_full = bars + 6;               // [8]  tx
  uint64_t* b_free = bars + 14;              // [8]  2 commits
  uint64_t* acc_full = bars + 22;            // [2] 2 commits
  uint64_t* acc_free = bars + 24;            // [2] epilogue (256)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 26);
  // N <= 128: two accumulator sets (2 x 2 M tiles x 128 columns), the MMAs of tile k + 1 run under the epilogue of tile k
  // (measured at N = 48, 5 K-steps: the issuer waited 44 % of a tile for the epilogue, the epilogue 57 % for the MMAs)
  const bool acc2 = NP <= 128;

  if (tid == 0) {
    for (int i = 0; i < kCvNA; ++i) { mbar_init(a_full + i, kNPT); mbar_init(a_free + i, 2); }
    for (int i = 0; i < kCvMaxNB; ++i) { mbar_init(b_full + i, 1); mbar_init(b_free + i, 2); }
    for (int i = 0; i < 2; ++i) { mbar_init(acc_full + i, 2); mbar_init(acc_free + i, 256); }
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc(tmem_slot, 512);
  for (int i = tid; i < NP; i += kCvThreads) s_bias[i] = (a.bias && i < a.cout) ? __ldg(a.bias + i) : 0.f;
  {   // zero the A ring once (positions >= 324 of every plane stay zero)
    uint4* z4 = reinterpret_cast<uint4*>(As);
    const int n4 = (int)((size_t)kCvNA * g.bufA / 16);
    for (int i = tid; i < n4; i += kCvThreads) z4[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int nmy = ((int)blockIdx.x < g.ntiles) ? (g.ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int nstage_tile = g.KS * 9;

  if (warp < 8) {
    // =========================================================== epilogue (256 threads)
    const int el = tid & 127, half = tid >> 7;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const float inv = (a.inv_scale ? __ldg(a.inv_scale) : 1.f) * (a.in_scale ? __ldg(a.in_scale + 1) : 1.f);
    const float gain = a.gain ? __ldg(a.gain) : 1.f;
    const bool vec_ok = (a.out_cstride & 3) == 0 && (a.out_coff & 3) == 0 && (reinterpret_cast<uintptr_t>(a.out) & 15) == 0;
    CVP_DECL();
    CvTileIt it;
    it.init(blockIdx.x, tiles_img);
    for (int k = 0; k < nmy; ++k) {
      int r0, c0;
      it.origin(g, r0, c0);
      const int b = it.b;
      it.advance(g, tiles_img);
      const int ai = acc2 ? (k & 1) : 0, an = acc2 ? (k >> 1) : k;       // accumulator set, its use count
      { CVP_T0(); mbar_wait(acc_full + ai, (uint32_t)(an & 1)); CVP_ADD(0); }
      tc_fence_after();
#pragma unroll 1
      for (int mt = 0; mt < 2; ++mt) {
        const int ir = r0 + (el >> 3), ic = c0 + 8 * mt + (el & 7);
        const bool valid = ir < a.H && ic < a.W;
        const size_t pix = valid ? (size_t)b * HW + (size_t)ir * a.W + ic : 0;
        const uint32_t trow = tmem_base + lane_base + (uint32_t)(acc2 ? ai * 256 + mt * 128 : mt * 256);
        if (a.lstm_R > 0) {
          // fused ConvLSTM cell (convLSTM.py:76-83): gates i,f,o,g at columns [0,R),[R,2R),[2R,3R),[3R,4R);
          // this thread owns recurrent channels [half*R/2, (half+1)*R/2) of its pixel
          const int R = a.lstm_R, Rh = R >> 1, rb = half * Rh;
          for (int q0 = 0; q0 < Rh; q0 += 16) {
            float gi[16], gf[16], go[16], gg[16], cp[16];
            tmem_ld16(trow + rb + q0, gi);
            tmem_ld16(trow + R + rb + q0, gf);
            tmem_ld16(trow + 2 * R + rb + q0, go);
            tmem_ld16(trow + 3 * R + rb + q0, gg);
            if (valid) {
              if (a.c_prev) {
                const float4* c4 = reinterpret_cast<const float4*>(a.c_prev + pix * R + rb + q0);
#pragma unroll
                for (int e = 0; e < 4; ++e) { const float4 t = __ldg(c4 + e); cp[4 * e] = t.x; cp[4 * e + 1] = t.y; cp[4 * e + 2] = t.z; cp[4 * e + 3] = t.w; }
              } else {
#pragma unroll
                for (int e = 0; e < 16; ++e) cp[e] = 0.f;
              }
              if (a.addend) {          // hoisted conditioning contribution to the four gates (same for every sample)
                const float* ad = a.addend + ((size_t)ir * a.W + ic) * a.addend_stride + rb + q0;
#pragma unroll
                for (int e4 = 0; e4 < 16; e4 += 4) {
                  const float4 ai = __ldg(reinterpret_cast<const float4*>(ad + e4)), af = __ldg(reinterpret_cast<const float4*>(ad + R + e4));
                  const float4 ao = __ldg(reinterpret_cast<const float4*>(ad + 2 * R + e4)), ag = __ldg(reinterpret_cast<const float4*>(ad + 3 * R + e4));
                  // accumulators carry the weight scale: add the (unscaled) term divided by inv, i.e. after the fmaf below
                  gi[e4] = fmaf(gi[e4], inv, ai.x); gi[e4 + 1] = fmaf(gi[e4 + 1], inv, ai.y); gi[e4 + 2] = fmaf(gi[e4 + 2], inv, ai.z); gi[e4 + 3] = fmaf(gi[e4 + 3], inv, ai.w);
                  gf[e4] = fmaf(gf[e4], inv, af.x); gf[e4 + 1] = fmaf(gf[e4 + 1], inv, af.y); gf[e4 + 2] = fmaf(gf[e4 + 2], inv, af.z); gf[e4 + 3] = fmaf(gf[e4 + 3], inv, af.w);
                  go[e4] = fmaf(go[e4], inv, ao.x); go[e4 + 1] = fmaf(go[e4 + 1], inv, ao.y); go[e4 + 2] = fmaf(go[e4 + 2], inv, ao.z); go[e4 + 3] = fmaf(go[e4 + 3], inv, ao.w);
                  gg[e4] = fmaf(gg[e4], inv, ag.x); gg[e4 + 1] = fmaf(gg[e4 + 1], inv, ag.y); gg[e4 + 2] = fmaf(gg[e4 + 2], inv, ag.z); gg[e4 + 3] = fmaf(gg[e4 + 3], inv, ag.w);
                }
              }
              const float inv_ = a.addend ? 1.f : inv;       // the scale was applied together with the addend
              float hn[16], cn[16];
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                const int r = rb + q0 + e;
                const float i_ = fast_sigm(fmaf(gi[e], inv_, s_bias[r]));
                const float f_ = fast_sigm(fmaf(gf[e], inv_, s_bias[R + r]));
                const float o_ = fast_sigm(fmaf(go[e], inv_, s_bias[2 * R + r]));
                const float g_ = fast_tanh(fmaf(gg[e], inv_, s_bias[3 * R + r]));
                cn[e] = fmaf(f_, cp[e], i_ * g_);
                hn[e] = o_ * fast_tanh(cn[e]);
              }
              float4* co = reinterpret_cast<float4*>(a.c_out + pix * R + rb + q0);
              float4* ho = reinterpret_cast<float4*>(a.h_out + pix * R + rb + q0);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                co[e] = make_float4(cn[4 * e], cn[4 * e + 1], cn[4 * e + 2], cn[4 * e + 3]);
                ho[e] = make_float4(hn[4 * e], hn[4 * e + 1], hn[4 * e + 2], hn[4 * e + 3]);
              }
            }
          }
        } else if (a.ndst > 0) {
          // data gradient: the columns are the input channels of the forward conv, each forward source padded to a
          // multiple of 4 columns; groups of 4 columns are routed to the tensor their channels came from (16-byte
          // accesses when the destination allows), gated by the ReLU of the forward input, optionally accumulated
          const int pb1 = (a.dst[0].nch + 3) & ~3, pb2 = pb1 + (a.ndst > 1 ? (a.dst[1].nch + 3) & ~3 : 0);
          const int pb3 = pb2 + (a.ndst > 2 ? (a.dst[2].nch + 3) & ~3 : 0);
          // 16-column chunks alternate between the two halves of the epilogue, continuing across the M tiles (odd chunk
          // counts stay balanced)
          for (int n0 = 0; n0 < NP; n0 += 16) {
            if (n0 >= pb3) break;
            if ((((n0 >> 4) + mt * (NP >> 4)) & 1) != half) continue;
            float v[16];
            tmem_ld16(trow + n0, v);
            if (valid) {
#pragma unroll
              for (int q = 0; q < 16; q += 4) {
                const int n = n0 + q;
                if (n >= pb3) break;
                const int d = n < pb1 ? 0 : (n < pb2 ? 1 : 2);
                const ConvDst& ds = a.dst[d];
                const int ch = n - (d == 0 ? 0 : (d == 1 ? pb1 : pb2));
                const int nv = min(4, ds.nch - ch);
                if (ds.p == nullptr || nv <= 0) continue;
                const size_t o = pix * ds.cstride + ds.coff + ch;
                float t[4] = {v[q] * inv, v[q + 1] * inv, v[q + 2] * inv, v[q + 3] * inv};
                const bool vec = nv == 4 && ((ds.cstride | ds.coff) & 3) == 0 && (reinterpret_cast<uintptr_t>(ds.p) & 15) == 0 &&
                                 (ds.mask == nullptr || (reinterpret_cast<uintptr_t>(ds.mask) & 15) == 0);
                if (vec) {
                  if (ds.mask) {
                    const float4 m4 = __ldg(reinterpret_cast<const float4*>(ds.mask + o));
                    if (!(m4.x > 0.f)) t[0] = 0.f;
                    if (!(m4.y > 0.f)) t[1] = 0.f;
                    if (!(m4.z > 0.f)) t[2] = 0.f;
                    if (!(m4.w > 0.f)) t[3] = 0.f;
                  }
                  float4* op = reinterpret_cast<float4*>(ds.p + o);
                  if (ds.accum) { const float4 p4 = *op; t[0] += p4.x; t[1] += p4.y; t[2] += p4.z; t[3] += p4.w; }
                  *op = make_float4(t[0], t[1], t[2], t[3]);
                } else {
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    if (e < nv) {
                      float x = t[e];
                      if (ds.mask && !(__ldg(ds.mask + o + e) > 0.f)) x = 0.f;
                      ds.p[o + e] = ds.accum ? ds.p[o + e] + x : x;
                    }
                  }
                }
              }
            }
          }
        } else {
          for (int n0 = 0; n0 < NP; n0 += 16) {
            if ((((n0 >> 4) + mt * (NP >> 4)) & 1) != half) continue;
            float v[16];
            tmem_ld16(trow + n0, v);
            if (valid) {
              float* op = a.out + pix * a.out_cstride + a.out_coff + n0;
              float ad[16];
              if (a.addend) {          // hoisted per-pixel term (rows padded to a multiple of 4 columns: 16-byte loads)
                const float* ap_ = a.addend + ((size_t)ir * a.W + ic) * a.addend_stride + n0;
#pragma unroll
                for (int e4 = 0; e4 < 16; e4 += 4) {
                  const float4 t4 = n0 + e4 < a.addend_stride ? __ldg(reinterpret_cast<const float4*>(ap_ + e4)) : make_float4(0.f, 0.f, 0.f, 0.f);
                  ad[e4] = t4.x; ad[e4 + 1] = t4.y; ad[e4 + 2] = t4.z; ad[e4 + 3] = t4.w;
                }
              } else {
#pragma unroll
                for (int e = 0; e < 16; ++e) ad[e] = 0.f;
              }
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                float t = fmaf(v[e], inv, s_bias[n0 + e]) + ad[e];
                if (a.gain) t *= gain;
                if (a.act == 1) t = fmaxf(t, 0.f);
                else if (a.act == 2) t = fminf(fmaxf(t, -2.f), kLog5);
                v[e] = t;
              }
              if (vec_ok) {      // 16-byte stores for every complete group of 4 columns, scalars only for a ragged tail
#pragma unroll
                for (int e = 0; e < 16; e += 4) {
                  if (n0 + e + 4 <= a.cout) *reinterpret_cast<float4*>(op + e) = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
                  else {
#pragma unroll
                    for (int e1 = e; e1 < e + 4; ++e1) if (n0 + e1 < a.cout) op[e1] = v[e1];
                  }
                }
              } else {
#pragma unroll
                for (int e = 0; e < 16; ++e) if (n0 + e < a.cout) op[e] = v[e];
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(acc_free + ai);
    }
#ifdef TMG_GT_PROFILE
    if (tid == 0) { cv_prof[blockIdx.x * 16 + 4] = clock64() - cvp_start; cv_prof[blockIdx.x * 16 + 5] = cvp_acc[0]; }
#endif
  } else if (warp < 10) {
    // =========================================================== MMA issue: warp 8 -> M tile 0, warp 9 -> M tile 1
    const int mt = warp - 8;
    if (elect_one()) {
      const uint32_t idesc = cv_idesc_f16(NP);
      const uint64_t hlA16 = g.hlA >> 4, hlB16 = g.hlB >> 4, tapB16 = (uint64_t)(g.nhl * g.hlB) >> 4;
      // ring slots / phase parities are carried incrementally and the tap loop is unrolled over the 9 taps: at narrow N one
      // MMA occupies the tensor pipe for ~24 cycles and the issue path (runtime div / mod, descriptor arithmetic) was longer
      int ua = 0, ub = 0; uint32_t pa = 0, pb = 0;
      CVP_DECL();
      for (int k = 0; k < nmy; ++k) {
        const int ai = acc2 ? (k & 1) : 0, an = acc2 ? (k >> 1) : k;
        const uint32_t tacc = tmem_base + (uint32_t)(acc2 ? ai * 256 + mt * 128 : mt * 256);
        { CVP_T0(); if (an >= 1) mbar_wait(acc_free + ai, (uint32_t)((an - 1) & 1)); CVP_ADD(0); }
        tc_fence_after();
        // One K-step = 9 taps, unrolled and branch-free for each of the three "taps per weight stage" layouts (the stage
        // boundaries are compile-time): a data-dependent branch per tap doubles the cycles per issued MMA (measured in
        // lstm_gate_f16.cu: 55 -> 120), and at narrow N the issue path, not the tensor pipe, is what is saturated.
        auto kstep = [&](auto tps_c, uint32_t acc0) {
          constexpr int TPS = decltype(tps_c)::value;
          { CVP_T0(); mbar_wait(a_full + ua, pa); CVP_ADD(1); }
          tc_fence_after();
          const uint64_t a0 = make_desc(smem_u32(As + (size_t)ua * g.bufA), kCvPLB, kCvRP * 16) + (uint64_t)(8 * mt);
          uint64_t bd = 0;
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            if (tap % TPS == 0) {
              { CVP_T0(); mbar_wait(b_full + ub, pb); CVP_ADD(2); }
              tc_fence_after();
              bd = make_desc(smem_u32(Bs + (size_t)ub * g.stageB), (uint32_t)NP * 16u, 128);
            }
            const int dr = tap / 3, dc = tap - 3 * dr;
            const uint64_t ad = a0 + (uint64_t)(dr * kCvRP + dc);
            cv_mma_f16(tacc, ad, bd, idesc, tap > 0 ? 1u : acc0);
            if (X3) {
              cv_mma_f16(tacc, ad + hlA16, bd, idesc, 1u);
              cv_mma_f16(tacc, ad, bd + hlB16, idesc, 1u);
            }
            bd += tapB16;
            if (tap % TPS == TPS - 1) {
              mma_commit(b_free + ub);
              if (++ub == g.nb) { ub = 0; pb ^= 1u; }
            }
          }
          mma_commit(a_free + ua);
          if (++ua == kCvNA) { ua = 0; pa ^= 1u; }
        };
        if (g.tps == 9) { for (int ks = 0; ks < g.KS; ++ks) kstep(std::integral_constant<int, 9>{}, ks > 0 ? 1u : 0u); }
        else if (g.tps == 3) { for (int ks = 0; ks < g.KS; ++ks) kstep(std::integral_constant<int, 3>{}, ks > 0 ? 1u : 0u); }
        else { for (int ks = 0; ks < g.KS; ++ks) kstep(std::integral_constant<int, 1>{}, ks > 0 ? 1u : 0u); }
        mma_commit(acc_full + ai);
      }
#ifdef TMG_GT_PROFILE
      if (mt == 0) { cv_prof[blockIdx.x * 16 + 0] = clock64() - cvp_start; for (int i = 0; i < 3; ++i) cv_prof[blockIdx.x * 16 + 1 + i] = cvp_acc[i]; }
#endif
    }
  } else if (warp == 10) {
    // =========================================================== weight streaming (one lane, cp.async.bulk)
    if (lane == 0) {
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(a.wpk);
      int jb = 0;
      const uint32_t tapB = g.nhl * g.hlB;
      for (int k = 0; k < nmy; ++k) {
        for (int st = 0; st < nstage_tile; st += g.tps, ++jb) {
          const int ub = jb % g.nb, use = jb / g.nb;
          if (use >= 1) mbar_wait(b_free + ub, (uint32_t)((use - 1) & 1));
          mbar_expect_tx(b_full + ub, g.stageB);
          uint8_t* dst = Bs + (size_t)ub * g.stageB;
          if (X3) {          // hi and lo of consecutive taps are contiguous in the packed layout
            bulk_g2s(dst, wsrc + (size_t)st * g.gstageB, g.stageB, b_full + ub);
          } else {
            for (int t = 0; t < g.tps; ++t) bulk_g2s(dst + (size_t)t * tapB, wsrc + (size_t)(st + t) * g.gstageB, tapB, b_full + ub);
          }
        }
      }
    }
  } else {
    // =========================================================== producers (128 threads): activation K-steps
    const int ptid = (tid - 11 * 32) & 127, pgrp = (tid - 11 * 32) >> 7;
    CvTileIt it;
    it.init(blockIdx.x, tiles_img);
    int ja = 0;
    CVP_DECL();
    const float isc = a.in_scale ? __ldg(a.in_scale) : 1.f;
    for (int k = 0; k < nmy; ++k) {
      int r0, c0;
      it.origin(g, r0, c0);
      const int b = it.b;
      it.advance(g, tiles_img);
      for (int ks = 0; ks < g.KS; ++ks, ++ja) {
        if (kNPG > 1 && (ja % kNPG) != pgrp) continue;
        const int ua = ja % kCvNA, use = ja / kCvNA;
        // item = (tile position, 16-byte quarter q of the K-step's 16 channels): 4 consecutive lanes read 64 contiguous bytes
        // of a pixel (8 pixels per request instead of 32 lines).  q = ptid & 3 is fixed per thread, so the source and channel
        // offset are chosen once per K-step, and ALL positions of the thread are requested before the first conversion: one
        // DRAM round trip per K-step (lstm_gate_f16.cu measured 10k -> 4k cycles per K-step with the same change).
        const int q = ptid & 3, plane = 2 * ks + (q >> 1);
        int si = 0;
        if (a.nsrc > 1 && plane >= g.plane0[1]) si = 1;
        if (a.nsrc > 2 && plane >= g.plane0[2]) si = 2;
        // (field-wise select: a dynamically indexed kernel parameter is copied to local memory)
        const float* sp = si == 0 ? a.src[0].p : (si == 1 ? a.src[1].p : a.src[2].p);
        const int s_cstride = si == 0 ? a.src[0].cstride : (si == 1 ? a.src[1].cstride : a.src[2].cstride);
        const int s_coff = si == 0 ? a.src[0].coff : (si == 1 ? a.src[1].coff : a.src[2].coff);
        const int s_nch = si == 0 ? a.src[0].nch : (si == 1 ? a.src[1].nch : a.src[2].nch);
        const bool s_relu = (si == 0 ? a.src[0].relu : (si == 1 ? a.src[1].relu : a.src[2].relu)) != 0;
        const bool s_shared = (si == 0 ? a.src[0].bshared : (si == 1 ? a.src[1].bshared : a.src[2].bshared)) != 0;
        const int ch = (plane - (si == 0 ? 0 : (si == 1 ? g.plane0[1] : g.plane0[2]))) * 8 + (q & 1) * 4;
        const int nv = min(4, s_nch - ch);
        const float* sbase = sp ? sp + (s_shared ? 0 : (size_t)b * HW) * s_cstride + s_coff + ch : nullptr;
        const bool live = sbase != nullptr && nv > 0;
        const bool vec4 = nv == 4 && (s_cstride & 3) == 0 && (reinterpret_cast<uintptr_t>(sbase) & 15) == 0;
        const bool vec2 = nv == 2 && (s_cstride & 1) == 0 && (reinterpret_cast<uintptr_t>(sbase) & 7) == 0;
        constexpr int kItems = (kCvNPOS + kNPT / 4 - 1) / (kNPT / 4);       // positions per thread: 324 / 32 -> 11
        float4 v[kItems];
#pragma unroll
        for (int u = 0; u < kItems; ++u) {
          const int p = (ptid >> 2) + u * (kNPT / 4);
          v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          const int rr = p / kCvRP, rc = p - rr * kCvRP;
          int r = r0 - 1 + rr, c = c0 - 1 + rc;
          bool inb = r >= 0 && r < a.H && c >= 0 && c < a.W;
          if (a.pad_replicate) { r = min(max(r, 0), a.H - 1); c = min(max(c, 0), a.W - 1); inb = true; }
          if (live && p < kCvNPOS && inb) {
            const float* ptr = sbase + ((size_t)r * a.W + c) * s_cstride;
            if (vec4) {
              v[u] = __ldg(reinterpret_cast<const float4*>(ptr));
            } else if (vec2) {
              const float2 t = __ldg(reinterpret_cast<const float2*>(ptr)); v[u].x = t.x; v[u].y = t.y;
            } else {
              v[u].x = __ldg(ptr);
              if (nv > 1) v[u].y = __ldg(ptr + 1);
              if (nv > 2) v[u].z = __ldg(ptr + 2);
              if (nv > 3) v[u].w = __ldg(ptr + 3);
            }
          }
        }
        { CVP_T0(); if (use >= 1) mbar_wait(a_free + ua, (uint32_t)((use - 1) & 1)); CVP_ADD(0); }
        uint8_t* Ab = As + (size_t)ua * g.bufA + (q >> 1) * kCvPLB + (q & 1) * 8;
        const float lo_ = s_relu ? 0.f : -60000.f;
#pragma unroll
        for (int u = 0; u < kItems; ++u) {
          const int p = (ptid >> 2) + u * (kNPT / 4);
          if (p >= kCvNPOS) continue;
          const float y0 = fmaxf(fminf(v[u].x * isc, 60000.f), lo_), y1 = fmaxf(fminf(v[u].y * isc, 60000.f), lo_);
          const float y2 = fmaxf(fminf(v[u].z * isc, 60000.f), lo_), y3 = fmaxf(fminf(v[u].w * isc, 60000.f), lo_);
          const __half2 h01 = __floats2half2_rn(y0, y1), h23 = __floats2half2_rn(y2, y3);
          uint8_t* dst = Ab + p * 16;
          *reinterpret_cast<uint2*>(dst) = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
          if (X3) {
            const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
            const __half2 l01 = __floats2half2_rn(y0 - f01.x, y1 - f01.y), l23 = __floats2half2_rn(y2 - f23.x, y3 - f23.y);
            *reinterpret_cast<uint2*>(dst + g.hlA) = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
          }
        }
        fence_proxy_async();
        mbar_arrive(a_full + ua);
      }
    }
#ifdef TMG_GT_PROFILE
    if (ptid == 0 && pgrp == 0) { cv_prof[blockIdx.x * 16 + 6] = clock64() - cvp_start; cv_prof[blockIdx.x * 16 + 7] = cvp_acc[0]; }
#endif
    // (no overflow detection here: the staging loop of this kernel is its bottleneck at N <= 64 -- measured +0.6 ms per
    // call for one compare per 8 values; the only unbounded input of these convolutions is the flow state, which the step
    // kernels that produce and consume it check: flow_step_f16.cu, flow_level_f16.cu)
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ------------------------------------------------------------------ power-of-two scale of a gradient tensor
__global__ void __launch_bounds__(256)
absmax_part_kernel(const float* __restrict__ g, int64_t npix, int cstride, int coff, int nch, float* __restrict__ part) {
  float m = 0.f;
  const int64_t total = npix * nch;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int64_t p = i / nch; const int c = (int)(i - p * nch);
    m = fmaxf(m, fabsf(__ldg(g + p * cstride + coff + c)));
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  __shared__ float s[8];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) m = fmaxf(m, s[w]);
    part[blockIdx.x] = m;
  }
}
__global__ void absmax_final_kernel(const float* __restrict__ part, int n, float* __restrict__ scale) {
  float m = 0.f;
  for (int i = threadIdx.x; i < n; i += 32) m = fmaxf(m, part[i]);
#pragma unroll
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (threadIdx.x == 0) {
    int ex = 0;
    float sc = 1.f;
    if (m > 0.f && m < 3.0e38f) { frexpf(m, &ex); sc = ldexpf(1.f, min(max(11 - ex, -100), 100)); }
    scale[0] = sc; scale[1] = 1.f / sc;
  }
}
// single-launch form: block maxima meet in an atomicMax (max is order-independent, so still deterministic); the last block
// to arrive derives the scale and resets the two words of `sync` (zero-initialised, library-owned) for the next use
__global__ void __launch_bounds__(256)
absmax_scale_kernel(const float* __restrict__ g, int64_t npix, int cstride, int coff, int nch, float* __restrict__ scale,
                    unsigned* __restrict__ sync) {
  float m = 0.f;
  const int64_t total = npix * nch;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int64_t p = i / nch; const int c = (int)(i - p * nch);
    m = fmaxf(m, fabsf(__ldg(g + p * cstride + coff + c)));
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  __shared__ float s[8];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) m = fmaxf(m, s[w]);
    if (!(m == m) || m > 3.0e38f) m = 3.0e38f;                       // NaN / inf gradients: keep the bit pattern ordered
    atomicMax(sync, __float_as_uint(m));                             // non-negative floats order like their bit patterns
    __threadfence();
    if (atomicAdd(sync + 1, 1u) == gridDim.x - 1) {
      __threadfence();
      const float mx = __uint_as_float(atomicExch(sync, 0u));
      sync[1] = 0u;
      int ex = 0;
      float sc = 1.f;
      if (mx > 0.f && mx < 3.0e38f) { frexpf(mx, &ex); sc = ldexpf(1.f, min(max(11 - ex, -100), 100)); }
      scale[0] = sc; scale[1] = 1.f / sc;
    }
  }
}
// absmax scale AND column sums of g (the bias gradient of the convolution g belongs to) in one launch: threads are laid
// out by column; per-block column sums go to part[block][nch], and the last block to arrive adds them up in block order
// (fixed order: deterministic) into gbias.
__global__ void __launch_bounds__(256)
absmax_colsum_kernel(const float* __restrict__ g, int64_t npix, int cstride, int coff, int nch, int cw, float* __restrict__ scale,
                     unsigned* __restrict__ sync, float* __restrict__ part, float* __restrict__ gbias, int accum) {
  __shared__ float s_sum[256], s_max[256];
  __shared__ unsigned s_last;
  const int c = threadIdx.x % cw, rowi = threadIdx.x / cw, rows = 256 / cw;
  float m = 0.f, acc0 = 0.f, acc1 = 0.f;
  if (c < nch) {
    // two independent chains per thread: the loads of consecutive iterations overlap
    const int64_t step = (int64_t)gridDim.x * rows;
    int64_t p = (int64_t)blockIdx.x * rows + rowi;
    for (; p + step < npix; p += 2 * step) {
      const float v0 = __ldg(g + p * cstride + coff + c), v1 = __ldg(g + (p + step) * cstride + coff + c);
      acc0 += v0; acc1 += v1; m = fmaxf(m, fmaxf(fabsf(v0), fabsf(v1)));
    }
    if (p < npix) { const float v = __ldg(g + p * cstride + coff + c); acc0 += v; m = fmaxf(m, fabsf(v)); }
  }
  float acc = acc0 + acc1;
  s_sum[threadIdx.x] = acc; s_max[threadIdx.x] = m;
  __syncthreads();
  if (rowi == 0 && c < nch) {
    for (int r = 1; r < rows; ++r) acc += s_sum[r * cw + c];
    part[(size_t)blockIdx.x * nch + c] = acc;
  }
  if (threadIdx.x == 0) {
    for (int i = 1; i < 256; ++i) m = fmaxf(m, s_max[i]);
    if (!(m == m) || m > 3.0e38f) m = 3.0e38f;
    atomicMax(sync, __float_as_uint(m));
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(sync + 1, 1u) == gridDim.x - 1 ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // last block: every column is summed over the blocks by `rows` threads (block b = rowi, rowi + rows, ...), then the
  // row partials in order -- a fixed order, whichever block happens to be last
  float sum = 0.f;
  if (c < nch) for (unsigned b = rowi; b < gridDim.x; b += rows) sum += __ldcg(part + (size_t)b * nch + c);
  s_sum[threadIdx.x] = sum;
  __syncthreads();
  if (rowi == 0 && c < nch) {
    for (int r = 1; r < rows; ++r) sum += s_sum[r * cw + c];
    gbias[c] = accum ? gbias[c] + sum : sum;
  }
  if (threadIdx.x == 0) {
    const float mx = __uint_as_float(atomicExch(sync, 0u));
    sync[1] = 0u;
    int ex = 0;
    float sc = 1.f;
    if (mx > 0.f && mx < 3.0e38f) { frexpf(mx, &ex); sc = ldexpf(1.f, min(max(11 - ex, -100), 100)); }
    scale[0] = sc; scale[1] = 1.f / sc;
  }
}
// part: >= 444 * nch floats
int launch_absmax_colsum(const float* g, int64_t npix, int cstride, int coff, int nch, float* scale, unsigned* sync, float* part,
                         float* gbias, int accum, cudaStream_t st) {
  if (nch > 256) { set_error("absmax_colsum: %d channels", nch); return TMG_ERR_UNSUPPORTED; }
  int cw = 1;
  while (cw < nch) cw <<= 1;
  const int nb = (int)std::max<int64_t>(1, std::min<int64_t>(444, npix / 64));
  absmax_colsum_kernel<<<nb, 256, 0, st>>>(g, npix, cstride, coff, nch, cw, scale, sync, part, gbias, accum);
  TMG_LAUNCH_CHECK();
  return TMG_OK;
}
int launch_absmax_scale(const float* g, int64_t npix, int cstride, int coff, int nch, float* scale, float* scratch, cudaStream_t st,
                        unsigned* sync) {
  const int64_t total = npix * nch;
  const int nb = (int)std::min<int64_t>(592, std::max<int64_t>(1, (total + 2047) / 2048));
  if (sync) {
    absmax_scale_kernel<<<nb, 256, 0, st>>>(g, npix, cstride, coff, nch, scale, sync);
    TMG_LAUNCH_CHECK();
    return TMG_OK;
  }
  absmax_part_kernel<<<nb, 256, 0, st>>>(g, npix, cstride, coff, nch, scratch);
  TMG_LAUNCH_CHECK();
  absmax_final_kernel<<<1, 32, 0, st>>>(scratch, nb, scale);
  TMG_LAUNCH_CHECK();
  return TMG_OK;
}

// ------------------------------------------------------------------ host side
int convf16_ksteps(const int* nch, int nsrc) {
  int planes = 0;
  for (int i = 0; i < nsrc; ++i) planes += (nch[i] + 7) / 8;
  return (planes + 1) / 2;
}
size_t convf16_packed_floats(const int* nch, int nsrc, int npad) {     // [ks][tap][hl][2 planes][npad][8 halves]
  return (size_t)convf16_ksteps(nch, nsrc) * 9 * 2 * 2 * npad * 16 / 4;
}

static int cv_sm_count() {
  int dev = 0, nsm = 148;
  cudaGetDevice(&dev);
  static int cached[64] = {0};
  if (dev >= 0 && dev < 64) {
    if (!cached[dev]) cudaDeviceGetAttribute(&cached[dev], cudaDevAttrMultiProcessorCount, dev);
    if (cached[dev] > 0) nsm = cached[dev];
  }
  return nsm;
}

static bool cv_geom(const ConvF16Args& a, ConvF16Geom& g, int grid) {
  if (a.npad % 16 || a.npad < 16 || a.npad > 256 || a.nsrc < 1 || a.nsrc > 3) return false;
  int pl = 0;
  for (int i = 0; i < 3; ++i) {
    g.plane0[i] = pl;
    g.nplanes[i] = i < a.nsrc ? (a.src[i].nch + 7) / 8 : 0;
    pl += g.nplanes[i];
  }
  g.KS = (pl + 1) / 2;
  g.nhl = a.x3 ? 2 : 1;
  g.hlA = 2 * kCvPLB; g.bufA = g.hlA * g.nhl;
  g.hlB = (uint32_t)2 * a.npad * 16; g.gstageB = 2 * g.hlB;
  // a weight stage holds 9, 3 or 1 taps: narrow-N convs have too little MMA work per tap to hide the barrier round
  // trip of a per-tap stage
  g.tps = g.hlB * g.nhl * 9 <= 32 * 1024 ? 9 : (g.hlB * g.nhl * 3 <= 32 * 1024 ? 3 : 1);
  g.stageB = g.hlB * g.nhl * g.tps;
  g.tiles_x = cdiv(a.W, 16); g.tiles_y = cdiv(a.H, 16);
  const int tiles_img = g.tiles_x * g.tiles_y;
  if (tiles_img >= 4096) return false;
  g.ntiles = tiles_img * a.B;
  g.step_b = grid / tiles_img; g.step_t = grid % tiles_img;
  g.inv_tx = (uint32_t)((65536 + g.tiles_x - 1) / g.tiles_x);
  uint32_t off = 0;
  auto take = [&](uint32_t n) { uint32_t o = off; off += (n + 127) / 128 * 128; return o; };
  g.oA = take((uint32_t)kCvNA * g.bufA);
  g.oMisc = take((uint32_t)a.npad * 4);
  g.oBar = take(27 * 8 + 16);
  const uint32_t fixed = off;
  int nb = (int)((220u * 1024u - fixed) / ((g.stageB + 127) / 128 * 128));
  nb = std::min(nb, g.tps == 1 ? kCvMaxNB : 4);
  if (nb < 2) return false;
  g.nb = nb;
  g.oB = take((uint32_t)nb * g.stageB);
  g.total = off;
  return g.total <= 227 * 1024;
}

bool convf16_supported(const ConvF16Args& a) {
  ConvF16Geom g{};
  return cv_geom(a, g, 148) && (a.lstm_R == 0 || (a.lstm_R % 32 == 0 && 4 * a.lstm_R == a.npad));
}

int launch_conv3x3_f16(const ConvF16Args& a, cudaStream_t st) {
  if (a.B <= 0 || a.H <= 0 || a.W <= 0) return TMG_OK;
  const int tiles = cdiv(a.W, 16) * cdiv(a.H, 16) * a.B;
  const int grid = std::min(tiles, cv_sm_count());
  ConvF16Geom g{};
  if (!cv_geom(a, g, grid) || (a.lstm_R != 0 && (a.lstm_R % 32 != 0 || 4 * a.lstm_R != a.npad))) {
    set_error("fp16 conv: unsupported shape (N=%d, %dx%d, %d sources)", a.npad, a.H, a.W, a.nsrc);
    return TMG_ERR_UNSUPPORTED;
  }
#define TMG_CV(XX, PW)                                                                                                      \
  {                                                                                                                         \
    TMG_SMEM_ATTR(conv3x3_f16_kernel<XX, PW>, 227 * 1024); \
    conv3x3_f16_kernel<XX, PW><<<grid, (11 + PW) * 32, g.total, st>>>(a, g);                                                \
  }
  // 4 producer warps for every shape.  Narrow N used 8 (two groups of 4 staging alternate K-steps) while a K-step still cost
  // several DRAM round trips; with one round trip per K-step the second group only takes issue slots from the MMA issuers
  // and registers from everybody (96 instead of 128 per thread): output conv 3.38 -> 3.23 ms with 4 (TMG_CV_PW8=1: A/B runs).
  static const bool pw8 = [] { const char* e = getenv("TMG_CV_PW8"); return e && e[0] == '1'; }();
  const bool wide = !pw8 || a.npad > 96 || a.lstm_R > 0;
  if (a.x3) { if (wide) TMG_CV(true, 4) else TMG_CV(true, 8) }
  else { if (wide) TMG_CV(false, 4) else TMG_CV(false, 8) }
#undef TMG_CV
  TMG_LAUNCH_CHECK();
#ifdef TMG_GT_PROFILE
  if (getenv("TMG_CV_PROF") && tiles >= 8 * grid) {
    static int left = atoi(getenv("TMG_CV_PROF"));
    if (left > 0) {
      --left;
      cudaStreamSynchronize(st);
      static long long h[148 * 16];
      cudaMemcpyFromSymbol(h, cv_prof, sizeof(h));
      double s8[8] = {0};
      for (int b = 0; b < grid; ++b) for (int i = 0; i < 8; ++i) s8[i] += (double)h[b * 16 + i] / grid;
      const double nt = (double)tiles / grid;
      fprintf(stderr, "conv_f16 %dx%d B=%d N=%d KS=%d nb=%d tps=%d lstm=%d: cycles/tile issuer %.0f (acc_free %.0f a_full %.0f b_full %.0f) | epilogue %.0f (acc_full %.0f) | producer %.0f (a_free %.0f)\n",
              a.H, a.W, a.B, a.npad, g.KS, g.nb, g.tps, a.lstm_R, s8[0] / nt, s8[1] / nt, s8[2] / nt, s8[3] / nt, s8[4] / nt, s8[5] / nt, s8[6] / nt, s8[7] / nt);
    }
  }
#endif
  return TMG_OK;
}

}  // namespace tmg
