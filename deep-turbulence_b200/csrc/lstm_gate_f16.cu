// ConvLSTM gate convolution + cell update (nn/modules/convLSTM.py:44,74-83) as a TWO-PASS tcgen05 implicit GEMM (sm_100a).
//
// conv3x3_f16.cu computes the four gates of a 16x16-pixel tile as two M tiles x N = 256 columns = all 512 TMEM columns, so
// the MMAs of the next tile wait for the cell-update epilogue of this one (ncu, profiles/r02_gate_*: tensor pipe 46 % busy,
// the rest is the epilogue).  Here the N dimension is split into two passes of 128 columns -- the gate rows are permuted at
// pack time so that pass p holds i, f, o, g of recurrent channels [32 p, 32 p + 32) -- and each pass owns half of TMEM: the
// epilogue of pass p runs while the tensor pipe works on pass 1 - p.  The weight stream is unchanged (every packed byte is
// still fetched once per tile); the staged activation K-steps of a tile stay resident until the second pass has read them
// (ring of KS + 1 .. KS + 2 buffers).
//
// The epilogue was bound by L1 wavefronts, not by math: one pixel per TMEM lane and pixel-major (NHWC) states put the 32
// lanes of every 16-byte access into 32 different 128-byte lines (34 wavefronts per request measured).  Now
//   * c_prev / c_out / h_out move between global and shared memory with 4 lanes per pixel (64 contiguous bytes, 8 lines per
//     request) and are transposed to the one-pixel-per-lane view through an XOR-swizzled, conflict-free staging buffer;
//   * the hoisted conditioning term (+ bias) is read from a plane-transposed table [column group of 4][pixel] float4:
//     8 neighbouring pixels of a tile row are one 128-byte line;
//   * the activation staging uses the same 4-lanes-per-pixel mapping (the h planes come first so that the two planes of a
//     K-step are 64 contiguous, aligned bytes).
#include <cuda_fp16.h>

#include <type_traits>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace tmg {

namespace {

constexpr int kGtRP = 18;                  // staged tile pitch: 16 + 2
constexpr int kGtNPOS = 324;
constexpr int kGtNPOSA = 332;              // plane pitch 5312 B = 16 banks mod 32: the two planes of a K-step do not collide
constexpr uint32_t kGtPLB = kGtNPOSA * 16;
constexpr int kGtMaxNA = 10, kGtMaxNB = 12;
constexpr int kGtNP = 128;                 // MMA N of one pass: 4 gates x 32 recurrent channels
constexpr int kGtPW = 4;                   // producer warps
constexpr int kGtThreads = (11 + kGtPW) * 32;
constexpr int kGtNPT = kGtPW * 32;
constexpr uint32_t kGtStageW = 4096;       // epilogue staging per warp: c block + h block, 32 pixels x 64 B each

struct GateGeom {
  int KS, NA, NB, nhl;
  int tiles_x, tiles_y, ntiles, step_b, step_t;
  uint32_t inv_tx;
  uint32_t hlA, bufA;      // bytes: hi -> lo distance inside an A K-step buffer, buffer size
  uint32_t stageB, hlB;    // bytes of one weight stage (one tap of one pass: hi [+ lo]); hi -> lo distance
  uint32_t oA, oB, oStage, oMisc, oBar, total;
  int plane0[3];
  int nplanes[3];
  int odd;                 // X3 and an odd number of K planes: the last K-step carries [x_hi | x_lo] / [x_hi | 0] (2 MMAs, see the issuer)
};

struct GtTileIt {
  int b, timg;
  __device__ __forceinline__ void init(int t, int tiles_img) { b = t / tiles_img; timg = t - b * tiles_img; }
  __device__ __forceinline__ void advance(const GateGeom& g, int tiles_img) {
    timg += g.step_t; b += g.step_b;
    if (timg >= tiles_img) { timg -= tiles_img; ++b; }
  }
  __device__ __forceinline__ void origin(const GateGeom& g, int& r0, int& c0) const {
    const int ty = (int)(((uint32_t)timg * g.inv_tx) >> 16);
    r0 = ty * 16; c0 = (timg - ty * g.tiles_x) * 16;
  }
};

__device__ __forceinline__ float exp2f_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// staging index of (pixel of the warp, 16-byte chunk): conflict-free for "4 lanes per pixel" and for "one pixel per lane"
__device__ __forceinline__ int gt_sw(int pw, int ch4) { return pw * 4 + (ch4 ^ ((pw >> 1) & 3)); }

// ConvLSTM cell (convLSTM.py:76-83) with 7 instead of 10 special-function operations: the sigmoid / tanh quotients of one
// update share a denominator,
//   c' = sigm(f) c + sigm(i) tanh(g) = (c A B + (e^2g - 1) D) / (D A B),  A = 1 + e^-i, D = 1 + e^-f, B = e^2g + 1
//   h' = sigm(o) tanh(c')            = (e^2c' - 1) / ((1 + e^-o)(e^2c' + 1))
// Pre-activations are clamped to +-28 (sigmoid and tanh are saturated to fp32 precision far before) so that the products
// stay finite: (1 + e^28)^3 = 3e36.  ex2.approx / rcp.approx: 2 ulp each, abs error ~3e-7.
__device__ __forceinline__ void gt_cell(float pi, float pf, float po, float pg, float cp, float& cn, float& hn) {
  constexpr float kL2E = 1.4426950408889634f;
  pi = fminf(fmaxf(pi, -28.f), 28.f); pf = fminf(fmaxf(pf, -28.f), 28.f); po = fminf(fmaxf(po, -28.f), 28.f);
  pg = fminf(fmaxf(pg, -14.f), 14.f);
  const float ei = exp2f_approx(-kL2E * pi), ef = exp2f_approx(-kL2E * pf), eg = exp2f_approx(2.f * kL2E * pg);
  const float A = 1.f + ei, D = 1.f + ef, Bq = eg + 1.f, N = eg - 1.f;
  const float AB = A * Bq;
  cn = __fdividef(fmaf(cp, AB, N * D), D * AB);
  const float cc = fminf(fmaxf(cn, -14.f), 14.f);
  const float eo = exp2f_approx(-kL2E * po), ec = exp2f_approx(2.f * kL2E * cc);
  hn = __fdividef(ec - 1.f, (1.f + eo) * (ec + 1.f));
}

#ifdef TMG_GT_PROFILE
// developer build (-DTMG_GT_PROFILE): cycles the M-tile-0 issuer / one epilogue warp / one producer warp spend in each wait
__device__ long long gt_prof[148 * 16];
#define GTP_T0() const long long gtp_t0 = clock64()
#define GTP_ADD(slot) gtp_acc[slot] += clock64() - gtp_t0
#else
#define GTP_T0()
#define GTP_ADD(slot)
#endif

}  // namespace

template <bool X3>
__global__ void __launch_bounds__(kGtThreads, 1)
lstm_gate_f16_kernel(ConvF16Args a, GateGeom g) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int HW = a.H * a.W;
  const int tiles_img = g.tiles_x * g.tiles_y;
  constexpr int R = 64, RH = 32;

  uint8_t* As = smem + g.oA;                 // NA K-step buffers [hl][2 planes][NPOSA][16 B]
  uint8_t* Bs = smem + g.oB;                 // NB weight stages [hl][2 planes][128][16 B]
  float* s_bias = reinterpret_cast<float*>(smem + g.oMisc);      // [2 passes][128], pass order
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + g.oBar);
  uint64_t* a_full = bars;                   // [10] producers (128)
  uint64_t* a_free = bars + 10;              // [10] 2 commits (second pass)
  uint64_t* b_full = bars + 20;              // [12] tx
  uint64_t* b_free = bars + 32;              // [12] 2 commits
  uint64_t* acc_full = bars + 44;            // [2]  2 commits
  uint64_t* acc_free = bars + 46;            // [2]  epilogue (256)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 48);

  if (tid == 0) {
    for (int i = 0; i < kGtMaxNA; ++i) { mbar_init(a_full + i, kGtNPT); mbar_init(a_free + i, 2); }
    for (int i = 0; i < kGtMaxNB; ++i) { mbar_init(b_full + i, 1); mbar_init(b_free + i, 2); }
    for (int i = 0; i < 2; ++i) { mbar_init(acc_full + i, 2); mbar_init(acc_free + i, 256); }
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc(tmem_slot, 512);
  for (int i = tid; i < 2 * kGtNP; i += kGtThreads) {
    // column n of pass p is gate row (n / 32) * R + 32 p + n % 32
    const int p = i >> 7, n = i & 127;
    s_bias[i] = a.bias ? __ldg(a.bias + (n >> 5) * R + p * RH + (n & 31)) : 0.f;
  }
  {   // zero the A ring once (positions >= 324 of every plane stay zero)
    uint4* z4 = reinterpret_cast<uint4*>(As);
    const int n4 = (int)((size_t)g.NA * g.bufA / 16);
    for (int i = tid; i < n4; i += kGtThreads) z4[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int nmy = ((int)blockIdx.x < g.ntiles) ? (g.ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp < 8) {
    // =========================================================== epilogue (256 threads): ConvLSTM cell, convLSTM.py:76-83
    const int el = tid & 127, half = tid >> 7, wq = warp & 3;
    const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
    const float inv = a.inv_scale ? __ldg(a.inv_scale) : 1.f;
    float4* stC = reinterpret_cast<float4*>(smem + g.oStage + (size_t)warp * kGtStageW);
    float4* stH = stC + 128;
    const float4* adT = reinterpret_cast<const float4*>(a.addend);
    const int sw = (lane >> 1) & 3;
#ifdef TMG_GT_PROFILE
    long long gtp_acc[4] = {0, 0, 0, 0};
    const long long gtp_start = clock64();
#endif
    // One step = (tile, pass p, M tile mt): 32 pixels x 16 recurrent channels per warp.  The loop is software-pipelined: the
    // c_prev block and the first half of the addend rows of step s + 1 are requested during the cell math of step s (the
    // states do not fit L2: an un-prefetched step waits a full DRAM round trip before its first instruction of math).
    GtTileIt it;
    it.init(blockIdx.x, tiles_img);
    int r0 = 0, c0 = 0, b = 0;
    if (nmy > 0) { it.origin(g, r0, c0); b = it.b; it.advance(g, tiles_img); }
    // per-lane pieces of the cooperative (4 lanes per pixel) view: pixel pw = 8 j + lane / 4 of the warp's 32
    const int cl_row = (wq * 32 + (lane >> 2)) >> 3, cl_col = (lane >> 2) & 7;       // rows advance by 1 per j
    auto coop_off = [&](int r0_, int c0_, int b_, int p_, int mt_, int j, bool& ok) -> size_t {
      const int irw = r0_ + cl_row + j, icw = c0_ + 8 * mt_ + cl_col;
      ok = irw < a.H && icw < a.W;
      return ((size_t)b_ * HW + (size_t)(ok ? irw * a.W + icw : 0)) * R + p_ * RH + half * 16 + (lane & 3) * 4;
    };
    auto own_pix = [&](int r0_, int c0_, int mt_) -> int {
      const int ir = r0_ + (el >> 3), ic = c0_ + 8 * mt_ + (el & 7);
      return (ir < a.H && ic < a.W) ? ir * a.W + ic : 0;
    };
    float4 ad[8], tc[4];
    if (nmy > 0) {        // prologue: requests of step 0
      if (adT) {
        const float4* ap = adT + (size_t)(half * 4) * HW + own_pix(r0, c0, 0);
#pragma unroll
        for (int g4 = 0; g4 < 4; ++g4) { ad[2 * g4] = __ldg(ap + (size_t)(8 * g4) * HW); ad[2 * g4 + 1] = __ldg(ap + (size_t)(8 * g4 + 1) * HW); }
      }
      if (a.c_prev) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { bool ok; const size_t o = coop_off(r0, c0, b, 0, 0, j, ok); tc[j] = ok ? __ldg(reinterpret_cast<const float4*>(a.c_prev + o)) : make_float4(0.f, 0.f, 0.f, 0.f); }
      }
    }
    const int nsteps = 4 * nmy;
#pragma unroll 1
    for (int sidx = 0; sidx < nsteps; ++sidx) {
      const int k = sidx >> 2, p = (sidx >> 1) & 1, mt = sidx & 1;
      if (mt == 0) {
        { GTP_T0(); mbar_wait(acc_full + p, (uint32_t)(k & 1)); GTP_ADD(0); }
        tc_fence_after();
      }
      const float4* ap = adT ? adT + (size_t)(p * 32 + half * 4) * HW + own_pix(r0, c0, mt) : nullptr;
      if (a.c_prev) {
#pragma unroll
        for (int j = 0; j < 4; ++j) stC[gt_sw(8 * j + (lane >> 2), lane & 3)] = tc[j];
      }
      __syncwarp();
      const uint32_t tcol = tmem_base + lane_base + (uint32_t)((p * 2 + mt) * kGtNP + half * 16);
      // the step after this one (same tile, or the first step of the CTA's next tile)
      int r0n = r0, c0n = c0, bn = b;
      if ((sidx & 3) == 3 && sidx + 1 < nsteps) { it.origin(g, r0n, c0n); bn = it.b; it.advance(g, tiles_img); }
      const int pn = ((sidx + 1) >> 1) & 1, mtn = (sidx + 1) & 1;
#pragma unroll
      for (int sub = 0; sub < 2; ++sub) {
        // 8 recurrent channels: gates i, f, o, g at columns +0, +32, +64, +96 of the pass
        uint32_t gr[4][8];
#pragma unroll
        for (int g4 = 0; g4 < 4; ++g4) tmem_ld_nowait<8>(tcol + (uint32_t)(32 * g4 + 8 * sub), gr[g4]);
        tmem_ld_wait();
        if (mt == 1 && sub == 1) {     // both M tiles of this pass are in registers: the tensor pipe may refill the accumulators
          tc_fence_before();
          mbar_arrive(acc_free + p);
        }
        float pre[4][8];
        if (adT) {
#pragma unroll
          for (int g4 = 0; g4 < 4; ++g4) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const float4 t = ad[2 * g4 + e];
              pre[g4][4 * e] = fmaf(__uint_as_float(gr[g4][4 * e]), inv, t.x); pre[g4][4 * e + 1] = fmaf(__uint_as_float(gr[g4][4 * e + 1]), inv, t.y);
              pre[g4][4 * e + 2] = fmaf(__uint_as_float(gr[g4][4 * e + 2]), inv, t.z); pre[g4][4 * e + 3] = fmaf(__uint_as_float(gr[g4][4 * e + 3]), inv, t.w);
            }
          }
          if (sub == 0) {             // second half of this step's table rows: in flight during the cell math below
#pragma unroll
            for (int g4 = 0; g4 < 4; ++g4) { ad[2 * g4] = __ldg(ap + (size_t)(8 * g4 + 2) * HW); ad[2 * g4 + 1] = __ldg(ap + (size_t)(8 * g4 + 3) * HW); }
          } else if (sidx + 1 < nsteps) {      // first half of the NEXT step's rows
            const float4* apn = adT + (size_t)(pn * 32 + half * 4) * HW + own_pix(r0n, c0n, mtn);
#pragma unroll
            for (int g4 = 0; g4 < 4; ++g4) { ad[2 * g4] = __ldg(apn + (size_t)(8 * g4) * HW); ad[2 * g4 + 1] = __ldg(apn + (size_t)(8 * g4 + 1) * HW); }
          }
        } else {
          const float* sb = s_bias + p * kGtNP + half * 16 + 8 * sub;
#pragma unroll
          for (int g4 = 0; g4 < 4; ++g4)
#pragma unroll
            for (int e = 0; e < 8; ++e) pre[g4][e] = fmaf(__uint_as_float(gr[g4][e]), inv, sb[32 * g4 + e]);
        }
        float cp[8];
        if (a.c_prev) {
#pragma unroll
          for (int e = 0; e < 2; ++e) { const float4 t = stC[lane * 4 + ((2 * sub + e) ^ sw)]; cp[4 * e] = t.x; cp[4 * e + 1] = t.y; cp[4 * e + 2] = t.z; cp[4 * e + 3] = t.w; }
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) cp[e] = 0.f;
        }
        if (sub == 1 && a.c_prev && sidx + 1 < nsteps) {     // c_prev block of the next step (kept in registers until its turn)
#pragma unroll
          for (int j = 0; j < 4; ++j) { bool ok; const size_t o = coop_off(r0n, c0n, bn, pn, mtn, j, ok); tc[j] = ok ? __ldg(reinterpret_cast<const float4*>(a.c_prev + o)) : make_float4(0.f, 0.f, 0.f, 0.f); }
        }
        float cn[8], hn[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) gt_cell(pre[0][e], pre[1][e], pre[2][e], pre[3][e], cp[e], cn[e], hn[e]);
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          stC[lane * 4 + ((2 * sub + e) ^ sw)] = make_float4(cn[4 * e], cn[4 * e + 1], cn[4 * e + 2], cn[4 * e + 3]);
          stH[lane * 4 + ((2 * sub + e) ^ sw)] = make_float4(hn[4 * e], hn[4 * e + 1], hn[4 * e + 2], hn[4 * e + 3]);
        }
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        bool ok;
        const size_t o = coop_off(r0, c0, b, p, mt, j, ok);
        if (ok) {
          const int si = gt_sw(8 * j + (lane >> 2), lane & 3);
          *reinterpret_cast<float4*>(a.c_out + o) = stC[si];
          *reinterpret_cast<float4*>(a.h_out + o) = stH[si];
        }
      }
      __syncwarp();
      r0 = r0n; c0 = c0n; b = bn;
    }
#ifdef TMG_GT_PROFILE
    if (tid == 0) { gt_prof[blockIdx.x * 16 + 4] = clock64() - gtp_start; gt_prof[blockIdx.x * 16 + 5] = gtp_acc[0]; }
#endif
  } else if (warp < 10) {
    // =========================================================== MMA issue: warp 8 -> M tile 0, warp 9 -> M tile 1
    const int mt = warp - 8;
    if (elect_one()) {
      const uint32_t idesc = cv_idesc_f16(kGtNP);
      const uint64_t hlA16 = g.hlA >> 4, hlB16 = g.hlB >> 4;
#ifdef TMG_GT_PROFILE
      long long gtp_acc[4] = {0, 0, 0, 0};
      const long long gtp_start = clock64();
      unsigned long long gtp_ns0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gtp_ns0));
#endif
      int ub = 0; uint32_t pb = 0;              // weight ring slot / phase parity (no runtime div / mod on the issue path)
      int ua0 = 0; uint32_t pa0 = 0;            // A ring slot / parity of the tile's first K-step
      for (int k = 0; k < nmy; ++k) {
        for (int p = 0; p < 2; ++p) {
          const uint32_t tacc = tmem_base + (uint32_t)((p * 2 + mt) * kGtNP);
          { GTP_T0(); if (k >= 1) mbar_wait(acc_free + p, (uint32_t)((k - 1) & 1)); GTP_ADD(0); }
          tc_fence_after();
          int ua = ua0; uint32_t pa = pa0;
          // One K-step = 9 taps, unrolled and branch-free: the issue path is as critical as the tensor pipe (a data-dependent
          // branch per tap doubled the cycles per issued MMA, 55 -> 120, and made the kernel issue-bound).  PAIR = the spare
          // plane variant of the last K-step: [x_hi | x_lo] x [W_hi ; W_hi] = hi*hi + lo*hi, [x_hi | 0] x [W_lo ; 0] = hi*lo.
          auto kstep = [&](auto pair_c, uint32_t acc0) {
            constexpr bool PAIR = decltype(pair_c)::value;
            { GTP_T0(); mbar_wait(a_full + ua, pa); GTP_ADD(1); }
            tc_fence_after();
            const uint64_t a0 = make_desc(smem_u32(As + (size_t)ua * g.bufA), kGtPLB, kGtRP * 16) + (uint64_t)(8 * mt);
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              { GTP_T0(); mbar_wait(b_full + ub, pb); GTP_ADD(2); }
              tc_fence_after();
              const uint64_t bd = make_desc(smem_u32(Bs + (size_t)ub * g.stageB), (uint32_t)kGtNP * 16u, 128);
              const int dr = tap / 3, dc = tap - 3 * dr;
              const uint64_t ad = a0 + (uint64_t)(dr * kGtRP + dc);
              cv_mma_f16(tacc, ad, bd, idesc, tap > 0 ? 1u : acc0);
              if (X3) {
                if (PAIR) {
                  cv_mma_f16(tacc, ad + hlA16, bd + hlB16, idesc, 1u);
                } else {
                  cv_mma_f16(tacc, ad + hlA16, bd, idesc, 1u);
                  cv_mma_f16(tacc, ad, bd + hlB16, idesc, 1u);
                }
              }
              mma_commit(b_free + ub);
              if (++ub == g.NB) { ub = 0; pb ^= 1u; }
            }
            if (p == 1) mma_commit(a_free + ua);
            if (++ua == g.NA) { ua = 0; pa ^= 1u; }
          };
          const int ks_full = g.odd ? g.KS - 1 : g.KS;
          for (int ks = 0; ks < ks_full; ++ks) kstep(std::false_type{}, ks > 0 ? 1u : 0u);
          if (g.odd) kstep(std::true_type{}, ks_full > 0 ? 1u : 0u);
          mma_commit(acc_full + p);
          if (p == 1) { ua0 = ua; pa0 = pa; }
        }
      }
#ifdef TMG_GT_PROFILE
      if (mt == 0) {
        gt_prof[blockIdx.x * 16 + 0] = clock64() - gtp_start;
        for (int i = 0; i < 3; ++i) gt_prof[blockIdx.x * 16 + 1 + i] = gtp_acc[i];
        unsigned long long gtp_ns1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gtp_ns1));
        gt_prof[blockIdx.x * 16 + 8] = (long long)(gtp_ns1 - gtp_ns0);
      }
#endif
    }
  } else if (warp == 10) {
    // =========================================================== weight streaming (one lane, cp.async.bulk)
    if (lane == 0) {
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(a.wpk);
      const int nstage_tile = 2 * g.KS * 9;
      const uint32_t gstage = 2 * g.hlB;          // packed stages always hold hi + lo
      int ub = 0; uint32_t pb = 1;               // parity of the PREVIOUS use of the slot; first round: nothing to wait for
      bool first = true;
      for (int k = 0; k < nmy; ++k) {
        for (int st = 0; st < nstage_tile; ++st) {
          if (!first) mbar_wait(b_free + ub, pb);
          mbar_expect_tx(b_full + ub, g.stageB);
          bulk_g2s(Bs + (size_t)ub * g.stageB, wsrc + (size_t)st * gstage, g.stageB, b_full + ub);
          if (++ub == g.NB) { ub = 0; pb ^= 1u; first = false; }
        }
      }
    }
  } else {
    // =========================================================== producers (128 threads): activation K-steps
    // item = (tile position, 16-byte quarter of the K-step's 16 channels): 4 consecutive lanes read 64 contiguous bytes
    const int ptid = tid - 11 * 32;
    GtTileIt it;
    it.init(blockIdx.x, tiles_img);
#ifdef TMG_GT_PROFILE
    long long gtp_acc[4] = {0, 0, 0, 0};
    const long long gtp_start = clock64();
#endif
    int ua = 0; uint32_t pa = 1;                 // parity of the previous use of the slot
    bool first = true;
    for (int k = 0; k < nmy; ++k) {
      int r0, c0;
      it.origin(g, r0, c0);
      const int b = it.b;
      it.advance(g, tiles_img);
      for (int ks = 0; ks < g.KS; ++ks) {
        // this thread's quarter q = ptid & 3 of the K-step's 16 channels is fixed (128 threads, 4 quarters per position):
        // one source, one channel offset per K-step; ALL its positions are loaded before the first conversion, so a K-step
        // costs one DRAM round trip (the states do not fit L2) instead of one per batch
        // odd plane count: the spare plane of the last K-step is fed from the same channels as its real plane -- the hi part
        // of the buffer holds [x_hi | x_lo], the lo part [x_hi | 0] (two MMAs instead of three, see the issuer)
        const bool dup = g.odd && ks == g.KS - 1;
        const int q = ptid & 3, plane = 2 * ks + (dup ? 0 : (q >> 1));
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
        constexpr int kItems = (kGtNPOS + kGtNPT / 4 - 1) / (kGtNPT / 4);       // positions per thread: 324 / 32 -> 11
        float4 v[kItems];
#pragma unroll
        for (int u = 0; u < kItems; ++u) {
          const int p = (ptid >> 2) + u * (kGtNPT / 4);
          v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          const int rr = p / kGtRP, rc = p - rr * kGtRP;
          const int r = r0 - 1 + rr, c = c0 - 1 + rc;
          if (live && p < kGtNPOS && r >= 0 && r < a.H && c >= 0 && c < a.W) {
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
        { GTP_T0(); if (!first) mbar_wait(a_free + ua, pa); GTP_ADD(0); }
        uint8_t* Ab = As + (size_t)ua * g.bufA + (q >> 1) * kGtPLB + (q & 1) * 8;
        const float lo_ = s_relu ? 0.f : -60000.f;
#pragma unroll
        for (int u = 0; u < kItems; ++u) {
          const int p = (ptid >> 2) + u * (kGtNPT / 4);
          if (p >= kGtNPOS) continue;
          const float y0 = fmaxf(fminf(v[u].x, 60000.f), lo_), y1 = fmaxf(fminf(v[u].y, 60000.f), lo_);
          const float y2 = fmaxf(fminf(v[u].z, 60000.f), lo_), y3 = fmaxf(fminf(v[u].w, 60000.f), lo_);
          const __half2 h01 = __floats2half2_rn(y0, y1), h23 = __floats2half2_rn(y2, y3);
          uint8_t* dst = Ab + p * 16;
          const uint2 hi2 = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
          if (X3) {
            const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
            const __half2 l01 = __floats2half2_rn(y0 - f01.x, y1 - f01.y), l23 = __floats2half2_rn(y2 - f23.x, y3 - f23.y);
            const uint2 lo2 = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
            if (!dup) {
              *reinterpret_cast<uint2*>(dst) = hi2;
              *reinterpret_cast<uint2*>(dst + g.hlA) = lo2;
            } else if ((q >> 1) == 0) {        // real plane: x_hi in both parts
              *reinterpret_cast<uint2*>(dst) = hi2;
              *reinterpret_cast<uint2*>(dst + g.hlA) = hi2;
            } else {                           // spare plane: x_lo in the hi part, zeros in the lo part
              *reinterpret_cast<uint2*>(dst) = lo2;
              *reinterpret_cast<uint2*>(dst + g.hlA) = make_uint2(0u, 0u);
            }
          } else {
            *reinterpret_cast<uint2*>(dst) = hi2;
          }
        }
        fence_proxy_async();
        mbar_arrive(a_full + ua);
        if (++ua == g.NA) { ua = 0; pa ^= 1u; first = false; }
      }
    }
#ifdef TMG_GT_PROFILE
    if (ptid == 0) { gt_prof[blockIdx.x * 16 + 6] = clock64() - gtp_start; gt_prof[blockIdx.x * 16 + 7] = gtp_acc[0]; }
#endif
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ------------------------------------------------------------------ hoisted conditioning table -> plane-transposed + bias
// src [HW][ncol] (columns in pass order, pack.cu JOB_SLICE with the gate permutation), dst [ncol / 4][HW] float4,
// bias (original gate-row order, may be null) added on the way: column n' = 128 p + 32 g + j  <-  bias[g R + 32 p + j]
__global__ void __launch_bounds__(256)
gate_addend_transpose_kernel(const float* __restrict__ src, const float* __restrict__ bias, float4* __restrict__ dst, int HW, int ncol, int R) {
  __shared__ float tile[32][33 * 4];
  const int g0 = blockIdx.y * 32;          // first column group of this block
  const int p0 = blockIdx.x * 32;          // first pixel
  const int ngroups = ncol / 4;
  // read: 32 pixels x 128 columns, coalesced along the columns
  for (int i = threadIdx.x; i < 32 * 32; i += 256) {
    const int px = i >> 5, gq = i & 31;
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p0 + px < HW && g0 + gq < ngroups) t = __ldg(reinterpret_cast<const float4*>(src + (size_t)(p0 + px) * ncol) + g0 + gq);
    float* d = &tile[px][gq * 4 + (gq >> 3)];   // pad one float per 8 groups
    d[0] = t.x; d[1] = t.y; d[2] = t.z; d[3] = t.w;
  }
  __syncthreads();
  const int RH = R / 2;
  for (int i = threadIdx.x; i < 32 * 32; i += 256) {
    const int gq = i >> 5, px = i & 31;
    if (p0 + px >= HW || g0 + gq >= ngroups) continue;
    const float* s = &tile[px][gq * 4 + (gq >> 3)];
    float4 t = make_float4(s[0], s[1], s[2], s[3]);
    if (bias) {
      const int n = (g0 + gq) * 4, p = n / (4 * RH), nn = n % (4 * RH);
      const float* bp = bias + (nn / RH) * R + p * RH + (nn % RH);
      t.x += __ldg(bp); t.y += __ldg(bp + 1); t.z += __ldg(bp + 2); t.w += __ldg(bp + 3);
    }
    dst[(size_t)(g0 + gq) * HW + p0 + px] = t;
  }
}

int launch_gate_addend_transpose(const float* src, const float* bias, float* dst, int HW, int ncol, int R, cudaStream_t st) {
  if (ncol % 4 || ncol != 4 * R) { set_error("gate addend transpose: %d columns, R = %d", ncol, R); return TMG_ERR_BAD_SHAPE; }
  gate_addend_transpose_kernel<<<dim3((unsigned)cdiv(HW, 32), (unsigned)cdiv(ncol / 4, 32)), 256, 0, st>>>(src, bias, reinterpret_cast<float4*>(dst), HW, ncol, R);
  TMG_LAUNCH_CHECK();
  return TMG_OK;
}

// ------------------------------------------------------------------ host side
size_t lstm_gate_packed_floats(const int* nch, int nsrc) {     // [pass][ks][tap][hl][2 planes][128][8 halves]
  return (size_t)2 * convf16_ksteps(nch, nsrc) * 9 * 2 * 2 * kGtNP * 16 / 4;
}

static bool gt_geom(const ConvF16Args& a, GateGeom& g, int grid) {
  if (a.lstm_R != 64 || a.nsrc < 1 || a.nsrc > 3 || a.pad_replicate || a.in_scale || a.ndst) return false;
  int pl = 0;
  for (int i = 0; i < 3; ++i) {
    g.plane0[i] = pl;
    g.nplanes[i] = i < a.nsrc ? (a.src[i].nch + 7) / 8 : 0;
    pl += g.nplanes[i];
  }
  g.KS = (pl + 1) / 2;
  g.nhl = a.x3 ? 2 : 1;
  static const bool no_odd = [] { const char* e = getenv("TMG_GT_NO_ODD"); return e && e[0] == '1'; }();   // A/B runs
  g.odd = (a.x3 && (pl & 1) && !no_odd) ? 1 : 0;
  g.hlA = 2 * kGtPLB; g.bufA = g.hlA * g.nhl;
  g.hlB = (uint32_t)2 * kGtNP * 16; g.stageB = g.hlB * g.nhl;
  g.tiles_x = cdiv(a.W, 16); g.tiles_y = cdiv(a.H, 16);
  const int tiles_img = g.tiles_x * g.tiles_y;
  if (tiles_img >= 4096) return false;
  g.ntiles = tiles_img * a.B;
  g.step_b = grid / tiles_img; g.step_t = grid % tiles_img;
  g.inv_tx = (uint32_t)((65536 + g.tiles_x - 1) / g.tiles_x);
  uint32_t off = 0;
  auto take = [&](uint32_t n) { uint32_t o = off; off += (n + 127) / 128 * 128; return o; };
  g.oStage = take(8 * kGtStageW);
  g.oMisc = take(2 * kGtNP * 4);
  g.oBar = take(49 * 8 + 16);
  const uint32_t budget = 227u * 1024u - off;
  // the K-steps of a tile stay staged through both passes: ring >= KS + 1; weight ring >= 3 stages
  static const int na_extra = [] { const char* e = getenv("TMG_GT_NA_EXTRA"); return e ? atoi(e) : 1; }();   // A/B runs
  int NA = std::min(g.KS + na_extra, kGtMaxNA);
  while (NA > g.KS && (uint32_t)NA * g.bufA + 3u * g.stageB > budget) --NA;
  if (NA < g.KS + 1 || NA > kGtMaxNA) return false;
  g.NA = NA;
  int NB = (int)((budget - (uint32_t)NA * g.bufA) / g.stageB);
  NB = std::min(NB, kGtMaxNB);
  if (NB < 3) return false;
  g.NB = NB;
  g.oA = take((uint32_t)NA * g.bufA);
  g.oB = take((uint32_t)NB * g.stageB);
  g.total = off;
  return g.total <= 227 * 1024;
}

bool lstm_gate_f16_supported(const ConvF16Args& a) {
  GateGeom g{};
  return gt_geom(a, g, 148);
}

int launch_lstm_gate_f16(const ConvF16Args& a, cudaStream_t st) {
  if (a.B <= 0 || a.H <= 0 || a.W <= 0) return TMG_OK;
  const int tiles = cdiv(a.W, 16) * cdiv(a.H, 16) * a.B;
  int dev = 0, nsm = 148;
  cudaGetDevice(&dev);
  static int cached[64] = {0};                  // SM count per device, queried once
  if (dev >= 0 && dev < 64) {
    if (!cached[dev]) cudaDeviceGetAttribute(&cached[dev], cudaDevAttrMultiProcessorCount, dev);
    if (cached[dev] > 0) nsm = cached[dev];
  }
  const int grid = std::min(tiles, nsm);
  GateGeom g{};
  if (!gt_geom(a, g, grid) || !a.h_out || !a.c_out) {
    set_error("two-pass ConvLSTM gate kernel: unsupported shape (R=%d, %dx%d, %d sources)", a.lstm_R, a.H, a.W, a.nsrc);
    return TMG_ERR_UNSUPPORTED;
  }
  if (a.x3) {
    TMG_SMEM_ATTR(lstm_gate_f16_kernel<true>, 227 * 1024);
    lstm_gate_f16_kernel<true><<<grid, kGtThreads, g.total, st>>>(a, g);
  } else {
    TMG_SMEM_ATTR(lstm_gate_f16_kernel<false>, 227 * 1024);
    lstm_gate_f16_kernel<false><<<grid, kGtThreads, g.total, st>>>(a, g);
  }
  TMG_LAUNCH_CHECK();
#ifdef TMG_GT_PROFILE
  if (getenv("TMG_GT_PROF") && tiles >= 8 * grid) {
    static int left = atoi(getenv("TMG_GT_PROF"));
    if (left > 0) {
      --left;
      cudaStreamSynchronize(st);
      static long long h[148 * 16];
      cudaMemcpyFromSymbol(h, gt_prof, sizeof(h));
      double s8[9] = {0};
      for (int b = 0; b < grid; ++b) for (int i = 0; i < 9; ++i) s8[i] += (double)h[b * 16 + i] / grid;
      const double nt = (double)tiles / grid;
      fprintf(stderr, "gate2p %dx%d B=%d KS=%d NA=%d NB=%d: cycles/tile issuer %.0f (acc_free %.0f a_full %.0f b_full %.0f) | epilogue %.0f (acc_full %.0f) | producer %.0f (a_free %.0f) | %.3f ms, SM clock %.0f MHz\n",
              a.H, a.W, a.B, g.KS, g.NA, g.NB, s8[0] / nt, s8[1] / nt, s8[2] / nt, s8[3] / nt, s8[4] / nt, s8[5] / nt, s8[6] / nt, s8[7] / nt, s8[8] * 1e-6, s8[0] / s8[8] * 1e3);
    }
  }
#endif
  return TMG_OK;
}

}  // namespace tmg
