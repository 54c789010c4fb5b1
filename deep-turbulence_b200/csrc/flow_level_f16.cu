// Level-resident fused flow steps (sm_100a): every plain step of one flow level in ONE launch, reverse (sampling) direction.
//
// flow_step_f16.cu runs one launch per flow step: the state [B,HW,C] makes a full HBM round trip per step, every tile is
// staged with a 3-pixel halo (484 staged positions for 256 outputs), the fp32 -> fp16 hi/lo conversion is done by producer
// warps from global memory, and a tile passes through six synchronisation stages.  Measured (S = 4096, round 1): level 0
// 964 us, level 1 449 us, level 2 365 us per step = 8x / 7x / 6x above the HBM time of the step.
//
// Here a CTA owns G whole samples of the level for all n-1 plain steps (flowLSTMBlock.py:348-359, reverse order):
//   * the fp32 flow state lives in the REGISTERS of the thread that owns the pixel (same thread in every step);
//   * the coupling-net input relu(x1) (+ d1, d2) lives in shared memory as fp16 hi/lo operand planes in a FLAT PADDED layout:
//     each sample is stored with a one-pixel replicated ring (pitch W+2), samples back to back, so a filter tap is a linear
//     offset of the operand descriptor and no halo is ever recomputed or re-read; M tiles are 128 consecutive positions;
//   * per step: E = un-shifted GEMM with the 9 taps of both Cout=1 dense layers in N (as in flow_step_f16.cu), the tap
//     partials are stored pre-shifted so that position p's row holds the nine terms of its 3x3 sum (zero padding = rows of
//     the ring never written); d1 -> d2 -> operand slots; Z = Conv2dZeros as tap-shifted implicit GEMM; epilogue = coupling,
//     1x1 mix, ActNorm on the register state, then the NEXT step's operand planes are written straight from registers;
//   * weights of step s+1 stream in (cp.async.bulk + mbarrier) while step s computes; the conditioning enters through the
//     hoisted tables dc / hc (model.cu run_hoist), exactly as in flow_step_f16.cu;
//   * HBM traffic per sample and level: read the state once, write it once (+ the hoisted tables from L2).
// Same arithmetic as flow_step_f16.cu (fp16 hi/lo split, three MMAs, fp32 accumulate, exact-fp32 epilogue); only summation
// order differs.  Reference: AffineCouplingLayer.reverse flowAffine.py:84-109, InvertibleConv1x1LU.reverse glowConv.py:196-222,
// ActNorm.reverse actNorm.py:69-85, step wrappers flowLSTMBlock.py:71-86,132-146.
#include <cuda_fp16.h>

#include <type_traits>
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace tmg {

constexpr int kLvMaxTiles = 20;

struct LevelGeom {
  int P, npos_s, G, T, M0, NA, Dt, NES, nzs, npass, zcols;
  uint32_t PLB, hlA;                     // bytes of one operand plane, hi -> lo distance
  uint32_t oA, oS1, oS2, oD1, oEdge, oY2, oWE, oWZ, oSm, oBar, total;
  uint32_t cpE_n, cpE_bytes, cpE_sstep, cpZ_n, cpZ_bytes, cpZ_sstep;   // bulk copies of one step's weights: count, size, source step
  uint32_t wE_hl, wZ_hl, wZ_tap;         // shared-memory strides of the weight copies
  uint32_t gE_hl, gZ_hl, gZ_tap;         // strides of the packed (global) weights
  uint32_t wE_stage, wZ_stage, sm_stage; // bytes per stage
  // MMA mix (C >= 24): NMX staging buffers for the state as fp16 hi/lo A operand [hl][K plane][128 rows][16 B], two stages of
  // the per-step W operand [hl][K plane][NP rows][16 B]
  int NMX;
  uint32_t oMXA, oMXW, mxa_bytes, mxw_stage;
  const int64_t* wmx;                    // device [nsteps]: packed-buffer offsets of the steps' mix operands
  float* unsq;                           // wide levels, optional: write the level's result un-squeezed (CheckerSqueeze.reverse,
  int unsq_cstride;                      //   flowUtils.py:124-145) into the next level's state [B, 2H, 2W, unsq_cstride] instead of y_out
};

__device__ __forceinline__ uint32_t lv_idesc(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24); }
__device__ __forceinline__ void lv_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void lv_ld32(uint32_t taddr, float* v) {
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

// relu(v) (v >= 0 after the max) as an fp16 hi/lo pair of pairs; values beyond the fp16 operand range are clamped and flagged
__device__ __forceinline__ void lv_split2(float y0, float y1, uint32_t& hi, uint32_t& lo, bool& ovf) {
  ovf = ovf || y0 > 60000.f || y1 > 60000.f;
  y0 = fminf(y0, 60000.f); y1 = fminf(y1, 60000.f);
  const __half2 h2 = __floats2half2_rn(y0, y1);
  const float2 hf = __half22float2(h2);
  const __half2 l2 = __floats2half2_rn(y0 - hf.x, y1 - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h2);
  lo = *reinterpret_cast<const uint32_t*>(&l2);
}

// signed variant (the flow state itself): clamped to +-6e4, flagged
__device__ __forceinline__ void lv_split2s(float y0, float y1, uint32_t& hi, uint32_t& lo, bool& ovf) {
  ovf = ovf || fabsf(y0) > 60000.f || fabsf(y1) > 60000.f;
  y0 = fmaxf(fminf(y0, 60000.f), -60000.f); y1 = fmaxf(fminf(y1, 60000.f), -60000.f);
  const __half2 h2 = __floats2half2_rn(y0, y1);
  const float2 hf = __half22float2(h2);
  const __half2 l2 = __floats2half2_rn(y0 - hf.x, y1 - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h2);
  lo = *reinterpret_cast<const uint32_t*>(&l2);
}

// developer profiling: cycles between marks, accumulated per role (one thread per role writes at the end)
// (compiled in only with -DTMG_LV_PROFILE: the counters cost 24 registers per thread)
#ifdef TMG_LV_PROFILE
#define LVP_DECL long long pt_ = 0, pacc_[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; const bool prof_on_ = a.prof != nullptr; if (prof_on_) pt_ = clock64();
#define LVP_MARK(i) if (prof_on_) { const long long n_ = clock64(); pacc_[i] += n_ - pt_; pt_ = n_; }
#define LVP_FLUSH(base) if (prof_on_) { for (int i_ = 0; i_ < 12; ++i_) a.prof[(size_t)blockIdx.x * 48 + (base) + i_] = pacc_[i_]; }
#else
#define LVP_DECL
#define LVP_MARK(i)
#define LVP_FLUSH(base)
#endif

// packed fp32 FMA (FFMA2 on sm_100a): (d.x, d.y) = (a.x, a.y) * (b, b) + (c.x, c.y) -- two outputs of the 1x1 mix per issue slot
__device__ __forceinline__ float2 lv_ffma2(float2 a, float b, float2 c) {
  const float2 bb = make_float2(b, b);
  unsigned long long ra = *reinterpret_cast<const unsigned long long*>(&a), rb = *reinterpret_cast<const unsigned long long*>(&bb),
                     rc = *reinterpret_cast<const unsigned long long*>(&c), rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}

// flag bits of an owned position
constexpr uint32_t kFValid = 1u, kFLeft = 2u, kFRight = 4u, kFTop = 8u, kFBottom = 16u, kFOk = 32u;   // bits 8..: sample in the pass

// CP ("compact, paired"; narrow levels, C/2 + 2 <= 8): the coupling-net input of a position is ONE 16-byte operand unit
// [x1 (<= 6 ch) | d1 | d2]; a K = 16 MMA contracts TWO filter taps -- its second K half is the unit of another position,
// reached through the descriptor's leading-dimension byte offset (for taps in the same row that offset is 16 bytes: the core
// matrices overlap) -- so Conv2dZeros takes 5 MMAs per M tile and operand pass instead of 9, the operand planes take 32
// instead of 64 bytes per position, and the tap partials of the dense layers are reduced along the row with warp shuffles
// before they go to shared memory (24 instead of 72 bytes per position): a whole 32 x 64 sample of level 0 fits on one SM.
// MX: the 1x1 mix u = W v as a tensor-core GEMM (M = 128 pixels of a tile, N = C, K = C, fp16 hi/lo split like the
// convolutions) instead of C^2 FMAs per pixel: each thread writes its state row as an A operand, the E issuer (idle most of
// a step) issues the MMAs into the tile's (already consumed) Conv2dZeros accumulator columns, the thread reads its row back.
// (A fifth service warp would take the kernel from 128 to 96 registers per thread: register allocation is per 4 warps.)
template <int C, bool X3, int NTG, int NT, bool CP, bool MX>
__global__ void __launch_bounds__((4 * NTG + 4) * 32, 1)
flow_level_kernel(LevelArgs a, LevelGeom g) {
  constexpr int NP = (C + 15) / 16 * 16;
  constexpr int KSy = CP ? 1 : (C / 2 + 2 + 15) / 16;
  constexpr int NPL = CP ? 1 : 2 * KSy;         // operand planes per hi / lo half
  constexpr int kd = CP ? 6 : KSy * 16 - 2;
  constexpr int NXP = (C / 2 + 7) / 8;          // operand planes that carry x1 channels
  constexpr int NE = 4 * NTG;
  constexpr int kEpi = NE * 32;
  constexpr int kThreads = (NE + 4) * 32;
  constexpr int CC = C * C;
  constexpr int CCs = MX ? 0 : CC;               // floats of W in the small-vector stage (MX: W lives in its own operand stage)
  constexpr int PM = (C / 8 + 1) / 2 * 2;        // MX: K planes of the mix operands (even)
  static_assert(!MX || (!CP && C % 8 == 0), "MMA mix: wide levels only");
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int HW = a.H * a.W, P = g.P, T = g.T;

  uint8_t* A = smem + g.oA;
  float* S1 = reinterpret_cast<float*>(smem + g.oS1);
  float* S2 = reinterpret_cast<float*>(smem + g.oS2);
  float* D1 = reinterpret_cast<float*>(smem + g.oD1);
  float* EDGE = reinterpret_cast<float*>(smem + g.oEdge);       // CP: row-sum terms that cross a warp boundary
  float* Y2S = reinterpret_cast<float*>(smem + g.oY2);          // CP: second half of the fp32 state [position][C/2]
  uint8_t* WE = smem + g.oWE;
  uint8_t* WZ = smem + g.oWZ;
  uint8_t* SM = smem + g.oSm;
  uint8_t* MXA = smem + g.oMXA;
  uint8_t* MXW = smem + g.oMXW;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + g.oBar);
  uint64_t* we_full = bars;                      // [2] E weights + small per-step vectors: loader
  uint64_t* we_free = bars + 2;                  // [2] E issuer commit + every epilogue thread
  uint64_t* wz_full = bars + 4;                  // [2]
  uint64_t* wz_free = bars + 6;                  // [2] both Z issuers
  uint64_t* d_ready = bars + 8;                  // [1] every epilogue thread, once per step
  uint64_t* a_ready = bars + 10;                 // [T] the 128 threads of a tile
  uint64_t* e_full = a_ready + kLvMaxTiles;      // [NES] commit
  uint64_t* e_free = e_full + kLvMaxTiles;       // [NES] 128
  uint64_t* z_full = e_free + kLvMaxTiles;       // [T] commit
  uint64_t* mx_ready = z_full + kLvMaxTiles;     // [T] MX: the 128 threads of a tile wrote their state rows
  uint64_t* mx_full = mx_ready + kLvMaxTiles;    // [T] MX: commit of the tile's mix MMAs
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mx_full + kLvMaxTiles);

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(we_full + i, 1); mbar_init(we_free + i, 1 + kEpi); mbar_init(wz_full + i, 1); mbar_init(wz_free + i, 2); }
    mbar_init(d_ready, kEpi);
    for (int i = 0; i < kLvMaxTiles; ++i) { mbar_init(a_ready + i, 128); mbar_init(e_full + i, 1); mbar_init(e_free + i, 128); mbar_init(z_full + i, 1); }
    if constexpr (MX) { for (int i = 0; i < kLvMaxTiles; ++i) { mbar_init(mx_ready + i, 128); mbar_init(mx_full + i, 1); } }
    fence_barrier_init();
  }
  if (warp == NE + 1) tmem_alloc(tmem_slot, 512);
  {
    // operand planes, tap-partial scratch and D1 start as zeros: padding channels, the margins and the rows / slots of the
    // ring that are never written stay finite zeros for the whole launch (zero padding of the dense layers relies on it)
    uint4* z4 = reinterpret_cast<uint4*>(smem + g.oA);
    const int n4 = (int)((g.oWE - g.oA) / 16);
    for (int i = tid; i < n4; i += kThreads) z4[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int npass_mine = ((int)blockIdx.x < g.npass) ? (g.npass - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int nsteps = a.nsteps;

  if (warp < NE) {
    // =========================================================== epilogue / state owners
    const int q = warp & 3, tg = warp >> 2;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const int P9 = P * 9;
    // static geometry of the owned positions
    int pix[NT]; uint32_t flg[NT];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const int mt = tg + j * NTG;
      pix[j] = 0; flg[j] = 0;
      if (mt < T) {
        const int pos = mt * 128 + q * 32 + lane;
        const int gi = pos / g.npos_s, rem = pos - gi * g.npos_s;
        const int row = rem / P, col = rem - row * P;
        if (gi < g.G && row >= 1 && row <= a.H && col >= 1 && col <= a.W) {
          flg[j] = kFValid | (col == 1 ? kFLeft : 0u) | (col == a.W ? kFRight : 0u) | (row == 1 ? kFTop : 0u) | (row == a.H ? kFBottom : 0u);
          pix[j] = (row - 1) * a.W + (col - 1);
          flg[j] |= (uint32_t)gi << 8;
        }
      }
    }
    // fp32 flow state of the owned pixels.  CP keeps only the first half (y1, the coupling-net input) in registers and the
    // second half in shared memory: four tiles per thread would otherwise leave no registers to prefetch into.
    constexpr int NSR = CP ? C / 2 : C;
    float st[NT][NSR];
    float ldacc[NT];
    bool ovf = false;
#define OKJ(j) ((flg[j] & kFOk) != 0u)
#define SAMPLE(j) (pass * g.G + (int)(flg[j] >> 8))

    // writes relu(x[0 .. C/2)) as hi/lo operand planes at position ap (and at the ring positions it replicates into)
    auto write_planes = [&](const float* x, int ap, uint32_t f) {
      uint32_t hi[NXP][4], lo[NXP][4];
#pragma unroll
      for (int p = 0; p < NXP; ++p) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c0 = 8 * p + 2 * e;
          const float y0 = c0 < C / 2 ? fmaxf(x[c0 < C / 2 ? c0 : 0], 0.f) : 0.f;
          const float y1 = c0 + 1 < C / 2 ? fmaxf(x[c0 + 1 < C / 2 ? c0 + 1 : 0], 0.f) : 0.f;
          lv_split2(y0, y1, hi[p][e], lo[p][e], ovf);
        }
      }
      auto put = [&](int at) {
#pragma unroll
        for (int p = 0; p < NXP; ++p) {
          uint8_t* dst = A + (size_t)p * g.PLB + (size_t)at * 16;
          *reinterpret_cast<uint4*>(dst) = make_uint4(hi[p][0], hi[p][1], hi[p][2], hi[p][3]);
          if (X3) *reinterpret_cast<uint4*>(dst + g.hlA) = make_uint4(lo[p][0], lo[p][1], lo[p][2], lo[p][3]);
        }
      };
      put(ap);
      if (f & (kFLeft | kFRight | kFTop | kFBottom)) {
        const int dx = (f & kFLeft) ? -1 : ((f & kFRight) ? 1 : 0);
        const int dy = (f & kFTop) ? -P : ((f & kFBottom) ? P : 0);
        if (dx) put(ap + dx);
        if (dy) put(ap + dy);
        if (dx && dy) put(ap + dx + dy);
      }
    };
    // the (d1, d2) operand slot: last two channels of plane kd / 8
    auto write_d = [&](float d1, float d2, int ap, uint32_t f) {
      uint32_t hi, lo;
      lv_split2(d1, d2, hi, lo, ovf);
      auto put = [&](int at) {
        uint8_t* dst = A + (size_t)(kd >> 3) * g.PLB + (size_t)at * 16 + (kd & 7) * 2;
        *reinterpret_cast<uint32_t*>(dst) = hi;
        if (X3) *reinterpret_cast<uint32_t*>(dst + g.hlA) = lo;
      };
      put(ap);
      if (f & (kFLeft | kFRight | kFTop | kFBottom)) {
        const int dx = (f & kFLeft) ? -1 : ((f & kFRight) ? 1 : 0);
        const int dy = (f & kFTop) ? -P : ((f & kFBottom) ? P : 0);
        if (dx) put(ap + dx);
        if (dy) put(ap + dy);
        if (dx && dy) put(ap + dx + dy);
      }
    };

    // CP: the 3x3 sum at position ap = three row sums (written pre-shifted by the source rows) + the boundary-lane terms
    auto row_sum3 = [&](const float* R, int layer, int ap) -> float {
      const float* r = R + ap * 3;
      float acc = (r[0] + r[1]) + r[2];
#pragma unroll
      for (int dyi = 0; dyi < 3; ++dyi) {
        const int c = ap - g.M0 + (dyi - 1) * P;                // tile-space position of the thread that wrote r[dyi]
        const int cl = c & 31;
        if (cl == 0) acc += EDGE[((c >> 5) + 1) * 12 + layer * 3 + dyi];
        else if (cl == 31) acc += EDGE[((c >> 5) + 1) * 12 + 6 + layer * 3 + dyi];
      }
      return acc;
    };
    int gs = 0;                                  // running step counter (all passes)
    LVP_DECL
    for (int ip = 0; ip < npass_mine; ++ip) {
      const int pass = blockIdx.x + ip * gridDim.x;
      // ---- load the state of this pass, write the first step's operand planes
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const int mt = tg + j * NTG;
        ldacc[j] = 0.f;
        flg[j] &= ~kFOk;
        if (mt < T) {
          const int ap = g.M0 + mt * 128 + q * 32 + lane;
          const int b = SAMPLE(j);
          if ((flg[j] & kFValid) && b < a.B) flg[j] |= kFOk;
          if (OKJ(j)) {
            const float4* y4 = reinterpret_cast<const float4*>(a.y_in + ((size_t)b * HW + pix[j]) * C);
#pragma unroll
            for (int c4 = 0; c4 < C / 4; ++c4) {
              const float4 t = __ldg(y4 + c4);
              const float tv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int c = 4 * c4 + e;
                if (c < NSR) st[j][c < NSR ? c : 0] = tv[e];
                else Y2S[(size_t)ap * (C / 2) + (c - C / 2)] = tv[e];
              }
            }
            write_planes(st[j], ap, flg[j]);
          } else {
#pragma unroll
            for (int c = 0; c < NSR; ++c) st[j][c] = 0.f;
          }
          fence_proxy_async();
          mbar_arrive(a_ready + mt);
        }
      }
      LVP_MARK(0)
      for (int is = 0; is < nsteps; ++is, ++gs) {
        const LevelStep* S = a.steps + is;
        const uint32_t par = (uint32_t)(gs & 1);
        const float* sm = reinterpret_cast<const float*>(SM + (size_t)(gs & 1) * g.sm_stage);
        const float* s_nb = sm + CCs + C;
        const float* s_rnw = sm + CCs + 2 * C;
        const float* s_b3 = sm + CCs + 3 * C;
        const float* s_w2d = sm + CCs + 4 * C;
        const int dc_off = __ldg(&S->dc_off), hc_off = __ldg(&S->hc_off);          // step index inside the hoisted tables
        // hoisted conditioning terms of d1 / d2 ([sample][step][pixel] float2): issued now, used after the gather barriers
        float2 dcv[NT];
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          const int pj = OKJ(j) ? pix[j] : 0;
          const int b = (OKJ(j) && a.hoist_bstride) ? SAMPLE(j) : 0;
          dcv[j] = __ldg(reinterpret_cast<const float2*>(a.dc) + ((size_t)b * a.nsteps_tab + dc_off) * HW + pj);
        }
        // ---- G1: E accumulators -> pre-shifted tap partials
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          const int mt = tg + j * NTG;
          if (mt < T) {
            const int ce = gs * T + mt, slot = ce % g.NES, use = ce / g.NES;
            LVP_MARK(1)
            mbar_wait_sleep(e_full + slot, (uint32_t)(use & 1));
            LVP_MARK(2)
            tc_fence_after();
            const int ap = g.M0 + mt * 128 + q * 32 + lane;
            if constexpr (!CP) {
              float v[32];
              lv_ld32(tmem_base + lane_base + (uint32_t)(g.zcols + slot * 32), v);
              tc_fence_before();
              mbar_arrive(e_free + slot);
              if (OKJ(j)) {
                float* s1 = S1 + (size_t)ap * 9;
                float* s2 = S2 + (size_t)ap * 9;
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                  const int dy = t / 3 - 1, dx = t % 3 - 1;
                  s1[t - dy * P9 - dx * 9] = v[t];
                  s2[t - dy * P9 - dx * 9] = v[16 + t];
                }
              }
            } else {
              // row sums R_dy(c) = E[c-1][(dy,-1)] + E[c][(dy,0)] + E[c+1][(dy,+1)] through warp shuffles; the term a
              // boundary lane cannot get from its own warp is left in EDGE by the neighbouring warp and added in G2 / G3.
              // Positions outside the image contribute zeros (zero padding of the dense layers).
              const int blk = mt * 4 + q;
              float* eb = EDGE + (blk + 1) * 12;
              // compact E columns: 0-8 dense layer 1, 9-17 dense layer 2 -> 18 of the slot's 32 columns are read
              uint32_t vr_[18];
              tmem_ld_nowait<16>(tmem_base + lane_base + (uint32_t)(g.zcols + slot * 32), vr_);
              tmem_ld_nowait<2>(tmem_base + lane_base + (uint32_t)(g.zcols + slot * 32 + 16), vr_ + 16);
              tmem_ld_wait();
              tc_fence_before();
              mbar_arrive(e_free + slot);
              const bool okj = OKJ(j);
#pragma unroll
              for (int layer = 0; layer < 2; ++layer) {
                float* R = layer ? S2 : S1;
#pragma unroll
                for (int dyi = 0; dyi < 3; ++dyi) {
                  const float vl = okj ? __uint_as_float(vr_[layer * 9 + 3 * dyi]) : 0.f;
                  const float vc = okj ? __uint_as_float(vr_[layer * 9 + 3 * dyi + 1]) : 0.f;
                  const float vr = okj ? __uint_as_float(vr_[layer * 9 + 3 * dyi + 2]) : 0.f;
                  float left = __shfl_up_sync(0xffffffffu, vl, 1);
                  float right = __shfl_down_sync(0xffffffffu, vr, 1);
                  if (lane == 0) { left = 0.f; eb[-12 + 6 + layer * 3 + dyi] = vr; }       // ER of the previous block
                  if (lane == 31) { right = 0.f; eb[12 + layer * 3 + dyi] = vl; }          // EL of the next block
                  if (okj) R[(ap - (dyi - 1) * P) * 3 + dyi] = (left + vc) + right;
                }
              }
            }
          }
        }
        LVP_MARK(3)
        named_bar_sync(1, kEpi);
        // the small per-step vectors are needed from here on
        mbar_wait_sleep(we_full + (gs & 1), (uint32_t)((gs >> 1) & 1));
        LVP_MARK(4)
        const float inv1 = s_w2d[9], inv2 = s_w2d[10], inv3 = s_w2d[11], gain = s_w2d[12];
        // ---- G2: d1 (zero outside the image: ring rows of D1 are never written)
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          if (OKJ(j)) {
            const int ap = g.M0 + (tg + j * NTG) * 128 + q * 32 + lane;
            float pre;
            if constexpr (!CP) {
              const float* e = S1 + (size_t)ap * 9;
              const float s0 = e[0] + e[1], s1 = e[2] + e[3], s2 = e[4] + e[5], s3 = e[6] + e[7];
              pre = ((s0 + s1) + (s2 + s3)) + e[8];
            } else {
              pre = row_sum3(S1, 0, ap);
            }
            D1[ap] = fmaxf(fmaf(pre, inv1, dcv[j].x), 0.f);
          }
        }
        LVP_MARK(5)
        named_bar_sync(1, kEpi);
        LVP_MARK(4)
        // ---- G3: d2, then the (d1, d2) operand slot
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          if (OKJ(j)) {
            const int ap = g.M0 + (tg + j * NTG) * 128 + q * 32 + lane;
            float sa = 0.f, sb = 0.f, da = dcv[j].y, db = 0.f;
            if constexpr (CP) sa = row_sum3(S2, 1, ap);
#pragma unroll
            for (int t = 0; t < 9; ++t) {
              const int off = (t / 3 - 1) * P + (t % 3 - 1);
              if (t & 1) { if constexpr (!CP) sb += S2[(size_t)ap * 9 + t]; db = fmaf(s_w2d[t], D1[ap + off], db); }
              else { if constexpr (!CP) sa += S2[(size_t)ap * 9 + t]; da = fmaf(s_w2d[t], D1[ap + off], da); }
            }
            const float d2 = fmaxf(fmaf(sa + sb, inv2, da + db), 0.f);
            write_d(D1[ap], d2, ap, flg[j]);
          }
        }
        fence_proxy_async();
        mbar_arrive(d_ready);
        LVP_MARK(6)
        // ---- F: Conv2dZeros output -> coupling, 1x1 mix, ActNorm on the register state; next step's operand planes
        const bool last = is == nsteps - 1;
        // hoisted conditioning part of h (hc): narrow levels prefetch it one tile ahead (the loads of tile j+1 are issued
        // after the coupling of tile j and fly during its mix); wide levels load it per 16 columns
        constexpr bool kPrefHC = C <= 24;
        constexpr int NHC = kPrefHC ? C / 4 : 1;
        float4 hcn[NHC];
        // (tables in the plane-transposed layout of hoist_transpose_kernel: [sample][step][C/4][pixel] float4 -- a warp's
        // 32 pixels read 512 contiguous bytes per load)
        auto hc_ptr = [&](int j) -> const float4* {
          const int pj = OKJ(j) ? pix[j] : 0;
          const int b = (OKJ(j) && a.hoist_bstride) ? SAMPLE(j) : 0;
          return reinterpret_cast<const float4*>(a.hc) + ((size_t)b * a.nsteps_tab + hc_off) * (size_t)(C / 4) * HW + pj;
        };
        if (kPrefHC) {
          const float4* hp0 = hc_ptr(0);
#pragma unroll
          for (int c4 = 0; c4 < NHC; ++c4) hcn[c4] = __ldg(hp0 + (size_t)c4 * HW);
        }
        // One tile of the F phase.  PH = 0: everything; PH = 1 / 2 (tensor-core mix with several tiles per thread): first the
        // coupling and the state rows of ALL tiles, then, per tile, the mix result and the next step's operand planes -- the round
        // trip through the MMA issuer of one tile runs under the coupling of the next.
        auto finish_tile = [&](int j, auto ph_c) {
          constexpr int PH = decltype(ph_c)::value;
          const int mt = tg + j * NTG;
          if (mt < T) {
            const int ap = g.M0 + mt * 128 + q * 32 + lane;
            if constexpr (PH != 2) {
            const float4* hcp = hc_ptr(j);
            // v = [y1 | y2]: registers (and shared memory for the second half in the CP layout)
            float v[C];
#pragma unroll
            for (int c = 0; c < NSR; ++c) v[c] = st[j][c];
            if constexpr (CP) {
              const float2* y2p = reinterpret_cast<const float2*>(Y2S + (size_t)ap * (C / 2));
#pragma unroll
              for (int c2 = 0; c2 < C / 4; ++c2) { const float2 t = y2p[c2]; v[C / 2 + 2 * c2] = t.x; v[C / 2 + 2 * c2 + 1] = t.y; }
            }
            LVP_MARK(7)
            mbar_wait_sleep(z_full + mt, par);
            LVP_MARK(8)
            tc_fence_after();
            float ldsum = 0.f;
#pragma unroll
            for (int n0 = 0; n0 < NP; n0 += 16) {
              float h[16];
              if constexpr (CP && C == 12) {           // 12 of the 16 accumulator columns carry channels
                uint32_t hr[12];
                tmem_ld_nowait<8>(tmem_base + lane_base + (uint32_t)(mt * NP), hr);
                tmem_ld_nowait<4>(tmem_base + lane_base + (uint32_t)(mt * NP + 8), hr + 8);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) h[i] = i < 12 ? __uint_as_float(hr[i < 12 ? i : 0]) : 0.f;
              } else {
                tmem_ld16(tmem_base + lane_base + (uint32_t)(mt * NP + n0), h);
              }
#pragma unroll
              for (int e4 = 0; e4 < 16; e4 += 4) {
                if (n0 + e4 < C) {
                  const float4 hc4 = kPrefHC ? hcn[(n0 + e4) / 4 < NHC ? (n0 + e4) / 4 : 0] : __ldg(hcp + (size_t)((n0 + e4) / 4) * HW);
                  const float hv[4] = {fmaf(h[e4], inv3, hc4.x), fmaf(h[e4 + 1], inv3, hc4.y), fmaf(h[e4 + 2], inv3, hc4.z),
                                       fmaf(h[e4 + 3], inv3, hc4.w)};
#pragma unroll
                  for (int e = 0; e < 4; e += 2) {
                    const int n = n0 + e4 + e;
                    const float shift = (hv[e] + s_b3[n]) * gain;                 // h[:, 0::2]
                    const float raw = (hv[e + 1] + s_b3[n + 1]) * gain;           // h[:, 1::2]
                    const float la = 2.f * __fdividef(raw, 1.f + fabsf(raw));     // 2*softsign
                    ldsum += la;
                    const int jj = C / 2 + n / 2;
                    v[jj] = fmaf(v[jj], __expf(-la), -shift);                     // y2 / exp(a) - shift
                  }
                }
              }
            }
            tc_fence_before();
            if (kPrefHC && j + 1 < NT && tg + (j + 1) * NTG < T) {               // next tile's hc: in flight during the mix
              const float4* hpn = hc_ptr(j + 1 < NT ? j + 1 : j);
#pragma unroll
              for (int c4 = 0; c4 < NHC; ++c4) hcn[c4] = __ldg(hpn + (size_t)c4 * HW);
            }
            if constexpr (MX) {
              if (OKJ(j)) ldacc[j] += ldsum;
              // u = W v on the tensor cores: this thread's row of the A operand (fp16 hi / lo, 8 channels per 16-byte unit)
              const int row = q * 32 + lane;
              if (mt >= g.NMX) mbar_wait_sleep(mx_full + (mt - g.NMX), par);       // the MMAs that read this buffer last are done
              uint8_t* mb = MXA + (size_t)(mt % g.NMX) * g.mxa_bytes + (size_t)row * 16;
              const bool okj = OKJ(j);
#pragma unroll
              for (int pl = 0; pl < C / 8; ++pl) {
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) lv_split2s(okj ? v[8 * pl + 2 * e] : 0.f, okj ? v[8 * pl + 2 * e + 1] : 0.f, hi[e], lo[e], ovf);
                *reinterpret_cast<uint4*>(mb + (size_t)pl * 2048) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                if (X3) *reinterpret_cast<uint4*>(mb + (size_t)(PM + pl) * 2048) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
              }
              fence_proxy_async();
              mbar_arrive(mx_ready + mt);
            }
            if constexpr (!MX) {
            if (OKJ(j)) {
              ldacc[j] += ldsum;
              // u = W v, out = (u - nb) / nw     (glowConv.py:219, actNorm.py:82)
              // (W is staged TRANSPOSED, Wt[k][r]: four consecutive outputs of one input channel per 16-byte load, two
              // outputs per packed FMA; C/2 independent accumulation chains)
              float2 acc2[C / 2];
#pragma unroll
              for (int r2 = 0; r2 < C / 2; ++r2) acc2[r2] = make_float2(0.f, 0.f);
#pragma unroll
              for (int k = 0; k < C; ++k) {
                const float4* w4 = reinterpret_cast<const float4*>(sm + k * C);
                const float vk = v[k];
#pragma unroll
                for (int r4 = 0; r4 < C / 4; ++r4) {
                  const float4 w = w4[r4];
                  acc2[2 * r4] = lv_ffma2(make_float2(w.x, w.y), vk, acc2[2 * r4]);
                  acc2[2 * r4 + 1] = lv_ffma2(make_float2(w.z, w.w), vk, acc2[2 * r4 + 1]);
                }
              }
              float o[C];
#pragma unroll
              for (int r2 = 0; r2 < C / 2; ++r2) {
                o[2 * r2] = (acc2[r2].x - s_nb[2 * r2]) * s_rnw[2 * r2];
                o[2 * r2 + 1] = (acc2[r2].y - s_nb[2 * r2 + 1]) * s_rnw[2 * r2 + 1];
              }
#pragma unroll
              for (int r = 0; r < NSR; ++r) st[j][r] = o[r];
              if constexpr (CP) {
                float2* y2p = reinterpret_cast<float2*>(Y2S + (size_t)ap * (C / 2));
#pragma unroll
                for (int c2 = 0; c2 < C / 4; ++c2) y2p[c2] = make_float2(o[C / 2 + 2 * c2], o[C / 2 + 2 * c2 + 1]);
              }
            }
            }
            }   // PH != 2
            if constexpr (PH != 1) {
            if constexpr (MX) {
              const bool okj = OKJ(j);
              mbar_wait_sleep(mx_full + mt, par);
              tc_fence_after();
              const float invw = s_w2d[13];
#pragma unroll
              for (int n0 = 0; n0 < NP; n0 += 16) {
                float u[16];
                tmem_ld16(tmem_base + lane_base + (uint32_t)(mt * NP + n0), u);
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                  const int r = n0 + e;
                  if (r < C) {
                    const float o = (u[e] * invw - s_nb[r]) * s_rnw[r];      // out = (u - nb) / nw   (glowConv.py:219, actNorm.py:82)
                    if (okj) st[j][r < NSR ? r : 0] = o;
                  }
                }
              }
              tc_fence_before();
            }
            // the Z MMAs of the following tiles read this tile's positions as neighbours: wait for them before overwriting
            LVP_MARK(9)
            for (int k = mt + 1; k <= mt + g.Dt && k < T; ++k) mbar_wait_sleep(z_full + k, par);
            LVP_MARK(10)
            if (!last) {
              if (OKJ(j)) write_planes(st[j], ap, flg[j]);
              fence_proxy_async();
              mbar_arrive(a_ready + mt);
            }
            }   // PH != 1
          }
        };
        if constexpr (MX && NT > 1) {
#pragma unroll
          for (int j = 0; j < NT; ++j) finish_tile(j, std::integral_constant<int, 1>{});
#pragma unroll
          for (int j = 0; j < NT; ++j) finish_tile(j, std::integral_constant<int, 2>{});
        } else {
#pragma unroll
          for (int j = 0; j < NT; ++j) finish_tile(j, std::integral_constant<int, 0>{});
        }
        mbar_arrive(we_free + (gs & 1));           // this thread is done with the step's small vectors
      }
      // ---- end of pass: state back to HBM, per-sample log-det
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        if constexpr (!CP) {
          if (OKJ(j) && g.unsq != nullptr) {
            // channel group k of this pixel is pixel (2y + dr, 2x + dc) of the next level: (0,0),(1,0),(1,1),(0,1) (flowUtils.py:117-120)
            constexpr int C4 = C / 4;
            const int b = SAMPLE(j), py = pix[j] / a.W, px = pix[j] - py * a.W;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int dr = (k == 1 || k == 2) ? 1 : 0, dc = k >= 2 ? 1 : 0;
              float* d = g.unsq + (((size_t)b * (2 * a.H) + 2 * py + dr) * (size_t)(2 * a.W) + 2 * px + dc) * g.unsq_cstride;
              if constexpr (C4 % 4 == 0) {
#pragma unroll
                for (int c = 0; c < C4; c += 4) *reinterpret_cast<float4*>(d + c) = make_float4(st[j][k * C4 + c], st[j][k * C4 + c + 1], st[j][k * C4 + c + 2], st[j][k * C4 + c + 3]);
              } else {
#pragma unroll
                for (int c = 0; c < C4; c += 2) *reinterpret_cast<float2*>(d + c) = make_float2(st[j][k * C4 + c], st[j][k * C4 + c + 1]);
              }
            }
            D1[g.M0 + (tg + j * NTG) * 128 + q * 32 + lane] = ldacc[j];
            continue;
          }
        }
        if (OKJ(j)) {
          const int b = SAMPLE(j);
          float4* y4 = reinterpret_cast<float4*>(a.y_out + ((size_t)b * HW + pix[j]) * C);
#pragma unroll
          for (int c4 = 0; c4 < C / 4; ++c4) {
            float tv[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int c = 4 * c4 + e;
              tv[e] = c < NSR ? st[j][c < NSR ? c : 0] : Y2S[(size_t)(g.M0 + (tg + j * NTG) * 128 + q * 32 + lane) * (C / 2) + (c - C / 2)];
            }
            y4[c4] = make_float4(tv[0], tv[1], tv[2], tv[3]);
          }
          D1[g.M0 + (tg + j * NTG) * 128 + q * 32 + lane] = ldacc[j];
        }
      }
      named_bar_sync(1, kEpi);
      if (a.ld_part) {
        for (int gi = warp; gi < g.G; gi += NE) {
          const int b = pass * g.G + gi;
          if (b < a.B) {
            double s = 0.0;
            const float* d = D1 + g.M0 + gi * g.npos_s;
            for (int i = lane; i < g.npos_s; i += 32) s += (double)d[i];      // ring entries are zero
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) a.ld_part[(size_t)b * a.ld_stride] = (float)s;
          }
        }
      }
      named_bar_sync(1, kEpi);
    }
    LVP_MARK(11)
    if (tid == 0) { LVP_FLUSH(0) }
    if (ovf && a.overflow) atomicOr(a.overflow, 1u);
  } else if (warp == NE) {
    // =========================================================== weight loader
    int gs = 0;
    for (int ip = 0; ip < npass_mine; ++ip) {
      for (int is = 0; is < nsteps; ++is, ++gs) {
        const LevelStep* S = a.steps + is;
        const int stg = gs & 1;
        if (gs >= 2) mbar_wait_sleep(we_free + stg, (uint32_t)(((gs >> 1) - 1) & 1));
        float* sm = reinterpret_cast<float*>(SM + (size_t)stg * g.sm_stage);
        [[maybe_unused]] const int64_t oW = S->W; const int64_t oNw = S->nw, oNb = S->nb, oB = S->bias, oM = S->misc, oG = S->gain;
        if constexpr (!MX) {
          for (int i = lane; i < CC; i += 32) sm[(i % C) * C + i / C] = __ldg(a.packed + oW + i);      // Wt[k][r] = W[r][k]
        }
        for (int i = lane; i < C; i += 32) {
          const float nw = oNw >= 0 ? __ldg(a.params + oNw + i) : 1.f;
          sm[CCs + i] = nw;
          sm[CCs + C + i] = oNw >= 0 ? __ldg(a.params + oNb + i) : 0.f;
          sm[CCs + 2 * C + i] = 1.f / nw;
          sm[CCs + 3 * C + i] = __ldg(a.params + oB + i);
        }
        if (lane < 12) sm[CCs + 4 * C + lane] = __ldg(a.packed + oM + lane);
        if (lane == 12) sm[CCs + 4 * C + 12] = __ldg(a.packed + oG);
        const int64_t oX = MX ? __ldg(g.wmx + is) : 0;
        if (MX && lane == 13) sm[CCs + 4 * C + 13] = __ldg(a.packed + oX);            // 1 / (power-of-two scale of the mix operand)
        __syncwarp();
        // NOTE: the copies are issued from rolled loops with a running source pointer.  With the (tap, hl) loops unrolled,
        // ptxas 12.9 formed the 64-bit source of some UBLKCP from a stale upper register (the mbarrier address): an
        // out-of-bounds global read on every launch (compute-sanitizer), cured by any perturbation of the code.
        if (lane == 0) {
          const uint8_t* src = reinterpret_cast<const uint8_t*>(a.packed + (CP ? S->wEc : S->wE));
          uint8_t* dst = WE + (size_t)stg * g.wE_stage;
          mbar_expect_tx(we_full + stg, g.cpE_n * g.cpE_bytes + (MX ? g.mxw_stage : 0u));
#pragma unroll 1
          for (uint32_t i = 0; i < g.cpE_n; ++i, src += g.cpE_sstep, dst += g.cpE_bytes) bulk_g2s(dst, src, g.cpE_bytes, we_full + stg);
          if constexpr (MX)      // mix operand: hi (and lo) planes are contiguous behind the 16-byte header of the block
            bulk_g2s(MXW + (size_t)stg * g.mxw_stage, reinterpret_cast<const uint8_t*>(a.packed + oX + 4), g.mxw_stage, we_full + stg);
        }
        const int zs = gs % g.nzs, zuse = gs / g.nzs;
        if (zuse >= 1) mbar_wait_sleep(wz_free + zs, (uint32_t)((zuse - 1) & 1));
        if (lane == 0) {
          // non-CP: global [tap][hl (always 2)][all planes][NP][16 B] -> shared [tap][hl (1 or 2)][NPL planes][NP][16 B]
          // CP:     global [pair][hl (always 2)][2 planes][NP][16 B]  -> shared [pair][hl (1 or 2)][2 planes][NP][16 B]
          const uint8_t* src = reinterpret_cast<const uint8_t*>(a.packed + (CP ? S->wZc : S->wZ));
          uint8_t* dst = WZ + (size_t)zs * g.wZ_stage;
          mbar_expect_tx(wz_full + zs, g.cpZ_n * g.cpZ_bytes);
#pragma unroll 1
          for (uint32_t i = 0; i < g.cpZ_n; ++i, src += g.cpZ_sstep, dst += g.cpZ_bytes) bulk_g2s(dst, src, g.cpZ_bytes, wz_full + zs);
        }
      }
    }
  } else if (warp == NE + 1) {
    // =========================================================== E issuer: one elected lane
    if (elect_one()) {
      const uint32_t idE = lv_idesc(32);
      const uint64_t hlA16 = g.hlA >> 4, wEhl16 = g.wE_hl >> 4;
      int gs = 0;
      LVP_DECL
      for (int ip = 0; ip < npass_mine; ++ip) {
        for (int is = 0; is < nsteps; ++is, ++gs) {
          const int stg = gs & 1;
          LVP_MARK(0)
          mbar_wait_sleep(we_full + stg, (uint32_t)((gs >> 1) & 1));
          LVP_MARK(1)
          const uint64_t bE0 = make_desc(smem_u32(WE + (size_t)stg * g.wE_stage), 512, 128);
          for (int mt = 0; mt < T; ++mt) {
            const int ce = gs * T + mt, slot = ce % g.NES, use = ce / g.NES;
            LVP_MARK(2)
            mbar_wait_sleep(a_ready + mt, (uint32_t)(gs & 1));
            LVP_MARK(3)
            if (use >= 1) mbar_wait_sleep(e_free + slot, (uint32_t)((use - 1) & 1));
            LVP_MARK(4)
            tc_fence_after();
            const uint32_t tE = tmem_base + (uint32_t)(g.zcols + slot * 32);
            // CP: one operand plane; the second K half (next position's unit, LBO = 16 bytes) meets zero weights
            const uint64_t aE0 = make_desc(smem_u32(A + (size_t)(g.M0 + mt * 128) * 16), CP ? 16u : g.PLB, 128);
#pragma unroll
            for (int ks = 0; ks < KSy; ++ks) {
              const uint64_t ad = aE0 + (uint64_t)ks * (uint64_t)((2 * g.PLB) >> 4);
              const uint64_t bd = bE0 + (uint64_t)ks * (uint64_t)(2 * 512 >> 4);
              lv_mma(tE, ad, bd, idE, ks > 0 ? 1u : 0u);
              if (X3) {
                lv_mma(tE, ad + hlA16, bd, idE, 1u);
                lv_mma(tE, ad, bd + wEhl16, idE, 1u);
              }
            }
            mma_commit(e_full + slot);
          }
          if constexpr (MX) {
            // the mixes of this step, tile by tile as their state rows arrive (the next step's a_ready follows each of them)
            const uint32_t idM = lv_idesc(NP);
            const uint64_t loA16 = (uint64_t)(PM * 2048) >> 4, loB16 = (uint64_t)(PM * NP * 16) >> 4;
            const uint64_t bW0 = make_desc(smem_u32(MXW + (size_t)stg * g.mxw_stage), (uint32_t)NP * 16u, 128);
            for (int mt = 0; mt < T; ++mt) {
              mbar_wait_sleep(mx_ready + mt, (uint32_t)(gs & 1));
              tc_fence_after();
              const uint32_t tM = tmem_base + (uint32_t)(mt * NP);
              const uint64_t aM0 = make_desc(smem_u32(MXA + (size_t)(mt % g.NMX) * g.mxa_bytes), 2048, 128);
#pragma unroll
              for (int ks = 0; ks < PM / 2; ++ks) {
                const uint64_t ad = aM0 + (uint64_t)ks * (uint64_t)((2 * 2048) >> 4);
                const uint64_t bd = bW0 + (uint64_t)ks * (uint64_t)(2 * NP);
                lv_mma(tM, ad, bd, idM, ks > 0 ? 1u : 0u);
                if (X3) {
                  lv_mma(tM, ad + loA16, bd, idM, 1u);
                  lv_mma(tM, ad, bd + loB16, idM, 1u);
                }
              }
              mma_commit(mx_full + mt);
            }
          }
          mma_commit(we_free + stg);
        }
      }
      LVP_MARK(2)
      LVP_FLUSH(12)
    }
  } else {
    // =========================================================== Z issuers (even / odd tiles): one elected lane each
    const int role = warp - (NE + 2);
    if (elect_one()) {
      const uint32_t idZ = lv_idesc(NP);
      const uint64_t hlA16 = g.hlA >> 4, wZhl16 = g.wZ_hl >> 4, wZtap16 = g.wZ_tap >> 4;
      int gs = 0;
      LVP_DECL
      for (int ip = 0; ip < npass_mine; ++ip) {
        for (int is = 0; is < nsteps; ++is, ++gs) {
          const int zs = gs % g.nzs, zuse = gs / g.nzs;
          LVP_MARK(0)
          mbar_wait_sleep(wz_full + zs, (uint32_t)(zuse & 1));
          LVP_MARK(1)
          mbar_wait_sleep(d_ready, (uint32_t)(gs & 1));
          LVP_MARK(2)
          tc_fence_after();
          const uint64_t bZ0 = make_desc(smem_u32(WZ + (size_t)zs * g.wZ_stage), (uint32_t)NP * 16u, 128);
          for (int mt = role; mt < T; mt += 2) {
            const uint32_t tZ = tmem_base + (uint32_t)(mt * NP);
            if constexpr (CP) {
              const uint32_t a0 = smem_u32(A + (size_t)(g.M0 + mt * 128) * 16);
#pragma unroll
              for (int pr = 0; pr < 5; ++pr) {
                const int ta = 2 * pr, tb = 2 * pr + 1;
                const int offa = (ta / 3 - 1) * P + (ta % 3 - 1);
                const int offb = tb < 9 ? (tb / 3 - 1) * P + (tb % 3 - 1) : offa + 1;     // pair 4: single tap, zero weights
                const uint64_t ad = make_desc(a0 + (uint32_t)(offa * 16), (uint32_t)((offb - offa) * 16), 128);
                const uint64_t bd = bZ0 + (uint64_t)pr * wZtap16;
                lv_mma(tZ, ad, bd, idZ, pr > 0 ? 1u : 0u);
                if (X3) {
                  lv_mma(tZ, ad + hlA16, bd, idZ, 1u);
                  lv_mma(tZ, ad, bd + wZhl16, idZ, 1u);
                }
              }
            } else {
            const uint64_t aZ0 = make_desc(smem_u32(A + (size_t)(g.M0 + mt * 128) * 16), g.PLB, 128);
#pragma unroll
            for (int ks = 0; ks < KSy; ++ks) {
              const uint64_t ak = aZ0 + (uint64_t)ks * (uint64_t)((2 * g.PLB) >> 4);
              const uint64_t bk = bZ0 + (uint64_t)ks * (uint64_t)(2 * NP);
#pragma unroll
              for (int tap = 0; tap < 9; ++tap) {
                const int off = (tap / 3 - 1) * P + (tap % 3 - 1);             // positions = 16-byte units
                const uint64_t ad = ak + (uint64_t)(int64_t)off;
                const uint64_t bd = bk + (uint64_t)tap * wZtap16;
                lv_mma(tZ, ad, bd, idZ, (tap > 0 || ks > 0) ? 1u : 0u);
                if (X3) {
                  lv_mma(tZ, ad + hlA16, bd, idZ, 1u);
                  lv_mma(tZ, ad, bd + wZhl16, idZ, 1u);
                }
              }
            }
            }
            mma_commit(z_full + mt);
          }
          mma_commit(wz_free + zs);
        }
      }
      LVP_MARK(0)
      if (role == 0) { LVP_FLUSH(24) }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == NE + 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ------------------------------------------------------------------ host side
// epilogue warp groups (4 warps own a tile) and tiles per group: the state of NT pixels lives in one thread's registers
template <int C> struct LvCfg;
// (C = 12, measured at S = 4096: 5 x 4 -> 14.5 ms for the three level launches, 4 x 5 (96 registers) 15.2 ms, 3 x 6 (128) 15.3 ms)
template <> struct LvCfg<12> { static constexpr int NTG = 5, NT = 4; static constexpr bool CP = true; };
template <> struct LvCfg<24> { static constexpr int NTG = 3, NT = 2; static constexpr bool CP = false; };
template <> struct LvCfg<48> { static constexpr int NTG = 3, NT = 1; static constexpr bool CP = false; };

static bool lv_geom(const LevelArgs& a, LevelGeom& g, int ntg, int nt, bool cp, bool mx) {
  const int C = a.C;
  const int NP = (C + 15) / 16 * 16;
  const int KSy = cp ? 1 : (C / 2 + 2 + 15) / 16, KS1 = (a.nch1 + 15) / 16;
  const int NPL = cp ? 1 : 2 * KSy, PLtot = 2 * (KSy + KS1);
  const int nhl = a.x3 ? 2 : 1;
  if (a.H < 1 || a.W < 1 || a.B < 1 || a.nsteps < 1) return false;
  if (cp && (C / 2 + 2 > 8 || !a.compact)) return false;
  g.P = a.W + 2;
  g.npos_s = (a.H + 2) * g.P;
  g.M0 = (g.P + 2 + 7) / 8 * 8;
  g.Dt = (g.P + 2 + 127) / 128;
  if (cp) {
    g.gE_hl = 2 * 512; g.gZ_hl = 2u * NP * 16; g.gZ_tap = 2 * g.gZ_hl;
    g.wE_hl = 2 * 512; g.wZ_hl = 2u * NP * 16; g.wZ_tap = nhl * g.wZ_hl;
    g.wE_stage = nhl * g.wE_hl; g.wZ_stage = 5u * g.wZ_tap;
    g.cpZ_n = 5u * nhl;
  } else {
    g.gE_hl = (uint32_t)PLtot * 512; g.gZ_hl = (uint32_t)PLtot * NP * 16; g.gZ_tap = 2 * g.gZ_hl;
    g.wE_hl = (uint32_t)NPL * 512; g.wZ_hl = (uint32_t)NPL * NP * 16; g.wZ_tap = nhl * g.wZ_hl;
    g.wE_stage = nhl * g.wE_hl; g.wZ_stage = 9u * g.wZ_tap;
    g.cpZ_n = 9u * nhl;
  }
  g.cpE_n = nhl; g.cpE_bytes = g.wE_hl; g.cpE_sstep = g.gE_hl;
  g.cpZ_bytes = g.wZ_hl; g.cpZ_sstep = a.x3 ? g.gZ_hl : g.gZ_tap;
  const int PM = (C / 8 + 1) / 2 * 2;
  g.sm_stage = (uint32_t)(((mx ? 0 : C * C) + 4 * C + 16) * 4 + 127) / 128 * 128;
  g.mxa_bytes = mx ? (uint32_t)nhl * PM * 2048u : 0u;
  g.mxw_stage = mx ? (uint32_t)nhl * PM * NP * 16u : 0u;
  const int maxT = std::min(ntg * nt, kLvMaxTiles);
  double best = -1.0;
  LevelGeom bg{};
  const int gmax = std::min(a.B, 64);
  for (int G = 1; G <= gmax; ++G) {
    const int npos = G * g.npos_s;
    const int T = (npos + 127) / 128;
    if (T > maxT) break;
    const int zcols = T * NP;
    int NES = std::min(T, (512 - zcols) / 32);
    if (NES < std::min(T, 2)) break;
    LevelGeom c = g;
    c.G = G; c.T = T; c.NES = NES; c.zcols = zcols;
    c.NA = (c.M0 + T * 128 + c.M0 + 7) / 8 * 8;
    c.PLB = (uint32_t)c.NA * 16; c.hlA = (uint32_t)NPL * c.PLB;
    bool placed = false;
    for (int nmx = mx ? std::min(T, 3) : 0; nmx >= (mx ? 1 : 0) && !placed; --nmx) {
    for (int nzs = 2; nzs >= 1; --nzs) {
      uint32_t off = 0;
      auto take = [&](uint32_t n) { uint32_t o = off; off += (n + 127) / 128 * 128; return o; };
      c.nzs = nzs; c.NMX = nmx;
      c.oA = take((uint32_t)nhl * c.hlA);
      c.oS1 = take((uint32_t)c.NA * (cp ? 12 : 36));
      c.oS2 = take((uint32_t)c.NA * (cp ? 12 : 36));
      c.oD1 = take((uint32_t)c.NA * 4);
      c.oEdge = take(cp ? (uint32_t)(T * 4 + 2) * 48 : 16u);
      c.oY2 = take(cp ? (uint32_t)c.NA * (uint32_t)(C / 2) * 4 : 16u);
      c.oMXA = take(mx ? (uint32_t)nmx * c.mxa_bytes : 16u);       // inside the range zeroed at kernel start (padding K plane)
      c.oWE = take(2u * c.wE_stage);
      c.oWZ = take((uint32_t)nzs * c.wZ_stage);
      c.oSm = take(2u * c.sm_stage);
      c.oMXW = take(mx ? 2u * c.mxw_stage : 16u);
      c.oBar = take((10 + 6 * kLvMaxTiles) * 8 + 16);
      c.total = off;
      if (c.total <= 227u * 1024u) {
        // fill of the M tiles by real pixels; larger groups amortise the per-step weight traffic (tie-break), two weight
        // stages for the Conv2dZeros weights hide their load
        const double fill = (double)G * a.H * a.W / (T * 128.0) + 0.02 * (nzs - 1) + 0.01 * nmx + 1e-4 * G;
        if (fill > best) { best = fill; bg = c; }
        placed = true;
        break;
      }
    }
    }
  }
  if (best < 0.0) return false;
  g = bg;
  g.npass = (a.B + g.G - 1) / g.G;
  return true;
}

static int lv_sm_count() {
  int dev = 0, nsm = 148;
  cudaGetDevice(&dev);
  static int cached[64] = {0};
  if (dev >= 0 && dev < 64) {
    if (!cached[dev]) cudaDeviceGetAttribute(&cached[dev], cudaDevAttrMultiProcessorCount, dev);
    if (cached[dev] > 0) nsm = cached[dev];
  }
  return nsm;
}

bool level_resident_supported(const LevelArgs& a, bool mix_mma) {
  LevelGeom g{};
  switch (a.C) {
    case 12: return lv_geom(a, g, LvCfg<12>::NTG, LvCfg<12>::NT, LvCfg<12>::CP, false);
    case 24: return lv_geom(a, g, LvCfg<24>::NTG, LvCfg<24>::NT, LvCfg<24>::CP, mix_mma);
    case 48: return lv_geom(a, g, LvCfg<48>::NTG, LvCfg<48>::NT, LvCfg<48>::CP, mix_mma);
    default: return false;
  }
}

template <int C>
static int lv_launch(const LevelArgs& a_in, const int64_t* wmx, float* unsq, int unsq_cstride, cudaStream_t st) {
  constexpr int NTG = LvCfg<C>::NTG, NT = LvCfg<C>::NT;
  constexpr bool CP = LvCfg<C>::CP;
  LevelArgs a = a_in;
  static const bool prof_env = getenv("TMG_LV_PROF") != nullptr;
  static long long* prof_buf = nullptr;
  static int prof_left = 0;
  if (prof_env && !prof_buf) { cudaMalloc(&prof_buf, 148 * 48 * sizeof(long long)); prof_left = atoi(getenv("TMG_LV_PROF")); }
  if (prof_env && prof_left > 0) { cudaMemsetAsync(prof_buf, 0, 148 * 48 * sizeof(long long), st); a.prof = prof_buf; }
  LevelGeom g{};
  const bool mx = !CP && wmx != nullptr;
  if (!lv_geom(a, g, NTG, NT, CP, mx)) { set_error("level-resident flow kernel: unsupported shape (C=%d, %dx%d)", a.C, a.H, a.W); return TMG_ERR_UNSUPPORTED; }
  g.wmx = wmx;
  g.unsq = CP ? nullptr : unsq; g.unsq_cstride = unsq_cstride;
  if (unsq && (CP || (unsq_cstride & 1) || (reinterpret_cast<uintptr_t>(unsq) & 15) || ((a.C / 4) % 4 == 0 && (unsq_cstride & 3)))) {
    set_error("level-resident flow kernel: un-squeezed output not supported for C=%d, stride %d", a.C, unsq_cstride);
    return TMG_ERR_UNSUPPORTED;
  }
  const int grid = std::min(g.npass, lv_sm_count());
#define TMG_LVK(XX, MM)                                                                                  \
  {                                                                                                      \
    TMG_SMEM_ATTR((flow_level_kernel<C, XX, NTG, NT, CP, MM>), 227 * 1024);                              \
    flow_level_kernel<C, XX, NTG, NT, CP, MM><<<grid, (4 * NTG + 4) * 32, g.total, st>>>(a, g);          \
  }
  if constexpr (CP) {
    if (a.x3) TMG_LVK(true, false) else TMG_LVK(false, false)
  } else {
    if (mx) { if (a.x3) TMG_LVK(true, true) else TMG_LVK(false, true) }
    else { if (a.x3) TMG_LVK(true, false) else TMG_LVK(false, false) }
  }
#undef TMG_LVK
  TMG_LAUNCH_CHECK();
  if (a.prof) {      // developer profiling: average cycles per role and phase over the CTAs of this launch, per step
    --prof_left;
    static long long h[148 * 48];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, prof_buf, sizeof(h), cudaMemcpyDeviceToHost);
    double s[48] = {0};
    for (int b = 0; b < grid; ++b) for (int i = 0; i < 48; ++i) s[i] += (double)h[b * 48 + i] / grid;
    const double ns = (double)g.npass / grid * a.nsteps;      // steps per CTA
    fprintf(stderr, "[level prof] C=%d x3=%d %dx%d B=%d G=%d T=%d NES=%d nzs=%d smem=%u passes/CTA=%.1f | cycles per step: "
            "EPI load %.0f pre %.0f wait_e %.0f scatter %.0f bar+G2/G3 %.0f (G2 %.0f) dslot %.0f preF %.0f wait_z %.0f finish %.0f wait_nb %.0f tail %.0f | "
            "E: gap %.0f wait_w %.0f wait_a %.0f wait_efree %.0f issue+loop %.0f | Z: gap %.0f wait_w %.0f wait_d %.0f issue %.0f\n",
            a.C, a.x3, a.H, a.W, a.B, g.G, g.T, g.NES, g.nzs, g.total, (double)g.npass / grid,
            s[0] / ns, s[1] / ns, s[2] / ns, s[3] / ns, s[4] / ns, s[5] / ns, s[6] / ns, s[7] / ns, s[8] / ns, s[9] / ns, s[10] / ns, s[11] / ns,
            s[12] / ns, s[13] / ns, s[15] / ns, s[16] / ns, s[14] / ns, s[24] / ns, s[25] / ns, s[26] / ns, 0.0);
  }
  return TMG_OK;
}

// Hoisted conditioning tables, re-laid out so that the 32 pixels of a warp read contiguous memory: in the layout run_hoist
// writes ([pixel][all steps]) a step's entries of neighbouring pixels are 128 B (dc) / 768 B (hc) apart -- every lane of a load
// touched its own cache line (measured: ~10 000 L1 wavefront cycles per level-0 step).
__global__ void hoist_transpose_kernel(const float* __restrict__ dc_all, int dstride, const float* __restrict__ hc_all, int hstride,
                                       float* __restrict__ dcT, float* __restrict__ hcT, int HW, int nsteps, int C) {
  extern __shared__ float tile[];                    // [32 pixels][ncol + 1]
  const int b = blockIdx.y, p0 = blockIdx.x * 32;
  const int np = min(32, HW - p0);
  {
    const int ncol = 2 * nsteps, ld = ncol + 1;
    for (int i = threadIdx.x; i < np * ncol; i += blockDim.x) {
      const int pp = i / ncol, c = i - pp * ncol;
      tile[pp * ld + c] = dc_all[((size_t)b * HW + p0 + pp) * dstride + c];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nsteps * np; i += blockDim.x) {
      const int s = i / np, pp = i - s * np;
      reinterpret_cast<float2*>(dcT)[((size_t)b * nsteps + s) * HW + p0 + pp] = make_float2(tile[pp * ld + 2 * s], tile[pp * ld + 2 * s + 1]);
    }
    __syncthreads();
  }
  {
    const int ncol = nsteps * C, ld = ncol + 1, c4n = C / 4;
    for (int i = threadIdx.x; i < np * ncol; i += blockDim.x) {
      const int pp = i / ncol, c = i - pp * ncol;
      tile[pp * ld + c] = hc_all[((size_t)b * HW + p0 + pp) * hstride + c];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nsteps * c4n * np; i += blockDim.x) {
      const int pp = i % np; const int t = i / np; const int c4 = t % c4n, s = t / c4n;
      const float* src = tile + pp * ld + s * C + 4 * c4;
      reinterpret_cast<float4*>(hcT)[(((size_t)b * nsteps + s) * c4n + c4) * HW + p0 + pp] = make_float4(src[0], src[1], src[2], src[3]);
    }
  }
}

int launch_hoist_transpose(const float* dc_all, int dstride, const float* hc_all, int hstride, float* dcT, float* hcT,
                           int Bx, int HW, int nsteps, int C, cudaStream_t st) {
  const size_t smem = (size_t)32 * (nsteps * C + 1) * sizeof(float);
  if (smem > 200 * 1024) { set_error("hoist transpose: %d steps x %d channels do not fit", nsteps, C); return TMG_ERR_UNSUPPORTED; }
  TMG_SMEM_ATTR(hoist_transpose_kernel, (int)smem);
  hoist_transpose_kernel<<<dim3((unsigned)cdiv(HW, 32), (unsigned)Bx), 256, smem, st>>>(dc_all, dstride, hc_all, hstride, dcT, hcT, HW, nsteps, C);
  TMG_LAUNCH_CHECK();
  return TMG_OK;
}

int launch_level_resident(const LevelArgs& a, const int64_t* wmx, float* unsq, int unsq_cstride, cudaStream_t st) {
  switch (a.C) {
    case 12: return lv_launch<12>(a, nullptr, nullptr, 0, st);
    case 24: return lv_launch<24>(a, wmx, unsq, unsq_cstride, st);
    case 48: return lv_launch<48>(a, wmx, unsq, unsq_cstride, st);
    default: set_error("level-resident flow kernel: %d channels not supported", a.C); return TMG_ERR_UNSUPPORTED;
  }
}

}  // namespace tmg
