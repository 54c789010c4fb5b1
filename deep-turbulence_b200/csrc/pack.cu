// Derives the kernel-side weights from the flat reference-layout parameter buffer in ONE launch:
// one CTA per job.  Reference semantics restated:
//   InvertibleConv1x1LU.weight / inv_weight   nn/modules/glowConv.py:151-174
//   log-det constants                         glowConv.py:186 (forward), :206-215 (reverse), actNorm.py:66,82
//   Conv2dZeros gain exp(clamp(scale,-4,ln4)) nn/modules/flowUtils.py:247
//   eval-mode BatchNorm2d folded to an affine nn/modules/denseBlock.py:49
#include <cuda_fp16.h>

#include "common.cuh"

namespace tmg {

__global__ void __launch_bounds__(256)
pack_kernel(const PackJob* jobs, const float* __restrict__ P, float* __restrict__ Q, int cmax) {
  extern __shared__ __align__(16) float sm[];
  const PackJob j = jobs[blockIdx.x];
  const int tid = threadIdx.x;
  if (j.type == JOB_CONVW) {
    // OIHW [O][I][3][3]  ->  tap-major [9][I][opad]  (zero-filled padding columns)
    const int O = j.a, I = j.b, OP = j.opad;
    const float* s = P + j.src[0];
    float* d = Q + j.dst[0];
    for (int i = tid; i < 9 * I * OP; i += blockDim.x) {
      int o = i % OP; int r = i / OP; int c = r % I; int tap = r / I;
      d[i] = o < O ? s[((size_t)o * I + c) * 9 + tap] : 0.f;
    }
  } else if (j.type == JOB_CONVW_TC) {
    // OIHW -> [chunk][tap][hi|lo][plane][npad][4] with the TF32 hi/lo split (conv3x3_tc.cu)
    const int O = j.a, I = j.b, NP = j.opad;
    const float* s = P + j.src[0];
    float* d = Q + j.dst[0];
    const size_t total = tc_packed_floats(I, NP);
    const size_t per = (total + j.nparts - 1) / j.nparts;
    const size_t lo_i = per * j.part, hi_i = lo_i + per < total ? lo_i + per : total;
    for (size_t i = lo_i + tid; i < hi_i; i += blockDim.x) {
      int e = i & 3;
      size_t t = i >> 2;
      int n = t % NP; t /= NP;
      int plane = t & 3; t >>= 2;
      int hl = t & 1; t >>= 1;
      int tap = t % 9;
      int chunk = (int)(t / 9);
      int c = chunk * 16 + plane * 4 + e;
      float v = (n < O && c < I) ? s[((size_t)n * I + c) * 9 + tap] : 0.f;
      float hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
      d[i] = hl ? (v - hi) : hi;
    }
  } else if (j.type == JOB_CPL_W12 || j.type == JOB_CPL_W3) {
    // Weights of the fused coupling kernel (coupling_tc.cu).  K is laid out in 4-channel planes with
    // every source starting on a plane boundary: [src0 planes][src1 planes][d plane][zero planes].
    //   W12: Cout=1 dense layer, taps in N:  [hi|lo][plane][16][4],  value w[0][c][tap n] for n < 9
    //   W3 : Conv2dZeros:                    [tap][hi|lo][plane][npad][4], value w[n][c][tap]
    const int O = j.a, I = j.b, NPL = j.nplanes, NP = j.type == JOB_CPL_W12 ? 16 : j.opad;
    const int p0 = (j.nch0 + 3) / 4, p1 = (j.nch1 + 3) / 4;
    const float* s = P + j.src[0];
    float* d = Q + j.dst[0];
    const size_t total = (size_t)(j.type == JOB_CPL_W12 ? 1 : 9) * 2 * NPL * NP * 4;
    for (size_t i = tid; i < total; i += blockDim.x) {
      int e = i & 3;
      size_t t = i >> 2;
      int n = t % NP; t /= NP;
      int plane = t % NPL; t /= NPL;
      int hl = t & 1; t >>= 1;
      int tap = (int)t;                       // W3 only
      int c = -1;                              // concatenated input channel
      if (plane < p0) { int q = plane * 4 + e; if (q < j.nch0) c = q; }
      else if (plane < p0 + p1) { int q = (plane - p0) * 4 + e; if (q < j.nch1) c = j.nch0 + q; }
      else if (plane == p0 + p1) { if (e < j.nd) c = j.nch0 + j.nch1 + e; }
      float v = 0.f;
      if (c >= 0 && c < I) {
        if (j.type == JOB_CPL_W12) { if (n < 9) v = s[(size_t)c * 9 + n]; }
        else if (n < O) v = s[((size_t)n * I + c) * 9 + tap];
      }
      float hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
      d[i] = hl ? (v - hi) : hi;
    }
  } else if (j.type == JOB_STEP2) {
    // Weights of the fp16 fused flow step (flow_step_f16.cu).  K layout: [src0 | pad | d1 d2][src1 | pad], 16 per
    // K-step, 8 per plane.  Each conv is scaled by a power of two so that max|w| lands in [2^10, 2^11): the hi/lo
    // fp16 pair then carries 22 mantissa bits of every weight within 2^13 of the largest one.
    //   wE [hl][plane][32][8]: col n < 9 -> dense layer 1 tap n, 16 <= n < 25 -> dense layer 2 tap n-16 (t rows only)
    //   wZ [tap][hl][plane][NP][8]: Conv2dZeros, col n < C
    //   wmisc: [0..8] layer 2's d1 row (fp32, applied on CUDA cores), [9..11] inverse scales
    const int C = j.a, I1 = j.b, NP = j.opad, nch0 = j.nch0, nch1 = j.nch1;
    int KSy, KS1, kd;
    KSy = (nch0 + 2 + 15) / 16; KS1 = (nch1 + 15) / 16; kd = KSy * 16 - 2;
    const int PL = 2 * (KSy + KS1);
    const float* w1 = P + j.src[0]; const float* w2 = P + j.src[1]; const float* w3 = P + j.src[2];
    const int n1 = 9 * I1, n2 = 9 * (I1 + 1), n3 = 9 * C * (I1 + 2);
    float m1 = 0.f, m2 = 0.f, m3 = 0.f;
    for (int i = tid; i < n1; i += blockDim.x) m1 = fmaxf(m1, fabsf(w1[i]));
    for (int i = tid; i < n2; i += blockDim.x) m2 = fmaxf(m2, fabsf(w2[i]));
    for (int i = tid; i < n3; i += blockDim.x) m3 = fmaxf(m3, fabsf(w3[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, o));
      m2 = fmaxf(m2, __shfl_xor_sync(0xffffffffu, m2, o));
      m3 = fmaxf(m3, __shfl_xor_sync(0xffffffffu, m3, o));
    }
    if ((tid & 31) == 0) { sm[(tid >> 5) * 3] = m1; sm[(tid >> 5) * 3 + 1] = m2; sm[(tid >> 5) * 3 + 2] = m3; }
    __syncthreads();
    float sc[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      float m = 0.f;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, sm[w * 3 + q]);
      int e = 0;
      if (m > 0.f && m < 3.0e38f) frexpf(m, &e);                // m = f * 2^e, f in [0.5, 1)
      sc[q] = m > 0.f ? ldexpf(1.f, min(max(11 - e, -20), 40)) : 1.f;   // scaled max in [2^10, 2^11)
    }
    auto chan = [&](int k) -> int {          // K index -> concatenated input channel (or -1)
      if (k < KSy * 16) {
        if (k < nch0) return k;
        if (k == kd) return nch0 + nch1;
        if (k == kd + 1) return nch0 + nch1 + 1;
        return -1;
      }
      const int q = k - KSy * 16;
      return q < nch1 ? nch0 + q : -1;
    };
    __half* dE = reinterpret_cast<__half*>(Q + j.dst[0]);
    const int totE = 2 * PL * 32 * 8;
    for (int i = tid; i < totE; i += blockDim.x) {
      const int e = i & 7; int t = i >> 3; const int n = t & 31; t >>= 5; const int plane = t % PL, hl = t / PL;
      const int c = chan(plane * 8 + e);
      float v = 0.f;
      if (c >= 0 && c < I1) {
        if (n < 9) v = w1[(size_t)c * 9 + n] * sc[0];
        else if (n >= 16 && n < 25) v = w2[(size_t)c * 9 + (n - 16)] * sc[1];
      }
      const __half hi = __float2half_rn(v);
      dE[i] = hl ? __float2half_rn(v - __half2float(hi)) : hi;
    }
    __half* dZ = reinterpret_cast<__half*>(Q + j.dst[1]);
    const int totZ = 9 * 2 * PL * NP * 8;
    for (int i = tid; i < totZ; i += blockDim.x) {
      const int e = i & 7; int t = i >> 3; const int n = t % NP; t /= NP; const int plane = t % PL; t /= PL;
      const int hl = t & 1, tap = t >> 1;
      const int c = chan(plane * 8 + e);
      float v = 0.f;
      if (c >= 0 && c < I1 + 2 && n < C) v = w3[((size_t)n * (I1 + 2) + c) * 9 + tap] * sc[2];
      const __half hi = __float2half_rn(v);
      dZ[i] = hl ? __float2half_rn(v - __half2float(hi)) : hi;
    }
    float* dm = Q + j.dst[2];
    if (tid < 9) dm[tid] = w2[(size_t)I1 * 9 + tid];
    if (tid >= 9 && tid < 12) dm[tid] = 1.f / sc[tid - 9];
  } else if (j.type == JOB_STEP2C) {
    // Compact weights of the level-resident flow kernel (flow_level_f16.cu) for narrow levels (C/2 + 2 <= 8): the whole
    // coupling-net input of a position -- x1 channels, d1 (slot 6), d2 (slot 7) -- is ONE 16-byte operand unit, and a
    // K = 16 MMA contracts TWO filter taps at once (its second K half reads another position through the descriptor's
    // leading-dimension offset).  Same power-of-two scales as JOB_STEP2 (its misc block carries their inverses).
    //   wE [hl][2 planes][32][8]: plane 0 = taps of dense layers 1 (cols 0-8) / 2 (cols 9-17), plane 1 = zeros
    //   wZ [pair 0..4][hl][2 planes][NP][8]: plane 0 = tap 2*pair, plane 1 = tap 2*pair + 1 (zeros for pair 4)
    const int C = j.a, I1 = j.b, NP = j.opad, nch0 = j.nch0, nch1 = j.nch1;
    const float* w1 = P + j.src[0]; const float* w2 = P + j.src[1]; const float* w3 = P + j.src[2];
    const int n1 = 9 * I1, n2 = 9 * (I1 + 1), n3 = 9 * C * (I1 + 2);
    float m1 = 0.f, m2 = 0.f, m3 = 0.f;
    for (int i = tid; i < n1; i += blockDim.x) m1 = fmaxf(m1, fabsf(w1[i]));
    for (int i = tid; i < n2; i += blockDim.x) m2 = fmaxf(m2, fabsf(w2[i]));
    for (int i = tid; i < n3; i += blockDim.x) m3 = fmaxf(m3, fabsf(w3[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, o));
      m2 = fmaxf(m2, __shfl_xor_sync(0xffffffffu, m2, o));
      m3 = fmaxf(m3, __shfl_xor_sync(0xffffffffu, m3, o));
    }
    if ((tid & 31) == 0) { sm[(tid >> 5) * 3] = m1; sm[(tid >> 5) * 3 + 1] = m2; sm[(tid >> 5) * 3 + 2] = m3; }
    __syncthreads();
    float sc[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      float m = 0.f;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, sm[w * 3 + q]);
      int e = 0;
      if (m > 0.f && m < 3.0e38f) frexpf(m, &e);
      sc[q] = m > 0.f ? ldexpf(1.f, min(max(11 - e, -20), 40)) : 1.f;
    }
    auto chan = [&](int k) -> int {          // slot inside the 8-channel unit -> concatenated input channel (or -1)
      if (k < nch0) return k;
      if (k == 6) return nch0 + nch1;
      if (k == 7) return nch0 + nch1 + 1;
      return -1;
    };
    __half* dE = reinterpret_cast<__half*>(Q + j.dst[0]);
    for (int i = tid; i < 2 * 2 * 32 * 8; i += blockDim.x) {
      const int e = i & 7; int t = i >> 3; const int n = t & 31; t >>= 5; const int plane = t & 1, hl = t >> 1;
      const int c = plane == 0 ? chan(e) : -1;
      float v = 0.f;
      if (c >= 0 && c < I1) {          // compact columns: 0-8 dense layer 1, 9-17 dense layer 2 (18 TMEM columns are read back)
        if (n < 9) v = w1[(size_t)c * 9 + n] * sc[0];
        else if (n < 18) v = w2[(size_t)c * 9 + (n - 9)] * sc[1];
      }
      const __half hi = __float2half_rn(v);
      dE[i] = hl ? __float2half_rn(v - __half2float(hi)) : hi;
    }
    __half* dZ = reinterpret_cast<__half*>(Q + j.dst[1]);
    for (int i = tid; i < 5 * 2 * 2 * NP * 8; i += blockDim.x) {
      const int e = i & 7; int t = i >> 3; const int n = t % NP; t /= NP; const int plane = t & 1; t >>= 1;
      const int hl = t & 1, pair = t >> 1;
      const int tap = 2 * pair + plane;
      const int c = chan(e);
      float v = 0.f;
      if (tap < 9 && c >= 0 && c < I1 + 2 && n < C) v = w3[((size_t)n * (I1 + 2) + c) * 9 + tap] * sc[2];
      const __half hi = __float2half_rn(v);
      dZ[i] = hl ? __float2half_rn(v - __half2float(hi)) : hi;
    }
  } else if (j.type == JOB_HOIST) {
    // Conditioning-only weight slices of every plain step of one level, concatenated along Cout, tap-major
    // [9][cond][opad] for conv3x3_ffma: the cond contribution to d1/d2 (kind 0: col 2s, 2s+1) or to the
    // Conv2dZeros output (kind 1: col s*C + n) of step s.  a = C, b = cond channels, nch0 = C/2 (first cond row),
    // nd = kind, part = step index, nparts = steps.  One job per step.
    const int C = j.a, CF = j.b, OP = j.opad, s = j.part, row0 = j.nch0;
    float* d = Q + j.dst[0];
    float* d2 = Q + j.dst[1];          // the same slice as OIHW [OP][CF][3][3] (source of the phase-2 fp16 packing)
    if (j.nd == 0) {
      const float* w1 = P + j.src[0]; const float* w2 = P + j.src[1];
      for (int i = tid; i < 9 * CF * 2; i += blockDim.x) {
        const int which = i & 1; int t = i >> 1; const int c = t % CF, tap = t / CF;
        const float v = (which ? w2 : w1)[(size_t)(row0 + c) * 9 + tap];
        d[((size_t)tap * CF + c) * OP + 2 * s + which] = v;
        d2[((size_t)(2 * s + which) * CF + c) * 9 + tap] = v;
      }
    } else {
      const float* w3 = P + j.src[2];
      const int I3 = row0 + CF + 2;
      for (int i = tid; i < 9 * CF * C; i += blockDim.x) {
        const int n = i % C; int t = i / C; const int c = t % CF, tap = t / CF;
        const float v = w3[((size_t)n * I3 + row0 + c) * 9 + tap];
        d[((size_t)tap * CF + c) * OP + s * C + n] = v;
        d2[((size_t)(s * C + n) * CF + c) * 9 + tap] = v;
      }
    }
  } else if (j.type == JOB_CONV_F16) {
    // OIHW -> fp16 [kstep][tap][hl][2 planes][npad][8] for conv3x3_f16.cu.  The input channels are the
    // concatenation of up to three sources (nch0, nch1, nd), each padded to a multiple of 8 in the K layout.
    // Scaled by a power of two so that max|w| lands in [2^10, 2^11); dst[1] receives the inverse scale.
    const int O = j.a, I = j.b, NP = j.opad;
    const int nch[3] = {j.nch0, j.nch1, j.nd};
    int pl0[4]; pl0[0] = 0;
    for (int q = 0; q < 3; ++q) pl0[q + 1] = pl0[q] + (nch[q] + 7) / 8;
    const int KS = (pl0[3] + 1) / 2;
    const float* w = P + j.src[0];
    float m = 0.f;
    for (int i = tid; i < O * I * 9; i += blockDim.x) m = fmaxf(m, fabsf(w[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((tid & 31) == 0) sm[tid >> 5] = m;
    __syncthreads();
    m = 0.f;
    for (int q = 0; q < (int)(blockDim.x >> 5); ++q) m = fmaxf(m, sm[q]);
    int ex = 0;
    if (m > 0.f && m < 3.0e38f) frexpf(m, &ex);
    const float sc = m > 0.f ? ldexpf(1.f, min(max(11 - ex, -20), 40)) : 1.f;
    __half* d = reinterpret_cast<__half*>(Q + j.dst[0]);
    const size_t total = (size_t)KS * 9 * 2 * 2 * NP * 8;
    for (size_t i = tid; i < total; i += blockDim.x) {
      const int e = (int)(i & 7); size_t t = i >> 3; const int n = (int)(t % NP); t /= NP;
      const int lp = (int)(t & 1); t >>= 1; const int hl = (int)(t & 1); t >>= 1;
      const int tap = (int)(t % 9); const int ks = (int)(t / 9);
      const int plane = 2 * ks + lp;
      int c = -1;
      for (int q = 0, base = 0; q < 3; ++q) {
        if (plane >= pl0[q] && plane < pl0[q + 1]) { const int ch = (plane - pl0[q]) * 8 + e; if (ch < nch[q]) c = base + ch; }
        base += nch[q];
        if (q == 0) base += j.part;          // j.part: input channels skipped after source 0 (conditioning hoisted out of the conv)
      }
      float v = 0.f;
      if (c >= 0 && c < I && n < O) v = w[((size_t)n * I + c) * 9 + tap] * sc;
      const __half hi = __float2half_rn(v);
      d[i] = hl ? __float2half_rn(v - __half2float(hi)) : hi;
    }
    if (tid == 0) Q[j.dst[1]] = 1.f / sc;
  } else if (j.type == JOB_GATE2P) {
    // ConvLSTM gate weights for lstm_gate_f16.cu: OIHW (O = 4 R gate rows) -> fp16 [pass][kstep][tap][hl][2 planes][4 RH][8],
    // RH = R / 2; column n of pass p is gate row (n / RH) R + RH p + n % RH (i, f, o, g of recurrent channels [RH p, RH p + RH)).
    // The K planes are the sources (first input channel src[1 + q], channels nch0 / nch1 / nd), each padded to a multiple
    // of 8 channels, in the order given -- not necessarily the concatenation order of the reference.  Same power-of-two scale.
    const int O = j.a, I = j.b, R = O / 4, RH = R / 2, NPP = 4 * RH;
    const int nch[3] = {j.nch0, j.nch1, j.nd};
    int pl0[4]; pl0[0] = 0;
    for (int q = 0; q < 3; ++q) pl0[q + 1] = pl0[q] + (nch[q] + 7) / 8;
    const int KS = (pl0[3] + 1) / 2;
    const float* w = P + j.src[0];
    float m = 0.f;
    for (int i = tid; i < O * I * 9; i += blockDim.x) m = fmaxf(m, fabsf(w[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((tid & 31) == 0) sm[tid >> 5] = m;
    __syncthreads();
    m = 0.f;
    for (int q = 0; q < (int)(blockDim.x >> 5); ++q) m = fmaxf(m, sm[q]);
    int ex = 0;
    if (m > 0.f && m < 3.0e38f) frexpf(m, &ex);
    const float sc = m > 0.f ? ldexpf(1.f, min(max(11 - ex, -20), 40)) : 1.f;
    __half* d = reinterpret_cast<__half*>(Q + j.dst[0]);
    const size_t total = (size_t)2 * KS * 9 * 2 * 2 * NPP * 8;
    for (size_t i = tid; i < total; i += blockDim.x) {
      const int e = (int)(i & 7); size_t t = i >> 3; const int n = (int)(t % NPP); t /= NPP;
      const int lp = (int)(t & 1); t >>= 1; const int hl = (int)(t & 1); t >>= 1;
      const int tap = (int)(t % 9); t /= 9; const int ks = (int)(t % KS); const int p = (int)(t / KS);
      // odd plane count: the spare plane of the last K-step repeats the real one in the hi part ([W_hi ; W_hi]) and is zero in
      // the lo part ([W_lo ; 0]) -- the kernel feeds it [x_hi | x_lo] / [x_hi | 0] and issues two MMAs instead of three
      const bool dup = (pl0[3] & 1) && ks == KS - 1;
      const int plane = 2 * ks + (dup ? 0 : lp);
      int c = -1;
      for (int q = 0; q < 3; ++q)
        if (plane >= pl0[q] && plane < pl0[q + 1]) { const int ch = (plane - pl0[q]) * 8 + e; if (ch < nch[q]) c = (int)j.src[1 + q] + ch; }
      const int o = (n / RH) * R + p * RH + n % RH;
      float v = 0.f;
      if (c >= 0 && c < I && o < O) v = w[((size_t)o * I + c) * 9 + tap] * sc;
      const __half hi = __float2half_rn(v);
      d[i] = hl ? ((dup && lp == 1) ? __float2half_rn(0.f) : __float2half_rn(v - __half2float(hi))) : hi;
    }
    if (tid == 0) Q[j.dst[1]] = 1.f / sc;
  } else if (j.type == JOB_SLICE) {
    // Input-channel slice [nch0, nch0 + nch1) of an OIHW weight as tap-major fp32 [9][nch1][opad] for conv3x3_ffma: the
    // conditioning rows of the ConvLSTM gate / output convolutions, whose contribution is the same for every sample of one
    // low-fidelity input and is evaluated once per call (run_hoist).
    const int O = j.a, I = j.b, OP = j.opad, c0 = j.nch0, n = j.nch1;
    const float* w = P + j.src[0];
    float* d = Q + j.dst[0];
    // j.part = RH > 0: the columns come in the pass order of lstm_gate_f16.cu (column 4 RH p + RH g + jj <- gate row g R + RH p + jj)
    const int RH = j.part;
    for (int i = tid; i < 9 * n * OP; i += blockDim.x) {
      int o = i % OP; int t = i / OP; const int c = t % n, tap = t / n;
      if (RH > 0 && o < O) { const int p = o / (4 * RH), nn = o % (4 * RH); o = (nn / RH) * 2 * RH + p * RH + nn % RH; }
      d[i] = o < O ? w[((size_t)o * I + c0 + c) * 9 + tap] : 0.f;
    }
  } else if (j.type == JOB_CONV_F16_T) {
    // Data-gradient weights for conv3x3_f16.cu: the gradient w.r.t. the input of a 3x3 convolution is the convolution
    // of the output gradient with the transposed, tap-flipped kernel: K = forward output channel o (one source of O
    // channels), N = forward input channel c (sources padded to 4 columns each), tap t' holds w[o][c][8 - t'].  Same
    // power-of-two scaling.
    const int O = j.a, I = j.b, NP = j.opad;
    const int KS = ((O + 7) / 8 + 1) / 2;
    const float* w = P + j.src[0];
    float m = 0.f;
    for (int i = tid; i < O * I * 9; i += blockDim.x) m = fmaxf(m, fabsf(w[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((tid & 31) == 0) sm[tid >> 5] = m;
    __syncthreads();
    m = 0.f;
    for (int q = 0; q < (int)(blockDim.x >> 5); ++q) m = fmaxf(m, sm[q]);
    int ex = 0;
    if (m > 0.f && m < 3.0e38f) frexpf(m, &ex);
    const float sc = m > 0.f ? ldexpf(1.f, min(max(11 - ex, -20), 40)) : 1.f;
    __half* d = reinterpret_cast<__half*>(Q + j.dst[0]);
    const size_t total = (size_t)KS * 9 * 2 * 2 * NP * 8;
    for (size_t i = tid; i < total; i += blockDim.x) {
      const int e = (int)(i & 7); size_t t = i >> 3; const int n = (int)(t % NP); t /= NP;
      const int lp = (int)(t & 1); t >>= 1; const int hl = (int)(t & 1); t >>= 1;
      const int tap = (int)(t % 9); const int ks = (int)(t / 9);
      const int o = (2 * ks + lp) * 8 + e;
      // column n -> forward input channel c: the channels of each forward source start on a 4-column boundary, so the
      // routed epilogue can use 16-byte accesses per destination
      const int pb1 = (j.nch0 + 3) / 4 * 4, pb2 = pb1 + (j.nch1 + 3) / 4 * 4;
      int c = -1;
      if (n < pb1) { if (n < j.nch0) c = n; }
      else if (n < pb2) { if (n - pb1 < j.nch1) c = j.nch0 + (n - pb1); }
      else if (n - pb2 < j.nd) c = j.nch0 + j.nch1 + (n - pb2);
      float v = 0.f;
      if (o < O && c >= 0 && c < I) v = w[((size_t)o * I + c) * 9 + (8 - tap)] * sc;
      const __half hi = __float2half_rn(v);
      d[i] = hl ? __float2half_rn(v - __half2float(hi)) : hi;
    }
    if (tid == 0) Q[j.dst[1]] = 1.f / sc;
  } else if (j.type == JOB_GAIN) {
    if (tid == 0) Q[j.dst[0]] = expf(fminf(fmaxf(P[j.src[0]], -4.f), kLog4));
  } else if (j.type == JOB_BN) {
    const float* w = P + j.src[0]; const float* b = P + j.src[1];
    const float* rm = P + j.src[2]; const float* rv = P + j.src[3];
    for (int c = tid; c < j.a; c += blockDim.x) {
      float sc = w[c] * rsqrtf(rv[c] + 1e-5f);
      Q[j.dst[0] + c] = sc;
      Q[j.dst[1] + c] = b[c] - rm[c] * sc;
    }
  } else if (j.type == JOB_1X1) {
    const int C = j.a;
    float* L = sm;                 // L  = l*l_mask + I          (unit lower)
    float* U = L + cmax * cmax;    // U  = u*u_mask + diag(sign*exp(log_s)) + 0.01 I
    float* Li = U + cmax * cmax;   // L^-1
    float* Ui = Li + cmax * cmax;  // U^-1
    const float* pl = P + j.src[0]; const float* pu = P + j.src[1]; const float* pls = P + j.src[2];
    const float* pp = P + j.src[3]; const float* psg = P + j.src[4]; const float* plm = P + j.src[5];
    const float* pum = P + j.src[6]; const float* pe = P + j.src[7];
    for (int i = tid; i < C * C; i += blockDim.x) {
      int r = i / C, c = i % C;
      L[i] = pl[i] * plm[i] + pe[i];
      float dg = (r == c) ? expf(pls[r]) * psg[r] : 0.f;
      U[i] = pu[i] * pum[i] + dg + 0.01f * pe[i];
    }
    __syncthreads();
    // triangular inverses, one column per thread, accumulated in double
    if (tid < C) {
      const int c = tid;
      for (int r = 0; r < C; ++r) {          // L * X = I   (forward substitution)
        double s = (r == c) ? 1.0 : 0.0;
        for (int k = 0; k < r; ++k) s -= (double)L[r * C + k] * (double)Li[k * C + c];
        Li[r * C + c] = (float)(s / (double)L[r * C + r]);
      }
      for (int r = C - 1; r >= 0; --r) {     // U * X = I   (back substitution)
        double s = (r == c) ? 1.0 : 0.0;
        for (int k = r + 1; k < C; ++k) s -= (double)U[r * C + k] * (double)Ui[k * C + c];
        Ui[r * C + c] = (float)(s / (double)U[r * C + r]);
      }
    }
    __syncthreads();
    float* W = Q + j.dst[0];
    float* Wi = Q + j.dst[1];
    // W = P (L U);   W^-1 = U^-1 L^-1 P^-1   (P is a permutation: P^-1 = P^T)
    for (int i = tid; i < C * C; i += blockDim.x) {
      int r = i / C, c = i % C;
      // row r of P picks row `pr` of (L U)
      double acc = 0.0;
      for (int q = 0; q < C; ++q) {
        float pv = pp[r * C + q];
        if (pv != 0.f) {
          double lu = 0.0;
          for (int k = 0; k < C; ++k) lu += (double)L[q * C + k] * (double)U[k * C + c];
          acc += (double)pv * lu;
        }
      }
      W[i] = (float)acc;
      double acc2 = 0.0;                      // (Ui Li)[r][q] * P^T[q][c] = (Ui Li)[r][q] * P[c][q]
      for (int q = 0; q < C; ++q) {
        float pv = pp[c * C + q];
        if (pv != 0.f) {
          double ul = 0.0;
          for (int k = 0; k < C; ++k) ul += (double)Ui[r * C + k] * (double)Li[k * C + q];
          acc2 += (double)pv * ul;
        }
      }
      Wi[i] = (float)acc2;
    }
    if (j.dst[3] >= 0) {
      // the same W as a tensor-core B operand for the level-resident kernel's mix (u = W v per pixel: M = pixels, N = output
      // row r, K = input channel k): fp16 hi + lo, K-major [hl][K plane of 8, padded to an even count][NP rows][8 halves],
      // scaled by a power of two so that max|w| lands in [2^10, 2^11); word 0 of the block = 1 / scale
      __syncthreads();
      __threadfence_block();
      float m = 0.f;
      for (int i = tid; i < C * C; i += blockDim.x) m = fmaxf(m, fabsf(W[i]));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      __syncthreads();
      if ((tid & 31) == 0) sm[tid >> 5] = m;
      __syncthreads();
      m = 0.f;
      for (int q = 0; q < (int)(blockDim.x >> 5); ++q) m = fmaxf(m, sm[q]);
      int ex = 0;
      if (m > 0.f && m < 3.0e38f) frexpf(m, &ex);
      const float sc = m > 0.f ? ldexpf(1.f, min(max(11 - ex, -20), 40)) : 1.f;
      const int PM = (C / 8 + 1) / 2 * 2, NPm = (C + 15) / 16 * 16;
      float* blk = Q + j.dst[3];
      __half* d = reinterpret_cast<__half*>(blk + 4);
      const int total = 2 * PM * NPm * 8;
      for (int i = tid; i < total; i += blockDim.x) {
        const int e = i & 7; int t = i >> 3; const int n = t % NPm; t /= NPm; const int pl = t % PM; const int hl = t / PM;
        const int k = pl * 8 + e;
        const float v = (n < C && k < C) ? W[n * C + k] * sc : 0.f;
        const __half hi = __float2half_rn(v);
        d[i] = hl ? __float2half_rn(v - __half2float(hi)) : hi;
      }
      if (tid == 0) { blk[0] = 1.f / sc; blk[1] = 0.f; blk[2] = 0.f; blk[3] = 0.f; }
    }
    if (tid == 0) {    // step constant: sum log|w_actnorm| - sum log_s   (multiplied by H*W at run time)
      double s = 0.0;
      for (int c = 0; c < C; ++c) s -= (double)pls[c];
      if (j.src[8] >= 0) {
        const float* nw = P + j.src[8];
        for (int c = 0; c < C; ++c) s += log(fabs((double)nw[c]));
      }
      Q[j.dst[2]] = (float)s;
    }
  }
}

int launch_pack(const PackJob* jobs_dev, int njobs, const float* params, float* packed, int cmax,
                cudaStream_t st) {
  size_t smem = (size_t)4 * cmax * cmax * sizeof(float);
  TMG_SMEM_ATTR(pack_kernel, (int)smem);
  pack_kernel<<<njobs, 256, smem, st>>>(jobs_dev, params, packed, cmax);
  TMG_LAUNCH_CHECK();
  return TMG_OK;
}

}  // namespace tmg
