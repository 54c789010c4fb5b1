// One flow step = ONE launch: the whole coupling network of an affine coupling layer
//   t = cat(x1, cond)  ->  d1 = conv3x3(relu(t))  ->  d2 = conv3x3(relu(cat(t,d1)))
//     ->  h = Conv2dZeros(relu(cat(t,d1,d2)))                      (flowAffine.py:49-55,73; denseBlock.py:135-150;
//                                                                    flowUtils.py:238-247)
// plus the coupling update, the invertible 1x1 convolution, ActNorm and the per-sample log-det
// partial (flowAffine.py:76-81,102-107; glowConv.py:193,219; actNorm.py:66,82) in the epilogue,
// on tcgen05 tensor cores with everything between the input tile and the output tile kept in
// shared memory / TMEM (the 40-channel concat, d1, d2 and h never touch HBM).
//
// Tile: a CH x CW block of pixels of one sample plus a 3-pixel halo, stored as 4-channel planes
// [hi|lo][plane][position][4 floats] (K-major, no swizzle; position = row*RP + col, RP = CW+6).
//   * The two Cout=1 dense layers put the 9 filter TAPS in the MMA N dimension:
//       D[q][tap] = sum_c u[q][c] * w[c][tap]     (one unshifted GEMM, N = 16)
//       d[p]      = sum_tap D[p + off(tap)][tap]  (9-term gather by the epilogue warps, zero padding =
//                                                  skipping out-of-image taps)
//     so a Cout=1 convolution costs K/8 MMAs per 128 positions instead of 9*K/8.
//   * Conv2dZeros is the shifted-descriptor implicit GEMM of conv3x3_tc.cu (replicate padding: the
//     planes hold clamped pixels), weights streamed per (tap, K-slice group) with cp.async.bulk.
//   * 3xTF32 operand split (fp32-grade) or single-pass TF32.
// Warps 0-3: staging + the three epilogues; warp 4: MMA issue; warp 5: weights.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace tmg {

struct CplGeom {
  int CH, CW, RP, RR, NPOS, NPOSA;
  int tiles_x, tiles_y;
  int PT, NK1, NPL, planes0;
  int n12, n3, rows3, i3;
  int sps, nsg, nstage3;
  uint32_t plane_bytes, stage3_bytes;
  uint32_t oA, oDsc, oW1, oW2, oWm, oW3, oBar, total;
};

template <int C>
__global__ void __launch_bounds__(192, 1)
coupling_tc_kernel(CouplingArgs a, CplGeom g) {
  constexpr int NP = (C + 15) / 16 * 16;               // MMA N of Conv2dZeros
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y;
  const int tile_r0 = (blockIdx.x / g.tiles_x) * g.CH, tile_c0 = (blockIdx.x % g.tiles_x) * g.CW;
  const int nhl = a.split3 ? 2 : 1;
  const int HW = a.H * a.W;

  uint8_t* A = smem + g.oA;                              // [hl][NPL][NPOSA][16 B]
  float* Dsc = reinterpret_cast<float*>(smem + g.oDsc);  // [9][NPOSA]
  uint8_t* W1 = smem + g.oW1;                            // [hl][NK1][16][16 B]
  uint8_t* W2 = smem + g.oW2;                            // [hl][NPL][16][16 B]
  float* Wm = reinterpret_cast<float*>(smem + g.oWm);    // [C*C] + nw[C] + nb[C] + bias3[C]
  uint8_t* W3 = smem + g.oW3;                            // nstage3 x stage3_bytes
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + g.oBar);
  uint64_t* t_full = bars + 0;    // 128
  uint64_t* w_full = bars + 1;    // 32
  uint64_t* acc1 = bars + 2;      // commit
  uint64_t* d1_ready = bars + 3;  // 128
  uint64_t* acc2 = bars + 4;
  uint64_t* d2_ready = bars + 5;
  uint64_t* acc3 = bars + 6;
  uint64_t* b_full = bars + 7;                 // [nstage3]
  uint64_t* b_free = b_full + g.nstage3;       // [nstage3]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_free + g.nstage3);

  if (tid == 0) {
    mbar_init(t_full, 128); mbar_init(w_full, 32);
    mbar_init(acc1, 1); mbar_init(d1_ready, 128); mbar_init(acc2, 1); mbar_init(d2_ready, 128); mbar_init(acc3, 1);
    for (int i = 0; i < g.nstage3; ++i) { mbar_init(b_full + i, 1); mbar_init(b_free + i, 1); }
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem3 = tmem_base + 64;          // conv3 accumulators start at column 64
  const uint32_t hl_bytes = (uint32_t)g.NPL * g.plane_bytes;

  if (warp < 4) {
    // ===================================================== stage t (replicate-clamped, ReLU, hi/lo)
    {
      const int items = g.NPOS * g.PT;
      for (int it0 = tid; it0 < items; it0 += 128 * 4) {
        float4 v[4];
        int pos[4], pl[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int it = it0 + u * 128;
          v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          pos[u] = -1;
          if (it < items) {
            const int p = it % g.NPOS, plane = it / g.NPOS;
            pos[u] = p; pl[u] = plane;
            const int rr = p / g.RP, rc = p - rr * g.RP;
            const int r = min(max(tile_r0 - 3 + rr, 0), a.H - 1), c = min(max(tile_c0 - 3 + rc, 0), a.W - 1);
            const bool s1 = plane >= g.planes0;
            const ConvSrc& s = a.src[s1 ? 1 : 0];
            const int ch = (s1 ? plane - g.planes0 : plane) * 4;
            const int nv = min(4, s.nch - ch);
            const float* ptr = s.p + ((s.bshared ? 0 : (size_t)b * HW) + (size_t)r * a.W + c) * s.cstride + s.coff + ch;
            float4 t;
            if ((reinterpret_cast<uintptr_t>(ptr) & 15) == 0) {
              t = __ldg(reinterpret_cast<const float4*>(ptr));
            } else {
              t.x = __ldg(ptr); t.y = nv > 1 ? __ldg(ptr + 1) : 0.f; t.z = nv > 2 ? __ldg(ptr + 2) : 0.f;
              t.w = nv > 3 ? __ldg(ptr + 3) : 0.f;
            }
            if (nv < 4) { t.w = 0.f; if (nv < 3) t.z = 0.f; if (nv < 2) t.y = 0.f; }
            if (s.relu) { t.x = fmaxf(t.x, 0.f); t.y = fmaxf(t.y, 0.f); t.z = fmaxf(t.z, 0.f); t.w = fmaxf(t.w, 0.f); }
            v[u] = t;
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (pos[u] < 0) continue;
          const float4 t = v[u];
          const float4 hi = make_float4(tf32_hi(t.x), tf32_hi(t.y), tf32_hi(t.z), tf32_hi(t.w));
          uint8_t* dst = A + (size_t)pl[u] * g.plane_bytes + (size_t)pos[u] * 16;
          *reinterpret_cast<float4*>(dst) = hi;
          if (a.split3) *reinterpret_cast<float4*>(dst + hl_bytes) = make_float4(t.x - hi.x, t.y - hi.y, t.z - hi.z, t.w - hi.w);
        }
      }
      // d plane and padding planes start as zeros (finite operands for the zero-weight K rows)
      const int zitems = g.NPOSA * (g.NPL - g.PT);
      for (int it = tid; it < zitems; it += 128) {
        const int p = it % g.NPOSA, plane = g.PT + it / g.NPOSA;
        uint8_t* dst = A + (size_t)plane * g.plane_bytes + (size_t)p * 16;
        *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a.split3) *reinterpret_cast<float4*>(dst + hl_bytes) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      // slack positions [NPOS, NPOSA) of the t planes
      const int sl = g.NPOSA - g.NPOS;
      for (int it = tid; it < sl * g.PT; it += 128) {
        const int p = g.NPOS + it % sl, plane = it / sl;
        uint8_t* dst = A + (size_t)plane * g.plane_bytes + (size_t)p * 16;
        *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a.split3) *reinterpret_cast<float4*>(dst + hl_bytes) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      fence_proxy_async();
      mbar_arrive(t_full);
    }

    // ===================================================== epilogues 1 and 2: tap partials -> d1 / d2
    for (int layer = 0; layer < 2; ++layer) {
      mbar_wait(layer == 0 ? acc1 : acc2, 0);
      tc_fence_after();
      for (int mt = 0; mt < g.n12; ++mt) {
        const int m0 = min(mt * 128, g.NPOSA - 128);
        float v[16];
        tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(mt * 16), v);
#pragma unroll
        for (int t = 0; t < 9; ++t) Dsc[t * g.NPOSA + m0 + tid] = v[t];
      }
      tc_fence_before();
      named_bar_sync(1, 128);
      uint8_t* dpl = A + (size_t)g.PT * g.plane_bytes;        // the d plane: channel 0 = d1, channel 1 = d2
      for (int p = tid; p < g.NPOS; p += 128) {
        const int rr = p / g.RP, rc = p - rr * g.RP;
        const int ir = tile_r0 - 3 + rr, ic = tile_c0 - 3 + rc;
        if (rr >= 1 && rr < g.RR - 1 && rc >= 1 && rc < g.RP - 1 && ir >= 0 && ir < a.H && ic >= 0 && ic < a.W) {
          float s = 0.f;
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            const int dr = t / 3 - 1, dc = t % 3 - 1;
            if (ir + dr >= 0 && ir + dr < a.H && ic + dc >= 0 && ic + dc < a.W)   // zero padding of the dense layers
              s += Dsc[t * g.NPOSA + p + dr * g.RP + dc];
          }
          s = fmaxf(s, 0.f);                                   // every consumer reads relu(d)
          const float hi = tf32_hi(s);
          *reinterpret_cast<float*>(dpl + (size_t)p * 16 + layer * 4) = hi;
          if (a.split3) *reinterpret_cast<float*>(dpl + hl_bytes + (size_t)p * 16 + layer * 4) = s - hi;
        }
      }
      named_bar_sync(1, 128);
      // replicate padding for Conv2dZeros: out-of-image halo positions copy the clamped pixel
      for (int p = tid; p < g.NPOS; p += 128) {
        const int rr = p / g.RP, rc = p - rr * g.RP;
        const int ir = tile_r0 - 3 + rr, ic = tile_c0 - 3 + rc;
        if (ir < 0 || ir >= a.H || ic < 0 || ic >= a.W) {
          const int cr = min(max(ir, 0), a.H - 1) - (tile_r0 - 3), cc = min(max(ic, 0), a.W - 1) - (tile_c0 - 3);
          if (cr >= 1 && cr < g.RR - 1 && cc >= 1 && cc < g.RP - 1) {
            const int pc = cr * g.RP + cc;
            *reinterpret_cast<float*>(dpl + (size_t)p * 16 + layer * 4) =
                *reinterpret_cast<const float*>(dpl + (size_t)pc * 16 + layer * 4);
            if (a.split3)
              *reinterpret_cast<float*>(dpl + hl_bytes + (size_t)p * 16 + layer * 4) =
                  *reinterpret_cast<const float*>(dpl + hl_bytes + (size_t)pc * 16 + layer * 4);
          }
        }
      }
      fence_proxy_async();
      mbar_arrive(layer == 0 ? d1_ready : d2_ready);
    }

    // ===================================================== epilogue 3: h -> coupling, 1x1, ActNorm, log-det
    mbar_wait(w_full, 0);
    mbar_wait(acc3, 0);
    tc_fence_after();
    const float* s_w = Wm;
    const float* s_nw = Wm + C * C;
    const float* s_nb = s_nw + C;
    const float* s_b3 = s_nb + C;
    const float gain = a.gain3 ? __ldg(a.gain3) : 1.f;
    float ldsum = 0.f;
    for (int mt = 0; mt < g.n3; ++mt) {
      const int st = (mt == g.n3 - 1 && g.rows3 >= 128) ? g.rows3 - 128 : mt * 128;
      const int i = g.i3 + st + tid;
      const int rr = i / g.RP, rc = i - rr * g.RP;
      const int ir = tile_r0 - 3 + rr, ic = tile_c0 - 3 + rc;
      const bool valid = st + tid >= mt * 128 && st + tid < g.rows3 && rr >= 3 && rr < 3 + g.CH && rc >= 3 && rc < 3 + g.CW &&
                         ir < a.H && ic < a.W;
      const size_t pix = valid ? (size_t)b * HW + (size_t)ir * a.W + ic : 0;
      float v[C];
      if (valid) {
        const float4* y4 = reinterpret_cast<const float4*>(a.y_in + pix * C);
#pragma unroll
        for (int k = 0; k < C / 4; ++k) { float4 t = __ldg(y4 + k); v[4 * k] = t.x; v[4 * k + 1] = t.y; v[4 * k + 2] = t.z; v[4 * k + 3] = t.w; }
      }
      const uint32_t trow = tmem3 + ((uint32_t)(warp * 32) << 16) + (uint32_t)(mt * NP);
#pragma unroll
      for (int n0 = 0; n0 < NP; n0 += 16) {
        float h[16];
        tmem_ld16(trow + n0, h);
        if (valid) {
#pragma unroll
          for (int k = 0; k < 16; k += 2) {
            if (n0 + k < C) {
              const float shift = (h[k] + s_b3[n0 + k]) * gain;          // h[:,0::2]
              const float raw = (h[k + 1] + s_b3[n0 + k + 1]) * gain;    // h[:,1::2]
              const float la = 2.f * (raw / (1.f + fabsf(raw)));
              ldsum += la;
              const float sc = expf(la);
              const int j = C / 2 + (n0 + k) / 2;
              v[j] = a.reverse ? v[j] / sc - shift : (v[j] + shift) * sc;
            }
          }
        }
      }
      if (valid) {
        if (!a.reverse && a.nw) {
#pragma unroll
          for (int k = 0; k < C; ++k) v[k] = fmaf(s_nw[k], v[k], s_nb[k]);
        }
        float* yo = a.y_out + pix * C;
        if (a.wmat) {
#pragma unroll 1
          for (int r4 = 0; r4 < C; r4 += 4) {
            float o[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float s = 0.f;
              const float4* w4 = reinterpret_cast<const float4*>(s_w + (r4 + q) * C);
#pragma unroll
              for (int k = 0; k < C / 4; ++k) {
                const float4 w = w4[k];
                s = fmaf(w.x, v[4 * k], s); s = fmaf(w.y, v[4 * k + 1], s);
                s = fmaf(w.z, v[4 * k + 2], s); s = fmaf(w.w, v[4 * k + 3], s);
              }
              if (a.reverse && a.nw) s = (s - s_nb[r4 + q]) / s_nw[r4 + q];
              o[q] = s;
            }
            *reinterpret_cast<float4*>(yo + r4) = make_float4(o[0], o[1], o[2], o[3]);
          }
        } else {
#pragma unroll
          for (int k = 0; k < C; k += 4) {
            float o[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) o[q] = (a.reverse && a.nw) ? (v[k + q] - s_nb[k + q]) / s_nw[k + q] : v[k + q];
            *reinterpret_cast<float4*>(yo + k) = make_float4(o[0], o[1], o[2], o[3]);
          }
        }
      }
    }
    if (a.ld_part) {
      // block reduction over the 128 epilogue threads
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ldsum += __shfl_xor_sync(0xffffffffu, ldsum, o);
      if (lane == 0) Dsc[warp] = ldsum;
      named_bar_sync(1, 128);
      if (tid == 0) a.ld_part[(size_t)b * a.ld_stride + blockIdx.x] = Dsc[0] + Dsc[1] + Dsc[2] + Dsc[3];
    }
  } else if (warp == 4) {
    // ===================================================== MMA issue
    if (elect_one()) {
      const uint32_t idesc16 = make_idesc_tf32(16);
      const uint32_t a_base = smem_u32(A);
      mbar_wait(t_full, 0);
      mbar_wait(w_full, 0);
      tc_fence_after();
      for (int layer = 0; layer < 2; ++layer) {
        if (layer == 1) { mbar_wait(d1_ready, 0); tc_fence_after(); }
        const int nks = (layer == 0 ? g.NK1 : g.NPL) / 2;
        const uint32_t w_base = smem_u32(layer == 0 ? W1 : W2);
        const uint32_t w_hl = (uint32_t)(layer == 0 ? g.NK1 : g.NPL) * 256u;
        for (int mt = 0; mt < g.n12; ++mt) {
          const uint32_t m0 = (uint32_t)min(mt * 128, g.NPOSA - 128);
          for (int ks = 0; ks < nks; ++ks) {
            const uint32_t ao = a_base + (uint32_t)(2 * ks) * g.plane_bytes + m0 * 16u;
            const uint64_t a_hi = make_desc(ao, g.plane_bytes, 128);
            const uint64_t b_hi = make_desc(w_base + (uint32_t)(2 * ks) * 256u, 256, 128);
            mma_tf32(tmem_base + (uint32_t)(mt * 16), a_hi, b_hi, idesc16, ks > 0 ? 1u : 0u);
            if (a.split3) {
              const uint64_t a_lo = make_desc(ao + hl_bytes, g.plane_bytes, 128);
              const uint64_t b_lo = make_desc(w_base + w_hl + (uint32_t)(2 * ks) * 256u, 256, 128);
              mma_tf32(tmem_base + (uint32_t)(mt * 16), a_lo, b_hi, idesc16, 1);
              mma_tf32(tmem_base + (uint32_t)(mt * 16), a_hi, b_lo, idesc16, 1);
            }
          }
        }
        mma_commit(layer == 0 ? acc1 : acc2);
      }
      mbar_wait(d2_ready, 0);
      tc_fence_after();
      const uint32_t idesc3 = make_idesc_tf32(a.npad);
      const uint32_t lbo3 = (uint32_t)a.npad * 16u;
      const uint32_t st_hl = (uint32_t)(2 * g.sps) * lbo3;         // hi part of a stage, then lo part
      int j = 0;
      for (int tap = 0; tap < 9; ++tap) {
        const int toff = (tap / 3 - 1) * g.RP + (tap % 3 - 1);
        for (int sg = 0; sg < g.nsg; ++sg, ++j) {
          const int si = j % g.nstage3;
          mbar_wait(b_full + si, (uint32_t)((j / g.nstage3) & 1));
          tc_fence_after();
          const uint32_t b_base = smem_u32(W3 + (size_t)si * g.stage3_bytes);
          const int ks_end = min(g.sps, g.NPL / 2 - sg * g.sps);
          for (int mt = 0; mt < g.n3; ++mt) {
            const int st = (mt == g.n3 - 1 && g.rows3 >= 128) ? g.rows3 - 128 : mt * 128;
            const uint32_t row0 = (uint32_t)(g.i3 + st + toff);
            for (int ks = 0; ks < ks_end; ++ks) {
              const uint32_t plane = (uint32_t)(2 * (sg * g.sps + ks));
              const uint32_t ao = a_base + plane * g.plane_bytes + row0 * 16u;
              const uint64_t a_hi = make_desc(ao, g.plane_bytes, 128);
              const uint64_t b_hi = make_desc(b_base + (uint32_t)(2 * ks) * lbo3, lbo3, 128);
              const uint32_t acc = (tap == 0 && sg == 0 && ks == 0) ? 0u : 1u;
              mma_tf32(tmem3 + (uint32_t)(mt * a.npad), a_hi, b_hi, idesc3, acc);
              if (a.split3) {
                const uint64_t a_lo = make_desc(ao + hl_bytes, g.plane_bytes, 128);
                const uint64_t b_lo = make_desc(b_base + st_hl + (uint32_t)(2 * ks) * lbo3, lbo3, 128);
                mma_tf32(tmem3 + (uint32_t)(mt * a.npad), a_lo, b_hi, idesc3, 1);
                mma_tf32(tmem3 + (uint32_t)(mt * a.npad), a_hi, b_lo, idesc3, 1);
              }
            }
          }
          mma_commit(b_free + si);
        }
      }
      mma_commit(acc3);
    }
  } else {
    // ===================================================== weights: small ones by the warp, w3 streamed by lane 0
    {
      const int n1 = nhl * g.NK1 * 16, n2 = nhl * g.NPL * 16;          // float4 counts (256 B = 16 float4 per plane)
      const float4* s1 = reinterpret_cast<const float4*>(a.w1);
      const float4* s2 = reinterpret_cast<const float4*>(a.w2);
      // packed w1/w2 always hold [hi|lo]; in single-pass mode only the hi half is copied
      for (int i = lane; i < n1; i += 32) reinterpret_cast<float4*>(W1)[i] = __ldg(s1 + i);
      for (int i = lane; i < n2; i += 32) reinterpret_cast<float4*>(W2)[i] = __ldg(s2 + i);
      for (int i = lane; i < C * C; i += 32) Wm[i] = a.wmat ? __ldg(a.wmat + i) : 0.f;
      for (int i = lane; i < C; i += 32) {
        Wm[C * C + i] = a.nw ? __ldg(a.nw + i) : 1.f;
        Wm[C * C + C + i] = a.nw ? __ldg(a.nb + i) : 0.f;
        Wm[C * C + 2 * C + i] = __ldg(a.bias3 + i);
      }
      fence_proxy_async();
      mbar_arrive(w_full);
    }
    if (lane == 0) {
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(a.w3);
      const uint32_t lbo3 = (uint32_t)a.npad * 16u;
      const size_t tap_stride = (size_t)2 * g.NPL * lbo3;            // [tap][hi|lo][plane][npad][4]
      int j = 0;
      for (int tap = 0; tap < 9; ++tap) {
        for (int sg = 0; sg < g.nsg; ++sg, ++j) {
          const int si = j % g.nstage3, use = j / g.nstage3;
          if (use >= 1) mbar_wait(b_free + si, (uint32_t)((use - 1) & 1));
          const int nks = min(g.sps, g.NPL / 2 - sg * g.sps);
          const uint32_t part = (uint32_t)(2 * nks) * lbo3;
          mbar_expect_tx(b_full + si, part * nhl);
          uint8_t* dst = W3 + (size_t)si * g.stage3_bytes;
          const uint8_t* src = wsrc + (size_t)tap * tap_stride + (size_t)(2 * sg * g.sps) * lbo3;
          bulk_g2s(dst, src, part, b_full + si);
          if (a.split3) bulk_g2s(dst + (size_t)(2 * g.sps) * lbo3, src + (size_t)g.NPL * lbo3, part, b_full + si);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) { tc_fence_after(); tmem_dealloc(tmem_base, 256); }
}

// ------------------------------------------------------------------ host side
static bool make_geom(const CouplingArgs& a, CplGeom& g) {
  const int nhl = a.split3 ? 2 : 1;
  g.planes0 = (a.src[0].nch + 3) / 4;
  g.PT = g.planes0 + (a.nsrc > 1 ? (a.src[1].nch + 3) / 4 : 0);
  g.NK1 = (g.PT + 1) / 2 * 2;
  g.NPL = (g.PT + 1 + 1) / 2 * 2;
  const int cands[4][2] = {{16, 16}, {8, 16}, {8, 8}, {4, 8}};
  for (int k = 0; k < 4; ++k) {
    g.CH = std::min(cands[k][0], a.H); g.CW = std::min(cands[k][1], a.W);
    g.RP = g.CW + 6; g.RR = g.CH + 6; g.NPOS = g.RR * g.RP;
    g.NPOSA = std::max(128, (g.NPOS + 7) / 8 * 8);
    g.plane_bytes = (uint32_t)g.NPOSA * 16u;
    g.n12 = cdiv(g.NPOSA, 128);
    g.i3 = 3 * g.RP + 3;
    g.rows3 = (g.CH - 1) * g.RP + g.CW;
    g.n3 = cdiv(g.rows3, 128);
    if (g.n12 * 16 > 64 || g.n3 * a.npad > 192) continue;
    // weight stages of Conv2dZeros: sps K-slices (of 2 planes) per stage
    const uint32_t slice_bytes = (uint32_t)nhl * 2u * a.npad * 16u;
    g.sps = std::max(1, std::min(g.NPL / 2, (int)(12 * 1024 / slice_bytes)));
    g.nsg = cdiv(g.NPL / 2, g.sps);
    g.stage3_bytes = slice_bytes * g.sps;
    g.nstage3 = 3;
    uint32_t off = 0;
    auto take = [&](uint32_t n) { uint32_t o = off; off += (n + 127) / 128 * 128; return o; };
    g.oA = take((uint32_t)nhl * g.NPL * g.plane_bytes);
    g.oDsc = take((uint32_t)9 * g.NPOSA * 4);
    g.oW1 = take((uint32_t)nhl * g.NK1 * 256);
    g.oW2 = take((uint32_t)nhl * g.NPL * 256);
    g.oWm = take((uint32_t)(a.C * a.C + 3 * a.C) * 4);
    g.oW3 = take(g.stage3_bytes * g.nstage3);
    g.oBar = take((7 + 2 * g.nstage3) * 8 + 16);
    g.total = off;
    if (g.total > 225 * 1024) { g.nstage3 = 2; off = g.oW3; take(g.stage3_bytes * 2); g.oBar = take((7 + 4) * 8 + 16); g.total = off; }
    if (g.total > 225 * 1024) continue;
    g.tiles_x = cdiv(a.W, g.CW); g.tiles_y = cdiv(a.H, g.CH);
    return true;
  }
  return false;
}

int coupling_tc_tiles(int H, int W) {       // upper bound of CTAs per sample (log-det partial slots)
  return cdiv(H, std::min(4, H)) * cdiv(W, std::min(8, W));
}

int launch_coupling_tc(const CouplingArgs& a, cudaStream_t st) {
  if (a.B <= 0) return TMG_OK;
  CplGeom g{};
  if (a.C % 4 || a.C > kMaxC || a.npad != (a.C + 15) / 16 * 16 || !make_geom(a, g)) {
    set_error("fused coupling step: unsupported shape (C=%d, %dx%d)", a.C, a.H, a.W);
    return TMG_ERR_UNSUPPORTED;
  }
  dim3 grid(g.tiles_x * g.tiles_y, a.B);
  switch (a.C) {
#define TMG_CASE(CC)                                                                                                   \
  case CC:                                                                                                             \
    TMG_SMEM_ATTR(coupling_tc_kernel<CC>, 227 * 1024); \
    coupling_tc_kernel<CC><<<grid, 192, g.total, st>>>(a, g);                                                          \
    break;
    TMG_CASE(4) TMG_CASE(8) TMG_CASE(12) TMG_CASE(16) TMG_CASE(24) TMG_CASE(32) TMG_CASE(48) TMG_CASE(64)
#undef TMG_CASE
    default:
      set_error("fused coupling step: %d channels not supported", a.C);
      return TMG_ERR_UNSUPPORTED;
  }
  TMG_LAUNCH_CHECK();
  return TMG_OK;
}

}  // namespace tmg
