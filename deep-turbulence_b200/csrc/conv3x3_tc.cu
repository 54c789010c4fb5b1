// 3x3 convolution as a tcgen05 implicit GEMM (sm_100a): D[128 pixels, N] += A[128, K] * B[K, N]
// with K = 9 taps x Cin, accumulators in TMEM, operands in shared memory.
//
// Replaces cuDNN/MKLDNN convolutions of the reference for the heavy instances of the path:
//   ConvLSTM gate conv          nn/modules/convLSTM.py:44,74      (N = 4*rec = 256)
//   LSTM_out_conv               nn/modules/convLSTM.py:129        (N = C/2+cond)
//   Conv2dZeros of coupling nets nn/modules/flowUtils.py:229,246  (N = C, replicate padding)
//
// Layout trick: the activation halo tile is staged ONCE per 16-channel chunk in the no-swizzle
// K-major canonical layout as 4-channel planes  [plane][padded pixel][4 floats]  where "padded
// pixel" is the linear index into the zero/replicate-padded image with row pitch P = W+2.  A tile
// is 128 consecutive padded pixels, so filter tap (dr,dc) is the SAME tile read through a shared
// memory descriptor whose start address is advanced by (dr*P+dc)*16 bytes: no im2col, no re-load.
// The two halo columns of every row produce garbage accumulator rows that are never stored.
//
// Precision: kind::tf32 with the 3xTF32 split (a = a_hi + a_lo, b = b_hi + b_lo, three MMAs per
// K-slice) gives fp32-grade results (~1e-6 relative); `split3 = 0` is the single-pass TF32 mode.
//
// Warp roles (192 threads): warps 0-3 stage A (global -> ReLU/pad/split -> st.shared), then run
// the epilogue (tcgen05.ld -> bias/gain/activation or the fused ConvLSTM cell update -> global);
// warp 4 lane 0 issues the MMAs; warp 5 lane 0 streams packed weights with cp.async.bulk.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace tmg {

constexpr int kTcThreads = 192;
constexpr int kTcKC = 16;          // channels per chunk (4 planes of 4)

// ------------------------------------------------------------------ the kernel
__global__ void __launch_bounds__(kTcThreads, 1)
conv3x3_tc_kernel(TcConvArgs a, int npos_pad, int nstage, int tps, int tmem_cols) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y;
  const int P = a.W + 2;
  const int q0 = P + blockIdx.x * 128;             // first output position (padded-input linear index)
  const int nhl = a.split3 ? 2 : 1;
  const int nchunks = (a.cin + kTcKC - 1) / kTcKC;
  const int groups = 9 / tps;                      // weight stages per chunk (tps in {1,3,9})
  const uint32_t plane_bytes = (uint32_t)npos_pad * 16u;
  const uint32_t abuf_bytes = (uint32_t)nhl * 4u * plane_bytes;
  const uint32_t tap_bytes = (uint32_t)nhl * 4u * a.npad * 16u;     // one tap of one chunk (hi[,lo])
  const uint32_t stage_bytes = tap_bytes * tps;

  uint8_t* a_smem = smem_raw;                                   // 2 buffers
  uint8_t* b_smem = a_smem + 2 * abuf_bytes;                    // nstage stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_smem + (size_t)nstage * stage_bytes);
  uint64_t* a_full = bars;            // [2]   count 128
  uint64_t* a_free = bars + 2;        // [2]   count 1 (tcgen05.commit)
  uint64_t* b_full = bars + 4;        // [nstage] tx
  uint64_t* b_free = b_full + nstage; // [nstage] count 1
  uint64_t* acc_full = b_free + nstage;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(a_full + i, 128); mbar_init(a_free + i, 1); }
    for (int i = 0; i < nstage; ++i) { mbar_init(b_full + i, 1); mbar_init(b_free + i, 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, (uint32_t)tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ============================ A staging ============================
    const int items = npos_pad * 4;
    for (int c = 0; c < nchunks; ++c) {
      const int buf = c & 1, use = c >> 1;
      if (use >= 1) { mbar_wait(a_free + buf, (uint32_t)((use - 1) & 1)); tc_fence_after(); }
      uint8_t* ab = a_smem + (size_t)buf * abuf_bytes;
      const int c0 = c * kTcKC;
      for (int it = tid; it < items; it += 128) {
        const int i = it % npos_pad, plane = it / npos_pad;
        const int Q = q0 - P - 1 + i;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (Q >= 0 && Q < (a.H + 2) * P) {
          int r = Q / P;
          int cc = Q - r * P - 1;
          r -= 1;
          bool inb = r >= 0 && r < a.H && cc >= 0 && cc < a.W;
          if (a.pad_replicate) { r = min(max(r, 0), a.H - 1); cc = min(max(cc, 0), a.W - 1); inb = true; }
          if (inb) {
            const size_t pix0 = (size_t)r * a.W + cc, pixb = (size_t)b * a.H * a.W + pix0;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              int ch = c0 + plane * 4 + e;
              if (ch < a.cin) {
                const ConvSrc* s = &a.src[0];
                if (ch >= s->nch && a.nsrc > 1) { ch -= s->nch; s = &a.src[1];
                  if (ch >= s->nch && a.nsrc > 2) { ch -= s->nch; s = &a.src[2]; } }
                if (ch < s->nch) {
                  float t = __ldg(s->p + (s->bshared ? pix0 : pixb) * s->cstride + s->coff + ch);
                  v[e] = s->relu ? fmaxf(t, 0.f) : t;
                }
              }
            }
          }
        }
        float4 hi = make_float4(tf32_hi(v[0]), tf32_hi(v[1]), tf32_hi(v[2]), tf32_hi(v[3]));
        *reinterpret_cast<float4*>(ab + (size_t)plane * plane_bytes + (size_t)i * 16) = hi;
        if (a.split3) {
          float4 lo = make_float4(v[0] - hi.x, v[1] - hi.y, v[2] - hi.z, v[3] - hi.w);
          *reinterpret_cast<float4*>(ab + (size_t)(4 + plane) * plane_bytes + (size_t)i * 16) = lo;
        }
      }
      fence_proxy_async();            // generic-proxy stores -> visible to the tensor core (async proxy)
      mbar_arrive(a_full + buf);
    }

    // ============================ epilogue ============================
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const int q = q0 + tid;                          // this thread's accumulator row (TMEM lane = tid)
    int r = q / P;
    const int cc = q - r * P - 1;
    r -= 1;
    const bool valid = r >= 0 && r < a.H && cc >= 0 && cc < a.W;
    const size_t pix = valid ? ((size_t)b * a.H + r) * a.W + cc : 0;
    const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
    if (a.lstm_R > 0) {
      // fused ConvLSTM cell (convLSTM.py:76-83): gates i,f,o,g at columns [0,R),[R,2R),[2R,3R),[3R,4R)
      const int R = a.lstm_R;
      for (int r0 = 0; r0 < R; r0 += 16) {
        float gi[16], gf[16], go[16], gg[16];
        tmem_ld16(trow + r0, gi);
        tmem_ld16(trow + R + r0, gf);
        tmem_ld16(trow + 2 * R + r0, go);
        tmem_ld16(trow + 3 * R + r0, gg);
        if (valid) {
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            if (r0 + k < R) {
              float i_ = sigm(gi[k] + __ldg(a.bias + r0 + k));
              float f_ = sigm(gf[k] + __ldg(a.bias + R + r0 + k));
              float o_ = sigm(go[k] + __ldg(a.bias + 2 * R + r0 + k));
              float g_ = tanhf(gg[k] + __ldg(a.bias + 3 * R + r0 + k));
              float cp = a.c_prev ? __ldg(a.c_prev + pix * R + r0 + k) : 0.f;
              float cn = f_ * cp + i_ * g_;
              a.c_out[pix * R + r0 + k] = cn;
              a.h_out[pix * R + r0 + k] = o_ * tanhf(cn);
            }
          }
        }
      }
    } else {
      const float gain = a.gain ? __ldg(a.gain) : 1.f;
      for (int n0 = 0; n0 < a.npad; n0 += 16) {
        float v[16];
        tmem_ld16(trow + n0, v);
        if (valid) {
          float* op = a.out + pix * a.out_cstride + a.out_coff + n0;
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            if (n0 + k < a.cout) {
              float t = v[k];
              if (a.bias) t += __ldg(a.bias + n0 + k);
              if (a.gain) t *= gain;
              if (a.act == 1) t = fmaxf(t, 0.f);
              else if (a.act == 2) t = fminf(fmaxf(t, -2.f), kLog5);
              op[k] = t;
            }
          }
        }
      }
    }
  } else if (warp == 4) {
    // ============================ MMA issue (one elected lane) ============================
    if (elect_one()) {
      const uint32_t idesc = make_idesc_tf32(a.npad);
      const uint32_t lbo_b = (uint32_t)a.npad * 16u;
      uint32_t acc = 0;
      int j = 0;                                       // running weight-stage counter
      for (int c = 0; c < nchunks; ++c) {
        const int buf = c & 1;
        mbar_wait(a_full + buf, (uint32_t)((c >> 1) & 1));
        tc_fence_after();
        const uint32_t a_base = smem_u32(a_smem + (size_t)buf * abuf_bytes);
        for (int g = 0; g < groups; ++g, ++j) {
          const int si = j % nstage;
          mbar_wait(b_full + si, (uint32_t)((j / nstage) & 1));
          tc_fence_after();
          const uint32_t b_base = smem_u32(b_smem + (size_t)si * stage_bytes);
          for (int t = 0; t < tps; ++t) {
            const int tap = g * tps + t;
            const uint32_t a_tap = a_base + (uint32_t)((tap / 3) * P + (tap % 3)) * 16u;
            const uint32_t b_tap = b_base + (uint32_t)t * tap_bytes;
#pragma unroll
            for (int s = 0; s < 2; ++s) {              // two K-slices of 8 channels
              const uint64_t a_hi = make_desc(a_tap + (uint32_t)(2 * s) * plane_bytes, plane_bytes, 128);
              const uint64_t b_hi = make_desc(b_tap + (uint32_t)(2 * s) * lbo_b, lbo_b, 128);
              mma_tf32(tmem_base, a_hi, b_hi, idesc, acc);
              acc = 1;
              if (a.split3) {
                const uint64_t a_lo = make_desc(a_tap + (uint32_t)(4 + 2 * s) * plane_bytes, plane_bytes, 128);
                const uint64_t b_lo = make_desc(b_tap + (uint32_t)(4 + 2 * s) * lbo_b, lbo_b, 128);
                mma_tf32(tmem_base, a_lo, b_hi, idesc, 1);
                mma_tf32(tmem_base, a_hi, b_lo, idesc, 1);
              }
            }
          }
          mma_commit(b_free + si);                     // weights of this stage consumed
        }
        mma_commit(a_free + buf);                      // activation chunk consumed
      }
      mma_commit(acc_full);                            // accumulator complete
    }
  } else {
    // ============================ weight streaming (one thread) ============================
    if (lane == 0) {
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(a.wpk);
      const size_t tap_stride = (size_t)2 * 4 * a.npad * 16;   // packed layout always holds hi and lo
      int j = 0;
      for (int c = 0; c < nchunks; ++c) {
        for (int g = 0; g < groups; ++g, ++j) {
          const int si = j % nstage, use = j / nstage;
          if (use >= 1) mbar_wait(b_free + si, (uint32_t)((use - 1) & 1));
          mbar_expect_tx(b_full + si, stage_bytes);
          uint8_t* dst = b_smem + (size_t)si * stage_bytes;
          for (int t = 0; t < tps; ++t) {
            const int tap = g * tps + t;
            bulk_g2s(dst + (size_t)t * tap_bytes, wsrc + ((size_t)c * 9 + tap) * tap_stride, tap_bytes, b_full + si);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) { tc_fence_after(); tmem_dealloc(tmem_base, (uint32_t)tmem_cols); }
}

// ------------------------------------------------------------------ host side
int launch_conv3x3_tc(const TcConvArgs& a, cudaStream_t st) {
  if (a.B <= 0 || a.H <= 0 || a.W <= 0) return TMG_OK;
  if (a.npad % 16 || a.npad < 16 || a.npad > 256) { set_error("tc conv: N=%d unsupported", a.npad); return TMG_ERR_UNSUPPORTED; }
  const int P = a.W + 2;
  const int npos = 128 + 2 * P + 2;
  const int npos_pad = (npos + 7) / 8 * 8;
  const int nhl = a.split3 ? 2 : 1;
  const size_t abuf = (size_t)nhl * 4 * npos_pad * 16;
  const size_t tap_bytes = (size_t)nhl * 4 * a.npad * 16;
  // taps per weight stage: whole chunk (9) when it is small, 3 or 1 for wide N
  int tps = tap_bytes * 9 <= 40 * 1024 ? 9 : (tap_bytes * 3 <= 40 * 1024 ? 3 : 1);
  int nstage = tps == 9 ? 2 : 4;
  size_t smem = 2 * abuf + (size_t)nstage * tps * tap_bytes + (4 + 2 * nstage + 1) * sizeof(uint64_t) + 16;
  while (smem > 220 * 1024 && nstage > 2) { --nstage; smem -= tps * tap_bytes; }
  if (smem > 220 * 1024) { set_error("tc conv: width %d / N %d needs %zu B of shared memory", a.W, a.npad, smem); return TMG_ERR_UNSUPPORTED; }
  int cols = 32;
  while (cols < a.npad) cols *= 2;
  TMG_SMEM_ATTR(conv3x3_tc_kernel, 220 * 1024);
  dim3 grid(cdiv(a.H * P, 128), a.B);
  conv3x3_tc_kernel<<<grid, kTcThreads, smem, st>>>(a, npos_pad, nstage, tps, cols);
  TMG_LAUNCH_CHECK();
  return TMG_OK;
}

// Weight packing for the tensor-core path:  OIHW -> [chunk][tap][hi|lo][plane][npad][4]
__global__ void pack_tc_kernel(const float* __restrict__ w, float* __restrict__ dst, int O, int I, int npad, size_t total) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int e = i & 3;
  size_t t = i >> 2;
  int n = t % npad; t /= npad;
  int plane = t & 3; t >>= 2;
  int hl = t & 1; t >>= 1;
  int tap = t % 9;
  int chunk = (int)(t / 9);
  int c = chunk * kTcKC + plane * 4 + e;
  float v = (n < O && c < I) ? w[((size_t)n * I + c) * 9 + tap] : 0.f;
  float hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
  dst[i] = hl ? (v - hi) : hi;
}

int launch_pack_tc(const float* w_oihw, float* dst, int O, int I, int npad, cudaStream_t st) {
  size_t total = tc_packed_floats(I, npad);
  pack_tc_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(w_oihw, dst, O, I, npad, total);
  TMG_LAUNCH_CHECK();
  return TMG_OK;
}

}  // namespace tmg
