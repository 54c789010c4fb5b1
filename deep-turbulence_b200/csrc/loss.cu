// TMGLowLoss (reference nn/trainFlowParallel.py:121-177) fused with its own backward: one pass over the predicted
// block yPred[B,T,3,H,W] yields the reverse-KL training loss
//     beta * (vPres + vDiv + vL1 + vRMS) + mean(logp) / (ln 2 * 3HW)
// and d loss / d yPred.  The PDE residuals are PhysConstrainedLES.calcPressurePoisson / calcDivergence
// (pc/physicsConstrained.py:43-94) on the un-normalised fields with the 3x3 smoothed finite differences of
// pc/grad1Filter.py:34-47 and pc/grad2Filter.py:33-45.  Only residuals at interior points enter the loss
// (trainFlowParallel.py:164,177), so the zero padding of the reference's filters never contributes; the divergence
// replicates the edge columns (physicsConstrained.py:58), which shows up here as a column clamp.
// HBM-bound streaming kernel: per (sample, 16x32 tile) CTA, y is read from L2/HBM and the gradient written once.
// Deterministic: per-CTA partial sums, fixed-order fp64 finish (no atomics).
#include "common.cuh"

namespace tmg {

constexpr int kLTW = 32, kLTH = 16, kLThreads = 256;
constexpr int kYW = kLTW + 4, kYH = kLTH + 4;      // un-normalised field tile, halo 2
constexpr int kCW = kLTW + 2, kCH = kLTH + 2;      // residual-adjoint tile, halo 1

struct LossArgs {
  const float* y; const float* tgt; const float* trms; const float* mu; const float* sd;
  int B, T, H, W;
  float dx, dy, dx2, dy2, dxdy;
  float c_pres, c_div, c_l1, c_rms;                // beta * 2 / (number of elements of the term's mean)
  float* gy;
  float* part;                                     // [ctas][4]: sums of pstar^2, ustar^2, (y-t)^2, (rms-rms_t)^2
};

__global__ void __launch_bounds__(kLThreads)
tmglow_loss_kernel(LossArgs a) {
  __shared__ float s_y[3][kYH][kYW];
  __shared__ float s_c[6][kCH][kCW];               // adjoint weights of ux, uy, vx, vy, laplace(p); divergence
  __shared__ float s_m[3][kLTH * kLTW], s_k[3][kLTH * kLTW], s_kd[3][kLTH * kLTW];
  __shared__ float s_red[kLThreads / 32][4];
  const int tid = threadIdx.x, b = blockIdx.z;
  const int i0 = blockIdx.y * kLTH, j0 = blockIdx.x * kLTW;
  const int H = a.H, W = a.W, T = a.T;
  const size_t plane = (size_t)H * W, frame = 3 * plane;
  const float* yb = a.y + (size_t)b * T * frame;
  const float* tb = a.tgt + (size_t)b * T * frame;
  float acc_p = 0.f, acc_d = 0.f, acc_l = 0.f, acc_r = 0.f;

  // ---- phase A: time mean / RMS of the fluctuation per (channel, pixel)  (trainFlowParallel.py:144-145)
  for (int it = tid; it < 3 * kLTH * kLTW; it += kLThreads) {
    const int c = it / (kLTH * kLTW), p = it - c * (kLTH * kLTW);
    const int gi = i0 + p / kLTW, gj = j0 + p % kLTW;
    float m = 0.f, k = 0.f, kd = 0.f;
    if (gi < H && gj < W) {
      const float* yp = yb + c * plane + (size_t)gi * W + gj;
      float s = 0.f;
      for (int t = 0; t < T; ++t) s += yp[t * frame];
      m = s / (float)T;
      float v = 0.f, sd = 0.f;
      for (int t = 0; t < T; ++t) { const float d = yp[t * frame] - m; v = fmaf(d, d, v); sd += d; }
      const float r = sqrtf(v / (float)T);
      const float diff = r - a.trms[((size_t)b * 3 + c) * plane + (size_t)gi * W + gj];
      acc_r = fmaf(diff, diff, acc_r);
      // d vRMS / d y_t = k * ((y_t - m) - mean_s(y_s - m)): the second term is the path through the time mean; it is
      // zero in exact arithmetic and cancels the rounding error of m when the fluctuation is tiny (autograd has it too).
      // inf/nan at r == 0, like autograd
      k = a.c_rms * diff / ((float)T * r);
      kd = k * (sd / (float)T);
    }
    s_m[c][p] = m; s_k[c][p] = k; s_kd[c][p] = kd;
  }
  const float sd0 = a.sd[0], sd1 = a.sd[1], sd2 = a.sd[2], mu0 = a.mu[0], mu1 = a.mu[1], mu2 = a.mu[2];

  for (int t = 0; t < T; ++t) {
    const float* yt = yb + (size_t)t * frame;
    __syncthreads();                                // previous time step done with s_y / s_c (and phase A with s_m / s_k)
    // ---- un-normalised fields, halo 2 (trainFlowParallel.py:160,173)
    for (int it = tid; it < 3 * kYH * kYW; it += kLThreads) {
      const int c = it / (kYH * kYW), r = (it / kYW) % kYH, cc = it % kYW;
      const int gi = i0 - 2 + r, gj = j0 - 2 + cc;
      float v = 0.f;
      if (gi >= 0 && gi < H && gj >= 0 && gj < W)
        v = fmaf(c == 0 ? sd0 : (c == 1 ? sd1 : sd2), yt[c * plane + (size_t)gi * W + gj], c == 0 ? mu0 : (c == 1 ? mu1 : mu2));
      s_y[c][r][cc] = v;
    }
    __syncthreads();
    // ---- residuals and their adjoint weights on the tile + halo 1
    for (int it = tid; it < kCH * kCW; it += kLThreads) {
      const int pi = it / kCW, pj = it - pi * kCW;
      const int gi = i0 - 1 + pi, gj = j0 - 1 + pj;
      const int r = pi + 1, cc = pj + 1;            // position inside s_y
      float wux = 0.f, wuy = 0.f, wvx = 0.f, wvy = 0.f, wp = 0.f, wd = 0.f;
      const bool rows = gi >= 1 && gi <= H - 2;
      const bool own = pi >= 1 && pi <= kLTH && pj >= 1 && pj <= kLTW;     // counted by this CTA (not its halo)
      if (rows && gj >= 0 && gj <= W - 1) {         // divergence: edge columns replicated (physicsConstrained.py:58)
        const int cm = max(gj - 1, 0) - (j0 - 2), cp = min(gj + 1, W - 1) - (j0 - 2);
        const float (*u)[kYW] = s_y[0];
        const float (*v)[kYW] = s_y[1];
        const float gx = ((u[r - 1][cp] - u[r - 1][cm]) + 2.f * (u[r][cp] - u[r][cm]) + (u[r + 1][cp] - u[r + 1][cm])) * 0.125f / a.dx;
        const float gy = ((v[r + 1][cm] - v[r - 1][cm]) + 2.f * (v[r + 1][cc] - v[r - 1][cc]) + (v[r + 1][cp] - v[r - 1][cp])) * 0.125f / a.dy;
        const float s = a.dx * (gy + gx);
        const float cs = fminf(fmaxf(s, -1.f), 1.f);
        if (own) acc_d = fmaf(cs, cs, acc_d);
        wd = (s >= -1.f && s <= 1.f) ? a.c_div * cs * a.dx : 0.f;
      }
      if (rows && gj >= 1 && gj <= W - 2) {         // pressure Poisson residual (physicsConstrained.py:83-94)
        const float (*u)[kYW] = s_y[0];
        const float (*v)[kYW] = s_y[1];
        const float (*p)[kYW] = s_y[2];
        const int cm = cc - 1, cp = cc + 1;
        const float ux = ((u[r - 1][cp] - u[r - 1][cm]) + 2.f * (u[r][cp] - u[r][cm]) + (u[r + 1][cp] - u[r + 1][cm])) * 0.125f / a.dx;
        const float vx = ((v[r - 1][cp] - v[r - 1][cm]) + 2.f * (v[r][cp] - v[r][cm]) + (v[r + 1][cp] - v[r + 1][cm])) * 0.125f / a.dx;
        const float uy = ((u[r + 1][cm] - u[r - 1][cm]) + 2.f * (u[r + 1][cc] - u[r - 1][cc]) + (u[r + 1][cp] - u[r - 1][cp])) * 0.125f / a.dy;
        const float vy = ((v[r + 1][cm] - v[r - 1][cm]) + 2.f * (v[r + 1][cc] - v[r - 1][cc]) + (v[r + 1][cp] - v[r - 1][cp])) * 0.125f / a.dy;
        const float pxx = ((p[r - 1][cm] - 2.f * p[r - 1][cc] + p[r - 1][cp]) + 2.f * (p[r][cm] - 2.f * p[r][cc] + p[r][cp]) +
                           (p[r + 1][cm] - 2.f * p[r + 1][cc] + p[r + 1][cp])) * 0.25f / a.dx2;
        const float pyy = ((p[r - 1][cm] - 2.f * p[r][cm] + p[r + 1][cm]) + 2.f * (p[r - 1][cc] - 2.f * p[r][cc] + p[r + 1][cc]) +
                           (p[r - 1][cp] - 2.f * p[r][cp] + p[r + 1][cp])) * 0.25f / a.dy2;
        const float q = a.dxdy * ((pxx + pyy) + (ux * ux + 2.f * uy * vx + vy * vy));
        const float cq = fminf(fmaxf(q, -1.f), 1.f);
        if (own) acc_p = fmaf(cq, cq, acc_p);
        const float gq = (q >= -1.f && q <= 1.f) ? a.c_pres * cq * a.dxdy : 0.f;
        wux = 2.f * gq * ux; wuy = 2.f * gq * vx; wvx = 2.f * gq * uy; wvy = 2.f * gq * vy; wp = gq;
      }
      s_c[0][pi][pj] = wux; s_c[1][pi][pj] = wuy; s_c[2][pi][pj] = wvx; s_c[3][pi][pj] = wvy; s_c[4][pi][pj] = wp;
      s_c[5][pi][pj] = wd;
    }
    __syncthreads();
    // ---- gradient: stencil adjoints (gather form) + MSE + RMS terms
    for (int p = tid; p < kLTH * kLTW; p += kLThreads) {
      const int li = p / kLTW, lj = p - li * kLTW;
      const int gi = i0 + li, gj = j0 + lj;
      if (gi >= H || gj >= W) continue;
      float gu = 0.f, gv = 0.f, gp = 0.f;
      float ax = 0.f, ay = 0.f, bx = 0.f, by = 0.f, dxu = 0.f, dyv = 0.f, lx = 0.f, ly = 0.f;
#pragma unroll
      for (int ta = 0; ta < 3; ++ta) {
#pragma unroll
        for (int tb2 = 0; tb2 < 3; ++tb2) {
          // the residual point (gi - ta + 1, gj - tb2 + 1) read this pixel through tap (ta, tb2)
          const int r = li + 2 - ta, cc = lj + 2 - tb2;
          const float rw = ta == 1 ? 2.f : 1.f, cw = tb2 == 1 ? 2.f : 1.f;
          const float wh = (tb2 == 0 ? -rw : (tb2 == 2 ? rw : 0.f));        // grad1 weight_h[ta][tb2] * 8
          const float wv = (ta == 0 ? -cw : (ta == 2 ? cw : 0.f));          // grad1 weight_v[ta][tb2] * 8
          const float w2h = rw * (tb2 == 1 ? -2.f : 1.f);                   // grad2 weight_h * 4
          const float w2v = cw * (ta == 1 ? -2.f : 1.f);                    // grad2 weight_v * 4
          ax = fmaf(wh, s_c[0][r][cc], ax); ay = fmaf(wv, s_c[1][r][cc], ay);
          bx = fmaf(wh, s_c[2][r][cc], bx); by = fmaf(wv, s_c[3][r][cc], by);
          lx = fmaf(w2h, s_c[4][r][cc], lx); ly = fmaf(w2v, s_c[4][r][cc], ly);
          dxu = fmaf(wh, s_c[5][r][cc], dxu); dyv = fmaf(wv, s_c[5][r][cc], dyv);
        }
      }
      // replicated edge columns of the divergence: the clamped tap of the edge point lands on the edge pixel itself
      if (gj == 0 || gj == W - 1) {
        const int tb2 = gj == 0 ? 0 : 2;
        // (W == 1 is rejected by the host; for gj == 0 == W-1 this would need both)
#pragma unroll
        for (int ta = 0; ta < 3; ++ta) {
          const int r = li + 2 - ta, cc = lj + 1;
          const float rw = ta == 1 ? 2.f : 1.f;
          const float wh = tb2 == 0 ? -rw : rw;
          const float wv = ta == 0 ? -1.f : (ta == 2 ? 1.f : 0.f);
          dxu = fmaf(wh, s_c[5][r][cc], dxu); dyv = fmaf(wv, s_c[5][r][cc], dyv);
        }
      }
      gu = (ax + dxu) * 0.125f / a.dx + ay * 0.125f / a.dy;
      gv = bx * 0.125f / a.dx + (by + dyv) * 0.125f / a.dy;
      gp = lx * 0.25f / a.dx2 + ly * 0.25f / a.dy2;
      const size_t off = (size_t)gi * W + gj;
      const float g3[3] = {gu * sd0, gv * sd1, gp * sd2};
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float yv = yt[c * plane + off];
        const float e = yv - tb[(size_t)t * frame + c * plane + off];
        acc_l = fmaf(e, e, acc_l);
        if (a.gy) a.gy[((size_t)b * T + t) * frame + c * plane + off] = g3[c] + a.c_l1 * e + (s_k[c][p] * (yv - s_m[c][p]) - s_kd[c][p]);
      }
    }
  }
  // ---- per-CTA partial sums, fixed order
  float v4[4] = {acc_p, acc_d, acc_l, acc_r};
#pragma unroll
  for (int k = 0; k < 4; ++k)
    for (int o = 16; o; o >>= 1) v4[k] += __shfl_xor_sync(0xffffffffu, v4[k], o);
  if ((tid & 31) == 0)
    for (int k = 0; k < 4; ++k) s_red[tid >> 5][k] = v4[k];
  __syncthreads();
  if (tid < 4) {
    float s = 0.f;
    for (int w = 0; w < kLThreads / 32; ++w) s += s_red[w][tid];
    const size_t cta = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    a.part[cta * 4 + tid] = s;
  }
}

// loss = beta * (vPres + vDiv + vL1 + vRMS) + mean(logp) / (ln2 * 3HW)   (trainFlowParallel.py:147-151)
__global__ void __launch_bounds__(256)
tmglow_loss_finish_kernel(const float* part, int ncta, const float* logp, int nlogp, double n_pres, double n_div, double n_l1,
                          double n_rms, double n_out, double beta, float* loss, float* terms, float* g_logp) {
  __shared__ double s[256];
  double out[5];
  for (int k = 0; k < 5; ++k) {
    double v = 0.0;
    if (k < 4) for (int i = threadIdx.x; i < ncta; i += 256) v += (double)part[(size_t)i * 4 + k];
    else for (int i = threadIdx.x; i < nlogp; i += 256) v += (double)logp[i];
    s[threadIdx.x] = v;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
      if ((int)threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
      __syncthreads();
    }
    out[k] = s[0];
    __syncthreads();
  }
  const double ln2 = 0.6931471805599453;
  const double vp = out[0] / n_pres, vd = out[1] / n_div, vl = out[2] / n_l1, vr = out[3] / n_rms;
  const double ne = out[4] / (double)nlogp / ln2 / n_out;
  if (threadIdx.x == 0) {
    if (terms) { terms[0] = (float)vp; terms[1] = (float)vd; terms[2] = (float)vl; terms[3] = (float)vr; terms[4] = (float)ne; }
    loss[0] = (float)(beta * (vp + vd + vl + vr) + ne);
  }
  if (g_logp) {
    const float g = (float)(1.0 / ((double)nlogp * ln2 * n_out));
    for (int i = threadIdx.x; i < nlogp; i += 256) g_logp[i] = g;
  }
}

static int loss_ctas(int B, int H, int W) { return cdiv(W, kLTW) * cdiv(H, kLTH) * B; }

}  // namespace tmg

using namespace tmg;

extern "C" size_t tmg_tmglow_loss_workspace_bytes(int B, int T, int H, int W) {
  (void)T;
  if (B <= 0 || H <= 0 || W <= 0) return 0;
  return (size_t)loss_ctas(B, H, W) * 4 * sizeof(float);
}

extern "C" int tmg_tmglow_loss(const float* y_pred, const float* logp, const float* target, const float* target_rms,
                               const float* out_mu, const float* out_std, int B, int T, int H, int W, int n_logp, double dx,
                               double dy, double beta, float* loss, float* terms, float* g_y, float* g_logp, void* workspace,
                               size_t workspace_bytes, void* stream) {
  if (!y_pred || !logp || !target || !target_rms || !out_mu || !out_std || !loss || !workspace) {
    set_error("tmg_tmglow_loss: null pointer");
    return TMG_ERR_NULL;
  }
  if (B <= 0 || T <= 0 || H < 3 || W < 3 || n_logp <= 0 || B > 65535) {
    set_error("tmg_tmglow_loss: bad shape B=%d T=%d H=%d W=%d n_logp=%d (needs H, W >= 3: residuals live on interior points)", B, T, H, W, n_logp);
    return TMG_ERR_BAD_SHAPE;
  }
  if (!(dx > 0.0) || !(dy > 0.0)) {
    set_error("tmg_tmglow_loss: dx, dy must be positive");
    return TMG_ERR_BAD_CONFIG;
  }
  const int ncta = loss_ctas(B, H, W);
  if (workspace_bytes < (size_t)ncta * 4 * sizeof(float)) {
    set_error("tmg_tmglow_loss: workspace too small");
    return TMG_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const double n_pres = (double)B * T * (H - 2) * (W - 2), n_div = (double)B * T * (H - 2) * W;
  const double n_l1 = (double)B * T * 3 * H * W, n_rms = (double)B * 3 * H * W;
  LossArgs a;
  a.y = y_pred; a.tgt = target; a.trms = target_rms; a.mu = out_mu; a.sd = out_std;
  a.B = B; a.T = T; a.H = H; a.W = W;
  a.dx = (float)dx; a.dy = (float)dy; a.dx2 = (float)(dx * dx); a.dy2 = (float)(dy * dy); a.dxdy = (float)(dx * dy);
  a.c_pres = (float)(beta * 2.0 / n_pres); a.c_div = (float)(beta * 2.0 / n_div);
  a.c_l1 = (float)(beta * 2.0 / n_l1); a.c_rms = (float)(beta * 2.0 / n_rms);
  a.gy = g_y; a.part = (float*)workspace;
  {
    ProfScope ps(st, PROF_MISC, 0.0, (double)(g_y ? 5 : 4) * n_l1 * 4.0);
    dim3 grid(cdiv(W, kLTW), cdiv(H, kLTH), B);
    tmglow_loss_kernel<<<grid, kLThreads, 0, st>>>(a);
    TMG_LAUNCH_CHECK();
    tmglow_loss_finish_kernel<<<1, 256, 0, st>>>(a.part, ncta, logp, n_logp, n_pres, n_div, n_l1, n_rms, 3.0 * H * W, beta, loss,
                                                 terms, g_logp);
    TMG_LAUNCH_CHECK();
  }
  return TMG_OK;
}
