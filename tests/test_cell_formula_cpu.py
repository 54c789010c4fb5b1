"""The ConvLSTM cell update of lstm_gate_f16.cu (`gt_cell`) shares denominators between its sigmoid / tanh quotients to use
7 instead of 10 special-function operations:

    c' = sigm(f) c + sigm(i) tanh(g) = (c A B + (e^2g - 1) D) / (D A B),   A = 1 + e^-i, D = 1 + e^-f, B = e^2g + 1
    h' = sigm(o) tanh(c')            = (e^2c' - 1) / ((1 + e^-o)(e^2c' + 1))

with the pre-activations clamped to +-28 (gates) / +-14 (tanh arguments) so that the products stay finite in fp32.  This
restates the same arithmetic in numpy float32 (exact exp2 / division instead of ex2.approx / rcp.approx: the algebra and the
clamps are what is checked here; the hardware approximations add ~2 ulp each and are covered by the GPU parity tests) and
compares it with the reference formulation (nn/modules/convLSTM.py:76-83) evaluated in float64."""
import numpy as np


def gt_cell_f32(pi, pf, po, pg, cp):
    f = np.float32
    l2e = f(1.4426950408889634)
    pi = np.clip(pi, f(-28), f(28)); pf = np.clip(pf, f(-28), f(28)); po = np.clip(po, f(-28), f(28))
    pg = np.clip(pg, f(-14), f(14))
    ei, ef, eg = np.exp2(-l2e * pi), np.exp2(-l2e * pf), np.exp2(f(2) * l2e * pg)
    A, D, B, N = f(1) + ei, f(1) + ef, eg + f(1), eg - f(1)
    AB = A * B
    cn = (cp * AB + N * D) / (D * AB)
    cc = np.clip(cn, f(-14), f(14))
    eo, ec = np.exp2(-l2e * po), np.exp2(f(2) * l2e * cc)
    hn = (ec - f(1)) / ((f(1) + eo) * (ec + f(1)))
    return cn.astype(np.float32), hn.astype(np.float32)


def cell_ref_f64(pi, pf, po, pg, cp):
    pi, pf, po, pg, cp = (np.asarray(v, dtype=np.float64) for v in (pi, pf, po, pg, cp))
    sig = lambda x: 1.0 / (1.0 + np.exp(-x))
    cn = sig(pf) * cp + sig(pi) * np.tanh(pg)
    return cn, sig(po) * np.tanh(cn)


def test_shared_denominator_cell_matches_reference_formulation():
    rng = np.random.default_rng(0)
    n = 200000
    pre = [rng.normal(0, 3, n).astype(np.float32) for _ in range(4)]
    cp = rng.normal(0, 2, n).astype(np.float32)
    with np.errstate(over="raise", invalid="raise", divide="raise"):
        cn, hn = gt_cell_f32(*pre, cp)
    cr, hr = cell_ref_f64(*pre, cp)
    # absolute error of the state relative to its scale, of the output absolutely (|h| < 1)
    assert np.max(np.abs(cn - cr) / (1.0 + np.abs(cr))) < 5e-7
    assert np.max(np.abs(hn - hr)) < 5e-7


def test_cell_is_finite_and_saturated_at_extreme_preactivations():
    f = np.float32
    ext = np.array([-1e30, -1e4, -100, -40, -28, -14, 0, 14, 28, 40, 100, 1e4, 1e30], dtype=np.float32)
    pi, pf, po, pg = (v.ravel() for v in np.meshgrid(ext, ext, ext, ext, indexing="ij"))
    for c0 in (f(0), f(3.5), f(-7), f(1e4), f(-1e4)):
        cp = np.full(pi.shape, c0, dtype=np.float32)
        with np.errstate(over="raise", invalid="raise", divide="raise"):
            cn, hn = gt_cell_f32(pi, pf, po, pg, cp)
        assert np.isfinite(cn).all() and np.isfinite(hn).all()
        cr, hr = cell_ref_f64(np.clip(pi, -60, 60), np.clip(pf, -60, 60), np.clip(po, -60, 60), np.clip(pg, -60, 60), cp)
        # beyond the clamps sigmoid / tanh are saturated to fp32 precision: the clamped evaluation is the exact one to ~1e-6 relative
        assert np.max(np.abs(cn - cr) / (1.0 + np.abs(cr))) < 2e-6
        assert np.max(np.abs(hn - hr)) < 2e-6
