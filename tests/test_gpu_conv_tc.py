"""The tcgen05 implicit-GEMM convolution (conv3x3_tc.cu) in isolation, against an fp64 torch CPU
convolution: same operator as nn.Conv2d(k=3, padding=1) of the reference call sites
(convLSTM.py:44,129; flowUtils.py:229,246).  Tolerances relative to max|out|:
    fp32 (CUDA-core FMA)  1e-5      tf32x3 (tcgen05, 3xTF32)  1e-5      tf32 (tcgen05, single pass)  4e-3
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

SHAPES = [
    # B, H,  W,  Cin, Cout, relu_in, replicate, act, bias
    (2, 32, 64, 102, 256, False, False, 0, True),    # ConvLSTM gate conv, block 0
    (2, 32, 64, 102, 38, False, False, 1, True),     # LSTM_out_conv, block 0
    (3, 32, 64, 40, 12, True, True, 0, True),        # Conv2dZeros, block 0
    (2, 16, 32, 46, 24, True, True, 0, True),        # Conv2dZeros, block 1
    (5, 8, 16, 58, 48, True, True, 2, True),         # Conv2dZeros, block 2 (+hardtanh epilogue)
    (1, 4, 8, 7, 5, False, False, 0, False),         # ragged channels, tiny map
    (2, 6, 10, 16, 16, True, False, 1, False),
]
TOL = {"fp32": 1e-5, "tf32x3": 1e-5, "tf32": 4e-3}


def _ref(x, w, b, relu_in, replicate, act):
    x, w = x.double(), w.double()
    if relu_in:
        x = F.relu(x)
    if replicate:
        y = F.conv2d(F.pad(x, (1, 1, 1, 1), mode="replicate"), w, None if b is None else b.double())
    else:
        y = F.conv2d(x, w, None if b is None else b.double(), padding=1)
    if act == 1:
        y = F.relu(y)
    elif act == 2:
        y = F.hardtanh(y, -2.0, 1.6094379124341003)
    return y


@pytest.mark.parametrize("mode", ["fp32", "tf32x3", "tf32"])
@pytest.mark.parametrize("shape", SHAPES)
def test_conv3x3_modes(shape, mode):
    from tmglow_b200 import ops
    B, H, W, Cin, Cout, relu_in, replicate, act, has_bias = shape
    g = torch.Generator().manual_seed(Cin * 1000 + Cout)
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / (3.0 * Cin ** 0.5)
    b = torch.randn(Cout, generator=g) if has_bias else None
    ref = _ref(x, w, b, relu_in, replicate, act)
    dev = torch.device("cuda:0")
    out = ops.conv3x3(x.to(dev), w.to(dev), None if b is None else b.to(dev), relu_in, replicate, act, mode)
    torch.cuda.synchronize()
    err = (out.cpu().double() - ref).abs().max().item() / ref.abs().max().item()
    assert err <= TOL[mode], "%s %s: rel err %.3e" % (mode, shape, err)
