"""Generate a checkpoint ("workspace") zip with the REAL reference's ``saveWorkspace`` (``utils/utils.py:55-93``).

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_workspace.py

A small reference TMGlow, ``Adam(model.parameters(), amsgrad=True, weight_decay=1e-8)`` as in ``main.py:78``, two
optimizer steps on random gradients (non-trivial moments), saved as ``tests/golden/nsWorkspace7.zip``
(the configuration is repeated in tests/test_workspace_cpu.py).
"""
import os
import sys
import types
import warnings

import numpy as np
import torch

REF = "/root/reference/tmglow"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    sys.path.insert(0, REF)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from nn.tmGlow import TMGlow
        from utils.utils import saveWorkspace
    torch.manual_seed(77); np.random.seed(77)
    cfg = dict(in_features=4, out_features=3, enc_blocks=[2, 2], glow_blocks=[3, 3], cond_features=8, cglow_upscale=2,
               growth_rate=4, init_features=8, rec_features=8)
    m = TMGlow(cfg["in_features"], cfg["out_features"], cfg["enc_blocks"], cfg["glow_blocks"], cond_features=cfg["cond_features"],
               cglow_upscale=cfg["cglow_upscale"], growth_rate=cfg["growth_rate"], init_features=cfg["init_features"],
               rec_features=cfg["rec_features"])
    opt = torch.optim.Adam(m.parameters(), lr=1e-3 * 0.995 ** 7, weight_decay=1e-8, amsgrad=True)
    g = torch.Generator().manual_seed(5)
    for _ in range(2):
        for p in m.parameters():
            p.grad = 0.01 * torch.randn(p.shape, generator=g)
        opt.step()
    args = types.SimpleNamespace(ckpt_dir=HERE, device="cpu", beta=200, dx=5. / 64, dy=5. / 64, lr=1e-3, epoch_start=7, epochs=10,
                                 enc_blocks=cfg["enc_blocks"], glow_blocks=cfg["glow_blocks"], batch_size=64, notes="golden")
    saveWorkspace(args, m, opt, file_id=7)
    print(os.path.getsize(os.path.join(HERE, "nsWorkspace7.zip")), "bytes")


if __name__ == "__main__":
    main()
