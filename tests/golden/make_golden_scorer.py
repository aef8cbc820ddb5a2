"""Generates tests/golden/scorer_*.npz from the UNMODIFIED reference scorer (run in the build container)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
from transkun.LayersTransformer import ScaledInnerProductIntervalScorer  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
for name, B, P, T, size, seed in [("scorer_B1_P3_T40_D64", 1, 3, 40, 64, 0), ("scorer_B2_P5_T70_D256", 2, 5, 70, 256, 1),
                                   # an odd number of 32-wide K chunks and a last track group of one (round 2 kernel)
                                   ("scorer_B1_P9_T130_D96", 1, 9, 130, 96, 2)]:
    torch.manual_seed(seed)
    m = ScaledInnerProductIntervalScorer(size, 1).eval()
    ctx = torch.randn(B, P, T, size)
    with torch.no_grad():
        S, b = m(ctx)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), ctx=ctx.numpy(), weight=m.map[0].weight.detach().numpy(),
                        bias=m.map[0].bias.detach().numpy(), S=S.numpy(), b=b.numpy())
    print(name, S.shape, float(S.abs().max()))
