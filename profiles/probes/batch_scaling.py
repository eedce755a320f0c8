#!/usr/bin/env python
"""Fixed cost per launch vs marginal bandwidth of the team kernel on the shapes that sit below the headline's
roofline fraction.  Each shape is timed at batches 1024 / 2048 / 4096 (same kernel, same team layout; only the number
of planes per team grows) and t(B) = t0 + B / rate is fitted by least squares: t0 is what one launch costs whatever
its size (launch gap, first loads, finish skew, drain of the last planes), `rate` the marginal samples/s, i.e. the
bandwidth the kernel body sustains.   python profiles/probes/batch_scaling.py [--overlap] [--only 0,2]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "profiles"))
import bench_configs as BC  # noqa: E402
from litehandnet_b200 import _lib as L  # noqa: E402

H = BC.heatmap_case
SHAPES = [
    ("no flip f32 64x64 (render + loss + DARK)", dict(K=21, H=64, W=64), 344064),
    ("headline f32 + flip 64x64", dict(K=21, H=64, W=64, flip=True), 688128),
    ("bf16 + flip 64x64", dict(K=21, H=64, W=64, flip=True, dtype=torch.bfloat16), 344064),
    ("56x56 f32 (render + loss + DARK)", dict(K=21, H=56, W=56), 21 * 56 * 56 * 4),
    ("cfg4 16x64x64 decode + fused counters", dict(K=16, H=64, W=64, refine=L.REFINE_SIGN, pck=True), 262144),
    ("cfg4 shape without counters", dict(K=16, H=64, W=64, refine=L.REFINE_SIGN, loss=False), 262144),
    ("cfg5 128x128 f32 (render + loss + DARK)", dict(K=21, H=128, W=128), 21 * 128 * 128 * 4),
]


def main():
    print("launch overlap:", "on" if BC.OVERLAP else "off")
    batches = [1024, 2048, 4096]
    only = sys.argv[sys.argv.index("--only") + 1].split(",") if "--only" in sys.argv else None
    for i, (name, kw, bps) in enumerate(SHAPES):
        if only is not None and str(i) not in only:
            continue
        ts = []
        batches = [256, 512, 1024] if kw["H"] >= 128 else [1024, 2048, 4096]
        for B in batches:
            r = H(name, B, sets=2 if B * bps > 300e6 else 3, **kw)
            ts.append(r["ms"] * 1e3)
            torch.cuda.empty_cache()
        A = np.stack([np.ones(len(batches)), np.array(batches, dtype=np.float64)], 1)
        (t0, slope), *_ = np.linalg.lstsq(A, np.array(ts), rcond=None)
        marg = bps / slope / 1e3          # GB/s: bytes per sample / us per sample
        print(f"{name:42s} " + "  ".join(f"B={b}: {t:7.1f} us ({b * bps / t / 1e3 / BC.PEAK * 100:5.1f}%)"
                                         for b, t in zip(batches, ts)) +
              f"  | fixed {t0:5.1f} us/launch, marginal {marg:6.0f} GB/s = {marg / BC.PEAK * 100:5.1f}% of {BC.PEAK:.0f}")


if __name__ == "__main__":
    main()
