"""Host time per eager call of the decode entry (BASELINE config 1 shape and the reference's eval batch of 32):
ctypes wrapper vs the torch-extension shim vs a bound launcher vs a CUDA-graph replay.  The GPU work of a call is
~10 us, so wall time per call over many back-to-back calls is max(host, device) per call."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from litehandnet_b200 import _lib as L, fused, ops, synth  # noqa: E402


def wall(fn, n=3000):
    for _ in range(50):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6


def main():
    dev = "cuda"
    for B in (64, 32):
        hm, _ = synth.blob_heatmaps(B, 21, 64, 64, seed=1, device=dev)
        c, s = synth.bbox_center_scale(B, seed=3, device=dev, fixed=True)
        args = (hm, L.MASK_NEG1, L.REFINE_SIGN, L.XFORM_CENTER_SCALE, c, s)
        t_ext = wall(lambda: ops.decode_heatmap(*args)) if L.ext() is not None else float("nan")
        t_ct = wall(lambda: ops._decode_heatmap_ctypes(*args))
        b = fused.BoundDecodeStep(hm, c, s, L.MASK_NEG1, L.REFINE_SIGN, L.XFORM_CENTER_SCALE)
        t_bound = wall(b.launch)
        b.capture()
        t_graph = wall(b.replay)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(200):
            b.replay()
        e1.record()
        torch.cuda.synchronize()
        print(f"batch {B} x 21 x 64 x 64 decode ('default'): eager ops.decode_heatmap via torch extension {t_ext:6.1f} us/call, "
              f"via ctypes {t_ct:6.1f}, bound launcher {t_bound:6.1f}, graph replay {t_graph:6.1f} (device time per replay "
              f"{e0.elapsed_time(e1) / 200 * 1e3:5.1f} us)", flush=True)


if __name__ == "__main__":
    main()
