#!/usr/bin/env python
"""Kernel-level throughput of the OTHER BASELINE configs (1, 3, 4, 5 and variants) on one B200: algorithmic
bytes / CUDA-event time of back-to-back launches over rotating buffers (> L2), as a fraction of the measured
HBM peak.  These are parity-test cases, not bench.py lines; this script is the evidence that the same kernels
hold their roofline fraction off the headline shape.   python profiles/bench_configs.py [--json out.json]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from litehandnet_b200 import _lib as L, ops, synth  # noqa: E402

DEV = torch.device("cuda", 0)
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0


def timed(fns, reps=20, warm=3):
    """fns: list of callables over distinct buffer sets, cycled; returns ms per call."""
    for i in range(warm):
        fns[i % len(fns)]()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fns[i % len(fns)]()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


OVERLAP = "--overlap" in sys.argv     # consecutive launches work on disjoint buffer sets: LHN_FLAG_OVERLAP_PREVIOUS


def as_graphs(fns):
    """Each launch captured once into a CUDA graph and replayed: takes the Python ops layer (20-35 us of host time per
    call, profiles/r01_configs.txt note) out of a launch-bound measurement."""
    reps = []
    side = torch.cuda.Stream()
    for f in fns:
        f()                                      # allocations + cudaFuncSetAttribute outside the capture
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            f()
        reps.append(g.replay)
    return reps


def heatmap_case(name, B, K, H, W, dtype=torch.float32, flip=False, refine=L.REFINE_DARK, loss=True, pck=False,
                 sets=None, graph=False):
    esz = torch.empty((), dtype=dtype).element_size()
    nbytes = B * K * H * W * esz * (2 if flip else 1)
    sets = sets or max(2, int(600e6 // nbytes) + 1)
    fns = []
    for r in range(sets):
        hm, cen = synth.blob_heatmaps(B, K, H, W, seed=r, device=DEV, sigma=2.0 * W / 64, dtype=dtype)
        hf = synth.flipped_blob_heatmaps(cen, H, W, seed=r + 50, device=DEV, dtype=dtype) if flip else None
        c, s = synth.bbox_center_scale(B, seed=3, device=DEV)
        if pck:
            gt, mask, wh = synth.pck_inputs(cen, seed=5, device=DEV)
            cnt = torch.zeros((1 + 20 + 4) * K, dtype=torch.int64, device=DEV)
            fns.append(lambda hm=hm, c=c, s=s, gt=gt, mask=mask, wh=wh, cnt=cnt:
                       ops.decode_heatmap_pck(hm, L.MASK_NEG1, refine, c, s, gt, mask, wh, cnt, overlap_previous=OVERLAP))
        elif loss:
            j, v = synth.hand_joints(B, K, (4 * W, 4 * H), seed=2, device=DEV)
            render = dict(loss_mode=L.LOSS_DISTANCE_BALANCE, image_size=(4 * W, 4 * H), sigma=2.0 * W / 64, unbiased=True)
            out = {}
            fns.append(lambda hm=hm, hf=hf, c=c, s=s, j=j, v=v, out=out:
                       out.update(ops.fused_render_loss_decode(hm, L.MASK_NEG1, refine, L.XFORM_CENTER_SCALE, c, s, render,
                                                               j, v, hm_flip=hf, blur_ksize=11, out=out or None,
                                                               overlap_previous=OVERLAP)))
        else:
            out = {}
            fns.append(lambda hm=hm, hf=hf, c=c, s=s, out=out:
                       out.update(ops.decode_heatmap(hm, L.MASK_NEG1, refine, L.XFORM_CENTER_SCALE, c, s, hm_flip=hf,
                                                     blur_ksize=11, out=out or None, overlap_previous=OVERLAP)))
    ms = timed(as_graphs(fns) if graph else fns)
    return dict(config=name, bytes=nbytes, ms=ms, gbs=nbytes / ms / 1e6, frac=nbytes / ms / 1e6 / PEAK,
                samples_per_s=B / ms * 1e3)


def simdr_case(name, B, K, Lv, k=2):
    nbytes = 2 * B * K * Lv * 4
    fns = []
    for r in range(3):
        xv, yv = synth.simdr_vectors(B, K, Lv, seed=r, device=DEV, k=k)
        c, s = synth.bbox_center_scale(B, seed=3, device=DEV)
        fns.append(lambda xv=xv, yv=yv, c=c, s=s: ops.decode_simdr(xv, yv, k, c, s, overlap_previous=OVERLAP))
    ms = timed(fns)
    return dict(config=name, bytes=nbytes, ms=ms, gbs=nbytes / ms / 1e6, frac=nbytes / ms / 1e6 / PEAK,
                samples_per_s=B / ms * 1e3)


def main():
    H = heatmap_case
    cases = [
        lambda: H("cfg1 decode argmax + quarter offset, 64x21x64x64 f32 (22 MB/launch: launch-bound)", 64, 21, 64, 64,
                  refine=L.REFINE_SIGN, loss=False, sets=32),
        lambda: H("cfg1 shape at batch 1024", 1024, 21, 64, 64, refine=L.REFINE_SIGN, loss=False),
        lambda: H("cfg2 without flip (render + loss + DARK), 1024x21x64x64 f32", 1024, 21, 64, 64),
        lambda: H("cfg2 headline (flip), 1024x21x64x64 f32", 1024, 21, 64, 64, flip=True),
        lambda: H("cfg2 headline, bf16 inputs", 1024, 21, 64, 64, dtype=torch.bfloat16, flip=True),
        lambda: simdr_case("cfg3 SimDR decode k=2, 2 x 4096x21x512 f32", 4096, 21, 512),
        lambda: H("cfg4 MPII 16x64x64 decode + fused PCK/AUC/EPE counters, batch 1024/GPU", 1024, 16, 64, 64,
                  refine=L.REFINE_SIGN, pck=True),
        lambda: H("      the same decode without the fused counters (16x64x64, batch 1024)", 1024, 16, 64, 64,
                  refine=L.REFINE_SIGN, loss=False),
        lambda: H("cfg5 21x128x128 render + loss + DARK, batch 1024/GPU f32", 1024, 21, 128, 128),
        lambda: H("56x56 (33 reference configs), render + loss + DARK, 1024x21 f32", 1024, 21, 56, 56),
        # appended (the --only indices of the rows above are used by the probe scripts)
        lambda: H("cfg1, the same launches replayed as CUDA graphs (no Python between launches)", 64, 21, 64, 64,
                  refine=L.REFINE_SIGN, loss=False, sets=32, graph=True),
    ]
    only = sys.argv[sys.argv.index("--only") + 1].split(",") if "--only" in sys.argv else None   # e.g. --only 5 (cfg3)
    rows = [c() for i, c in enumerate(cases) if only is None or str(i) in only]
    print("launch overlap (LHN_FLAG_OVERLAP_PREVIOUS over rotating buffer sets):", "on" if OVERLAP else "off")
    for r in rows:
        print(f"{r['config']:95s} {r['ms'] * 1e3:9.1f} us  {r['gbs']:8.1f} GB/s  {r['frac'] * 100:5.1f}% of {PEAK:.0f}  "
              f"{r['samples_per_s'] / 1e6:7.2f} M samples/s")
    if "--json" in sys.argv:
        json.dump(rows, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)


if __name__ == "__main__":
    main()
