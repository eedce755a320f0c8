#!/usr/bin/env python
"""Kernel-level throughput of the un-fused / training-side kernels: algorithmic bytes per launch /
CUDA-event time over rotating buffers (> L2), as a fraction of the measured HBM copy peak, at B = 1024 x 21 x 64 x 64.

  loss forward   lhn_loss_mse_multi (one launch)  vs  lhn_loss_partials -> lhn_loss_reduce -> lhn_loss_finalize (three)
  loss backward  lhn_loss_backward (2 reads + 1 write per element), lhn_render_loss_backward (1 read + 1 write)
  SimDR          lhn_simdr_smoothl1 (4 vector reads), lhn_simdr_smoothl1_backward (4 reads + 2 writes)
  render         lhn_render_targets (1 write), lhn_render_simdr
  metrics        lhn_pck_accumulate, lhn_evaluate_pck (2 heatmap reads)
  SRHandNet      four scales in one launch vs the old 4 x 3
python profiles/bench_kernels.py [--json out.json]"""
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
    """(device ms, eager ms) per call.  Eager: events around a python loop of calls — what a caller sees, bounded below by
    the host's launch rate (allocations + ctypes, 20-50 us).  Device: the same `reps` calls captured into one CUDA graph
    and replayed, so the events see the kernels back to back; None when a wrapper cannot be captured."""
    for i in range(warm):
        fns[i % len(fns)]()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fns[i % len(fns)]()
    e1.record()
    torch.cuda.synchronize()
    eager = e0.elapsed_time(e1) / reps
    dev = None
    if os.environ.get("LHN_BENCH_GRAPH", "1") != "0":
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for i in range(reps):
                    fns[i % len(fns)]()
            g.replay()
            torch.cuda.synchronize()
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            dev = e0.elapsed_time(e1) / reps
            del g
        except Exception as e:                                        # a wrapper that synchronises cannot be captured
            print(f"# graph capture failed ({type(e).__name__}); eager time only", flush=True)
            torch.cuda.synchronize()
    return dev, eager


def row(name, nbytes, fns):
    dev, eager = timed(fns)
    ms = dev if dev is not None else eager
    return dict(kernel=name, bytes=nbytes, us=ms * 1e3, gbs=nbytes / ms / 1e6, frac=nbytes / ms / 1e6 / PEAK,
                us_eager=eager * 1e3, timed="graph replay" if dev is not None else "eager loop")


def main():
    B, K, H, W = 1024, 21, 64, 64
    R = 3
    plane_bytes = B * K * H * W * 4
    sets = []
    for r in range(R):
        j, v = synth.hand_joints(B, K, (256, 256), seed=r, device=DEV)
        t, tw = ops.render_targets(j, v, (256, 256), (W, H), 2.0, True)
        o = t + torch.randn(t.shape, generator=torch.Generator(device=DEV).manual_seed(r), device=DEV) * 0.05
        sets.append((o, t, tw, j, v))
    rows = []
    mode = L.LOSS_DISTANCE_BALANCE
    rows.append(row("loss fwd, ONE launch (lhn_loss_mse_multi), DistanceLoss balance", 2 * plane_bytes,
                    [lambda s=s: ops.loss_mse_multi([s[0]], [s[1]], [s[2]], mode) for s in sets]))

    def three(s):
        p = ops.loss_partials(s[0], s[1], s[2], mode)
        return ops.loss_finalize(ops.loss_reduce(p), mode)

    rows.append(row("loss fwd, three launches (partials -> reduce -> finalize)", 2 * plane_bytes, [lambda s=s: three(s) for s in sets]))
    sums = ops.loss_mse_multi([sets[0][0]], [sets[0][1]], [sets[0][2]], mode)[1][0]
    rows.append(row("loss bwd (lhn_loss_backward): 2 reads + 1 write", 3 * plane_bytes,
                    [lambda s=s: ops.loss_backward(s[0], s[1], s[2], mode, sums) for s in sets]))
    render = dict(loss_mode=mode, image_size=(256, 256), sigma=2.0, unbiased=True)
    rows.append(row("loss bwd, target re-rendered (lhn_render_loss_backward): 1 read + 1 write", 2 * plane_bytes,
                    [lambda s=s: ops.render_loss_backward(s[0], s[3], s[4], render, sums) for s in sets]))
    rows.append(row("render targets (lhn_render_targets): 1 write", plane_bytes,
                    [lambda s=s: ops.render_targets(s[3], s[4], (256, 256), (W, H), 2.0, True) for s in sets]))
    # bf16 forward
    bsets = [(s[0].bfloat16(), s[1].bfloat16(), s[2]) for s in sets]
    rows.append(row("loss fwd, ONE launch, bf16 tensors", plane_bytes,
                    [lambda s=s: ops.loss_mse_multi([s[0]], [s[1]], [s[2]], mode) for s in bsets]))
    del bsets
    # SRHandNet: 4 scales [16,16,32,64] x 24 channels, batch 64 (the reference's training batch)
    Bs = 64
    sr = []
    for r in range(R):
        outs, tgts, ws = [], [], []
        for hw in (16, 16, 32, 64):
            j, v = synth.hand_joints(Bs, 21, (256, 256), seed=10 * r + hw, device=DEV)
            t, tw = ops.render_targets(j, v, (256, 256), (hw, hw), 2.0 if hw > 16 else 1.0, True)
            outs.append(t + 0.01); tgts.append(t); ws.append(tw)
        sr.append((outs, tgts, ws))
    sr_bytes = 2 * sum(Bs * 21 * hw * hw * 4 for hw in (16, 16, 32, 64))
    rows.append(row("SRHandNetLoss, 4 scales in ONE launch (batch 64)", sr_bytes,
                    [lambda s=s: ops.loss_mse_multi(s[0], s[1], s[2], mode, loss_weights=[0.3, 0.3, 0.5, 1.0]) for s in sr]))

    def twelve(s):
        tot = 0
        for o, t, w in zip(*s):
            tot = tot + ops.loss_finalize(ops.loss_reduce(ops.loss_partials(o, t, w, mode)), mode)
        return tot

    rows.append(row("SRHandNetLoss, 4 x 3 launches (round-1 path, batch 64)", sr_bytes, [lambda s=s: twelve(s) for s in sr]))
    del sets
    torch.cuda.empty_cache()
    # SimDR loss
    Bq, Lv = 2048, 512
    sd = []
    for r in range(R):
        xv, yv = synth.simdr_vectors(Bq, 21, Lv, seed=r, device=DEV)
        j, v = synth.hand_joints(Bq, 21, (256, 256), seed=r + 5, device=DEV)
        tx, ty = ops.render_simdr(j, v, (256, 256), 2, 2)
        sd.append((xv, yv, tx, ty, v[..., :1].contiguous(), j, v))
    vb = Bq * 21 * Lv * 4
    rows.append(row("KLDiscretLoss fwd (lhn_simdr_smoothl1): 4 vector reads", 4 * vb,
                    [lambda s=s: ops.simdr_smoothl1(s[0], s[1], s[2], s[3], s[4]) for s in sd]))
    rows.append(row("KLDiscretLoss bwd (lhn_simdr_smoothl1_backward): 4 reads + 2 writes", 6 * vb,
                    [lambda s=s: ops.simdr_smoothl1_backward(s[0], s[1], s[2], s[3], s[4]) for s in sd]))
    rows.append(row("render SimDR targets (lhn_render_simdr): 2 vector writes", 2 * vb,
                    [lambda s=s: ops.render_simdr(s[5], s[6], (256, 256), 2, 2) for s in sd]))
    del sd
    torch.cuda.empty_cache()
    # metrics
    N = 1 << 20
    g = torch.Generator(device=DEV).manual_seed(1)
    pm = []
    for r in range(R):
        pred = torch.rand(N, 16, 2, generator=g, device=DEV) * 256
        gt = pred + torch.randn(N, 16, 2, generator=g, device=DEV) * 6
        mask = torch.rand(N, 16, generator=g, device=DEV) < 0.9
        nor = (torch.rand(N, 1, generator=g, device=DEV) * 140 + 60).expand(N, 2).contiguous()
        pm.append((pred, gt, mask, nor))
    rows.append(row("lhn_pck_accumulate, 1 Mi samples x 16 joints (pred, gt f32, mask u8, normalize f32)",
                    N * 16 * (8 + 8 + 1) + N * 8, [lambda s=s: ops.pck_accumulate(s[0], s[1], s[2], [0.2], s[3]) for s in pm]))
    del pm
    torch.cuda.empty_cache()
    ev = []
    for r in range(R):
        hm, cen = synth.blob_heatmaps(1024, 21, 64, 64, seed=r, device=DEV)
        gh, _ = synth.blob_heatmaps(1024, 21, 64, 64, seed=r + 9, device=DEV, centers=cen)
        wh = torch.rand(1024, 2, generator=g, device=DEV) * 140 + 60
        ev.append((hm, gh, wh))
    rows.append(row("lhn_evaluate_pck (two argmax passes + per-image ratio), 1024 x 21 x 64 x 64", 2 * plane_bytes,
                    [lambda s=s: ops.evaluate_pck(s[0], s[1], s[2], None, (256, 256), 0.2) for s in ev]))
    for r in rows:
        print(f"{r['kernel']:92s} {r['us']:9.1f} us  {r['gbs']:8.1f} GB/s  {r['frac'] * 100:5.1f}% of {PEAK:.0f}"
              f"  ({r['timed']}; eager loop {r['us_eager']:.1f} us)")
    if "--json" in sys.argv:
        json.dump(rows, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)


if __name__ == "__main__":
    main()
