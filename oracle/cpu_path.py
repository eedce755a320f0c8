"""The reference's CPU path for the headline workload, driven through the numpy oracle on all host
cores (TEST/BENCH INFRASTRUCTURE: the `cpu_baseline` leg and `bench.py --impl reference`).

Pipeline per BASELINE.md §3 config 2 — exactly the calls the fused kernel replaces:
  TopDownGenerateTarget per sample -> DistanceLoss(L2, balance=True) -> flip_back + average ->
  keypoints_from_heatmaps('unbiased', kernel=11).
The reference itself is single-process Python; to give the CPU "all the host threads it can use" the
batch is cut into contiguous shards that run in forked worker processes (inputs shared copy-on-write),
and the balanced loss is finalised from the summed shard sums (equal to the monolithic value up to
f64 summation order).
"""
import multiprocessing as mp
import os
import time

import numpy as np

from . import np_oracle as O

_G = {}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def _shard(bounds):
    a, b = bounds
    g = _G
    hm = g["hm"][a:b]
    with np.errstate(all="ignore"):
        tg, tw = O.render_targets(g["joints"][a:b], g["vis"][a:b], g["image_size"], (hm.shape[3], hm.shape[2]),
                                  g["sigma"], True)
        sums = O.distance_loss_l2_sums(hm, tg, tw)
        avg = hm if g["hf"] is None else O.flip_average(hm, g["hf"][a:b], g["pairs"])
        hm_preds, preds, maxvals = O.keypoints_from_heatmaps(avg, g["center"][a:b], g["scale"][a:b], "unbiased", g["kernel"])
        idx = avg.reshape(avg.shape[0], avg.shape[1], -1).argmax(-1).astype(np.int32)   # _get_max_preds' np.argmax
    return a, b, np.concatenate([preds, maxvals], axis=2), sums, idx


class FusedCpuRunner:
    """Holds the host inputs and a persistent pool of forked workers; run(n) processes the first n
    samples once and returns (preds [n,K,3], loss, seconds)."""

    def __init__(self, hm, hf, joints, vis, center, scale, image_size=(256, 256), sigma=2, kernel=11,
                 pairs=(), workers=None):
        self.B, self.K = hm.shape[0], hm.shape[1]
        self.workers = max(1, min(workers or host_cores(), self.B))
        _G.update(hm=hm, hf=hf, joints=joints, vis=vis, center=center, scale=scale, image_size=image_size,
                  sigma=sigma, kernel=kernel, pairs=pairs)
        self.pool = None
        if self.workers > 1:
            self.pool = mp.get_context("fork").Pool(self.workers)     # forked AFTER _G is filled

    def run(self, n=None):
        n = self.B if n is None else min(n, self.B)
        parts = max(1, min(self.workers, n))
        edges = [round(i * n / parts) for i in range(parts + 1)]
        bounds = [(a, b) for a, b in zip(edges[:-1], edges[1:]) if b > a]
        t0 = time.perf_counter()
        results = self.pool.map(_shard, bounds, chunksize=1) if self.pool else [_shard(bd) for bd in bounds]
        preds = np.zeros((n, self.K, 3), np.float32)
        idx = np.zeros((n, self.K), np.int32)
        sums = np.zeros(4, np.float64)
        for a, b, p, s, ix in results:
            preds[a:b] = p
            idx[a:b] = ix
            sums += s
        self.last_idx, self.last_sums = idx, sums       # argmax indices / f64 loss sums of the last run (parity checks)
        loss = O.distance_loss_from_sums(sums, True)
        return preds, loss, time.perf_counter() - t0

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()
            self.pool = None


def run_fused(hm, hf, joints, vis, center, scale, image_size=(256, 256), sigma=2, kernel=11, pairs=(),
              workers=None):
    """One-shot convenience: returns (preds, loss, seconds, workers_used)."""
    r = FusedCpuRunner(hm, hf, joints, vis, center, scale, image_size, sigma, kernel, pairs, workers)
    try:
        preds, loss, dt = r.run()
    finally:
        r.close()
    return preds, loss, dt, r.workers
