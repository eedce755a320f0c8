"""The reference's OWN functions on the host CPU, per BASELINE.md §3 config (TEST/BENCH INFRASTRUCTURE:
`bench.py --impl reference` and the `cpu_baseline` leg; nothing under litehandnet_b200/ imports this).

The unmodified reference sources are executed — from oracle/_ref/ (staged by oracle/build_ref.py; travels to
the GPU box) or from /root/reference — through oracle/ref_loader.py.  Per config, exactly the calls the CUDA
path replaces:

  1  top_down_eval.keypoints_from_heatmaps(post_process='default')                       (:375-463)
  2  TopDownGenerateTarget per sample (generateTarget.py:245-300) -> DistanceLoss(L2, balance=True)
     (heatmapLoss.py:242-265) -> flip_back + average (utils/transforms.py:78-92) ->
     keypoints_from_heatmaps(post_process='unbiased', kernel=11)
  3  top_down_eval.keypoints_from_simdr(k=2)                                              (:466-500)
  4  config-1 decode on 16 joints -> keypoint_pck_accuracy(0.2) / keypoint_auc(30) / keypoint_epe
  5  as 2 at 128x128 without the flip plane

The reference is single-process Python.  To give it "all the host threads it can use" the per-sample Python
loops (render, decode — the reference itself renders in DataLoader workers) run on contiguous shards in forked
worker processes, and the torch parts (DistanceLoss on the whole sample, so N_pos stays batch-global) run in the
parent with torch's intra-op threads on every core.  The metric functions of config 4 run once on the whole sample.
"""
import multiprocessing as mp
import os
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
STAGED = os.path.join(HERE, "_ref")

_G = {}
_REF = None


def reference_root():
    if os.path.exists(os.path.join(STAGED, "MANIFEST.sha256")):
        return STAGED
    if os.path.isdir("/root/reference/utils"):
        return "/root/reference"
    return None


def available():
    return reference_root() is not None


def ref():
    global _REF
    if _REF is None:
        from . import ref_loader
        ref_loader.REF_ROOT = reference_root()
        _REF = ref_loader.load()
    return _REF


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def _shared_f32(shape):
    """Anonymous shared memory visible to the forked workers (they write their shard of the rendered targets)."""
    n = int(np.prod(shape))
    raw = mp.RawArray("f", n)
    return np.frombuffer(raw, dtype=np.float32, count=n).reshape(shape)


# ---- shard functions (run in the workers; inputs in _G, shared copy-on-write) -----------------------------
def _shard_decode_default(bounds):
    a, b = bounds
    g = _G
    with np.errstate(all="ignore"):
        _, preds, maxvals = g["T"].keypoints_from_heatmaps(g["hm"][a:b].copy(), g["center"][a:b], g["scale"][a:b],
                                                          post_process="default")
    return a, b, np.concatenate([preds, maxvals], axis=2)


def _shard_render_decode(bounds):
    a, b = bounds
    g = _G
    hm = g["hm"]
    K, H, W = hm.shape[1:]
    gen, ann = g["gen"], g["ann"]
    for i in range(a, b):
        out = gen(dict(joints_3d=g["joints"][i], joints_3d_visible=g["vis"][i], ann_info=ann))
        g["target"][i] = out["target"]
        g["weight"][i] = out["target_weight"]
    with np.errstate(all="ignore"):
        if g["hf"] is not None:
            avg = (hm[a:b] + g["flip_back"](g["hf"][a:b].copy(), g["pairs"])) * 0.5
        else:
            avg = hm[a:b].copy()
        _, preds, maxvals = g["T"].keypoints_from_heatmaps(avg, g["center"][a:b], g["scale"][a:b],
                                                          post_process="unbiased", kernel=g["kernel"])
    return a, b, np.concatenate([preds, maxvals], axis=2)


def _shard_simdr(bounds):
    a, b = bounds
    g = _G
    return a, b, g["T"].keypoints_from_simdr(g["xv"][a:b], g["yv"][a:b], g["center"][a:b], g["scale"][a:b], g["k"])


class RefRunner:
    """run(n) executes the reference pipeline of one config on the first n samples; returns (result, seconds)."""

    def __init__(self, cfg_id, inputs, image_size=(256, 256), sigma=2, kernel=11, pairs=(), k=2, workers=None):
        import torch
        r = ref()
        self.cfg_id = cfg_id
        self.torch = torch
        T = r.top_down_eval
        _G.clear()
        _G.update(T=T, kernel=kernel, pairs=list(pairs), k=k)
        if cfg_id in (1, 4):
            hm, center, scale = inputs[:3]
            self.B, self.K = hm.shape[:2]
            _G.update(hm=hm, center=center, scale=scale)
            self.shard_fn = _shard_decode_default
            if cfg_id == 4:
                self.gt, self.mask, self.bbox_wh = inputs[3:6]
        elif cfg_id in (2, 5):
            hm, hf, joints, vis, center, scale = inputs
            self.B, self.K = hm.shape[:2]
            H, W = hm.shape[2:]
            gen = r.generateTarget.TopDownGenerateTarget(sigma=sigma, unbiased_encoding=True)
            ann = dict(num_joints=self.K, image_size=np.array(image_size), heatmap_size=np.array([W, H]),
                       joint_weights=None, use_different_joint_weights=False)
            _G.update(hm=hm, hf=hf, joints=joints, vis=vis, center=center, scale=scale, gen=gen, ann=ann,
                      flip_back=r.transforms.flip_back, target=_shared_f32(hm.shape),
                      weight=_shared_f32((self.B, self.K, 1)))
            self.crit = r.loss.DistanceLoss("L2", "mean", True)
            self.shard_fn = _shard_render_decode
        elif cfg_id == 3:
            xv, yv, center, scale = inputs
            self.B, self.K = xv.shape[:2]
            _G.update(xv=xv, yv=yv, center=center, scale=scale)
            self.shard_fn = _shard_simdr
        else:
            raise ValueError(cfg_id)
        self.workers = max(1, min(workers or host_cores(), self.B))
        torch.set_num_threads(host_cores())
        self.pool = mp.get_context("fork").Pool(self.workers) if self.workers > 1 else None   # after _G is filled

    def run(self, n=None):
        torch = self.torch
        n = self.B if n is None else min(int(n), self.B)
        parts = max(1, min(self.workers, n))
        edges = [round(i * n / parts) for i in range(parts + 1)]
        bounds = [(a, b) for a, b in zip(edges[:-1], edges[1:]) if b > a]
        t0 = time.perf_counter()
        results = self.pool.map(self.shard_fn, bounds, chunksize=1) if self.pool else [self.shard_fn(bd) for bd in bounds]
        preds = np.zeros((n, self.K, 3), np.float32)
        for a, b, p in results:
            preds[a:b] = p
        out = dict(preds=preds)
        if self.cfg_id in (2, 5):
            with torch.no_grad():
                out["loss"] = float(self.crit(torch.from_numpy(_G["hm"][:n]), torch.from_numpy(_G["target"][:n]),
                                              torch.from_numpy(_G["weight"][:n])).item())
        elif self.cfg_id == 4:
            T = _G["T"]
            p64 = preds[..., :2].astype(np.float64)             # _report_metric reads the predictions back from JSON
            t = np.max(self.bbox_wh[:n].astype(np.float64), axis=1)
            with np.errstate(all="ignore"):
                _, pck, _ = T.keypoint_pck_accuracy(p64, self.gt[:n], self.mask[:n], 0.2, np.stack([t, t], axis=1))
                out.update(PCK=float(pck), AUC=float(T.keypoint_auc(p64, self.gt[:n], self.mask[:n], 30)),
                           EPE=float(T.keypoint_epe(p64, self.gt[:n], self.mask[:n])))
        return out, time.perf_counter() - t0

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()
            self.pool = None
