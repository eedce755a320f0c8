#!/usr/bin/env python
"""Randomised parity sweep of the region-map / box kernels against the oracle (test infrastructure, GPU box):
box NMS on random overlapping boxes with tied scores, heatmap / vector NMS with NaN, +-inf and plateaus, the fused
region decode in its three modes, the window-restricted decode.   python profiles/probes/fuzz_region.py [n_seeds]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import assert_coords_close  # noqa: E402
from litehandnet_b200 import _lib as L, ops, synth  # noqa: E402
from oracle import np_oracle as O  # noqa: E402

F32 = np.float32
DEV = "cuda"


def cu(a):
    return torch.as_tensor(np.ascontiguousarray(a)).to(DEV)


def arrays(lists, max_num):
    boxes = np.zeros((len(lists), max_num, 5), F32)
    counts = np.zeros(len(lists), np.int32)
    for i, l in enumerate(lists):
        if l is not None:
            counts[i] = len(l)
            boxes[i, :len(l)] = np.asarray(l, F32)
    return boxes, counts


def fuzz_box_nms(rng):
    B, N = 64, int(rng.integers(1, 33))
    c = np.zeros((B, N, 5), F32)
    c[..., 0:2] = rng.uniform(0, 256, (B, N, 2))
    c[..., 2:4] = rng.uniform(0, 120, (B, N, 2))
    c[..., 4] = np.round(rng.uniform(0, 1, (B, N)), int(rng.integers(1, 4)))       # coarse scores: many ties
    c[rng.random((B, N)) < 0.05, 2] = 1.0                                            # below min_wh
    if N > 2:
        c[:, 1, :4] = c[:, 0, :4]                                                    # exact duplicates: IoU == 1
        c[:, 2, :4] = c[:, 0, :4] + F32(0.5)
    det, iou, mx = float(rng.uniform(0, 0.6)), float(rng.choice([0.1, 0.3, 0.5, 0.6, 0.9])), int(rng.integers(1, N + 1))
    want_b, want_c = arrays(O.box_nms(c, det, iou, mx), mx)
    boxes, counts = ops.box_nms(cu(c), det, iou, mx)
    assert np.array_equal(counts.cpu().numpy(), want_c), "box_nms counts"
    assert np.array_equal(boxes.cpu().numpy(), want_b), "box_nms boxes"


def fuzz_nms_maps(rng):
    B, C = 2, 3
    H, W = int(rng.choice([16, 32, 56, 64])), int(rng.choice([16, 32, 56, 64]))
    k = int(rng.choice([3, 5, 11]))
    hm = rng.random((B, C, H, W)).astype(F32)
    hm[rng.random(hm.shape) < 0.002] = np.nan
    hm[rng.random(hm.shape) < 0.002] = np.inf
    hm[rng.random(hm.shape) < 0.002] = -np.inf
    hm[0, 0, H // 2, 2:6] = 3.0
    hm = np.round(hm, 2) if rng.random() < 0.5 else hm                               # plateaus
    with np.errstate(all="ignore"):
        want = O.heatmap_nms(hm, k, (k - 1) // 2)
    assert np.array_equal(ops.heatmap_nms(cu(hm), k).cpu().numpy(), want, equal_nan=True), "heatmap_nms"
    v = rng.random((3, 5, int(rng.choice([36, 100, 448, 512])))).astype(F32)
    v[rng.random(v.shape) < 0.003] = np.nan
    v = np.round(v, 1) if rng.random() < 0.5 else v
    with np.errstate(all="ignore"):
        want = O.vector_nms(v)
    assert np.array_equal(ops.vector_nms(cu(v)).cpu().numpy(), want, equal_nan=True), "vector_nms"


def region_maps(rng, B, H, W):
    c = (rng.random((B, 1, H, W)) * 0.3).astype(F32)
    ys, xs = np.meshgrid(np.arange(H, dtype=F32), np.arange(W, dtype=F32), indexing="ij")
    for i in range(B):
        for _ in range(int(rng.integers(1, 5))):
            cx, cy, a = rng.random() * W, rng.random() * H, 0.4 + rng.random()
            c[i, 0] += (a * np.exp(-((xs - cx) ** 2 + (ys - cy) ** 2) / 8)).astype(F32)
    return c, rng.random((B, 2, H, W)).astype(F32)


def fuzz_region(rng, mode):
    B = 3
    H = W = int(rng.choice([32, 56, 64]))
    c, s = region_maps(rng, B, H, W)
    if mode != "sh":
        s = (s * F32(30)).astype(F32)
    N = int(rng.integers(1, 21))
    mx = int(rng.integers(1, N + 1))
    isz = (4.0 * W, 4.0 * H)
    with np.errstate(all="ignore"):
        nms = O.heatmap_nms(c) if mode != "cs" else c
        oc = O.candidate_bbox(nms, s, mode, isz, (4, 4), num_candidates=N, thr=0.25)
        ob, on = arrays(O.box_nms(oc, 0.25, 0.4, mx), mx)
    r = ops.region_bbox_decode(cu(c), cu(s), dict(sh=L.REGION_SH, rp=L.REGION_RP, cs=L.REGION_CS)[mode],
                               nms_kernel=0 if mode == "cs" else 11, num_candidates=N, max_num_bbox=mx,
                               refine=L.REFINE_DARK_LEGACY if mode == "rp" else L.REFINE_NONE, image_size=isz,
                               cand_thr=0.25, det_thr=0.25, iou_thr=0.4)
    cand = r["candidates"].cpu().numpy()
    assert np.array_equal(cand[..., 4], oc[..., 4]), f"{mode} confidences"
    if mode == "sh":
        assert np.array_equal(cand, oc), "sh candidates"
    else:
        assert_coords_close(cand, oc, what=f"{mode} candidates")
    assert np.array_equal(r["counts"].cpu().numpy(), on), f"{mode} counts"
    assert_coords_close(r["boxes"].cpu().numpy(), ob, what=f"{mode} boxes")


def fuzz_roi(rng, seed):
    B, K, H, W = 4, 3, 64, 64
    roi = np.zeros((B, 4), np.int32)
    centers = np.zeros((B, K, 2), F32)
    for b in range(B):
        x0, y0 = rng.integers(0, W - 12), rng.integers(0, H - 12)
        roi[b] = (x0, y0, rng.integers(x0 + 10, W + 1), rng.integers(y0 + 10, H + 1))
        centers[b, :, 0] = (roi[b, 0] + roi[b, 2]) / 2 + (rng.random(K) - 0.5) * 0.6 * (roi[b, 2] - roi[b, 0])
        centers[b, :, 1] = (roi[b, 1] + roi[b, 3]) / 2 + (rng.random(K) - 0.5) * 0.6 * (roi[b, 3] - roi[b, 1])
    hm = synth.blob_heatmaps(B, K, H, W, seed=seed, centers=torch.from_numpy(centers))[0].numpy()
    for refine, dark in ((L.REFINE_OFFSET_HALF, False), (L.REFINE_DARK_LEGACY, True)):
        out = ops.decode_heatmap_roi(cu(hm), cu(roi), refine, scale_xy=(4.0, 4.0)).cpu().numpy()
        for b in range(B):
            x0, y0, x1, y1 = [int(v) for v in roi[b]]
            with np.errstate(all="ignore"):
                k = O.get_pred_kpt(hm[b:b + 1, :, y0:y1, x0:x1], dark=dark)
            k[..., :2] += np.asarray([x0, y0], F32)
            k[..., :2] *= F32(4)
            assert_coords_close(out[b], k[0], what=f"roi refine={refine}")


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    t0 = time.time()
    done = dict(box=0, maps=0, sh=0, rp=0, cs=0, roi=0)
    for seed in range(n):
        rng = np.random.default_rng(1000 + seed)
        for _ in range(4):
            fuzz_box_nms(rng); done["box"] += 1
        fuzz_nms_maps(rng); done["maps"] += 1
        for mode in ("sh", "cs"):
            fuzz_region(rng, mode); done[mode] += 1
        if seed % 4 == 0:
            fuzz_region(rng, "rp"); done["rp"] += 1
            fuzz_roi(rng, seed); done["roi"] += 1
    print(f"fuzz_region: all equal — {done} cases in {time.time() - t0:.1f} s")


if __name__ == "__main__":
    main()
