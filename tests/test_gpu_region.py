"""GPU: the bbox branch of the legacy parsers — centre-map NMS, top-k candidates, size lookup,
legacy-DARK candidate refinement, box NMS, bbox-restricted keypoint decode, the +-0.25 rule at given points —
against the golden vectors of the executed reference (tests/golden/region_bbox.npz) and against the oracle on
seeded inputs.  Integer/box-selection results bit-exact; float coordinates within 1e-5 relative."""
import numpy as np
import pytest
import torch

from conftest import assert_coords_close, canon_candidates, load_golden
from oracle import np_oracle as O
from litehandnet_b200 import _lib as L
from litehandnet_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda"
F32 = np.float32


def cu(a):
    return torch.as_tensor(np.ascontiguousarray(a)).to(DEV)


def _arrays(lists, max_num):
    B = len(lists)
    boxes = np.zeros((B, max_num, 5), F32)
    counts = np.zeros(B, np.int32)
    for i, l in enumerate(lists):
        if l is not None:
            counts[i] = len(l)
            boxes[i, :len(l)] = np.asarray(l, F32)
    return boxes, counts


def region_maps(B, seed, H=64, W=64, npk=3):
    rng = np.random.default_rng(seed)
    c = (rng.random((B, 1, H, W)) * 0.3).astype(F32)
    ys, xs = np.meshgrid(np.arange(H, dtype=F32), np.arange(W, dtype=F32), indexing="ij")
    for i in range(B):
        for _ in range(npk):
            cx, cy, a = rng.random() * W, rng.random() * H, 0.4 + rng.random()
            c[i, 0] += (a * np.exp(-((xs - cx) ** 2 + (ys - cy) ** 2) / 8)).astype(F32)
    s = rng.random((B, 2, H, W)).astype(F32)
    return c, s


def cfg(dark):
    return dict(image_size=[256, 256], hm_size=[64, 64], model="litehandnet", simdr_split_ratio=2, bbox_alpha=1.0,
                with_region_map=False, cycle_detection_reduction=1, DARK=dark)


# ---- heatmap_nms / vector_nms --------------------------------------------------------------------------------
def test_heatmap_nms_golden_and_inplace():
    from litehandnet_b200.decode import HeatmapParser_SH, ResultParser
    g = load_golden("region_bbox.npz")
    c = cu(g["center"])
    ret = HeatmapParser_SH().heatmap_nms(c)
    assert ret is c, "the reference returns its (mutated) argument"
    assert np.array_equal(c.cpu().numpy(), g["ref_nms"])
    # out of place through the ops layer, on a channel-sliced view and in bf16
    region = torch.cat([cu(g["center"]), cu(g["size"])], 1)
    out = ops.heatmap_nms(region[:, 0:1], 11)
    assert np.array_equal(out.cpu().numpy(), g["ref_nms"]) and np.array_equal(region[:, 0:1].cpu().numpy(), g["center"])
    hb = cu(g["center"]).to(torch.bfloat16)
    ob = ops.heatmap_nms(hb, 11)
    assert np.array_equal(ob.float().cpu().numpy(), O.heatmap_nms(hb.float().cpu().numpy()))
    # CPU input: result on the CPU, argument untouched
    cpu_in = torch.from_numpy(g["center"].copy())
    r = ResultParser(cfg(True)).heatmap_nms(cpu_in)
    assert not r.is_cuda and np.array_equal(r.numpy(), g["ref_nms"])


@pytest.mark.parametrize("shape,k", [((3, 21, 64, 64), 11), ((2, 4, 56, 56), 5), ((2, 3, 17, 33), 3), ((1, 2, 128, 128), 11)])
def test_heatmap_nms_shapes_nan_ties(shape, k):
    rng = np.random.default_rng(5)
    hm = rng.random(shape).astype(F32)
    hm[0, 0, 3, 4:7] = 2.0                       # plateau
    hm[0, 1, 5, 5] = np.nan                      # NaN: stays NaN, wipes its window
    hm[-1, -1] = 0                               # all-equal plane: everything survives (0 * 1)
    hm[-1, 0, 2, 2] = -np.inf
    with np.errstate(all="ignore"):
        ref = O.heatmap_nms(hm, k, (k - 1) // 2)
    out = ops.heatmap_nms(cu(hm), k).cpu().numpy()
    assert np.array_equal(out, ref, equal_nan=True)


def test_vector_nms_kernel():
    from litehandnet_b200.decode import ResultParser
    rng = np.random.default_rng(7)
    v = rng.random((5, 21, 512)).astype(F32)
    v[0, 0, 10:13] = 3.0
    v[1, 2, 100] = np.nan
    v[2, 3, 0] = 9.0; v[2, 3, -1] = 9.0
    with np.errstate(all="ignore"):
        ref = O.vector_nms(v)
    t = cu(v)
    out = ResultParser(cfg(False)).vector_nms(t)
    got = out.cpu().numpy()
    # NaN * 0 = NaN in torch; the oracle's (vmax == v) mask gives NaN * 0 too
    assert np.array_equal(got, ref, equal_nan=True)
    assert np.array_equal(t.cpu().numpy(), v, equal_nan=True), "returns a new tensor"
    assert np.array_equal(ops.vector_nms(cu(v[0, 0, :7])).cpu().numpy(), O.vector_nms(v[0, 0, :7]))


# ---- HeatmapParser_SH ------------------------------------------------------------------------------------------
def test_sh_parse_with_region_maps_golden():
    """utils/SPheatmapParser.py:169-206 with centre/size maps, as the reference's __main__ block calls it."""
    from litehandnet_b200.decode import HeatmapParser_SH
    g = load_golden("region_bbox.npz")
    c = cu(g["center"])
    kpt, boxes = HeatmapParser_SH().parse(cu(g["kpt_hm"]), c, cu(g["size"]), (256, 256))
    assert not kpt.is_cuda and np.array_equal(kpt.numpy(), g["ref_parse_kpt"])
    assert np.array_equal(c.cpu().numpy(), g["ref_parse_center_after"]), "centre maps are masked in place"
    bx, ct = _arrays(boxes, 1)
    assert np.array_equal(ct, g["ref_parse_counts"])
    tie_free = [0, 2, 3, 4, 5]                   # image 1 holds a three-way tie: torch.topk's order is unspecified
    assert np.array_equal(bx[tie_free], g["ref_parse_boxes"][tie_free])
    assert boxes[2] is None
    # image 1: the lowest index of the tie (documented rule), everything else as the reference
    assert bx[1, 0, 0] == 40.0 and bx[1, 0, 1] == 40.0 and bx[1, 0, 4] == g["ref_parse_boxes"][1, 0, 4] == 2.0


def test_sh_known_answer():
    """The reference's own fixture (utils/SPheatmapParser.py:221-233): peak (3,3), size maps 1 on [0,7)^2."""
    from litehandnet_b200.decode import HeatmapParser_SH
    g = load_golden("region_bbox.npz")
    k_hm = torch.zeros((2, 4, 64, 64), device=DEV); k_hm[..., 3, 3] = 1; k_hm[..., 3, 2] = 0.5; k_hm[..., 2, 3] = 0.5
    c_hm = torch.zeros(2, 1, 64, 64, device=DEV); c_hm[..., 3, 3] = 1
    s_hm = torch.zeros(2, 2, 64, 64, device=DEV); s_hm[..., 0:7, 0:7] = 1.
    k, b = HeatmapParser_SH().parse(k_hm, c_hm, s_hm, (256, 256))
    assert np.array_equal(k[0, 0].numpy(), np.array([11.0, 11.0, 1.0], F32))
    assert np.array_equal(_arrays(b, 1)[0], g["ref_main_boxes"])
    assert b[0] == [[12.0, 12.0, 253.44000244140625, 253.44000244140625, 1.0]]
    k2, b2 = HeatmapParser_SH().parse(k_hm, None, None, (256, 256))
    assert b2 is None and torch.equal(k, k2)


def test_sh_candidates_and_nms_golden():
    from litehandnet_b200.decode import HeatmapParser_SH
    g = load_golden("region_bbox.npz")
    P = HeatmapParser_SH()
    cand = P.candidate_bbox(cu(g["ref_nms"]), cu(g["size"]), (256, 256))
    assert not cand.is_cuda and cand.shape == (6, 10, 5)
    (a, da), (b, db) = canon_candidates(cand.numpy()), canon_candidates(g["ref_sh_cand"])
    assert np.array_equal(da, db) and np.array_equal(a[da], b[db])
    # the kernel's own order: lowest index first among equal values == the oracle's rule, bit for bit
    assert np.array_equal(cand.numpy(), O.candidate_bbox(g["ref_nms"], g["size"], "sh", (256, 256)))
    # box NMS on the reference's candidate order, several thresholds and box budgets
    bx, ct = _arrays(P.non_max_suppression(cu(g["ref_sh_cand"])), 1)
    assert np.array_equal(ct, g["ref_sh_counts"]) and np.array_equal(bx, g["ref_sh_boxes"])
    P.max_num_bbox = 10
    for thr in (0.6, 0.3, 0.1):
        P.iou_threshold = thr
        bx, ct = _arrays(P.non_max_suppression(torch.from_numpy(g["ref_sh_cand"])), 10)
        assert np.array_equal(ct, g[f"ref_sh_counts10_iou{int(thr * 10)}"])
        assert np.array_equal(bx, g[f"ref_sh_boxes10_iou{int(thr * 10)}"])


# ---- ResultParser (DARK) ---------------------------------------------------------------------------------------
def test_rp_get_pred_bbox_golden():
    from litehandnet_b200.decode import ResultParser
    g = load_golden("region_bbox.npz")
    s40 = (g["size"] * F32(40)).astype(F32)
    RP = ResultParser(cfg(True))
    RP.num_candidates = 1
    cand1 = RP.candidate_bbox(cu(g["ref_nms"]), cu(s40))
    tie_free = [0, 2, 3, 4, 5]
    assert_coords_close(cand1.numpy()[tie_free], g["ref_rp_cand1"][tie_free], what="rp cand1")
    RP.num_candidates = 10
    cand = RP.candidate_bbox(cu(g["ref_nms"]), cu(s40)).numpy()
    ref10 = np.concatenate([g["ref_rp_xy10"], np.zeros_like(g["ref_rp_xy10"]), g["ref_rp_topk_val"][..., None]], axis=2)
    mine = cand.copy(); mine[..., 2:4] = 0
    (a, da), (b, db) = canon_candidates(mine), canon_candidates(ref10)
    assert np.array_equal(da, db) and np.array_equal(a[da][:, 4], b[db][:, 4])
    assert_coords_close(a[da][:, :2], b[db][:, :2], what="rp xy10")
    with np.errstate(all="ignore"):
        oc = O.candidate_bbox(g["ref_nms"], s40, "rp", num_candidates=10)
    assert_coords_close(cand, oc, what="rp cand vs oracle")
    assert (np.abs(cand[..., :2] / 4 - np.round(cand[..., :2] / 4)) > 1e-3).sum() > 10, "DARK offsets must be exercised"
    # the whole call: region map in, list of boxes out, channel 0 masked in place
    region = torch.cat([cu(g["center"]), cu(s40)], 1)
    boxes = RP.get_pred_bbox(region)
    assert np.array_equal(region[:, 0:1].cpu().numpy(), g["ref_nms"])
    assert np.array_equal(region[:, 1:3].cpu().numpy(), s40)
    bx, ct = _arrays(boxes, 1)
    with np.errstate(all="ignore"):
        obx, oct_ = _arrays(O.box_nms(oc, 0.1, 0.6, 1), 1)
    assert np.array_equal(ct, oct_) and np.array_equal(ct, g["ref_rp_counts1"])
    assert_coords_close(bx, obx, what="rp boxes")
    assert_coords_close(bx[tie_free], g["ref_rp_boxes1"][tie_free], what="rp boxes vs reference")
    # DARK off: the reference raises TypeError inside candidate_bbox (torch.from_numpy of a tensor)
    with pytest.raises(TypeError):
        ResultParser(cfg(False)).get_pred_bbox(region)


def test_rp_first_result_and_group_keypoints_golden():
    from litehandnet_b200.decode import ResultParser
    g = load_golden("region_bbox.npz")
    hm = cu(g["kpt_hm"])
    for dark in (0, 1):
        RP = ResultParser(cfg(bool(dark)))
        for j, bb in enumerate(g["first_result_boxes"]):
            r = RP._get_first_result([float(v) for v in bb], hm, j % 6)
            assert r.is_cuda and r.shape == (1, 5, 3)
            assert_coords_close(r[0].cpu().numpy(), g[f"ref_first_result_dark{dark}"][j], what=f"first_result dark={dark} box {j}")
        # batched: one bbox slot for every image in one launch
        bbox_list = [[[float(v) for v in g["first_result_boxes"][i]]] for i in range(6)]
        bbox_list[3] = None
        out = RP.get_group_keypoints(None, None, bbox_list, hm)
        assert out.shape == (6, 1, 5, 3) and not out.is_cuda
        for i in range(6):
            if i == 3:
                assert float(out[i].abs().sum()) == 0.0
            else:
                assert_coords_close(out[i, 0].numpy(), g[f"ref_first_result_dark{dark}"][i], what=f"group dark={dark} image {i}")
    assert torch.equal(hm, cu(g["kpt_hm"])), "heatmaps must not be modified"
    RPc = ResultParser(dict(cfg(False), with_region_map=True))
    with pytest.raises(NotImplementedError):
        RPc.get_group_keypoints(None, None, [None] * 6, hm)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_roi_decode_random_windows_vs_oracle(seed):
    """Random windows; the blobs of an image lie inside its window (a window holding only noise makes the Taylor
    step of DARK ill-conditioned — it amplifies 1-ulp differences of logf — so there is no parity to test there)."""
    from litehandnet_b200 import synth
    rng = np.random.default_rng(seed)
    B, K, H, W = 7, 4, 64, 64
    roi = np.zeros((B, 4), np.int32)
    for b in range(B):
        x0, y0 = rng.integers(0, W - 12), rng.integers(0, H - 12)
        roi[b] = (x0, y0, rng.integers(x0 + 10, W + 1), rng.integers(y0 + 10, H + 1))
    roi[1] = (0, 0, W, H)
    roi[2] = (10, 10, 10, 30)                                   # empty: falls back to the whole plane
    roi[3] = (60, 5, 64, 9)                                     # 4x4 window: the DARK guard never passes
    centers = np.zeros((B, K, 2), F32)
    for b in range(B):
        x0, y0, x1, y1 = [int(v) for v in roi[b]]
        if x1 <= x0 or y1 <= y0:
            x0, y0, x1, y1 = 0, 0, W, H
        centers[b, :, 0] = (x0 + x1) / 2 + (rng.random(K) - 0.5) * 0.6 * (x1 - x0)
        centers[b, :, 1] = (y0 + y1) / 2 + (rng.random(K) - 0.5) * 0.6 * (y1 - y0)
    hm = synth.blob_heatmaps(B, K, H, W, seed=seed, centers=torch.from_numpy(centers))[0].numpy()
    hm[0, 0, int(centers[0, 0, 1]), int(centers[0, 0, 0]):int(centers[0, 0, 0]) + 2] = 5.0     # a tie inside a window
    for refine, dark in ((L.REFINE_OFFSET_HALF, False), (L.REFINE_DARK_LEGACY, True), (L.REFINE_NONE, None)):
        out, idx = ops.decode_heatmap_roi(cu(hm), cu(roi), refine, scale_xy=(4.0, 4.0), want_idx=True)
        out = out.cpu().numpy(); idx = idx.cpu().numpy()
        for b in range(B):
            x0, y0, x1, y1 = [int(v) for v in roi[b]]
            if x1 <= x0 or y1 <= y0:
                x0, y0, x1, y1 = 0, 0, W, H
            part = hm[b:b + 1, :, y0:y1, x0:x1]
            with np.errstate(all="ignore"):
                if dark is None:
                    p, mv, _ = O.max_preds(part, "none")
                    k = np.concatenate([p, mv], 2)
                else:
                    k = O.get_pred_kpt(part, dark=dark)
            k[..., :2] += np.asarray([x0, y0], F32)
            k[..., :2] *= F32(4)
            assert np.array_equal(idx[b], np.argmax(part.reshape(K, -1), 1)), (refine, b)
            assert_coords_close(out[b], k[0], what=f"roi refine={refine} image {b}")


# ---- evaluation.py ---------------------------------------------------------------------------------------------
def test_cs_from_region_map_golden():
    from litehandnet_b200.decode import cs_from_region_map, non_max_suppression
    g = load_golden("region_bbox.npz")
    s40 = (g["size"] * F32(40)).astype(F32)
    region = torch.cat([cu(g["center"]), cu(s40)], 1)
    before = region.clone()
    for K, thr in ((20, 0.1), (5, 0.5)):
        cc = cs_from_region_map(region, 256, K, thr)
        assert not cc.is_cuda and cc.shape == (6, K, 5)
        (a, da), (b, db) = canon_candidates(cc.numpy()), canon_candidates(g[f"ref_cs_cand_k{K}"])
        assert np.array_equal(da, db)
        assert_coords_close(a[da], b[db], what="cs cand")
        bx, ct = _arrays(non_max_suppression(cu(g[f"ref_cs_cand_k{K}"]), 0.6, 0.1, 3), 3)
        assert np.array_equal(ct, g[f"ref_cs_counts_k{K}"]) and np.array_equal(bx, g[f"ref_cs_boxes_k{K}"])
    assert torch.equal(region, before), "cs_from_region_map does not touch the region map"
    # defaults of the reference signature (max_num = 100 > number of candidates)
    assert len(non_max_suppression(cu(g["ref_cs_cand_k20"]))) == 6


@pytest.mark.parametrize("mode", ["sh", "rp", "cs"])
@pytest.mark.parametrize("shape", [(5, 64, 64), (3, 56, 56), (2, 32, 32)])
def test_region_decode_vs_oracle_seeded(mode, shape):
    """The fused call against the oracle pipeline on seeded maps: candidates bit-exact (sh) / 1e-5 (rp, cs),
    kept boxes and counts identical."""
    B, H, W = shape
    c, s = region_maps(B, seed=100 + H, H=H, W=W)
    s = (s * F32(30)).astype(F32) if mode != "sh" else s
    isz = (4.0 * W, 4.0 * H)
    with np.errstate(all="ignore"):
        nms = O.heatmap_nms(c) if mode != "cs" else c
        oc = O.candidate_bbox(nms, s, mode, isz, (4, 4), num_candidates=10, thr=0.2)
        ob, on = _arrays(O.box_nms(oc, 0.2, 0.5, 4), 4)
    r = ops.region_bbox_decode(cu(c), cu(s), dict(sh=L.REGION_SH, rp=L.REGION_RP, cs=L.REGION_CS)[mode],
                               nms_kernel=0 if mode == "cs" else 11, num_candidates=10, max_num_bbox=4,
                               refine=L.REFINE_DARK_LEGACY if mode == "rp" else L.REFINE_NONE, image_size=isz,
                               cand_thr=0.2, det_thr=0.2, iou_thr=0.5, want_nms=True)
    cand = r["candidates"].cpu().numpy()
    if mode == "sh":
        assert np.array_equal(cand, oc)
    else:
        assert np.array_equal(cand[..., 4], oc[..., 4])
        assert_coords_close(cand, oc, what=f"{mode} candidates")
    assert np.array_equal(r["nms"].cpu().numpy(), nms)
    assert np.array_equal(r["counts"].cpu().numpy(), on)
    assert_coords_close(r["boxes"].cpu().numpy(), ob, what=f"{mode} boxes")


def test_region_decode_is_deterministic_and_rejects_bad_calls():
    c, s = region_maps(8, seed=3)
    a = ops.region_bbox_decode(cu(c), cu(s), L.REGION_SH)
    b = ops.region_bbox_decode(cu(c), cu(s), L.REGION_SH)
    assert torch.equal(a["candidates"], b["candidates"]) and torch.equal(a["boxes"], b["boxes"])
    with pytest.raises(L.LhnError):
        ops.region_bbox_decode(cu(c), cu(s), L.REGION_SH, num_candidates=33)
    with pytest.raises(L.LhnError):
        ops.region_bbox_decode(cu(c), cu(s), L.REGION_SH, nms_kernel=4)
    with pytest.raises(L.LhnError):
        ops.region_bbox_decode(torch.from_numpy(c), cu(s), L.REGION_SH)       # no CPU fallback


# ---- HeatmapParser.adjust_keypoints at given points ------------------------------------------------------------
def test_heatmap_parser_adjust_keypoints_points():
    """utils/HeatmapParser.py:197-223 — grouped candidates (not plane argmaxima), list-of-lists in and out."""
    from litehandnet_b200.decode import HeatmapParser
    rng = np.random.default_rng(11)
    B, K, H, W = 3, 6, 64, 64
    hm = rng.random((B, K + 1, H, W)).astype(F32)
    kps = [[[[float(rng.integers(0, W)), float(rng.integers(0, H)), 0.5] for _ in range(K)] for _ in range(2)] for _ in range(B)]
    kps[1][1] = []                                                # a bbox without keypoints
    kps[0][0][0][:2] = [0.0, 63.0]; kps[0][0][1][:2] = [63.0, 0.0]  # corners: clamped neighbours
    for off in (0, 1):
        import copy
        mine = HeatmapParser(channel_offset=off).adjust_keypoints(copy.deepcopy(kps), cu(hm))
        for b in range(B):
            for g_ in range(2):
                for j, joint in enumerate(kps[b][g_]):
                    x, y = joint[:2]
                    xx, yy = int(x), int(y)
                    t = hm[b, j + off]
                    ex = x + (0.25 if t[yy, min(xx + 1, W - 1)] > t[yy, max(xx - 1, 0)] else -0.25)
                    ey = y + (0.25 if t[min(yy + 1, H - 1), xx] > t[max(yy - 1, 0), xx] else -0.25)
                    assert mine[b][g_][j][0] == ex and mine[b][g_][j][1] == ey and mine[b][g_][j][2] == 0.5
        assert mine[1][1] == []


# ---- SRHandNet multi-scale targets with the region map (R3b) --------------------------------------------------
def test_srhandnet_generate_target_golden():
    """datasets/data_pipeline/generateTarget.py:303-426 — dict in, lists of NumPy arrays out, as the pipeline calls it."""
    from litehandnet_b200.render import SRHandNetGenerateTarget, render_srhandnet_targets
    g = load_golden("render_srhandnet.npz")
    hs = [list(x) for x in g["heatmap_sizes"]]
    N, K = g["joints_3d"].shape[:2]
    for pred_bbox in (1, 0):
        for unb in (0, 1):
            T = SRHandNetGenerateTarget(pred_bbox=bool(pred_bbox), sigma=[2, 2, 2, 2], unbiased_encoding=bool(unb))
            for n in range(N):
                res = dict(joints_3d=g["joints_3d"][n].copy(), joints_3d_visible=g["joints_3d_visible"][n].copy(),
                           bbox=g["bbox"][n].copy(),
                           ann_info=dict(image_size=np.array([256, 256]), heatmap_size=np.array(hs), num_joints=K,
                                         use_different_joint_weights=False))
                out = T(res)
                assert isinstance(out["target"], list) and len(out["target"]) == 4
                for i in range(4):
                    rt = g[f"ref_t_bbox{pred_bbox}_unb{unb}_s{i}"][n]
                    t = out["target"][i]
                    assert isinstance(t, np.ndarray) and t.shape == rt.shape and t.dtype == np.float32
                    np.testing.assert_allclose(t, rt, rtol=1e-5, atol=2e-7)
                    assert np.array_equal(out["target_weight"][i], g[f"ref_w_bbox{pred_bbox}_unb{unb}_s{i}"][n])
                    if pred_bbox:       # the width / height planes are exact (one f32 constant inside the window)
                        assert np.array_equal(t[K + 1:], rt[K + 1:])
            # the batched entry: every sample in one go
            t, w = render_srhandnet_targets(cu(g["joints_3d"]), cu(g["joints_3d_visible"]), g["bbox"], (256, 256), hs,
                                            [2, 2, 2, 2], bool(pred_bbox), bool(unb))
            for i in range(4):
                np.testing.assert_allclose(t[i].cpu().numpy(), g[f"ref_t_bbox{pred_bbox}_unb{unb}_s{i}"], rtol=1e-5, atol=2e-7)
                assert np.array_equal(w[i].cpu().numpy(), g[f"ref_w_bbox{pred_bbox}_unb{unb}_s{i}"])


def test_adjust_keypoints_by_offset_at_arbitrary_positions():
    """utils/heatmap_post_processing.py:6-33 / SPheatmapParser.py:140-167 refine around int(keypoints) wherever the
    caller puts them; positions that are not the plane argmax go through lhn_refine_points."""
    from litehandnet_b200.decode import HeatmapParser_SH, adjust_keypoints_by_offset
    rng = np.random.default_rng(21)
    B, K, H, W = 4, 6, 64, 64
    hm = rng.random((B, K, H, W)).astype(F32)
    kp = np.zeros((B, K, 3), F32)
    kp[..., 0] = rng.integers(0, W, (B, K)); kp[..., 1] = rng.integers(0, H, (B, K)); kp[..., 2] = rng.random((B, K))
    kp[0, 0, :2] = (0, 0); kp[0, 1, :2] = (W - 1, H - 1)
    got = adjust_keypoints_by_offset(cu(kp), cu(hm))
    assert np.array_equal(got.cpu().numpy(), O.refine_offset_clamped(kp, hm, plus_half=True))
    got2 = HeatmapParser_SH.adjust_keypoints(cu(kp), cu(hm))
    assert np.array_equal(got2.cpu().numpy(), O.refine_offset_clamped(kp, hm, plus_half=False))
    # and the argmax call pattern still takes the fused kernel with the same answer
    p, mv, _ = O.max_preds(hm, "none")
    ka = np.concatenate([p, mv], 2)
    assert np.array_equal(adjust_keypoints_by_offset(cu(ka), cu(hm)).cpu().numpy(), O.refine_offset_clamped(ka, hm, plus_half=True))


def test_adjust_keypoints_by_dark_at_arbitrary_positions():
    """utils/heatmap_post_processing.py:35-76 refines around int(keypoints) wherever the caller puts them
    (ResultParser.candidate_bbox passes top-k candidates): positions one or two pixels off the blob maximum, on the
    border (guard fails: unchanged) and the exact argmax (fused-kernel path) — against the oracle."""
    from litehandnet_b200 import synth
    from litehandnet_b200.decode import adjust_keypoints_by_DARK
    B, K, H, W = 3, 5, 64, 64
    hm, centers = synth.blob_heatmaps(B, K, H, W, seed=31)
    hm = hm.numpy()
    p, mv, _ = O.max_preds(hm, "none")
    rng = np.random.default_rng(32)
    kp = np.concatenate([p + rng.integers(-2, 3, p.shape).astype(F32), mv], 2)
    kp[..., :2] = np.clip(kp[..., :2], 0, 63)
    kp[0, 0, :2] = (1, 30); kp[0, 1, :2] = (30, 62)                      # guard fails
    with np.errstate(all="ignore"):
        want = O.refine_dark(kp, hm, 19, legacy=True)
    got = adjust_keypoints_by_DARK(cu(kp), cu(hm))
    assert isinstance(got, np.ndarray) and got.shape == kp.shape
    assert_coords_close(got, want, what="dark at points")
    assert np.array_equal(got[0, :2], kp[0, :2]) and np.array_equal(got[..., 2], kp[..., 2])
    ka = np.concatenate([p, mv], 2)
    with np.errstate(all="ignore"):
        assert_coords_close(adjust_keypoints_by_DARK(cu(ka), cu(hm)), O.refine_dark(ka, hm, 19, legacy=True), what="dark at argmax")
    # bf16 planes, strided channel view
    big = torch.cat([cu(hm), torch.rand(B, 2, H, W, device=DEV)], 1).to(torch.bfloat16)
    hb = big[:, :K]
    with np.errstate(all="ignore"):
        want_b = O.refine_dark(kp, hb.float().cpu().numpy(), 19, legacy=True)
    assert_coords_close(adjust_keypoints_by_DARK(cu(kp), hb), want_b, rtol=1e-5, atol=5e-5, what="dark at points, bf16")
