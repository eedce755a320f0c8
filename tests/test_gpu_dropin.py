"""GPU: the drop-in mirrors of the reference interface (same names, arguments, return structure)
against the golden vectors of the executed reference — these read like the reference's own call sites
(test.py:114-135, train/topdown_trainer.py:34, utils/SPheatmapParser.py:220-240)."""
import numpy as np
import pytest
import torch

from conftest import assert_coords_close, load_golden
from oracle import np_oracle as O
from oracle.ref_loader import _AttrDict
from litehandnet_b200 import synth

pytestmark = pytest.mark.gpu
DEV = "cuda"


def cu(a):
    return torch.as_tensor(np.ascontiguousarray(a)).to(DEV)


def make_cfg(K=21, hm=(64, 64), img=(256, 256), unbiased=True, k=0, model="litehandnet", loss_weight=(1.0, 1.0)):
    return _AttrDict(
        DATASET=dict(image_size=list(img), heatmap_size=list(hm), num_joints=K),
        PIPELINE=dict(unbiased_encoding=unbiased, kernel=(11, 11), use_udp=False, simdr_split_ratio=k, sigma=2),
        MODEL=dict(name=model), LOSS=dict(type="TopdownHeatmapLoss", loss_weight=list(loss_weight), auto_weight=False))


# ---- decode ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["decode_64.npz", "decode_56.npz", "decode_mpii16.npz"])
def test_keypoints_from_heatmaps_numpy_in_numpy_out(case):
    from litehandnet_b200.decode import keypoints_from_heatmaps, _get_max_preds
    g = load_golden(case)
    before = g["hm"].copy()
    for pp, tag in (("default", "default"), ("unbiased", "unbiased"), (None, "none")):
        hp, p, mv = keypoints_from_heatmaps(g["hm"], g["center"], g["scale"], post_process=pp, kernel=11)
        assert isinstance(p, np.ndarray) and p.dtype == np.float32 and p.shape == g[f"ref_g2_{tag}_preds"].shape
        assert_coords_close(hp, g[f"ref_g2_{tag}_hm_preds"], what=tag)
        assert_coords_close(p, g[f"ref_g2_{tag}_preds"], what=tag)
        assert np.array_equal(mv, g[f"ref_g2_{tag}_maxvals"], equal_nan=True)
    p2, mv2 = keypoints_from_heatmaps(g["hm"], g["center"], g["scale"], only_original_preds=True)
    assert np.array_equal(p2, g["ref_g2_default_preds"], equal_nan=True)
    assert np.array_equal(g["hm"], before, equal_nan=True), "inputs must not be modified"
    preds, maxvals = _get_max_preds(g["hm"])
    assert np.array_equal(preds, g["ref_g2_none_hm_preds"]) and maxvals.shape == (g["hm"].shape[0], g["hm"].shape[1], 1)
    with pytest.raises(AssertionError):
        _get_max_preds(g["hm"][0])


def test_keypoints_from_heatmaps_cuda_in_cuda_out():
    from litehandnet_b200.decode import keypoints_from_heatmaps
    g = load_golden("decode_64.npz")
    hp, p, mv = keypoints_from_heatmaps(cu(g["hm"]), cu(g["center"]), cu(g["scale"]), post_process="default")
    assert p.is_cuda and np.array_equal(p.cpu().numpy(), g["ref_g2_default_preds"], equal_nan=True)


def test_topdown_decoder_facade():
    from litehandnet_b200.decode import TopDownDecoder
    g = load_golden("decode_64.npz")
    N, K = g["hm"].shape[:2]
    out = torch.cat([cu(g["hm"]), torch.rand(N, 3, 64, 64, device=DEV)], 1)     # extra channels are sliced off
    meta = dict(bbox_score=torch.ones(N), bbox_id=torch.arange(N), image_file=[f"{i}.jpg" for i in range(N)],
                center=torch.from_numpy(g["center"]), scale=torch.from_numpy(g["scale"]))
    for unbiased, tag in ((True, "unbiased"), (False, "default")):
        r = TopDownDecoder(make_cfg(K=K, unbiased=unbiased)).decode(meta, out)
        assert set(r) == {"preds", "hm_preds", "boxes", "image_paths", "bbox_ids", "output_heatmap"}
        assert r["preds"].shape == (N, K, 3) and r["preds"].dtype == np.float32
        assert_coords_close(r["preds"][..., :2], g[f"ref_g2_{tag}_preds"], what="decoder preds")
        assert np.array_equal(r["preds"][..., 2:], g[f"ref_g2_{tag}_maxvals"], equal_nan=True)
        assert_coords_close(r["hm_preds"][..., :2], g[f"ref_g2_{tag}_hm_preds"] * 4, what="decoder hm_preds")
        want_boxes = np.zeros((N, 6), np.float32)
        want_boxes[:, 0:2] = g["center"]; want_boxes[:, 2:4] = g["scale"]
        want_boxes[:, 4] = np.prod(g["scale"] * 200.0, axis=1); want_boxes[:, 5] = 1
        assert np.array_equal(r["boxes"], want_boxes) and r["bbox_ids"] == list(range(N))
        assert np.array_equal(r["output_heatmap"], g["hm"], equal_nan=True)


def test_decode_simdr_facade():
    from litehandnet_b200.decode import TopDownDecoder, keypoints_from_simdr
    g = load_golden("render_loss_64.npz")
    out = keypoints_from_simdr(g["simdr_xv"], g["simdr_yv"], g["center"], g["scale"], 2)
    assert isinstance(out, np.ndarray) and np.array_equal(out, g["ref_simdr_decode"])
    N = g["simdr_xv"].shape[0]
    meta = dict(bbox_score=torch.ones(N), bbox_id=torch.arange(N), image_file=["a"] * N,
                center=torch.from_numpy(g["center"]), scale=torch.from_numpy(g["scale"]),
                simdr_x=torch.from_numpy(g["simdr_xv"]), simdr_y=torch.from_numpy(g["simdr_yv"]))
    r = TopDownDecoder(make_cfg(K=8, k=2)).decode_simdr(meta, torch.zeros(N, 8, 64, 64, device=DEV))
    assert np.array_equal(r["preds"], g["ref_simdr_decode"])


@pytest.mark.parametrize("case", ["decode_64.npz", "decode_56.npz"])
def test_legacy_parsers(case):
    from litehandnet_b200.decode import (ResultParser, HeatmapParser_SH, adjust_keypoints_by_offset,
                                         adjust_keypoints_by_DARK, get_max_preds, get_final_preds, flip_back,
                                         get_coordinates_from_heatmap)
    g = load_golden(case)
    hm = cu(g["hm"])
    H, W = g["hm"].shape[2:]
    isz = [int(v) for v in g["image_size"]]
    cfg = dict(image_size=isz, hm_size=[W, H], model="litehandnet", simdr_split_ratio=2, bbox_alpha=1.0,
               with_region_map=False, cycle_detection_reduction=1, DARK=False)
    rp = ResultParser(cfg)
    kpts = rp.get_coordinates_from_heatmaps(hm)
    assert kpts.is_cuda and np.array_equal(kpts.cpu().numpy()[..., :2], O.max_preds(g["hm"], "none")[0])
    assert np.array_equal(rp.get_pred_kpt(hm).cpu().numpy(), g["ref_legacy_offset_hm"], equal_nan=True)
    assert np.array_equal(rp.get_pred_kpt(hm, resized=True).cpu().numpy(), g["ref_legacy_offset_img"], equal_nan=True)
    rp_dark = ResultParser(dict(cfg, DARK=True))
    assert_coords_close(rp_dark.get_pred_kpt(hm, resized=True).cpu().numpy(), g["ref_legacy_dark_img"], what="DARK")
    # the two-step legacy form: argmax -> adjust
    adj = adjust_keypoints_by_offset(kpts.clone(), hm)
    assert np.array_equal(adj.cpu().numpy(), g["ref_legacy_offset_hm"], equal_nan=True)
    dk = adjust_keypoints_by_DARK(kpts.clone(), hm)
    assert isinstance(dk, np.ndarray)
    assert_coords_close(dk, g["ref_legacy_dark_hm"], what="adjust DARK")
    assert np.array_equal(hm.cpu().numpy(), g["hm"], equal_nan=True), "heatmaps must stay untouched (CUDA semantics)"
    k, bb = HeatmapParser_SH().parse(hm, image_size=tuple(isz))
    assert bb is None and not k.is_cuda and np.array_equal(k.numpy(), g["ref_parse_sh"], equal_nan=True)
    c = HeatmapParser_SH.get_coordinates(hm)
    assert not c.is_cuda
    a = HeatmapParser_SH.adjust_keypoints(c.clone(), hm)
    assert np.array_equal((a.numpy()[..., :2] * (np.array(isz, np.float32) / np.array([W, H], np.float32))),
                          g["ref_parse_sh"][..., :2], equal_nan=True)
    p, mv = get_max_preds(g["hm"])
    assert np.array_equal(p, g["ref_a3_preds"]) and np.array_equal(mv, g["ref_a3_maxvals"], equal_nan=True)
    p, mv = get_coordinates_from_heatmap(hm)
    assert np.array_equal(p.cpu().numpy(), g["ref_a1_preds"])
    assert_coords_close(get_final_preds(hm, g["center"], g["scale"]).cpu().numpy(), g["ref_final_preds"],
                        rtol=1e-5, atol=1e-4, what="final preds")
    pairs = [tuple(int(v) for v in pr) for pr in g["flip_pairs"]]
    assert np.array_equal(flip_back(g["hm"], pairs), g["ref_flip_back"], equal_nan=True)
    # positions that are not the plane argmax are refined where they are (lhn_refine_points), as the reference does
    moved = kpts.clone()
    moved[..., :2] = torch.clamp(moved[..., :2] + 1.0, 0, min(H, W) - 1)
    with np.errstate(all="ignore"):
        want = O.refine_offset_clamped(moved.cpu().numpy(), g["hm"], plus_half=True)
    assert np.array_equal(adjust_keypoints_by_offset(moved, hm).cpu().numpy(), want, equal_nan=True)


def test_sp_parser_main_fixture():
    """utils/SPheatmapParser.py:221-233."""
    from litehandnet_b200.decode import HeatmapParser_SH
    kpt_hm = torch.zeros((2, 4, 64, 64)); kpt_hm[..., 3, 3] = 1; kpt_hm[..., 3, 2] = 0.5; kpt_hm[..., 2, 3] = 0.5
    k, b = HeatmapParser_SH().parse(kpt_hm, image_size=(256, 256))
    assert b is None and np.array_equal(k.numpy(), load_golden("sp_parser_main.npz")["ref_kpt"])
    assert k[0, 0].tolist() == [11.0, 11.0, 1.0]


def test_vector_nms_and_coordinates_from_vectors():
    from litehandnet_b200.decode import ResultParser
    cfg = dict(image_size=[256, 256], hm_size=[64, 64], model="litehandnet", simdr_split_ratio=2, bbox_alpha=1.0,
               with_region_map=False, cycle_detection_reduction=1, DARK=False)
    rp = ResultParser(cfg)
    xv, yv = synth.simdr_vectors(4, 21, 512, seed=5)
    assert np.array_equal(rp.vector_nms(xv.to(DEV)).cpu().numpy(), O.vector_nms(xv.numpy()))
    bboxes = [[[128.0, 120.0, 100.0, 90.0, 1.0]], None, [[60.2, 200.7, 50.5, 80.0, 0.9]], [[250.0, 10.0, 40.0, 40.0, 0.5]]]
    got = rp.get_coordinates_from_vectors(xv.to(DEV), yv.to(DEV), bboxes).cpu().numpy()
    assert got.shape == (4, 1, 21, 3)
    ranges = np.zeros((4, 4), np.int64)
    for i, bb in enumerate(bboxes):
        if bb is None:
            continue
        b = np.round(np.array(bb[0]) * 2)
        x1, y1 = b[:2] - b[2:4] / 2; x2, y2 = b[:2] + b[2:4] / 2
        ranges[i] = (max(int(x1), 0), min(int(x2), 512), max(int(y1), 0), min(int(y2), 512))
    want = O.coordinates_from_vectors(xv.numpy(), yv.numpy(), ranges, 2)
    want[1] = 0
    assert np.array_equal(got[:, 0], want)


# ---- losses ------------------------------------------------------------------------------------------
def test_loss_modules_golden():
    from litehandnet_b200.loss import DistanceLoss, JointsDistanceLoss, KLDiscretLoss
    g = load_golden("render_loss_64.npz")
    hm = cu(np.nan_to_num(load_golden("decode_64.npz")["hm"], nan=0.25, posinf=1.0, neginf=-1.0))
    for tag in ("unbiased", "int"):
        t, w = cu(g[f"ref_target_{tag}"]), cu(g[f"ref_weight_{tag}"])
        for bal, bt in ((True, "bal"), (False, "nobal")):
            out = DistanceLoss("L2", "mean", bal)(hm, t, w)
            assert out.dim() == 0 and out.dtype == torch.float32 and out.is_cuda
            np.testing.assert_allclose(out.item(), g[f"ref_distance_loss_{tag}_{bt}"], rtol=1e-5)
        np.testing.assert_allclose(JointsDistanceLoss()(hm, t, w).item(), g[f"ref_joints_mse_{tag}"], rtol=1e-5)
    with pytest.raises(NameError):
        JointsDistanceLoss()(hm, cu(g["ref_target_int"]), None)
    with pytest.raises(AssertionError):
        DistanceLoss(reduction="avg")
    s = DistanceLoss("L2", "sum", True)(hm, cu(g["ref_target_unbiased"]), cu(g["ref_weight_unbiased"]))
    np.testing.assert_allclose(s.item(), float(g["ref_distance_loss_unbiased_bal"]) * hm.numel(), rtol=1e-5)
    # 5-D hourglass shape
    t5, w5 = O.render_targets(g["joints_3d"], g["joints_3d_visible"], (256, 256), (64, 64), [2, 2])
    o5 = torch.stack([hm, hm * 0.5], 1)
    np.testing.assert_allclose(DistanceLoss()(o5, cu(t5), cu(w5)).item(), g["ref_distance_loss_5d_bal"], rtol=1e-5)
    # fused entry == explicit-target entry
    fl, fw = DistanceLoss().forward_fused(hm, cu(g["joints_3d"]), cu(g["joints_3d_visible"]), (256, 256), 2, True)
    np.testing.assert_allclose(fl.item(), g["ref_distance_loss_unbiased_bal"], rtol=1e-5)
    assert np.array_equal(fw.cpu().numpy(), g["ref_weight_unbiased"])
    sx, sy = O.render_simdr_batch(g["joints_3d"], g["joints_3d_visible"], (256, 256), 2, 2)
    kl = KLDiscretLoss()(cu(g["simdr_xv"]), cu(g["simdr_yv"]), cu(sx), cu(sy), cu(g["ref_weight_unbiased"]))
    np.testing.assert_allclose(kl.item(), g["ref_kld_loss"], rtol=1e-5)


def test_topdown_heatmap_loss_wrapper():
    from litehandnet_b200.loss import get_loss
    g = load_golden("render_loss_64.npz")
    hm = cu(np.nan_to_num(load_golden("decode_64.npz")["hm"], nan=0.25, posinf=1.0, neginf=-1.0))
    cfg = make_cfg(K=8, loss_weight=(0.5, 1.0))
    crit = get_loss(cfg).cuda()
    meta = dict(target=torch.from_numpy(g["ref_target_unbiased"]), target_weight=torch.from_numpy(g["ref_weight_unbiased"]))
    loss, ld = crit(hm, meta)                                    # meta tensors on CPU, as the dataloader hands them
    np.testing.assert_allclose(loss.item(), 0.5 * float(g["ref_distance_loss_unbiased_bal"]), rtol=1e-5)
    assert set(ld) == {"heatmap"} and isinstance(ld["heatmap"], float)
    loss2, _ = crit(hm, dict(joints_3d=torch.from_numpy(g["joints_3d"]), joints_3d_visible=torch.from_numpy(g["joints_3d_visible"])))
    np.testing.assert_allclose(loss2.item(), loss.item(), rtol=1e-5)
    cfg_att = make_cfg(K=8, model="atthandnet")
    loss3, _ = get_loss(cfg_att).cuda()(hm, meta)
    np.testing.assert_allclose(loss3.item(), g["ref_distance_loss_unbiased_nobal"], rtol=1e-5)
    # SimDR branch: two nn.Linear heads (torch/cuBLAS) + our SmoothL1 reduction
    cfg_s = make_cfg(K=8, k=2)
    crit_s = get_loss(cfg_s).cuda()
    sx, sy = O.render_simdr_batch(g["joints_3d"], g["joints_3d_visible"], (256, 256), 2, 2)
    meta_s = dict(meta, simdr_x=torch.from_numpy(sx), simdr_y=torch.from_numpy(sy))
    loss4, ld4 = crit_s(hm, meta_s)
    with torch.no_grad():
        px = crit_s.simdr_loss.x_shared_decoder(hm.flatten(2)); py = crit_s.simdr_loss.y_shared_decoder(hm.flatten(2))
    want = O.kl_discret_loss(px.cpu().numpy(), py.cpu().numpy(), sx, sy, g["ref_weight_unbiased"])
    np.testing.assert_allclose(ld4["simdr"], want, rtol=1e-5)
    np.testing.assert_allclose(loss4.item(), ld4["heatmap"] + ld4["simdr"], rtol=1e-6)


def test_srhandnet_loss_multiscale():
    from litehandnet_b200.loss import SRHandNetLoss
    cfg = _AttrDict(MODEL=dict(output_channel=21, pred_bbox=False), LOSS=dict(loss_weight=[0.3, 0.3, 0.5, 1.0]))
    sizes = [16, 16, 32, 64]
    j, v = synth.hand_joints(3, 21, seed=7)
    outs, tgs, tws, want = [], [], [], 0.0
    for i, s in enumerate(sizes):
        o, _ = synth.blob_heatmaps(3, 21, s, s, seed=40 + i, margin=2.0)
        t, w = O.render_targets(j.numpy(), v.numpy(), (256, 256), (s, s), 2, False)
        outs.append(o.to(DEV)); tgs.append(torch.from_numpy(t)); tws.append(torch.from_numpy(w))
        want += float(O.distance_loss_l2(o.numpy(), t, w, True)) * cfg.LOSS.loss_weight[i]
    loss, ld = SRHandNetLoss(cfg)(outs, dict(target=tgs, target_weight=tws))
    np.testing.assert_allclose(loss.item(), want, rtol=1e-5)
    assert set(ld) == {"kpt_loss"}


# ---- metrics -------------------------------------------------------------------------------------------
def test_metric_functions_golden():
    from litehandnet_b200.metrics import (keypoint_pck_accuracy, keypoint_auc, keypoint_epe, report_metric,
                                          evaluate_pck)
    g = load_golden("metrics_16.npz")
    p64 = g["preds"].astype(np.float64)
    t = g["bbox_wh"].max(1).astype(np.float64)
    nor = np.stack([t, t], 1)
    acc, avg, cnt = keypoint_pck_accuracy(p64, g["gt"], g["mask"], 0.2, nor)
    assert np.array_equal(acc, g["ref_pck_acc"]) and avg == g["ref_pck_avg"] and cnt == g["ref_pck_cnt"]
    hs = g["head_size"]
    acc, avg, _ = keypoint_pck_accuracy(p64, g["gt"], g["mask"], 0.5, np.stack([hs, hs], 1))
    assert np.array_equal(acc, g["ref_pckh_acc"]) and avg == g["ref_pckh_avg"]
    assert keypoint_auc(p64, g["gt"], g["mask"], 30) == g["ref_auc"]
    np.testing.assert_allclose(keypoint_epe(p64, g["gt"], g["mask"]), g["ref_epe"], rtol=1e-5)
    info = dict(report_metric(p64, g["gt"], g["mask"], bbox_wh=g["bbox_wh"], head_size=hs,
                              metrics=("PCK", "PCKh", "AUC", "EPE")))
    assert info["PCK"] == g["ref_pck_avg"] and info["PCKh"] == g["ref_pckh_avg"] and info["AUC"] == g["ref_auc"]
    with np.errstate(all="ignore"):
        a = evaluate_pck(torch.from_numpy(g["pck_pred_hm"]), torch.from_numpy(g["pck_gt_hm"]), torch.from_numpy(g["pck_bbox"]),
                         256, torch.from_numpy(g["pck_tw"]), 0.2)
        b = evaluate_pck(cu(g["pck_pred_hm"]), cu(g["pck_gt_hm"]), cu(g["pck_bbox"]), 256, None, 0.02)
    np.testing.assert_allclose(a, g["ref_evaluate_pck_w"], rtol=1e-6, equal_nan=True)
    np.testing.assert_allclose(b, g["ref_evaluate_pck_now"], rtol=1e-6, equal_nan=True)


def test_metric_accumulator_sharded_equals_monolithic():
    """BASELINE config 4: MPII-style 16-joint decode + PCK@0.2/AUC/EPE, batch-sharded (8 shards here on
    one GPU; the NCCL all-reduce of the same int64 block is covered by the gloo test on CPU)."""
    from litehandnet_b200.metrics import MetricAccumulator
    N, K = 256, 16
    hm, cen = synth.blob_heatmaps(N, K, 64, 64, seed=101)
    center, scale = synth.bbox_center_scale(N, fixed=True)
    gt, mask, wh = synth.pck_inputs(cen, seed=102)
    hm, center, scale, gt, mask, wh = [t.to(DEV) for t in (hm, center, scale, gt, mask, wh)]
    mono = MetricAccumulator(K)
    r = mono.update_from_heatmaps(hm, center, scale, gt, mask, wh)
    shard = MetricAccumulator(K)
    for s in np.array_split(np.arange(N), 8):
        sl = slice(int(s[0]), int(s[-1]) + 1)
        shard.update_from_heatmaps(hm[sl], center[sl], scale[sl], gt[sl], mask[sl], wh[sl])
    assert torch.equal(mono.counters, shard.counters)
    from_preds = MetricAccumulator(K)
    from_preds.update_from_preds(r["kpts"], gt, mask, wh)
    assert torch.equal(mono.counters, from_preds.counters)
    out = mono.compute()
    preds = r["kpts"][..., :2].cpu().numpy().astype(np.float64)
    want = dict(O.report_metric(preds, gt.cpu().numpy(), mask.cpu().numpy(), bbox_wh=wh.cpu().numpy().astype(np.float64)))
    assert out["PCK"] == want["PCK"] and out["AUC"] == want["AUC"]
    np.testing.assert_allclose(out["EPE"], want["EPE"], rtol=1e-5)


# ---- render ---------------------------------------------------------------------------------------------
def test_render_transforms_dict_in_dict_out():
    from litehandnet_b200.render import TopDownGenerateTarget, GenerateSimDR
    g = load_golden("render_loss_64.npz")
    ann = dict(num_joints=8, image_size=np.array([256, 256]), heatmap_size=np.array([64, 64]), joint_weights=None,
               use_different_joint_weights=False)
    for unb, tag in ((True, "unbiased"), (False, "int")):
        res = TopDownGenerateTarget(sigma=2, unbiased_encoding=unb)(
            dict(joints_3d=g["joints_3d"][0], joints_3d_visible=g["joints_3d_visible"][0], ann_info=ann))
        assert res["target"].shape == (8, 64, 64) and res["target_weight"].shape == (8, 1)
        np.testing.assert_allclose(res["target"], g[f"ref_target_{tag}"][0], rtol=1e-5, atol=1e-7)
        assert np.array_equal(res["target_weight"], g[f"ref_weight_{tag}"][0])
    res = TopDownGenerateTarget(sigma=[2, 2], unbiased_encoding=True)(
        dict(joints_3d=g["joints_3d"][1], joints_3d_visible=g["joints_3d_visible"][1], ann_info=ann))
    assert res["target"].shape == (2, 8, 64, 64) and res["target_weight"].shape == (2, 8, 1)
    res = GenerateSimDR(sigma=2, k=2)(dict(joints_3d=g["joints_3d"][0], joints_3d_visible=g["joints_3d_visible"][0], ann_info=ann))
    np.testing.assert_allclose(res["simdr_x"][0], g["ref_simdr_x_row0"][0], rtol=1e-5, atol=1e-7)


# ---- the headline call at full size: size-independent properties -------------------------------------------
def test_full_size_config2_properties():
    """B=1024 x 21 x 64 x 64 (BASELINE config 2): the oracle is too slow here, so check properties:
    batch-permutation equivariance, shard-additivity of the loss sums, flip symmetry, determinism."""
    from litehandnet_b200 import fused
    B, K = 1024, 21
    hm, cen = synth.blob_heatmaps(B, K, 64, 64, seed=7, device=DEV)
    hf = synth.flipped_blob_heatmaps(cen, 64, 64, seed=8, device=DEV)
    j, v = synth.hand_joints(B, K, seed=9, device=DEV)
    c, s = synth.bbox_center_scale(B, seed=10, device=DEV)
    step = fused.FusedHeatmapStep()
    a = step(hm, j, v, c, s, hf)
    b = step(hm, j, v, c, s, hf)
    assert torch.equal(a["preds"], b["preds"]) and torch.equal(a["loss"], b["loss"]), "must be deterministic"
    perm = torch.randperm(B, device=DEV)
    p = step(hm[perm], j[perm], v[perm], c[perm], s[perm], hf[perm])
    assert torch.equal(p["preds"], a["preds"][perm]) and torch.equal(p["idx"], a["idx"][perm])
    np.testing.assert_allclose(p["loss"].item(), a["loss"].item(), rtol=1e-6)
    sums = torch.zeros(4, dtype=torch.float64, device=DEV)
    for sl in (slice(0, 300), slice(300, 1024)):
        sums += step(hm[sl], j[sl], v[sl], c[sl], s[sl], hf[sl])["loss_sums"]
    assert torch.allclose(sums, a["loss_sums"], rtol=1e-12)
    # argmax of the average equals argmax computed by torch on the same average (first-index ties)
    avg = (hm + hf.flip(-1)) * 0.5
    assert torch.equal(a["idx"].long(), avg.flatten(2).argmax(-1))
    # a slice of the batch against the oracle
    n = 6
    with np.errstate(all="ignore"):
        ref = O.fused_render_loss_decode(hm[:n].cpu().numpy(), hf[:n].cpu().numpy(), j[:n].cpu().numpy(), v[:n].cpu().numpy(),
                                         c[:n].cpu().numpy(), s[:n].cpu().numpy())
    assert_coords_close(a["preds"][:n, :, :2].cpu().numpy(), ref["preds"], what="config-2 preds")


def test_evaluate_results_like_dataset_evaluate():
    """test.py:114-135: results.append(decoder.decode(meta, outputs)); dataset.evaluate(results, out, metric) —
    on shuffled, partly duplicated batches (the last batch of a DistributedSampler repeats samples), against
    the reference's _report_metric arithmetic (oracle) on the de-duplicated, id-sorted predictions."""
    from litehandnet_b200 import decode as D, metrics as M
    N, K = 96, 21
    hm, cen = synth.blob_heatmaps(N, K, 64, 64, seed=91)
    c, s = synth.bbox_center_scale(N, seed=92)
    gt, mask, wh = synth.pck_inputs(cen, seed=93)
    db = [dict(joints_3d=np.concatenate([gt[i].numpy(), np.zeros((K, 1), np.float32)], 1),
               joints_3d_visible=np.repeat(mask[i].numpy().astype(np.float32)[:, None], 3, 1),
               bbox=np.array([10.0, 20.0, float(wh[i, 0]), float(wh[i, 1])]), bbox_id=i) for i in range(N)]
    dec = D.TopDownDecoder(make_cfg(K=K, unbiased=False))             # post_process='default'
    perm = torch.randperm(N, generator=torch.Generator().manual_seed(0))
    perm = torch.cat([perm, perm[:16]])                               # 16 duplicates at the end
    results = []
    for b in range(0, len(perm), 32):
        idx = perm[b:b + 32]
        meta = dict(center=c[idx], scale=s[idx], image_file=[f"img{int(i)}.jpg" for i in idx],
                    bbox_score=torch.ones(len(idx)), bbox_id=[int(i) for i in idx])
        results.append(dec.decode(meta, hm[idx].to(DEV)))
    got = M.evaluate_results(results, db, ["PCK", "AUC", "EPE"])
    _, preds, _ = O.keypoints_from_heatmaps(hm.numpy(), c.numpy(), s.numpy(), "default", 11)
    p64 = preds.astype(np.float64)
    t = wh.max(1).values.double().numpy()
    _, pck, _ = O.keypoint_pck_accuracy(p64, gt.numpy(), mask.numpy(), 0.2, np.stack([t, t], 1))
    assert list(got.keys()) == ["PCK", "AUC", "EPE"]
    np.testing.assert_allclose(got["PCK"], pck, rtol=1e-12)
    np.testing.assert_allclose(got["AUC"], O.keypoint_auc(p64, gt.numpy(), mask.numpy(), 30), rtol=1e-12)
    np.testing.assert_allclose(got["EPE"], O.keypoint_epe(p64, gt.numpy(), mask.numpy()), rtol=1e-5)


def test_get_final_preds_uses_scale0_only():
    """utils/transforms.py:18-44,112-152: the cv2-affine back-transform of the legacy path is a similarity built
    from scale[0] alone; anisotropic scales and a non-square heatmap (oracle checked live against the reference)."""
    from litehandnet_b200.decode import get_final_preds
    for shape, seed in (((5, 16, 64, 64), 150), ((4, 16, 64, 48), 151)):
        hm, _ = synth.blob_heatmaps(*shape, seed=seed, zero_frac=0.1, tie_frac=0.1)
        rng = np.random.default_rng(seed)
        c = rng.uniform(60, 200, (shape[0], 2)).astype(np.float32)
        s = np.stack([rng.uniform(0.6, 1.6, shape[0]), rng.uniform(0.6, 1.6, shape[0])], 1).astype(np.float32)
        got = get_final_preds(hm.to(DEV), cu(c), cu(s)).cpu().numpy()
        assert np.array_equal(got, O.get_final_preds(hm.numpy(), c, s))


def test_mpii_evaluate_golden_and_sharded():
    """TopDownMpiiDataset.evaluate arithmetic (topdown_mpii_dataset.py:126-249) from shuffled, partly duplicated
    result batches; counters of two shards add up to the monolithic ones."""
    from litehandnet_b200 import metrics as M
    g = load_golden("mpii_pckh.npz")
    gt = dict(dataset_joints=g["dataset_joints"], jnt_missing=g["jnt_missing"], pos_gt_src=g["pos_gt_src"],
              headboxes_src=g["headboxes_src"])
    ids = [int(i) for i in g["bbox_ids"]]
    results = [dict(preds=g["preds"][ids[a:a + 16]], bbox_ids=ids[a:a + 16]) for a in range(0, len(ids), 16)]
    out = M.mpii_evaluate(results, gt)
    assert list(out.keys()) == list(g["ref_names"])
    np.testing.assert_allclose(np.array([float(v) for v in out.values()]), g["ref_values"], rtol=1e-12)
    thr = [0.5, 0.1]
    mono = M.mpii_pckh_counters(g["preds"], gt, thr)
    parts = None
    for sl in (slice(0, 17), slice(17, 40)):
        sub = dict(dataset_joints=g["dataset_joints"], jnt_missing=g["jnt_missing"][:, sl],
                   pos_gt_src=g["pos_gt_src"][:, :, sl], headboxes_src=g["headboxes_src"][:, :, sl])
        parts = M.mpii_pckh_counters(g["preds"][sl], sub, thr, counters=parts)
    assert torch.equal(mono, parts)


def test_udp_decode_golden_and_oracle():
    """keypoints_from_heatmaps(..., use_udp=True) (top_down_eval.py:427-431 -> post_dark_udp) against the executed
    reference and, on a larger random batch and with the flip average, against the oracle."""
    from litehandnet_b200 import decode as D
    g = load_golden("decode_udp.npz")
    for k in (11, 17):
        hp, p, mv = D.keypoints_from_heatmaps(g["hm"], g["center"], g["scale"], post_process="default", kernel=k, use_udp=True)
        assert np.array_equal(mv, g["ref_udp_maxvals"])
        assert_coords_close(hp, g[f"ref_udp_hm_preds_k{k}"], what=f"udp hm k={k}")
        assert_coords_close(p, g[f"ref_udp_preds_k{k}"], what=f"udp img k={k}")
    for shape in ((40, 21, 64, 64), (9, 16, 56, 56), (6, 5, 28, 28)):
        hm, cen = synth.blob_heatmaps(*shape, seed=101, zero_frac=0.05, tie_frac=0.03)
        hf = synth.flipped_blob_heatmaps(cen, shape[2], shape[3], seed=102)
        c, s = synth.bbox_center_scale(shape[0], seed=103)
        hp, p, mv = D.keypoints_from_heatmaps(hm.to(DEV), c, s, kernel=11, use_udp=True, heatmaps_flipped=hf.to(DEV))
        avg = O.flip_average(hm.numpy(), hf.numpy(), ())
        with np.errstate(all="ignore"):
            rhp, rp_, rmv = O.keypoints_from_heatmaps_udp(avg, c.numpy(), s.numpy(), 11)
        assert np.array_equal(mv.cpu().numpy(), rmv)
        assert_coords_close(hp.cpu().numpy(), rhp, what="udp hm vs oracle")
        assert_coords_close(p.cpu().numpy(), rp_, what="udp img vs oracle")


def test_udp_target_transform_dict_in_dict_out():
    """TopDownGenerateTarget(encoding='UDP') as the dataset pipeline calls it (generateTarget.py:245-300)."""
    from litehandnet_b200 import render as R
    g = load_golden("render_udp.npz")
    isz, hsz = g["image_size"], g["heatmap_size"]
    for sg, tag in ((2, "s2"), ([2, 3], "list")):
        gen = R.TopDownGenerateTarget(sigma=sg, encoding="UDP", target_type="GaussianHeatmap")
        for b in range(g["joints_3d"].shape[0]):
            out = gen(dict(joints_3d=g["joints_3d"][b], joints_3d_visible=g["joints_3d_visible"][b],
                           ann_info=dict(num_joints=8, image_size=isz, heatmap_size=hsz, joint_weights=None,
                                         use_different_joint_weights=False)))
            assert np.array_equal(out["target_weight"], g[f"ref_weight_{tag}"][b])
            assert np.abs(out["target"] - g[f"ref_target_{tag}"][b]).max() <= 1.2e-7


def test_full_size_config1_and_3_properties():
    """BASELINE config 1 (64 x 21 x 64 x 64, argmax + quarter offset) against the oracle in full, and config 3
    (SimDR, 2 x [4096, 21, 512], k = 2) through torch.argmax as an independent first-index argmax plus the oracle
    on a slice."""
    from litehandnet_b200 import decode as D
    hm, _ = synth.blob_heatmaps(64, 21, 64, 64, seed=11, zero_frac=0.02, tie_frac=0.01)
    c, s = synth.bbox_center_scale(64, seed=12)
    hp, p, mv = D.keypoints_from_heatmaps(hm.numpy(), c.numpy(), s.numpy(), post_process="default")
    rhp, rp_, rmv = O.keypoints_from_heatmaps(hm.numpy(), c.numpy(), s.numpy(), "default", 11)
    assert np.array_equal(hp, rhp) and np.array_equal(p, rp_) and np.array_equal(mv, rmv)
    rpk = D.ResultParser(dict(image_size=[256, 256], hm_size=[64, 64], model="litehandnet", simdr_split_ratio=2,
                              bbox_alpha=1.0, with_region_map=False, cycle_detection_reduction=1, DARK=False))
    k = rpk.get_pred_kpt(hm.to(DEV), resized=True)
    assert np.array_equal(k.cpu().numpy(), O.get_pred_kpt(hm.numpy(), dark=False, resized=True, feature_stride=(4, 4)))
    B, K, Lv = 4096, 21, 512
    xv, yv = synth.simdr_vectors(B, K, Lv, seed=13, device=DEV)
    c, s = synth.bbox_center_scale(B, seed=14, device=DEV)
    out = D.keypoints_from_simdr(xv, yv, c, s, k=2)
    ix, iy = xv.argmax(-1), yv.argmax(-1)
    score = (xv.amax(-1) + yv.amax(-1)) / 2
    assert torch.equal(out[..., 2], score)
    # undo transform_preds on a slice through the oracle
    n = 64
    ref = O.keypoints_from_simdr(xv[:n].cpu().numpy(), yv[:n].cpu().numpy(), c[:n].cpu().numpy(), s[:n].cpu().numpy(), 2)
    assert np.array_equal(out[:n].cpu().numpy(), ref)
    # and the raw indices (bit-exact against torch's first-index argmax) through the ops layer
    from litehandnet_b200 import ops as OPS
    r = OPS.decode_simdr(xv, yv, 2, None, None, want_idx=True)
    idx = r[1] if isinstance(r, tuple) else r["idx"]
    assert torch.equal(idx[..., 0].long().reshape(B, K), ix) and torch.equal(idx[..., 1].long().reshape(B, K), iy)


def test_full_size_config4_sharded_counters():
    """BASELINE config 4: MPII 16 x 64 x 64, B = 8192 in 8 shards of 1024 — fused decode + PCK/AUC/EPE counters per
    shard, summed, must equal the monolithic counters bit for bit and the metric functions on the decoded points."""
    from litehandnet_b200 import metrics as M
    B, K = 8192, 16
    hm, cen = synth.blob_heatmaps(B, K, 64, 64, seed=15, device=DEV)
    c, s = synth.bbox_center_scale(B, seed=16, device=DEV)
    gt, mask, wh = synth.pck_inputs(cen, seed=17, device=DEV)
    mono = M.MetricAccumulator(K, device=DEV)
    mono.update_from_heatmaps(hm, c, s, gt, mask, wh, post_process="default")
    shard = M.MetricAccumulator(K, device=DEV)
    for r in range(8):
        sl = slice(r * 1024, (r + 1) * 1024)
        part = M.MetricAccumulator(K, device=DEV)
        part.update_from_heatmaps(hm[sl], c[sl], s[sl], gt[sl], mask[sl], wh[sl], post_process="default")
        shard.counters += part.counters
    assert torch.equal(mono.counters, shard.counters), "integer counters must be shard-additive"
    res = mono.compute(("PCK", "AUC", "EPE"))
    # independent route: decode, then the metric functions on the points (float64 like the JSON round trip)
    from litehandnet_b200 import decode as D
    _, preds, _ = D.keypoints_from_heatmaps(hm, c, s, post_process="default")
    p64 = preds.double().cpu().numpy()
    t = wh.max(1).values.double().cpu().numpy()
    acc, avg, cnt = O.keypoint_pck_accuracy(p64, gt.cpu().numpy(), mask.cpu().numpy(), 0.2, np.stack([t, t], 1))
    got = dict(res)
    np.testing.assert_allclose(got["PCK"], avg, rtol=1e-12)
    np.testing.assert_allclose(got["AUC"], O.keypoint_auc(p64, gt.cpu().numpy(), mask.cpu().numpy(), 30), rtol=1e-12)
    np.testing.assert_allclose(got["EPE"], O.keypoint_epe(p64, gt.cpu().numpy(), mask.cpu().numpy()), rtol=1e-5)


def test_full_size_config5_properties():
    """BASELINE config 5 shape, one GPU's shard: 1024 x 21 x 128 x 128 f32 (1.4 GB), fused render + loss + DARK."""
    from litehandnet_b200 import fused
    B, K, H, W = 1024, 21, 128, 128
    hm, cen = synth.blob_heatmaps(B, K, H, W, seed=18, device=DEV, sigma=4.0)
    j, v = synth.hand_joints(B, K, (512, 512), seed=19, device=DEV)
    c, s = synth.bbox_center_scale(B, seed=20, device=DEV)
    step = fused.FusedHeatmapStep(image_size=(512, 512), sigma=4)
    a = step(hm, j, v, c, s)
    assert torch.equal(a["idx"].long(), hm.flatten(2).argmax(-1)), "argmax must equal torch's first-index argmax"
    b = step(hm, j, v, c, s)
    assert torch.equal(a["preds"], b["preds"]) and torch.equal(a["loss"], b["loss"])
    sums = torch.zeros(4, dtype=torch.float64, device=DEV)
    for sl in (slice(0, 500), slice(500, 1024)):
        sums += step(hm[sl], j[sl], v[sl], c[sl], s[sl])["loss_sums"]
    assert torch.allclose(sums, a["loss_sums"], rtol=1e-12)
    n = 3
    with np.errstate(all="ignore"):
        ref = O.fused_render_loss_decode(hm[:n].cpu().numpy(), None, j[:n].cpu().numpy(), v[:n].cpu().numpy(),
                                         c[:n].cpu().numpy(), s[:n].cpu().numpy(), image_size=(512, 512), sigma=4)
    assert_coords_close(a["preds"][:n, :, :2].cpu().numpy(), ref["preds"], what="config-5 preds")
    sub = step(hm[:n], j[:n], v[:n], c[:n], s[:n])
    np.testing.assert_allclose(sub["loss"].item(), float(ref["loss"]), rtol=1e-5)


# ---- autograd through the drop-in losses -----------------------------------------------
def _close(a, b, what, rtol=1e-5):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert np.abs(a - b).max() <= rtol * np.abs(b).max() + 1e-12, (what, np.abs(a - b).max(), np.abs(b).max())


def test_loss_backward_golden():
    """loss.backward() through DistanceLoss / JointsDistanceLoss / KLDiscretLoss against torch autograd through
    the executed reference modules (tests/golden/loss_backward.npz)."""
    from litehandnet_b200 import loss as LS
    g = load_golden("loss_backward.npz")
    tg, tw = cu(g["target"]), cu(g["target_weight"])
    for tag, crit in (("bal", LS.DistanceLoss(balance=True)), ("nobal", LS.DistanceLoss(balance=False)),
                      ("sum", LS.DistanceLoss(balance=True, reduction="sum")), ("jmse", LS.JointsDistanceLoss())):
        x = cu(g["hm"]).requires_grad_(True)
        loss = crit(x, tg, tw)
        (loss * 0.7).backward()
        np.testing.assert_allclose(loss.item(), g[f"ref_loss_{tag}"], rtol=1e-5)
        _close(x.grad.cpu().numpy(), g[f"ref_grad_{tag}"], tag)
    px, py = cu(g["simdr_out_x"]).requires_grad_(True), cu(g["simdr_out_y"]).requires_grad_(True)
    loss = LS.KLDiscretLoss()(px, py, cu(g["simdr_tgt_x"]), cu(g["simdr_tgt_y"]), tw)
    (loss * 1.3).backward()
    np.testing.assert_allclose(loss.item(), g["ref_loss_simdr"], rtol=1e-5)
    _close(px.grad.cpu().numpy(), g["ref_grad_simdr_x"], "simdr x")
    _close(py.grad.cpu().numpy(), g["ref_grad_simdr_y"], "simdr y")
    # fused entry: the target rendered in-kernel from the joints must give the same gradient
    x = cu(g["hm"]).requires_grad_(True)
    w_frac = g["target_weight"].copy()
    vis = g["joints_3d_visible"].copy()
    loss, w = LS.DistanceLoss(balance=True).forward_fused(x, cu(g["joints_3d"]), cu(vis), tuple(int(v) for v in g["image_size"]),
                                                          sigma=float(g["sigma"]))
    (loss * 0.7).backward()
    # the golden case carries one hand-edited fractional weight (0.5): compare on the other planes
    keep = (w_frac[..., 0] != 0.5)
    ref_w = O.render_targets(g["joints_3d"], vis, tuple(int(v) for v in g["image_size"]), (32, 32), float(g["sigma"]), True)[1]
    assert np.array_equal(w.cpu().numpy(), ref_w)
    want = O.distance_loss_l2_grad(g["hm"], g["target"], ref_w, True, grad_out=0.7)
    _close(x.grad.cpu().numpy(), want, "fused backward")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(16, 21, 64, 64), (3, 2, 21, 64, 64), (5, 6, 56, 56)])
def test_loss_backward_vs_oracle(shape, dtype):
    from litehandnet_b200 import loss as LS
    stacked = len(shape) == 5
    N, K, H, W = (shape[0], shape[2], shape[3], shape[4]) if stacked else shape
    S = shape[1] if stacked else 1
    hm, _ = synth.blob_heatmaps(N, S * K, H, W, seed=71)
    joints, vis = synth.hand_joints(N, K, (4 * W, 4 * H), seed=72, vis_prob=0.9, outside_frac=0.05)
    sig = [2.0, 3.0][:S] if stacked else 2
    tg, tw = O.render_targets(joints.numpy(), vis.numpy(), (4 * W, 4 * H), (W, H), sig, True)
    hm = hm.reshape(tg.shape).to(dtype)
    hm32 = hm.float().numpy()
    tol = 1e-5 if dtype == torch.float32 else 8e-3          # bf16 gradient storage: 2^-8 relative
    for bal in (True, False):
        x = hm.to(DEV).requires_grad_(True)
        loss = LS.DistanceLoss(balance=bal)(x, cu(tg).to(dtype), cu(tw))
        loss.backward()
        assert x.grad.dtype == dtype and x.grad.shape == x.shape
        tgq = cu(tg).to(dtype).float().cpu().numpy()
        _close(x.grad.float().cpu().numpy(), O.distance_loss_l2_grad(hm32, tgq, tw, bal), f"balance={bal}", tol)
    # fused entry (render in-kernel) against the same oracle on the f32 target
    x = hm.to(DEV).requires_grad_(True)
    loss, w = LS.DistanceLoss(balance=True).forward_fused(x, joints.to(DEV), vis.to(DEV), (4 * W, 4 * H), sigma=sig)
    loss.backward()
    np.testing.assert_allclose(loss.item(), O.distance_loss_l2(hm32, tg, tw, True), rtol=1e-5)
    _close(x.grad.float().cpu().numpy(), O.distance_loss_l2_grad(hm32, tg, tw, True), "fused", tol)


def test_training_step_through_dropin_criterion():
    """train_one_epoch's pattern (train/topdown_trainer.py:68-87): loss, _ = criterion(outputs, meta);
    loss.backward(); optimizer.step() — with a tiny conv 'model', both the explicit-target and the fused meta."""
    from litehandnet_b200 import loss as LS
    torch.manual_seed(0)
    N, K, H, W = 4, 21, 64, 64
    crit = LS.get_loss(make_cfg(K=K, hm=(W, H)))
    model = torch.nn.Conv2d(3, K, 3, padding=1).to(DEV)
    opt = torch.optim.SGD(model.parameters(), lr=0.5)
    img = torch.randn(N, 3, H, W, device=DEV)
    joints, vis = synth.hand_joints(N, K, (256, 256), seed=81)
    tg, tw = O.render_targets(joints.numpy(), vis.numpy(), (256, 256), (W, H), 2, True)
    metas = [dict(target=torch.from_numpy(tg), target_weight=torch.from_numpy(tw)),
             dict(joints_3d=joints, joints_3d_visible=vis)]
    grads = []
    for meta in metas:
        opt.zero_grad()
        loss, d = crit(model(img), meta)
        loss.backward()
        grads.append([p.grad.clone() for p in model.parameters()])
        assert isinstance(d["heatmap"], float) and np.isfinite(d["heatmap"])
    for a, b in zip(*grads):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-7), "explicit-target and fused gradients must agree"
    # reference arithmetic in plain torch on the same graph
    opt.zero_grad()
    out = model(img)
    t, w = torch.from_numpy(tg).to(DEV), torch.from_numpy(tw).to(DEV)
    ell = (out - t) ** 2 * w.unsqueeze(-1)
    pos = t > 0.5
    ref_loss = 0.1 * ell[pos].sum() / (pos.sum() + 1) + ell[~pos].sum() / ((~pos).sum() + 1)
    ref_loss.backward()
    for a, p in zip(grads[0], model.parameters()):
        assert torch.allclose(a, p.grad, rtol=1e-4, atol=1e-7)
    l0 = float(ref_loss.detach())
    for _ in range(5):
        opt.zero_grad()
        loss, _ = crit(model(img), metas[1])
        loss.backward()
        opt.step()
    assert float(loss.detach()) < l0, "five SGD steps through the drop-in criterion must reduce the loss"
