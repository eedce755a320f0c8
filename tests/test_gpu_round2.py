"""GPU: round-2 additions — lhn_metrics_finalize, the single-launch un-fused criterion, multi-tensor losses."""
import numpy as np
import pytest
import torch

from oracle import np_oracle as O
from litehandnet_b200 import _lib as L
from litehandnet_b200 import metrics as M
from litehandnet_b200 import ops, synth

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("K,N,mask_prob", [(16, 512, 0.9), (21, 300, 0.5), (7, 40, 0.9), (16, 64, 0.0)])
def test_metrics_finalize_equals_host_expressions(K, N, mask_prob):
    """lhn_metrics_finalize (device) == _report_metric's NumPy expressions on the same counters, bit for bit
    (datasets/base_dataset.py:193-261, top_down_eval.py:44-62,167-196), and == the oracle on the decoded points."""
    hm, cen = synth.blob_heatmaps(N, K, 64, 64, seed=11, device=DEV)
    c, s = synth.bbox_center_scale(N, seed=12, device=DEV)
    gt, mask, wh = synth.pck_inputs(cen, seed=13, device=DEV, mask_prob=mask_prob)
    if K == 7:
        mask[:, 3] = False                                  # a joint without any valid sample: acc = -1, left out of the mean
    acc = M.MetricAccumulator(K, device=DEV)
    r = acc.update_from_heatmaps(hm, c, s, gt, mask, wh, post_process="default")
    host = acc.compute()
    dev = acc.compute(on_device=True)
    for name in ("PCK", "AUC", "EPE"):
        assert float(host[name]) == float(dev[name]), (name, host[name], dev[name])
    full = acc.compute_device().cpu().numpy()
    cnt = acc.counters.cpu().numpy().reshape(-1, K)
    want_acc = np.array([h / v if v > 0 else -1 for h, v in zip(cnt[0], cnt[1])])
    assert np.array_equal(full[3:], want_acc)
    if mask_prob > 0:
        p64 = r["kpts"][..., :2].double().cpu().numpy()
        t = wh.max(1).values.double().cpu().numpy()
        _, pck, _ = O.keypoint_pck_accuracy(p64, gt.cpu().numpy(), mask.cpu().numpy(), 0.2, np.stack([t, t], 1))
        np.testing.assert_allclose(dev["PCK"], pck, rtol=1e-12)
        np.testing.assert_allclose(dev["AUC"], O.keypoint_auc(p64, gt.cpu().numpy(), mask.cpu().numpy(), 30), rtol=1e-12)
    else:
        assert dev["PCK"] == 0 and dev["AUC"] == 0 and dev["EPE"] == 0


def test_metrics_finalize_rejects_bad_arguments():
    cnt = torch.zeros(25 * 16, dtype=torch.int64, device=DEV)
    with pytest.raises(L.LhnError):
        ops.metrics_finalize(cnt, 15)
    with pytest.raises(L.LhnError):
        ops.metrics_finalize(cnt.int(), 16)
