"""Drop-in mirrors of the target renderers (per-sample dict-in/dict-out transforms of the dataset
pipeline) plus the batched GPU entry points.

  datasets/data_pipeline/generateTarget.py:33   TopDownGenerateTarget (MSRA unbiased + integer-centre, UDP
                                                GaussianHeatmap, sigma lists -> stacked targets)
  datasets/data_pipeline/generateTarget.py:303  SRHandNetGenerateTarget (one target per scale, optional region map:
                                                centre Gaussian + width / height ratio planes)
  datasets/data_pipeline/generate_simder.py:3   GenerateSimDR

The fused training path never materialises targets (see loss.TopdownHeatmapLoss / fused.py); these
classes exist so a pipeline that wants the tensors still gets them, rendered on the GPU.
"""
import numpy as np
import torch

from . import ops
from .decode import _device, _np


def render_targets(joints_3d, joints_3d_visible, image_size, heatmap_size, sigma=2, unbiased_encoding=True):
    """Batched render: joints_3d [B,K,3], joints_3d_visible [B,K,3] (CUDA tensors) ->
    target [B,(S,)K,H,W] f32, target_weight [B,(S,)K,1] f32 on the GPU."""
    return ops.render_targets(joints_3d, joints_3d_visible, image_size, heatmap_size, sigma, unbiased_encoding)


def render_simdr(joints_3d, joints_3d_visible, image_size, k=2, sigma=2):
    """Batched SimDR render -> simdr_x [B,K,Wi*k], simdr_y [B,K,Hi*k] f32 on the GPU."""
    return ops.render_simdr(joints_3d, joints_3d_visible, image_size, k, sigma)


class TopDownGenerateTarget:
    def __init__(self, sigma=2, kernel=(11, 11), target_type='GaussianHeatmap', encoding='MSRA',
                 unbiased_encoding=False):
        self.sigma = sigma
        self.unbiased_encoding = unbiased_encoding
        self.kernel = kernel
        self.target_type = target_type
        self.encoding = encoding

    def __call__(self, results):
        """results['joints_3d'] [K,3], ['joints_3d_visible'] [K,3], ['ann_info'] -> adds 'target'
        [K,H,W] (or [S,K,H,W]) and 'target_weight' [K,1] (or [S,K,1]) as NumPy arrays."""
        assert self.encoding in ['MSRA', 'UDP']
        if self.encoding == 'UDP' and self.target_type.lower() != 'gaussianheatmap':
            raise NotImplementedError("UDP 'CombinedTarget' (offset maps) is not produced by the reference either "
                                      "(generateTarget.py:162-243 renders the GaussianHeatmap only)")
        cfg = results['ann_info']
        if cfg.get('use_different_joint_weights', False):
            raise NotImplementedError("joint_weights are forced off by every hand dataset (freihand_dataset.py:62)")
        dev = _device()
        j = torch.as_tensor(np.asarray(results['joints_3d'], dtype=np.float32)).to(dev)[None]
        v = torch.as_tensor(np.asarray(results['joints_3d_visible'], dtype=np.float32)).to(dev)[None]
        sigma = list(self.sigma) if isinstance(self.sigma, (list, tuple)) else self.sigma
        t, w = ops.render_targets(j, v, tuple(cfg['image_size']), tuple(cfg['heatmap_size']), sigma,
                                  'udp' if self.encoding == 'UDP' else self.unbiased_encoding)
        results['target'] = _np(t[0])
        results['target_weight'] = _np(w[0])
        return results


def _region_window(bbox, image_size, heatmap_size):
    """generateTarget.py:334-365, the per-sample scalars (host logic on 4 numbers, in the array's own dtype as
    NumPy evaluates it): box centre, clipped width / height ratios and the 5x5 window around the centre, which
    the reference uses as a Python slice (a negative upper bound counts from the end).
    -> (centre [2], (x1, x2, y1, y2), (gamma_x, gamma_y))"""
    bbox = np.asarray(bbox)
    image_size = np.asarray(image_size)
    heatmap_size = np.asarray(heatmap_size)
    center = bbox[:2] + bbox[2:] / 2
    gamma_x, gamma_y = np.clip(bbox[2:] / image_size, 0, 1)
    x, y = center * (heatmap_size / image_size)
    W, H = int(heatmap_size[0]), int(heatmap_size[1])
    x1, x2, _ = slice(max(0, int(x - 2)), min(int(x + 3), W)).indices(W)
    y1, y2, _ = slice(max(0, int(y - 2)), min(int(y + 3), H)).indices(H)
    return center, (x1, max(x2, x1), y1, max(y2, y1)), (np.float32(gamma_x), np.float32(gamma_y))


def render_srhandnet_targets(joints_3d, joints_3d_visible, bbox, image_size, heatmap_sizes, sigmas, pred_bbox=True,
                             unbiased_encoding=False):
    """Batched SRHandNetGenerateTarget (generateTarget.py:369-426): joints_3d / joints_3d_visible [B,K,3] CUDA
    tensors, bbox [B,4] (lx, ly, w, h; host array) -> lists over the scales of target [B,K(+3),H,W] and
    target_weight [B,K(+3),1] on the GPU.  Channels K..K+2 of each scale are the region map: a Gaussian at the box
    centre (lhn_render_targets) and the width / height ratio planes (lhn_render_region_wh)."""
    dev = joints_3d.device
    B, K = joints_3d.shape[:2]
    bbox = np.asarray(bbox)
    targets, weights = [], []
    for hs, sg in zip(heatmap_sizes, sigmas):
        W, H = int(hs[0]), int(hs[1])
        t, w = ops.render_targets(joints_3d, joints_3d_visible, tuple(image_size), (W, H), sg, unbiased_encoding)
        if pred_bbox:
            out = torch.empty((B, K + 3, H, W), dtype=torch.float32, device=dev)
            out[:, :K] = t
            cj = np.ones((B, 1, 3), np.float32)
            rect = np.zeros((B, 4), np.int32)
            gamma = np.zeros((B, 2), np.float32)
            for b in range(B):
                c, rect[b], gamma[b] = _region_window(bbox[b], image_size, (W, H))
                cj[b, 0, :2] = c
            cjt = torch.from_numpy(cj).to(dev)
            ct, _ = ops.render_targets(cjt, torch.ones_like(cjt), tuple(image_size), (W, H), sg, unbiased_encoding)
            out[:, K:K + 1] = ct
            ops.render_region_wh(torch.from_numpy(rect).to(dev), torch.from_numpy(gamma).to(dev), H, W, out=out[:, K + 1:K + 3])
            t = out
            w = torch.cat([w, torch.ones((B, 3, 1), dtype=torch.float32, device=dev)], dim=1)
        targets.append(t)
        weights.append(w)
    return targets, weights


class SRHandNetGenerateTarget(TopDownGenerateTarget):
    """datasets/data_pipeline/generateTarget.py:303-426: results['ann_info']['heatmap_size'] is a list of sizes
    (e.g. [[16,16],[16,16],[32,32],[64,64]]), one sigma per scale; adds results['target'] / ['target_weight'] as
    LISTS of NumPy arrays [K(+3),H,W] / [K(+3),1]."""

    def __init__(self, pred_bbox=True, sigma=[2, 2, 2, 2], kernel=(11, 11), target_type='GaussianHeatmap',
                 encoding='MSRA', unbiased_encoding=False):
        super().__init__(sigma=sigma, kernel=kernel, target_type=target_type, encoding=encoding,
                         unbiased_encoding=unbiased_encoding)
        self.pred_bbox = pred_bbox

    def __call__(self, results):
        assert self.encoding in ['MSRA', 'UDP']
        cfg = results['ann_info']
        heatmap_size = cfg['heatmap_size']
        assert len(heatmap_size) == len(self.sigma)
        dev = _device()
        j = torch.as_tensor(np.asarray(results['joints_3d'], dtype=np.float32)).to(dev)[None]
        v = torch.as_tensor(np.asarray(results['joints_3d_visible'], dtype=np.float32)).to(dev)[None]
        bbox = np.asarray(results['bbox'])[None] if self.pred_bbox else None
        t, w = render_srhandnet_targets(j, v, bbox, cfg['image_size'], heatmap_size, self.sigma, self.pred_bbox,
                                        'udp' if self.encoding == 'UDP' else self.unbiased_encoding)
        results['target'] = [_np(x[0]) for x in t]
        results['target_weight'] = [_np(x[0]) for x in w]
        return results


class GenerateSimDR:
    def __init__(self, sigma=2, k=2):
        self.sigma = sigma
        self.k = int(k)
        self.with_simdr = k > 0 and not isinstance(sigma, (list, tuple))

    def __call__(self, results):
        if self.with_simdr:
            dev = _device()
            j = torch.as_tensor(np.asarray(results['joints_3d'], dtype=np.float32)).to(dev)[None]
            v = torch.as_tensor(np.asarray(results['joints_3d_visible'], dtype=np.float32)).to(dev)[None]
            sx, sy = ops.render_simdr(j, v, tuple(results['ann_info']['image_size']), self.k, self.sigma)
            results['simdr_x'] = _np(sx[0])
            results['simdr_y'] = _np(sy[0])
        return results
