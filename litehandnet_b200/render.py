"""Drop-in mirrors of the target renderers (per-sample dict-in/dict-out transforms of the dataset
pipeline) plus the batched GPU entry points.

  datasets/data_pipeline/generateTarget.py:33   TopDownGenerateTarget (MSRA unbiased + integer-centre, UDP
                                                GaussianHeatmap, sigma lists -> stacked targets)
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
