"""Drop-in mirrors of the reference's loss modules (loss/heatmapLoss.py, loss/centernet_simdr_loss.py,
loss/loss.py, loss/__init__.py) — same class names, constructor arguments, call signatures, return
types and error behaviour; the arithmetic runs in the CUDA library.

Scope: L2-type DistanceLoss (balance on/off, mean/sum), JointsDistanceLoss mse,
KLDiscretLoss, SimDRLoss, TopdownHeatmapLoss, SRHandNetLoss (heatmap-only branch), get_loss.

Autograd: every loss is a torch.autograd.Function whose backward is one streaming CUDA
kernel (lhn_loss_backward / lhn_render_loss_backward / lhn_simdr_smoothl1_backward) driven by the forward's
f64 sums, so ``loss, d = criterion(outputs, meta); loss.backward()`` works as in
train/topdown_trainer.py:68-87.  Gradients flow to the network output only (targets and weights are data).
"""
import torch
from torch import nn

from . import _lib as L
from . import ops


def _as_cuda(t, device):
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(t)
    return t.to(device)          # the reference moves meta['target'] to the output's device (loss.py:97-98)


class _HeatmapLossFn(torch.autograd.Function):
    """forward: ONE launch (lhn_loss_mse_multi: streaming sums, fixed-order two-level reduction, finalisation);
    backward: lhn_loss_backward."""

    @staticmethod
    def forward(ctx, output, target, weight, mode, value, reduction):
        loss, sums, _ = ops.loss_mse_multi([output], [target], [weight], mode, value, reduction)
        sums = sums[0]
        ctx.save_for_backward(output.detach(), target, weight, sums)
        ctx.cfg = (mode, value, reduction)
        return loss[0]

    @staticmethod
    def backward(ctx, grad_out):
        output, target, weight, sums = ctx.saved_tensors
        mode, value, reduction = ctx.cfg
        g = ops.loss_backward(output, target, weight, mode, sums, value, reduction, 1.0, grad_out)
        return g, None, None, None, None, None


class _MultiHeatmapLossFn(torch.autograd.Function):
    """sum_i loss_weight[i] * DistanceLoss(outputs[i], targets[i], w[i]) in ONE launch (SRHandNetLoss's four scales,
    loss/loss.py:59-66); backward: one lhn_loss_backward per tensor, scaled by its loss weight."""

    @staticmethod
    def forward(ctx, mode, value, reduction, loss_weights, n, *tensors):
        outputs, targets, weights = tensors[:n], tensors[n:2 * n], tensors[2 * n:3 * n]
        loss, sums, _ = ops.loss_mse_multi(list(outputs), list(targets), list(weights), mode, value, reduction, loss_weights)
        ctx.save_for_backward(*[o.detach() for o in outputs], *targets, *weights, sums)
        ctx.cfg = (mode, value, reduction, [float(x) for x in loss_weights], n)
        return loss[0]

    @staticmethod
    def backward(ctx, grad_out):
        mode, value, reduction, lw, n = ctx.cfg
        saved = ctx.saved_tensors
        outputs, targets, weights, sums = saved[:n], saved[n:2 * n], saved[2 * n:3 * n], saved[3 * n]
        grads = []
        for i in range(n):
            if ctx.needs_input_grad[5 + i]:
                grads.append(ops.loss_backward(outputs[i], targets[i], weights[i], mode, sums[i], value, reduction, lw[i], grad_out))
            else:
                grads.append(None)
        return (None, None, None, None, None, *grads, *([None] * (2 * n)))


class _FusedHeatmapLossFn(torch.autograd.Function):
    """The fused entry: target rendered in-kernel from the joints, one launch forward
    (lhn_fused_render_loss_decode without a refinement), lhn_render_loss_backward backward."""

    @staticmethod
    def forward(ctx, output, joints, vis, render, reduction):
        r = ops.fused_render_loss_decode(output.detach(), L.MASK_NEG1, L.REFINE_NONE, L.XFORM_NONE, None, None,
                                         render, joints, vis, want_idx=False, reduction=reduction)
        ctx.save_for_backward(output.detach(), joints, vis, r["sums"])
        ctx.cfg = (render, reduction)
        ctx.mark_non_differentiable(r["weight"])
        return r["loss"][0], r["weight"]

    @staticmethod
    def backward(ctx, grad_out, _grad_w):
        output, joints, vis, sums = ctx.saved_tensors
        render, reduction = ctx.cfg
        return ops.render_loss_backward(output, joints, vis, render, sums, reduction, 1.0, grad_out), None, None, None, None


class _SimDRLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, out_x, out_y, tgt_x, tgt_y, weight):
        ctx.save_for_backward(out_x.detach(), out_y.detach(), tgt_x, tgt_y, weight)
        return ops.simdr_smoothl1(out_x.detach(), out_y.detach(), tgt_x, tgt_y, weight)[0]

    @staticmethod
    def backward(ctx, grad_out):
        out_x, out_y, tgt_x, tgt_y, weight = ctx.saved_tensors
        gx, gy = ops.simdr_smoothl1_backward(out_x, out_y, tgt_x, tgt_y, weight, 1.0, grad_out)
        return gx, gy.to(out_y.dtype), None, None, None


class _SimDRHeadsLossFn(torch.autograd.Function):
    """SimDRLoss.forward with the two linear heads fused into the loss kernel (lhn_simdr_heads_loss).  When a
    gradient is needed the kernel's epilogue also stores g = d SmoothL1 / d pred = clamp(pred - target, -1, 1); the
    backward scales its rows by grad * mean_b(w[b, j]) / (B * L * K) and takes the three plain GEMMs of a linear layer
    (d heatmap = G W, d W = G^T A, d b = column sums of G) from the library."""

    @staticmethod
    def forward(ctx, heatmap, wx, bx, wy, by, tgt_x, tgt_y, weight, w_split):
        need = any(ctx.needs_input_grad[:5])
        bias = torch.cat([bx.detach(), by.detach()])
        loss, dpred, _ = ops.simdr_heads_loss(heatmap.detach(), w_split, bias, tgt_x, tgt_y, weight, want_dpred=need)
        if need:
            ctx.save_for_backward(heatmap.detach(), wx.detach(), wy.detach(), weight, dpred)
            ctx.Lx, ctx.Ly = tgt_x.shape[-1], tgt_y.shape[-1]
        return loss[0]

    @staticmethod
    def backward(ctx, grad_out):
        heatmap, wx, wy, weight, g = ctx.saved_tensors
        B, K = heatmap.shape[:2]
        Lx, Ly = ctx.Lx, ctx.Ly
        A = heatmap.reshape(B * K, -1).float()
        mw = weight.reshape(B, K).float().mean(0)                          # mean_b w[b, j]
        row = (grad_out.float() * mw / float(K * B)).repeat(B)              # coefficient of row b*K + j, without 1/L
        g = g * row[:, None]
        gx, gy = g[:, :Lx] * (1.0 / Lx), g[:, Lx:] * (1.0 / Ly)
        d_hm = d_wx = d_bx = d_wy = d_by = None
        if ctx.needs_input_grad[0]:
            d_hm = (gx @ wx.float() + gy @ wy.float()).reshape(heatmap.shape).to(heatmap.dtype)
        if ctx.needs_input_grad[1]:
            d_wx = gx.t() @ A
        if ctx.needs_input_grad[2]:
            d_bx = gx.sum(0)
        if ctx.needs_input_grad[3]:
            d_wy = gy.t() @ A
        if ctx.needs_input_grad[4]:
            d_by = gy.sum(0)
        return d_hm, d_wx, d_bx, d_wy, d_by, None, None, None, None


class DistanceLoss(nn.Module):
    """loss/heatmapLoss.py:228-265.  Only loss_type 'L2' is on the path (every config uses it)."""

    def __init__(self, loss_type='L2', reduction='mean', balance=True, value=0.5):
        super().__init__()
        assert reduction in ['mean', 'sum', None], f"Error: {reduction=}"
        if loss_type.lower() != 'l2':
            raise NotImplementedError("only DistanceLoss(loss_type='L2') is on the B200 hot path "
                                      "(SURVEY.md §2 row 1); L1/SmoothL1 are unused by every config")
        if reduction is None:
            raise NotImplementedError("reduction=None (unreduced loss map) is not on the hot path")
        self.reduction = reduction
        self.value = value
        self.balance = balance

    @property
    def _mode(self):
        return L.LOSS_DISTANCE_BALANCE if self.balance else L.LOSS_DISTANCE

    def forward(self, output, target, target_weight):
        """output/target [N,K,H,W] or [N,S,K,H,W]; target_weight [N,K,1] / [N,S,K,1] -> 0-dim f32."""
        L.require_cuda(output, "output")
        target = _as_cuda(target, output.device)
        target_weight = _as_cuda(target_weight, output.device)
        return _HeatmapLossFn.apply(output, target, target_weight, self._mode, self.value, self.reduction)

    def forward_fused(self, output, joints_3d, joints_3d_visible, image_size, sigma=2,
                      unbiased_encoding=True):
        """Additive fused entry: render the target in-kernel from the joints instead of
        reading a target tensor.  Returns (loss, target_weight [.., 1])."""
        L.require_cuda(output, "output")
        render = dict(loss_mode=self._mode, image_size=image_size, sigma=sigma, unbiased=unbiased_encoding,
                      pos_value=self.value)
        loss, weight = _FusedHeatmapLossFn.apply(output, _as_cuda(joints_3d, output.device),
                                                 _as_cuda(joints_3d_visible, output.device), render, self.reduction)
        return loss, weight.unsqueeze(-1)


class JointsDistanceLoss(nn.Module):
    """loss/heatmapLoss.py:175-225 (HRNet per-joint 0.5*MSE; the weight multiplies both operands)."""

    def __init__(self, use_target_weight=True, loss_type='mse'):
        super().__init__()
        assert loss_type.lower() in ['mse', 'mae', 'smoothl1']
        if loss_type.lower() != 'mse':
            raise NotImplementedError("only JointsDistanceLoss(loss_type='mse') is on the hot path")
        self.use_target_weight = use_target_weight

    def forward(self, output, target, target_weight=None):
        L.require_cuda(output, "output")
        if self.use_target_weight:
            if target_weight is None:
                raise NameError
            w = _as_cuda(target_weight, output.device)
        else:
            w = torch.ones(output.shape[:2], device=output.device)
        return _HeatmapLossFn.apply(output, _as_cuda(target, output.device), w, L.LOSS_JOINTS_MSE, 0.5, "mean")


class KLDiscretLoss(nn.Module):
    """loss/centernet_simdr_loss.py:6-39 — despite the name: per-joint SmoothL1 mean, times the
    batch-mean of the weight."""

    def forward(self, output_x, output_y, target_x, target_y, target_weight):
        L.require_cuda(output_x, "output_x")
        dev = output_x.device
        return _SimDRLossFn.apply(output_x, _as_cuda(output_y, dev), _as_cuda(target_x, dev),
                                  _as_cuda(target_y, dev), _as_cuda(target_weight, dev))


class SimDRLoss(nn.Module):
    """loss/centernet_simdr_loss.py:42-69.  Same parameters as the reference (two nn.Linear heads, so checkpoints
    load unchanged); the forward is ONE tcgen05 kernel that multiplies the flattened heatmaps with both heads and
    takes KLDiscretLoss in its epilogue (lhn_simdr_heads_loss: the [B*K, k*W_img] predictions never reach HBM).
    ``fused=False`` keeps the reference's structure (cuBLAS heads, then lhn_simdr_smoothl1); shapes outside the
    fused kernel's envelope (H*W not a multiple of 64, non-f32 heatmaps) take that route too."""

    def __init__(self, cfg=None):
        super().__init__()
        image_size = cfg.DATASET.image_size
        heatmap_size = cfg.DATASET.heatmap_size
        k = cfg.PIPELINE.simdr_split_ratio
        self.simdr_width = int(k * image_size[0])
        self.simdr_height = int(k * image_size[1])
        in_features = int(heatmap_size[0] * heatmap_size[1])
        self.x_shared_decoder = nn.Linear(in_features, self.simdr_width)
        self.y_shared_decoder = nn.Linear(in_features, self.simdr_height)
        self.loss = KLDiscretLoss()
        self.fused = True
        self._split = None                   # (key, (w_hi, w_lo)): bf16 pair of cat([Wx, Wy]), rebuilt when the weights change

    def _weight_split(self):
        wx, wy = self.x_shared_decoder.weight, self.y_shared_decoder.weight
        key = (wx.data_ptr(), wx._version, wy.data_ptr(), wy._version, wx.device)
        if self._split is None or self._split[0] != key:
            self._split = (key, ops.split_bf16(torch.cat([wx.detach(), wy.detach()]).float().contiguous()))
        return self._split[1]

    def _fusable(self, heatmap):
        n = self.simdr_width + self.simdr_height
        return (self.fused and heatmap.is_cuda and heatmap.dtype == torch.float32 and heatmap.dim() == 4 and
                heatmap[0, 0].numel() % 64 == 0 and n % 64 == 0 and self.simdr_width % 4 == 0 and self.simdr_height % 4 == 0
                and self.x_shared_decoder.weight.dtype == torch.float32)

    def forward(self, heatmap, simdr_x, simdr_y, target_weight):
        if self._fusable(heatmap):
            dev = heatmap.device
            return _SimDRHeadsLossFn.apply(heatmap, self.x_shared_decoder.weight, self.x_shared_decoder.bias,
                                           self.y_shared_decoder.weight, self.y_shared_decoder.bias,
                                           _as_cuda(simdr_x, dev), _as_cuda(simdr_y, dev), _as_cuda(target_weight, dev),
                                           self._weight_split())
        pred_x = self.x_shared_decoder(heatmap.flatten(start_dim=2))
        pred_y = self.y_shared_decoder(heatmap.flatten(start_dim=2))
        return self.loss(pred_x, pred_y, simdr_x, simdr_y, target_weight)


class TopdownHeatmapLoss(nn.Module):
    """loss/loss.py:69-114.  criterion(output, meta) -> (loss, {name: float}).

    Fused entry (additive): when ``meta`` has no 'target' but has 'joints_3d' / 'joints_3d_visible',
    the target is rendered in-kernel (cfg.PIPELINE sigma / unbiased_encoding, cfg.DATASET.image_size)."""

    def __init__(self, cfg):
        super().__init__()
        loss_type = cfg.LOSS.get('dl_type', 'L2')
        balance = cfg.MODEL.name != 'atthandnet'
        self.heatmap_loss = DistanceLoss(loss_type=loss_type, reduction='mean', balance=balance)
        if cfg.PIPELINE.simdr_split_ratio > 0:
            self.simdr_loss = SimDRLoss(cfg)
        else:
            self.simdr_loss = None
        self.loss_weight = cfg.LOSS.loss_weight
        self.auto_weight = cfg.LOSS.auto_weight
        if self.auto_weight:
            params = torch.ones(len(self.loss_weight), requires_grad=True)
            self.p = nn.Parameter(params, requires_grad=True)
        self._image_size = cfg.DATASET.get('image_size', None)
        self._sigma = cfg.PIPELINE.get('sigma', 2)
        self._unbiased = cfg.PIPELINE.get('unbiased_encoding', True)

    def forward(self, output, meta):
        loss_dict = {}
        device = output.device
        if 'target' in meta:
            target = meta['target'].to(device)
            target_weight = meta['target_weight'].to(device)
            hl = self.heatmap_loss(output, target, target_weight)
        else:
            # the in-kernel render multiplies no per-joint weights into target_weight (generateTarget.py:156-157);
            # every hand dataset forces them off (freihand_dataset.py:62) — refuse rather than silently ignore
            ann = meta.get('ann_info', {}) if hasattr(meta, 'get') else {}
            if meta.get('use_different_joint_weights', False) or (hasattr(ann, 'get') and ann.get('use_different_joint_weights', False)):
                raise NotImplementedError("use_different_joint_weights: render the targets with TopDownGenerateTarget "
                                          "and pass meta['target'] / meta['target_weight'] instead of the fused entry")
            hl, target_weight = self.heatmap_loss.forward_fused(
                output, meta['joints_3d'], meta['joints_3d_visible'], self._image_size, self._sigma,
                self._unbiased)
        loss_dict['heatmap'] = self.loss_weight[0] * hl
        if self.simdr_loss is not None:
            simdr_x = meta['simdr_x'].to(device)
            simdr_y = meta['simdr_y'].to(device)
            loss_dict['simdr'] = self.loss_weight[1] * self.simdr_loss(output, simdr_x, simdr_y, target_weight)
        loss = 0
        for k, v in loss_dict.items():
            loss += v
            loss_dict[k] = v.item()          # .item() syncs: the reference contract
        return loss, loss_dict


class SRHandNetLoss(nn.Module):
    """loss/loss.py:7-66, heatmap-only branch (_forward_only_heatmap): sum_i loss_weight[i] *
    DistanceLoss(outputs[i], targets[i], w[i]) over the 4 scales.  The region-map branch
    (pred_bbox with 24 channels) is out of scope."""

    def __init__(self, cfg):
        super().__init__()
        out_c = cfg.MODEL.get('output_channel', 24)
        pred_bbox = cfg.MODEL.get('pred_bbox', False)
        self.mse_loss = DistanceLoss(loss_type='L2', reduction='mean')
        if pred_bbox and out_c == 24:
            raise NotImplementedError("SRHandNetLoss region-map branch is out of the hot-path scope")
        self.smoothl1_loss = None
        self.num_out = 4
        self.loss_weight = cfg.LOSS.loss_weight
        assert len(self.loss_weight) == self.num_out

    def forward(self, outputs, meta):
        target = meta['target']
        target_weight = meta['target_weight']
        return self._forward_only_heatmap(outputs, target, target_weight)

    def _forward_only_heatmap(self, outputs, targets, target_weight):
        device = outputs[-1].device
        ws = [_as_cuda(target_weight[i] if isinstance(target_weight, (list, tuple)) else target_weight, device)
              for i in range(self.num_out)]
        tg = [_as_cuda(targets[i], device) for i in range(self.num_out)]
        outs = [outputs[i] for i in range(self.num_out)]
        for o in outs:
            L.require_cuda(o, "output")
        # the four scales in ONE launch (it used to be 4 x 3)
        loss = _MultiHeatmapLossFn.apply(self.mse_loss._mode, self.mse_loss.value, self.mse_loss.reduction,
                                         [float(x) for x in self.loss_weight], self.num_out, *outs, *tg, *ws)
        return loss, dict(kpt_loss=loss.item())


srhandnetloss = SRHandNetLoss
topdownheatmaploss = TopdownHeatmapLoss


def get_loss(cfg):
    """loss/__init__.py:18-19."""
    return {"srhandnetloss": SRHandNetLoss, "topdownheatmaploss": TopdownHeatmapLoss}[cfg.LOSS.type.lower()](cfg)
