"""The headline fused path (BASELINE config 2): Gaussian target render + target-weight masked MSE
loss + flip-test average + argmax + DARK refinement + affine back-transform, with every heatmap
plane read from HBM exactly once.

It replaces this chain of reference calls:
  datasets/data_pipeline/generateTarget.py:245  TopDownGenerateTarget.__call__      (per sample, CPU)
  loss/heatmapLoss.py:242                       DistanceLoss.forward                (~9 eager kernels)
  utils/transforms.py:78                        flip_back + (a + b) * 0.5
  utils/post_processing/evaluation/top_down_eval.py:375  keypoints_from_heatmaps('unbiased')
"""
import torch

from . import _lib as L
from . import ops


def flip_index_from_pairs(num_joints, flip_pairs, device):
    """flip_back's channel swap (utils/transforms.py:86-90) as a gather index."""
    idx = list(range(num_joints))
    for a, b in flip_pairs or ():
        idx[a], idx[b] = idx[b], idx[a]
    return torch.tensor(idx, dtype=torch.int32, device=device)


class FusedHeatmapStep:
    """Reusable launcher for one configuration; a call is ONE kernel launch (decode + render + loss sums +
    fixed-order reduction + finalisation, lhn_fused_render_loss_decode)."""

    def __init__(self, image_size=(256, 256), sigma=2, unbiased_encoding=True, balance=True,
                 post_process="unbiased", kernel=11, flip_pairs=(), loss_weight=1.0, pos_value=0.5,
                 loss_type="distance"):
        self.image_size = tuple(image_size)
        self.sigma = sigma
        self.unbiased = bool(unbiased_encoding)
        if loss_type == "joints_mse":
            self.loss_mode = L.LOSS_JOINTS_MSE
        elif loss_type is None:
            self.loss_mode = L.LOSS_NONE
        else:
            self.loss_mode = L.LOSS_DISTANCE_BALANCE if balance else L.LOSS_DISTANCE
        self.refine = {"unbiased": L.REFINE_DARK, "default": L.REFINE_SIGN, None: L.REFINE_NONE}[post_process]
        self.kernel = int(kernel)
        self.flip_pairs = tuple(flip_pairs)
        self.loss_weight = float(loss_weight)
        self.pos_value = float(pos_value)
        self._flip_index = {}

    def __call__(self, hm, joints_3d, joints_3d_visible, center, scale, hm_flip=None):
        dev = hm.device
        K = hm.shape[-3]
        fi = None
        if hm_flip is not None and self.flip_pairs:
            key = (K, dev)
            if key not in self._flip_index:
                self._flip_index[key] = flip_index_from_pairs(K, self.flip_pairs, dev)
            fi = self._flip_index[key]
        render = None
        if self.loss_mode != L.LOSS_NONE:
            render = dict(loss_mode=self.loss_mode, image_size=self.image_size, sigma=self.sigma,
                          unbiased=self.unbiased, pos_value=self.pos_value)
        if render is None:
            r = ops.decode_heatmap(hm, L.MASK_NEG1, self.refine, L.XFORM_CENTER_SCALE, center, scale,
                                   hm_flip=hm_flip, flip_index=fi, blur_ksize=self.kernel)
            return dict(preds=r["kpts"], hm_preds=r["hm_kpts"], idx=r["idx"], maxvals=r["kpts"][..., 2:])
        # one launch: decode + render + loss sums + fixed-order reduction + finalisation
        r = ops.fused_render_loss_decode(hm, L.MASK_NEG1, self.refine, L.XFORM_CENTER_SCALE, center, scale,
                                         render, joints_3d, joints_3d_visible, hm_flip=hm_flip, flip_index=fi,
                                         blur_ksize=self.kernel, loss_scale=self.loss_weight)
        return dict(preds=r["kpts"], hm_preds=r["hm_kpts"], idx=r["idx"], maxvals=r["kpts"][..., 2:],
                    loss_sums=r["sums"], loss=r["loss"][0], target_weight=r["weight"].unsqueeze(-1))


def fused_render_loss_decode(hm, joints_3d, joints_3d_visible, center, scale, hm_flip=None,
                             image_size=(256, 256), sigma=2, unbiased_encoding=True, balance=True,
                             post_process="unbiased", kernel=11, flip_pairs=(), loss_weight=1.0):
    """hm [B,K,H,W] (f32/bf16/f16, CUDA), joints_3d [B,K,3] image pixels, joints_3d_visible [B,K,3],
    center/scale [B,2] (scale = bbox/200), hm_flip = network output on the mirrored image or None.

    Returns dict(loss 0-dim f32, preds [B,K,3] (X,Y,score) in image coords, hm_preds [B,K,3] in
    heatmap coords, idx [B,K] int32 argmax, maxvals [B,K,1], target_weight [B,K,1], loss_sums f64[4]).
    """
    step = FusedHeatmapStep(image_size, sigma, unbiased_encoding, balance, post_process, kernel,
                            flip_pairs, loss_weight)
    return step(hm, joints_3d, joints_3d_visible, center, scale, hm_flip)


class HostPipeline:
    """End-to-end entry for HOST buffers (the call a reference user makes today hands CPU arrays to
    the loss/decoder): the batch is cut into chunks; chunk i+1's host->device copy (pinned memory, copy
    stream) overlaps chunk i's fused kernel (compute stream); the per-plane loss partials of all chunks
    land in one buffer so the balanced loss still uses batch-global counts; the [B,K,3] predictions and
    the loss scalar are copied back device->host at the end.

    h2d_bytes / d2h_bytes report what one call moves.  Inputs may be pinned host tensors (fastest: truly
    asynchronous copies), pageable host tensors or NumPy arrays (what a reference caller holds; the driver then
    stages each copy itself).  want_idx adds the int32 argmax indices to the device->host read."""

    def __init__(self, step, B, K, H, W, dtype=torch.float32, flip=True, chunks=8, device="cuda", want_idx=False,
                 metrics=None):
        """metrics: None or dict(pck_thr=0.2, auc_nor=30.0, auc_steps=20) — decode-only steps then also accumulate
        the PCK/AUC/EPE counters per chunk (lhn_decode_heatmap_pck; __call__ takes gt, mask, bbox_wh) and the call
        returns the finalised (PCK, AUC, EPE) from lhn_metrics_finalize in `self.h_metrics` (BASELINE config 4)."""
        self.step, self.B, self.K, self.H, self.W = step, B, K, H, W
        self.metrics = None
        if metrics is not None:
            if step.loss_mode != L.LOSS_NONE or flip:
                raise L.LhnError("HostPipeline: fused metrics go with a decode-only step without a flip plane")
            self.metrics = dict(pck_thr=0.2, auc_nor=30.0, auc_steps=20)
            self.metrics.update(metrics)
            T = int(self.metrics["auc_steps"])
            dev_ = torch.device(device)
            self.d_counters = torch.zeros((T + 5) * K, dtype=torch.int64, device=dev_)
            self.d_gt = torch.empty((B, K, 2), device=dev_)
            self.d_mask = torch.empty((B, K), dtype=torch.uint8, device=dev_)
            self.d_wh = torch.empty((B, 2), device=dev_)
            self.d_metrics = torch.empty(3 + K, dtype=torch.float64, device=dev_)
            self.h_metrics = torch.empty(3 + K, dtype=torch.float64, pin_memory=True)
            self.h_counters = torch.empty((T + 5) * K, dtype=torch.int64, pin_memory=True)
        self.flip = flip
        self.dev = torch.device(device)
        chunks = max(1, min(chunks, B))
        edges = [round(i * B / chunks) for i in range(chunks + 1)]
        self.bounds = [(a, b) for a, b in zip(edges[:-1], edges[1:]) if b > a]
        cb = max(b - a for a, b in self.bounds)
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        # double-buffered device staging for the heatmap chunks
        self.d_hm = [torch.empty((cb, K, H, W), dtype=dtype, device=self.dev) for _ in range(2)]
        self.d_hf = [torch.empty((cb, K, H, W), dtype=dtype, device=self.dev) for _ in range(2)] if flip else None
        self.d_joints = torch.empty((B, K, 3), device=self.dev)
        self.d_vis = torch.empty((B, K, 3), device=self.dev)
        self.d_center = torch.empty((B, 2), device=self.dev)
        self.d_scale = torch.empty((B, 2), device=self.dev)
        self.d_partials = torch.empty((B * K, 4), dtype=torch.float64, device=self.dev)
        self.d_preds = torch.empty((B, K, 3), device=self.dev)
        self.d_hmk = torch.empty((B, K, 3), device=self.dev)
        self.d_weight = torch.empty((B, K), device=self.dev)
        self.d_loss = torch.empty(1, device=self.dev)
        self.flip_index = flip_index_from_pairs(K, step.flip_pairs, self.dev) if (flip and step.flip_pairs) else None
        self.want_idx = bool(want_idx)
        self.d_idx = torch.empty((B, K), dtype=torch.int32, device=self.dev) if want_idx else None
        self.h_idx = torch.empty((B, K), dtype=torch.int32, pin_memory=True) if want_idx else None
        self.has_loss = step.loss_mode != L.LOSS_NONE
        self.h_preds = torch.empty((B, K, 3), pin_memory=True)
        self.h_loss = torch.empty(1, pin_memory=True)
        self.copied = [torch.cuda.Event() for _ in self.bounds]
        self.freed = [torch.cuda.Event() for _ in range(2)]
        esz = torch.empty((), dtype=dtype).element_size()
        self.h2d_bytes = B * K * H * W * esz * (2 if flip else 1) + B * (K * 3 * 4 * 2 + 16)
        if not self.has_loss:
            self.h2d_bytes = B * K * H * W * esz * (2 if flip else 1) + B * 16
        if self.metrics is not None:
            self.h2d_bytes += B * K * 8 + B * K + B * 8
        self.d2h_bytes = B * K * 3 * 4 + (4 if self.has_loss else 0) + (B * K * 4 if want_idx else 0)
        if self.metrics is not None:
            self.d2h_bytes += (3 + K) * 8 + self.d_counters.numel() * 8
        self.launches = 0

    def __call__(self, hm, hm_flip, joints, vis, center, scale, gt=None, mask=None, bbox_wh=None):
        """All arguments are HOST tensors / arrays (hm_flip None when flip=False; joints and vis may be None for a
        decode-only step).  Returns (h_preds [B,K,3], h_loss [1] or None) host tensors — and h_idx in `self.h_idx`
        when want_idx — valid after the call returns (it synchronises on the final device->host copy)."""
        step = self.step

        def host(t):
            return t if (t is None or isinstance(t, torch.Tensor)) else torch.as_tensor(t)

        hm, hm_flip, joints, vis, center, scale = (host(t) for t in (hm, hm_flip, joints, vis, center, scale))
        gt, mask, bbox_wh = host(gt), host(mask), host(bbox_wh)
        if mask is not None and mask.dtype == torch.bool:
            mask = mask.view(torch.uint8)
        if hm.shape[0] != self.B or (self.flip and (hm_flip is None or hm_flip.shape != hm.shape)):
            raise L.LhnError("HostPipeline: heatmap batch does not match the bound shape")
        comp = torch.cuda.current_stream(self.dev)
        cs = self.copy_stream
        cs.wait_stream(comp)
        with torch.cuda.stream(cs):
            if self.has_loss:
                self.d_joints.copy_(joints, non_blocking=True)
                self.d_vis.copy_(vis, non_blocking=True)
            self.d_center.copy_(center, non_blocking=True)
            self.d_scale.copy_(scale, non_blocking=True)
            if self.metrics is not None:
                self.d_gt.copy_(gt, non_blocking=True)
                self.d_mask.copy_(mask, non_blocking=True)
                self.d_wh.copy_(bbox_wh, non_blocking=True)
        if self.metrics is not None:
            self.d_counters.zero_()
        render = dict(loss_mode=step.loss_mode, image_size=step.image_size, sigma=step.sigma,
                      unbiased=step.unbiased, pos_value=step.pos_value) if self.has_loss else None
        self.launches = 1 if self.metrics is not None else 0
        for i, (a, b) in enumerate(self.bounds):
            slot = i & 1
            n = b - a
            with torch.cuda.stream(cs):
                if i >= 2:
                    cs.wait_event(self.freed[slot])       # the kernel that read this slot is done
                self.d_hm[slot][:n].copy_(hm[a:b], non_blocking=True)
                if self.flip:
                    self.d_hf[slot][:n].copy_(hm_flip[a:b], non_blocking=True)
                self.copied[i].record(cs)
            comp.wait_event(self.copied[i])
            out = dict(kpts=self.d_preds[a:b], hm_kpts=self.d_hmk[a:b])
            if self.has_loss:
                out.update(partials=self.d_partials[a * self.K:b * self.K], weight=self.d_weight[a:b])
            if self.want_idx:
                out["idx"] = self.d_idx[a:b]
            if self.metrics is not None:
                m = self.metrics
                ops.decode_heatmap_pck(self.d_hm[slot][:n], L.MASK_NEG1, step.refine, self.d_center[a:b], self.d_scale[a:b],
                                       self.d_gt[a:b], self.d_mask[a:b], self.d_wh[a:b], self.d_counters, m["pck_thr"],
                                       m["auc_nor"], m["auc_steps"], step.kernel, out=out)
                self.freed[slot].record(comp)
                self.launches += 1
                continue
            ops.decode_heatmap(self.d_hm[slot][:n], L.MASK_NEG1, step.refine, L.XFORM_CENTER_SCALE,
                               self.d_center[a:b], self.d_scale[a:b],
                               hm_flip=self.d_hf[slot][:n] if self.flip else None, flip_index=self.flip_index,
                               blur_ksize=step.kernel, want_idx=self.want_idx, render=render,
                               joints=self.d_joints[a:b] if self.has_loss else None,
                               vis=self.d_vis[a:b] if self.has_loss else None, out=out)
            self.freed[slot].record(comp)
            self.launches += 1
        if self.has_loss:
            sums = ops.loss_reduce(self.d_partials)
            ops.loss_finalize(sums, step.loss_mode, "mean", step.loss_weight, out=self.d_loss)
            self.launches += 2
            self.h_loss.copy_(self.d_loss, non_blocking=True)
        self.h_preds.copy_(self.d_preds, non_blocking=True)
        if self.want_idx:
            self.h_idx.copy_(self.d_idx, non_blocking=True)
        if self.metrics is not None:
            ops.metrics_finalize(self.d_counters, self.K, self.metrics["auc_steps"], out=self.d_metrics)
            self.launches += 1
            self.h_metrics.copy_(self.d_metrics, non_blocking=True)
            self.h_counters.copy_(self.d_counters, non_blocking=True)
        comp.synchronize()
        return self.h_preds, (self.h_loss if self.has_loss else None)


class BoundFusedStep:
    """A FusedHeatmapStep bound to fixed device buffers: every ctypes argument is built once, outputs are
    preallocated, and a steady-state step is ONE raw C-ABI launch (lhn_fused_render_loss_decode: the kernel
    reduces and finalises the loss itself).  ``capture()`` records it into a CUDA graph (the inner loop is
    launch-bound at ~0.1 ms of GPU work per step)."""

    def __init__(self, step, hm, joints_3d, joints_3d_visible, center, scale, hm_flip=None,
                 finalize=True, overlap_previous=False, spare_sms=0, accumulate_into=None, outputs=None,
                 exchange=None):
        """overlap_previous: this step shares no buffer with the step launched just before it on the stream
        (rotating input/output sets), so its kernel may start while that one drains (LHN_FLAG_OVERLAP_PREVIOUS);
        it is still ordered after every EARLIER launch, so two rotating sets are enough.
        accumulate_into: f32 [1] device tensor that receives ``+= loss`` of every step (the epoch sum that
        train_one_epoch keeps in loss_dict['sum'], left on the device; LHN_FLAG_ACCUMULATE_LOSS).
        exchange: a dist.PeerExchange — the four loss sums are all-reduced over the ranks INSIDE the kernel (peer-mapped
        mailboxes over NVLink, lhn_fused_render_loss_decode_xch): `sums` / `loss` are then the batch-global values a
        single process would compute on the concatenated batch, still one launch per step and no NCCL call.
        outputs: optional dict of caller-owned contiguous output tensors (hm_preds / preds f32 [B,K,3], idx int32
        [B,K], weight f32 [B,K], sums f64 [4], loss f32 [1]; larger leading sizes are fine) — e.g. the other
        output set of a rotation whose steps have different batch sizes."""
        import ctypes as C
        self.step = step
        lib = L.lib()
        self._lib = lib
        hm, B, Cc, H, W, sb, sc = ops._plane_view(hm, "heatmaps")
        dev = hm.device
        self.dev = dev
        for name, t in (("hm_flip", hm_flip), ("joints_3d", joints_3d), ("joints_3d_visible", joints_3d_visible),
                        ("center", center), ("scale", scale), ("accumulate_into", accumulate_into)):
            if t is not None and t.device != dev:
                raise L.LhnError(f"{name} is on {t.device}, the heatmaps on {dev}")
        self.B, self.K, self.H, self.W = B, Cc, H, W
        fb = fc = 0
        if hm_flip is not None:
            hm_flip, _, _, _, _, fb, fc = ops._plane_view(hm_flip, "flipped heatmaps")
        fi = None
        if hm_flip is not None and step.flip_pairs:
            fi = flip_index_from_pairs(Cc, step.flip_pairs, dev)
        self._keep = [hm, hm_flip, fi]
        joints = ops._f32c(joints_3d, "joints")
        vis = ops._f32c(joints_3d_visible, "vis")
        if vis.dim() == 2:
            vis = vis.unsqueeze(-1).contiguous()
        center, scale = ops._f32c(center, "center"), ops._f32c(scale, "scale")
        self._keep += [joints, vis, center, scale]
        self.dp = ops._decode_params(L.MASK_NEG1, step.refine, L.XFORM_CENTER_SCALE, blur_ksize=step.kernel,
                                     flags=(L.FLAG_OVERLAP_PREVIOUS if overlap_previous else 0) |
                                     (L.FLAG_ACCUMULATE_LOSS if accumulate_into is not None else 0) |
                                     ((int(spare_sms) & 0xff) << 8))
        self.rp = ops._render_params(step.loss_mode, step.image_size, step.sigma, step.unbiased, step.pos_value)
        outputs = outputs or {}

        def _out(name, shape, dtype):
            t = outputs.get(name)
            if t is None:
                return torch.empty(shape, dtype=dtype, device=dev)
            n = 1
            for d in shape:
                n *= d
            if t.dtype != dtype or t.device != dev or not t.is_contiguous() or t.numel() < n:
                raise L.LhnError(f"output '{name}' has the wrong dtype/device/layout or is too small")
            return t

        self.hm_preds = _out("hm_preds", (B, Cc, 3), torch.float32)
        self.preds = _out("preds", (B, Cc, 3), torch.float32)
        self.idx = _out("idx", (B, Cc), torch.int32)
        self.weight = _out("weight", (B, Cc), torch.float32)
        self.partials = torch.empty((B * Cc, 4), dtype=torch.float64, device=dev)
        self.sums = _out("sums", (4,), torch.float64)
        self.loss = _out("loss", (1,), torch.float32) if accumulate_into is None else accumulate_into
        self.finalize = finalize
        # one-launch step: the kernel reduces the loss sums itself; `finalize=False` (multi-GPU) leaves the
        # f64 sums for the cross-rank all-reduce and finalises afterwards with launch_finalize()
        with L.on_device(dev):       # the workspace size depends on the device's SM count
            nbytes = int(lib.lhn_fused_workspace_bytes(B, Cc // self.rp.num_stacks, self.rp.num_stacks))
        self.workspace = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
        self._fused_args = (
            L.ptr(hm), L.ptr(hm_flip), L.ptr(fi), L.dtype_code(hm), B, Cc // self.rp.num_stacks, H, W, sb, sc, fb, fc,
            L.ptr(center), L.ptr(scale), C.byref(self.dp), L.ptr(self.hm_preds), L.ptr(self.preds),
            L.ptr(self.idx), C.byref(self.rp), L.ptr(joints), joints.shape[2], L.ptr(vis), vis.shape[2],
            L.ptr(self.weight), None, L.ptr(self.workspace), self.workspace.numel(), L.ptr(self.sums), 0,
            float(step.loss_weight), L.ptr(self.loss) if finalize else None)
        self.exchange = exchange
        if exchange is not None:
            self.xch = exchange.struct()
            self._xch_args = self._fused_args + (C.byref(self.xch),)
        self._decode_args = (
            L.ptr(hm), L.ptr(hm_flip), L.ptr(fi), L.dtype_code(hm), B, Cc, H, W, sb, sc, fb, fc,
            L.ptr(center), L.ptr(scale), C.byref(self.dp), L.ptr(self.hm_preds), L.ptr(self.preds),
            L.ptr(self.idx), C.byref(self.rp), L.ptr(joints), joints.shape[2], L.ptr(vis), vis.shape[2],
            L.ptr(self.weight), L.ptr(self.partials))
        self._p_partials, self._p_sums, self._p_loss = L.ptr(self.partials), L.ptr(self.sums), L.ptr(self.loss)
        self.n_planes = B * Cc
        self.graph = None
        self.launches_per_step = 1 if finalize else 2     # N > 1: + finalize after the all-reduce

    def launch_kernel(self, stream):
        if self.exchange is not None:
            self.xch.seq = self.exchange.next_seq()
            L.check(self._lib.lhn_fused_render_loss_decode_xch(*self._xch_args, stream), "lhn_fused_render_loss_decode_xch")
            return
        L.check(self._lib.lhn_fused_render_loss_decode(*self._fused_args, stream), "lhn_fused_render_loss_decode")

    def launch_kernel_partials(self, stream):
        """The three-launch form (per-plane partials -> reduce -> finalise), kept for comparison tests."""
        L.check(self._lib.lhn_decode_heatmap(*self._decode_args, stream), "lhn_decode_heatmap")

    def launch_reduce(self, stream):
        L.check(self._lib.lhn_loss_reduce(self._p_partials, self.n_planes, self._p_sums, 0, stream), "lhn_loss_reduce")

    def launch_finalize(self, stream):
        L.check(self._lib.lhn_loss_finalize(self._p_sums, self.step.loss_mode, 0, self.step.loss_weight,
                                            self._p_loss, 0, stream), "lhn_loss_finalize")

    def stream(self):
        """torch's current stream on this step's device, as the raw handle the C ABI takes."""
        return L.stream(self.dev)

    def launch(self, events=None):
        """Issue the step on the current stream of the step's device.  events=(before, after) brackets the kernel."""
        with L.on_device(self.dev):
            st = L.stream(self.dev)
            if events is not None:
                events[0].record()
            self.launch_kernel(st)
            if events is not None:
                events[1].record()

    def capture(self):
        """Record the step into a CUDA graph (warm-up launch first so attributes are set eagerly)."""
        self.launch()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.launch()
        self.graph = g
        return g

    def replay(self):
        self.graph.replay()


class SimdrHostPipeline:
    """End-to-end SimDR decode for HOST vectors (keypoints_from_simdr, top_down_eval.py:466-500, which takes NumPy
    arrays): chunked host->device copies on a copy stream overlapped with the decode kernel, [B,K,3] read back."""

    def __init__(self, B, K, Lx, Ly, k=2, dtype=torch.float32, chunks=8, device="cuda"):
        self.B, self.K, self.Lx, self.Ly, self.k = B, K, Lx, Ly, int(k)
        self.dev = torch.device(device)
        chunks = max(1, min(chunks, B))
        edges = [round(i * B / chunks) for i in range(chunks + 1)]
        self.bounds = [(a, b) for a, b in zip(edges[:-1], edges[1:]) if b > a]
        cb = max(b - a for a, b in self.bounds)
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.d_x = [torch.empty((cb, K, Lx), dtype=dtype, device=self.dev) for _ in range(2)]
        self.d_y = [torch.empty((cb, K, Ly), dtype=dtype, device=self.dev) for _ in range(2)]
        self.d_center = torch.empty((B, 2), device=self.dev)
        self.d_scale = torch.empty((B, 2), device=self.dev)
        self.d_out = torch.empty((B, K, 3), device=self.dev)
        self.h_out = torch.empty((B, K, 3), pin_memory=True)
        self.copied = [torch.cuda.Event() for _ in self.bounds]
        self.freed = [torch.cuda.Event() for _ in range(2)]
        esz = torch.empty((), dtype=dtype).element_size()
        self.h2d_bytes = B * K * (Lx + Ly) * esz + B * 16
        self.d2h_bytes = B * K * 3 * 4
        self.launches = 0

    def __call__(self, x_vec, y_vec, center, scale):
        def host(t):
            return t if isinstance(t, torch.Tensor) else torch.as_tensor(t)

        x_vec, y_vec, center, scale = host(x_vec), host(y_vec), host(center), host(scale)
        comp = torch.cuda.current_stream(self.dev)
        cs = self.copy_stream
        cs.wait_stream(comp)
        with torch.cuda.stream(cs):
            self.d_center.copy_(center, non_blocking=True)
            self.d_scale.copy_(scale, non_blocking=True)
        self.launches = 0
        for i, (a, b) in enumerate(self.bounds):
            slot, n = i & 1, b - a
            with torch.cuda.stream(cs):
                if i >= 2:
                    cs.wait_event(self.freed[slot])
                self.d_x[slot][:n].copy_(x_vec[a:b], non_blocking=True)
                self.d_y[slot][:n].copy_(y_vec[a:b], non_blocking=True)
                self.copied[i].record(cs)
            comp.wait_event(self.copied[i])
            ops.decode_simdr(self.d_x[slot][:n], self.d_y[slot][:n], self.k, self.d_center[a:b], self.d_scale[a:b],
                             out=self.d_out[a:b])
            self.freed[slot].record(comp)
            self.launches += 1
        self.h_out.copy_(self.d_out, non_blocking=True)
        comp.synchronize()
        return self.h_out


class BoundDecodeStep:
    """A decode-only launch (argmax + refinement + back-transform[, fused PCK/AUC/EPE counters]) bound to fixed
    device buffers: every ctypes argument is built once, a step is ONE raw C-ABI launch (lhn_decode_heatmap, or
    lhn_decode_heatmap_pck when `metrics` is given).  The eager ops wrappers cost 20-30 us of Python per call,
    which is more than the kernel at the reference's batch sizes (BASELINE config 1: batch 64)."""

    def __init__(self, hm, center, scale, mask_mode=L.MASK_NEG1, refine=L.REFINE_SIGN, transform=L.XFORM_CENTER_SCALE,
                 hm_flip=None, flip_pairs=(), kernel=11, scale_xy=(1.0, 1.0), overlap_previous=False, metrics=None,
                 outputs=None):
        """metrics: None or dict(gt [B,K,2] f32, mask [B,K] bool/u8, bbox_wh [B,2] f32, counters int64
        [(auc_steps+5)*K], pck_thr=0.2, auc_nor=30.0, auc_steps=20), or with the in-kernel cross-rank exchange
        dict(gt, mask, bbox_wh, exchange=dist.PeerExchange, totals=int64 [(auc_steps+5)*K], ...): every launch accumulates
        into one of the exchange's rotating per-step blocks and all-reduces the PREVIOUS launch's block into `totals`
        inside the kernel (lhn_decode_heatmap_pck_xch); exchange.flush() completes the last step."""
        import ctypes as C
        lib = L.lib()
        self._lib = lib
        hm, B, Cc, H, W, sb, sc = ops._plane_view(hm, "heatmaps")
        dev = hm.device
        self.dev, self.B, self.K, self.H, self.W = dev, B, Cc, H, W
        fb = fc = 0
        fi = None
        if hm_flip is not None:
            hm_flip, _, _, _, _, fb, fc = ops._plane_view(hm_flip, "flipped heatmaps")
            if flip_pairs:
                fi = flip_index_from_pairs(Cc, flip_pairs, dev)
        center, scale = ops._f32c(center, "center"), ops._f32c(scale, "scale")
        self.dp = ops._decode_params(mask_mode, refine, transform, scale_xy, kernel,
                                     flags=L.FLAG_OVERLAP_PREVIOUS if overlap_previous else 0)
        outputs = outputs or {}

        def _out(name, shape, dtype):
            t = outputs.get(name)
            if t is None:
                return torch.empty(shape, dtype=dtype, device=dev)
            n = 1
            for d in shape:
                n *= d
            if t.dtype != dtype or t.device != dev or not t.is_contiguous() or t.numel() < n:
                raise L.LhnError(f"output '{name}' has the wrong dtype/device/layout or is too small")
            return t

        self.hm_preds = _out("hm_preds", (B, Cc, 3), torch.float32)
        self.preds = _out("preds", (B, Cc, 3), torch.float32)
        self.idx = _out("idx", (B, Cc), torch.int32)
        self._keep = [hm, hm_flip, fi, center, scale]
        self.counters = None
        if metrics is None:
            self._fn, self._name = lib.lhn_decode_heatmap, "lhn_decode_heatmap"
            self._args = (L.ptr(hm), L.ptr(hm_flip), L.ptr(fi), L.dtype_code(hm), B, Cc, H, W, sb, sc, fb, fc,
                          L.ptr(center), L.ptr(scale), C.byref(self.dp), L.ptr(self.hm_preds), L.ptr(self.preds),
                          L.ptr(self.idx), None, None, 0, None, 0, None, None)
        else:
            if hm_flip is not None:
                raise L.LhnError("the fused-metrics decode takes no flip plane")
            gt = ops._f32c(metrics["gt"], "gt")
            mask = ops._mask_u8(metrics["mask"])
            wh = ops._f32c(metrics["bbox_wh"], "bbox_wh")
            steps = int(metrics.get("auc_steps", 20))
            self.exchange = metrics.get("exchange")
            self.counters = metrics["counters"] if self.exchange is None else \
                self.exchange.step_blocks((steps + 5) * Cc)[0][:(steps + 5) * Cc]
            if self.counters.dtype != torch.int64 or self.counters.numel() != (steps + 5) * Cc or \
                    not self.counters.is_contiguous() or self.counters.device != dev:
                raise L.LhnError("counters must be a contiguous int64 tensor of (auc_steps+5)*K entries on the heatmaps' device")
            self._keep += [gt, mask, wh, self.counters]
            self._fn, self._name = lib.lhn_decode_heatmap_pck, "lhn_decode_heatmap_pck"
            self._args = [L.ptr(hm), L.dtype_code(hm), B, Cc, H, W, sb, sc, L.ptr(center), L.ptr(scale), C.byref(self.dp),
                          L.ptr(self.hm_preds), L.ptr(self.preds), L.ptr(self.idx), L.ptr(gt), L.ptr(mask), L.ptr(wh),
                          float(metrics.get("pck_thr", 0.2)), float(metrics.get("auc_nor", 30.0)), steps,
                          L.ptr(self.counters)]
            if self.exchange is not None:
                totals = metrics["totals"]
                if totals.dtype != torch.int64 or totals.numel() != self.counters.numel() or not totals.is_contiguous() \
                        or totals.device != dev:
                    raise L.LhnError("totals must be a contiguous int64 tensor of (auc_steps+5)*K entries")
                self.totals, self._n_cnt = totals, (steps + 5) * Cc
                self._keep.append(totals)
                self.xch = self.exchange.struct()
                self._fn, self._name = lib.lhn_decode_heatmap_pck_xch, "lhn_decode_heatmap_pck_xch"
                self._args = self._args + [L.ptr(totals), C.byref(self.xch)]
        self.graph = None

    def stream(self):
        return L.stream(self.dev)

    def launch_kernel(self, stream):
        if getattr(self, "exchange", None) is not None:
            # this step's block; the two previous steps' blocks for this launch to publish / add up
            self._args[20] = self.exchange.begin_step(self._n_cnt, self.totals, self.xch)
        L.check(self._fn(*self._args, stream), self._name)

    def launch(self):
        with L.on_device(self.dev):
            self.launch_kernel(L.stream(self.dev))

    def capture(self):
        self.launch()
        torch.cuda.synchronize(self.dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.launch()
        self.graph = g
        return g

    def replay(self):
        self.graph.replay()


class BoundSimdrStep:
    """lhn_decode_simdr_flags bound to fixed buffers (one raw C-ABI launch per step; BASELINE config 3)."""

    def __init__(self, x_vec, y_vec, k=2, center=None, scale=None, overlap_previous=False, out=None):
        L.require_cuda(x_vec, "x_vectors")
        L.require_cuda(y_vec, "y_vectors")
        x_vec, y_vec = x_vec.contiguous(), y_vec.to(x_vec.dtype).contiguous()
        B, K, Lx = x_vec.shape
        Ly = y_vec.shape[2]
        self.dev, self.B, self.K = x_vec.device, B, K
        center, scale = ops._f32c(center, "center"), ops._f32c(scale, "scale")
        self.out = out if out is not None else torch.empty((B, K, 3), dtype=torch.float32, device=self.dev)
        self.idx = torch.empty((B, K, 2), dtype=torch.int32, device=self.dev)
        self._keep = [x_vec, y_vec, center, scale]
        self._lib = L.lib()
        self._args = (L.ptr(x_vec), L.ptr(y_vec), L.dtype_code(x_vec), B, K, Lx, Ly, int(k), L.ptr(center), L.ptr(scale),
                      0, None, L.ptr(self.out), L.ptr(self.idx), L.FLAG_OVERLAP_PREVIOUS if overlap_previous else 0)

    def stream(self):
        return L.stream(self.dev)

    def launch_kernel(self, stream):
        L.check(self._lib.lhn_decode_simdr_flags(*self._args, stream), "lhn_decode_simdr_flags")

    def launch(self):
        with L.on_device(self.dev):
            self.launch_kernel(L.stream(self.dev))
