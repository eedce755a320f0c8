"""Drop-in mirrors of the reference's decode entry points — same names, arguments and return
structure; the per-plane work (argmax, refinement, flip average, back-transform) runs in one CUDA
kernel on the device-resident heatmaps.

Mirrored (file:line relative to the reference root):
  utils/post_processing/evaluation/top_down_eval.py:199  _get_max_preds
  utils/post_processing/evaluation/top_down_eval.py:375  keypoints_from_heatmaps
  utils/post_processing/evaluation/top_down_eval.py:466  keypoints_from_simdr
  utils/post_processing/decoder.py:9                     TopDownDecoder (.decode / .decode_simdr)
  utils/heatmap_post_processing.py:6,35                  adjust_keypoints_by_offset / _by_DARK
  utils/result_parser.py:14                              ResultParser (.get_coordinates_from_heatmaps,
                                                         .get_pred_kpt, .vector_nms, .get_coordinates_from_vectors,
                                                         .heatmap_nms, .candidate_bbox, .non_max_suppression,
                                                         .get_pred_bbox, ._get_first_result, .get_group_keypoints)
  utils/SPheatmapParser.py:12                            HeatmapParser_SH (.get_coordinates, .adjust_keypoints, .parse,
                                                         .heatmap_nms, .candidate_bbox, .non_max_suppression)
  utils/evaluation.py:94,170                             cs_from_region_map, non_max_suppression
  utils/HeatmapParser.py:197                             HeatmapParser.adjust_keypoints (list-of-lists form)
  utils/transforms.py:18,47,78                           get_final_preds, get_max_preds, flip_back
  utils/evaluation.py:62                                 get_coordinates_from_heatmap

Array convention: the reference takes/returns NumPy for the Gen-2 functions and torch tensors for the
legacy ones.  Here every function accepts CUDA tensors (preferred: no copies) and also NumPy arrays /
CPU tensors, which are uploaded; the result comes back in the kind the reference returns when the
input was NumPy/CPU, and as CUDA tensors when the input was a CUDA tensor.
"""
import numpy as np
import torch

from . import _lib as L
from . import ops
from .fused import flip_index_from_pairs

_DEV = None


def _device():
    global _DEV
    if _DEV is None:
        if not torch.cuda.is_available():
            raise L.LhnError("no CUDA device: the B200 hot path has no CPU fallback")
        _DEV = torch.device("cuda", torch.cuda.current_device())
    return _DEV


def _up(x, dtype=None):
    """-> (cuda tensor, was_cuda)."""
    if isinstance(x, torch.Tensor):
        was = x.is_cuda
        t = x.detach() if was else x.detach().to(_device())
    else:
        was = False
        t = torch.as_tensor(np.ascontiguousarray(x)).to(_device())
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t, was


def _hm(x):
    t, was = _up(x)
    if t.dtype not in (torch.float32, torch.bfloat16, torch.float16):
        t = t.float()
    return t, was


def _np(t):
    return t.detach().cpu().numpy()


# ---- Gen-2 (mmpose-style) ---------------------------------------------------------------------------
def _get_max_preds(heatmaps):
    """top_down_eval.py:199-231 -> (preds [N,K,2], maxvals [N,K,1]); coords = -1 where max <= 0."""
    if not isinstance(heatmaps, (np.ndarray, torch.Tensor)):
        raise AssertionError('heatmaps should be numpy.ndarray')
    assert heatmaps.ndim == 4, 'batch_images should be 4-ndim'
    t, was = _hm(heatmaps)
    r = ops.decode_heatmap(t, L.MASK_NEG1, L.REFINE_NONE, want_idx=False)
    preds, maxvals = r["hm_kpts"][..., :2], r["hm_kpts"][..., 2:]
    return (preds, maxvals) if was else (_np(preds), _np(maxvals))


def keypoints_from_heatmaps(heatmaps, center, scale, post_process='default', kernel=11, use_udp=False,
                            target_type='GaussianHeatmap', only_original_preds=False,
                            heatmaps_flipped=None, flip_pairs=()):
    """top_down_eval.py:375-463.  Returns (hm_preds [N,K,2], preds [N,K,2], maxvals [N,K,1]) or
    (preds, maxvals) with only_original_preds.  The input heatmaps are never modified.

    Additive: ``heatmaps_flipped`` (+ ``flip_pairs``) fuses the flip-test average
    (heatmaps + flip_back(heatmaps_flipped, flip_pairs)) * 0.5 into the same pass."""
    if use_udp and target_type.lower() != 'gaussianheatmap':
        raise NotImplementedError("UDP 'CombinedTarget' decoding: the reference's keypoints_from_heatmaps leaves "
                                  "hm_preds undefined for it (top_down_eval.py:427-431)")
    if post_process == 'unbiased':
        assert kernel > 0
    if post_process == 'megvii':
        raise NotImplementedError("'megvii' post-processing is not in the reference either")
    t, was = _hm(heatmaps)
    c, _ = _up(center, torch.float32)
    s, _ = _up(scale, torch.float32)
    refine = L.REFINE_DARK if post_process == 'unbiased' else (L.REFINE_SIGN if post_process is not None else L.REFINE_NONE)
    if use_udp:
        refine = L.REFINE_DARK_UDP            # top_down_eval.py:427-431: post_dark_udp whatever post_process says
    hf = fi = None
    if heatmaps_flipped is not None:
        hf, _ = _hm(heatmaps_flipped)
        hf = hf.to(t.dtype)
        if flip_pairs:
            fi = flip_index_from_pairs(t.shape[1], flip_pairs, t.device)
    r = ops.decode_heatmap(t, L.MASK_NEG1, refine, L.XFORM_CENTER_SCALE, c, s, hm_flip=hf, flip_index=fi,
                           blur_ksize=kernel, use_udp=bool(use_udp), want_idx=False)
    hm_preds, preds, maxvals = r["hm_kpts"][..., :2], r["kpts"][..., :2], r["kpts"][..., 2:]
    if not was:
        hm_preds, preds, maxvals = _np(hm_preds), _np(preds), _np(maxvals)
    if only_original_preds:
        return preds, maxvals
    return hm_preds, preds, maxvals


def keypoints_from_simdr(x_vectors, y_vectors, center, scale, k=2):
    """top_down_eval.py:466-500 -> [B,K,3] (x, y, score)."""
    assert k > 0, f"Error: {k=}"
    xv, was = _hm(x_vectors)
    yv, _ = _hm(y_vectors)
    c, _ = _up(center, torch.float32)
    s, _ = _up(scale, torch.float32)
    out = ops.decode_simdr(xv, yv, int(k), c, s)
    return out if was else _np(out)


class TopDownDecoder:
    """utils/post_processing/decoder.py:9-107."""

    def __init__(self, cfg):
        self.image_size = np.array(cfg.DATASET.image_size)
        self.heatmap_size = np.array(cfg.DATASET.heatmap_size)
        self.num_joints = cfg.DATASET.num_joints
        self.post_process = 'unbiased' if cfg.PIPELINE.unbiased_encoding else 'default'
        self.kernel = cfg.PIPELINE.kernel[0]
        self.use_udp = cfg.PIPELINE.use_udp
        self.k = cfg.PIPELINE.get('simdr_split_ratio', 0)

    @staticmethod
    def _boxes(center, scale, score):
        n = center.shape[0]
        all_boxes = np.zeros((n, 6), dtype=np.float32)
        all_boxes[:, 0:2] = center[:, 0:2]
        all_boxes[:, 2:4] = scale[:, 0:2]
        all_boxes[:, 4] = np.prod(scale * 200.0, axis=1)
        all_boxes[:, 5] = score
        return all_boxes

    def decode(self, meta, model_output, model_output_flipped=None, flip_pairs=()):
        """-> dict(preds [N,K,3], hm_preds [N,K,3], boxes [N,6], image_paths, bbox_ids,
        output_heatmap).  ``model_output`` stays on the GPU; only the [N,K,3] results (and the
        heatmap copy the reference's dict carries) cross to the host."""
        score = _np(torch.as_tensor(meta['bbox_score']))
        bbox_ids = _np(torch.as_tensor(meta['bbox_id']))
        center_t = torch.as_tensor(meta['center'])
        scale_t = torch.as_tensor(meta['scale'])
        hm = model_output[:, :self.num_joints]                     # a view: no copy (strided planes)
        hf = None if model_output_flipped is None else model_output_flipped[:, :self.num_joints]
        hm_preds, preds, maxvals = keypoints_from_heatmaps(
            hm if hm.is_cuda else hm.to(_device()), center_t, scale_t, post_process=self.post_process,
            kernel=self.kernel, use_udp=self.use_udp, heatmaps_flipped=hf, flip_pairs=flip_pairs)
        hm_preds, preds, maxvals = _np(hm_preds), _np(preds), _np(maxvals)
        center, scale = _np(center_t).astype(np.float32), _np(scale_t).astype(np.float32)
        batch_size = model_output.shape[0]
        all_preds = np.zeros((batch_size, self.num_joints, 3), dtype=np.float32)
        all_preds[:, :, 0:2] = preds[:, :, 0:2]
        all_preds[:, :, 2:3] = maxvals
        result = {}
        result['preds'] = all_preds
        result['hm_preds'] = np.concatenate([hm_preds[:, :, 0:2] * 4, maxvals], axis=2)   # hard-coded x4 (decoder.py:65)
        result['boxes'] = self._boxes(center, scale, score)
        result['image_paths'] = meta['image_file']
        result['bbox_ids'] = bbox_ids.tolist()
        result['output_heatmap'] = _np(hm.float())
        return result

    def decode_simdr(self, meta, model_output):
        score = _np(torch.as_tensor(meta['bbox_score']))
        bbox_ids = _np(torch.as_tensor(meta['bbox_id']))
        center_t = torch.as_tensor(meta['center'])
        scale_t = torch.as_tensor(meta['scale'])
        sx, sy = torch.as_tensor(meta['simdr_x']), torch.as_tensor(meta['simdr_y'])
        all_preds = _np(torch.as_tensor(keypoints_from_simdr(sx.to(_device()), sy.to(_device()),
                                                             center_t, scale_t, self.k)))
        center, scale = _np(center_t).astype(np.float32), _np(scale_t).astype(np.float32)
        result = {}
        result['preds'] = all_preds
        result['boxes'] = self._boxes(center, scale, score)
        result['image_paths'] = meta['image_file']
        result['bbox_ids'] = bbox_ids.tolist()
        result['output_heatmap'] = _np(model_output[:, :self.num_joints].float())
        return result


# ---- legacy (Gen-1) ---------------------------------------------------------------------------------
# config/__init__.py:4-24 (the post-processing constants the legacy parsers read at construction)
pcfg = {
    "num_candidates": 10, "max_num_bbox": 1, "nms_kernel": 11, "nms_stride": 1, "nms_padding": 5,
    "detection_threshold": 0.1, "iou_threshold": 0.6, "bbox_factor": 1.3, "region_avg_kernel": 3,
    "region_avg_stride": 1, "blue_kernel": 19, "cd_iou": 0.3, "cd_ratio": 0,
}


def _boxes_to_lists(boxes, counts):
    """(boxes [B,M,5], counts [B]) on the device -> the reference's list over images: None or
    [[x, y, w, h, conf], ...] (x[index].tolist())."""
    bx = boxes.cpu().tolist()
    ct = counts.cpu().tolist()
    return [bx[i][:n] if n > 0 else None for i, n in enumerate(ct)]


def _region_inplace(center_maps):
    """The reference's heatmap_nms masks its argument in place; do the same when the argument is a CUDA tensor
    with contiguous planes (a channel slice of a region map qualifies).  -> (cuda tensor, inplace)"""
    t, was = _hm(center_maps)
    ok = was and t is not None and t.dim() == 4 and t.stride(3) == 1 and t.stride(2) == t.shape[3] \
        and isinstance(center_maps, torch.Tensor) and t.data_ptr() == center_maps.data_ptr()
    return t, ok


def adjust_keypoints_by_offset(keypoints, heatmaps):
    """utils/heatmap_post_processing.py:6-33: keypoints [B,K,3] (x,y,conf) -> +-0.25 towards the higher clamped
    neighbour, then +0.5.  Returns a new tensor (callers pass .clone() to the reference); the confidence column is
    preserved.  Keypoints that are the per-plane argmax (every call site of the reference) take the fused decode
    kernel; any other positions are refined where they are (lhn_refine_points)."""
    kp, was = _up(keypoints, torch.float32)
    t, _ = _hm(heatmaps)
    out = kp.clone()
    out[..., :2] = _refine_offset(kp, t, L.REFINE_OFFSET_HALF)
    return out if was else (out.cpu() if isinstance(keypoints, torch.Tensor) else _np(out))


def _refine_offset(kp, hm, refine):
    """The +-0.25 rule of D1 / D2 around int(kp[..., :2]): the decode kernel's own result when kp is the plane
    argmax, else the rule evaluated at the given positions."""
    r = ops.decode_heatmap(hm, L.MASK_NONE, refine, want_idx=False)["hm_kpts"][..., :2]
    r0 = ops.decode_heatmap(hm, L.MASK_NONE, L.REFINE_NONE, want_idx=False)["hm_kpts"][..., :2]
    if torch.equal(kp[..., :2], r0):
        return r
    B, K = kp.shape[:2]
    xy = kp[..., :2].reshape(-1, 2).contiguous().clone()
    ops.refine_points(hm, _plane_index(B, K, kp.device), xy, plus_half=(refine == L.REFINE_OFFSET_HALF))
    return xy.reshape(B, K, 2)


def _plane_index(B, K, device):
    return torch.stack(torch.meshgrid(torch.arange(B, device=device, dtype=torch.int32),
                                      torch.arange(K, device=device, dtype=torch.int32), indexing="ij"), -1).reshape(-1, 2)


def adjust_keypoints_by_DARK(keypoints, heatmaps):
    """utils/heatmap_post_processing.py:35-54: blur k=pcfg['blue_kernel']=19 (f64), log, Taylor around
    int(keypoints).  Returns a NumPy array like the reference; the input heatmap is left untouched (the reference's
    CUDA semantics — on CPU inputs the reference blurs the caller's array in place).  Keypoints that are the
    per-plane argmax (get_pred_kpt's call pattern) take the fused decode kernel; any other positions — e.g.
    candidate_bbox's top-k candidates — are refined where they are (lhn_dark_refine_points)."""
    kp, _ = _up(keypoints, torch.float32)
    t, _ = _hm(heatmaps)
    out = kp.clone()
    r0 = ops.decode_heatmap(t, L.MASK_NONE, L.REFINE_NONE, want_idx=False)["hm_kpts"][..., :2]
    if torch.equal(kp[..., :2], r0):
        out[..., :2] = ops.decode_heatmap(t, L.MASK_NONE, L.REFINE_DARK_LEGACY, want_idx=False)["hm_kpts"][..., :2]
    else:
        B, K = kp.shape[:2]
        xy = kp[..., :2].reshape(-1, 2).contiguous().clone()
        ops.dark_refine_points(t, _plane_index(B, K, kp.device), xy, pcfg['blue_kernel'])
        out[..., :2] = xy.reshape(B, K, 2)
    return _np(out)


class ResultParser:
    """utils/result_parser.py:14 — keypoint branch (get_coordinates_from_heatmaps, get_pred_kpt, vector_nms,
    get_coordinates_from_vectors) and bbox branch (heatmap_nms, candidate_bbox, non_max_suppression,
    get_pred_bbox, _get_first_result, get_group_keypoints without cycle detection — the second pass re-runs the
    model, which is outside the path)."""

    def __init__(self, cfg):
        self.cfg = cfg
        self.nms_kernel = pcfg["nms_kernel"]
        if pcfg["nms_stride"] != 1 or 2 * pcfg["nms_padding"] + 1 != pcfg["nms_kernel"]:
            raise L.LhnError("only the size-preserving centre-map max-pool (stride 1, kernel = 2 padding + 1) is supported")
        self.avg_kernel = pcfg["region_avg_kernel"]
        self.num_candidates = pcfg["num_candidates"]
        self.max_num_bbox = pcfg["max_num_bbox"]
        self.detection_threshold = pcfg["detection_threshold"]
        self.iou_threshold = pcfg["iou_threshold"]
        self.bbox_factor = pcfg["bbox_factor"]
        self.image_area = cfg['image_size'][0] * cfg['image_size'][1]
        self.bbox_alpha = cfg.get('bbox_alpha')
        self.cd_enabled = cfg.get('with_region_map', False)
        self.cd_reduction = cfg.get('cycle_detection_reduction')
        self.image_size = torch.tensor(cfg['image_size'])
        if cfg['model'] == 'srhandnet':
            self.heatmap_size = torch.tensor([cfg['hm_size'][-1], cfg['hm_size'][-1]])
        else:
            self.heatmap_size = torch.tensor(cfg['hm_size'])
        self.feature_stride = torch.div(self.image_size, self.heatmap_size, rounding_mode='trunc')
        self.simdr_split_ratio = cfg['simdr_split_ratio']

    def get_coordinates_from_heatmaps(self, heatmaps):
        """topk(k=1) -> kpts [B,K,3] (x, y, score), no mask (result_parser.py:76-90)."""
        t, _ = _hm(heatmaps)
        return ops.decode_heatmap(t, L.MASK_NONE, L.REFINE_NONE, want_idx=False)["hm_kpts"]

    def get_pred_kpt(self, heatmap, resized=False):
        """result_parser.py:231-249 -> Tensor [B,K,3] on the heatmap's device."""
        t, was = _hm(heatmap)
        refine = L.REFINE_DARK_LEGACY if self.cfg['DARK'] else L.REFINE_OFFSET_HALF
        fs = self.feature_stride.tolist()
        r = ops.decode_heatmap(t, L.MASK_NONE, refine, L.XFORM_SCALE if resized else L.XFORM_NONE,
                               scale_xy=(float(fs[0]), float(fs[1])), want_idx=False)
        out = r["kpts"] if resized else r["hm_kpts"]
        return out if was else out.cpu()

    def vector_nms(self, vector):
        """result_parser.py:61-74 (returns a new tensor; the reference masks its argument in place)."""
        v, was = _hm(vector)
        out = ops.vector_nms(v)
        return out if was else out.cpu()

    # ---- bbox branch ----
    def heatmap_nms(self, heatmaps):
        """result_parser.py:50-59: hm * eq(maxpool11(hm), hm), in place on a CUDA tensor like the reference."""
        t, inplace = _region_inplace(heatmaps)
        out = ops.heatmap_nms(t, self.nms_kernel, inplace=inplace)
        return heatmaps if inplace else (out if isinstance(heatmaps, torch.Tensor) and heatmaps.is_cuda else out.cpu())

    def _region(self, center_maps, size_maps, nms_kernel, inplace):
        if not self.cfg['DARK']:
            # result_parser.py:161-166: with DARK off the reference passes the tensor adjust_keypoints_by_offset
            # returned to torch.from_numpy and raises — same error here
            raise TypeError("expected np.ndarray (got Tensor)")
        fs = self.feature_stride.tolist()
        return ops.region_bbox_decode(center_maps, size_maps, L.REGION_RP, nms_kernel=nms_kernel,
                                      num_candidates=self.num_candidates, max_num_bbox=self.max_num_bbox,
                                      avg_kernel=self.avg_kernel, refine=L.REFINE_DARK_LEGACY,
                                      blur_ksize=pcfg['blue_kernel'], stride=(float(fs[0]), float(fs[1])),
                                      det_thr=self.detection_threshold, iou_thr=self.iou_threshold,
                                      nms_inplace=inplace)

    def candidate_bbox(self, center_maps, size_maps):
        """result_parser.py:131-175 on an already NMS'd centre map -> candidates [B, k, 5] (a CPU tensor, as the
        reference allocates it).  Every candidate is refined on the same map (the reference's CUDA semantics)."""
        c, _ = _hm(center_maps)
        sm, _ = _hm(size_maps)
        return self._region(c, sm.to(c.dtype), 0, False)["candidates"].cpu()

    def non_max_suppression(self, candidates):
        """result_parser.py:177-215 -> list over images: None or [[x, y, w, h, conf], ...]."""
        c, _ = _up(candidates, torch.float32)
        return _boxes_to_lists(*ops.box_nms(c, self.detection_threshold, self.iou_threshold, self.max_num_bbox))

    def get_pred_bbox(self, region_map):
        """result_parser.py:217-229: region_map [B,3,H,W] -> list of bboxes; ONE launch (NMS + top-k + size
        lookup + DARK refinement + box NMS); channel 0 of a CUDA region_map is masked in place as in the reference."""
        t, _ = _hm(region_map)
        c, inplace = _region_inplace(region_map[:, 0:1]) if isinstance(region_map, torch.Tensor) else (t[:, 0:1], False)
        if not inplace:
            c = t[:, 0:1]
        r = self._region(c, t[:, 1:3], self.nms_kernel, inplace)
        return _boxes_to_lists(r["boxes"], r["counts"])

    def _first_result_roi(self, bbox, H, W):
        """result_parser.py:290-303: the enlarged bbox as a heatmap window, in f32 like the reference's 0-dim tensors."""
        f32 = np.float32
        stride = f32(int(self.feature_stride[0]))
        xc, yc, wb, hb = [f32(f32(v) / stride) for v in bbox[:4]]
        wb = int(f32(wb * f32(self.bbox_factor)))
        hb = int(f32(hb * f32(self.bbox_factor)))
        ul_x = max(0, int(f32(f32(xc - f32(wb / 2)) + f32(0.5))))
        ul_y = max(0, int(f32(f32(yc - f32(hb / 2)) + f32(0.5))))
        br_x, br_y = min(ul_x + wb, W), min(ul_y + hb, H)
        if br_x <= ul_x or br_y <= ul_y:
            return 0, 0, W, H
        return ul_x, ul_y, br_x, br_y

    def _decode_rois(self, heatmaps, rois):
        refine = L.REFINE_DARK_LEGACY if self.cfg['DARK'] else L.REFINE_OFFSET_HALF
        fs = self.feature_stride.tolist()
        roi = torch.tensor(rois, dtype=torch.int32).to(heatmaps.device)
        return ops.decode_heatmap_roi(heatmaps, roi, refine, scale_xy=(float(fs[0]), float(fs[1])),
                                      blur_ksize=pcfg['blue_kernel'] if self.cfg['DARK'] else None)

    def _get_first_result(self, bbox, heatmaps, img_idx: int):
        """result_parser.py:288-306: keypoints of image img_idx decoded inside the enlarged bbox only -> [1,K,3]."""
        t, was = _hm(heatmaps)
        H, W = t.shape[2:]
        out = self._decode_rois(t[img_idx:img_idx + 1], [self._first_result_roi(bbox, H, W)])
        return out if was else out.cpu()

    def get_group_keypoints(self, model, img, bbox_list, heatmaps):
        """result_parser.py:251-273 without cycle detection: one launch per bbox slot over the whole batch
        -> [B, max_num_bbox, K, 3] (zeros where an image has no such bbox; a CPU tensor, as the reference's)."""
        if self.cd_enabled:
            raise NotImplementedError("cycle detection re-runs the model on a crop (result_parser.py:308-352): "
                                      "outside the heatmap decode path")
        t, _ = _hm(heatmaps)
        B, K, H, W = t.shape
        pred = torch.zeros((B, self.max_num_bbox, K, 3), device=t.device)
        for j in range(self.max_num_bbox):
            has = [bl is not None and len(bl) > j for bl in bbox_list]
            if not any(has):
                continue
            rois = [self._first_result_roi(bl[j], H, W) if h else (0, 0, W, H) for bl, h in zip(bbox_list, has)]
            out = self._decode_rois(t, rois)
            pred[:, j] = out * torch.tensor(has, device=t.device)[:, None, None]
        return pred.cpu()

    def get_coordinates_from_vectors(self, x_vectors, y_vectors, pred_bboxes):
        """result_parser.py:92-129 with max_num_bbox = 1: vector_nms, bbox-masked first-max argmax, /k,
        mean score -> [B, 1, K, 3].  Samples whose bbox list is None stay zero."""
        xv, was = _hm(x_vectors)
        yv, _ = _hm(y_vectors)
        B, K, w = xv.shape
        h = yv.shape[2]
        ranges = np.zeros((B, 4), np.int32)
        valid = np.zeros(B, bool)
        for i in range(B):
            if pred_bboxes[i] is not None and len(pred_bboxes[i]) > 0:
                bbox = np.round(np.array(pred_bboxes[i][0]) * self.simdr_split_ratio)
                x1, y1 = bbox[:2] - bbox[2:4] / 2
                x2, y2 = bbox[:2] + bbox[2:4] / 2
                ranges[i] = (max(int(x1), 0), min(int(x2), w), max(int(y1), 0), min(int(y2), h))
                valid[i] = True
        out = ops.decode_simdr(xv, yv, int(self.simdr_split_ratio), nms=True,
                               ranges=torch.from_numpy(ranges).to(xv.device))
        out = out * torch.from_numpy(valid).to(xv.device)[:, None, None]
        out = out.unsqueeze(1)
        return out if was else out.cpu()


class HeatmapParser_SH:
    """utils/SPheatmapParser.py:12 — keypoint branch (get_coordinates, adjust_keypoints) and bbox branch
    (heatmap_nms, candidate_bbox, non_max_suppression), both behind parse()."""

    @staticmethod
    def get_coordinates(heatmaps):
        t, _ = _hm(heatmaps)
        return ops.decode_heatmap(t, L.MASK_NONE, L.REFINE_NONE, want_idx=False)["hm_kpts"].cpu()   # reference allocates on CPU (:50)

    @staticmethod
    def adjust_keypoints(keypoints, heatmaps):
        kp, _ = _up(keypoints, torch.float32)
        t, _ = _hm(heatmaps)
        out = kp.clone()
        out[..., :2] = _refine_offset(kp, t, L.REFINE_OFFSET)
        return out.cpu() if not (isinstance(keypoints, torch.Tensor) and keypoints.is_cuda) else out

    def __init__(self):
        self.nms_kernel = pcfg["nms_kernel"]
        self.avg_kernel = pcfg["region_avg_kernel"]
        self.num_candidates = pcfg["num_candidates"]
        self.max_num_bbox = pcfg["max_num_bbox"]
        self.detection_threshold = pcfg["detection_threshold"]
        self.iou_threshold = pcfg["iou_threshold"]

    def heatmap_nms(self, heatmaps):
        """SPheatmapParser.py:32-41 (in place on a CUDA tensor, like the reference)."""
        t, inplace = _region_inplace(heatmaps)
        out = ops.heatmap_nms(t, self.nms_kernel, inplace=inplace)
        return heatmaps if inplace else (out if isinstance(heatmaps, torch.Tensor) and heatmaps.is_cuda else out.cpu())

    def _region(self, c, sm, image_size, nms_kernel, inplace):
        isz = [float(v) for v in (image_size.tolist() if isinstance(image_size, torch.Tensor) else image_size)]
        return ops.region_bbox_decode(c, sm.to(c.dtype), L.REGION_SH, nms_kernel=nms_kernel,
                                      num_candidates=self.num_candidates, max_num_bbox=self.max_num_bbox,
                                      avg_kernel=self.avg_kernel, image_size=isz, det_thr=self.detection_threshold,
                                      iou_thr=self.iou_threshold, nms_inplace=inplace)

    def candidate_bbox(self, center_maps, size_maps, image_size=(352, 352)):
        """SPheatmapParser.py:58-99 on an already NMS'd centre map -> candidates [B, k, 5] (CPU tensor)."""
        c, _ = _hm(center_maps)
        sm, _ = _hm(size_maps)
        return self._region(c, sm, image_size, 0, False)["candidates"].cpu()

    def non_max_suppression(self, candidates):
        """SPheatmapParser.py:101-138 -> list over images: None or [[x, y, w, h, conf], ...]."""
        c, _ = _up(candidates, torch.float32)
        return _boxes_to_lists(*ops.box_nms(c, self.detection_threshold, self.iou_threshold, self.max_num_bbox))

    def parse(self, heatmaps, center_maps=None, size_maps=None, image_size=(256, 256), scale_factor=1):
        """SPheatmapParser.py:169-206 -> (kpt [B,K,3] on CPU, pred_bboxes).  With centre/size maps the bbox
        branch runs as ONE launch (centre NMS in place on a CUDA centre map, top-k, avg-pool size lookup, box NMS)."""
        pred_bboxes = None
        if center_maps is not None and size_maps is not None:
            c, inplace = _region_inplace(center_maps)
            sm, _ = _hm(size_maps)
            r = self._region(c, sm, image_size, self.nms_kernel, inplace)
            pred_bboxes = _boxes_to_lists(r["boxes"], r["counts"])
        t, _ = _hm(heatmaps)
        H, W = t.shape[2:]
        isz = torch.tensor(image_size, dtype=torch.float32)
        f = (isz / torch.tensor([W, H], dtype=torch.float32)).tolist()
        r = ops.decode_heatmap(t, L.MASK_NONE, L.REFINE_OFFSET, L.XFORM_SCALE, scale_xy=(f[0], f[1]), want_idx=False)
        return r["kpts"].cpu(), pred_bboxes


class HeatmapParser:
    """utils/HeatmapParser.py:197-223 — only adjust_keypoints (list-of-lists form).  The reference
    indexes ``heatmaps[batch, joint_id]`` on an (n_joints+1)-channel tensor whose channel 0 is the
    centre map (an off-by-one quirk); ``channel_offset`` reproduces it (default 0 = as the code)."""

    def __init__(self, channel_offset=0):
        self.channel_offset = channel_offset

    def adjust_keypoints(self, keypoints, heatmaps):
        """keypoints[batch][bbox] = list of [x, y, ...] per joint (grouped candidates, not plane argmaxima): all
        points are refined in ONE launch (lhn_refine_points) and written back into the lists."""
        t, _ = _hm(heatmaps)
        bc, xy, where = [], [], []
        for batch_id, kpt_list in enumerate(keypoints):
            for bbox_id, kpt in enumerate(kpt_list):
                for joint_id, joint in enumerate(kpt):
                    bc.append((batch_id, joint_id + self.channel_offset))
                    xy.append((float(joint[0]), float(joint[1])))
                    where.append((batch_id, bbox_id, joint_id))
        if not xy:
            return keypoints
        out = ops.refine_points(t, torch.tensor(bc, dtype=torch.int32).to(t.device),
                                torch.tensor(xy, dtype=torch.float32).to(t.device), plus_half=False).cpu().tolist()
        for (b, g, j), (x, y) in zip(where, out):
            keypoints[b][g][j][0] = x
            keypoints[b][g][j][1] = y
        return keypoints


# ---- region maps (utils/evaluation.py) ------------------------------------------------------------------
def cs_from_region_map(batch_region_maps, image_size=256, k=20, thr=0.8):
    """utils/evaluation.py:94-138: top-k of the (un-suppressed) centre map, size = mean of the size maps over the
    clipped window around each centre; candidates [B, k, 5] (a CPU tensor, as the reference allocates it)."""
    t, _ = _hm(batch_region_maps)
    r = ops.region_bbox_decode(t[:, 0:1], t[:, 1:3], L.REGION_CS, nms_kernel=0, num_candidates=k, max_num_bbox=1,
                               image_size=(float(image_size), float(image_size)), cand_thr=thr)
    return r["candidates"].cpu()


def non_max_suppression(prediction, iou_threshold=0.8, conf_threshold=0.8, max_num=100):
    """utils/evaluation.py:170-212 -> list over images: None or [[x, y, w, h, conf], ...]."""
    c, _ = _up(prediction, torch.float32)
    return _boxes_to_lists(*ops.box_nms(c, conf_threshold, iou_threshold, max_num))


# ---- HRNet-style helpers (utils/transforms.py) ---------------------------------------------------------
def get_max_preds(batch_heatmaps):
    """utils/transforms.py:47-75 (coords * 0 where max <= 0)."""
    if not isinstance(batch_heatmaps, (np.ndarray, torch.Tensor)):
        raise AssertionError('batch_heatmaps should be numpy.ndarray')
    assert batch_heatmaps.ndim == 4, 'batch_images should be 4-ndim'
    t, was = _hm(batch_heatmaps)
    r = ops.decode_heatmap(t, L.MASK_ZERO, L.REFINE_NONE, want_idx=False)["hm_kpts"]
    return (r[..., :2], r[..., 2:]) if was else (_np(r[..., :2]), _np(r[..., 2:]))


def get_coordinates_from_heatmap(batch_heatmaps):
    """utils/evaluation.py:62-89 (torch in, torch out on the same device)."""
    assert isinstance(batch_heatmaps, torch.Tensor), "batch_heatmaps should be torch.Tensor"
    assert batch_heatmaps.dim() == 4, "batch_heatmaps should be 4-ndim"
    t, was = _hm(batch_heatmaps)
    r = ops.decode_heatmap(t, L.MASK_ZERO, L.REFINE_NONE, want_idx=False)["hm_kpts"]
    preds, maxvals = r[..., :2].contiguous(), r[..., 2:].contiguous()
    return (preds, maxvals) if was else (preds.cpu(), maxvals.cpu())


def get_final_preds(batch_heatmaps, center, scale):
    """utils/transforms.py:18-44 (A3 -> floor(x+0.5) guarded quarter shift -> back-transform).  The
    reference's cv2-affine transform_preds (utils/transforms.py:112-152) is, for rot = 0, the similarity
    X = (x - W/2) f + cx, Y = (y - H/2) f + cy with f = 200 scale[0] / W — scale[1] is never used — which is
    T1 with scale_y := scale[0] * H / W."""
    t, was = _hm(batch_heatmaps)
    c, _ = _up(center, torch.float32)
    s, _ = _up(scale, torch.float32)
    H, W = t.shape[-2:]
    # tensor / tensor: torch turns division by a Python scalar on CUDA into a multiplication by its reciprocal
    s = torch.stack([s[:, 0], (s[:, 0] * float(H)) / torch.full_like(s[:, 0], float(W))], dim=1).contiguous()
    r = ops.decode_heatmap(t, L.MASK_ZERO, L.REFINE_SIGN_ROUND, L.XFORM_CENTER_SCALE, c, s, want_idx=False)
    preds = r["kpts"][..., :2]
    return preds if was else _np(preds)


def flip_back(output_flipped, matched_parts):
    """utils/transforms.py:78-92: reverse W and swap the matched channels."""
    assert output_flipped.ndim == 4, 'output_flipped should be [batch_size, num_joints, height, width]'
    t, was = _up(output_flipped)
    fi = flip_index_from_pairs(t.shape[1], matched_parts, t.device)
    out = ops.flip_back(t, fi)
    return out if was else (_np(out) if isinstance(output_flipped, np.ndarray) else out.cpu())
