"""Seeded synthetic inputs for the parity tests and bench.py.

There is no dataset on the GPU box, so every workload is synthetic: Gaussian-blob heatmaps with
uniform noise, joints distributed like FreiHAND keypoints, bbox centre/scale as the top-down
pipeline produces them.  All generators are plain torch ops on the requested device with an
explicit generator, so the CPU tests and the oracle see bit-identical inputs.
"""
import torch

MPII_FLIP_PAIRS = ((0, 5), (1, 4), (2, 3), (10, 15), (11, 14), (12, 13))  # dataset_configs/mpii.py:13-112


def _gen(seed, device):
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    return g


def blob_heatmaps(B, K, H, W, seed=0, device="cpu", sigma=2.0, noise=0.02, margin=4.0,
                  zero_frac=0.0, tie_frac=0.0, dtype=torch.float32, centers=None):
    """[B,K,H,W] heatmaps: amplitude U(0.3,1) * exp(-r^2/(2 sigma^2)) at a uniform sub-pixel centre
    in [margin, size-margin) plus U(0, noise).  `zero_frac` of the planes are all-zero and
    `tie_frac` carry an exact two-way tie of the maximum (parity cases).  Returns (hm, centers[B,K,2])."""
    g = _gen(seed, device)
    if centers is None:
        cx = torch.rand(B, K, generator=g, device=device) * (W - 2 * margin) + margin
        cy = torch.rand(B, K, generator=g, device=device) * (H - 2 * margin) + margin
    else:
        cx, cy = centers[..., 0].to(device), centers[..., 1].to(device)
    amp = torch.rand(B, K, generator=g, device=device) * 0.7 + 0.3
    xs = torch.arange(W, device=device, dtype=torch.float32)
    ys = torch.arange(H, device=device, dtype=torch.float32)
    ex = torch.exp(-(xs[None, None, :] - cx[..., None]) ** 2 / (2 * sigma * sigma))   # [B,K,W]
    ey = torch.exp(-(ys[None, None, :] - cy[..., None]) ** 2 / (2 * sigma * sigma))   # [B,K,H]
    hm = amp[..., None, None] * ey[..., :, None] * ex[..., None, :]
    hm = hm + torch.rand(B, K, H, W, generator=g, device=device) * noise
    if zero_frac > 0:
        z = torch.rand(B, K, generator=g, device=device) < zero_frac
        hm[z] = 0.0
    if tie_frac > 0:
        t = torch.rand(B, K, generator=g, device=device) < tie_frac
        flat = hm.view(B, K, -1)
        m, i = flat.max(dim=2)
        j = (i + 1 + (torch.rand(B, K, generator=g, device=device) * (H * W - 2)).long()) % (H * W)
        sel = t.nonzero(as_tuple=False)
        if sel.numel():
            flat[sel[:, 0], sel[:, 1], j[sel[:, 0], sel[:, 1]]] = m[sel[:, 0], sel[:, 1]]
    return hm.to(dtype), torch.stack([cx, cy], dim=-1)


def flipped_blob_heatmaps(centers, H, W, seed=1, device="cpu", flip_pairs=(), **kw):
    """The network's output on the mirrored image: the blob of joint flip_index[k] mirrored along W
    (so that flip_back() lands it on joint k), with independent noise."""
    B, K = centers.shape[:2]
    idx = list(range(K))
    for a, b in flip_pairs:
        idx[a], idx[b] = idx[b], idx[a]
    c = centers[:, idx].clone()
    c[..., 0] = (W - 1) - c[..., 0]
    hm, _ = blob_heatmaps(B, K, H, W, seed=seed, device=device, centers=c, **kw)
    return hm


def hand_joints(B, K, image_size=(256, 256), seed=2, device="cpu", vis_prob=0.95, spread=0.18,
                outside_frac=0.01):
    """joints_3d [B,K,3] f32 (x, y, 0) in image pixels and joints_3d_visible [B,K,3] f32:
    a per-sample hand centre plus per-joint jitter (the FreiHAND crops of split_testset/freihand_20
    have this shape: 21 keypoints inside a bbox covering ~1/3 of the crop); `outside_frac` of the
    joints lie up to 40 px outside the image (visibility-rule edge case)."""
    g = _gen(seed, device)
    Wi, Hi = image_size
    c = torch.rand(B, 1, 2, generator=g, device=device) * 0.4 + 0.3
    j = c + torch.randn(B, K, 2, generator=g, device=device) * spread
    j = j * torch.tensor([Wi, Hi], device=device, dtype=torch.float32)
    out = torch.rand(B, K, generator=g, device=device) < outside_frac
    far = (torch.rand(B, K, 2, generator=g, device=device) * 2 - 1) * 40.0
    far = torch.where(far < 0, far, far + torch.tensor([Wi, Hi], device=device, dtype=torch.float32))
    j = torch.where(out[..., None], far, j)
    joints = torch.zeros(B, K, 3, device=device)
    joints[..., :2] = j
    v = (torch.rand(B, K, generator=g, device=device) < vis_prob).float()
    vis = torch.stack([v, v, torch.zeros_like(v)], dim=-1)
    return joints, vis


def bbox_center_scale(B, seed=3, device="cpu", fixed=False):
    """center [B,2] f32, scale [B,2] f32 (= bbox w,h / 200).  fixed=True gives the identity crop of
    a 256x256 image (center 128, scale 1.28) used by BASELINE configs 1 and 3."""
    if fixed:
        return (torch.full((B, 2), 128.0, device=device), torch.full((B, 2), 1.28, device=device))
    g = _gen(seed, device)
    center = torch.rand(B, 2, generator=g, device=device) * 160 + 48
    s = torch.rand(B, 1, generator=g, device=device) * 1.2 + 0.5
    return center, s.expand(B, 2).contiguous()


def simdr_vectors(B, K, L, seed=4, device="cpu", k=2, noise=0.02):
    """x/y SimDR vectors [B,K,L]: 1-D Gaussian sigma=2k at U(0,L) + U(0,noise)."""
    g = _gen(seed, device)
    out = []
    for _ in range(2):
        mu = torch.rand(B, K, 1, generator=g, device=device) * L
        pos = torch.arange(L, device=device, dtype=torch.float32)
        v = torch.exp(-(pos - mu) ** 2 / (2 * (2.0 * k) ** 2))
        out.append(v + torch.rand(B, K, L, generator=g, device=device) * noise)
    return out[0], out[1]


def pck_inputs(centers_hm, stride=4.0, seed=5, device="cpu", mask_prob=0.9, gt_noise=6.0):
    """gt [B,K,2] = blob centre * stride + N(0, gt_noise px); mask Bernoulli; bbox w,h U(60,200)."""
    g = _gen(seed, device)
    B, K = centers_hm.shape[:2]
    gt = centers_hm.to(device) * stride + torch.randn(B, K, 2, generator=g, device=device) * gt_noise
    mask = torch.rand(B, K, generator=g, device=device) < mask_prob
    wh = torch.rand(B, 2, generator=g, device=device) * 140 + 60
    return gt.float(), mask, wh.float()
