"""Randomised parity sweep of the fused kernel against the numpy oracle: random shapes (incl. H != W, odd K, stacked
sigmas), dtypes, flip / flip pairs, refinements, mask modes, loss modes, encodings.  Not a test (minutes of oracle
time); run on a GPU box:  python profiles/probes/fuzz_fused.py [n_cases] [seed]"""
import sys

import numpy as np
import torch

sys.path.insert(0, '/root/repo')
from litehandnet_b200 import _lib as L, ops, synth   # noqa: E402
from oracle import np_oracle as O                     # noqa: E402

DEV = 'cuda'
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
cu = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(DEV)
ONLY = int(sys.argv[3]) if len(sys.argv) > 3 else -1      # re-run one case of a sweep with a detailed dump
bad = 0
for case in range(n_cases):
    H, W = [(64, 64), (56, 56), (64, 48), (32, 32), (28, 28), (16, 16), (14, 14), (128, 128), (40, 72)][rng.integers(9)]
    K = int(rng.integers(1, 23))
    N = int(rng.integers(1, 40 if H * W <= 4096 else 6))
    flip = bool(rng.integers(2))
    dt = [torch.float32, torch.float32, torch.bfloat16, torch.float16][rng.integers(4)]
    mode = [L.LOSS_NONE, L.LOSS_DISTANCE_BALANCE, L.LOSS_DISTANCE, L.LOSS_JOINTS_MSE][rng.integers(4)]
    enc = [True, False, 'udp'][rng.integers(3)]
    refine = [L.REFINE_NONE, L.REFINE_OFFSET_HALF, L.REFINE_OFFSET, L.REFINE_SIGN, L.REFINE_SIGN_ROUND, L.REFINE_DARK,
              L.REFINE_DARK_LEGACY, L.REFINE_DARK_UDP][rng.integers(8)]
    if refine == L.REFINE_DARK_UDP and (min(H, W) <= 7 or K == 1):   # K == 1: the reference's .squeeze() breaks too
        refine = L.REFINE_DARK
    mask = [L.MASK_NONE, L.MASK_ZERO, L.MASK_NEG1][rng.integers(3)]
    if refine in (L.REFINE_DARK, L.REFINE_DARK_UDP, L.REFINE_SIGN):
        mask = L.MASK_NEG1
    seed = int(rng.integers(1 << 30))
    sig = float([1.5, 2.0, 3.0][rng.integers(3)]) * W / 64 if W >= 32 else 1.0
    pairs = ((0, 1),) if (flip and K >= 2 and rng.integers(2)) else ()
    if ONLY >= 0 and case != ONLY:
        continue                                         # every rng draw of the case has been consumed
    hm, cen = synth.blob_heatmaps(N, K, H, W, seed=seed, zero_frac=0.05, tie_frac=0.05, sigma=max(1.0, 2.0 * W / 64), dtype=dt)
    hf = synth.flipped_blob_heatmaps(cen, H, W, seed=seed + 1, flip_pairs=pairs, dtype=dt) if flip else None
    c, s = synth.bbox_center_scale(N, seed=seed + 2)
    s = s * torch.tensor([1.0, 0.6 + 0.1 * (seed % 9)])                 # anisotropic boxes
    j, v = synth.hand_joints(N, K, (4 * W, 4 * H), seed=seed + 3, vis_prob=0.9, outside_frac=0.05)
    hm32 = hm.float().numpy()
    hf32 = None if hf is None else hf.float().numpy()
    render = None if mode == L.LOSS_NONE else dict(loss_mode=mode, image_size=(4 * W, 4 * H), sigma=sig, unbiased=enc)
    fi = None
    if pairs:
        from litehandnet_b200.fused import flip_index_from_pairs
        fi = flip_index_from_pairs(K, pairs, DEV)
    try:
        r = ops.decode_heatmap(hm.to(DEV), mask, refine, L.XFORM_CENTER_SCALE, cu(c.numpy()), cu(s.numpy()),
                               hm_flip=None if hf is None else hf.to(DEV), flip_index=fi,
                               use_udp=(refine == L.REFINE_DARK_UDP), render=render, joints=cu(j.numpy()), vis=cu(v.numpy()))
    except L.LhnError as e:
        print(case, 'rejected:', e, (N, K, H, W), dt, refine)
        continue
    torch.cuda.synchronize()
    avg = hm32 if hf32 is None else O.flip_average(hm32, hf32, pairs)
    ok = True
    idx_ref = O.argmax_planes(avg)[0]
    if not np.array_equal(r['idx'].cpu().numpy(), idx_ref):
        ok = False; print(case, 'ARGMAX mismatch')
    mm = {L.MASK_NONE: 'none', L.MASK_ZERO: 'zero', L.MASK_NEG1: 'neg1'}[mask]
    p0, mv, _ = O.max_preds(avg, mm)
    with np.errstate(all='ignore'):
        if refine == L.REFINE_NONE: want = p0
        elif refine == L.REFINE_OFFSET_HALF: want = O.refine_offset_clamped(p0, avg, True)
        elif refine == L.REFINE_OFFSET: want = O.refine_offset_clamped(p0, avg, False)
        elif refine == L.REFINE_SIGN: want = O.refine_sign_guarded(p0, avg)
        elif refine == L.REFINE_SIGN_ROUND: want = O.refine_sign_guarded(p0, avg, round_half=True)
        elif refine == L.REFINE_DARK: want = O.refine_dark(p0, avg, 11, False)
        elif refine == L.REFINE_DARK_LEGACY: want = O.refine_dark(p0, avg, 19, True)
        else: want = O.post_dark_udp(p0, avg.copy(), 11)
    got = r['hm_kpts'].cpu().numpy()[..., :2]
    err = np.abs(got - want)
    tol = 1e-5 * np.maximum(np.abs(want), 1.0) + 2e-5
    if refine in (L.REFINE_DARK, L.REFINE_DARK_LEGACY, L.REFINE_DARK_UDP):
        tol = tol * 50        # ill-conditioned Hessians on noise planes amplify 1-ulp differences of log()
    if not np.all((err <= tol) | ~np.isfinite(want)):
        ok = False; print(case, 'COORD mismatch', float(np.nanmax(err)), (N, K, H, W), dt, refine, mm)
        if ONLY >= 0:
            w = np.argwhere(~((err <= tol) | ~np.isfinite(want)))
            for b_, k_, _ in w[:4]:
                print(' plane', b_, k_, 'got', got[b_, k_], 'want', want[b_, k_], 'p0', p0[b_, k_], 'max', mv[b_, k_])
                np.save('/root/repo/gpurun_out/fuzz_plane.npy', avg[b_, k_])
    # image-space coordinates: transform_preds of the oracle's heatmap coordinates (T1, both scale components)
    gi = r['kpts'].cpu().numpy()[..., :2]
    wi = np.stack([O.transform_preds(want[i], c.numpy()[i], s.numpy()[i], [W, H], use_udp=(refine == L.REFINE_DARK_UDP))[:, :2]
                   for i in range(N)])
    ei = np.abs(gi - wi)
    toli = (1e-5 * np.maximum(np.abs(wi), 1.0) + 2e-4) * (50 if refine in (L.REFINE_DARK, L.REFINE_DARK_LEGACY, L.REFINE_DARK_UDP) else 1)
    if not np.all((ei <= toli) | ~np.isfinite(wi)):
        ok = False; print(case, 'IMAGE-COORD mismatch', float(np.nanmax(ei)), (N, K, H, W), dt, refine, mm)
    if render is not None:
        tg, tw = O.render_targets(j.numpy(), v.numpy(), (4 * W, 4 * H), (W, H), sig, enc)
        if mode == L.LOSS_JOINTS_MSE: wl = O.joints_distance_loss_mse(hm32, tg, tw)
        else: wl = O.distance_loss_l2(hm32, tg, tw, balance=(mode == L.LOSS_DISTANCE_BALANCE))
        gl = ops.loss_finalize(ops.loss_reduce(r['partials']), mode).item()
        if not np.array_equal(r['weight'].cpu().numpy(), tw.reshape(N, K)):
            ok = False; print(case, 'WEIGHT mismatch')
        if abs(gl - float(wl)) > 1e-5 * abs(float(wl)) + 1e-12:
            ok = False; print(case, 'LOSS mismatch', gl, float(wl), (N, K, H, W), dt, mode, enc, sig)
    bad += (not ok)
print(f'{n_cases} cases, {bad} with mismatches')
