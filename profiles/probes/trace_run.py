import ctypes, sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from litehandnet_b200 import fused, synth, _lib as L
dev = torch.device('cuda', 0)
B, K, H, W = 1024, 21, 64, 64
step = fused.FusedHeatmapStep((256, 256), sigma=2, unbiased_encoding=True, balance=True, post_process='unbiased', kernel=11)
sets = []
for r in range(2):
    hm, cen = synth.blob_heatmaps(B, K, H, W, seed=10 * r, device=dev)
    hf = synth.flipped_blob_heatmaps(cen, H, W, seed=10 * r + 1, device=dev)
    j, v = synth.hand_joints(B, K, (256, 256), seed=10 * r + 2, device=dev)
    c, s = synth.bbox_center_scale(B, seed=10 * r + 3, device=dev)
    sets.append(fused.BoundFusedStep(step, hm, j, v, c, s, hm_flip=hf))
for i in range(6):
    sets[i % 2].launch()
torch.cuda.synchronize()
lib = L.lib()
buf = np.zeros(148 * 6 * 16 * 16, dtype=np.int64)
rc = lib.lhn_debug_trace(buf.ctypes.data_as(ctypes.c_void_p))
print('rc', rc)
T = buf.reshape(148 * 6, 16, 16).astype(np.float64)
t = T[:, :, :7]
names = ['S1->wait', 'wait(mbar)', 'sweep', 'reduce+S2', 'resolve+tile+pos', 'S3', 'epilogue(w0)']
its = slice(3, 15)
d = np.diff(t[:, its, :], axis=2)            # [team, it, 6]
for i, n in enumerate(names[1:]):
    x = d[:, :, i]
    print(f'{n:18s} mean {x.mean():8.0f}  p10 {np.percentile(x,10):8.0f}  p50 {np.percentile(x,50):8.0f}  p90 {np.percentile(x,90):8.0f}')
cyc = t[:, 4:15, 0] - t[:, 3:14, 0]
print(f'cycle per plane    mean {cyc.mean():8.0f}  p10 {np.percentile(cyc,10):8.0f} p50 {np.percentile(cyc,50):8.0f} p90 {np.percentile(cyc,90):8.0f}')
gap = t[:, 4:15, 0] - t[:, 3:14, 6]
print(f'epilogue end -> next S1 pass  mean {gap.mean():8.0f} p50 {np.percentile(gap,50):8.0f}')

# finer epilogue stamps (warp 0): 5 = S3 passed, 7 = TMA re-armed, 8 = blur window done, 9 = slow-path check done,
# 10 = log + Taylor done, 11 = transform + keypoint stores done, 6 = loss partials / counters done
order = [5, 7, 8, 9, 10, 11, 6]
lab = ['re-arm TMA', 'row+col blur', 'clamp check', 'log+Taylor', 'transform+stores', 'loss partial']
for a_, b_, n in zip(order[:-1], order[1:], lab):
    x = T[:, its, b_] - T[:, its, a_]
    print(f'  epi {n:18s} mean {x.mean():8.0f}  p50 {np.percentile(x,50):8.0f}  p90 {np.percentile(x,90):8.0f}')
