import ctypes, sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from litehandnet_b200 import fused, synth, _lib as L
dev = torch.device('cuda', 0)
MODE = sys.argv[1] if len(sys.argv) > 1 else 'headline'      # headline | noflip | bf16 | p56 | cfg5 (last two: no flip)
B, K, H, W = {'p56': (1024, 21, 56, 56), 'cfg5': (256, 21, 128, 128)}.get(MODE, (1024, 21, 64, 64))
DT = torch.bfloat16 if MODE == 'bf16' else torch.float32
print('mode', MODE)
step = fused.FusedHeatmapStep((4 * W, 4 * H), sigma=2.0 * W / 64, unbiased_encoding=True, balance=True, post_process='unbiased', kernel=11)
sets = []
for r in range(2):
    hm, cen = synth.blob_heatmaps(B, K, H, W, seed=10 * r, device=dev, dtype=DT)
    hf = None if MODE in ('noflip', 'p56', 'cfg5') else synth.flipped_blob_heatmaps(cen, H, W, seed=10 * r + 1, device=dev, dtype=DT)
    j, v = synth.hand_joints(B, K, (4 * W, 4 * H), seed=10 * r + 2, device=dev)
    c, s = synth.bbox_center_scale(B, seed=10 * r + 3, device=dev)
    sets.append(fused.BoundFusedStep(step, hm, j, v, c, s, hm_flip=hf))
for i in range(6):
    sets[i % 2].launch()
torch.cuda.synchronize()
lib = L.lib()
buf = np.zeros(148 * 12 * 16 * 16, dtype=np.int64)
rc = lib.lhn_debug_trace(buf.ctypes.data_as(ctypes.c_void_p))
print('rc', rc)
T = buf.reshape(148 * 12, 16, 16).astype(np.float64)
T = T[T[:, 0, 15] > 0]                              # teams that ran
print('teams traced', T.shape[0], '=', T.shape[0] // 148, 'per SM')
its = slice(3, 15)
def stat(name, x):
    print(f'{name:34s} mean {x.mean():8.0f}  p10 {np.percentile(x,10):8.0f}  p50 {np.percentile(x,50):8.0f}  p90 {np.percentile(x,90):8.0f}')
# sweepers (first sweeper warp, lane 0): 0 = S1 passed, 1 = plane landed, 2 = sweep done, 3 = S2 passed,
# 4 = row sums / positives / record staged, 5 = S3 passed + stage re-armed
for a_, b_, n in [(0, 1, 'sweeper wait for plane (mbar)'), (1, 2, 'sweep'), (2, 3, 'warp reduce + S2'),
                  (3, 4, 'resolve + window + positives'), (3, 6, '  resolve argmax'), (6, 7, '  window staging'),
                  (7, 4, '  positives + record'), (4, 5, 'S3 + re-arm TMA')]:
    stat(n, T[:, its, b_] - T[:, its, a_])
stat('wait for tables/empty (next plane)', T[:, 4:15, 0] - T[:, 3:14, 5])
stat('sweeper cycle per plane', T[:, 4:15, 0] - T[:, 3:14, 0])
# epilogue warp: 8 = record received, 9 = blur / clamp check / log done, 10 = Taylor done, 11 = plane finished,
# 12 = side ring + tables of plane n+2 done
for a_, b_, n in [(8, 9, 'epilogue: blur + log'), (9, 10, 'epilogue: Taylor'), (10, 11, 'epilogue: transform + stores + loss'),
                  (11, 12, 'epilogue: side ring + tables n+2')]:
    stat(n, T[:, its, b_] - T[:, its, a_])
stat('epilogue busy per plane', T[:, its, 12] - T[:, its, 8])
stat('epilogue wait for record', T[:, 4:15, 8] - T[:, 3:14, 12])

# whole-kernel view per team (clock64 is per SM: only differences within a team are meaningful)
entry, finish = T[:, 0, 15], T[:, 0, 14]
stat('team: entry -> first plane landed', T[:, 0, 1] - entry)
stat('team: entry -> tables ready (it 0)', T[:, 0, 0] - entry)
stat('team: first 3 planes (entry -> it3)', T[:, 3, 0] - entry)
stat('team: total (entry -> finish)', finish - entry)
per_sm = (finish - entry).reshape(148, -1)
stat('SM: slowest team total', per_sm.max(axis=1))
print('kernel cycles if every SM ran at its slowest team:', per_sm.max())

# wall-clock view (globaltimer, ns): SM clock actually seen by the kernel and its span across the chip
g0, g1 = T[:, 0, 13], T[:, 1, 13]
mhz = (finish - entry) / (g1 - g0) * 1e3
stat('SM clock seen by the kernel (MHz)', mhz)
print(f'first entry -> last finish: {(g1.max() - g0.min()) / 1e3:.1f} us; entry skew {(g0.max() - g0.min()) / 1e3:.1f} us; '
      f'finish skew {(g1.max() - g1.min()) / 1e3:.1f} us')
