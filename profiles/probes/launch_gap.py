"""How much of a step is launch gap?  K1 launched back to back: eager, one graph per step, one graph of all steps."""
import sys, torch
sys.path.insert(0, '/root/repo')
from litehandnet_b200 import fused, synth, _lib as L
dev = torch.device('cuda', 0)
B, K, H, W = 1024, 21, 64, 64
OVERLAP = len(sys.argv) > 1 and sys.argv[1] == 'overlap'
print('overlap_previous =', OVERLAP)
step = fused.FusedHeatmapStep((256, 256), sigma=2, unbiased_encoding=True, balance=True, post_process='unbiased', kernel=11)
bound = []
for r in range(2):
    hm, cen = synth.blob_heatmaps(B, K, H, W, seed=10 * r, device=dev)
    hf = synth.flipped_blob_heatmaps(cen, H, W, seed=10 * r + 1, device=dev)
    j, v = synth.hand_joints(B, K, (256, 256), seed=10 * r + 2, device=dev)
    c, s = synth.bbox_center_scale(B, seed=10 * r + 3, device=dev)
    bound.append(fused.BoundFusedStep(step, hm, j, v, c, s, hm_flip=hf, overlap_previous=OVERLAP))
N = 50
def timed(fn, n=N):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
st = L.stream()
print('eager back-to-back      : %.1f us/launch' % timed(lambda: [bound[i % 2].launch_kernel(st) for i in range(N)]))
for b in bound: b.capture()
print('one graph per step      : %.1f us/launch' % timed(lambda: [bound[i % 2].replay() for i in range(N)]))
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for i in range(N):
        bound[i % 2].launch_kernel(L.stream())
print('one graph of %d launches : %.1f us/launch' % (N, timed(g.replay)))

# results must not depend on the launch mode
torch.cuda.synchronize()
print('loss', [float(b.loss.item()) for b in bound], 'sums', bound[0].sums.tolist())
