"""Worker of tests/test_gpu_exchange.py::test_exchange_across_gpus_under_torchrun (one process per GPU, NCCL group):
the in-kernel exchange over real peer mappings against NCCL all-reduces of the same blocks."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from litehandnet_b200 import _lib as L  # noqa: E402
from litehandnet_b200 import fused, metrics as M, synth  # noqa: E402
from litehandnet_b200.dist import PeerExchange  # noqa: E402


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    x = PeerExchange(dev)
    K, T, B, R, steps = 16, 20, 512, 2, 10
    sets = []
    for i in range(R):
        hm, cen = synth.blob_heatmaps(B, K, 64, 64, seed=100 * rank + i, device=dev, zero_frac=0.02)
        c, s = synth.bbox_center_scale(B, seed=100 * rank + i + 1, device=dev)
        gt, mask, wh = synth.pck_inputs(cen, seed=100 * rank + i + 2, device=dev)
        sets.append((hm, c, s, gt, mask, wh))
    totals = torch.zeros((T + 5) * K, dtype=torch.int64, device=dev)
    bound = [fused.BoundDecodeStep(s[0], s[1], s[2], L.MASK_NEG1, L.REFINE_SIGN, L.XFORM_CENTER_SCALE, overlap_previous=True,
                                   metrics=dict(gt=s[3], mask=s[4], bbox_wh=s[5], auc_steps=T, exchange=x, totals=totals))
             for i, s in enumerate(sets)]
    for step in range(steps):
        bound[step % R].launch()
    x.flush()                                    # the exchange runs one launch behind
    torch.cuda.synchronize()
    assert int(x.status.item()) == 0, "timed out waiting for a peer"
    want = torch.zeros_like(totals)
    for step in range(steps):
        s = sets[step % R]
        acc = M.MetricAccumulator(K, device=dev)
        acc.update_from_heatmaps(s[0], s[1], s[2], s[3], s[4], s[5], post_process="default")
        want += acc.counters
    dist.all_reduce(want)
    assert torch.equal(totals, want), f"rank {rank}: in-kernel totals differ from the NCCL all-reduce"

    # loss sums: global-batch loss on every rank, identical bits
    cfg = fused.FusedHeatmapStep((256, 256), sigma=2, post_process="unbiased", kernel=11)
    hm, cen = synth.blob_heatmaps(256, 21, 64, 64, seed=7 + rank, device=dev)
    hf = synth.flipped_blob_heatmaps(cen, 64, 64, seed=8 + rank, device=dev)
    j, v = synth.hand_joints(256, 21, seed=9 + rank, device=dev)
    c, s = synth.bbox_center_scale(256, seed=10 + rank, device=dev)
    local_step = fused.BoundFusedStep(cfg, hm, j, v, c, s, hm_flip=hf, finalize=False)
    local_step.launch()
    sums = local_step.sums.clone()
    dist.all_reduce(sums)
    x2 = PeerExchange(dev)                       # the loss sums exchange immediately (its own mailbox and step numbers)
    xb = fused.BoundFusedStep(cfg, hm, j, v, c, s, hm_flip=hf, exchange=x2)
    for _ in range(3):
        xb.launch()
    torch.cuda.synchronize()
    assert int(x2.status.item()) == 0
    assert torch.allclose(xb.sums, sums, rtol=1e-12), (xb.sums, sums)
    gathered = [torch.zeros_like(xb.sums) for _ in range(world)]
    dist.all_gather(gathered, xb.sums)
    assert all(torch.equal(g, gathered[0]) for g in gathered), "ranks disagree on the global sums"
    dist.barrier()
    if rank == 0:
        print(f"exchange ok ({x.how}; world {world})")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
