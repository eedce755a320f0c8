"""In-kernel timing of the NVLink exchange (torchrun, one process per GPU): runs config-4-like steps with the per-step
counter exchange and prints, per rank, the last launches' stamps kept in the mailbox control page:
publish = t_published - t_enter (bulk copy of the block to every peer + system fence + flags), wait = the time the
consumer (two launches later) spent polling for the peers' flags of that step.
    python -m torch.distributed.run --nproc-per-node 2 profiles/probes/xch_timing.py"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from litehandnet_b200 import _lib as L  # noqa: E402
from litehandnet_b200 import fused, synth  # noqa: E402
from litehandnet_b200.dist import PeerExchange  # noqa: E402


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    x = PeerExchange(dev)
    K, T, B, R = 16, 20, 1024, 2
    sets = []
    for i in range(R):
        hm, cen = synth.blob_heatmaps(B, K, 64, 64, seed=100 * rank + i, device=dev)
        c, s = synth.bbox_center_scale(B, seed=100 * rank + i + 1, device=dev)
        gt, mask, wh = synth.pck_inputs(cen, seed=100 * rank + i + 2, device=dev)
        sets.append((hm, c, s, gt, mask, wh))
    totals = torch.zeros((T + 5) * K, dtype=torch.int64, device=dev)
    bound = [fused.BoundDecodeStep(s[0], s[1], s[2], L.MASK_NEG1, L.REFINE_SIGN, L.XFORM_CENTER_SCALE, overlap_previous=True,
                                   metrics=dict(gt=s[3], mask=s[4], bbox_wh=s[5], auc_steps=T, exchange=x, totals=totals))
             for i, s in enumerate(sets)]
    for step in range(10):
        bound[step % R].launch()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 60
    import time
    e0.record()
    th = time.perf_counter()
    for step in range(n):
        bound[step % R].launch()
    host_us = (time.perf_counter() - th) / n * 1e6
    e1.record()
    x.flush()
    torch.cuda.synchronize()
    ctrl = x.mailbox[L.XCH_SLOTS * L.XCH_MAX_RANKS * L.XCH_PAYLOAD_BYTES + 2048:][:4 * 8 * 8].view(torch.int64).cpu().view(4, 8)
    rows = sorted(ctrl.tolist(), key=lambda r: r[0])
    out = [f"rank {rank}: {e0.elapsed_time(e1) / n * 1e3:.1f} us per step over {n} steps, host time per launch {host_us:.1f} us ({x.how})"]
    prev = None
    for r in rows:
        gap = "" if prev is None else f" since previous launch's arrival {(r[1] - prev) / 1e3:7.1f} us"
        out.append(f"  seq {r[0]:4d}: publish {(r[2] - r[1]) / 1e3:6.1f} us, consumer waited for peers {(r[3] - r[4]) / 1e3:7.1f} us, "
                   f"publish -> consumed {(r[3] - r[2]) / 1e3:7.1f} us{gap}")
        prev = r[3]
    for rk in range(world):
        if rk == rank:
            print("\n".join(out), flush=True)
        dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
