"""SimDRLoss heads: the fused tcgen05 kernel (lhn_simdr_heads_loss) against the reference's structure on the same box
(two torch.nn.Linear = cuBLAS fp32 SGEMM, then the SmoothL1 reduction lhn_simdr_smoothl1).  CUDA events, 30 iterations
after 5 warm-ups.  Prints one JSON object per shape; `frac_bf16_sustained` = executed tensor flops (3 bf16 MMAs per
product) / time / MEASURED_PEAKS bf16_tflops_sustained."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from litehandnet_b200 import ops, synth  # noqa: E402


def timed(fn, iters=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3      # us


def main():
    dev = "cuda"
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]
    except Exception:
        peak = 1387.3
    torch.backends.cuda.matmul.allow_tf32 = False           # torch's default: the reference's nn.Linear runs in fp32
    shapes = ((64, 21, 4096, 512, 512), (32, 21, 4096, 512, 512), (256, 21, 4096, 512, 512), (64, 21, 3136, 448, 448))
    if "--only" in sys.argv:
        shapes = (shapes[int(sys.argv[sys.argv.index("--only") + 1])],)
    for B, K, HW, Lx, Ly in shapes:
        side = int(round(HW ** 0.5))
        hm, _ = synth.blob_heatmaps(B, K, side, side, seed=1, device=dev)
        lin_x = torch.nn.Linear(HW, Lx).to(dev)
        lin_y = torch.nn.Linear(HW, Ly).to(dev)
        j, v = synth.hand_joints(B, K, (Lx // 2, Ly // 2), seed=2, device=dev)
        tx, ty = ops.render_simdr(j, v, (Lx // 2, Ly // 2), 2, 2)
        w = v[..., :1].contiguous()
        wcat = torch.cat([lin_x.weight.detach(), lin_y.weight.detach()]).contiguous()
        bias = torch.cat([lin_x.bias.detach(), lin_y.bias.detach()]).contiguous()
        split = ops.split_bf16(wcat)

        def unfused():
            with torch.no_grad():
                a = hm.flatten(2)
                return ops.simdr_smoothl1(lin_x(a), lin_y(a), tx, ty, w)

        def unfused_tf32():
            torch.backends.cuda.matmul.allow_tf32 = True
            try:
                return unfused()
            finally:
                torch.backends.cuda.matmul.allow_tf32 = False

        def fused():
            return ops.simdr_heads_loss(hm, split, bias, tx, ty, w)[0]

        l0, l1 = float(unfused().item()), float(fused().item())
        t_un, t_tf32, t_fu = timed(unfused), timed(unfused_tf32), timed(fused)
        t_split = timed(lambda: ops.split_bf16(hm.reshape(B * K, HW)))

        def graphed(fn):
            """device time per call: the call captured once into a CUDA graph and replayed (no host work between kernels)"""
            fn()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
            return timed(g.replay)

        t_un_dev, t_fu_dev, t_split_dev = graphed(unfused), graphed(fused), graphed(lambda: ops.split_bf16(hm.reshape(B * K, HW)))
        M, N = B * K, Lx + Ly
        flops = 2.0 * M * N * HW
        print(json.dumps({
            "shape": f"B={B} K={K} HW={HW} N={N}", "loss_cublas_fp32": l0, "loss_fused": l1, "rel": abs(l0 - l1) / abs(l0),
            "us_cublas_fp32_plus_smoothl1": t_un, "us_cublas_tf32_plus_smoothl1": t_tf32, "us_fused_total": t_fu,
            "us_of_which_split_of_A": t_split, "speedup_vs_fp32": t_un / t_fu,
            "device_us_cublas_fp32_plus_smoothl1": t_un_dev, "device_us_fused_total": t_fu_dev, "device_us_split_of_A": t_split_dev,
            "device_speedup_vs_fp32": t_un_dev / t_fu_dev,
            "device_frac_bf16_sustained": 3 * flops / ((t_fu_dev - t_split_dev) * 1e-6) / 1e12 / peak,
            "useful_tflops_fused": flops / (t_fu * 1e-6) / 1e12,
            "executed_tflops_fused_kernel": 3 * flops / ((t_fu - t_split) * 1e-6) / 1e12,
            "frac_bf16_sustained": 3 * flops / ((t_fu - t_split) * 1e-6) / 1e12 / peak,
            "pred_bytes_not_written_and_read": 2 * M * N * 4}), flush=True)


if __name__ == "__main__":
    main()
