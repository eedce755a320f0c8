#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native heatmap hot path.

Metric (BASELINE.json): heatmap render+loss+decode samples/s @ 21x64x64 (+ % of HBM roofline).
Workload at every N: BASELINE config[1] — FreiHAND 21x64x64 fused Gaussian target render +
target-weight MSE loss (DistanceLoss L2, balance=True) + flip-test average + DARK decode (k=11) +
affine back-transform, batch 1024 PER GPU (weak scaling: the batch is sharded by rank, the only
collective is an all-reduce of the four f64 loss sums).  A step is ONE launch of the persistent fused kernel
(lhn_fused_render_loss_decode: it also reduces and finalises the loss); at N > 1 the kernel leaves the f64 sums,
NCCL all-reduces them and lhn_loss_finalize runs as a second, tiny launch.

  python bench.py --gpus N --steps K --warmup W           (torchrun launches N ranks for N > 1)
  python bench.py --impl reference ...                    the reference's CPU path (oracle port) on the
                                                          box's host cores, same metric and config

One JSON line on stdout (rank 0).  `value` = device-resident throughput (CUDA events, max over ranks);
`e2e` = the same step through the public host-buffer API (pinned host inputs, H2D + kernels + D2H);
`roofline` = the fused kernel's algorithmic bytes / its event-timed duration vs the measured HBM peak;
`cpu_baseline` = the oracle port timed on this box's host cores on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "heatmap render+loss+decode samples/s @21x64x64"
UNIT = "samples/s"
K_JOINTS, H, W = 21, 64, 64
IMAGE_SIZE = (256, 256)
FALLBACK_HBM_GBS = 6650.0            # /opt/skills/guides/B200_PROFILING.md fallback


def workload_name(batch):
    return (f"FreiHAND 21x64x64 fused Gaussian target render + target-weight MSE loss + DARK decode "
            f"with flip-test, batch {batch} per GPU (BASELINE configs[1])")


def algorithmic_bytes_per_sample(flip=True, esz=4):
    """every heatmap element read once (x2 with the flip plane) + O(K) side data:
    joints (x,y) + visibility, center/scale, and the per-plane outputs (2x[3] f32 keypoints, idx,
    weight, 4 f64 loss partials)."""
    planes = K_JOINTS * H * W * esz * (2 if flip else 1)
    side_in = K_JOINTS * 3 * 4 + 16
    side_out = K_JOINTS * (12 + 12 + 4 + 4 + 32)
    return planes + side_in + side_out


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


def recorded_traffic(batch):
    """dram bytes per launch of the fused kernel from the committed ncu --set full capture, if the
    capture was taken on this workload (profiles/traffic.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        if int(t.get("batch", -1)) == int(batch):
            return float(t["dram_bytes_per_launch"])
    except Exception:
        pass
    return None


class ClockSampler:
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, index, interval_s=0.0005):
        self.interval_s = interval_s
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for n, bit in names.items():
                    if r & bit:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(self.interval_s)

    def start(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        if self._thread is not None:
            self._stop.set()
            self._thread.join()
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def physical_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


# ---- CPU arm -----------------------------------------------------------------------------------------
def cpu_inputs(n, seed=0):
    from litehandnet_b200 import synth
    hm, cen = synth.blob_heatmaps(n, K_JOINTS, H, W, seed=seed)
    hf = synth.flipped_blob_heatmaps(cen, H, W, seed=seed + 1)
    joints, vis = synth.hand_joints(n, K_JOINTS, IMAGE_SIZE, seed=seed + 2)
    center, scale = synth.bbox_center_scale(n, seed=seed + 3)
    return [t.numpy() for t in (hm, hf, joints, vis, center, scale)]


def cpu_sample_size(runner, requested, budget_s):
    """Bounded sample: a short probe gives the CPU rate; the sample is sized to ~budget_s per pass."""
    if requested:
        return min(int(requested), runner.B)
    probe = min(runner.B, max(16, 2 * runner.workers))
    runner.run(probe)                                     # warm the workers
    _, _, dt = runner.run(probe)
    rate = probe / max(dt, 1e-6)
    return int(max(16, min(runner.B, 0.5 * rate * budget_s)))


def run_reference_arm(args, rank, world):
    """The reference's CPU path (numpy oracle port, all host cores); rank 0 only."""
    if rank != 0:
        return
    from oracle import cpu_path
    runner = cpu_path.FusedCpuRunner(*cpu_inputs(1024), image_size=IMAGE_SIZE, sigma=2, kernel=11)
    # keep the whole run within ~2.5 minutes: (steps + warmup) passes of `sample` samples each
    budget = max(0.05, min(2.0, 150.0 / max(1, args.steps + args.warmup)))
    sample = cpu_sample_size(runner, args.cpu_sample, budget)
    for _ in range(args.warmup):
        runner.run(sample)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        runner.run(sample)
    dt = time.perf_counter() - t0
    runner.close()
    value = sample * args.steps / dt
    used = runner.workers
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(args.batch), "sample_per_step": sample,
                   "note": "each step is a bounded sample of the workload on the host CPU"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": "port",
                         "sample": f"{sample} samples/step of the batch-{args.batch} workload; numpy oracle "
                                   f"port of the reference pipeline in {used} forked workers"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def parity_block(bound, sets, epoch_local, steps, R):
    """Outputs of the MEASURED path (the bound steps' buffers as the timed loop left them) against the reference's
    CPU pipeline (oracle.cpu_path.FusedCpuRunner: TopDownGenerateTarget -> DistanceLoss(balance) -> flip average ->
    keypoints_from_heatmaps('unbiased')) on every sample of every rotating set: argmax indices bit-exact,
    coordinates |a-b| <= 2e-5 + 1e-5 |b| element-wise, loss 1e-5 relative (north_star)."""
    import numpy as np
    from oracle import cpu_path
    idx_equal, coord_ok, max_rel, loss_rel, n = True, True, 0.0, 0.0, 0
    ref_losses = []
    for r in range(R):
        b = bound[r]
        runner = cpu_path.FusedCpuRunner(*[t.cpu().numpy() for t in sets[r]], image_size=IMAGE_SIZE, sigma=2, kernel=11)
        try:
            with np.errstate(all="ignore"):
                preds, loss, _ = runner.run()
            ridx = runner.last_idx
        finally:
            runner.close()
        ref_losses.append(float(loss))
        got = b.preds.cpu().numpy()
        idx_equal &= bool(np.array_equal(b.idx.cpu().numpy(), ridx))
        d = np.abs(got[..., :2].astype(np.float64) - preds[..., :2])
        coord_ok &= bool((d <= 2e-5 + 1e-5 * np.abs(preds[..., :2])).all())
        coord_ok &= bool(np.array_equal(got[..., 2], preds[..., 2], equal_nan=True))
        max_rel = max(max_rel, float((d / np.maximum(np.abs(preds[..., :2]), 1.0)).max()))
        n += preds.shape[0]
        if epoch_local is None:
            loss_rel = max(loss_rel, abs(float(b.loss.item()) - float(loss)) / abs(float(loss)))
    if epoch_local is not None:       # accumulated over the timed steps: sum of the per-step reference losses
        want = sum(ref_losses[i % R] for i in range(steps))
        loss_rel = abs(epoch_local - want) / abs(want)
    ok = idx_equal and coord_ok and loss_rel <= (1e-5 if epoch_local is None else 5e-5)
    return {"ok": ok, "idx_equal": idx_equal, "coord_max_rel": max_rel, "coord_within_1e-5": coord_ok,
            "loss_rel": loss_rel, "samples": n,
            "against": "oracle.cpu_path.FusedCpuRunner (reference CPU pipeline restated) on every sample of the "
                       "rotating sets, outputs as left by the timed loop, rank 0"}


# ---- GPU arm -----------------------------------------------------------------------------------------
def run_gpu_arm(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from litehandnet_b200 import fused, synth

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    step_cfg = fused.FusedHeatmapStep(IMAGE_SIZE, sigma=2, unbiased_encoding=True, balance=True,
                                      post_process="unbiased", kernel=11)

    # ---- synthetic inputs, device resident; R rotating sets so no step re-reads L2-resident data --
    R = args.rotate
    sets = []
    for r in range(R):
        seed = 1000 * rank + 10 * r
        hm, cen = synth.blob_heatmaps(B, K_JOINTS, H, W, seed=seed, device=dev)
        hf = synth.flipped_blob_heatmaps(cen, H, W, seed=seed + 1, device=dev)
        joints, vis = synth.hand_joints(B, K_JOINTS, IMAGE_SIZE, seed=seed + 2, device=dev)
        center, scale = synth.bbox_center_scale(B, seed=seed + 3, device=dev)
        sets.append((hm, hf, joints, vis, center, scale))
    # N > 1, default: the reference's DDP semantics — every rank finalises the loss of its own shard (local N_pos,
    # loss/heatmapLoss.py:253-258 sees only the rank's batch) inside the one-launch kernel, adds it into a
    # device-resident epoch sum (train_one_epoch's loss_dict['sum'] += v) and the ranks all-reduce that scalar
    # ONCE at the end of the timed region (train/distributed_utils.py:65-76 reduce_value).  --global-loss keeps
    # global N_pos instead: the kernel leaves the f64 sums, NCCL all-reduces them every step (pipelined behind the
    # next step's kernel) and lhn_loss_finalize runs as a second launch.
    global_loss = world > 1 and args.global_loss
    epoch_loss = torch.zeros(1, dtype=torch.float32, device=dev) if (world > 1 and not global_loss) else None
    # consecutive steps work on disjoint buffer sets (R >= 2), so each launch may overlap the tail of the previous one
    bound = [fused.BoundFusedStep(step_cfg, s[0], s[2], s[3], s[4], s[5], hm_flip=s[1], finalize=not global_loss,
                                  overlap_previous=(R >= 2 and not args.no_overlap),
                                  spare_sms=(args.spare_sms if global_loss else 0), accumulate_into=epoch_loss)
             for s in sets]
    use_graph = (world == 1) and args.graph
    if use_graph:
        for b in bound:
            b.capture()
    # --graph-all: the K steps of the timed region as ONE CUDA graph of K kernel launches (no Python between launches).
    # It was the N > 1 default until the A/B of profiles/r01_graph_vs_eager.txt: a replayed graph keeps almost none of
    # the launch overlap (N = 1: 109.0 us per step against 103.0 eager; N = 8: 108.3 against 104.8), and eight
    # Python processes issuing one launch per 100 us do not starve anything once NVML is initialised before the barrier.
    graph_all = None
    if args.graph_all and not global_loss and not args.no_graph_all:
        for i in range(3):
            bound[i % R].launch_kernel(fused.L.stream())
        torch.cuda.synchronize()
        graph_all = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph_all):
            for i in range(args.steps):
                bound[i % R].launch_kernel(fused.L.stream())
        if epoch_loss is not None:
            torch.cuda.synchronize()
            epoch_loss.zero_()

    pending = []                                          # (async all-reduce, step) awaiting finalisation

    def flush():
        while pending:
            work, pb_ = pending.pop(0)
            work.wait()                                   # compute stream waits for the NCCL stream
            pb_.launch_finalize(fused.L.stream())

    def one_step(i, events=None):
        b = bound[i % R]
        if use_graph and events is None:
            b.replay()
        else:
            b.launch(events)
        if global_loss:
            # Global N_pos / sums for the balanced loss: the 32-byte all-reduce of step i runs on the NCCL stream
            # while the kernel of step i+1 runs; step i is finalised right after that kernel is queued.
            flush()
            pending.append((dist.all_reduce(b.sums, async_op=True), b))

    def end_of_epoch():
        flush()
        if epoch_loss is not None:
            dist.all_reduce(epoch_loss)                   # the one collective of the timed region

    def fence():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- warm-up, then the timed region ------------------------------------------------------------
    for i in range(args.warmup):
        one_step(i)
    end_of_epoch()
    # NVML is initialised and the sampler thread started BEFORE the barrier: nvmlInit from N processes at once takes
    # milliseconds and would otherwise skew the ranks' entry into the timed region (the final all-reduce then waits
    # for the last rank).  Sampling period: 0.5 ms at N = 1, 2 ms at N > 1.
    sampler = ClockSampler(physical_gpu_index(local_rank), args.clock_interval_ms * 1e-3 if args.clock_interval_ms > 0
                           else (0.0005 if world == 1 else 0.002))
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fence()
    if epoch_loss is not None:
        epoch_loss.zero_()
        fence()
    sampler.samples.clear()
    e0.record()
    if graph_all is not None:
        graph_all.replay()
    else:
        for i in range(args.steps):
            one_step(i)
    end_of_epoch()
    e1.record()
    fence()
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    if epoch_loss is not None:
        loss_val = float(epoch_loss.item()) / (world * args.steps)      # mean over ranks and steps
    else:
        loss_val = float(bound[(args.steps - 1) % R].loss.item())

    # ---- the fused kernel alone: a second pass of K back-to-back launches (no collective, no finalise)
    #      between two events on the launching stream; average launch duration = elapsed / K ---------------
    fence()
    st = fused.L.stream()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(min(3, args.steps)):
        bound[i % R].launch_kernel(st)
    k0.record()
    for i in range(args.steps):
        bound[i % R].launch_kernel(st)
    k1.record()
    fence()
    kernel_ms = k0.elapsed_time(k1) / args.steps

    # ---- end to end through the host-buffer API ---------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        s0 = sets[0]
        host = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t) for t in s0]
        pipe = fused.HostPipeline(step_cfg, B, K_JOINTS, H, W, flip=True, chunks=args.chunks, device=dev)
        for _ in range(max(1, min(args.warmup, 3))):
            pipe(host[0], host[1], host[2], host[3], host[4], host[5])
        fence()
        t0 = time.perf_counter()
        e2e_launches = 0
        for _ in range(args.steps):
            pipe(host[0], host[1], host[2], host[3], host[4], host[5])
            e2e_launches += pipe.launches
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        e2e = (e2e_s, pipe.h2d_bytes, pipe.d2h_bytes, e2e_launches)
        del pipe, host

    # ---- max over ranks ------------------------------------------------------------------------------------
    stats = torch.tensor([ms_total, kernel_ms, e2e[0] if e2e else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    ms_total, kernel_ms, e2e_s = [float(v) for v in stats.tolist()]

    if rank == 0:
        value = world * B * args.steps / (ms_total * 1e-3)
        peak, peak_src = measured_peak()
        bytes_launch = algorithmic_bytes_per_sample() * B
        achieved = bytes_launch / (kernel_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(B), "batch_per_gpu": B, "global_batch": B * world,
                       "joints": K_JOINTS, "heatmap": f"{H}x{W}", "flip_test": True, "loss": "DistanceLoss L2 balance=True",
                       "decode": "argmax + DARK k=11 + transform_preds", "parallelism": f"batch-shard x{world}" + (
                           "" if world == 1 else
                           (f"; global-N_pos loss: one 32-byte NCCL all-reduce per step, {args.spare_sms} SMs left to NCCL"
                            if global_loss else
                            "; per-rank loss (reference DDP semantics), ONE NCCL all-reduce of the epoch loss sum per timed region")),
                       "l2_policy": f"inputs {2 * B * K_JOINTS * H * W * 4 / 1e6:.0f} MB/step > 126 MB L2, "
                                    f"{R} rotating input sets, L2 evict_first loads",
                       "launch": ("one CUDA graph of all K launches" if graph_all is not None else
                                  "CUDA graph replay" if use_graph else "eager C-ABI launches, one per step") +
                                 ("" if args.no_overlap or R < 2 else
                                  "; LHN_FLAG_OVERLAP_PREVIOUS (programmatic dependent launch over rotating buffer sets)"),
                       "loss_check": loss_val},
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": recorded_traffic(B),
                         "kernel": "heatmap_team_kernel<f32,64x64,TW=4,FLIP,LOSS,KS=11>", "kernel_ms": kernel_ms,
                         "algorithmic_bytes_per_launch": bytes_launch, "peak_source": peak_src},
            "gpu_launches": args.steps * (2 if global_loss else 1),
        }
        if e2e:
            line["e2e"] = {"value": world * B * args.steps / e2e_s, "unit": UNIT,
                           "h2d_bytes_per_step": e2e[1], "d2h_bytes_per_step": e2e[2],
                           "api": "litehandnet_b200.fused.HostPipeline (pinned host buffers, chunked H2D overlapped with the fused kernel)",
                           "gpu_launches": e2e[3]}
        if world == 1 and not args.no_cpu_baseline:
            from oracle import cpu_path
            runner = cpu_path.FusedCpuRunner(*[t.cpu().numpy() for t in sets[0]], image_size=IMAGE_SIZE,
                                             sigma=2, kernel=11)
            sample = cpu_sample_size(runner, args.cpu_sample, 6.0)
            dts = [runner.run(sample)[2] for _ in range(2)]
            runner.close()
            line["cpu_baseline"] = {"value": sample / min(dts), "unit": UNIT, "cores": runner.workers, "kind": "port",
                                    "sample": f"first {sample} samples of the batch-{B} workload, best of 2 passes; numpy "
                                              f"oracle port of the reference pipeline in {runner.workers} forked workers"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="samples per GPU per step")
    ap.add_argument("--rotate", type=int, default=2, help="distinct device-resident input sets")
    ap.add_argument("--chunks", type=int, default=8, help="H2D/compute pipeline chunks of the e2e path")
    ap.add_argument("--cpu-sample", type=int, default=0)
    ap.add_argument("--graph", action="store_true", help="replay one CUDA graph per step instead of eager launches")
    ap.add_argument("--no-graph", action="store_true", help="(default now; kept for old command lines)")
    ap.add_argument("--global-loss", action="store_true",
                    help="N > 1: balanced loss with batch-global N_pos (an all-reduce of the f64 sums every step) "
                         "instead of the reference's per-rank loss")
    ap.add_argument("--spare-sms", type=int, default=4,
                    help="N > 1: SMs the persistent kernel leaves free so the NCCL all-reduce of the previous step "
                         "can run beside it")
    ap.add_argument("--clock-interval-ms", type=float, default=0.0, help="NVML sampling period (0 = automatic)")
    ap.add_argument("--graph-all", action="store_true",
                    help="run the timed steps as one CUDA graph of K launches instead of eager launches (A/B runs)")
    ap.add_argument("--no-graph-all", action="store_true",
                    help="(default now: eager launches at every N; kept for old command lines)")
    ap.add_argument("--no-overlap", action="store_true", help="do not let a launch overlap the previous one's tail")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_gpu_arm(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
