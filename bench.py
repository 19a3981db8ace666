#!/usr/bin/env python
"""Headline benchmark: one training step of the STswinCL Swin head (the window-attention hot
path, BASELINE.json north_star) on synthetic EndoVis18-shaped clips.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step = forward + backward + optimizer step of ``SwinTransformerLayerv5`` (dim 512, 64x80
tokens, 4 heads, 12 blocks + PatchMerging -- seg18/net/Ours/swin_512.py:280-327) over one batch
of ``--clips`` clips x 4 frames of OS-8 features [clips, 4, 512, 64, 80] (configs[1] of
BASELINE.json: batch 8, bf16, 1 B200).  frames/s counts input frames (clips x 4 per step).

Prints ONE JSON line (rank 0).  ``value`` has the inputs resident in HBM; ``e2e`` feeds the same
module from pinned HOST buffers (H2D copy every step, loss read back every step).  ``roofline``
is measured with CUDA events around every launch of the dominant kernel family during extra,
separately timed steps; ``cpu_baseline`` is the CPU oracle (a port of the reference path, see
oracle/) timed on this box's host cores on a bounded sample (N=1, rank 0 only).

``--impl reference`` times the reference's CPU path (the oracle port: the reference is pure
Python/PyTorch with no compiled artefact to build; /root/reference does not exist on the GPU
box) with all host threads and prints the same line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "STswin-T train frames/s at 1/2/4/8 B200; window-attn TFLOP/s vs tensor peak"
DIM, RES, HEADS, T = 512, (64, 80), 4, 4


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--clips", type=int, default=8, help="clips per GPU per step (reference recipe: batch 8)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="do not capture the step in a CUDA graph")
    ap.add_argument("--dp", default="flat", choices=["flat", "ddp"],
                    help="N > 1: 'flat' = forward + backward replayed from a CUDA graph, ONE NCCL all-reduce (average) of the "
                         "flattened fp32 gradients, optimizer step; 'ddp' = torch DistributedDataParallel, eager steps")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tflops_burst": d["bf16_tflops"], "tflops_sustained": d["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.path = index, None, None

    def __enter__(self):
        if self.index is None:
            return self
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        return False

    def summary(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.path or not os.path.exists(self.path):
            return out
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ----------------------------------------------------------------------------------------------
# CPU path (oracle port of the reference) -- cpu_baseline leg and --impl reference
# ----------------------------------------------------------------------------------------------
def cpu_step_factory():
    import torch
    from oracle import swin_oracle as so          # checker / CPU baseline only (never on the product path)
    torch.set_num_threads(os.cpu_count() or 1)
    params = so.make_layer_params(DIM, RES, HEADS, seed=0)
    leaf = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and not k.endswith("attn_mask") else v)
            for k, v in params.items()}
    x = so.make_features(1, 1, T, DIM, RES[0], RES[1])
    g1 = so.make_features(2, 1, T, DIM, RES[0], RES[1]) - 0.4
    g2 = so.make_features(3, 1, T, 2 * DIM, RES[0] // 2, RES[1] // 2) - 0.4

    def step():
        for v in leaf.values():
            if v.is_floating_point() and v.requires_grad:
                v.grad = None
        y1, y2 = so.swin_layer_v5(x, leaf, DIM, RES, HEADS)
        loss = (y1 * g1).sum() + (y2 * g2).sum()
        loss.backward()
        return float(loss)

    return step, torch.get_num_threads()


SAMPLE = "1 clip (4 frames) of the same workload [1,4,512,64,80], fp32, forward+backward, oracle port of the reference"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    step, threads = cpu_step_factory()
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    fps = T / dt
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.clips),
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port", "sample": SAMPLE},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(clips, graph=None):
    return {"cuda_graph": graph, "workload": "configs[1]: Swin head train step (SwinTransformerLayerv5 dim 512, 64x80 tokens, 4 heads; "
                        "fwd+bwd+Adam) on EndoVis18-shaped OS-8 features, bf16",
            "clips_per_gpu": clips, "frames_per_clip": T, "feature_shape": [clips, T, DIM, RES[0], RES[1]],
            "optimizer": "torch.optim.Adam(fused=True) on the 96.6M Swin-head parameters",
            "l2": "inputs larger than L2 (168 MB of features per step, >10 GB of activations touched)"}


# ----------------------------------------------------------------------------------------------
# GPU path
# ----------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from stswincl_b200 import ops, swin

    torch.manual_seed(0)
    model = swin.SwinTransformerLayerv5(dim=DIM, input_resolution=RES, num_heads=HEADS).to(dev)
    for blk in model.modules():
        if isinstance(blk, swin.WindowAttention):       # sigma 0.02 init is nearly a no-op bias; use a visible one
            torch.nn.init.normal_(blk.relative_position_bias_table, std=0.5)
    net = model
    side = torch.cuda.Stream(device=dev)
    flat_dp = world > 1 and args.dp == "flat"
    params = list(model.parameters())
    if world > 1 and not flat_dp:
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], gradient_as_bucket_view=True,
                                                        bucket_cap_mb=64, broadcast_buffers=False)
    flat = None
    if flat_dp:
        from stswincl_b200 import dist as sdist
        sdist.broadcast_parameters(params)                 # replicas start identical (DDP's constructor does the same)
        flat = sdist.FlatGradients(params)
    opt = torch.optim.Adam(model.parameters(), lr=3e-5, fused=True, capturable=True)
    B = args.clips
    g = torch.Generator(device=dev).manual_seed(1 + rank)
    x_dev = torch.relu(torch.randn(B, T, DIM, RES[0], RES[1], generator=g, device=dev)).to(torch.bfloat16)
    g1 = (torch.randn(B, T, DIM, RES[0], RES[1], generator=g, device=dev) * 0.1).to(torch.bfloat16)
    g2 = (torch.randn(B, T, 2 * DIM, RES[0] // 2, RES[1] // 2, generator=g, device=dev) * 0.1).to(torch.bfloat16)

    def fwd_bwd(x):
        opt.zero_grad(set_to_none=True)
        y1, y2 = net(x)
        loss = torch.sum(y1 * g1, dtype=torch.float32) + torch.sum(y2 * g2, dtype=torch.float32)
        loss.backward()
        if flat_dp:                                        # gather the gradients into the all-reduce buffer
            flat.gather()
        return loss

    def reduce_and_update():
        if flat_dp:                                        # the data-parallel exchange: one all-reduce over NVLink
            flat.all_reduce()
            flat.bind()
        opt.step()

    def step(x):
        loss = fwd_bwd(x)
        reduce_and_update()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / n

    # ---- device-resident throughput
    for _ in range(args.warmup):
        step(x_dev)
    launches0 = ops.LAUNCHES
    step(x_dev)
    launches_per_step = ops.LAUNCHES - launches0

    # The step has no host synchronisation and static shapes, so it is captured once into a CUDA graph and replayed:
    # on one GPU the whole forward + backward + optimizer step; on N GPUs (--dp flat) forward + backward + the gather of
    # the gradients into one flat buffer, followed -- outside the graph -- by one NCCL all-reduce and the optimizer step.
    # torch DDP steps (--dp ddp) run eagerly: capturing them (DDP on a side stream, 11 eager iterations first) measured
    # 1580 vs 1557 frames/s at 2 GPUs but the process then hangs in the NCCL teardown.
    graph, static_x, static_loss = None, None, None
    captured = fwd_bwd if flat_dp else step
    if (world == 1 or flat_dp) and not args.no_graph:
        try:
            static_x = x_dev.clone()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    step(static_x)
            torch.cuda.current_stream().wait_stream(side)
            opt.zero_grad(set_to_none=True)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_loss = captured(static_x)
            graph.replay()
            if flat_dp:
                reduce_and_update()
            torch.cuda.synchronize()
        except Exception as e:                      # capture is an optimisation of the harness, not of the product
            print(f"bench.py: CUDA graph capture failed ({type(e).__name__}: {e}); timing eager steps", file=sys.stderr)
            graph = None
            torch.cuda.synchronize()
        if flat_dp:          # every rank replays, or none does
            ok = torch.tensor([1 if graph is not None else 0], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok) == 0:
                graph = None

    def run_step(x):
        if graph is None:
            return step(x)
        if x is not static_x:
            static_x.copy_(x, non_blocking=True)
        graph.replay()
        if flat_dp:
            reduce_and_update()
        return static_loss

    for _ in range(2):
        run_step(x_dev)
    # nvidia-smi is started by rank 0 only and well before the timed steps: its start-up (NVML initialisation over
    # all GPUs of the box) stalls kernel launches of every process for ~0.1-0.2 s -- measured at 4 GPUs, where one
    # sampler per rank starting inside the timed region cost the eager DDP steps 58 instead of 41 ms
    with ClockSampler(local if rank == 0 else None) as clk:
        for _ in range(12):
            run_step(static_x if graph is not None else x_dev)
        ms_step = timed(lambda i: run_step(static_x if graph is not None else x_dev), args.steps)
    launches = launches_per_step * args.steps
    clocks = clk.summary()

    # ---- end to end: pinned host features -> H2D (copy stream, double buffered) -> step -> loss.item()
    x_host = x_dev.cpu().pin_memory()
    bufs = [torch.empty_like(x_dev), torch.empty_like(x_dev)]
    copy_stream = torch.cuda.Stream(device=dev)
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    d2h_bytes = 4

    def prefetch(i):
        b = i & 1
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[b])
            bufs[b].copy_(x_host, non_blocking=True)
            ready[b].record(copy_stream)

    def e2e_step(i):
        b = i & 1
        if i == 0:
            prefetch(0)
        prefetch(i + 1)                      # next step's features travel while this step computes
        torch.cuda.current_stream().wait_event(ready[b])
        loss = run_step(bufs[b])
        consumed[b].record(torch.cuda.current_stream())
        return loss.item()                   # D2H read of the step's result

    for b in range(2):
        consumed[b].record(torch.cuda.current_stream())
    for i in range(max(2, args.warmup)):
        e2e_step(i)
    ms_e2e = timed(e2e_step, args.steps)

    # ---- roofline of the dominant kernel family: CUDA events around every launch (separate steps)
    pk = peaks()
    # The profiled steps run eagerly.  A replay of the captured step (one launch for the host, a whole step of real
    # work for the GPU) is queued in front of each, so the host stays ahead of the GPU and every event pair brackets
    # its kernel back to back with its neighbours on a GPU as warm as in the timed region -- without the head start
    # the short kernels' times include the host's launch gaps.
    prof = ops.EventProfiler()
    for _ in range(2):
        if graph is not None:
            graph.replay()
        else:
            step(x_dev)
        ops.set_profiler(prof)
        step(x_dev)
        ops.set_profiler(None)
    fam = prof.summary()
    total_ms = sum(v["ms"] for v in fam.values()) or 1.0
    shares = {k: round(v["ms"] / total_ms, 4) for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])}
    dom = max(("gemm", "gemm_wgrad"), key=lambda k: fam.get(k, {"ms": 0})["ms"])
    gsum = {"ms": fam["gemm"]["ms"] + fam["gemm_wgrad"]["ms"], "work": fam["gemm"]["work"] + fam["gemm_wgrad"]["work"],
            "launches": fam["gemm"]["launches"] + fam["gemm_wgrad"]["launches"]}
    achieved = gsum["work"] / (gsum["ms"] * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("gemm_dram_bytes_per_launch")
    roofline = {"kernel": "stswin::gemm_kernel (tcgen05 dense layers, fwd + dgrad + wgrad)", "bound": "tensor",
                "achieved": achieved, "peak": pk["tflops_sustained"], "unit": "TFLOP/s",
                "frac": achieved / pk["tflops_sustained"], "traffic": traffic, "peak_source": pk["source"] + ", sustained",
                "launches_timed": gsum["launches"], "avg_launch_ms": gsum["ms"] / gsum["launches"],
                "share_of_step": round(gsum["ms"] / total_ms, 4), "family_time_shares": shares}
    # secondary, HBM-bound: the window-attention core kernels and the row-wise / layout kernels
    for k in ("winattn_fwd", "winattn_bwd", "layernorm_fwd", "layernorm_bwd", "transpose", "copy"):
        if k in fam:
            gbs = fam[k]["work"] / (fam[k]["ms"] * 1e-3) / 1e9
            roofline[k] = {"bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"],
                           "avg_launch_ms": fam[k]["ms"] / fam[k]["launches"]}

    # ---- "window-attn TFLOP/s vs tensor peak" (second half of the metric): the WindowAttention module (qkv Linear +
    # shifted-window core + proj Linear, SURVEY 8d: 8C^2 + 4LC FLOPs per token forward, 2x that backward) at this
    # workload's batch, stage-1 (C 512, 64x80, ws 8, shift 4) and stage-2 (C 1024, 32x40, ws 4, shift 2) geometry,
    # replayed from a CUDA graph and timed alone -> against the measured burst bf16 peak
    window_attn = None
    if rank == 0 and world == 1:
        window_attn = {}
        for name, dim, res, ws in (("stage1", DIM, RES, 8), ("stage2", 2 * DIM, (RES[0] // 2, RES[1] // 2), 4)):
            attn = swin.WindowAttention(dim, (ws, ws), HEADS).to(dev)
            torch.nn.init.normal_(attn.relative_position_bias_table, std=0.5)
            xa = torch.relu(torch.randn(2 * B, 2, res[0] * res[1], dim, device=dev)).to(torch.bfloat16).requires_grad_(True)
            ga = (torch.randn_like(xa) * 0.1)
            geom = (res[0], res[1], HEADS, ws, ws // 2, 0.0)

            def attn_fwd_bwd():
                xa.grad = None
                attn.zero_grad(set_to_none=True)
                y = swin._AttentionFn.apply(xa, attn.relative_position_bias_table, attn.qkv.weight, attn.qkv.bias,
                                            attn.proj.weight, attn.proj.bias, geom)
                y.backward(ga)

            def attn_fwd():
                with torch.no_grad():
                    swin._AttentionFn.apply(xa, attn.relative_position_bias_table, attn.qkv.weight, attn.qkv.bias,
                                            attn.proj.weight, attn.proj.bias, geom)

            res_ms = {}
            for tag, fn in (("fwd", attn_fwd), ("fwd_bwd", attn_fwd_bwd)):
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                side = torch.cuda.Stream(device=dev)
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    fn()
                torch.cuda.current_stream().wait_stream(side)
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr):
                    fn()
                for _ in range(3):
                    gr.replay()
                res_ms[tag] = timed(lambda i: gr.replay(), 10)
            tok = 2 * B * 2 * res[0] * res[1]
            f_fwd = (8.0 * dim * dim + 4.0 * (2 * ws * ws) * dim) * tok
            window_attn[name] = {
                "tokens": tok, "gflop_fwd": f_fwd / 1e9, "fwd_ms": res_ms["fwd"], "fwd_bwd_ms": res_ms["fwd_bwd"],
                "fwd_tflops": f_fwd / res_ms["fwd"] / 1e9, "fwd_bwd_tflops": 3 * f_fwd / res_ms["fwd_bwd"] / 1e9,
                "peak_tflops": pk["tflops_burst"], "fwd_frac_of_peak": f_fwd / res_ms["fwd"] / 1e9 / pk["tflops_burst"],
                "fwd_bwd_frac_of_peak": 3 * f_fwd / res_ms["fwd_bwd"] / 1e9 / pk["tflops_burst"]}
            del attn, xa, ga

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cstep, threads = cpu_step_factory()
        cstep()                                           # warm-up (page-in, thread pool)
        t0 = time.perf_counter(); cstep(); dt = time.perf_counter() - t0
        cpu = {"value": T / dt, "unit": "frames/s", "cores": threads, "kind": "port", "sample": SAMPLE, "seconds": dt}

    if rank == 0:
        frames = B * T * world
        line = {"metric": METRIC, "value": frames / (ms_step * 1e-3), "unit": "frames/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": dict(workload_config(B, graph is not None),
                               parallelism=("dp%d: %s" % (world, "one flat fp32 gradient all-reduce (NCCL, average) per step" if flat_dp
                                                          else "torch DDP, bucketed all-reduce overlapped with backward")) if world > 1 else "single GPU"),
                "clocks": clocks,
                "e2e": {"value": frames / (ms_e2e * 1e-3), "unit": "frames/s", "ms_per_step": ms_e2e,
                        "h2d_bytes_per_step": x_host.numel() * x_host.element_size(), "d2h_bytes_per_step": d2h_bytes},
                "gpu_launches": launches, "roofline": roofline, "window_attn": window_attn, "cpu_baseline": cpu,
                "clips_per_s": B * world / (ms_step * 1e-3)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # convenience: self-launch one rank per GPU (the driver launches torchrun itself)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
