#!/usr/bin/env python
"""bench.py — instances/sec of IST-Net's per-instance forward+backward hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on the host cores

A "step" = train-mode forward + SupervisedLoss + backward over one synthetic batch of 32 instances per GPU
(1024 points + 192x192 RGB each; BASELINE.json configs[1]; weak scaling: 32 per GPU, configs[2] at N=8), plus the
NCCL gradient all-reduce when N > 1.  One process per GPU (torchrun env), CUDA-event timing, max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

LABELS = ("qo", "rotation_label", "translation_label", "size_label")
# dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel's largest launch (up_1 conv, B=32) from one
# `ncu --set full` capture (profiles/), bytes per launch; None until captured for the current kernel version
ROOFLINE_TRAFFIC = 398.0e6  # profiles/r1_conv_up1_ns2.txt (up_1 forward as it runs now, 2 operand planes): 337.0 MB read + 61.0 MB written (algorithmic: 302 MB operand planes + 75.5 MB output)
MODEL_IN = ("rgb", "pts", "choose", "category_label", "qo")
WORKLOADS = {
    "cfg1": dict(model="ist_net", batch=32, npts=1024, img=192, desc="ist_net_default.yaml train fwd+bwd, 32 x (1024 pts + 192x192 RGB) per GPU"),
    "cfg3": dict(model="posenet_gt", batch=64, npts=1024, img=192, desc="posenet_gt_default.yaml train fwd+bwd, 64 x (1024 pts + 192x192 RGB) per GPU"),
    "cfg4": dict(model="ist_net", batch=16, npts=4096, img=192, desc="ist_net_default.yaml train fwd+bwd, 16 x (4096 pts + 192x192 RGB) per GPU"),
}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p, "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None, "reasons": reasons, "samples": len(self.rows)}


def build_model(kind, device):
    from istnet_b200 import model as M

    torch.manual_seed(1)  # rd_seed, config/ist_net_default.yaml:60
    if kind == "ist_net":
        m, loss = M.IST_Net(6, False), M.SupervisedLoss(M.LossCfg(1.0, 10.0, False))
    else:
        m, loss = M.PoseNetGT(6), M.PoseNetGTLoss()
    return m.to(device).train(), loss


# ----------------------------------------------------------------------------------------------- reference arm
def cpu_reference_rate(wl, sample_batch, steps, warmup):
    """Reference CPU path = oracle port (plain torch FP32 + C point ops) on all host threads; returns inst/s."""
    from istnet_b200 import model as M
    from istnet_b200.synth import make_batch
    from oracle import istnet_port as port

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(1)
    mod = M.IST_Net(6, False) if wl["model"] == "ist_net" else M.PoseNetGT(6)
    sd = {k: v.clone() for k, v in mod.state_dict().items()}
    for k, v in sd.items():
        if v.is_floating_point() and "running_" not in k:
            v.requires_grad_(True)
    data = make_batch(sample_batch, wl["npts"], wl["img"], seed=1)
    times = []
    for it in range(warmup + steps):
        for v in sd.values():
            v.grad = None
        t0 = time.perf_counter()
        if wl["model"] == "ist_net":
            loss = port.ist_net_loss(port.ist_net_forward(sd, data, True), data)
        else:
            loss = port.posenet_gt_loss(port.posenet_gt_forward(sd, data, True), data)
        loss.backward()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return sample_batch * len(times) / sum(times), cores


def gpu_reference_rate(wl, dev, steps=20, warmup=3, tf32=True, graph=True):
    """Secondary bar: what the UNMODIFIED reference does on this GPU — its module dataflow (oracle port =
    the reference's torch ops, verified bit-identical to the reference modules) on cuDNN/cuBLAS + the reference's own CUDA
    extension compiled from /root/reference (oracle/_ref), the whole step captured in a CUDA graph like this repo's arm.
    tf32=True: PyTorch defaults (TF32 convolutions; misses the 1e-4 bar), tf32=False: the parity-grade FP32 arithmetic.
    Context only.  Returns (inst/s, "cuda-graph" | "eager") or None when oracle/_ref is absent."""
    import importlib.util

    so = os.path.join(ROOT, "oracle", "_ref", "pointnet2_ref", "_ext.so")
    if not os.path.exists(so):
        return None
    from istnet_b200 import model as M
    from istnet_b200.synth import make_batch
    from oracle import istnet_port as port

    spec = importlib.util.spec_from_file_location("_ext", so)
    ref_ext = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_ext)
    torch.manual_seed(1)
    mod = M.IST_Net(6, False) if wl["model"] == "ist_net" else M.PoseNetGT(6)
    sd = {k: v.clone().to(dev) for k, v in mod.state_dict().items()}
    for k, v in sd.items():
        if v.is_floating_point() and "running_" not in k:
            v.requires_grad_(True)
    data = {k: v.to(dev) for k, v in make_batch(wl["batch"], wl["npts"], wl["img"], seed=1).items()}

    def one_step():
        for v in sd.values():
            v.grad = None
        if wl["model"] == "ist_net":
            loss = port.ist_net_loss(port.ist_net_forward(sd, data, True, ops=ref_ext), data)
        else:
            loss = port.posenet_gt_loss(port.posenet_gt_forward(sd, data, True, ops=ref_ext), data)
        loss.backward()
        return loss

    with torch.backends.cudnn.flags(enabled=True, allow_tf32=tf32):
        mode, g = "eager", None
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                one_step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if graph:
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    one_step()
                mode = "cuda-graph"
            except Exception as e:  # e.g. an op of the reference dataflow that synchronises or allocates outside the capture
                g = None
                mode = "eager (capture failed: " + " ".join(str(e).split())[:160] + ")"
                torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(steps):
            if g is not None:
                g.replay()
            else:
                one_step()
        ev1.record()
        torch.cuda.synchronize()
    return wl["batch"] * steps / (ev0.elapsed_time(ev1) / 1000.0), mode


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = max(1, min(8, wl["batch"]))  # bounded sample of the workload: ~10 s of host work at ~6.5 inst/s
    steps, warmup = min(args.steps, 6), min(args.warmup, 1)
    rate, cores = cpu_reference_rate(wl, sample, steps, warmup)
    line = {
        "impl": "reference", "metric": "instances/sec", "value": rate, "unit": "instances/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": 1000.0 * sample / rate, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.config}: {wl['desc']}", "sample_batch": sample},
        "cpu_baseline": {"value": rate, "unit": "instances/s", "cores": cores, "kind": "port",
                         "sample": f"{steps} timed steps of fwd+loss+bwd on a batch of {sample} (same shapes as the workload)"},
        "e2e": {"value": rate, "unit": "instances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- own arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg1", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eager", action="store_true", help="do not capture the step in a CUDA graph")
    ap.add_argument("--ref-gpu-probe", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--no-optimizer", action="store_true", help="time forward + loss + backward (+ all-reduce) without the Adam step")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.config])
    if args.batch:
        wl["batch"] = args.batch
    if args.impl == "reference":
        return run_reference(args, wl)
    if args.ref_gpu_probe:
        dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
        torch.cuda.set_device(dev)
        ref = {}
        for name, tf32 in (("tf32_default", True), ("fp32_parity_grade", False)):
            r = gpu_reference_rate(wl, dev, tf32=tf32)
            if r is not None:
                ref[name] = {"value": r[0], "launch": r[1]}
        if ref:
            ref.update({"unit": "instances/s", "what": "reference dataflow (oracle port) on cuDNN/cuBLAS + the reference's own CUDA extension "
                        "(oracle/_ref), fwd+loss+bwd, 20 steps, same GPU; context only (tf32_default misses the 1e-4 parity bar)"})
            print(json.dumps(ref), flush=True)
        return

    from istnet_b200 import _C
    from istnet_b200.parallel import DataParallelStep, FlatAdam, GradAllReducer, broadcast_module
    from istnet_b200.synth import flops_per_instance, make_batch

    assert torch.cuda.is_available(), "bench.py (own arm) needs a GPU; there is no CPU fallback"
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    warmup = max(args.warmup, 3)
    model, loss_fn = build_model(wl["model"], dev)
    broadcast_module(model)
    reducer = GradAllReducer(model)
    B = wl["batch"]
    host = make_batch(B, wl["npts"], wl["img"], seed=1 + rank)
    host = {k: v.pin_memory() for k, v in host.items()}
    resident = {k: v.to(dev) for k, v in host.items()}
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())

    opt = None  # built after the first step (the flat buckets come from the parameters that received a gradient)

    def eager_step(data):
        reducer.zero_grad()
        ep = model({k: data[k] for k in MODEL_IN})
        ep.update({k: data[k] for k in LABELS})
        loss = loss_fn(ep)
        loss.backward()
        reducer.finish()
        if opt is not None:
            opt.step()
        return loss

    # one eager step first: discovers the gradient buckets and warms everything up
    eager_step(resident)
    if not args.no_optimizer:
        # the reference's optimizer (utils/solver.py:41-46: Adam, default betas/eps, CyclicLR from base_lr 1e-5) on the flat buckets
        opt = FlatAdam(reducer, lr=1e-5, weight_decay=0.0)
    l_probe = _C.LAUNCHES
    eager_step(resident)  # counts the kernel launches of a step (the graph replays exactly these)
    launches_per_step = _C.LAUNCHES - l_probe
    trace_stages = os.environ.get("ISTNET_TRACE_STAGES") == "1"

    def stage(msg):  # debugging aid for multi-GPU bring-up: where does a rank stop?
        if trace_stages:
            torch.cuda.synchronize()
            print(f"[rank {rank}] {msg}", file=sys.stderr, flush=True)

    stage("eager steps done")
    graphed = None
    # ISTNET_GRAPH_NCCL=1 captures the per-bucket NCCL all-reduce (communication stream) and Adam inside the step's graph; the
    # default keeps NCCL out of the capture: the graph ends with the bucket packing, all-reduce + Adam follow each replay
    nccl_in_graph = os.environ.get("ISTNET_GRAPH_NCCL", "0") == "1"
    if not args.eager:
        from istnet_b200.graph import GraphedTrainStep

        # the captured step = forward + loss + backward + per-bucket gradient packing / NCCL all-reduce (communication stream,
        # first bucket behind the image branch's backward) + Adam on the flat buckets
        dp = DataParallelStep(reducer, opt, nccl_in_graph=nccl_in_graph)
        graphed = GraphedTrainStep(model, loss_fn, resident, MODEL_IN, LABELS, before_forward=dp.before, after_backward=dp.after,
                                   before_replay=dp.before_replay, after_replay=dp.after_replay)

    def step(data):
        if graphed is None:
            return eager_step(data)
        return graphed()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # e2e pipeline (graph path): the host->device copy of batch i+1 runs on a copy stream into a staging buffer while step i
    # computes; a device-to-device copy moves it into the graph's static inputs at the start of step i+1; the loss of step i is read
    # back (pinned, asynchronous) while step i+1 is already enqueued.  Every step still copies its own inputs from pinned host
    # memory and has its loss read on the host inside the timed region — only the waiting is overlapped.
    copy_stream = torch.cuda.Stream()
    staging = {k: torch.empty_like(v, device=dev) for k, v in host.items()}
    loss_pinned = torch.zeros(2, dtype=torch.float32).pin_memory()

    def timed(n, e2e):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        if e2e and graphed is not None:
            cur = torch.cuda.current_stream()
            ev_copy, ev_loaded, ev_loss = torch.cuda.Event(), torch.cuda.Event(), [torch.cuda.Event(), torch.cuda.Event()]

            def prefetch():
                with torch.cuda.stream(copy_stream):
                    for k, v in host.items():
                        staging[k].copy_(v, non_blocking=True)
                    ev_copy.record(copy_stream)

            copy_stream.wait_stream(cur)
            prefetch()
            for i in range(n):
                cur.wait_event(ev_copy)
                graphed.load(staging)          # device-to-device into the graph's static inputs
                ev_loaded.record(cur)
                if i + 1 < n:
                    copy_stream.wait_event(ev_loaded)
                    prefetch()                 # batch i+1 travels while step i computes
                loss = step(None)
                loss_pinned[i % 2 : i % 2 + 1].copy_(loss.reshape(1), non_blocking=True)
                ev_loss[i % 2].record(cur)
                if i > 0:
                    ev_loss[(i - 1) % 2].synchronize()
                    _ = float(loss_pinned[(i - 1) % 2])  # the previous step's result on the host
            ev_loss[(n - 1) % 2].synchronize()
            _ = float(loss_pinned[(n - 1) % 2])
        else:
            for _ in range(n):
                if e2e:
                    loss = step({k: v.to(dev, non_blocking=True) for k, v in host.items()})
                    _ = loss.item()  # device->host read of the step's result
                else:
                    step(resident)
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    stage("step built (graph captured)" if graphed is not None else "eager mode")
    for i_ in range(warmup):
        step(resident)
        stage(f"warm-up step {i_} done")
    barrier()
    l0 = _C.LAUNCHES
    with ClockSampler(local) as clk:
        ms = timed(args.steps, e2e=False)
    launches = (_C.LAUNCHES - l0) if graphed is None else launches_per_step * args.steps
    ms_e2e = timed(args.steps, e2e=True)
    value = world * B * args.steps / (ms / 1000.0)
    e2e_value = world * B * args.steps / (ms_e2e / 1000.0)

    # ---- roofline of the dominant kernel (conv_gemm_tc_kernel: ~31 % of device time, profiles/): one extra eager step,
    # outside the timed region, with CUDA events on the launching stream around every tensor-core launch (each re-issued
    # 3x back to back between the events so the device-side duration is measured, not the host's launch latency)
    from istnet_b200 import nhwc

    from istnet_b200 import model as model_mod

    nhwc.PROFILE = []
    from istnet_b200 import pointnet2 as pn2_mod

    model_mod.USE_SIDE_STREAMS, nhwc.WGRAD_SIDE_STREAM, pn2_mod.SA_FORK = False, False, False  # time each kernel alone on its launching stream
    reducer.zero_grad()
    ep_ = model({k: resident[k] for k in MODEL_IN})
    ep_.update({k: resident[k] for k in LABELS})
    loss_fn(ep_).backward()
    torch.cuda.synchronize()
    prof, nhwc.PROFILE = nhwc.PROFILE, None
    kstat = {}
    table = []
    for name, e0, e1, reps, fl, ns, desc in prof:
        d = kstat.setdefault(name, {"ms": 0.0, "flop": 0.0, "mma_flop": 0.0, "n": 0})
        table.append((e0.elapsed_time(e1) / reps, name, ns, fl, desc))
        d["ms"] += e0.elapsed_time(e1) / reps
        d["flop"] += fl
        d["mma_flop"] += fl * (ns * (ns + 1) // 2)
        d["n"] += 1

    if rank == 0 and os.environ.get("ISTNET_KERNEL_TABLE"):  # per-launch table of the tensor-core kernels (tools / profiles)
        with open(os.environ["ISTNET_KERNEL_TABLE"], "w") as f:
            for ms_, name, ns, fl, desc in table:
                f.write(f"{ms_ * 1000:9.1f} us  {name:22s} ns={ns}  {fl * (ns * (ns + 1) // 2) / (ms_ * 1e-3) / 1e12:7.1f} MMA-TFLOP/s  {desc}\n")
    if rank == 0:
        pk, pk_src = peaks()
        gflop = flops_per_instance(wl["model"], wl["npts"], True)
        step_tflops = (value / world) * gflop / 1000.0  # TFLOP/s per GPU, algorithmic, whole step
        dom = kstat["conv_gemm_tc_kernel"]
        achieved = dom["flop"] / (dom["ms"] * 1e-3) / 1e12
        peak = pk["bf16_tflops"]
        line = {
            "metric": "instances/sec", "value": value, "unit": "instances/s", "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"{args.config}: {wl['desc']}", "per_gpu_batch": B, "global_batch": B * world,
                       "parallelism": f"dp{world}", "launch": "eager" if graphed is None else "cuda-graph (whole step)",
                       "step": "forward + SupervisedLoss + backward" + (" + NCCL gradient all-reduce" if world > 1 else "") +
                               ("" if opt is None else " + Adam (flat buckets, utils/solver.py:41-46)") +
                               ((" [all inside the graph]" if nccl_in_graph or world == 1 else " [all-reduce + Adam after the replay]") if graphed is not None else ""),
                       "l2": "per-step activation working set (>1 GB) exceeds the 126 MB L2; no explicit flush"},
            "e2e": {"value": e2e_value, "unit": "instances/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps,
                    "how": "every step copies its batch from pinned host memory and its loss is read on the host; the copy of batch i+1 "
                           "(copy stream -> staging buffer) and the read of loss i overlap step i / i+1" if graphed is not None else
                           "host->device copies and loss.item() serial with the step"},
            "gpu_launches": launches,
            "clocks": clk.summary(),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": ROOFLINE_TRAFFIC,
                         "kernel": "conv_gemm_tc_kernel (implicit-GEMM conv / 1x1 / dgrad, tcgen05)", "launches_per_step": dom["n"],
                         "avg_launch_ms": dom["ms"] / dom["n"], "algorithmic_gflop_per_launch": dom["flop"] / dom["n"] / 1e9,
                         "executed_mma_tflops": dom["mma_flop"] / (dom["ms"] * 1e-3) / 1e12,
                         "peak_source": f"bf16 dense burst (kernel timed alone), {pk_src}",
                         "note": "achieved counts one multiply-add per reference MAC; the tensor pipe executes 6x (3 bf16 operand planes: PointNet++ and lower ResNet layers) or 3x (2 planes: up_1..3, layer4, pose heads, every backward GEMM) that for FP32-level accuracy (DESIGN.md section 2)",
                         "wgrad_tc_kernel": {"achieved": kstat["wgrad_tc_kernel"]["flop"] / (kstat["wgrad_tc_kernel"]["ms"] * 1e-3) / 1e12,
                                             "launches_per_step": kstat["wgrad_tc_kernel"]["n"]},
                         "whole_step": {"achieved": step_tflops, "frac_of_sustained_peak": step_tflops / pk["bf16_tflops_sustained"],
                                        "gflop_per_instance": gflop}},
        }
        if not args.no_cpu_baseline and world == 1:
            rate, cores = cpu_reference_rate(wl, 4, 3, 1)
            line["cpu_baseline"] = {"value": rate, "unit": "instances/s", "cores": cores, "kind": "port",
                                    "sample": "3 timed steps of fwd+loss+bwd on a batch of 4 (same shapes), oracle port on host threads"}
            # context only, in a CHILD process with a time limit: the reference's extension exit()s on a CUDA error
            # (cuda_utils.h:35-44) and a failed capture must not take this process (and its JSON line) with it
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--ref-gpu-probe", "--config", args.config, "--batch", str(B)],
                                   capture_output=True, text=True, timeout=240)
                probe = [l for l in r.stdout.splitlines() if l.startswith("{")]
                if probe:
                    line["reference_gpu_path"] = json.loads(probe[-1])
                else:
                    line["reference_gpu_path_error"] = (r.stderr or r.stdout)[-200:]
            except Exception as e:
                line["reference_gpu_path_error"] = str(e)[:200]
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
