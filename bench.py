#!/usr/bin/env python
"""Benchmark of the conditional RealNVP hot path (BASELINE.json metric: images/sec for log-likelihood
eval and sampling).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N ...            # the reference's CPU path (oracle port)

Workload (N=1): BASELINE config 2 — conv_cINN class-conditioned 28x28x1 image + 1 label plane
(io_shape [28,28,2], x_d 1, lists [0,1,0,0]/[3,3,3,3]/[64,64,32,32]/[8,8,4,4]), batch 256 per GPU,
synthetic inputs and "trained-like" random weights.  One STEP = one log-likelihood
evaluation (`log_loss`) of a batch + one sampling pass (`call(zy,-1)`) of a batch; images/sec counts the
images through both (2*B per step).  N>1: one process per GPU (torchrun), the batch dimension is sharded
(weak scaling: 256 images per GPU), no data-path collective; time is the max over ranks.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CFG2 = dict(io_shape=[28, 28, 2], x_d=1, squeeze_factor_block_list=[0, 1, 0, 0], ResNeXt_block_list=[3, 3, 3, 3],
            num_kernels_list=[64, 64, 32, 32], cardinality_list=[8, 8, 4, 4])
WORKLOAD = "cfg2: conv_cINN class-conditioned 28x28x1 + label plane (io 28x28x2), batch 256/GPU, log_loss + sampling"
METRIC = "images/sec (log-likelihood eval + sampling)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="images per GPU per step")
    ap.add_argument("--cpu-sample-batch", type=int, default=32)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# clocks during the timed region (B200_PROFILING.md recipe)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference's CPU path
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(batch, steps, warmup):
    import numpy as np
    import torch
    from oracle.flow_torch import FlowOracle
    from oracle.weights import init_weights, synth_inputs
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    o = FlowOracle(**CFG2, dtype=torch.float32)
    o.set_weights(init_weights(o.plan, 'rand', seed=0))
    x = synth_inputs('cfg2', batch, seed=0)
    z = synth_inputs('noise:28x28x2', batch, seed=1)
    z[..., 1:] = x[..., 1:]
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        o.log_loss(x)
        o.call(z, -1)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    total = sum(times)
    return {"value": 2 * batch * steps / total, "ms_per_step": 1e3 * total / steps, "cores": cores,
            "torch_threads": torch.get_num_threads(),
            "sample": f"{steps} steps of log_loss + call(-1) on a batch of {batch} cfg2 images "
                      f"(oracle port, torch {torch.__version__} CPU fp32, {warmup} warm-up)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 3))
    r = cpu_reference_run(args.cpu_sample_batch, steps, 1)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "images/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": 1, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "reference CPU path: TensorFlow is not installable here, so this is "
                   "the repo's oracle port of the reference (restated reference on CPU, not TensorFlow); each step is "
                   f"a bounded sample of {args.cpu_sample_batch} images"},
        "cpu_baseline": {"value": r["value"], "unit": "images/s", "cores": r["cores"], "kind": "port",
                         "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from arl_conditional_normalizing_flows_b200.conv_cINN_make_model import cFlow
    from arl_conditional_normalizing_flows_b200 import _lib
    from oracle.weights import init_weights, synth_inputs
    from oracle.planner import plan_flow

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    B = args.batch
    H, W, D = CFG2["io_shape"]
    model = cFlow(**CFG2, device=dev)
    plan = plan_flow(CFG2["io_shape"], CFG2["x_d"], CFG2["squeeze_factor_block_list"], CFG2["ResNeXt_block_list"],
                     CFG2["num_kernels_list"], CFG2["cardinality_list"])
    model.set_weights(init_weights(plan, 'rand', seed=0))

    # rotating synthetic inputs (different batch every step; shards differ per rank)
    NBUF = 8
    xs, zs = [], []
    for i in range(NBUF):
        x = synth_inputs('cfg2', B, seed=1000 * rank + i)
        z = synth_inputs('noise:28x28x2', B, seed=5000 + 1000 * rank + i)
        z[..., 1:] = x[..., 1:]
        xs.append(torch.from_numpy(x))
        zs.append(torch.from_numpy(z))
    xs_d = [t.to(dev) for t in xs]
    zs_d = [t.to(dev) for t in zs]
    xs_h = [t.pin_memory() for t in xs]
    zs_h = [t.pin_memory() for t in zs]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device(i):
        model.log_loss(xs_d[i % NBUF])
        model(zs_d[i % NBUF], -1)

    host_out = {"loss": torch.empty(4).pin_memory(), "ps": torch.empty((3, (B + 3) & ~3)).pin_memory(),
                "x": torch.empty((B, H, W, D)).pin_memory()}
    dbuf = {"x": torch.empty((B, H, W, D), device=dev), "z": torch.empty((B, H, W, D), device=dev)}

    def step_e2e(i):
        # host buffers in, host results out, through the public API
        dbuf["x"].copy_(xs_h[i % NBUF], non_blocking=True)
        four = model.log_loss(dbuf["x"])
        host_out["loss"].copy_(torch.stack(list(four)), non_blocking=True)
        host_out["ps"][0, :B].copy_(model.last_per_sample["ll_z"], non_blocking=True)
        host_out["ps"][1, :B].copy_(model.last_per_sample["ll_y"], non_blocking=True)
        host_out["ps"][2, :B].copy_(model.last_per_sample["logdet"], non_blocking=True)
        dbuf["z"].copy_(zs_h[i % NBUF], non_blocking=True)
        s = model(dbuf["z"], -1)
        host_out["x"].copy_(s, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    K, Wm = args.steps, max(args.warmup, 3)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_total = timed(step_device, K, Wm)
    clocks = sampler.stop() if rank == 0 else None
    ms_eval = timed(lambda i: model.log_loss(xs_d[i % NBUF]), K, 1)
    ms_samp = timed(lambda i: model(zs_d[i % NBUF], -1), K, 1)
    ms_e2e = timed(step_e2e, K, 2)

    # ---- dominant kernel (profiles/r01i_summary.md: 42 % of the step): the fused grouped dilated 3x3 convs of the 28x28x64
    # channel layers (gconv_oct_kernel, all dilation branches in one launch), timed ALONE with CUDA events on the stream it
    # is launched on; the 1x1-conv kernel (pw_tc3_kernel<64>, tcgen05 3xTF32) is timed the same way and reported beside it.
    layer = model.coupling_layers[2]        # block 0, mask 2: h,w,nk = 28,28,64
    info = layer._info
    hw, nk, cat = info.h * info.w, info.nk, info.cat
    u1c = torch.randn(B, info.h, info.w, info.c1, device=dev)
    layer.A_wrapper(u1c)                    # fills the layer workspace (X, Y1, Y2, LayerNorm statistics)
    ws = layer._workspace(B)

    def time_pw(which, reps=20):
        br = _lib.Borrowed()
        pp, pw_ = br(layer.params), br(ws)
        for _ in range(3):
            _lib.check(_lib.lib.cnf_measure_stage(layer._h, pp, pw_, B, which, _lib.stream_ptr()))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            _lib.check(_lib.lib.cnf_measure_stage(layer._h, pp, pw_, B, which, _lib.stream_ptr()))
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    ms_pw1 = time_pw(0)
    ms_pw2 = time_pw(1)
    ms_gc = time_pw(2)
    # algorithmic bytes per launch (both nets): read the input once, write the output once (+ residual read);
    # gamma/beta (0.4 MB per net, L2-resident across the batch) and W (16-28 KB) are not counted
    pw1_bytes = 2 * B * hw * (nk + nk) * 4
    pw2_bytes = 2 * B * hw * (cat + nk + nk) * 4
    pw_flops = 2 * 2 * B * hw * nk * (nk + cat)
    gc_bytes = 2 * B * hw * (nk + cat) * 4
    gc_macs = 2 * B * hw * sum(info.groups[i] * info.group_in[i] * info.group_out[i] * 9 for i in range(info.n_branches))
    hbm_peak, peak_src = peaks()
    pw_gbs = (pw1_bytes + pw2_bytes) / ((ms_pw1 + ms_pw2) * 1e-3) / 1e9
    ach = gc_bytes / (ms_gc * 1e-3) / 1e9
    roof = {"bound": "hbm", "kernel": "gconv_oct_kernel: grouped dilated 3x3 convs (d = 1, 2, 4; 64 -> 112 channels) of a 28x28x64 "
            "channel layer with LReLU+LayerNorm-on-load, fp32 FFMA2, both nets, B=256, one launch",
            "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
            # dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full (profiles/r01i_gconv_oct_ncu_full_summary.csv);
            # below the algorithmic bytes because part of the written tensor is still in the 126 MB L2 when the kernel ends
            "traffic": 235.8e6 if B == 256 else None,
            "peak_source": peak_src, "ms_per_launch": ms_gc, "algorithmic_bytes_per_launch": gc_bytes,
            "note": "this kernel is bound by fp32 FFMA issue / shared-memory operand traffic, not by HBM (DESIGN.md section 3): "
                    "fp32 rate below, against 148 SMs x 128 FMA/clk x 2 x 1.965 GHz = 74.5 TFLOP/s",
            "tflops_fp32": 2 * gc_macs / (ms_gc * 1e-3) / 1e12, "frac_of_fp32_ffma_peak": 2 * gc_macs / (ms_gc * 1e-3) / 74.5e12,
            "pw_tc3_kernel<64>": {"what": "1x1 convs (64->64 and 112->64 + residual), tcgen05 3xTF32, timed alone",
                                  "GB/s": pw_gbs, "frac_of_hbm_peak": pw_gbs / hbm_peak,
                                  "ms_per_launch": {"pw1_64to64": ms_pw1, "pw2_112to64_res": ms_pw2},
                                  "algorithmic_bytes_per_launch": {"pw1": pw1_bytes, "pw2": pw2_bytes},
                                  "traffic": {"pw1": 165.2e6, "pw2": 357.4e6},
                                  "tflops_fp32_equivalent": pw_flops / ((ms_pw1 + ms_pw2) * 1e-3) / 1e12}}

    # ---- the stand-alone fused coupling-law kernel (mask addressing + affine law + per-sample log-det), the
    # HBM-bound kernel the north star names: 12 bytes per element of u
    Bl, Hl, Wl, Dl = 512, 128, 128, 4
    ul = torch.randn(Bl, Hl, Wl, Dl, device=dev)
    sl = 0.1 * torch.randn(Bl, Hl, Wl, Dl // 2, device=dev)
    tl_ = torch.randn(Bl, Hl, Wl, Dl // 2, device=dev)
    vl = torch.empty_like(ul)
    ldl = torch.empty(Bl, device=dev)

    def law():
        br = _lib.Borrowed()
        _lib.check(_lib.lib.cnf_coupling_law(br(ul), br(sl), br(tl_), 2, 0, br(vl), br(ldl), _lib.stream_ptr()))
    for _ in range(3):
        law()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        law()
    e1.record()
    torch.cuda.synchronize()
    ms_law = e0.elapsed_time(e1) / 10
    law_gbs = 12 * ul.numel() / (ms_law * 1e-3) / 1e9
    del ul, sl, tl_, vl

    # ---- training step: loss_and_grad + gradient all-reduce (NCCL, N>1) + fused Adam on the same
    # batch shape.  Reported in `config` beside the headline metric; run last because it updates the weights.
    from arl_conditional_normalizing_flows_b200.conv_cINN_make_model import Adam
    model.compile(optimizer=Adam(3e-4))
    Kt = max(1, min(K, 10))
    ms_train = timed(lambda i: model.train_step(xs_d[i % NBUF]), Kt, 2)
    train_ws_gib = model._train_ws.numel() / 2 ** 30

    n_coupling = len(model.coupling_layers)
    launches_per_dir = sum(2 + 3 * l._info.R for l in model.coupling_layers)
    launches_step = (launches_per_dir + 3) + launches_per_dir     # fwd (+logdet, prior, finalize) + inverse

    if rank == 0:
        imgs = 2 * B * world
        line = {
            "metric": METRIC, "value": imgs * K / (ms_total * 1e-3), "unit": "images/s", "n_gpus": world, "steps": K,
            "warmup": Wm, "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": B, "images_per_step": imgs,
                       "weights": "trained-like random (SURVEY 8d W-rand), seed 0",
                       "l2": "no explicit flush: each step streams ~52 MB of weights and >300 MB of s/t-net activations "
                             "(> 126 MB L2) and inputs rotate over 8 distinct batches",
                       "eval_images_per_s": B * world * K / (ms_eval * 1e-3),
                       "sample_images_per_s": B * world * K / (ms_samp * 1e-3),
                       "coupling_law_kernel": {"shape": [512, 128, 128, 4], "ms": ms_law, "GB/s": law_gbs,
                                               "frac_of_hbm_peak": law_gbs / hbm_peak, "bytes_per_element": 12},
                       "train_step": {"images_per_s": B * world * Kt / (ms_train * 1e-3), "ms_per_step": ms_train / Kt,
                                      "steps": Kt, "what": "cFlow.train_step: forward with saved activations, hand-written "
                                      "backward, flat-gradient all-reduce (NCCL when N>1), fused Adam; batch "
                                      f"{B}/GPU", "activation_workspace_GiB": train_ws_gib},
                       "parallelism": f"batch-sharded x{world}; eval/sampling: no collective; training: one gradient "
                                      "all-reduce per step"},
            "e2e": {"value": imgs * K / (ms_e2e * 1e-3), "unit": "images/s",
                    "h2d_bytes_per_step": 2 * B * H * W * D * 4, "d2h_bytes_per_step": (4 + 3 * B) * 4 + B * H * W * D * 4},
            "gpu_launches": launches_step * K,
            "clocks": clocks,
            "roofline": roof,
        }
        if not args.no_cpu_baseline and world == 1:
            r = cpu_reference_run(args.cpu_sample_batch, 2, 1)
            line["cpu_baseline"] = {"value": r["value"], "unit": "images/s", "cores": r["cores"], "kind": "port",
                                    "sample": r["sample"]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
