#!/usr/bin/env python
"""Benchmark of the conditional RealNVP hot path (BASELINE.json metric: images/sec for log-likelihood
eval and sampling).

    python bench.py --gpus N --steps K --warmup W [--config 2|3|4|4-heavy|5|5-heavy]   # this repo's CUDA path
    python bench.py --impl reference --gpus N ...                                       # the reference's CPU path (oracle port)

Default workload (N=1): BASELINE config 2 - conv_cINN class-conditioned 28x28x1 image + 1 label plane
(io_shape [28,28,2], x_d 1, lists [0,1,0,0]/[3,3,3,3]/[64,64,32,32]/[8,8,4,4]), batch 256 per GPU,
synthetic inputs and "trained-like" random weights.  One STEP = one log-likelihood
evaluation (`log_loss`) of a batch + one sampling pass (`call(zy,-1)`) of a batch; images/sec counts the
images through both (2*B per step).  N>1: one process per GPU (torchrun), the batch dimension is sharded
(weak scaling: the config's batch per GPU), no data-path collective; time is the max over ranks.
--config selects the other BASELINE configs (builder-chosen hyper-parameters where the reference
wires none): 3 = 32x32x4 class-conditioned, 4 = 64x64x6 super-resolution (light lists; 4-heavy: 128/64 kernels),
5 = 128x128x4 noise pre-training (light; 5-heavy: 256/128 kernels).  Their line has the same keys; `config.train_step`
is the training throughput (the named workload of configs 4 and 5).
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

_L = dict(squeeze_factor_block_list=[0, 1, 0, 0], ResNeXt_block_list=[3, 3, 3, 3])
CONFIGS = {
    "2": dict(cfg=dict(io_shape=[28, 28, 2], x_d=1, num_kernels_list=[64, 64, 32, 32], cardinality_list=[8, 8, 4, 4], **_L),
              batch=256, synth="cfg2",
              workload="cfg2: conv_cINN class-conditioned 28x28x1 + label plane (io 28x28x2), batch 256/GPU, log_loss + sampling"),
    "3": dict(cfg=dict(io_shape=[32, 32, 4], x_d=3, num_kernels_list=[64, 64, 32, 32], cardinality_list=[8, 8, 4, 4], **_L),
              batch=256, synth="cfg3",
              workload="cfg3: conv_cINN class-conditioned 32x32x3 + label plane (io 32x32x4), batch 256/GPU, log_loss + sampling"),
    "4": dict(cfg=dict(io_shape=[64, 64, 6], x_d=3, num_kernels_list=[64, 64, 32, 32], cardinality_list=[4, 4, 2, 2], **_L),
              batch=64, synth="cfg4",
              workload="cfg4 (light lists 64/32 kernels, cardinality 4/2): conv_cINN super-resolution 64x64x3 with 8x8 condition "
                       "(io 64x64x6), batch 64/GPU, log_loss + sampling; training in config.train_step"),
    "4-heavy": dict(cfg=dict(io_shape=[64, 64, 6], x_d=3, num_kernels_list=[128, 128, 64, 64], cardinality_list=[8, 8, 4, 4], **_L),
                    batch=32, synth="cfg4",
                    workload="cfg4 heavy (128/64 kernels, cardinality 8/4): super-resolution 64x64x3 (io 64x64x6), batch 32/GPU"),
    "5": dict(cfg=dict(io_shape=[128, 128, 4], x_d=3, num_kernels_list=[64, 64, 32, 32], cardinality_list=[2, 2, 2, 2], **_L),
              batch=16, synth="cfg5",
              workload="cfg5 (light lists 64/32 kernels, cardinality 2): noise pre-training 128x128x3 + label plane (io 128x128x4), "
                       "batch 16/GPU, log_loss + sampling; training in config.train_step"),
    "5-heavy": dict(cfg=dict(io_shape=[128, 128, 4], x_d=3, num_kernels_list=[256, 256, 128, 128], cardinality_list=[8, 8, 4, 4], **_L),
                    batch=4, synth="cfg5",
                    workload="cfg5 heavy (256/128 kernels, cardinality 8/4): noise pre-training 128x128x4, batch 4/GPU"),
}
METRIC = "images/sec (log-likelihood eval + sampling)"
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12      # FFMA2: 128 fp32 FMA lanes per SM and clock; computed, not measured


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="2", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=0, help="images per GPU per step (0: the config's batch)")
    ap.add_argument("--streams", type=int, default=2, choices=[1, 2],
                    help="2: log_loss and sampling of a step on two CUDA streams (cFlow.log_loss_and_sample); 1: one after the other")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step measurement")
    ap.add_argument("--quick", action="store_true", help="headline + e2e + roofline (+ training) only: skips the eval-only / "
                    "sampling-only / layer-per-kernel / strong-scaling side measurements (multi-GPU lines of the other configs)")
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


def ncu_metric(kernel, key):
    """one metric of the committed ncu capture of `kernel` (profiles/r02_traffic.json), or None"""
    p = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if not os.path.exists(p):
        return None
    return json.load(open(p)).get(kernel, {}).get(key)


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel`, from the committed ncu capture
    (profiles/r02_traffic.json, written by tools/ncu_traffic.py from an `ncu --set full` report); None if absent."""
    p = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if not os.path.exists(p):
        return None, None
    d = json.load(open(p))
    e = d.get(kernel)
    return (e["dram_bytes"], e["source"]) if e else (None, None)


def synth_host(kind, cfg, B, seed):
    """synthetic xy batch as a host tensor, generated with torch (the product's bench owns its inputs)"""
    import torch
    g = torch.Generator().manual_seed(seed)
    H, W, D = cfg["io_shape"]
    xd = cfg["x_d"]
    if kind in ("cfg2", "cfg3"):
        img = 0.98 * torch.rand((B, H, W, xd), generator=g) + 0.02 * torch.randn((B, H, W, xd), generator=g)
        label = float(torch.randint(0, 10, (1,), generator=g)) / 9.0
        lab = 0.98 * torch.full((B, H, W, D - xd), label) + 0.02 * torch.randn((B, H, W, D - xd), generator=g)
        return torch.cat([img, lab], 3).contiguous()
    if kind == "cfg4":
        hr = torch.rand((B, H, W, xd), generator=g)
        y = hr.reshape(B, H // 8, 8, W // 8, 8, xd).mean((2, 4), keepdim=True).expand(B, H // 8, 8, W // 8, 8, xd).reshape(B, H, W, xd)
        xy = torch.cat([hr - y, y], 3)
        return (0.98 * xy + 0.02 * torch.randn(xy.shape, generator=g)).contiguous()
    return torch.randn((B, H, W, D), generator=g)


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference's CPU path
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(key, batch, steps, warmup, budget_s=150.0):
    import torch
    from oracle.flow_torch import FlowOracle
    from oracle.weights import init_weights
    C = CONFIGS[key]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    o = FlowOracle(**C["cfg"], dtype=torch.float32)
    o.set_weights(init_weights(o.plan, 'rand', seed=0))
    x = synth_host(C["synth"], C["cfg"], batch, 0).numpy()
    z = torch.randn(x.shape, generator=torch.Generator().manual_seed(1)).numpy()
    z[..., C["cfg"]["x_d"]:] = x[..., C["cfg"]["x_d"]:]
    times = []
    t_start = time.perf_counter()
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        o.log_loss(x)
        o.call(z, -1)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if times and time.perf_counter() - t_start > budget_s:
            break
    total = sum(times)
    return {"value": 2 * batch * len(times) / total, "ms_per_step": 1e3 * total / len(times), "cores": cores,
            "steps": len(times), "torch_threads": torch.get_num_threads(),
            "sample": f"{len(times)} steps of log_loss + call(-1) on a batch of {batch} {C['synth']} images, the bench's own "
                      f"config and batch (oracle port, torch {torch.__version__} CPU fp32, {cores} threads, {warmup} warm-up)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    C = CONFIGS[args.config]
    B = args.batch or C["batch"]
    r = cpu_reference_run(args.config, B, max(1, min(args.steps, 3)), 1)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "images/s", "n_gpus": args.gpus,
        "steps": r["steps"], "warmup": 1, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": C["workload"], "batch_per_gpu": B, "images_per_step": 2 * B,
                   "note": "reference CPU path: TensorFlow is not installable here, so this is the repo's oracle port of the "
                           "reference (restated reference on CPU, not TensorFlow) on the bench's own config and batch; the "
                           "number of timed steps is bounded (at most 3, at most ~150 s)"},
        "cpu_baseline": {"value": r["value"], "unit": "images/s", "cores": r["cores"], "kind": "port",
                         "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from arl_conditional_normalizing_flows_b200.conv_cINN_make_model import cFlow
    from arl_conditional_normalizing_flows_b200 import _lib

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

    C = CONFIGS[args.config]
    CFG = C["cfg"]
    B = args.batch or C["batch"]
    H, W, D = CFG["io_shape"]
    xd = CFG["x_d"]
    model = cFlow(**CFG, device=dev)
    model.randomize_weights(seed=0)          # trained-like random weights, identical on every rank

    # rotating synthetic inputs (different batch every step; shards differ per rank)
    NBUF = 8 if B * H * W * D * 4 < (64 << 20) else 3
    xs, zs = [], []
    for i in range(NBUF):
        x = synth_host(C["synth"], CFG, B, 1000 * rank + i)
        z = torch.randn(x.shape, generator=torch.Generator().manual_seed(5000 + 1000 * rank + i))
        z[..., xd:] = x[..., xd:]
        xs.append(x)
        zs.append(z)
    xs_d = [t.to(dev) for t in xs]
    zs_d = [t.to(dev) for t in zs]
    xs_h = [t.pin_memory() for t in xs]
    zs_h = [t.pin_memory() for t in zs]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_serial(i):
        model.log_loss(xs_d[i % NBUF])
        model(zs_d[i % NBUF], -1)

    def step_device(i):
        # the same two calls issued on two CUDA streams (cFlow.log_loss_and_sample): one pass's CTAs fill the SMs the
        # other pass leaves idle at every kernel boundary
        if args.streams == 2:
            model.log_loss_and_sample(xs_d[i % NBUF], zs_d[i % NBUF])
        else:
            step_serial(i)

    host_out = {"loss": torch.empty(4).pin_memory(), "ps": torch.empty((3, (B + 3) & ~3)).pin_memory(),
                "x": torch.empty((B, H, W, D)).pin_memory()}
    dbuf = {"x": torch.empty((B, H, W, D), device=dev), "z": torch.empty((B, H, W, D), device=dev)}

    def step_e2e(i):
        # host buffers in, host results out, through the public API
        dbuf["x"].copy_(xs_h[i % NBUF], non_blocking=True)
        if args.streams == 2:
            dbuf["z"].copy_(zs_h[i % NBUF], non_blocking=True)
            four, s = model.log_loss_and_sample(dbuf["x"], dbuf["z"])
        else:
            four = model.log_loss(dbuf["x"])
        host_out["loss"].copy_(torch.stack(list(four)), non_blocking=True)
        host_out["ps"][0, :B].copy_(model.last_per_sample["ll_z"], non_blocking=True)
        host_out["ps"][1, :B].copy_(model.last_per_sample["ll_y"], non_blocking=True)
        host_out["ps"][2, :B].copy_(model.last_per_sample["logdet"], non_blocking=True)
        if args.streams != 2:
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
    ms_e2e = timed(step_e2e, K, 2)
    ms_serial = timed(step_serial, K, 1) if args.streams == 2 else ms_total     # the two calls one after the other
    Bs = max(1, B // world)
    ms_eval = ms_samp = ms_unfused = ms_strong = float("nan")
    if not args.quick:
        ms_eval = timed(lambda i: model.log_loss(xs_d[i % NBUF]), K, 1)
        ms_samp = timed(lambda i: model(zs_d[i % NBUF], -1), K, 1)
        # the same step with every layer on the layer-per-kernel path (what the activation-resident launches replace)
        model.set_fusion(0)
        ms_unfused = timed(step_serial, K, 1)
        model.set_fusion(1)
        # strong scaling: the config's batch as the GLOBAL batch, split over the ranks
        ms_strong = timed(lambda i: (model.log_loss(xs_d[i % NBUF][:Bs]), model(zs_d[i % NBUF][:Bs], -1)), K, 1)

    # ---- dominant kernel: the grouped dilated 3x3 convs of the first full-resolution channel layer (config 2: 28x28x64,
    # gconv_oct_kernel, all dilation branches in one launch; profiles/r02_summary.md), timed ALONE with CUDA events on the
    # stream it is launched on; the 1x1-conv kernel (pw_tc3_kernel, tcgen05 3xTF32) is timed the same way beside it.
    layer = model.coupling_layers[2]        # block 0, mask 2: full resolution, nk = num_kernels_list[0]
    info = layer._info
    hw, nk, cat = info.h * info.w, info.nk, info.cat
    u1c = torch.randn(B, info.h, info.w, info.c1, device=dev)
    layer.A_wrapper(u1c)                    # fills the layer workspace (X, Y1, Y2, LayerNorm statistics)
    ws = layer._workspace(B)

    def time_stage(which, reps=20):
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

    ms_pw1 = time_stage(0)
    ms_pw2 = time_stage(1)
    ms_gc = time_stage(2)
    # algorithmic bytes per launch (both nets): read the input once, write the output once (+ residual read);
    # gamma/beta (L2-resident across the batch) and W are not counted
    pw1_bytes = 2 * B * hw * (nk + nk) * 4
    pw2_bytes = 2 * B * hw * (cat + nk + nk) * 4
    pw_flops = 2 * 2 * B * hw * nk * (nk + cat)
    gc_bytes = 2 * B * hw * (nk + cat) * 4
    gc_macs = 2 * B * hw * sum(info.groups[i] * info.group_in[i] * info.group_out[i] * 9 for i in range(info.n_branches))
    hbm_peak, peak_src = peaks()
    pw_gbs = (pw1_bytes + pw2_bytes) / ((ms_pw1 + ms_pw2) * 1e-3) / 1e9
    gc_gbs = gc_bytes / (ms_gc * 1e-3) / 1e9
    gc_tflops = 2 * gc_macs / (ms_gc * 1e-3) / 1e12
    is_cfg2 = args.config == "2" and B == 256
    tr_gc, tr_src = ncu_traffic("gconv_oct_kernel") if is_cfg2 else (None, None)
    tr_pw1, _ = ncu_traffic("pw_tc3_kernel<64> pw1") if is_cfg2 else (None, None)
    tr_pw2, _ = ncu_traffic("pw_tc3_kernel<64> pw2") if is_cfg2 else (None, None)
    roof = {"bound": "fp32-issue",
            "kernel": f"grouped dilated 3x3 convs ({info.n_branches} dilation branches, {nk} -> {cat} channels) of the {info.h}x{info.w}x{nk} "
                      "channel layer with LReLU+LayerNorm-on-load, fp32 FFMA2, both nets, one stage of one coupling layer "
                      "(gconv_oct_kernel where the image tile fits shared memory, else one gconv3_kernel launch per branch)",
            "achieved": gc_tflops, "peak": FP32_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": gc_tflops / FP32_PEAK_TFLOPS,
            "peak_source": "computed: 148 SMs x 128 fp32 FMA lanes x 2 x 1.965 GHz (MEASURED_PEAKS.json has no fp32 figure; the "
                           "kernel is bound by fp32 issue / shared-memory operand traffic, not by HBM or the tensor pipe, "
                           "DESIGN.md section 3)",
            "traffic": tr_gc, "traffic_source": tr_src,
            "ms_per_launch": ms_gc, "algorithmic_flops_per_launch": 2 * gc_macs,
            "hbm": {"achieved": gc_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gc_gbs / hbm_peak, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": gc_bytes},
            "pw_tc3_kernel": {"what": f"1x1 convs ({nk}->{nk} and {cat}->{nk} + residual) of the same layer, tcgen05 3xTF32 where N is "
                                      "16/32/64, timed alone; bound: HBM",
                              "GB/s": pw_gbs, "frac_of_hbm_peak": pw_gbs / hbm_peak,
                              "ms_per_launch": {"pw1": ms_pw1, "pw2_res": ms_pw2},
                              "algorithmic_bytes_per_launch": {"pw1": pw1_bytes, "pw2": pw2_bytes},
                              "traffic": {"pw1": tr_pw1, "pw2": tr_pw2},
                              # tensor-pipe duty of the committed ncu capture (the north star's metric for the convs): the
                              # kernel is paced by its operand hand-offs and HBM, not by the tensor pipe
                              "ncu_tensor_pipe_active_pct_of_peak": {
                                  w: ncu_metric(f"pw_tc3_kernel<64> {w}", "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed")
                                  for w in ("pw1", "pw2")} if is_cfg2 else None,
                              "ncu_umma_operand_read_active_pct": {
                                  w: ncu_metric(f"pw_tc3_kernel<64> {w}", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed")
                                  for w in ("pw1", "pw2")} if is_cfg2 else None,
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
    train = None
    if not args.no_train:
        from arl_conditional_normalizing_flows_b200.conv_cINN_make_model import Adam
        model.compile(optimizer=Adam(3e-4))
        Kt = max(1, min(K, 10))
        try:
            ms_train = timed(lambda i: model.train_step(xs_d[i % NBUF]), Kt, 2)
            train = {"images_per_s": B * world * Kt / (ms_train * 1e-3), "ms_per_step": ms_train / Kt, "steps": Kt,
                     "what": "cFlow.train_step: forward with saved activations, hand-written backward, gradient all-reduce in "
                             f"buckets of >= {model.grad_bucket_bytes >> 20} MiB of coupling layers issued while the backward "
                             f"pass runs (NCCL when N>1), fused Adam; batch {B}/GPU",
                     "activation_workspace_GiB": model._train_ws.numel() / 2 ** 30}
            if world > 1:     # the same step with ONE all-reduce after the backward pass, for comparison
                bucket = model.grad_bucket_bytes
                model.grad_bucket_bytes = 0
                ms_flat = timed(lambda i: model.train_step(xs_d[i % NBUF]), Kt, 2)
                model.grad_bucket_bytes = bucket
                train["ms_per_step_flat_allreduce"] = ms_flat / Kt
        except NotImplementedError as e:      # e.g. group widths above 8 (config 5): backward kernels not built
            train = {"unsupported": str(e)}

    # kernels of this repo launched per step: per pass one launch per activation-resident layer, 2 + 3R per other layer; the
    # forward pass adds the state copy, log-det finalize, prior/L1 and loss-finalize kernels, the inverse pass the state copy
    per_pass = model.launches_per_pass()
    launches_step = (per_pass + 4) + (per_pass + 1)

    if rank == 0:
        imgs = 2 * B * world
        line = {
            "metric": METRIC, "value": imgs * K / (ms_total * 1e-3), "unit": "images/s", "n_gpus": world, "steps": K,
            "warmup": Wm, "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": C["workload"], "batch_per_gpu": B, "images_per_step": imgs,
                       "weights": "trained-like random (SURVEY 8d W-rand; cFlow.randomize_weights, seed 0)",
                       "l2": "no explicit flush: each step streams the model's weights and the s/t-net activations of the "
                             f"full-resolution layers (> 126 MB L2 at this batch) and inputs rotate over {NBUF} distinct batches",
                       "streams": f"{args.streams}: " + ("log_loss and sampling of a step issued on two CUDA streams "
                                                         "(cFlow.log_loss_and_sample); the same kernels one call after the other: "
                                                         "serial_images_per_s" if args.streams == 2 else "one call after the other"),
                       "serial_images_per_s": imgs * K / (ms_serial * 1e-3),
                       "eval_images_per_s": None if args.quick else B * world * K / (ms_eval * 1e-3),
                       "sample_images_per_s": None if args.quick else B * world * K / (ms_samp * 1e-3),
                       "layer_per_kernel_path_images_per_s": None if args.quick else imgs * K / (ms_unfused * 1e-3),
                       "resident_layers": f"{sum(l.resident_kernel_eligible() for l in model.coupling_layers)} of "
                                          f"{len(model.coupling_layers)} coupling layers run as one activation-resident launch",
                       "strong_scaling": None if args.quick else {"global_batch": Bs * world, "batch_per_gpu": Bs,
                                                                  "images_per_s": 2 * Bs * world * K / (ms_strong * 1e-3)},
                       "coupling_law_kernel": {"shape": [512, 128, 128, 4], "ms": ms_law, "GB/s": law_gbs,
                                               "frac_of_hbm_peak": law_gbs / hbm_peak, "bytes_per_element": 12},
                       "train_step": train,
                       "parallelism": f"batch-sharded x{world}; eval/sampling: no collective; training: gradient "
                                      "all-reduce per bucket of coupling layers, overlapped with the backward pass"},
            "e2e": {"value": imgs * K / (ms_e2e * 1e-3), "unit": "images/s",
                    "h2d_bytes_per_step": 2 * B * H * W * D * 4, "d2h_bytes_per_step": (4 + 3 * B) * 4 + B * H * W * D * 4},
            "gpu_launches": launches_step * K,
            "clocks": clocks,
            "roofline": roof,
        }
        if not args.no_cpu_baseline and world == 1:
            r = cpu_reference_run(args.config, B, 2, 1, budget_s=60.0)
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
