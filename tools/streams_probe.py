"""How much of the idle time at kernel boundaries do more CUDA streams recover?  One evaluation + sampling step of config 2
(batch 256) issued as 1, 2 (cFlow.log_loss_and_sample) or 4 streams (each pass split into two half batches).
usage: python tools/streams_probe.py [B]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arl_conditional_normalizing_flows_b200.conv_cINN_make_model import cFlow  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda:0")
cfg = dict(io_shape=[28, 28, 2], x_d=1, squeeze_factor_block_list=[0, 1, 0, 0], ResNeXt_block_list=[3, 3, 3, 3],
           num_kernels_list=[64, 64, 32, 32], cardinality_list=[8, 8, 4, 4])
m = cFlow(**cfg, device=dev)
m.randomize_weights(seed=0)
g = torch.Generator().manual_seed(0)
xs = [torch.randn(B, 28, 28, 2, generator=g).to(dev) for _ in range(4)]
zs = [torch.randn(B, 28, 28, 2, generator=g).to(dev) for _ in range(4)]
side = [torch.cuda.Stream(device=dev) for _ in range(3)]


def serial(i):
    m.log_loss(xs[i % 4])
    m(zs[i % 4], -1)


def two(i):
    m.log_loss_and_sample(xs[i % 4], zs[i % 4])


def four(i):
    main = torch.cuda.current_stream()
    x, z = xs[i % 4], zs[i % 4]
    h = B // 2
    work = [(lambda: m(z[:h], -1)), (lambda: m(z[h:], -1)), (lambda: m.log_loss(x[h:]))]
    for s, fn in zip(side, work):
        s.wait_stream(main)
        with torch.cuda.stream(s):
            fn()
    m.log_loss(x[:h])
    for s in side:
        main.wait_stream(s)


def timed(fn, steps=20, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(warm + i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


for name, fn in (("1 stream", serial), ("2 streams", two), ("4 streams (half batches)", four), ("2 streams", two)):
    ms = timed(fn)
    print(f"{name}: {ms:.3f} ms per step, {2 * B / ms * 1e3:.0f} images/s", flush=True)
