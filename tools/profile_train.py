"""Times one config-2 training step (loss_and_grad + Adam) at batch B and prints the forward-only time beside it.
Used under ncu for the launch list:  python tools/profile_train.py [B] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from arl_conditional_normalizing_flows_b200.conv_cINN_make_model import cFlow, Adam
from oracle.planner import plan_flow  # noqa: F401  (weights helper needs the plan)
from oracle.weights import init_weights, synth_inputs
from oracle.flow_torch import FlowOracle

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
CFG2 = dict(io_shape=[28, 28, 2], x_d=1, squeeze_factor_block_list=[0, 1, 0, 0], ResNeXt_block_list=[3, 3, 3, 3],
            num_kernels_list=[64, 64, 32, 32], cardinality_list=[8, 8, 4, 4])
m = cFlow(**CFG2, device="cuda:0")
m.set_weights(init_weights(FlowOracle(**CFG2).plan, 'rand', seed=0))
m.compile(optimizer=Adam(3e-4))
xy = torch.from_numpy(synth_inputs('cfg2', B, seed=0)).cuda()


def timed(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


if len(sys.argv) > 3 and sys.argv[3] == 'once':     # under ncu: one warm-up step + one step
    m.train_step(xy)
    m.train_step(xy)
    torch.cuda.synchronize()
    sys.exit(0)
t_f = timed(lambda: m.log_loss(xy))
t_g = timed(lambda: m.loss_and_grad(xy))
t_s = timed(lambda: m.train_step(xy))
print(f"B={B}: log_loss {t_f:.2f} ms, loss_and_grad {t_g:.2f} ms, train_step {t_s:.2f} ms "
      f"-> {B / t_s * 1e3:.0f} images/s training; train ws {m._train_ws.numel() / 2**30:.2f} GiB")
