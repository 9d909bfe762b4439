"""One warm-up + one training step of a BASELINE config (for the ncu launch list): python tools/profile_train_cfg.py 4|5 [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from arl_conditional_normalizing_flows_b200.conv_cINN_make_model import cFlow, Adam
key = sys.argv[1] if len(sys.argv) > 1 else "4"
C = bench.CONFIGS[key]
B = int(sys.argv[2]) if len(sys.argv) > 2 else C["batch"]
m = cFlow(**C["cfg"], device="cuda:0")
m.randomize_weights(seed=0)
m.compile(optimizer=Adam(3e-4))
x = bench.synth_host(C["synth"], C["cfg"], B, 0).cuda()
m.train_step(x)
m.train_step(x)
torch.cuda.synchronize()
