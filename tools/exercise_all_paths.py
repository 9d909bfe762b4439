"""One small invocation of every kernel family of libcnf on cuda:0 -- the activation-resident coupling-layer kernel, the
layer-per-kernel path (stem, tcgen05 1x1, octet grouped conv with TMA tiles, head) with each kernel family excluded in turn,
the tensor-core grouped conv (16-wide groups), the three training modes with the layer hook and the Adam step, the toy model
and the data kernels.  Written as the workload for compute-sanitizer; that tool is closed on this GPU pool (it answers
"closed on this pool"), so out-of-bounds writes are checked with guard bands instead (tests/test_gpu_guard_bands.py) and
this script remains as a quick all-paths smoke run:   python tools/exercise_all_paths.py [cfg2 paths wide train toy data]"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from arl_conditional_normalizing_flows_b200 import _lib  # noqa: E402
from arl_conditional_normalizing_flows_b200.conv_cINN_make_model import Adam, cFlow  # noqa: E402

DEV = torch.device("cuda:0")
WHAT = set(sys.argv[1:]) or {"cfg2", "paths", "wide", "train", "toy", "data"}

CFG2 = dict(io_shape=[28, 28, 2], x_d=1, squeeze_factor_block_list=[0, 1, 0, 0], ResNeXt_block_list=[1, 1, 1, 1],
            num_kernels_list=[64, 64, 32, 32], cardinality_list=[8, 8, 4, 4])
# 16-wide groups at 32x32 (tensor-core grouped conv), 124-channel concat (K % 8 == 4 on the tcgen05 1x1 kernel)
WIDE = dict(io_shape=[32, 32, 4], x_d=3, squeeze_factor_block_list=[0, 1], ResNeXt_block_list=[1, 1],
            num_kernels_list=[64, 32], cardinality_list=[2, 2])
SMALL = dict(io_shape=[8, 8, 3], x_d=2, squeeze_factor_block_list=[1, 0], ResNeXt_block_list=[2, 1],
             num_kernels_list=[16, 8], cardinality_list=[2, 2])


def batch(cfg, B, seed=0):
    return torch.randn(B, *cfg['io_shape'], generator=torch.Generator().manual_seed(seed)).to(DEV)


def model(cfg, seed=3):
    m = cFlow(**cfg, device=DEV)
    m.randomize_weights(seed=seed)
    return m


def infer(m, x, tag):
    four = [float(t) for t in m.log_loss(x)]
    z = m(x, +1)
    z = z[0] if isinstance(z, (tuple, list)) else z
    s = m(z, -1)
    torch.cuda.synchronize()
    err = float((s - x).abs().max())
    print(f"{tag}: loss {four[0]:.5f}, round trip max err {err:.2e}", flush=True)


if "cfg2" in WHAT:
    m = model(CFG2)
    infer(m, batch(CFG2, 3), "cfg2 shapes (resident + layer-per-kernel, octet TMA tiles, tcgen05 1x1)")
if "paths" in WHAT:
    m = model(CFG2)
    x = batch(CFG2, 2)
    for name, bits in (("no resident", _lib.CNF_PATH_NO_RESIDENT),
                       ("no resident, no tcgen05", _lib.CNF_PATH_NO_RESIDENT | _lib.CNF_PATH_NO_TCGEN05),
                       ("no resident, no octet", _lib.CNF_PATH_NO_RESIDENT | _lib.CNF_PATH_NO_OCTET),
                       ("generic kernels only", _lib.CNF_PATH_ALL)):
        m.set_kernel_paths(bits)
        infer(m, x, f"cfg2 shapes, {name}")
if "wide" in WHAT:
    m = model(WIDE)
    infer(m, batch(WIDE, 2), "16-wide groups (tensor-core grouped conv)")
if "train" in WHAT:
    for cfg, B, tag in ((CFG2, 3, "cfg2 shapes"), (WIDE, 2, "16-wide groups"), (SMALL, 4, "small")):
        for mode in (0, 1, 2):
            m = model(cfg)
            m.recompute_activations = mode == 1
            m.recover_states_by_inverse = mode == 2
            m.compile(optimizer=Adam(3e-4))
            x = batch(cfg, B, seed=1)
            seen = []
            m.loss_and_grad(x, on_layer_grads=lambda layer, off, count: seen.append(layer))
            logs = [float(m.train_step(x)['loss']) for _ in range(2)]
            torch.cuda.synchronize()
            print(f"train {tag}, mode {mode}: losses {logs[0]:.5f} {logs[1]:.5f}, hook saw {len(seen)} layers", flush=True)
def guarded(fn):
    try:
        fn()
    except Exception as e:      # noqa: BLE001 -- a host-side usage error of this script must not hide the tool's report
        print(f"{fn.__name__}: host-side error {type(e).__name__}: {e}", flush=True)


def toy():
    from arl_conditional_normalizing_flows_b200.TOYcINN_make_model import cINN_affine
    t = cINN_affine(3, 2, 6, 32, 3, 'glorot_uniform', device=DEV, seed=0)
    xy = torch.randn(64, 3, generator=torch.Generator().manual_seed(2)).to(DEV)
    four = [float(v) for v in t.log_loss(xy)]
    t.call(xy, -1)
    t.compile(optimizer=Adam(1e-3))
    t.train_step(xy)
    torch.cuda.synchronize()
    print(f"toy: loss {four[0]:.5f}", flush=True)


def data():
    from arl_conditional_normalizing_flows_b200 import conv_cINN_base_functions as F
    hr = torch.rand(2, 64, 64, 3, generator=torch.Generator().manual_seed(5)).to(DEV)
    lo = F.down(hr, 3)
    F.up(lo, 3)
    F.preprocess_dataset_SR(hr, 'cINN', RESIDUAL=True, levels=(0, 3))
    F.preprocess_dataset_class(torch.rand(4, 28, 28, 1, generator=torch.Generator().manual_seed(6)).to(DEV), LOGITS=True)
    F.instance_noise(hr, 0.05, seed=7)
    torch.cuda.synchronize()
    print("data kernels done", flush=True)


if "toy" in WHAT:
    guarded(toy)
if "data" in WHAT:
    guarded(data)
print("sanitize workload done", flush=True)
