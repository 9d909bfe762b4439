"""Out-of-bounds WRITES through the C-ABI, checked with guard bands (compute-sanitizer is not available on the GPU pool):
every buffer libcnf writes - outputs, per-sample vectors, gradients and above all the workspace, which is sized by
cnf_plan_*workspace_bytes and carved into dozens of stage buffers by the library - is a view into a larger allocation whose
64 KB on either side are filled with a byte pattern.  After forward / inverse / log_loss / the three training modes on odd
batch sizes the bands must be untouched, and the workspace is given EXACTLY the advertised number of bytes."""
import pytest
import torch

pytestmark = pytest.mark.gpu

GUARD = 64 << 10
PATTERN = 0xA5

CFG2_R1 = dict(io_shape=[28, 28, 2], x_d=1, squeeze_factor_block_list=[0, 1, 0, 0], ResNeXt_block_list=[1] * 4,
               num_kernels_list=[64, 64, 32, 32], cardinality_list=[8, 8, 4, 4])
# 32 / 16-wide groups (tensor-core grouped conv), cat = 124 / 62, planes 32x32 and 16x16
WIDE = dict(io_shape=[32, 32, 4], x_d=3, squeeze_factor_block_list=[0, 1], ResNeXt_block_list=[1, 1],
            num_kernels_list=[64, 32], cardinality_list=[2, 2])
SMALL = dict(io_shape=[8, 8, 3], x_d=2, squeeze_factor_block_list=[1, 0], ResNeXt_block_list=[2, 1],
             num_kernels_list=[16, 8], cardinality_list=[2, 2])
# a 64x64 level: the octet kernel's whole-image tile does not fit, the per-branch kernels and the halo-tile head run
LARGE = dict(io_shape=[64, 64, 2], x_d=1, squeeze_factor_block_list=[0], ResNeXt_block_list=[1],
             num_kernels_list=[32], cardinality_list=[2])


class Guarded:
    """`nbytes` usable bytes with GUARD pattern bytes on either side (the usable part starts 64 KB into a torch allocation,
    so it keeps the allocator's alignment)."""

    def __init__(self, nbytes, dev):
        self.n = int(nbytes)
        self.big = torch.full((2 * GUARD + self.n,), PATTERN, dtype=torch.uint8, device=dev)

    def bytes(self):
        return self.big[GUARD:GUARD + self.n]

    def floats(self, *shape):
        return self.bytes().view(torch.float32).view(*shape)

    def intact(self):
        front = bool((self.big[:GUARD] == PATTERN).all())
        back = bool((self.big[GUARD + self.n:] == PATTERN).all())
        return front and back


def gfloats(dev, *shape):
    n = 4
    for s in shape:
        n *= int(s)
    g = Guarded(n, dev)
    return g, g.floats(*shape)


@pytest.mark.parametrize("cfg,B", [(CFG2_R1, 5), (CFG2_R1, 1), (WIDE, 3), (SMALL, 7), (LARGE, 3)])
@pytest.mark.parametrize("paths", [0, 1, 255])      # fastest path, no activation-resident launches, generic kernels only
def test_inference_entry_points_stay_inside_their_buffers(cfg, B, paths):
    from arl_conditional_normalizing_flows_b200._lib import Borrowed, check, lib, stream_ptr
    from arl_conditional_normalizing_flows_b200.conv_cINN_make_model import cFlow
    dev = torch.device("cuda:0")
    m = cFlow(**cfg, device=dev)
    m.randomize_weights(seed=3)
    m.set_kernel_paths(paths)
    shape = (B, *cfg['io_shape'])
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(0)).to(dev)
    ws = Guarded(int(lib.cnf_plan_workspace_bytes(m._plan, B)), dev)
    gz, zy = gfloats(dev, *shape)
    gl, ld = gfloats(dev, B + 1)
    gx, xr = gfloats(dev, *shape)
    gp, pers = gfloats(dev, 3, (B + 3) & ~3)
    g4, loss4 = gfloats(dev, 4)
    guards = [ws, gz, gl, gx, gp, g4]
    br = Borrowed()
    check(lib.cnf_flow_forward(m._plan, br(x), br(m.params), br(zy), br(ld), br(ws.bytes()), stream_ptr()))
    check(lib.cnf_flow_inverse(m._plan, br(zy), br(m.params), br(xr), br(ws.bytes()), stream_ptr()))
    check(lib.cnf_flow_log_loss(m._plan, br(x), br(m.params), br(zy), br(pers[0, :B]), br(pers[1, :B]), br(pers[2, :B]),
                                br(loss4), br(ws.bytes()), stream_ptr()))
    torch.cuda.synchronize()
    assert all(g.intact() for g in guards), [g.intact() for g in guards]
    assert torch.isfinite(zy).all() and torch.isfinite(loss4).all()
    # and the calls did what the public API does
    want = m(x, +1)[0]
    assert torch.equal(zy, want)


@pytest.mark.parametrize("cfg,B", [(CFG2_R1, 5), (WIDE, 3), (SMALL, 7), (LARGE, 2)])
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_training_entry_points_stay_inside_their_buffers(cfg, B, mode):
    from arl_conditional_normalizing_flows_b200._lib import Borrowed, check, lib, stream_ptr
    from arl_conditional_normalizing_flows_b200.conv_cINN_make_model import cFlow
    dev = torch.device("cuda:0")
    m = cFlow(**cfg, device=dev)
    m.randomize_weights(seed=3)
    shape = (B, *cfg['io_shape'])
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(1)).to(dev)
    need = [lib.cnf_plan_train_workspace_bytes, lib.cnf_plan_train_workspace_bytes_recompute,
            lib.cnf_plan_train_workspace_bytes_invert][mode](m._plan, B)
    ws = Guarded(int(need), dev)
    gz, zy = gfloats(dev, *shape)
    gg, grads = gfloats(dev, m.params.numel())
    gp, pers = gfloats(dev, 3, (B + 3) & ~3)
    g4, loss4 = gfloats(dev, 4)
    gm, mom = gfloats(dev, 2, m.params.numel())
    mom.zero_()
    guards = [ws, gz, gg, gp, g4, gm]
    seen = []
    from arl_conditional_normalizing_flows_b200._lib import LAYER_GRADS_READY_FN
    cb = LAYER_GRADS_READY_FN(lambda _u, layer, off, count: seen.append(layer))
    br = Borrowed()
    check(lib.cnf_flow_loss_and_grad_hooked(m._plan, br(x), br(m.params), br(grads), br(zy), br(pers[0, :B]), br(pers[1, :B]),
                                            br(pers[2, :B]), br(loss4), br(ws.bytes()), stream_ptr(), mode, cb, None))
    params = m.params.clone()
    check(lib.cnf_adam_step(br(params), br(grads), br(mom[0]), br(mom[1]), 1, 3e-4, 0.9, 0.999, 1e-7, 1.0, stream_ptr()))
    torch.cuda.synchronize()
    assert all(g.intact() for g in guards), [g.intact() for g in guards]
    assert len(seen) == len(m.coupling_layers)
    assert torch.isfinite(grads).all() and torch.isfinite(loss4).all() and torch.isfinite(params).all()
