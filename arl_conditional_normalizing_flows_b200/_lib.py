"""ctypes binding of libcnf.so (the C-ABI declared in include/cnf.h).

Tensors cross the boundary zero-copy as DLPack capsules: `torch.utils.dlpack.to_dlpack(t)` yields a
PyCapsule named "dltensor" whose pointer is a `DLManagedTensor*`; libcnf borrows it for the call
and never runs the deleter (the capsule's own destructor frees it when Python drops it).

There is NO fallback: if the shared library is missing, importing this module raises.
"""
import ctypes
import os
from ctypes import (POINTER, Structure, byref, c_char, c_char_p, c_double, c_int, c_int64, c_uint64, c_void_p,
                    py_object)

import torch
from torch.utils.dlpack import to_dlpack

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcnf.so")

CNF_MAX_BRANCHES = 8
CNF_NAME_CAP = 64
# kernel families a coupling layer must NOT use (include/cnf.h, cnf_*_set_kernel_paths)
(CNF_PATH_NO_RESIDENT, CNF_PATH_NO_TCGEN05, CNF_PATH_NO_PW_FFMA, CNF_PATH_NO_OCTET, CNF_PATH_NO_BRANCH,
 CNF_PATH_NO_GCONV2, CNF_PATH_NO_STEM2, CNF_PATH_NO_HEAD2) = (1, 2, 4, 8, 16, 32, 64, 128)
CNF_PATH_ALL = 255

(CNF_OK, CNF_ERR_ARG, CNF_ERR_SHAPE, CNF_ERR_DTYPE, CNF_ERR_DEVICE, CNF_ERR_LAYOUT, CNF_ERR_CUDA,
 CNF_ERR_UNSUPPORTED, CNF_ERR_WORKSPACE) = range(9)

# error class -> the exception type the reference raises for the same condition
_EXC = {
    CNF_ERR_ARG: AssertionError,          # python `assert` in the reference
    CNF_ERR_SHAPE: ValueError,            # tf.ensure_shape
    CNF_ERR_DTYPE: TypeError,
    CNF_ERR_DEVICE: RuntimeError,
    CNF_ERR_LAYOUT: ValueError,
    CNF_ERR_CUDA: RuntimeError,
    CNF_ERR_UNSUPPORTED: NotImplementedError,
    CNF_ERR_WORKSPACE: ValueError,
}


class CouplingInfo(Structure):
    _fields_ = [(n, c_int) for n in ("H", "W", "D", "mask", "mask_complement", "R", "cardinality", "nk",
                                     "ksize", "layer_norm", "h", "w", "c1", "c2", "cat", "n_branches")] + [
        ("dilation", c_int * CNF_MAX_BRANCHES), ("branch_channels", c_int * CNF_MAX_BRANCHES),
        ("groups", c_int * CNF_MAX_BRANCHES), ("group_in", c_int * CNF_MAX_BRANCHES),
        ("group_out", c_int * CNF_MAX_BRANCHES), ("n_ln", c_int), ("net_stride", c_int64),
        ("param_count", c_int64), ("n_entries", c_int)]


class PlanInfo(Structure):
    _fields_ = [(n, c_int) for n in ("n_blocks", "n_coupling", "n_layers", "H", "W", "D", "x_d", "ksize",
                                     "layer_norm")] + [("lambda_y", c_double), ("param_count", c_int64)]


class BlockInfo(Structure):
    _fields_ = [("scale", c_int), ("num_prev_factors", c_int), ("H", c_int), ("W", c_int), ("D", c_int),
                ("n_checkerboard", c_int), ("checkerboard", c_int * CNF_MAX_BRANCHES),
                ("n_channelwise", c_int), ("channelwise", c_int * CNF_MAX_BRANCHES)]


# cnf_layer_grads_ready_fn (include/cnf.h): void (*)(void* user, int layer, int64_t param_offset, int64_t param_count)
LAYER_GRADS_READY_FN = ctypes.CFUNCTYPE(None, c_void_p, c_int, c_int64, c_int64)


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU or PyTorch fallback for this path.")
    lib = ctypes.CDLL(LIB_PATH)
    P, I, I64, D, V = c_void_p, c_int, c_int64, c_double, c_void_p
    IP = POINTER(c_int)
    sig = {
        "cnf_version": (I, []),
        "cnf_last_error": (c_char_p, []),
        "cnf_coupling_create": (I, [IP, I, I, I, I, I, I, IP, I, POINTER(P)]),
        "cnf_coupling_destroy": (None, [P]),
        "cnf_coupling_get_info": (I, [P, POINTER(CouplingInfo)]),
        "cnf_coupling_param_entry": (I, [P, I, POINTER(c_char * CNF_NAME_CAP), POINTER(I64), IP,
                                         POINTER(I64 * 4), IP]),
        "cnf_coupling_workspace_bytes": (I64, [P, I64]),
        "cnf_plan_create": (I, [IP, I, I, IP, IP, IP, IP, D, I, I, I, POINTER(P)]),
        "cnf_plan_destroy": (None, [P]),
        "cnf_plan_get_info": (I, [P, POINTER(PlanInfo)]),
        "cnf_plan_block_info": (I, [P, I, POINTER(BlockInfo)]),
        "cnf_plan_layer": (I, [P, I, IP, IP]),
        "cnf_plan_coupling": (P, [P, I]),
        "cnf_plan_coupling_param_offset": (I64, [P, I]),
        "cnf_plan_coupling_level": (I, [P, I]),
        "cnf_plan_workspace_bytes": (I64, [P, I64]),
        "cnf_plan_set_fusion": (I, [P, I]),
        "cnf_coupling_set_fusion": (I, [P, I]),
        "cnf_plan_set_kernel_paths": (I, [P, I]),
        "cnf_coupling_set_kernel_paths": (I, [P, I]),
        "cnf_coupling_resident_eligible": (I, [P]),
        "cnf_flow_forward": (I, [P, P, P, P, P, P, V]),
        "cnf_flow_inverse": (I, [P, P, P, P, P, V]),
        "cnf_flow_log_loss": (I, [P, P, P, P, P, P, P, P, P, V]),
        "cnf_prior_loss": (I, [P, P, P, I, D, P, P, P, V]),
        "cnf_plan_train_workspace_bytes": (I64, [P, I64]),
        "cnf_flow_loss_and_grad": (I, [P, P, P, P, P, P, P, P, P, P, V]),
        "cnf_plan_train_workspace_bytes_recompute": (I64, [P, I64]),
        "cnf_flow_loss_and_grad_recompute": (I, [P, P, P, P, P, P, P, P, P, P, V]),
        "cnf_plan_train_workspace_bytes_invert": (I64, [P, I64]),
        "cnf_flow_loss_and_grad_invert": (I, [P, P, P, P, P, P, P, P, P, P, V]),
        "cnf_flow_loss_and_grad_hooked": (I, [P, P, P, P, P, P, P, P, P, P, V, I, LAYER_GRADS_READY_FN, V]),
        "cnf_adam_step": (I, [P, P, P, P, I64, D, D, D, D, D, V]),
        "cnf_coupling_forward": (I, [P, P, P, P, P, P, V]),
        "cnf_coupling_backward": (I, [P, P, P, P, P, V]),
        "cnf_coupling_nets": (I, [P, P, P, P, P, P, V]),
        "cnf_coupling_law": (I, [P, P, P, I, I, P, P, V]),
        "cnf_residual_block": (I, [P, I, P, P, P, P, V]),
        "cnf_measure_stage": (I, [P, P, P, I64, I, V]),
        "cnf_mask": (I, [P, I, I, P, V]),
        "cnf_decompress_mask": (I, [P, I, P, V]),
        "cnf_space_to_depth": (I, [P, P, V]),
        "cnf_depth_to_space": (I, [P, P, V]),
        "cnf_down": (I, [P, I, P, V]),
        "cnf_up": (I, [P, I, P, V]),
        "cnf_sr_preprocess": (I, [P, I, I, I, P, V]),
        "cnf_logit_scale": (I, [P, D, I, P, V]),
        "cnf_instance_noise": (I, [P, D, c_uint64, c_uint64, P, V]),
        "cnf_toy_param_count": (I64, [I, I, I]),
        "cnf_toy_layer_offset": (I64, [I, I, I]),
        "cnf_toy_call": (I, [P, P, IP, I, I, I, I, P, P, V]),
        "cnf_toy_log_loss": (I, [P, P, IP, I, I, I, I, D, P, P, P, P, P, V]),
        "cnf_toy_loss_and_grad": (I, [P, P, IP, I, I, I, I, D, P, P, P, P, P, P, V]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)          # AttributeError here == header/library mismatch
        fn.restype = res
        fn.argtypes = args
    return lib, sig


lib, SIGNATURES = _load()

_capsule_ptr = ctypes.pythonapi.PyCapsule_GetPointer
_capsule_ptr.restype = c_void_p
_capsule_ptr.argtypes = [py_object, c_char_p]


class Borrowed:
    """Holds DLPack capsules alive for one C call and yields their DLManagedTensor* values."""

    def __init__(self):
        self._caps = []

    def __call__(self, t):
        if t is None:
            return None
        cap = to_dlpack(t)
        self._caps.append(cap)
        return _capsule_ptr(cap, b"dltensor")


def check(rc):
    if rc != CNF_OK:
        msg = lib.cnf_last_error().decode("utf-8", "replace")
        raise _EXC.get(rc, RuntimeError)(msg)


def stream_ptr():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def int_array(values):
    values = [int(v) for v in values]
    return (c_int * len(values))(*values)


def require_cuda(t, name="tensor"):
    """Python-side mirror of the checks libcnf makes, so that a CPU tensor fails loudly and early."""
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise RuntimeError(f"{name}: this path only runs on CUDA (sm_100a) tensors; got a {t.device} tensor. "
                           "There is no CPU fallback.")
    if t.dtype != torch.float32:
        raise TypeError(f"{name}: expected float32, got {t.dtype}")
    return t.contiguous()


def coupling_info(handle):
    info = CouplingInfo()
    check(lib.cnf_coupling_get_info(handle, byref(info)))
    return info


def coupling_entries(handle, n_entries):
    """[(name, offset, shape tuple, role)] of one net of a coupling layer."""
    out = []
    name = (c_char * CNF_NAME_CAP)()
    off = c_int64()
    ndim = c_int()
    shape = (c_int64 * 4)()
    role = c_int()
    for i in range(n_entries):
        check(lib.cnf_coupling_param_entry(handle, i, byref(name), byref(off), byref(ndim), byref(shape),
                                           byref(role)))
        out.append((name.value.decode(), int(off.value), tuple(int(shape[j]) for j in range(ndim.value)),
                    int(role.value)))
    return out
