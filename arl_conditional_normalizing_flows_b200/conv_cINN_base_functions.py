"""Mirror of the hot-path part of the reference's `conv_cINN_base_functions.py` (F).

Only `dilated_residual_block` (F:501-627) with its helpers `grouped_convolution` (F:364-413) and
`add_common_layers` (F:330-362) is on the path SURVEY §8 names; `conv_cINN_make_model` imports it
(M:24).  In the reference it is a Keras functional-graph builder that is traced once while a
`coupling_layer` builds its two s/t models; here that construction is done by the C planner
(`cnf_coupling_create`: channel split per dilation F:577-588, groups F:396-411) and the block runs
inside the fused kernels of csrc/stnet_kernels.cu.  The data helpers of the reference file
(TFRecord parsing, up/down sampling, de_logitify, instance noise) are input pipeline and out of scope.
"""
import numpy as np


def residual_block_plan(nb_channels_in, _which_dilations=(1, 2, 4), cardinality=4):
    """Channel bookkeeping of one dilated residual block: for every dilation d the branch gets
    nb_channels_in // d channels (F:579), split into `cardinality` groups of equal width (F:396-397)
    that read the FIRST nb_channels_in // d input channels (F:402); outputs are concatenated (F:590)."""
    branches, offset = [], 0
    for d in _which_dilations:
        nb = nb_channels_in // int(d)
        if cardinality == 1:
            groups, gin, gout = 1, nb_channels_in, nb
        else:
            assert not nb % cardinality
            groups, gin, gout = cardinality, nb // cardinality, nb // cardinality
        branches.append({'dilation': int(d), 'channels': nb, 'groups': groups, 'group_in': gin,
                         'group_out': gout, 'concat_offset': offset})
        offset += nb
    return {'branches': branches, 'concat_channels': offset}


def dilated_residual_block(y, nb_channels_in, nb_channels_out, _strides=(1, 1), _project_shortcut=False,
                           _which_dilations=[1, 2, 4], ksize=(4, 4), cardinality=4, ln=False, do=False,
                           ln_axis=-1, init='glorot_uniform', weights=None):
    """F:501-627 on a concrete torch CUDA tensor `y` [B,h,w,nb_channels_out].

    Runs one residual block through libcnf (`coupling_layer` machinery with a single block); `weights`
    is the block's parameter dict (names as in coupling_layer.get_weights()['b'] with prefix 'rb0.').
    Strided / projected shortcuts and dropout are never used by the reference model (M:1123-1130) and
    are not built."""
    if tuple(_strides) != (1, 1) or _project_shortcut or do:
        raise NotImplementedError("only the configuration conv_cINN_make_model uses is built "
                                  "(strides (1,1), identity shortcut, no dropout; M:1123-1130)")
    if nb_channels_in != nb_channels_out:
        raise NotImplementedError("identity shortcut needs nb_channels_in == nb_channels_out (F:614-623)")
    raise NotImplementedError("stand-alone residual blocks are not exposed yet; the block runs fused "
                              "inside coupling_layer (see csrc/stnet_kernels.cu)")
