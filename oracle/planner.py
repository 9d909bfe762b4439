"""Oracle restatement of the architecture planner in `cFlow.__init__`.

Follows /root/reference/conv_cINN_make_model.py:
  * validation asserts            M:1459-1484
  * scale / num_prev_factors      M:1493-1518
  * io_shape_list                 M:1521-1536
  * dilation search               M:1553-1617
  * layer order                   M:1636-1689
  * per-coupling-layer shapes     M:420-433 (kernel halving, complement),
                                  M:474-498 (compressed shape),
                                  M:1093-1104 (uv2 depth for odd D)
Pure Python ints; TEST INFRASTRUCTURE (see oracle/__init__.py).
"""
import math


def plan_flow(io_shape, x_d, squeeze_factor_block_list, ResNeXt_block_list,
              num_kernels_list, cardinality_list, ksize=3, DILATIONS=True):
    sq = list(squeeze_factor_block_list)
    # M:1459-1484
    assert len(sq) == len(ResNeXt_block_list) == len(num_kernels_list) == len(cardinality_list)
    assert not io_shape[0] % 2 and not io_shape[1] % 2
    for nk in num_kernels_list:
        assert not nk % 2
    for c in cardinality_list:
        assert not c % 2
    for s in sq:
        assert s in [0, 1]
    nblocks = len(sq)

    # M:1493-1518
    scale_list, npf_list = [], []
    scale_flag, npf = 0, 0
    for i in range(nblocks):
        s = 0 if i == 0 else sq[i - 1]
        if not scale_flag:
            scale_list.append(1)
            scale_flag = 1
        else:
            scale_list.append(2 ** s * scale_list[-1])
        npf += s
        npf_list.append(npf)

    # M:1521-1536
    io_shape_list = []
    for i in range(nblocks):
        scale = scale_list[i]
        assert not io_shape[0] % (scale * 2) and not io_shape[1] % (scale * 2)
        io_shape_list.append([int(io_shape[0] / scale), int(io_shape[1] / scale), io_shape[2] * scale])

    # M:1553-1617
    assert DILATIONS, "reference never defines dilations_list when DILATIONS is False (M:1553)"
    dilations_list = []
    for shp in io_shape_list:
        d = {'checkerboard': [], 'channelwise': []}
        s_ch = min(shp[0], shp[1])
        s_cb = s_ch / 2
        d_ip1 = 1
        sanity = 0
        dk = ksize
        if dk > (s_ch + 1) / 2:
            d['channelwise'].append(d_ip1)
            d['checkerboard'].append(d_ip1)
        else:
            while dk < (s_ch + 1) / 2:
                assert sanity < 10
                d['channelwise'].append(d_ip1)
                if d_ip1 < (s_cb + 1) / 2:
                    d['checkerboard'].append(d_ip1)
                dk = (ksize - 1) * (dk - 1) + 1
                d_ip1 = ((dk - ksize) / (ksize - 1)) + 1
                sanity += 1
        dilations_list.append(d)
    for i in range(nblocks):
        nkc = num_kernels_list[i] / cardinality_list[i]
        for dil in dilations_list[i]['channelwise']:
            assert not nkc % dil

    # M:1636-1689: layer order; each entry is a dict describing one layer
    layers = []
    for i in range(nblocks):
        H, W, D = io_shape_list[i]
        for mask in [0, 1, 2, 3]:
            dil = dilations_list[i]['checkerboard' if mask in (0, 1) else 'channelwise']
            layers.append(plan_coupling([H, W, D], mask, ResNeXt_block_list[i], cardinality_list[i],
                                        num_kernels_list[i], ksize, dil))
        if sq[i] == 1:
            layers.append({'type': 'squeeze'})
            layers.append({'type': 'factor', 'num_prev_factors': npf_list[i]})
    return {
        'io_shape': list(io_shape), 'x_d': x_d, 'scale_list': scale_list,
        'num_prev_factors_list': npf_list, 'io_shape_list': io_shape_list,
        'dilations_list': dilations_list, 'layers': layers,
    }


def plan_coupling(in_shape, which_mask, num_res_blocks, cardinality, num_kernels, ksize, which_dilations):
    H, W, D = in_shape
    assert H % 2 == 0 and W % 2 == 0                       # M:415-417
    nk = int(num_kernels / 2) if which_mask in (0, 1) else num_kernels   # M:420-423
    comp = {0: 1, 1: 0, 2: 3, 3: 2}[which_mask]             # M:426-433
    if which_mask in (0, 1):                                # M:474-498
        h, w, c1 = int(H / 2), int(W / 2), 2 * D
    else:
        h, w = H, W
        c1 = int(math.ceil(D / 2)) if which_mask == 2 else int(math.floor(D / 2))
    if D % 2 and which_mask == 2:                           # M:1093-1104
        c2 = c1 - 1
    elif D % 2 and which_mask == 3:
        c2 = c1 + 1
    else:
        c2 = c1
    dil = [int(d) for d in which_dilations]
    branches = []
    for d in dil:                                           # F:577-588, F:396-397
        nb = nk // d
        if cardinality != 1:
            assert not nb % cardinality
        g = nb // cardinality
        branches.append({'dilation': d, 'channels': nb, 'group_width': g})
    cat = sum(b['channels'] for b in branches)
    return {'type': 'coupling', 'in_shape': [H, W, D], 'mask': which_mask, 'mask_complement': comp,
            'R': int(num_res_blocks), 'cardinality': int(cardinality), 'nk': nk, 'ksize': ksize,
            'h': h, 'w': w, 'c1': c1, 'c2': c2, 'dilations': dil, 'branches': branches, 'cat': cat}
