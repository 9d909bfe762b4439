"""Oracle: literal NumPy transcription of the reference's index permutations.

Follows /root/reference/conv_cINN_make_model.py op by op (slices, scatter_nd,
transposes as written), NOT a simplified closed form, so it can pin the closed-form
addressing used by the CUDA kernels:
  * coupling_layer.mask            M:500-761   (uncompressed M:632-717, compressed M:723-759)
  * coupling_layer.decompress_mask M:763-1073
  * squeeze_layer                  M:155-217   (tf.nn.space_to_depth / depth_to_space, block 2, NHWC)
  * factor_out_zy_layer            M:256-329
TEST INFRASTRUCTURE (see oracle/__init__.py).
"""
import numpy as np


def _scatter_nd(indices, updates, shape):
    """tf.scatter_nd for indices of shape (n, 1): rows of `updates` are ADDED into a zero
    tensor of `shape` at first-axis positions `indices[:, 0]`."""
    out = np.zeros(tuple(int(s) for s in shape), dtype=updates.dtype)
    np.add.at(out, indices[:, 0], updates)
    return out


def space_to_depth2(x):
    """tf.nn.space_to_depth(x, 2) for NHWC: out[b,i,j,(dy*2+dx)*C+c] = in[b,2i+dy,2j+dx,c]."""
    B, H, W, C = x.shape
    x = x.reshape(B, H // 2, 2, W // 2, 2, C)
    x = x.transpose(0, 1, 3, 2, 4, 5)
    return np.ascontiguousarray(x.reshape(B, H // 2, W // 2, 4 * C))


def depth_to_space2(x):
    """tf.nn.depth_to_space(x, 2) for NHWC (exact inverse of space_to_depth2)."""
    B, h, w, C4 = x.shape
    C = C4 // 4
    x = x.reshape(B, h, w, 2, 2, C)
    x = x.transpose(0, 1, 3, 2, 4, 5)
    return np.ascontiguousarray(x.reshape(B, 2 * h, 2 * w, C))


def squeeze_forward(u, zy):                                  # M:155-185
    assert u.shape[1] % 2 == 0 and u.shape[2] % 2 == 0
    v = space_to_depth2(u)
    if zy is not None:
        zy = space_to_depth2(zy)
    return v, zy


def squeeze_backward(v, zy):                                 # M:191-217
    assert v.shape[3] % 4 == 0
    u = depth_to_space2(v)
    if zy is not None:
        zy = depth_to_space2(zy)
    return u, zy


def factor_forward(u, zy):                                   # M:256-288
    split = u.shape[3] // 2
    factored = u[..., :split]
    v = u[..., split:]
    zy = factored if zy is None else np.concatenate([zy, factored], axis=3)
    return v, zy


def factor_backward(v, zy, num_prev_factors):                # M:294-329
    if v is None:
        split = zy.shape[3] // (2 ** num_prev_factors)
    else:
        split = v.shape[3]
    re_v = zy[..., -split:]
    zy = zy[..., :-split]
    assert re_v.shape[3] == split
    u = re_v if v is None else np.concatenate([re_v, v], axis=3)
    return u, zy


def mask(uv, which_mask_index, compress):                    # M:500-761
    B, H, W, D = uv.shape
    if not compress:                                         # M:632-717
        if which_mask_index == 0:
            ones, zeros = np.ones(D, uv.dtype), np.zeros(D, uv.dtype)
            m = np.stack([np.stack([ones, zeros]), np.stack([zeros, ones])])
        elif which_mask_index == 1:
            ones, zeros = np.ones(D, uv.dtype), np.zeros(D, uv.dtype)
            m = np.stack([np.stack([zeros, ones]), np.stack([ones, zeros])])
        elif which_mask_index == 2:
            idx = np.arange(0, D, 2)[:, None]
            mm = _scatter_nd(idx, np.ones(int(np.ceil(D / 2)), uv.dtype), [D])
            m = np.stack([np.stack([mm, mm]), np.stack([mm, mm])])
        elif which_mask_index == 3:
            idx = np.arange(1, D, 2)[:, None]
            mm = _scatter_nd(idx, np.ones(int(np.floor(D / 2)), uv.dtype), [D])
            m = np.stack([np.stack([mm, mm]), np.stack([mm, mm])])
        m = np.tile(m, [int(H / 2), int(W / 2), 1])           # (H, W, D)
        return np.einsum('jkl,ijkl->ijkl', m, uv)
    if which_mask_index in (0, 1):                           # M:723-748
        if which_mask_index == 0:
            c0, c1 = uv[:, 0::2, 0::2, :], uv[:, 1::2, 1::2, :]
        else:
            c0, c1 = uv[:, 0::2, 1::2, :], uv[:, 1::2, 0::2, :]
        return np.concatenate([c0, c1], axis=-1)
    if which_mask_index == 2:                                # M:753-759
        return np.ascontiguousarray(uv[..., 0::2])
    return np.ascontiguousarray(uv[..., 1::2])


def decompress_mask(uvc, which_mask_index, out_shape):       # M:763-1073
    B, hc, wc, dc = uvc.shape
    _, H, W, D = out_shape
    if which_mask_index in (0, 1):
        assert dc % 2 == 0
        c0, c1 = uvc[..., :D], uvc[..., D:]
        ih0 = np.arange(0, 2 * hc, 2)[:, None]
        ih1 = np.arange(1, 2 * hc + 1, 2)[:, None]
        iw0 = np.arange(0, 2 * wc, 2)[:, None]
        iw1 = np.arange(1, 2 * wc + 1, 2)[:, None]

        def place(c, ih, iw):
            upd = c.transpose(1, 2, 3, 0)                                     # (h/2,w/2,d,B)
            sc = _scatter_nd(ih, upd, np.array([2, 1, 1, 1]) * upd.shape)     # (h,w/2,d,B)
            upd = sc.transpose(1, 0, 2, 3)                                    # (w/2,h,d,B)
            sc = _scatter_nd(iw, upd, np.array([2, 1, 1, 1]) * upd.shape)     # (w,h,d,B)
            return sc.transpose(3, 1, 0, 2)                                   # (B,h,w,d)

        if which_mask_index == 0:
            return place(c0, ih0, iw0) + place(c1, ih1, iw1)
        return place(c0, ih0, iw1) + place(c1, ih1, iw0)
    if which_mask_index == 2:
        idx = np.arange(0, D, 2)[:, None]
    else:
        idx = np.arange(1, D, 2)[:, None]
    upd = uvc.transpose(3, 1, 2, 0)
    shape = np.array([2, 1, 1, 1]) * upd.shape
    if D % 2:                                                # M:1050-1060
        if which_mask_index == 2:
            shape = shape - [1, 0, 0, 0]
        if which_mask_index == 3:
            shape = shape + [1, 0, 0, 0]
    sc = _scatter_nd(idx, upd, shape)
    return sc.transpose(3, 1, 2, 0)
