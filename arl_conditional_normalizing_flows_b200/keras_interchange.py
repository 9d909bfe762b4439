"""Weight interchange with the reference's Keras model.

The reference saves `model.save_weights(...h5)` (C:639-641, P:146-147) and reloads with `model.load_weights` after one
call that builds the model (C:572-579).  h5py / TensorFlow are not part of this image, so the interchange file is a
NumPy `.npz` keyed by the Keras VARIABLE NAMES, which a maintainer writes on a TensorFlow box with

    np.savez(path, **{w.name: w.numpy() for w in model.weights})            # reference side, TF >= 2.7

and reads back with `{w.name: w for w in model.weights}[name].assign(npz[name])`.

Name map.  Keras numbers layers of one class in creation order inside a fresh process: `conv2d`, `conv2d_1`, ...;
`layer_normalization`, `layer_normalization_1`, ...; `tanh_scaling_layer`, ... .  cFlow.__init__ creates the coupling
layers in `layers_list` order (M:1636-1689); each builds its two functional models in `coupling_function` (M:1076-1213)
in this order: **b-net first, then A-net**; per net: stem conv (M:1106-1111), per residual block [LN1, 1x1 conv, LN2,
for every dilation (for every group: conv), LN3, 1x1 conv] (F:552-612, F:387-411; cardinality 1: one conv per dilation,
F:389-395), final LN (M:1133-1141), head conv (M:1143-1148); the A-net ends with `tanh_scaling_layer` whose scalar
weight is created by `add_weight` without a name, i.e. `tanh_scaling_layer_k/Variable:0` (M:109-112).
Shapes are identical on both sides (HWIO kernels, flat h*w*C LayerNorm vectors), so no transposition is involved.
The map assumes the model was the first thing built in the TensorFlow process; `offsets=` shifts the three counters if it
was not.
"""
import re

import numpy as np

_SUFFIX = {'kernel': 'kernel:0', 'bias': 'bias:0', 'gamma': 'gamma:0', 'beta': 'beta:0'}


def _keras_layer_name(base, k):
    return base if k == 0 else f"{base}_{k}"


def _net_entries(names):
    """our parameter names of one net, grouped into Keras layers in creation order: [(kind, [(our_name, weight)])]"""
    def has(n):
        return n in names
    layers = [('conv2d', [('stem.kernel', 'kernel'), ('stem.bias', 'bias')])]
    r = 0
    while has(f'rb{r}.pw1.kernel'):
        if has(f'rb{r}.ln1.gamma'):
            layers.append(('layer_normalization', [(f'rb{r}.ln1.gamma', 'gamma'), (f'rb{r}.ln1.beta', 'beta')]))
        layers.append(('conv2d', [(f'rb{r}.pw1.kernel', 'kernel'), (f'rb{r}.pw1.bias', 'bias')]))
        if has(f'rb{r}.ln2.gamma'):
            layers.append(('layer_normalization', [(f'rb{r}.ln2.gamma', 'gamma'), (f'rb{r}.ln2.beta', 'beta')]))
        # grouped convs: dilations in list order, groups in index order (F:577-588, F:399-410)
        gc = {}
        for n in names:
            m = re.fullmatch(rf'rb{r}\.gc\.d(\d+)\.g(\d+)\.kernel', n)
            if m:
                gc[(int(m.group(1)), int(m.group(2)))] = n[:-len('.kernel')]
        for key in sorted(gc):
            layers.append(('conv2d', [(gc[key] + '.kernel', 'kernel'), (gc[key] + '.bias', 'bias')]))
        if has(f'rb{r}.ln3.gamma'):
            layers.append(('layer_normalization', [(f'rb{r}.ln3.gamma', 'gamma'), (f'rb{r}.ln3.beta', 'beta')]))
        layers.append(('conv2d', [(f'rb{r}.pw2.kernel', 'kernel'), (f'rb{r}.pw2.bias', 'bias')]))
        r += 1
    if has('lnf.gamma'):
        layers.append(('layer_normalization', [('lnf.gamma', 'gamma'), ('lnf.beta', 'beta')]))
    layers.append(('conv2d', [('head.kernel', 'kernel'), ('head.bias', 'bias')]))
    if has('tanh_scale'):
        layers.append(('tanh_scaling_layer', [('tanh_scale', 'Variable:0')]))
    return layers


def keras_weight_names(layer_weight_names, offsets=None):
    """layer_weight_names: for every coupling layer (layers_list order) {'A': iterable of names, 'b': iterable of names}
    (e.g. `[{n: w.keys() for n, w in lw.items()} for lw in model.get_weights()]`).
    Returns [(keras_variable_name, coupling_index, net, our_name)] in Keras creation order."""
    counters = dict(conv2d=0, layer_normalization=0, tanh_scaling_layer=0)
    counters.update(offsets or {})
    out = []
    for ci, nets in enumerate(layer_weight_names):
        for net in ('b', 'A'):                                     # coupling_function builds b first (M:1106 vs M:1150)
            for kind, ws in _net_entries(set(nets[net])):
                lname = _keras_layer_name(kind, counters[kind])
                counters[kind] += 1
                for our, w in ws:
                    out.append((f"{lname}/{_SUFFIX.get(w, w)}", ci, net, our))
    return out


def export_keras_npz(model, path, offsets=None):
    """Write the model's weights keyed by the reference's Keras variable names."""
    W = model.get_weights()
    table = keras_weight_names([{n: list(w) for n, w in lw.items()} for lw in W], offsets)
    np.savez(path, **{k: np.asarray(W[ci][net][our], dtype=np.float32) for k, ci, net, our in table})
    return [k for k, *_ in table]


def import_keras_npz(model, path, offsets=None):
    """Load a `.npz` written on the reference side (see the module docstring) into `model` (shape-checked)."""
    W = model.get_weights()
    table = keras_weight_names([{n: list(w) for n, w in lw.items()} for lw in W], offsets)
    with np.load(path) as z:
        missing = [k for k, *_ in table if k not in z.files]
        if missing:
            raise KeyError(f"import_keras_npz: {len(missing)} variables missing, first: {missing[:3]}")
        for k, ci, net, our in table:
            a = np.asarray(z[k], dtype=np.float32)
            want = np.shape(W[ci][net][our])
            if a.shape != tuple(want):
                raise ValueError(f"import_keras_npz: {k} -> layer {ci} net {net} {our}: shape {a.shape}, expected {tuple(want)}")
            W[ci][net][our] = a
    model.set_weights(W)
    return len(table)
