import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_flow(name):
    d = np.load(os.path.join(GOLDEN, f"flow_{name}.npz"))
    cfg = json.loads(str(d["config"]))
    n = 1 + max(int(k.split(".")[1]) for k in d.files if k.startswith("w."))
    W = [{"A": {}, "b": {}} for _ in range(n)]
    for k in d.files:
        if k.startswith("w."):
            _, i, net, pname = k.split(".", 3)
            W[int(i)][net][pname] = d[k]
    return cfg, W, d


def load_toy():
    d = np.load(os.path.join(GOLDEN, "toy.npz"))
    n = int(d["n"])
    W = []
    for j in range(n):
        e = {}
        for net in ("A", "b"):
            layers, i = [], 0
            while f"w.{j}.{net}.{i}.W" in d.files:
                layers.append((d[f"w.{j}.{net}.{i}.W"], d[f"w.{j}.{net}.{i}.b"]))
                i += 1
            e[net] = layers
        W.append(e)
    return W, d
