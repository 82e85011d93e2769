import glob
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


GOLDEN_DIR = os.path.join(HERE, "golden")
_ALL_NPZ = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
GOLDEN_CASES = [c for c in _ALL_NPZ if not c.startswith(("loop_", "loopblk_", "snap_", "stargas_", "brute_"))]   # operator-level fixtures (make_golden.py)
SNAP_CASES = [c[5:] for c in _ALL_NPZ if c.startswith("snap_")]        # "su" snapshot files (make_snapshot_golden.py)
LOOP_CASES = [c for c in _ALL_NPZ if c.startswith("loop_")]            # main-loop fixtures (make_loop_golden.py)
LOOPBLK_CASES = [c for c in _ALL_NPZ if c.startswith("loopblk_")]      # main loop with block time steps (make_loopblk_golden.py)


def load_golden(name):
    """-> (params dict with the reference's .dat names, dict of arrays)"""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    params, data = {}, {}
    for k in z.files:
        if k.startswith("param."):
            params[k[6:]] = str(z[k])
        else:
            data[k] = z[k]
    for k in ("ndim", "tabulated_kernel", "self_gravity", "hydro_forces", "Nleafmax"):
        if k in params:
            params[k] = int(params[k])
    return params, data


def stage(data, name):
    """fields of one dump stage as a dict, e.g. stage(d, 'dens_in')['r']"""
    n = len(name) + 1
    aux = ("neib_", "tree_", "direct_", "smooth_", "gravcell", "leaf_of")
    return {k[n:]: v for k, v in data.items()
            if k.startswith(name + ".") and not k[n:].startswith(aux)}


def rel_err(a, b, floor=0.0):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = np.maximum(np.abs(b), floor)
    scale[scale == 0] = 1.0
    return np.abs(a - b) / scale


def scaled_err(a, b):
    """|a-b| / max|b| (1 if b is identically zero): for fields with sign changes / exact zeros"""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = np.abs(b).max() if b.size else 1.0
    if scale == 0:
        scale = 1.0
    return np.abs(a - b) / scale


def vec_rel_err(a, b):
    """relative error of vectors measured against the vector magnitude (not per component)"""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    nb = np.linalg.norm(b.reshape(len(b), -1), axis=1)
    nb[nb == 0] = 1.0
    return np.linalg.norm((a - b).reshape(len(b), -1), axis=1) / nb


@pytest.fixture(params=GOLDEN_CASES)
def golden(request):
    return (request.param,) + load_golden(request.param)
