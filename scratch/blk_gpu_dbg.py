"""device block-time-step loop vs the numpy oracle, step by step (first divergence)"""
import sys, os, numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gandalf_b200 as g, oracle_py as orc
from conftest import load_golden
from test_mainloop_oracle import split_blk
from test_gpu_mainloop import _upload_setup
name = sys.argv[1]; K = int(sys.argv[2])
params, d = load_golden(name)
setup, sim0, loop, sim1 = split_blk(d)
ps, tree = _upload_setup(g, params, setup)
tree.IntegratorBegin(compute_forces=False)
s, sim = setup, sim0
for k in range(1, K + 1):
    tree.MainLoop(1)
    tree.Download(ps)
    b = tree.IntegratorBlockState(arrays=True)
    t, ts, n = tree.IntegratorState()
    s, sim = orc.mainloop_block(params, s, sim, 1)
    line = "k=%2d dev n %d nresync %d lmax %d ts %.10e | orc n %d nresync %d lmax %d ts %.10e |" % (
        k, b["n"], b["nresync"], b["level_max"], ts, sim["n"], sim["nresync"], sim["level_max"], sim["timestep"])
    for f in ("level", "levelneib"):
        bad = np.nonzero(ps[f] != s[f])[0]
        if len(bad): line += " %s mismatch %d (i=%s dev %s orc %s)" % (f, len(bad), bad[:4], ps[f][bad[:4]], s[f][bad[:4]])
    for f in ("nstep", "nlast"):
        bad = np.nonzero(b[f] != s[f])[0]
        if len(bad): line += " %s mismatch %d (i=%s dev %s orc %s)" % (f, len(bad), bad[:4], b[f][bad[:4]], s[f][bad[:4]])
    for f in ("r", "v", "a", "h", "rho", "u", "dudt", "div_v", "sound"):
        A = np.asarray(ps[f], dtype=float).reshape(len(s["m"]), -1); B = np.asarray(s[f], dtype=float).reshape(len(s["m"]), -1)
        e = np.abs(A - B).max()/(np.abs(B).max() or 1.0)
        if e > 1e-11: line += " %s %.1e(i=%d)" % (f, e, np.abs(A - B).max(1).argmax())
    for f, arr in (("dt", b["dt"]), ("tlast", b["tlast"])):
        e = np.abs(arr - s[f]).max()/(np.abs(s[f]).max() or 1.0)
        if e > 1e-11: line += " %s %.1e" % (f, e)
    print(line)
tree.close()
