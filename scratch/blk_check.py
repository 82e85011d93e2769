"""development check: numpy block-time-step main loop vs the reference's own MainLoop after k steps"""
import sys, os, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import refio, oracle_py as orc
from conftest import load_golden, stage
BASE = {"run_id": "LOOP", "sim": "gradhsph", "ic": "python", "dimensionless": 1, "Nlevels": 1,
        "Nstepsmax": 1000, "tend": 1.0e10, "dt_snap": 1.0e10, "noutputstep": 1000000, "ndiagstep": 1000000,
        "out_file_form": "su", "sink_particles": 0, "sph_integration": "lfkdk", "neib_search": "kdtree"}
name = sys.argv[1]; Nlevels = int(sys.argv[2]); ks = [int(x) for x in sys.argv[3].split(",")]
params, d = load_golden(name)
p = dict(BASE); p.update(params); p.update({k: BASE[k] for k in ("run_id","ic","Nstepsmax","tend","neib_search")})
p["Nlevels"] = Nlevels
s = stage(d, "dens_in"); order = np.argsort(s["iorig"]); ic = {k: s[k][order] for k in ("r","v","m","h","u")}
p.setdefault("Nhydromax", 8*len(ic["m"]))
F = ["r","v","a","atree","m","h","u","rho","invomega","zeta","hfactor","hrangesqd","sound","pressure","alpha","dalphadt","dudt","div_v","gpot",
     "ptype","level","levelneib","nstep","nlast","dt","dt_next","tlast","r0","v0","a0","u0","dudt0"]
for k in ks:
    out, info = refio.run_mainloop(p, ic, k, gpu=False, kind="std", dumps="setup")
    o = np.argsort(out["setup.iorig"])
    setup = {f: out["setup."+f][o] for f in F}
    st, sim = orc.mainloop_block(p, setup, info["setup"], k)
    o2 = np.argsort(out["loop.iorig"])
    ref = {f: out["loop."+f][o2] for f in F}
    print("k=%d ref: n %d nresync %d level_max %d ts %.6e t %.6e | ours: n %d nresync %d level_max %d ts %.6e t %.6e" % (
        k, info["n"], info["nresync"], info["level_max"], info["timestep"], info["t"], sim["n"], sim["nresync"], sim["level_max"], sim["timestep"], sim["t"]))
    for f in ("level","levelneib","nstep","nlast"):
        bad = np.nonzero(ref[f] != st[f])[0]
        if len(bad): print("   MISMATCH", f, len(bad), bad[:8], ref[f][bad[:8]], st[f][bad[:8]])
    for f in ("r","v","a","h","rho","u","dudt","dt","tlast","alpha"):
        a = np.asarray(st[f], dtype=float).reshape(len(ref["m"]), -1); b = ref[f].reshape(len(ref["m"]), -1)
        sc = np.abs(b).max() or 1.0
        print("   %-6s max err %.2e" % (f, np.abs(a-b).max()/sc), end="")
    print()
    if os.environ.get("BLKDBG"):
        f = os.environ["BLKDBG"]
        a = np.asarray(st[f], dtype=float).reshape(len(ref["m"]), -1); b = ref[f].reshape(len(ref["m"]), -1)
        allbad = np.nonzero(np.abs(a-b).max(1) > 1e-9*np.abs(b).max())[0]
        print("   %d particles differ; their levels" % len(allbad), np.bincount(ref["level"][allbad]), "nlast", np.unique(ref["nlast"][allbad]))
        a = a[:, 0]; b = b[:, 0]
        for i in allbad[:6]:
            print("   i=%d %s ours %.6e ref %.6e | level %d/%d nstep %d/%d nlast %d/%d setup level %d nstep %d dt %.6e dt_next %.6e" % (
                i, f, a[i], b[i], st["level"][i], ref["level"][i], st["nstep"][i], ref["nstep"][i], st["nlast"][i], ref["nlast"][i],
                setup["level"][i], setup["nstep"][i], setup["dt"][i], setup["dt_next"][i]))
