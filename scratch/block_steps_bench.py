"""device-resident main loop with hierarchical block time steps on the bench's cloud: ms per MainLoop call and
active particle-updates per second (only the particles at the end of their own step pass through the kernels)"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gandalf_b200 as g
from gandalf_b200 import workloads

NP = int(sys.argv[1]) if len(sys.argv) > 1 else 16777216
nlev = int(sys.argv[2]) if len(sys.argv) > 2 else 5
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 32
params, ic = workloads.WORKLOADS["cloud"](NP, device="cuda:0")
params = dict(params, Nlevels=nlev)
N = int(ic["m"].shape[0])
ps = g.ParticleSet(N, 3)
for k in ("r", "v"):
    ps[k][...] = ic[k].T.cpu().numpy()
for k in ("m", "h", "u"):
    ps[k][...] = ic[k].cpu().numpy()
tree = g.GpuSphTree(params)
tree.IntegratorBegin(ps, compute_forces=True)
b0 = tree.IntegratorBlockState()
tree.MainLoop(2)
tree.Synchronize()
act, ms = [], []
for s in range(steps):
    tree.TimerStart()
    tree.MainLoop(1)
    ms.append(tree.TimerStop())
    act.append(int(tree.Stats().n_active))
tree.Download(ps, only=("level",))
b1 = tree.IntegratorBlockState()
lv, cnt = np.unique(ps["level"], return_counts=True)
out = {"workload": "turbulent self-gravitating cloud, %d particles, Nlevels=%d" % (N, nlev), "steps": steps,
       "ms_per_mainloop": float(np.mean(ms)), "ms_min": float(np.min(ms)), "ms_max": float(np.max(ms)),
       "active_fraction_mean": float(np.mean(act))/N, "active_updates_per_s": float(np.sum(act))/(np.sum(ms)*1e-3),
       "levels": {int(a): int(b) for a, b in zip(lv, cnt)}, "hierarchy_begin": b0, "hierarchy_end": b1,
       "note": "n_active = particles active when the step's tree was built (woken particles not counted)"}
print(json.dumps(out))
tree.close()
