import sys, numpy as np, torch
import gandalf_b200 as g
from gandalf_b200 import workloads
NP = int(sys.argv[1]) if len(sys.argv) > 1 else 67108864
params, ic = workloads.WORKLOADS["cloud"](NP, device="cuda:0")
N = int(ic["m"].shape[0])
ps = g.ParticleSet(N, 3, "soa", fields={"r", "v", "m", "h", "u"})
for k in ("r", "v"):
    for c in range(3):
        ps[k][:, c] = ic[k][c].cpu().numpy()
for k in ("m", "h", "u"):
    ps[k][...] = ic[k].cpu().numpy()
del ic
tree = g.GpuSphTree(params)
tree.Upload(ps, only=("r", "v", "m", "h", "u"))
for world in (2, 4, 8):
    for rep in range(2):
        tree.Upload(ps, only=("r", "m", "h"))
        tree.Synchronize()
        tree.BuildTreeShardBegin(world - 1, world)
        st = tree.Stats()
    print("world", world, "begin ms", st.ms_build_tree, "select+sort", st.ms_tree_sort, "levels", st.ms_tree_levels)
