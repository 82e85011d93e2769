"""One rank's share of a multi-GPU step, emulated on ONE GPU (no NCCL): every phase bracketed by a
synchronise + wall clock.  The other slices' fields are valid from a full single-GPU pass before."""
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gandalf_b200 as g
from gandalf_b200 import workloads, domain
NP = int(sys.argv[1]) if len(sys.argv) > 1 else 67108864
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
rank = int(sys.argv[3]) if len(sys.argv) > 3 else world - 1
params, ic = workloads.WORKLOADS["cloud"](NP, device="cuda:0")
N = int(ic["m"].shape[0])
ps = g.ParticleSet(N, 3, "soa", fields={"r", "v", "m", "h", "u"})
for k in ("r", "v"):
    for c in range(3):
        ps[k][:, c] = ic[k][c].cpu().numpy()
for k in ("m", "h", "u"):
    ps[k][...] = ic[k].cpu().numpy()
del ic
torch.cuda.empty_cache()
tree = g.GpuSphTree(params)
IN = ("r", "v", "m", "h", "u")
tree.Upload(ps, only=IN)
tree.BuildTree(True, 0, 1, 1, 0.0)
tree.UpdateAllSphProperties()
tree.Download(ps, only=("h",))
tree.Upload(ps, only=IN)
def sync():
    tree.Synchronize(); torch.cuda.synchronize()
T = {}
def timed(name, fn):
    sync(); t0 = time.perf_counter(); fn(); sync()
    T.setdefault(name, []).append(1e3*(time.perf_counter() - t0))
lo, hi, chunk = domain.slice_bounds(N, world, rank)
for rep in range(3):
    tree.Upload(ps, only=IN)
    timed("shard_begin", lambda: tree.BuildTreeShardBegin(rank, world))
    timed("shard_finish", lambda: tree.BuildTreeShardFinish())
    timed("set_slice", lambda: tree.SetActiveSlice(lo, hi))
    timed("density_pass", lambda: tree.DensityPass())
    timed("derive", lambda: tree.DeriveDensityFields(lo, hi))
    timed("refresh_hmax", lambda: tree.RefreshHmax())
    timed("zero", lambda: tree.ZeroAccelerations())
    timed("forces_begin", lambda: tree.UpdateAllSphForcesBegin())
    timed("forces_finish", lambda: tree.UpdateAllSphForcesFinish())
    st = tree.Stats()
    T.setdefault("k_density", []).append(st.ms_kernel_density); T.setdefault("k_hydro", []).append(st.ms_kernel_hydro)
    T.setdefault("k_gravity", []).append(st.ms_kernel_gravity)
    T.setdefault("t_sort", []).append(st.ms_tree_sort); T.setdefault("t_levels", []).append(st.ms_tree_levels)
    T.setdefault("t_stock", []).append(st.ms_tree_stock)
print("N %d world %d rank %d" % (N, world, rank))
tot = 0.0
for k, v in T.items():
    print("%-14s %s  min %.2f" % (k, " ".join("%7.2f" % x for x in v), min(v)))
    if not k.startswith(("k_", "t_")):
        tot += min(v)
print("sum of phases (min) %.2f ms" % tot)
