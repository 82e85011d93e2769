import sys, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import gandalf_b200 as g, oracle_py as orc
rng = np.random.default_rng(7)
for N in (7, 13, 100, 1237):
    r = rng.normal(size=(N, 3))*np.where(rng.uniform(size=(N, 1)) < 0.3, 0.1, 1.0)
    m = np.full(N, 1.0/N); h = np.full(N, 0.5); u = np.ones(N)
params = {"ndim": 3, "kernel": "m4", "tabulated_kernel": 1}
p = orc.params_from_dict(params)
ps = g.ParticleSet(N, 3).set(r=r, m=m, h=h, u=u)
tree = g.GpuSphTree(params)
tree.BuildTree(True, 0, 1, 1, 0.0, ps)
tree.UpdateAllSphProperties(); tree.Download(ps)
od = orc.density(p, r, m, h, u)
e = np.abs(ps["h"]/od["h"]-1)
bad = np.where(e > 1e-9)[0]
print("bad", bad, e[bad], od["niter"][bad], "gpu h", ps["h"][bad], "orc h", od["h"][bad])
st = tree.Stats(); print("retries", st.density_retries, "itmax", st.hiter_max, "oracle itmax", od["niter"].max())
leaf = tree.TreeLeafOf()
for b in bad[:5]:
    mates = np.where(leaf == leaf[b])[0]
    print("leaf mates of", b, mates, "h:", ps["h"][mates], od["h"][mates], od["niter"][mates])
# self-consistency: rho(h_last) relation. recompute rho at GPU h by brute force with w0 table
tab = orc.kernel_table(p, 0)
def rho_at(i, hh):
    d2 = ((r - r[i])**2)
    d2 = d2[:,0] + d2[:,1] + d2[:,2]
    s2 = d2/(hh*hh)
    idx = np.minimum((s2*250.0).astype(int), 999)
    w = np.where(s2 < 4.0, tab[idx], 0.0)
    return (m*w).sum()/hh**3, int((s2<4.0).sum())
for b in bad[:6]:
    # GPU final h = 1.2*(m/rho)^(1/3)  => rho_gpu known
    print(b, "gpu rho", ps["rho"][b], "orc rho", od["rho"][b], "h_gpu", ps["h"][b], "h_orc", od["h"][b])
    # the rho was computed at the last iterate h_last, unknown; scan: find h_last s.t. brute rho == gpu rho
    for hh in (0.5,):
        print("   brute rho at h0=0.5:", rho_at(b, 0.5))
