"""BASELINE.json configurations 2-4 at their FULL sizes (config 1 = the adsod goldens, config 5 = bench.py): the oracle is
O(N^2), so parity here means size-independent properties of the reference's scheme that a wrong neighbour set, a
missed pair or a wrong term would break -- h = h_fac (m/rho)^(1/ndim) at convergence, pairwise antisymmetry of the
momentum exchange, the energy budget sum m (v.a + du/dt) = 0 of grad-h SPH, symmetry of the neighbour relation,
linearity of gravity in the masses (exact for a factor of 2) and agreement of two multipole orders within the MAC error.
Synthetic particle sets of the named shapes (gandalf_b200/workloads.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _particles(g, ic, ndim, fields):
    N = int(ic["m"].shape[0])
    ps = g.ParticleSet(N, ndim, "soa", fields=set(fields))
    for k in ("r", "v"):
        ps[k][...] = ic[k].T.cpu().numpy()
    for k in ("m", "h", "u"):
        ps[k][...] = ic[k].cpu().numpy()
    return ps, N


FIELDS = ("r", "v", "m", "h", "u", "rho", "a", "dudt", "div_v", "atree", "gpot", "levelneib", "hrangesqd")


def _budget(ps):
    """(|sum m a| / sum |m a| per component, |sum m (v.a + dudt)| / sum |m v.a|)"""
    ma = ps["m"][:, None]*ps["a"]
    mom = np.abs(ma.sum(0))/np.abs(ma).sum(0)
    va = (ps["v"]*ps["a"]).sum(1)
    en = abs((ps["m"]*(va + ps["dudt"])).sum())/np.abs(ps["m"]*va).sum()
    return mom, en


def test_config2_khi_million_2d_quintic():
    import gandalf_b200 as g
    from gandalf_b200 import workloads
    params, ic = workloads.khi(983552)
    ps, N = _particles(g, ic, 2, FIELDS)
    assert 900000 < N < 1100000
    tree = g.GpuSphTree(params)
    tree.BuildTree(True, 0, 1, 1, 0.0, ps)
    tree.UpdateAllSphProperties(); tree.ZeroAccelerations(); tree.UpdateAllSphHydroForces()
    tree.Download(ps)
    st = tree.Stats()
    assert st.n_active == N and 30 < st.ngb_mean < 60          # pi (3 h_fac)^2 = 40.7 for the 2-D quintic
    assert np.abs(ps["h"]/(1.2*np.sqrt(ps["m"]/ps["rho"])) - 1).max() < 1e-13
    inner = np.abs(ps["r"][:, 1]) < 0.2
    outer = np.abs(ps["r"][:, 1]) > 0.3
    assert abs(ps["rho"][inner].mean() - 2.0) < 0.02 and abs(ps["rho"][outer].mean() - 1.0) < 0.02
    mom, en = _budget(ps)
    assert np.all(mom < 1e-11), mom
    assert en < 1e-10, en
    off, idx = tree.NeighbourLists()
    rows = np.repeat(np.arange(N), np.diff(off))
    assert np.array_equal(np.sort(rows.astype(np.int64)*N + idx), np.sort(idx.astype(np.int64)*N + rows))
    tree.close()


def test_config3_bossbodenheimer_4m_gravity():
    import gandalf_b200 as g
    from gandalf_b200 import workloads
    params, ic = workloads.bossbodenheimer(4000000)
    ps, N = _particles(g, ic, 3, FIELDS)
    assert N == 4000000

    def forces(par, mass_factor=1.0, hydro=True):
        q = dict(par, hydro_forces=1 if hydro else 0)
        ps["m"][...] = ic["m"].cpu().numpy()*mass_factor
        ps["h"][...] = ic["h"].cpu().numpy()
        tree = g.GpuSphTree(q)
        tree.BuildTree(True, 0, 1, 1, 0.0, ps)
        tree.UpdateAllSphProperties(); tree.ZeroAccelerations(); tree.UpdateAllSphForces()
        tree.Download(ps)
        st = tree.Stats()
        tree.close()
        return {k: ps[k].copy() for k in ("a", "atree", "gpot", "h", "rho", "dudt", "m", "v")}, st

    full, st = forces(params)
    assert st.n_active == N and 50 < st.ngb_mean < 70 and st.ncell_mean > 0 and st.ndirect_mean > st.ngb_mean
    assert np.abs(full["h"]/(1.2*np.cbrt(full["m"]/full["rho"])) - 1).max() < 1e-13
    assert np.all(full["gpot"] > 0)                                 # the reference's sign convention: gpot = +sum m/r
    # tree gravity: momentum is conserved to the MAC error, not to rounding
    mat = full["m"][:, None]*full["atree"]
    assert np.all(np.abs(mat.sum(0)) < 2e-3*np.abs(mat).sum(0))
    # gravity only: linear in the masses -- a factor of 2 is exact in binary arithmetic, every list is the same
    g1, _ = forces(params, 1.0, hydro=False)
    g2, _ = forces(params, 2.0, hydro=False)
    assert np.array_equal(g2["atree"], 2.0*g1["atree"]) and np.array_equal(g2["gpot"], 2.0*g1["gpot"])
    assert np.array_equal(g2["h"], g1["h"]) and np.array_equal(g2["rho"], 2.0*g1["rho"])
    # two multipole orders, same opening criterion: they differ by the truncation error only
    mono, _ = forces(dict(params, multipole="monopole"), hydro=False)
    quad, _ = forces(dict(params, multipole="quadrupole"), hydro=False)
    amag = np.linalg.norm(quad["atree"], axis=1)
    for other, bar in ((mono, 1e-2), (g1, 1e-2)):                  # g1 = fast_monopole; theta^2 = 0.15: ~0.4 % median
        err = np.linalg.norm(other["atree"] - quad["atree"], axis=1)/amag
        assert np.median(err) < bar and np.percentile(err, 99) < 5*bar, (np.median(err), np.percentile(err, 99))


def test_config4_glass_16m_hydro():
    import gandalf_b200 as g
    from gandalf_b200 import workloads
    params, ic = workloads.glass(252**3)
    ps, N = _particles(g, ic, 3, ("r", "v", "m", "h", "u", "rho", "a", "dudt"))
    assert N == 252**3
    tree = g.GpuSphTree(params)
    tree.BuildTree(True, 0, 1, 1, 0.0, ps)
    tree.UpdateAllSphProperties(); tree.ZeroAccelerations(); tree.UpdateAllSphHydroForces()
    tree.Download(ps, only=("h", "rho", "a", "dudt"))
    st = tree.Stats()
    assert st.n_active == N and 57 < st.ngb_mean < 66               # gather count 57.9; the force pass also
                                                                     # counts the neighbours that only see the particle
    assert np.abs(ps["h"]/(1.2*np.cbrt(ps["m"]/ps["rho"])) - 1).max() < 1e-13
    assert abs(ps["rho"].mean() - 1.0) < 0.03                       # (the SPH estimate of a jittered lattice is biased high)
    mom, en = _budget(ps)
    assert np.all(mom < 1e-11), mom
    assert en < 1e-10, en
    tree.close()
