"""Pin the CPU oracle (oracle/gsph_oracle.c) against golden vectors dumped from the UNMODIFIED
reference (tests/golden/make_golden.py).  CPU only.

Tolerances: neighbour lists bit-exact; floating point 1e-10 relative (the north-star bar is 1e-6;
the oracle and the reference differ only in summation order and in libm call sites).
"""
import numpy as np
import pytest

import oracle_py as orc
from conftest import load_golden, stage, rel_err, vec_rel_err, scaled_err, GOLDEN_CASES

TOL = 1e-10


def _state_by_iorig(st):
    """reorder a dumped stage so that row i is the particle with iorig == i"""
    order = np.argsort(st["iorig"])
    return {k: v[order] for k, v in st.items()}


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_kernel_tables_match_dumped_reference_fields(name):
    """hfactor/hrangesqd in the dumps are simple functions of h: checks the dump itself."""
    params, d = load_golden(name)
    s = stage(d, "dens_out")
    nd = params["ndim"]
    kr = 2.0 if params["kernel"] == "m4" else 3.0
    assert np.array_equal(s["hrangesqd"], kr * kr * s["h"] * s["h"])
    assert rel_err(s["hfactor"], (1.0 / s["h"]) ** (nd + 1)).max() < 1e-14


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_neighbour_lists_bit_exact(name):
    params, d = load_golden(name)
    if "force_in.neib_off" not in d:
        pytest.skip("no neighbour lists dumped for mirror boundaries (the lists hold ghost images)")
    p = orc.params_from_dict(params)
    s = _state_by_iorig(stage(d, "force_in"))
    off, idx = orc.neighbour_lists(p, s["r"], s["hrangesqd"])
    assert np.array_equal(off, d["force_in.neib_off"])
    assert np.array_equal(idx, d["force_in.neib_idx"])


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_density(name):
    params, d = load_golden(name)
    p = orc.params_from_dict(params)
    si = stage(d, "dens_in")
    so = stage(d, "dens_out")
    out = orc.density(p, si["r"], si["m"], si["h"], si["u"], stars=d.get("stars"))
    if "stars" in d:
        # the star terms are a sizeable part of zeta near the stars (otherwise this case pins nothing)
        plain = orc.density(p, si["r"], si["m"], si["h"], si["u"])
        assert (np.abs(plain["zeta"] - out["zeta"]) > 1e-3*np.abs(out["zeta"])).sum() > 100
    for f in ("h", "rho", "invomega", "zeta", "hfactor", "hrangesqd", "u", "sound", "pressure"):
        floor = np.abs(so[f]).max() * 1e-6 if f == "zeta" else 0.0
        e = rel_err(out[f], so[f], floor).max()
        assert e < TOL, "%s: %s rel err %.3e" % (name, f, e)
    if params.get("time_dependent_avisc", "none") == "cd2010":
        # the Cullen & Dehnen switch at the end of ComputeH (Sph.h:364-456)
        al, da = orc.cd2010_switch(p, si["r"], si["v"], si["a"], si["m"], out, si["alpha"])
        assert scaled_err(al, so["alpha"]).max() < TOL and scaled_err(da, so["dalphadt"]).max() < 1e-9


@pytest.mark.parametrize("name", [c for c in GOLDEN_CASES if not c.startswith("sphere_grav")])
def test_hydro_forces(name):
    params, d = load_golden(name)
    p = orc.params_from_dict(params)
    si = stage(d, "force_in")
    so = stage(d, "force_out")
    out = orc.hydro_forces(p, si)
    assert vec_rel_err(out["a"], so["a"]).max() < TOL
    for f in ("dudt", "div_v", "dalphadt"):
        assert scaled_err(out[f], so[f]).max() < TOL, f
    assert np.array_equal(out["levelneib"], so["levelneib"])


def _mac_inputs(d):
    """atree / gpot at tree-build time (the reference stocks amin / macfactor from them) by iorig"""
    if "dens_in.atree" not in d:
        return None, None
    order = np.argsort(d["dens_in.iorig"])
    return d["dens_in.atree"][order], d["dens_in.gpot"][order]


@pytest.mark.parametrize("name", [c for c in GOLDEN_CASES if c.startswith("sphere_grav")])
def test_kdtree_matches_reference(name):
    params, d = load_golden(name)
    p = orc.params_from_dict(params)
    si = _state_by_iorig(stage(d, "force_in"))
    t = orc.Tree(p, si["r"], si["v"], si["m"], si["h"])
    hdr = d["force_in.tree_hdr"]
    assert t.ncell == hdr[0] and t.ltot == hdr[1]
    # identical leaf membership (the tree is built on the array in iorig order in both cases)
    assert np.array_equal(t.leaf_of(), d["force_in.tree_leaf_of_iorig"])
    c = t.cells()
    assert np.array_equal(c["N"], d["force_in.tree_meta"][:, 2])
    live = c["N"] > 0
    assert rel_err(c["box"][live], d["force_in.tree_box"][live]).max() < 1e-13
    assert np.abs(c["com"][live] - d["force_in.tree_com"][live]).max() < 1e-13
    assert rel_err(c["scalars"][live][:, :4], d["force_in.tree_scalars"][live][:, :4]).max() < 1e-12
    if params["multipole"] in ("quadrupole", "fast_quadrupole"):
        qs = np.abs(d["force_in.tree_q"]).max()
        assert (np.abs(c["q"] - d["force_in.tree_q"]) / qs).max() < 1e-12


@pytest.mark.parametrize("name", [c for c in GOLDEN_CASES if c.startswith("sphere_grav")])
def test_tree_gravity_forces(name):
    params, d = load_golden(name)
    p = orc.params_from_dict(params)
    si = _state_by_iorig(stage(d, "force_in"))
    so = _state_by_iorig(stage(d, "force_out"))
    atree, gpot = _mac_inputs(d)
    t = orc.Tree(p, si["r"], si["v"], si["m"], si["h"], atree, gpot)
    out = t.hydro_grav_forces(si)
    if "stars" in d:                 # ComputeStarGravForces after the tree forces (GradhSphTree.cpp:600-607)
        st = d["stars"]
        sa, sat, sg = orc.star_grav_forces(p, si["r"], si["h"], st[:, :3], st[:, 3], st[:, 4])
        out["a"] = out["a"] + sa
        out["atree"] = out["atree"] + sat
        out["gpot"] = out["gpot"] + sg
    # interaction-list sizes equal the reference's
    nd_ref = np.diff(d["grav_in.direct_off"])
    assert np.array_equal(out["ndirect"], nd_ref)
    assert vec_rel_err(out["atree"], so["atree"]).max() < TOL
    assert vec_rel_err(out["a"], so["a"]).max() < TOL
    assert rel_err(out["gpot"], so["gpot"]).max() < TOL
    assert scaled_err(out["dudt"], so["dudt"]).max() < TOL
    assert scaled_err(out["div_v"], so["div_v"]).max() < TOL


@pytest.mark.parametrize("name", [c for c in GOLDEN_CASES if c.startswith("sphere_grav")])
def test_direct_gravity_close_to_tree(name):
    """exact direct summation agrees with the reference's tree force within the MAC error."""
    params, d = load_golden(name)
    p = orc.params_from_dict(params)
    si = _state_by_iorig(stage(d, "force_in"))
    so = _state_by_iorig(stage(d, "force_out"))
    out = orc.hydro_grav_forces_direct(p, si)
    if "stars" in d:
        st = d["stars"]
        out["atree"] = out["atree"] + orc.star_grav_forces(p, si["r"], si["h"], st[:, :3], st[:, 3], st[:, 4])[1]
    e = vec_rel_err(out["atree"], so["atree"])
    assert np.sqrt(np.mean(e ** 2)) < 2e-2


def test_star_gas_forces_match_reference():
    """HydroTree::UpdateAllStarGasForces: gravity of the gas on the star particles (a, gpot, and adot of
    the Hermite-4 class) from the reference's own run (tests/golden/stargas_hermite4.npz)"""
    params, d = load_golden("stargas_hermite4")
    p = orc.params_from_dict(params)
    o = np.argsort(d["force_in.iorig"])
    r, v, m, h = (d["force_in." + f][o] for f in ("r", "v", "m", "h"))
    t = orc.Tree(p, r, v, m, h)
    a, adot, gpot, counts = t.star_gas_forces(d["stargas.r"], d["stargas.v"], d["stargas.h"])
    assert (counts[:, 0] > 0).any() and (counts[:, 2] > 0).all()      # smoothed neighbours and cells both occur
    assert vec_rel_err(a, d["stargas.a"]).max() < TOL
    assert rel_err(gpot, d["stargas.gpot"]).max() < TOL
    assert vec_rel_err(adot, d["stargas.adot"]).max() < 1e-6           # (single-precision powf in the reference)


def test_bruteforce_reference_is_the_all_opened_tree():
    """the reference's SECOND, tree-independent path (neib_search = bruteforce -> GradhSphBruteForce: every pair is a
    kernel neighbour or an exact direct sum; golden brute_sphere_grav from make_golden.py) pins (a) the oracle's exact
    direct summation and (b) the tree path with an opening criterion that opens every cell (thetamaxsqd -> 0): the
    direct-sum and softened pair terms are the ones the tree path uses, without any cell in between"""
    params, d = load_golden("brute_sphere_grav")
    assert params["neib_search"] == "bruteforce"
    p = orc.params_from_dict(params)
    si = _state_by_iorig(stage(d, "force_in"))
    so = _state_by_iorig(stage(d, "force_out"))
    out = orc.hydro_grav_forces_direct(p, si)
    assert vec_rel_err(out["a"], so["a"]).max() < TOL
    assert rel_err(out["gpot"], so["gpot"]).max() < TOL
    assert scaled_err(out["dudt"], so["dudt"]).max() < TOL
    p2 = orc.params_from_dict(dict(params, thetamaxsqd=1e-30))
    t = orc.Tree(p2, si["r"], si["v"], si["m"], si["h"])
    out2 = t.hydro_grav_forces(si)
    assert out2["ncell"].max() == 0 if "ncell" in out2 else True
    assert vec_rel_err(out2["a"], so["a"]).max() < TOL
    assert vec_rel_err(out2["atree"], so["atree"]).max() < TOL
    assert rel_err(out2["gpot"], so["gpot"]).max() < TOL
