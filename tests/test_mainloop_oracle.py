"""The main-loop restatement (tests/oracle_py.py: mainloop, global_timestep) against the reference's
own SphSimulation::MainLoop: tests/golden/loop_*.npz hold the reference's particle state after its
setup and after N of its own steps (generator: tests/golden/make_loop_golden.py).  This pins the
checker of the device-resident main loop (gsph_integrator_*, tests/test_gpu_mainloop.py)."""
import numpy as np
import pytest

import oracle_py as orc
from conftest import load_golden, LOOP_CASES, rel_err

CPU_CASES = [c for c in LOOP_CASES if c != "loop_sphere_grav_quad"]   # (that one mixes rebuild / restock / extrapolate)


def split(d):
    out = {"setup": {}, "loop": {}}
    for k, v in d.items():
        st, f = k.split(".", 1)
        out[st][f] = v
    return out["setup"], out["loop"]


def vec_err(a, b):
    a = a.reshape(len(a), -1); b = b.reshape(len(b), -1)
    return np.linalg.norm(a - b, axis=1)/np.maximum(np.linalg.norm(b, axis=1), np.abs(b).max()*1e-6)


def check_state(s, ref, tol):
    for f in ("r", "v", "a"):
        assert vec_err(np.asarray(s[f]), ref[f]).max() < tol, f
    for f in ("h", "rho", "u", "alpha"):
        assert rel_err(np.asarray(s[f]), ref[f]).max() < tol, f
    sc = np.abs(ref["dudt"]).max() + 1e-300
    assert (np.abs(np.asarray(s["dudt"]) - ref["dudt"])/sc).max() < tol


def test_loop_fixtures_exist():
    assert len(LOOP_CASES) >= 5


@pytest.mark.parametrize("name", LOOP_CASES)
def test_first_timestep_matches_reference(name):
    """Simulation::ComputeGlobalTimestep on the reference's post-setup state: its first MainLoop
    advances t by exactly that step"""
    params, d = load_golden(name)
    setup, loop = split(d)
    p = orc.params_from_dict(params)
    ts = orc.global_timestep(p, setup, 0.15, 0.3, 0.4, params.get("gas_eos", "energy_eqn") == "energy_eqn")
    assert 0 < ts < float(loop["t"])
    assert float(loop["t"]) / int(loop["steps"]) < 3*ts


@pytest.mark.parametrize("name", CPU_CASES)
def test_mainloop_restatement_matches_reference(name):
    params, d = load_golden(name)
    setup, loop = split(d)
    s, t, ts = orc.mainloop(params, setup, int(loop["steps"]))
    assert abs(t - float(loop["t"])) <= 1e-12*abs(t)
    assert abs(ts - float(loop["timestep"])) <= 1e-9*abs(ts)
    check_state(s, loop, 1e-9)
    if "star_r" in loop:                     # star particles moved by the N-body leapfrog
        for k in ("star_r", "star_v", "star_a"):
            assert vec_err(np.asarray(s[k]), loop[k]).max() < 1e-9, k
        if params.get("nbody", "hermite4") == "hermite4":     # NbodyHermite4: jerk, snap, crackle (single-precision
            for k in ("star_adot", "star_a2dot", "star_a3dot"):   # powf in the reference's wmean and dt^3)
                assert vec_err(np.asarray(s[k]), loop[k].reshape(np.asarray(s[k]).shape)).max() < 1e-6, k


# ---------------------------------------------------------------------------------------------------
# hierarchical block time steps (Nlevels > 1): tests/golden/loopblk_*.npz, make_loopblk_golden.py
# ---------------------------------------------------------------------------------------------------
from conftest import LOOPBLK_CASES  # noqa: E402

SIM_KEYS = ("n", "nresync", "level_max", "level_step", "dt_max", "timestep", "t")
INT_STATE = ("level", "levelneib", "nstep", "nlast")


def split_blk(d):
    """-> (setup arrays, setup scalars, loop arrays, loop scalars)"""
    arr = {"setup": {}, "loop": {}}
    sim = {"setup": {}, "loop": {}}
    for k, v in d.items():
        st, f = k.split(".", 1)
        if f.startswith("sim_"):
            sim[st][f[4:]] = v.item()
        elif f != "steps":
            arr[st][f] = v
    return arr["setup"], sim["setup"], arr["loop"], sim["loop"]


def check_block_state(s, sim, ref, refsim, tol):
    for k in ("n", "nresync", "level_max", "level_step"):
        assert int(sim[k]) == int(refsim[k]), k
    for k in ("dt_max", "timestep", "t"):
        assert abs(float(sim[k]) - float(refsim[k])) <= 1e-11*abs(float(refsim[k])), k
    for f in INT_STATE:
        assert np.array_equal(np.asarray(s[f]).astype(np.int64), ref[f].astype(np.int64)), f
    check_state(s, ref, tol)
    for f in ("dt", "tlast"):
        assert np.abs(np.asarray(s[f]) - ref[f]).max() <= 1e-11*np.abs(ref[f]).max(), f


def check_block_stars(s, ref, params, tol):
    """star particles on their own block time steps: identical levels and integer times, state to `tol` (the jerk and
    the higher derivatives of Hermite-4 pass through single-precision powf in the reference: 1e-5)"""
    if "star_r" not in ref:
        return
    for k in ("level", "nstep", "nlast"):
        assert np.array_equal(np.asarray(s["star_" + k]).astype(np.int64), ref["star_" + k].astype(np.int64)), k
    assert np.abs(np.asarray(s["star_tlast"]) - ref["star_tlast"]).max() <= 1e-11*np.abs(ref["star_tlast"]).max()
    for k in ("r", "v", "a"):
        a = np.asarray(s["star_" + k])
        assert vec_err(a, ref["star_" + k].reshape(a.shape)).max() < tol, k
    if params.get("nbody", "hermite4") == "hermite4":
        for k in ("adot", "a2dot", "a3dot"):
            a = np.asarray(s["star_" + k])
            assert vec_err(a, ref["star_" + k].reshape(a.shape)).max() < max(tol, 1e-5), k


def test_block_fixtures_exist():
    assert len(LOOPBLK_CASES) >= 3


@pytest.mark.parametrize("name", LOOPBLK_CASES)
def test_block_timestep_restatement_matches_reference(name):
    """oracle_py.mainloop_block (AdvanceParticles / CheckTimesteps / ComputeBlockTimesteps / EndTimestep with
    Nlevels = 4) reproduces 20 of the reference's own MainLoop calls: the same levels, integer times and step
    sizes for every particle -- through added and removed levels and a resynchronisation -- and the same state"""
    params, d = load_golden(name)
    setup, sim0, loop, sim1 = split_blk(d)
    single = int(params.get("sph_single_timestep", 0)) == 1    # (then the gas shares ONE level by construction)
    assert int(params["Nlevels"]) > 1 and (len(np.unique(loop["level"])) >= 3 or single)
    steps = int(d["loop.steps"])
    if not name.startswith("loopblk_adsod"):
        steps = 20                     # (kept: about 10 s each)
    s, sim = orc.mainloop_block(params, setup, sim0, steps)
    check_block_state(s, sim, loop, sim1, 1e-9)
    check_block_stars(s, loop, params, 1e-9)
