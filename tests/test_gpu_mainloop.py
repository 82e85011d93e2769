"""Device-resident main loop (gsph_integrator_begin / _step) against the reference's own
SphSimulation::MainLoop: the golden fixtures tests/golden/loop_*.npz hold the reference's particle state
after its setup and after N of its own steps (lfkdk, one global time step).  The GPU starts from the
setup state (compute_forces = 0: nothing is recomputed, like a restart) and runs the same N steps
without the particles ever leaving the device; positions, velocities, accelerations, h, rho, u, alpha,
the simulation time and the next time step must agree.  Everything goes through the C ABI."""
import numpy as np
import pytest

import oracle_py as orc
from conftest import load_golden, LOOP_CASES, rel_err
from test_mainloop_oracle import split, check_state, vec_err

pytestmark = pytest.mark.gpu
TOL = 1e-8          # asserted after 4-6 steps (rounding differences of the force sums feed back through the
                    # integration; the single operator calls are held to 1e-9); the north-star bar is 1e-6


def _upload_setup(g, params, setup):
    N = len(setup["m"])
    nd = int(params["ndim"])
    ps = g.ParticleSet(N, nd)
    for f in ("r", "v", "a", "atree"):
        ps[f][...] = setup[f].reshape(N, nd)
    for f in ("m", "h", "u", "sound", "dudt", "div_v", "alpha", "dalphadt", "gpot"):
        ps[f][...] = setup[f]
    for f in ("rho", "invomega", "zeta", "hfactor", "hrangesqd", "pressure"):
        if f in setup:                                       # block time steps: inactive particles keep these
            ps[f][...] = setup[f]
    ps["ptype"][...] = setup["ptype"]
    ps["level"][...] = setup["level"]
    tree = g.GpuSphTree(params)
    tree.Upload(ps)
    return ps, tree


@pytest.mark.parametrize("name", LOOP_CASES)
def test_device_mainloop_matches_reference(name):
    import gandalf_b200 as g
    params, d = load_golden(name)
    setup, loop = split(d)
    ps, tree = _upload_setup(g, params, setup)
    hermite = "star_r" in setup and params.get("nbody", "hermite4") == "hermite4"
    if "star_r" in setup:                    # star particles: the reference's N-body integrator around the device passes
        extra = {k: setup["star_" + k] for k in ("adot", "a2dot", "a3dot")} if hermite else {}
        tree.IntegratorSetStars(setup["star_r"], setup["star_v"], setup["star_m"], setup["star_h"], a=setup["star_a"], **extra)
    tree.IntegratorBegin(compute_forces=False)
    t0, ts0, n0 = tree.IntegratorState()
    assert t0 == 0.0 and n0 == 0 and ts0 > 0
    steps = int(loop["steps"])
    tree.MainLoop(steps)
    t, ts, n = tree.IntegratorState()
    assert n == steps
    # same time-step sequence.  (Hermite-4 stars: the Aarseth step follows the jerk, and the reference's gas-on-star jerk
    # goes through single-precision powf -- 1e-7 -- where the device kernel works in double precision.)
    assert abs(t - float(loop["t"])) <= (1e-8 if hermite else 1e-12)*abs(t)
    assert abs(ts - float(loop["timestep"])) <= (1e-6 if hermite else 1e-9)*abs(ts)
    tree.Download(ps)
    check_state({k: ps[k] for k in ("r", "v", "a", "h", "rho", "u", "alpha", "dudt")}, loop, TOL)
    if "star_r" in setup:
        st = tree.IntegratorGetStars()
        for k in ("r", "v", "a"):
            assert vec_err(st[k], loop["star_" + k]).max() < TOL, k
        assert rel_err(st["gpot"], loop["star_gpot"]).max() < TOL
        if hermite:                          # jerk etc.: the reference's gas-on-star jerk goes through single-precision powf
            for k in ("adot", "a2dot", "a3dot"):
                assert vec_err(st[k], loop["star_" + k].reshape(st[k].shape)).max() < 1e-5, k
    tree.close()


def test_device_hermite4_stars_from_initial_conditions():
    """compute_forces = 1 with nbody = hermite4 (the reference's default): IntegratorBegin computes the stars' a, adot
    (gas + stars) and the start-up a2dot, a3dot (NbodyHermite4::CalculateAllStartupQuantities) itself -- they must be
    the ones the reference's setup left, and the run must then follow the reference"""
    import gandalf_b200 as g
    params, d = load_golden("loop_sphere_grav_stars_hermite4")
    setup, loop = split(d)
    assert params["nbody"] == "hermite4"
    N = len(setup["m"])
    ps = g.ParticleSet(N, 3)
    for f in ("r", "v"):
        ps[f][...] = setup[f].reshape(N, 3)
    for f in ("m", "h", "u"):
        ps[f][...] = setup[f]
    tree = g.GpuSphTree(params)
    tree.IntegratorSetStars(setup["star_r"], setup["star_v"], setup["star_m"], setup["star_h"])
    tree.IntegratorBegin(ps, compute_forces=True)
    st = tree.IntegratorGetStars()
    # (begin converges h once more from the reference's converged h: the fixed point moves inside h_converge, so the
    #  gas state -- and with it the force on a star inside the cloud -- agrees to ~1e-4, not to rounding)
    for k, tol in (("a", 1e-3), ("adot", 1e-2), ("a2dot", 1e-2), ("a3dot", 1e-2)):
        ref = setup["star_" + k].reshape(st[k].shape)
        assert np.abs(st[k] - ref).max() <= tol*np.abs(ref).max(), (k, np.abs(st[k] - ref).max()/np.abs(ref).max())
    t0, ts0, _ = tree.IntegratorState()
    steps = int(loop["steps"])
    tree.MainLoop(steps)
    t, ts, n = tree.IntegratorState()
    assert n == steps and abs(t - float(loop["t"])) <= 1e-2*abs(t)
    st = tree.IntegratorGetStars()
    for k in ("r", "v", "a"):
        assert vec_err(st[k], loop["star_" + k]).max() < 2e-2, k
    tree.close()


def test_device_mainloop_step_by_step_equals_one_call():
    """MainLoop(1) x N == MainLoop(N) bit for bit, and the state can be downloaded in between"""
    import gandalf_b200 as g
    params, d = load_golden("loop_khi2d_quintic")
    setup, loop = split(d)
    res = []
    for chunks in ((4,), (1, 1, 2)):
        ps, tree = _upload_setup(g, params, setup)
        tree.IntegratorBegin(compute_forces=False)
        for c in chunks:
            tree.MainLoop(c)
            tree.Download(ps)
        res.append(({k: ps[k].copy() for k in ("r", "v", "a", "h", "u")}, tree.IntegratorState()))
        tree.close()
    assert res[0][1] == res[1][1]
    for k in res[0][0]:
        assert np.array_equal(res[0][0][k], res[1][0][k]), k


def test_device_mainloop_from_initial_conditions():
    """compute_forces = 1: begin runs the first density + force pass itself; the first step then equals
    the oracle's main loop started from the GPU's own post-setup state"""
    import gandalf_b200 as g
    params, d = load_golden("loop_box3d_mm97")
    setup, loop = split(d)
    N = len(setup["m"])
    ps = g.ParticleSet(N, 3)
    for f in ("r", "v"):
        ps[f][...] = setup[f].reshape(N, 3)
    for f in ("m", "h", "u", "alpha"):
        ps[f][...] = setup[f]
    tree = g.GpuSphTree(params)
    tree.IntegratorBegin(ps, compute_forces=True)
    tree.Download(ps)
    st = {k: ps[k].copy() for k in ("r", "v", "a", "atree", "m", "h", "u", "sound", "dudt", "div_v", "alpha", "dalphadt",
                                    "gpot", "ptype", "level")}
    t0, ts0, _ = tree.IntegratorState()
    p = orc.params_from_dict(params)
    assert abs(ts0 - orc.global_timestep(p, st, 0.15, 0.3, 0.4, True)) <= 1e-14*ts0
    tree.MainLoop(2)
    t, ts, n = tree.IntegratorState()
    tree.Download(ps)
    so, to, tso = orc.mainloop(params, st, 2)
    assert abs(t - to) <= 1e-12*t and abs(ts - tso) <= 1e-9*ts
    check_state({k: ps[k] for k in ("r", "v", "a", "h", "rho", "u", "alpha", "dudt")}, so, TOL)
    tree.close()


def test_device_mainloop_errors_are_loud():
    import gandalf_b200 as g
    params, d = load_golden("loop_adsod")
    setup, loop = split(d)
    tree = g.GpuSphTree(params)
    with pytest.raises(g.GsphError):
        tree.MainLoop(1)                                   # before IntegratorBegin
    with pytest.raises(g.GsphError):
        tree.IntegratorBegin(compute_forces=False)         # before Upload
    ps, tree2 = _upload_setup(g, params, setup)
    tree2.SetStars(np.zeros((1, 1)), np.ones(1), np.ones(1))
    with pytest.raises(g.GsphError):
        tree2.IntegratorBegin(compute_forces=False)        # stars are not integrated on the device
    with pytest.raises(ValueError):
        g.GpuSphTree(dict(params, sink_particles=1)).IntegratorBegin(ps)   # no sink search / accretion on this path
    tree.close(); tree2.close()


def test_snapshot_of_a_device_resident_run_and_restart(tmp_path):
    """su snapshot straight from the device (gandalf_b200/snapshot.py): equals the downloaded
    state, carries t and the step count, and a run restarted from it (the reference's restart semantics:
    the first density / force pass is redone from the snapshot's h) steps on"""
    import gandalf_b200 as g
    from gandalf_b200 import snapshot
    params, d = load_golden("loop_box3d_mm97")
    setup, loop = split(d)
    ps, tree = _upload_setup(g, params, setup)
    tree.IntegratorBegin(compute_forces=False)
    tree.MainLoop(2)
    path = str(tmp_path / "run.su.00002")
    tree.WriteSnapshot(path)
    tree.Download(ps)
    t, ts, n = tree.IntegratorState()
    s = snapshot.read_su(path)
    assert s["t"] == t and s["Nsteps"] == n == 2 and s["N"] == len(setup["m"])
    o = np.argsort(s["iorig"])
    for f in ("m", "h", "rho", "u"):
        assert np.array_equal(s[f][o], ps[f]), f
    assert np.array_equal(s["r"][o], ps["r"]) and np.array_equal(s["v"][o], ps["v"])
    tree.MainLoop(1)                                   # the uninterrupted run, one step further
    tree.Download(ps)
    tree.close()
    ps2, hdr = g.GpuSphTree.ReadSnapshot(path)
    tree2 = g.GpuSphTree(params)
    tree2.IntegratorBegin(ps2, compute_forces=True, t=hdr["t"])
    tree2.MainLoop(1)
    t2, _, _ = tree2.IntegratorState()
    assert t2 > hdr["t"]
    tree2.Download(ps2)
    # the restarted run follows the uninterrupted one (it lost alpha, dalpha/dt and the old accelerations,
    # as a restart of the reference does): same flow one step later
    assert rel_err(ps2["rho"], ps["rho"]).max() < 0.05
    assert np.abs(ps2["r"] - ps["r"]).max() < 0.1*ps["h"].min()
    tree2.close()


def test_snapshot_keeps_the_star_particles(tmp_path):
    """a device-resident run with star particles writes them into its su snapshot (the reference always writes
    nbody->stardata, SimulationIO.hpp:2212-2250) and a restart from that file carries them on"""
    import gandalf_b200 as g
    params, d = load_golden("loop_sphere_grav_stars")
    setup, loop = split(d)
    ps, tree = _upload_setup(g, params, setup)
    tree.IntegratorSetStars(setup["star_r"], setup["star_v"], setup["star_m"], setup["star_h"], a=setup["star_a"])
    tree.IntegratorBegin(compute_forces=False)
    tree.MainLoop(2)
    path = str(tmp_path / "stars.su.00002")
    tree.WriteSnapshot(path)
    st = tree.IntegratorGetStars()
    tree.close()
    ps2, hdr = g.GpuSphTree.ReadSnapshot(path)
    assert "stars" in hdr and len(hdr["stars"]["m"]) == len(setup["star_m"])
    for k in ("r", "v", "m", "h"):
        assert np.array_equal(np.asarray(hdr["stars"][k]).reshape(np.asarray(st[k]).shape), st[k]), k
    tree2 = g.GpuSphTree(params)
    tree2.Upload(ps2)
    tree2.IntegratorSetStars(hdr["stars"]["r"], hdr["stars"]["v"], hdr["stars"]["m"], hdr["stars"]["h"])
    tree2.IntegratorBegin(compute_forces=True, t=hdr["t"])
    tree2.MainLoop(1)
    st2 = tree2.IntegratorGetStars()
    assert len(st2["m"]) == len(st["m"]) and np.all(np.isfinite(st2["r"]))
    assert np.abs(st2["r"] - st["r"]).max() < 0.1*np.abs(st["r"]).max()        # they moved on from where they were
    tree2.close()


def test_operator_calls_inside_a_resident_run_are_consistent():
    """after MainLoop the tree-order copies are older than the host-order state (v, u were corrected):
    an operator call without a new BuildTree is refused; after BuildTree its results do come back through
    Download (they used to be dropped because the resident loop skips the tree-order -> host-order scatter);
    sinks are refused up front"""
    import gandalf_b200 as g
    params, d = load_golden("loop_box3d_mm97")
    setup, loop = split(d)
    ps, tree = _upload_setup(g, params, setup)
    tree.IntegratorBegin(compute_forces=False)
    tree.MainLoop(1)
    with pytest.raises(g.GsphError):
        tree.UpdateAllSphProperties()
    tree.Download(ps)
    rho1 = ps["rho"].copy()
    tree.BuildTree(True, 0, 1, 1, 0.0)                 # tree-order copies of the CURRENT host-order state
    tree.UpdateAllSphProperties()
    ps["rho"][...] = -1.0
    tree.Download(ps)
    assert rel_err(ps["rho"], rho1).max() < 1e-12      # same positions, h converged -> same density, and it arrived
    tree.MainLoop(1)                                   # the loop carries on
    tree.close()
    with pytest.raises(ValueError):
        g.GpuSphTree(dict(params, sink_particles=1)).IntegratorBegin(ps)
    with pytest.raises(ValueError):
        g.GpuSphTree(dict(params, create_sinks=1)).IntegratorBegin(ps)


# ---------------------------------------------------------------------------------------------------
# hierarchical block time steps (Nlevels > 1)
# ---------------------------------------------------------------------------------------------------
from conftest import LOOPBLK_CASES  # noqa: E402
from test_mainloop_oracle import split_blk, check_block_state  # noqa: E402


@pytest.mark.parametrize("name", LOOPBLK_CASES)
def test_device_block_timesteps_match_reference(name):
    """the device-resident loop with Nlevels = 4 against 20 of the reference's own MainLoop calls: IntegratorBegin
    rebuilds the hierarchy the reference's setup chose (same level / nstep / dt for every particle), then levels are
    added and removed, steps are cut short by the neighbour-level limiter, the hierarchy is resynchronised -- the
    integer state of every particle is the reference's, the physical state agrees to rounding"""
    import gandalf_b200 as g
    params, d = load_golden(name)
    setup, sim0, loop, sim1 = split_blk(d)
    ps, tree = _upload_setup(g, params, setup)
    stars = "star_r" in setup
    hermite = stars and params.get("nbody", "hermite4") == "hermite4"
    if stars:                                # star particles on their own block time steps (N-body integrator on the host)
        extra = {k: setup["star_" + k] for k in ("adot", "a2dot", "a3dot")} if hermite else {}
        tree.IntegratorSetStars(setup["star_r"], setup["star_v"], setup["star_m"], setup["star_h"], a=setup["star_a"], **extra)
    tree.IntegratorBegin(compute_forces=False)
    b0 = tree.IntegratorBlockState(arrays=True)
    if stars:
        st0 = tree.IntegratorGetStars()
        assert np.array_equal(st0["level"], setup["star_level"]) and np.array_equal(st0["nstep"], setup["star_nstep"])
    tree.Download(ps, only=("level", "levelneib"))
    for k in ("n", "nresync", "level_max", "level_step"):
        assert int(b0[k]) == int(sim0[k]), k
    assert abs(b0["dt_max"] - sim0["dt_max"]) <= 1e-12*sim0["dt_max"]
    assert np.array_equal(ps["level"], setup["level"]) and np.array_equal(b0["nstep"], setup["nstep"])
    assert np.abs(b0["dt"] - setup["dt"]).max() <= 1e-12*np.abs(setup["dt"]).max()
    steps = int(d["loop.steps"])
    tree.MainLoop(steps)
    t, ts, n = tree.IntegratorState()
    assert n == steps
    b1 = tree.IntegratorBlockState(arrays=True)
    tree.Download(ps)
    state = {k: ps[k] for k in ("r", "v", "a", "h", "rho", "u", "alpha", "dudt", "level", "levelneib")}
    state.update(nstep=b1["nstep"], nlast=b1["nlast"], dt=b1["dt"], tlast=b1["tlast"])
    sim = dict(b1, timestep=ts, t=t)
    # integer state exact; physical state to the north-star bar after 20 steps (per-particle relative errors of
    # vector components that pass through zero in the shock tube reach 7e-8; scaled by the field maximum: 1e-11)
    check_block_state(state, sim, loop, sim1, 1e-6)
    if stars:
        st = tree.IntegratorGetStars()
        for k in ("level", "nstep", "nlast"):
            assert np.array_equal(st[k], loop["star_" + k]), k
        for k in ("r", "v", "a"):
            assert vec_err(st[k], loop["star_" + k].reshape(st[k].shape)).max() < 1e-6, k
    for f in ("r", "v", "a", "h", "rho", "u"):
        tol = 1e-7 if hermite else 1e-9          # (Hermite-4: the reference's gas-on-star jerk is single precision)
        assert np.abs(np.asarray(state[f]).reshape(-1) - loop[f].reshape(-1)).max() <= tol*np.abs(loop[f]).max(), f
    # the start-of-step state of every particle (most of them in the middle of their own step) is the reference's
    s0 = tree.IntegratorStartState()
    energy = params.get("gas_eos", "energy_eqn") == "energy_eqn"       # (SphLeapfrogKDK touches u0, dudt0 only then)
    for f in ("r0", "v0", "a0", "dt_next") + (("u0", "dudt0") if energy else ()):
        if f in loop:
            assert np.abs(s0[f].reshape(-1) - loop[f].reshape(-1)).max() <= 1e-9*max(np.abs(loop[f]).max(), 1e-300), f
    tree.close()


def test_device_block_timesteps_step_by_step_and_one_level():
    """MainLoop(1) x N == MainLoop(N) bit for bit with block time steps, and Nlevels = 1 still takes the global-step
    path (levels all 0)"""
    import gandalf_b200 as g
    params, d = load_golden("loopblk_adsod")
    setup, sim0, loop, sim1 = split_blk(d)
    res = []
    for chunks in ((12,), (5, 1, 6)):
        ps, tree = _upload_setup(g, params, setup)
        tree.IntegratorBegin(compute_forces=False)
        for c in chunks:
            tree.MainLoop(c)
            tree.Download(ps)
        res.append(({k: ps[k].copy() for k in ("r", "v", "a", "h", "u", "level")}, tree.IntegratorState(),
                    tree.IntegratorBlockState()))
        tree.close()
    assert res[0][1] == res[1][1] and res[0][2] == res[1][2]
    for k in res[0][0]:
        assert np.array_equal(res[0][0][k], res[1][0][k]), k
    ps, tree = _upload_setup(g, dict(params, Nlevels=1), setup)
    tree.IntegratorBegin(compute_forces=False)
    tree.MainLoop(2)
    assert tree.IntegratorBlockState()["level_max"] == 0
    tree.close()
