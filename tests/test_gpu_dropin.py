"""Drop-in check (parity stage P6): the UNMODIFIED reference's own main loop
(SphSimulation::MainLoop, src/Hydrodynamics/SphSimulation.cpp:574-880) is run twice from the same
initial conditions -- once with its CPU GradhSphTree, once with sim->sphneib replaced by our
GpuSphTree subclass (gandalf_b200/host/GpuSphTree.h -> libgsph.so) -- and the particle state after
several time steps is compared.  Everything else (integrator, time-step control, ghosts, EOS object,
zeroing of accelerations) is the reference's own code in both runs.

The binary oracle/_ref/ref_harness_gpu is prebuilt in the build container (it needs the reference
headers) and travels to the GPU box; nothing here reads /root/reference."""
import numpy as np
import pytest

import refio
from conftest import load_golden, stage, vec_rel_err, rel_err, scaled_err

pytestmark = pytest.mark.gpu

BASE = {"run_id": "DROPIN", "sim": "gradhsph", "ic": "python", "dimensionless": 1, "Nlevels": 1,
        "Nstepsmax": 1000, "tend": 1.0e10, "dt_snap": 1.0e10, "noutputstep": 1000000, "ndiagstep": 1000000,
        "out_file_form": "su", "sink_particles": 0, "sph_integration": "lfkdk", "neib_search": "kdtree"}


def _case(name):
    params, d = load_golden(name)
    p = dict(BASE)
    p.update(params)
    p.update({k: BASE[k] for k in ("run_id", "ic", "Nstepsmax", "tend", "neib_search")})
    s = stage(d, "dens_in")
    order = np.argsort(s["iorig"])
    ic = {k: s[k][order] for k in ("r", "v", "m", "h", "u")}
    p.setdefault("Nhydromax", 8*len(ic["m"]))
    return p, ic


@pytest.mark.parametrize("name,steps", [("adsod", 6), ("khi2d_quintic", 4), ("box3d_m4tab", 4),
                                        ("box3d_mm97", 3), ("sphere_grav_quad", 4), ("sphere_grav_fastmono", 4)])
def test_reference_mainloop_with_gpu_tree(name, steps):
    if refio.harness_path("gpu") is None:
        pytest.skip("oracle/_ref/ref_harness_gpu not built")
    p, ic = _case(name)
    cpu, ci = refio.run_mainloop(p, ic, steps, gpu=False)
    gpu, gi = refio.run_mainloop(p, ic, steps, gpu=True)
    assert ci["steps"] == gi["steps"] == steps and gi["gpu"] and not ci["gpu"]
    assert abs(ci["t"] - gi["t"]) <= 1e-12*abs(ci["t"])          # same time-step sequence
    oc, og = np.argsort(cpu["loop.iorig"]), np.argsort(gpu["loop.iorig"])
    tol = 1e-6                                                   # north-star bar (double precision)
    for f in ("r", "v", "a"):
        a, b = gpu["loop." + f][og], cpu["loop." + f][oc]
        a = a.reshape(len(a), -1); b = b.reshape(len(b), -1)
        e = np.linalg.norm(a - b, axis=1)/np.maximum(np.linalg.norm(b, axis=1), np.abs(b).max()*1e-6)
        assert e.max() < tol, (f, e.max())
    for f in ("h", "rho", "u"):
        assert rel_err(gpu["loop." + f][og], cpu["loop." + f][oc]).max() < tol, f
    assert scaled_err(gpu["loop.dudt"][og], cpu["loop.dudt"][oc]).max() < tol


@pytest.mark.parametrize("name,steps,resident", [("adsod", 6, 6), ("box3d_mm97", 4, 2), ("sphere_grav_quad", 4, 3),
                                                 ("khi2d_quintic", 4, 4)])
def test_reference_hands_over_to_the_device_main_loop(name, steps, resident):
    """C++ twin, MainLoop level: after the reference's own setup the particles go to the device once
    (GpuSphTree::ResidentBegin), `resident` steps run in the device-resident main loop
    (gsph_integrator_step), the particles come home (ResidentFetch restores r0, v0, a0, u0, dudt0, tlast,
    dt as SphLeapfrogKDK::EndTimestep leaves them) and the reference's OWN MainLoop carries on for the
    remaining steps.  The result must match the all-CPU run: same time-step sequence, state to 1e-6."""
    if refio.harness_path("gpu") is None:
        pytest.skip("oracle/_ref/ref_harness_gpu not built")
    p, ic = _case(name)
    cpu, ci = refio.run_mainloop(p, ic, steps, gpu=False)
    dev, di = refio.run_mainloop(p, ic, steps, gpu=True, resident=resident)
    assert ci["steps"] == di["steps"] == steps
    assert abs(ci["t"] - di["t"]) <= 1e-12*abs(ci["t"])
    assert abs(ci["timestep"] - di["timestep"]) <= 1e-9*abs(ci["timestep"])
    oc, og = np.argsort(cpu["loop.iorig"]), np.argsort(dev["loop.iorig"])
    tol = 1e-6
    for f in ("r", "v", "a"):
        a, b = dev["loop." + f][og], cpu["loop." + f][oc]
        a = a.reshape(len(a), -1); b = b.reshape(len(b), -1)
        e = np.linalg.norm(a - b, axis=1)/np.maximum(np.linalg.norm(b, axis=1), np.abs(b).max()*1e-6)
        assert e.max() < tol, (f, e.max())
    for f in ("h", "rho", "u", "alpha"):
        assert rel_err(dev["loop." + f][og], cpu["loop." + f][oc]).max() < tol, f
    assert scaled_err(dev["loop.dudt"][og], cpu["loop.dudt"][oc]).max() < tol


@pytest.mark.parametrize("name,steps,resident", [("adsod", 14, 9), ("box3d_mm97", 10, 5), ("sphere_grav_fastmono", 10, 7)])
def test_reference_hands_over_with_block_time_steps(name, steps, resident):
    """the same hand-over with hierarchical block time steps (Nlevels = 4): the device loop starts the hierarchy like
    PostInitialConditionsSetup, runs `resident` steps, and ResidentFetch brings home every particle in the middle
    of its own step -- level, nstep, nlast, dt, dt_next, tlast, r0, v0, a0, u0, dudt0 -- plus n, nresync, level_max,
    level_step, dt_max of the hierarchy; the reference's own MainLoop carries on.  Must equal the all-CPU run:
    identical integer state, values to 1e-6."""
    if refio.harness_path("gpu") is None:
        pytest.skip("oracle/_ref/ref_harness_gpu not built")
    p, ic = _case(name)
    p = dict(p, Nlevels=4)
    cpu, ci = refio.run_mainloop(p, ic, steps, gpu=False)
    dev, di = refio.run_mainloop(p, ic, steps, gpu=True, resident=resident)
    assert ci["steps"] == di["steps"] == steps
    for k in ("n", "nresync", "level_max", "level_step"):
        assert ci[k] == di[k], (k, ci[k], di[k])
    assert abs(ci["t"] - di["t"]) <= 1e-12*abs(ci["t"])
    assert abs(ci["dt_max"] - di["dt_max"]) <= 1e-9*abs(ci["dt_max"])
    oc, og = np.argsort(cpu["loop.iorig"]), np.argsort(dev["loop.iorig"])
    assert len(np.unique(cpu["loop.level"])) > 1, "one level only: not a block-step case"
    for f in ("level", "nstep", "nlast"):
        assert np.array_equal(dev["loop." + f][og], cpu["loop." + f][oc]), f
    tol = 1e-6
    for f in ("r", "v", "a", "r0", "v0", "a0"):
        a, b = dev["loop." + f][og], cpu["loop." + f][oc]
        a = a.reshape(len(a), -1); b = b.reshape(len(b), -1)
        e = np.linalg.norm(a - b, axis=1)/np.maximum(np.linalg.norm(b, axis=1), np.abs(b).max()*1e-6)
        assert e.max() < tol, (f, e.max())
    energy = p.get("gas_eos", "energy_eqn") == "energy_eqn"       # (SphLeapfrogKDK touches u0 only then)
    for f in ("h", "rho", "u", "dt", "tlast") + (("u0",) if energy else ()):
        assert rel_err(dev["loop." + f][og], cpu["loop." + f][oc]).max() < tol, f
    assert scaled_err(dev["loop.dudt"][og], cpu["loop.dudt"][oc]).max() < tol


@pytest.mark.parametrize("nbody", ["hermite4", "lfkdk"])
def test_reference_hands_over_with_star_particles(nbody):
    """the hand-over with three star particles: GpuSphTree::ResidentBegin(sph, nbody, ...) ships the stars of the
    reference's own N-body object (r, v, a and -- Hermite-4 -- adot, a2dot, a3dot), the device loop integrates them
    with the reference's integrator for 2 steps, ResidentFetch(sph, nbody, ...) restores them as Nbody*::EndTimestep
    leaves them and the reference's own MainLoop carries on.  Gas and stars == the all-CPU run."""
    if refio.harness_path("gpu") is None:
        pytest.skip("oracle/_ref/ref_harness_gpu not built")
    p, ic = _case("sphere_grav_stars")
    p = dict(p, nbody=nbody)
    _, d = load_golden("sphere_grav_stars")
    cpu, ci = refio.run_mainloop(p, ic, 4, gpu=False, stars=d["stars"])
    dev, di = refio.run_mainloop(p, ic, 4, gpu=True, resident=2, stars=d["stars"])
    assert ci["steps"] == di["steps"] == 4
    # (Hermite-4: the Aarseth step follows the jerk; the reference's gas-on-star jerk is single precision)
    assert abs(ci["t"] - di["t"]) <= (1e-8 if nbody == "hermite4" else 1e-12)*abs(ci["t"])
    oc, og = np.argsort(cpu["loop.iorig"]), np.argsort(dev["loop.iorig"])
    for f in ("r", "v", "a"):
        a, b = dev["loop." + f][og], cpu["loop." + f][oc]
        e = np.linalg.norm(a - b, axis=1)/np.maximum(np.linalg.norm(b, axis=1), np.abs(b).max()*1e-6)
        assert e.max() < 1e-6, (f, e.max())
    for f in ("h", "rho", "u", "gpot"):
        assert rel_err(dev["loop." + f][og], cpu["loop." + f][oc]).max() < 1e-6, f
    for f in ("star_r", "star_v", "star_a"):
        a, b = dev["loop." + f], cpu["loop." + f]
        assert np.abs(a - b).max() <= 1e-6*np.abs(b).max(), f


def test_reference_mainloop_with_stars_and_gpu_tree():
    """three star particles: the gas feels them (ComputeStarGravForces, star terms of zeta) and they feel
    the gas through GpuSphTree::UpdateAllStarGasForces; the reference's own N-body integrator moves them.
    Gas state after 3 steps == the all-CPU run."""
    if refio.harness_path("gpu") is None:
        pytest.skip("oracle/_ref/ref_harness_gpu not built")
    p, ic = _case("sphere_grav_stars")
    _, d = load_golden("sphere_grav_stars")
    cpu, ci = refio.run_mainloop(p, ic, 3, gpu=False, stars=d["stars"])
    gpu, gi = refio.run_mainloop(p, ic, 3, gpu=True, stars=d["stars"])
    assert abs(ci["t"] - gi["t"]) <= 1e-12*abs(ci["t"])
    oc, og = np.argsort(cpu["loop.iorig"]), np.argsort(gpu["loop.iorig"])
    for f in ("r", "v", "a"):
        a, b = gpu["loop." + f][og], cpu["loop." + f][oc]
        e = np.linalg.norm(a - b, axis=1)/np.maximum(np.linalg.norm(b, axis=1), np.abs(b).max()*1e-6)
        assert e.max() < 1e-6, (f, e.max())
    for f in ("h", "rho", "u", "zeta", "gpot"):
        assert rel_err(gpu["loop." + f][og], cpu["loop." + f][oc]).max() < 1e-6, f
