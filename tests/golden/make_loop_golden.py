#!/usr/bin/env python
"""Golden fixtures for the device-resident main loop (gsph_integrator_*): tests/golden/loop_*.npz.

Each one holds the particle state of the UNMODIFIED reference right after its own setup
(SphSimulation::PostInitialConditionsSetup) and after `steps` calls of its own
SphSimulation::MainLoop (lfkdk, Nlevels = 1), both sorted by iorig, plus the simulation time and the
next time step.  Initial conditions are those of the operator goldens (dens_in.* of <name>.npz).

    python tests/golden/make_loop_golden.py [name ...]      (needs oracle/_ref/ref_harness_std)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import refio  # noqa: E402
from conftest import load_golden, stage  # noqa: E402

BASE = {"run_id": "LOOP", "sim": "gradhsph", "ic": "python", "dimensionless": 1, "Nlevels": 1,
        "Nstepsmax": 1000, "tend": 1.0e10, "dt_snap": 1.0e10, "noutputstep": 1000000, "ndiagstep": 1000000,
        "out_file_form": "su", "sink_particles": 0, "sph_integration": "lfkdk", "neib_search": "kdtree"}

# name -> (steps, extra parameters)
CASES = {
    "adsod": (6, {}),
    "khi2d_quintic": (4, {}),
    "box3d_mm97": (4, {}),                                   # alpha is integrated too (mm97)
    # mirror walls.  (Not rti2d_mirror: with a periodic AND a mirror dimension the reference's force pass reads
    # an uninitialised dr_corr component in GhostNeighbourFinder::PeriodicDistanceCorrection; inside its
    # MainLoop the stack garbage is not benign and four corner particles get arbitrary accelerations.)
    "box3d_mirror": (4, {}),
    "sphere_grav_fastmono": (4, {}),                         # barotropic: u follows from the EOS
    "sphere_grav_quad": (5, {"ntreebuildstep": 3, "ntreestockstep": 2}),   # rebuild / extrapolate / restock mix
    "sphere_grav_stars": (4, {}),                            # three star particles moved by the N-body leapfrog
    # ... and by the reference's DEFAULT N-body integrator, the 4th-order Hermite scheme (NbodyHermite4.cpp)
    "sphere_grav_stars+hermite4": (4, {"nbody": "hermite4"}),
}

# setup: what a restart needs (the first step recomputes every density-pass field); loop: what is compared
FIELDS = {"setup": ["r", "v", "a", "atree", "m", "h", "u", "sound", "dudt", "div_v", "alpha", "dalphadt", "gpot",
                    "ptype", "level"],
          "loop": ["r", "v", "a", "h", "u", "rho", "dudt", "div_v", "alpha", "gpot"]}


def main():
    for name in (sys.argv[1:] or list(CASES)):
        steps, extra = CASES[name]
        params, d = load_golden(name.split("+")[0])
        p = dict(BASE)
        p.update(params)
        p.update({k: BASE[k] for k in ("run_id", "ic", "Nstepsmax", "tend", "neib_search")})
        p.update(extra)
        s = stage(d, "dens_in")
        order = np.argsort(s["iorig"])
        ic = {k: s[k][order] for k in ("r", "v", "m", "h", "u")}
        p.setdefault("Nhydromax", 8*len(ic["m"]))
        out, info = refio.run_mainloop(p, ic, steps, gpu=False, kind="std", dumps="setup", stars=d.get("stars"))
        assert info["steps"] == steps
        fix = {"param." + k: np.array(str(v)) for k, v in p.items()}
        for st in ("setup", "loop"):
            o = np.argsort(out[st + ".iorig"])
            for f in FIELDS[st]:
                fix["%s.%s" % (st, f)] = out["%s.%s" % (st, f)][o]
            for k in out:
                if k.startswith(st + ".star_"):
                    fix[k] = out[k]
        fix["loop.t"] = np.float64(info["t"])
        fix["loop.timestep"] = np.float64(info["timestep"])
        fix["loop.steps"] = np.int64(steps)
        path = os.path.join(HERE, "loop_" + name.replace("+", "_") + ".npz")
        np.savez_compressed(path, **fix)
        print("%-28s N=%-6d steps=%d t=%.6e  %7.1f kB" % ("loop_" + name, len(ic["m"]), steps, info["t"],
                                                         os.path.getsize(path)/1e3))


if __name__ == "__main__":
    main()
