#!/usr/bin/env python
"""Golden fixtures for the device-resident main loop with hierarchical BLOCK time steps (Nlevels > 1):
tests/golden/loopblk_*.npz.

Each one holds the particle state of the UNMODIFIED reference right after its own setup and after `steps` calls of
its own SphSimulation::MainLoop (lfkdk, Nlevels = 4), sorted by iorig, incl. the integrator state of every particle
(level, levelneib, nstep, nlast, dt, dt_next, tlast, r0, v0, a0, u0, dudt0) and the scalars of the block structure
(n, nresync, level_max, level_step, dt_max, timestep, t) at both points.  The step counts are chosen so that levels
are added, removed and the hierarchy is resynchronised at least once.

    python tests/golden/make_loopblk_golden.py [name ...]      (needs oracle/_ref/ref_harness_std)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import refio  # noqa: E402
from conftest import load_golden, stage  # noqa: E402
from make_loop_golden import BASE  # noqa: E402

CASES = {"adsod": (20, 4), "box3d_mm97": (20, 4), "sphere_grav_fastmono": (20, 4), "adsod+cd2010": (20, 4),
         "sphere_grav_stars+hermite4": (20, 4), "sphere_grav_stars+lfkdk": (20, 4),   # star particles on their own levels
         # sph_single_timestep = 1: every gas particle on the deepest gas level, the stars below it
         "sphere_grav_stars+hermite4+single": (20, 4), "adsod+single": (12, 4),
         "box3d_mirror": (20, 4)}                              # three mirror walls: images raise levelneib of their originals       # name -> (steps, Nlevels)
FIELDS = ["r", "v", "a", "atree", "m", "h", "u", "rho", "invomega", "zeta", "hfactor", "hrangesqd", "sound", "pressure",
          "alpha", "dalphadt", "dudt", "div_v", "gpot", "ptype", "level", "levelneib", "nstep", "nlast", "dt", "dt_next",
          "tlast", "r0", "v0", "a0", "u0", "dudt0"]
SIM = ["n", "nresync", "level_max", "level_step", "dt_max", "timestep", "t"]


def main():
    for name in (sys.argv[1:] or list(CASES)):
        steps, nlev = CASES[name]
        params, d = load_golden(name.split("+")[0])
        p = dict(BASE)
        p.update(params)
        opts = name.split("+")[1:]
        for o in opts:
            if o in ("hermite4", "lfkdk"):                   # the N-body integrator of the three stars of that golden
                p["nbody"] = o
            if o == "single":
                p["sph_single_timestep"] = 1
        if "cd2010" in opts:                         # the Cullen & Dehnen switch on the same initial conditions
            p.update(time_dependent_avisc="cd2010", alpha_visc_min=0.1)
        p.update({k: BASE[k] for k in ("run_id", "ic", "Nstepsmax", "tend", "neib_search")})
        p["Nlevels"] = nlev
        s = stage(d, "dens_in")
        order = np.argsort(s["iorig"])
        ic = {k: s[k][order] for k in ("r", "v", "m", "h", "u")}
        p.setdefault("Nhydromax", 8*len(ic["m"]))
        out, info = refio.run_mainloop(p, ic, steps, gpu=False, kind="std", dumps="setup", stars=d.get("stars"))
        assert info["steps"] == steps
        fix = {"param." + k: np.array(str(v)) for k, v in p.items()}
        for st in ("setup", "loop"):
            o = np.argsort(out[st + ".iorig"])
            for f in FIELDS:
                fix["%s.%s" % (st, f)] = out["%s.%s" % (st, f)][o]
            for k in out:
                if k.startswith(st + ".star_"):
                    fix[k] = out[k]
            src = info["setup"] if st == "setup" else info
            for k in SIM:
                fix["%s.sim_%s" % (st, k)] = np.float64(src[k]) if k in ("dt_max", "timestep", "t") else np.int64(src[k])
        fix["loop.steps"] = np.int64(steps)
        path = os.path.join(HERE, "loopblk_" + name.replace("+", "_") + ".npz")
        np.savez_compressed(path, **fix)
        print("%-32s N=%-6d steps=%d n=%d nresync=%d level_max=%d  %7.1f kB" % (
            "loopblk_" + name, len(ic["m"]), steps, info["n"], info["nresync"], info["level_max"], os.path.getsize(path)/1e3))


if __name__ == "__main__":
    main()
