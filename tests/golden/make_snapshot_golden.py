#!/usr/bin/env python
"""Golden "su" snapshot files written by the UNMODIFIED reference's own writer
(Simulation::WriteSerenUnformSnapshotFile) right after its setup, next to the particle state dumped
at full precision (tests/golden/snap_<case>.su + snap_<case>.npz).  They pin gandalf_b200/snapshot.py.

    python tests/golden/make_snapshot_golden.py        (needs oracle/_ref/ref_harness_std)
"""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import refio  # noqa: E402
from conftest import load_golden, stage  # noqa: E402
from make_golden import STARS  # noqa: E402

BASE = {"run_id": "SNAP", "sim": "gradhsph", "ic": "python", "dimensionless": 1, "Nlevels": 1, "Nstepsmax": 1,
        "tend": 1.0e-10, "dt_snap": 1.0e10, "noutputstep": 1000000, "ndiagstep": 1000000, "out_file_form": "su",
        "sink_particles": 0, "sph_integration": "lfkdk", "neib_search": "kdtree"}
CASES = {"adsod": None, "khi2d_quintic": None, "sphere_grav_stars": STARS}


def main():
    for name, stars in CASES.items():
        params, d = load_golden(name)
        p = dict(BASE); p.update(params)
        p.update({k: BASE[k] for k in ("run_id", "ic", "Nstepsmax", "tend", "neib_search")})
        s = stage(d, "dens_in"); order = np.argsort(s["iorig"])
        ic = {k: s[k][order] for k in ("r", "v", "m", "h", "u")}
        p.setdefault("Nhydromax", 8*len(ic["m"]))
        work = tempfile.mkdtemp(prefix="gsph_snap_")
        try:
            dat = os.path.join(work, "params.dat"); refio.write_dat(dat, p)
            icf = os.path.join(work, "ic.bin")
            refio.write_ic(icf, int(p["ndim"]), ic["r"], ic["v"], ic["m"], ic["h"], ic["u"])
            snap = os.path.join(HERE, "snap_%s.su" % name)
            cmd = [refio.harness_path("std"), dat, os.path.join(work, "out"), "--dump", "setup", "--quiet", "--in", icf,
                   "--snapshot", snap]
            if name == "sphere_grav_stars" or name == "adsod":
                cmd += ["--snapshot-column", os.path.join(HERE, "snap_%s.column" % name)]
                cmd += ["--snapshot-sf", os.path.join(HERE, "snap_%s.sf" % name)]
            if stars is not None:
                sf = os.path.join(work, "stars.bin")
                with open(sf, "wb") as fh:
                    fh.write(np.int64(len(stars)).tobytes())
                    fh.write(np.ascontiguousarray(np.asarray(stars, dtype=np.float64).T).tobytes())
                cmd += ["--stars", sf]
            subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            out = refio.load_dump(os.path.join(work, "out"))
            keep = {k[6:]: out[k] for k in out if k.startswith("setup.") and k[6:] in
                    ("r", "v", "m", "h", "rho", "u", "iorig", "ptype")}
            keep["ndim"] = np.int32(p["ndim"])
            if stars is not None:
                keep["stars"] = stars
            np.savez_compressed(os.path.join(HERE, "snap_%s.npz" % name), **keep)
            print("snap_%s: N=%d  %.1f kB" % (name, len(keep["m"]), os.path.getsize(snap)/1e3))
        finally:
            shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
