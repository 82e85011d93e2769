"""Helpers to drive the reference harness (oracle/_ref/ref_harness_*) -- TEST INFRASTRUCTURE ONLY.

The harness binary is the UNMODIFIED reference compiled from /root/reference by oracle/Makefile
plus our dump driver oracle/ref_harness.cpp.  It exists only where it was prebuilt
(this container; it travels to the GPU box inside oracle/_ref/).  Nothing here reads
/root/reference at run time.
"""
import glob
import json
import os
import shutil
import struct
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def harness_path(kind="std"):
    p = os.path.join(REF_DIR, "ref_harness_%s" % kind)
    return p if os.path.exists(p) else None


def write_dat(path, params):
    """Write a GANDALF parameter file ('comment : key = value', Parameters.cpp:126-158)."""
    with open(path, "w") as f:
        for k, v in params.items():
            f.write("%s : %s = %s\n" % (k, k, v))


def write_ic(path, ndim, r, v, m, h, u):
    """IC.bin format read by oracle/ref_harness.cpp."""
    N = len(m)
    r = np.asarray(r, dtype=np.float64).reshape(N, ndim)
    v = np.asarray(v, dtype=np.float64).reshape(N, ndim)
    with open(path, "wb") as f:
        f.write(b"GSPHIC1\0")
        f.write(struct.pack("<iiq", ndim, 0, N))
        for k in range(ndim):
            f.write(np.ascontiguousarray(r[:, k]).tobytes())
        for k in range(ndim):
            f.write(np.ascontiguousarray(v[:, k]).tobytes())
        for a in (m, h, u):
            f.write(np.ascontiguousarray(a, dtype=np.float64).tobytes())


def load_dump(outdir):
    out = {}
    for f in glob.glob(os.path.join(outdir, "*.npy")):
        out[os.path.basename(f)[:-4]] = np.load(f)
    return out


def run_harness(params, ic=None, dumps="setup,density,forces,neib", perturb_h=0.0, kind="std",
                bench_steps=0, threads=None, keep=None, datfile=None, extra_sets=None, timeout=3600,
                stars=None):
    """Run the reference harness; returns (dict of dumped arrays, bench json or None).

    params : dict of .dat parameters (ignored if datfile is given)
    ic     : None (use the .dat's own ic) or dict r,v,m,h,u (h may be 0 => reference guesses h)
    """
    exe = harness_path(kind)
    if exe is None:
        raise FileNotFoundError("oracle/_ref/ref_harness_%s not built (make -C oracle ref)" % kind)
    work = keep or tempfile.mkdtemp(prefix="gsph_ref_")
    os.makedirs(work, exist_ok=True)
    try:
        dat = datfile
        if dat is None:
            dat = os.path.join(work, "params.dat")
            write_dat(dat, params)
        cmd = [exe, dat, os.path.join(work, "out"), "--dump", dumps, "--quiet"]
        if ic is not None:
            ndim = int(params["ndim"]) if params else ic["r"].shape[1]
            icf = os.path.join(work, "ic.bin")
            write_ic(icf, ndim, ic["r"], ic["v"], ic["m"], ic["h"], ic["u"])
            cmd += ["--in", icf]
        if stars is not None:            # (Nstar, ndim + 2) rows {r, m, h}: star particles at rest
            sf = os.path.join(work, "stars.bin")
            with open(sf, "wb") as fh:
                fh.write(np.int64(len(stars)).tobytes())
                fh.write(np.ascontiguousarray(np.asarray(stars, dtype=np.float64).T).tobytes())
            cmd += ["--stars", sf]
        if perturb_h:
            cmd += ["--perturb-h", repr(float(perturb_h))]
        if bench_steps:
            cmd += ["--bench", str(int(bench_steps))]
        for k, v in (extra_sets or {}).items():
            cmd += ["--set", "%s=%s" % (k, v)]
        env = dict(os.environ)
        if threads:
            env["OMP_NUM_THREADS"] = str(threads)
            env["OMP_PROC_BIND"] = "close"
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env, timeout=timeout)
        if res.returncode != 0:
            raise RuntimeError("ref_harness failed (%d): %s" % (res.returncode, res.stderr.decode()[-2000:]))
        bench = None
        for line in res.stdout.decode().splitlines():
            if line.startswith("{\"ref_bench\""):
                bench = json.loads(line)
        return load_dump(os.path.join(work, "out")), bench
    finally:
        if keep is None:
            shutil.rmtree(work, ignore_errors=True)


def csr_rows(off, idx):
    return [idx[off[i]:off[i + 1]] for i in range(len(off) - 1)]


def run_mainloop(params, ic, steps, gpu, timeout=1800, kind="gpu", dumps="", resident=0, stars=None):
    """N steps of the reference's own SphSimulation::MainLoop through oracle/_ref/ref_harness_gpu;
    gpu=True swaps sim->sphneib for gandalf_b200/host/GpuSphTree.h first.  -> (arrays, info json)"""
    exe = harness_path(kind)
    if exe is None:
        raise FileNotFoundError("oracle/_ref/ref_harness_%s not built (make -C oracle ref_%s)" % (kind, kind))
    work = tempfile.mkdtemp(prefix="gsph_loop_")
    try:
        dat = os.path.join(work, "params.dat")
        write_dat(dat, params)
        icf = os.path.join(work, "ic.bin")
        write_ic(icf, int(params["ndim"]), ic["r"], ic["v"], ic["m"], ic["h"], ic["u"])
        cmd = [exe, dat, os.path.join(work, "out"), "--dump", dumps, "--quiet", "--in", icf, "--mainloop", str(int(steps))]
        if gpu:
            cmd.append("--gpu")
        if resident:
            cmd += ["--resident", str(int(resident))]      # the first `resident` steps in the device main loop
        if stars is not None:
            sf = os.path.join(work, "stars.bin")
            with open(sf, "wb") as fh:
                fh.write(np.int64(len(stars)).tobytes())
                fh.write(np.ascontiguousarray(np.asarray(stars, dtype=np.float64).T).tobytes())
            cmd += ["--stars", sf]
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=timeout)
        if res.returncode != 0:
            raise RuntimeError("ref_harness_gpu failed (%d): %s %s" % (res.returncode, res.stdout.decode()[-1000:],
                                                                     res.stderr.decode()[-2000:]))
        info = None
        setup_info = None
        for line in res.stdout.decode().splitlines():
            if line.startswith("{\"ref_mainloop\""):
                info = json.loads(line)
            if line.startswith("{\"ref_setup\""):
                setup_info = json.loads(line)
        if info is not None and setup_info is not None:
            info["setup"] = setup_info                     # n, nresync, level_max, ... right after the reference's setup
        return load_dump(os.path.join(work, "out")), info
    finally:
        shutil.rmtree(work, ignore_errors=True)
