"""gandalf_b200/snapshot.py (SEREN unformatted "su" snapshots) against the reference's own
writer and reader: tests/golden/snap_*.su were written by Simulation::WriteSerenUnformSnapshotFile of
the UNMODIFIED reference, snap_*.npz hold the same particles dumped at full precision."""
import os

import numpy as np
import pytest

import refio
from conftest import GOLDEN_DIR, SNAP_CASES
from gandalf_b200 import snapshot


def _load(name):
    z = np.load(os.path.join(GOLDEN_DIR, "snap_%s.npz" % name))
    return {k: z[k] for k in z.files}, os.path.join(GOLDEN_DIR, "snap_%s.su" % name)


def test_snapshot_fixtures_exist():
    assert len(SNAP_CASES) >= 3


@pytest.mark.parametrize("name", SNAP_CASES)
def test_read_the_reference_snapshot(name):
    d, path = _load(name)
    s = snapshot.read_su(path)
    nd = int(d["ndim"])
    assert s["ndim"] == nd and s["N"] == len(d["m"]) and s["t"] == 0.0 and s["units"] == []
    assert np.array_equal(s["iorig"], d["iorig"]) and np.array_equal(s["ptype"], d["ptype"])
    for f in ("r", "v"):
        assert np.array_equal(s[f].reshape(len(d["m"]), nd), d[f].reshape(len(d["m"]), nd)), f
    for f in ("m", "h", "rho", "u"):
        assert np.array_equal(s[f], d[f]), f
    if "stars" in d:
        st = d["stars"]
        assert s["Nstar"] == len(st)
        assert np.array_equal(s["stars"]["r"], st[:, :nd]) and np.array_equal(s["stars"]["m"], st[:, nd])
        assert np.array_equal(s["stars"]["h"], st[:, nd + 1])
    else:
        assert s["Nstar"] == 0


@pytest.mark.parametrize("name", SNAP_CASES)
def test_written_file_is_byte_identical_to_the_reference(name, tmp_path):
    d, path = _load(name)
    hdr = snapshot.read_su(path)
    out = str(tmp_path / "mine.su")
    snapshot.write_su(out, d["r"], d["v"], d["m"], d["h"], d["rho"], d["u"], iorig=d["iorig"], ptype=d["ptype"],
                      t=hdr["t"], Nsteps=hdr["Nsteps"], Noutsnap=hdr["Noutsnap"], h_fac=hdr["h_fac"],
                      tsnaplast=hdr["tsnaplast"], mmean=hdr["mmean"], stars=hdr.get("stars"))
    assert open(out, "rb").read() == open(path, "rb").read()


def test_types_are_grouped_like_the_reference(tmp_path):
    """particles of several types are written grouped icm, gas, cdm, dust, each group in array order
    (WriteSerenUnformArrayScalar, SimulationIO.hpp:1962-1981)"""
    rng = np.random.default_rng(5)
    N = 50
    ptype = rng.choice([0, 1, 2, 3], size=N)
    r = rng.normal(size=(N, 3)); v = rng.normal(size=(N, 3))
    m, h, rho, u = (rng.uniform(1, 2, N) for _ in range(4))
    out = str(tmp_path / "t.su")
    snapshot.write_su(out, r, v, m, h, rho, u, ptype=ptype, t=1.5, Nsteps=7)
    s = snapshot.read_su(out)
    order = np.concatenate([np.nonzero(ptype == pt)[0] for pt in (1, 0, 2, 3)])
    assert np.array_equal(s["iorig"], order) and np.array_equal(s["ptype"], ptype[order])
    assert np.array_equal(s["r"], r[order]) and np.array_equal(s["u"], u[order])
    assert s["t"] == 1.5 and s["Nsteps"] == 7
    with pytest.raises(ValueError):
        snapshot.write_su(out, r, v, m, h, rho, u, ptype=np.full(N, 4))      # boundary_type is not a file type
    open(out, "r+b").write(b"NOTASNAPSHOT")
    with pytest.raises(ValueError):
        snapshot.read_su(out)


@pytest.mark.parametrize("name", ["adsod", "sphere_grav_stars"])
def test_reference_reads_our_snapshot(name, tmp_path):
    """the other direction: a file written by snapshot.write_su is the initial condition (ic = file,
    in_file_form = su) of the unmodified reference; its particles after setup carry our r, v, m, u, iorig"""
    if refio.harness_path("std") is None:
        pytest.skip("oracle/_ref/ref_harness_std not built")
    d, path = _load(name)
    nd = int(d["ndim"])
    mine = str(tmp_path / "ic.su")
    st = None
    if "stars" in d:
        st = {"r": d["stars"][:, :nd], "m": d["stars"][:, nd], "h": d["stars"][:, nd + 1]}
    snapshot.write_su(mine, d["r"], d["v"], d["m"], d["h"], d["rho"], d["u"], iorig=d["iorig"], ptype=d["ptype"], stars=st)
    params = {"run_id": "SNAPIN", "sim": "gradhsph", "ic": "file", "in_file": mine, "in_file_form": "su", "ndim": nd,
              "dimensionless": 1, "Nlevels": 1, "Nstepsmax": 1, "tend": 1.0e-10, "dt_snap": 1.0e10,
              "noutputstep": 1000000, "ndiagstep": 1000000, "out_file_form": "su", "sink_particles": 0,
              "sph_integration": "lfkdk", "neib_search": "kdtree", "Nhydromax": 8*len(d["m"]),
              "self_gravity": 1 if "stars" in d else 0, "boundary_lhs[0]": "open", "boundary_rhs[0]": "open"}
    out, _ = refio.run_harness(params, ic=None, dumps="setup")
    o = np.argsort(out["setup.iorig"])
    assert np.array_equal(out["setup.iorig"][o], np.sort(d["iorig"]))
    mo = np.argsort(d["iorig"])
    N = len(d["m"])
    for f in ("r", "v"):
        assert np.array_equal(out["setup." + f][o].reshape(N, nd), d[f].reshape(N, nd)[mo]), f
    for f in ("m", "u"):
        assert np.array_equal(out["setup." + f][o], d[f][mo]), f


@pytest.mark.parametrize("name", ["adsod", "sphere_grav_stars"])
def test_column_format_matches_the_reference(name, tmp_path):
    """the text format: our file from the same particles is byte-identical to the one the reference's
    WriteColumnSnapshotFile wrote, and reading it returns the particles to the 6 digits it carries"""
    d, _ = _load(name)
    ref = os.path.join(GOLDEN_DIR, "snap_%s.column" % name)
    nd = int(d["ndim"])
    N = len(d["m"])
    st = None
    if "stars" in d:
        st = {"r": d["stars"][:, :nd], "m": d["stars"][:, nd], "h": d["stars"][:, nd + 1]}
    out = str(tmp_path / "mine.column")
    snapshot.write_column(out, d["r"].reshape(N, nd), d["v"].reshape(N, nd), d["m"], d["h"], d["rho"], d["u"], stars=st)
    assert open(out).read() == open(ref).read()
    s = snapshot.read_column(ref)
    assert s["N"] == N and s["ndim"] == nd and s["Nstar"] == (0 if st is None else len(st["m"]))
    assert np.allclose(s["r"], d["r"].reshape(N, nd), rtol=1e-5, atol=1e-12)
    assert np.allclose(s["rho"], d["rho"], rtol=1e-5) and np.allclose(s["u"], d["u"], rtol=1e-5)


@pytest.mark.parametrize("name", ["adsod", "sphere_grav_stars"])
def test_formatted_sf_matches_the_reference(name, tmp_path):
    """the formatted text twin of su: byte-identical to the reference's WriteSerenFormSnapshotFile, and
    read back to the 11 digits it carries"""
    d, _ = _load(name)
    ref = os.path.join(GOLDEN_DIR, "snap_%s.sf" % name)
    s = snapshot.read_sf(ref)
    nd, N = int(d["ndim"]), len(d["m"])
    assert s["N"] == N and s["ndim"] == nd and np.array_equal(s["iorig"], d["iorig"])
    assert np.allclose(s["r"].reshape(N, nd), d["r"].reshape(N, nd), rtol=1e-10, atol=1e-14)
    assert np.allclose(s["u"], d["u"], rtol=1e-10) and np.allclose(s["rho"], d["rho"], rtol=1e-10)
    out = str(tmp_path / "mine.sf")
    snapshot.write_sf(out, d["r"].reshape(N, nd), d["v"].reshape(N, nd), d["m"], d["h"], d["rho"], d["u"],
                      iorig=d["iorig"], ptype=d["ptype"], t=s["t"], Nsteps=s["Nsteps"], Noutsnap=s["Noutsnap"],
                      h_fac=s["h_fac"], tsnaplast=s["tsnaplast"], mmean=snapshot.read_su(os.path.join(
                          GOLDEN_DIR, "snap_%s.su" % name))["mmean"], stars=s.get("stars"))
    assert open(out).read() == open(ref).read()


def test_round_trips_random_particle_sets(tmp_path):
    """write -> read is the identity for su (bit for bit), and exact to the digits carried for sf / column,
    for ragged sizes, every dimension, mixed particle types, with and without stars, including N = 0"""
    rng = np.random.default_rng(11)
    for trial in range(12):
        nd = 1 + trial % 3
        N = [0, 1, 7, 100, 333][trial % 5]
        r = rng.normal(size=(N, nd)); v = rng.normal(size=(N, nd))
        m, h, rho, u = (rng.uniform(0.1, 2.0, N) for _ in range(4))
        ptype = rng.choice([0, 1, 2, 3], size=N) if trial % 2 else np.zeros(N, dtype=np.int32)
        iorig = rng.permutation(N).astype(np.int32)
        ns = trial % 3
        stars = None if ns == 0 else {"r": rng.normal(size=(ns, nd)), "v": rng.normal(size=(ns, nd)),
                                      "m": rng.uniform(1, 2, ns), "h": rng.uniform(0.1, 0.2, ns)}
        order = np.concatenate([np.nonzero(ptype == pt)[0] for pt in (1, 0, 2, 3)]).astype(int)
        p = str(tmp_path / ("t%d.su" % trial))
        snapshot.write_su(p, r, v, m, h, rho, u, iorig=iorig, ptype=ptype, t=0.25*trial, Nsteps=trial, stars=stars)
        s = snapshot.read_su(p)
        assert s["N"] == N and s["ndim"] == nd and s["Nstar"] == ns and s["t"] == 0.25*trial and s["Nsteps"] == trial
        if N:
            assert np.array_equal(s["iorig"], iorig[order]) and np.array_equal(s["ptype"], ptype[order])
            assert np.array_equal(s["r"].reshape(N, nd), r[order]) and np.array_equal(s["v"].reshape(N, nd), v[order])
            for k, a in (("m", m), ("h", h), ("rho", rho), ("u", u)):
                assert np.array_equal(s[k], a[order]), k
        if ns:
            assert np.array_equal(s["stars"]["r"], stars["r"]) and np.array_equal(s["stars"]["v"], stars["v"])
            assert np.array_equal(s["stars"]["m"], stars["m"]) and np.array_equal(s["stars"]["h"], stars["h"])
        p2 = str(tmp_path / ("t%d.sf" % trial))
        snapshot.write_sf(p2, r, v, m, h, rho, u, iorig=iorig, ptype=ptype, t=0.25*trial, Nsteps=trial, stars=stars)
        s2 = snapshot.read_sf(p2)
        assert s2["N"] == N and s2["Nstar"] == ns and s2["Nsteps"] == trial
        if N:
            assert np.array_equal(s2["iorig"], iorig[order])
            assert np.allclose(s2["r"].reshape(N, nd), r[order], rtol=1e-10, atol=1e-12)
            assert np.allclose(s2["rho"], rho[order], rtol=1e-10)
        if N and (ptype == 0).all():
            p3 = str(tmp_path / ("t%d.column" % trial))
            snapshot.write_column(p3, r, v, m, h, rho, u, t=0.25*trial, stars=stars)
            s3 = snapshot.read_column(p3)
            assert s3["N"] == N and s3["Nstar"] == ns
            assert np.allclose(s3["r"], r, rtol=1e-5, atol=1e-7) and np.allclose(s3["u"], u, rtol=1e-5)
