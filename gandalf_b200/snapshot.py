"""SEREN unformatted ("su") snapshots, the reference's binary wire format, straight from / to SoA
arrays: lets 16-64 Mi particle inputs and the state of a device-resident run be produced
and consumed without the host AoS detour.

Layout (src/Common/SimulationIO.hpp:1655-2262 writer, :1290-1650 reader; all little endian):
    char[20]  "SERENBINARYDUMPV3" left-justified, blank padded          (:42-44)
    int32     8                      size of FLOAT (the double-precision build)
    int32 x3  ndim, ndim, ndim
    int32[50] idata     [0] Nhydro  [1] Nstar  [3] N_icm [4] N_gas [5] N_cdm [6] N_dust
                        [19] nunit  [20] ndata
    int64[50] ilpdata   [0] Noutsnap [1] Nsteps [10] Noutlitesnap
    f64[50]   rdata     [0] h_fac
    f64[50]   ddata     [0] t [1] tsnaplast [2] mmean [10] tlitesnaplast
    char[20] x nunit    unit strings (none when dimensionless)
    char[20] x ndata    array ids: porig r m h v rho u [sink_v1]
    int32[5] x ndata    per array: width, first, last (count), type code (2 int, 4 real, 7 sink), unit index
    porig int32[N]; r f64[N][ndim]; m; h; v f64[N][ndim]; rho; u      particles grouped by type in the
                                                                      order icm, gas, cdm, dust (:2143-2166)
    stars: int32[6] {2, 2, 0, 12 + 2 ndim, 0, 0}, then per star int32[4] {1, 1, i + 1, 0} and
           f64[12 + 2 ndim]: [1..ndim] r, [1+ndim..2 ndim] v, [1+2ndim] m, [2+2ndim] h, [3+2ndim] radius
"""
import numpy as np

TAG = b"SERENBINARYDUMPV3"
STRLEN = 20
# ptype codes (src/Headers/Particle.h:43) in the order the file groups them
GAS, ICM, CDM, DUST = 0, 1, 2, 3
FILE_ORDER = (ICM, GAS, CDM, DUST)


def _pad(s):
    b = s if isinstance(s, bytes) else s.encode()
    return b.ljust(STRLEN)[:STRLEN]


def write_su(path, r, v, m, h, rho, u, iorig=None, ptype=None, t=0.0, Nsteps=0, Noutsnap=0, h_fac=1.2,
             tsnaplast=0.0, mmean=None, stars=None, units=None):
    """Write one snapshot.  r, v: (N, ndim); the rest (N,).  stars: dict r (Ns, ndim), v, m, h[, radius].
    units: None (dimensionless) or the 21 unit strings of SimulationIO.hpp:1700-1722 (values are written
    as given: scale them to output units beforehand)."""
    r = np.ascontiguousarray(r, dtype="<f8")
    N, ndim = r.shape
    v = np.ascontiguousarray(v, dtype="<f8").reshape(N, ndim)
    cols = [np.ascontiguousarray(a, dtype="<f8").reshape(N) for a in (m, h, rho, u)]
    iorig = np.arange(N, dtype="<i4") if iorig is None else np.ascontiguousarray(iorig, dtype="<i4")
    ptype = np.zeros(N, dtype=np.int32) if ptype is None else np.asarray(ptype, dtype=np.int32)
    if not np.isin(ptype, FILE_ORDER).all():
        raise ValueError("SerenUnformReader: Type not recognised")
    # stable grouping by type in file order
    order = np.concatenate([np.nonzero(ptype == pt)[0] for pt in FILE_ORDER]) if (ptype != GAS).any() else None
    counts = [int((ptype == pt).sum()) for pt in FILE_ORDER]
    nstar = 0 if stars is None else len(stars["m"])
    ids = ([b"porig", b"r", b"m", b"h", b"v", b"rho", b"u"] if N > 0 else []) + ([b"sink_v1"] if nstar else [])
    typedata = []
    if N > 0:
        typedata += [[1, 1, N, 2, 0], [ndim, 1, N, 4, 1], [1, 1, N, 4, 2], [1, 1, N, 4, 1], [ndim, 1, N, 4, 4],
                     [1, 1, N, 4, 6], [1, 1, N, 4, 20]]
    if nstar:
        typedata += [[1, 1, nstar, 7, 0]]
    idata = np.zeros(50, dtype="<i4")
    idata[0], idata[1] = N, nstar
    idata[3:7] = counts
    idata[19] = 0 if units is None else len(units)
    idata[20] = len(ids)
    ilp = np.zeros(50, dtype="<i8"); ilp[0], ilp[1] = Noutsnap, Nsteps
    rdata = np.zeros(50, dtype="<f8"); rdata[0] = h_fac
    ddata = np.zeros(50, dtype="<f8")
    ddata[0], ddata[1] = t, tsnaplast
    ddata[2] = (float(cols[0].mean()) if N else 0.0) if mmean is None else mmean

    def grouped(a):
        return a if order is None else a[order]

    with open(path, "wb") as f:
        f.write(_pad(TAG))
        np.array([8, ndim, ndim, ndim], dtype="<i4").tofile(f)
        idata.tofile(f); ilp.tofile(f); rdata.tofile(f); ddata.tofile(f)
        for s in (units or []):
            f.write(_pad(s))
        for s in ids:
            f.write(_pad(s))
        np.array(typedata, dtype="<i4").reshape(-1, 5).tofile(f)
        if N > 0:
            grouped(iorig).tofile(f)
            grouped(r).tofile(f)
            grouped(cols[0]).tofile(f)          # m
            grouped(cols[1]).tofile(f)          # h
            grouped(v).tofile(f)
            grouped(cols[2]).tofile(f)          # rho
            grouped(cols[3]).tofile(f)          # u
        if nstar:
            L = 12 + 2*ndim
            np.array([2, 2, 0, L, 0, 0], dtype="<i4").tofile(f)
            sr = np.asarray(stars["r"], dtype="<f8").reshape(nstar, ndim)
            sv = np.asarray(stars.get("v", np.zeros((nstar, ndim))), dtype="<f8").reshape(nstar, ndim)
            for i in range(nstar):
                np.array([1, 1, i + 1, 0], dtype="<i4").tofile(f)
                sd = np.zeros(L, dtype="<f8")
                sd[1:1 + ndim] = sr[i]; sd[1 + ndim:1 + 2*ndim] = sv[i]
                sd[1 + 2*ndim] = stars["m"][i]; sd[2 + 2*ndim] = stars["h"][i]
                sd[3 + 2*ndim] = stars.get("radius", np.zeros(nstar))[i]
                sd.tofile(f)


def read_su(path):
    """-> dict: ndim, N, Nstar, t, Nsteps, Noutsnap, h_fac, tsnaplast, mmean, units, arrays
    iorig r m h v rho u ptype (file order: grouped by type) and stars (r v m h radius) if present."""
    with open(path, "rb") as f:
        tag = f.read(STRLEN)
        if tag.rstrip() != TAG:
            raise ValueError("Incorrect format of snapshot file %s: %r instead of %r" % (path, tag, TAG))
        prec, n1, n2, n3 = np.fromfile(f, dtype="<i4", count=4)
        if prec != 8:
            raise ValueError("Incorrect precision of snapshot file %s: %d (this build reads doubles)" % (path, prec))
        if not (n1 == n2 == n3) or not 1 <= n1 <= 3:
            raise ValueError("Incorrect NDIM in file %s" % path)
        ndim = int(n1)
        idata = np.fromfile(f, dtype="<i4", count=50)
        ilp = np.fromfile(f, dtype="<i8", count=50)
        rdata = np.fromfile(f, dtype="<f8", count=50)
        ddata = np.fromfile(f, dtype="<f8", count=50)
        N, nstar, nunit, ndata = int(idata[0]), int(idata[1]), int(idata[19]), int(idata[20])
        units = [f.read(STRLEN).decode().strip() for _ in range(nunit)]
        ids = [f.read(STRLEN).decode().strip() for _ in range(ndata)]
        typedata = np.fromfile(f, dtype="<i4", count=5*ndata).reshape(ndata, 5)
        out = {"ndim": ndim, "N": N, "Nstar": nstar, "t": float(ddata[0]), "tsnaplast": float(ddata[1]),
               "mmean": float(ddata[2]), "Nsteps": int(ilp[1]), "Noutsnap": int(ilp[0]), "h_fac": float(rdata[0]),
               "units": units, "data_id": ids}
        names = {"porig": "iorig"}
        for name, td in zip(ids, typedata):
            width, first, last, code = int(td[0]), int(td[1]), int(td[2]), int(td[3])
            n = last - first + 1
            if code == 2:
                a = np.fromfile(f, dtype="<i4", count=n*width)
            elif code == 4:
                a = np.fromfile(f, dtype="<f8", count=n*width)
            elif code == 7:
                hdr = np.fromfile(f, dtype="<i4", count=6)
                L = int(hdr[3])
                st = {k: [] for k in ("r", "v", "m", "h", "radius")}
                for _ in range(n):
                    np.fromfile(f, dtype="<i4", count=4)
                    sd = np.fromfile(f, dtype="<f8", count=L)
                    st["r"].append(sd[1:1 + ndim]); st["v"].append(sd[1 + ndim:1 + 2*ndim])
                    st["m"].append(sd[1 + 2*ndim]); st["h"].append(sd[2 + 2*ndim]); st["radius"].append(sd[3 + 2*ndim])
                out["stars"] = {k: np.array(x) for k, x in st.items()}
                continue
            else:
                raise ValueError("unknown array type code %d for %s" % (code, name))
            if len(a) != n*width:
                raise ValueError("snapshot file %s is truncated in array %s" % (path, name))
            out[names.get(name, name)] = a.reshape(n, width) if width > 1 else a
        counts = idata[3:7]
        out["ptype"] = np.concatenate([np.full(int(c), pt, dtype=np.int32) for c, pt in zip(counts, FILE_ORDER)]) \
            if counts.sum() == N else np.zeros(N, dtype=np.int32)
    return out


def particle_set_from_su(path, alloc=None):
    """-> (ParticleSet in SoA layout with r v m h u rho ptype iorig filled, header dict)"""
    from .sphtree import ParticleSet
    d = read_su(path)
    ps = ParticleSet(d["N"], d["ndim"], "soa", alloc=alloc)
    for k in ("r", "v"):
        ps[k][...] = d[k].reshape(d["N"], d["ndim"])
    for k in ("m", "h", "u", "rho", "ptype", "iorig"):
        ps[k][...] = d[k]
    return ps, d


def write_su_from_particle_set(path, ps, **header):
    """snapshot of a ParticleSet (e.g. after GpuSphTree.Download of a device-resident run)"""
    dead = (np.asarray(ps["flags"]) & 1) != 0
    keep = ~dead
    write_su(path, ps["r"][keep], ps["v"][keep], ps["m"][keep], ps["h"][keep], ps["rho"][keep], ps["u"][keep],
             iorig=ps["iorig"][keep], ptype=ps["ptype"][keep], **header)


# ---------------------------------------------------------------------------------------------------
# "column" text snapshots (SimulationIO.hpp:162-545): Nhydro, Nstar, ndim, t on four lines, then one
# row per particle  r[ndim] v[ndim] m h rho u  (three blanks between the numbers, C++ default stream
# formatting = %g with 6 digits), then one row per star  r v m h 0 0.
# ---------------------------------------------------------------------------------------------------
def write_column(path, r, v, m, h, rho, u, t=0.0, stars=None):
    r = np.asarray(r, dtype=np.float64)
    N, ndim = r.shape
    v = np.asarray(v, dtype=np.float64).reshape(N, ndim)
    nstar = 0 if stars is None else len(stars["m"])
    rows = np.column_stack([r, v, m, h, rho, u])
    with open(path, "w") as f:
        f.write("%d\n%d\n%d\n%g\n" % (N, nstar, ndim, t))
        tail = "   \n" if ndim == 1 else "\n"       # the 1-D particle rows of the reference end with a separator (:466-472)
        for row in rows:
            f.write("   ".join("%g" % x for x in row) + tail)
        if nstar:
            sr = np.asarray(stars["r"], dtype=np.float64).reshape(nstar, ndim)
            sv = np.asarray(stars.get("v", np.zeros((nstar, ndim))), dtype=np.float64).reshape(nstar, ndim)
            for i in range(nstar):
                row = list(sr[i]) + list(sv[i]) + [stars["m"][i], stars["h"][i], 0.0, 0.0]
                f.write("   ".join("%g" % x for x in row) + "\n")


def read_column(path):
    """-> dict ndim N Nstar t r v m h rho u [stars]; values carry the 6 significant digits of the file"""
    with open(path) as f:
        N, nstar, ndim = int(f.readline()), int(f.readline()), int(f.readline())
        t = float(f.readline())
        body = np.loadtxt(f, ndmin=2) if N + nstar else np.zeros((0, 2*ndim + 4))
    if body.shape != (N + nstar, 2*ndim + 4):
        raise ValueError("column snapshot %s: expected %d rows of %d numbers" % (path, N + nstar, 2*ndim + 4))
    p = body[:N]
    out = {"ndim": ndim, "N": N, "Nstar": nstar, "t": t, "r": p[:, :ndim], "v": p[:, ndim:2*ndim],
           "m": p[:, 2*ndim], "h": p[:, 2*ndim + 1], "rho": p[:, 2*ndim + 2], "u": p[:, 2*ndim + 3]}
    if nstar:
        s = body[N:]
        out["stars"] = {"r": s[:, :ndim], "v": s[:, ndim:2*ndim], "m": s[:, 2*ndim], "h": s[:, 2*ndim + 1]}
    return out


# ---------------------------------------------------------------------------------------------------
# SEREN formatted ("sf") snapshots: the same records as "su" as text (SimulationIO.hpp:989-1240 writer,
# :600-980 reader): tag line, precision 4 and ndim x3 in width 2, the four 50-word header blocks
# (integers width 10, reals %18.10e), unit and array ids one per line, the type table (5 x width 10 per
# line), then the arrays (one particle per line) and the star records (integers width 8 there).
# ---------------------------------------------------------------------------------------------------
ASCII_TAG = "SERENASCIIDUMPV2"


def write_sf(path, r, v, m, h, rho, u, iorig=None, ptype=None, t=0.0, Nsteps=0, Noutsnap=0, h_fac=1.2,
             tsnaplast=0.0, mmean=None, stars=None, units=None):
    r = np.asarray(r, dtype=np.float64)
    N, ndim = r.shape
    v = np.asarray(v, dtype=np.float64).reshape(N, ndim)
    cols = [np.asarray(a, dtype=np.float64).reshape(N) for a in (m, h, rho, u)]
    iorig = np.arange(N, dtype=np.int32) if iorig is None else np.asarray(iorig, dtype=np.int32)
    ptype = np.zeros(N, dtype=np.int32) if ptype is None else np.asarray(ptype, dtype=np.int32)
    if not np.isin(ptype, FILE_ORDER).all():
        raise ValueError("SerenFormReader: Type not recognised")
    order = np.concatenate([np.nonzero(ptype == pt)[0] for pt in FILE_ORDER])
    nstar = 0 if stars is None else len(stars["m"])
    ids = (["porig", "r", "m", "h", "v", "rho", "u"] if N > 0 else []) + (["sink_v1"] if nstar else [])
    typedata = []
    if N > 0:
        typedata += [[1, 1, N, 2, 0], [ndim, 1, N, 4, 1], [1, 1, N, 4, 2], [1, 1, N, 4, 1], [ndim, 1, N, 4, 4],
                     [1, 1, N, 4, 6], [1, 1, N, 4, 20]]
    if nstar:
        typedata += [[1, 1, nstar, 7, 0]]
    idata = [0]*50
    idata[0], idata[1] = N, nstar
    idata[3:7] = [int((ptype == pt).sum()) for pt in FILE_ORDER]
    idata[19], idata[20] = (0 if units is None else len(units)), len(ids)
    ilp = [0]*50; ilp[0], ilp[1] = Noutsnap, Nsteps
    rdata = [0.0]*50; rdata[0] = h_fac
    ddata = [0.0]*50
    ddata[0], ddata[1] = t, tsnaplast
    ddata[2] = (float(cols[0].mean()) if N else 0.0) if mmean is None else mmean
    fl = "%18.10e"
    with open(path, "w") as f:
        f.write(ASCII_TAG + "\n")
        f.write("%2d\n" % 4 + ("%2d\n" % ndim)*3)
        for x in idata + ilp:
            f.write("%10d\n" % x)
        for x in rdata + ddata:
            f.write(fl % x + "\n")
        for s_ in (units or []):
            f.write(s_ + "\n")
        for s_ in ids:
            f.write(s_ + "\n")
        for td in typedata:
            f.write("".join("%10d" % x for x in td) + "\n")
        if N > 0:
            f.write("".join("%10d\n" % x for x in iorig[order]))
            f.write("".join("".join(fl % x for x in row) + "\n" for row in r[order]))
            f.write("".join(fl % x + "\n" for x in cols[0][order]))
            f.write("".join(fl % x + "\n" for x in cols[1][order]))
            f.write("".join("".join(fl % x for x in row) + "\n" for row in v[order]))
            f.write("".join(fl % x + "\n" for x in cols[2][order]))
            f.write("".join(fl % x + "\n" for x in cols[3][order]))
        if nstar:
            L = 12 + 2*ndim
            f.write("".join("%10d" % x for x in (2, 2, 0, L, 0, 0)) + "\n")
            sr = np.asarray(stars["r"], dtype=np.float64).reshape(nstar, ndim)
            sv = np.asarray(stars.get("v", np.zeros((nstar, ndim))), dtype=np.float64).reshape(nstar, ndim)
            for i in range(nstar):
                f.write("TT\n")
                f.write("%8d%8d\n" % (i + 1, 0))
                sd = np.zeros(L)
                sd[1:1 + ndim] = sr[i]; sd[1 + ndim:1 + 2*ndim] = sv[i]
                sd[1 + 2*ndim] = stars["m"][i]; sd[2 + 2*ndim] = stars["h"][i]
                sd[3 + 2*ndim] = stars.get("radius", np.zeros(nstar))[i]
                f.write("".join(fl % x for x in sd) + "\n")


def read_sf(path):
    """-> the same dict as read_su (values carry the 11 significant digits of the text)"""
    with open(path) as f:
        tok = f.read().split()
    if tok[0] not in (ASCII_TAG, "SERENASCIIDUMPV3"):
        raise ValueError("Incorrect format of snapshot file %s: %s" % (path, tok[0]))
    pos = [1]

    def take(n, conv):
        out = [conv(x) for x in tok[pos[0]:pos[0] + n]]
        if len(out) != n:
            raise ValueError("snapshot file %s is truncated" % path)
        pos[0] += n
        return out
    _, n1, n2, n3 = take(4, int)
    if not (n1 == n2 == n3):
        raise ValueError("Incorrect NDIM in file %s" % path)
    ndim = n1
    idata, ilp = take(50, int), take(50, int)
    rdata, ddata = take(50, float), take(50, float)
    N, nstar, nunit, ndata = idata[0], idata[1], idata[19], idata[20]
    units, ids = take(nunit, str), take(ndata, str)
    typedata = [take(5, int) for _ in range(ndata)]
    out = {"ndim": ndim, "N": N, "Nstar": nstar, "t": ddata[0], "tsnaplast": ddata[1], "mmean": ddata[2],
           "Nsteps": ilp[1], "Noutsnap": ilp[0], "h_fac": rdata[0], "units": units, "data_id": ids}
    for name, td in zip(ids, typedata):
        width, n, code = td[0], td[2] - td[1] + 1, td[3]
        if code == 2:
            out["iorig" if name == "porig" else name] = np.array(take(n*width, int), dtype=np.int32)
        elif code == 4:
            a = np.array(take(n*width, float))
            out[name] = a.reshape(n, width) if width > 1 else a
        elif code == 7:
            L = take(6, int)[3]
            st = {k: [] for k in ("r", "v", "m", "h", "radius")}
            for _ in range(n):
                take(1, str); take(2, int)
                sd = np.array(take(L, float))
                st["r"].append(sd[1:1 + ndim]); st["v"].append(sd[1 + ndim:1 + 2*ndim])
                st["m"].append(sd[1 + 2*ndim]); st["h"].append(sd[2 + 2*ndim]); st["radius"].append(sd[3 + 2*ndim])
            out["stars"] = {k: np.array(x) for k, x in st.items()}
    counts = idata[3:7]
    out["ptype"] = np.concatenate([np.full(int(c), pt, dtype=np.int32) for c, pt in zip(counts, FILE_ORDER)]) \
        if sum(counts) == N else np.zeros(N, dtype=np.int32)
    return out
