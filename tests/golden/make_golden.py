#!/usr/bin/env python
"""Generate the golden fixtures tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (needs oracle/_ref/ref_harness_std, i.e. `make -C oracle ref_std`,
which compiles /root/reference/src in place):

    python tests/golden/make_golden.py [case ...]

Each fixture holds, for one small configuration, the inputs and the reference's own outputs of
    UpdateAllSphProperties  (dens_in.* -> dens_out.*)      src/GradhSph/GradhSphTree.cpp:83-271
    UpdateAllSphHydroForces / UpdateAllSphForces (force_in.* -> force_out.*)   :280-657
and the reference's neighbour lists for the force pass (force_in.neib_off/idx, rows by iorig,
entries = iorig sorted), taken from its own tree walk + NeighbourManager.  Gravity cases add the
KD-tree cell table and the per-leaf interaction lists.  All random inputs are seeded.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import refio  # noqa: E402

REFERENCE = "/root/reference"

KEEP_PARTICLE = ["r", "v", "m", "h", "u", "rho", "invomega", "zeta", "hfactor", "hrangesqd", "sound",
                 "pressure", "alpha", "iorig", "level", "ptype"]
KEEP_FORCE_OUT = ["a", "atree", "gpot", "dudt", "div_v", "dalphadt", "levelneib", "iorig"]


def base_params(ndim, **kw):
    p = {
        "run_id": "GOLD", "sim": "gradhsph", "ic": "python", "ndim": ndim, "dimensionless": 1,
        "Nlevels": 1, "Nstepsmax": 1, "tend": 1.0e-10, "dt_snap": 1.0e10, "noutputstep": 100000,
        "ndiagstep": 100000, "out_file_form": "su", "hydro_forces": 1, "self_gravity": 0,
        "gas_eos": "energy_eqn", "gamma_eos": 1.6666666666666667, "kernel": "m4", "tabulated_kernel": 1,
        "h_fac": 1.2, "h_converge": 0.01, "avisc": "mon97", "acond": "none", "alpha_visc": 1.0,
        "beta_visc": 2.0, "neib_search": "kdtree", "Nleafmax": 6, "sink_particles": 0,
        "sph_integration": "lfkdk",
    }
    p.update(kw)
    return p


def periodic_box(p, ndim, lo, hi):
    for k in range(ndim):
        p["boxmin[%d]" % k] = lo[k]
        p["boxmax[%d]" % k] = hi[k]
        p["boundary_lhs[%d]" % k] = "periodic"
        p["boundary_rhs[%d]" % k] = "periodic"


def smooth_field(rng, r, ncomp, amp, L):
    """Smooth seeded pseudo-random field from a few Fourier modes."""
    out = np.zeros((r.shape[0], ncomp))
    for c in range(ncomp):
        for _ in range(4):
            k = rng.integers(1, 3, size=r.shape[1]) * 2 * np.pi / L
            ph = rng.uniform(0, 2 * np.pi)
            out[:, c] += amp * rng.uniform(0.3, 1.0) * np.sin(r @ k + ph)
    return out


def case_adsod():
    """The reference's own adiabatic Sod test (tests/hydro_tests/adsod.dat), verbatim."""
    dat = os.path.join(REFERENCE, "tests", "hydro_tests", "adsod.dat")
    d, _ = refio.run_harness(None, datfile=dat, perturb_h=0.1)
    params = {"ndim": 1, "kernel": "m4", "tabulated_kernel": 0, "h_fac": 1.2, "h_converge": 0.01,
              "avisc": "mon97", "acond": "none", "alpha_visc": 1.0, "beta_visc": 2.0, "gas_eos": "energy_eqn",
              "gamma_eos": 1.4, "boxmin[0]": -2.0, "boxmax[0]": 2.0, "boundary_lhs[0]": "periodic",
              "boundary_rhs[0]": "periodic", "neib_search": "bruteforce", "Nleafmax": 6, "self_gravity": 0,
              "hydro_forces": 1}
    return params, d


def case_box3d_m4tab():
    rng = np.random.default_rng(12345)
    n = 16
    g = (np.arange(n) + 0.5) / n
    r = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    r = (r + 0.2 / n * rng.uniform(-1, 1, r.shape)) % 1.0
    N = r.shape[0]
    v = smooth_field(rng, r, 3, 0.5, 1.0)
    m = np.full(N, 1.0 / N) * (1.0 + 0.1 * rng.uniform(-1, 1, N))
    u = 1.5 * (1.0 + 0.3 * smooth_field(rng, r, 1, 0.5, 1.0)[:, 0])
    h = np.full(N, 1.2 / n)
    p = base_params(3, Nhydromax=8 * N)
    periodic_box(p, 3, [0, 0, 0], [1, 1, 1])
    d, _ = refio.run_harness(p, ic=dict(r=r, v=v, m=m, h=h, u=u), perturb_h=0.15,
                             dumps="setup,density,forces,neib,tree")
    return p, d


def case_khi2d_quintic():
    """2-D, quintic tabulated, wadsley2008 conductivity, Nleafmax 12, alpha 0.1 beta 0.2: the
    option set of tests/hydro_tests/khi.dat on a small two-density layered box."""
    rng = np.random.default_rng(777)
    pts = []
    for (n, ylo, yhi) in ((32, -0.5, -0.25), (48, -0.25, 0.25), (32, 0.25, 0.5)):
        ny = int(round(n * (yhi - ylo)))
        gx = (np.arange(n) + 0.5) / n - 0.5
        gy = ylo + (np.arange(ny) + 0.5) / n
        q = np.stack(np.meshgrid(gx, gy, indexing="ij"), -1).reshape(-1, 2)
        q += 0.15 / n * rng.uniform(-1, 1, q.shape)
        pts.append((q, n))
    r = np.concatenate([q for q, _ in pts])
    m = np.concatenate([np.full(len(q), 1.0 / 48 ** 2) for q, n in pts])   # equal masses, rho 1 / 2.25
    r[:, 0] = (r[:, 0] + 0.5) % 1.0 - 0.5
    r[:, 1] = (r[:, 1] + 0.5) % 1.0 - 0.5
    N = len(m)
    v = np.zeros((N, 2))
    v[:, 0] = np.where(np.abs(r[:, 1]) < 0.25, 0.5, -0.5)
    v[:, 1] = 0.1 * np.sin(4 * np.pi * r[:, 0])
    rho0 = np.where(np.abs(r[:, 1]) < 0.25, 2.25, 1.0)
    u = 2.5 / rho0 / (5.0 / 3.0 - 1.0)
    h = 1.2 * np.sqrt(m / rho0)
    p = base_params(2, kernel="quintic", acond="wadsley2008", alpha_visc=0.1, beta_visc=0.2, Nleafmax=12,
                    Nhydromax=10 * N)
    periodic_box(p, 2, [-0.5, -0.5], [0.5, 0.5])
    d, _ = refio.run_harness(p, ic=dict(r=r, v=v, m=m, h=h, u=u), perturb_h=0.1)
    return p, d


def case_box3d_cd2010():
    """3-D periodic box with the Cullen & Dehnen (2010) viscosity switch (Sph.h:364-456): the density
    pass reads the neighbours' v and the accelerations of the reference's own setup force pass."""
    rng = np.random.default_rng(2010)
    n = 11
    g = (np.arange(n) + 0.5) / n
    r = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    r = (r + 0.25 / n * rng.uniform(-1, 1, r.shape)) % 1.0
    N = r.shape[0]
    v = smooth_field(rng, r, 3, 1.5, 1.0)
    m = np.full(N, 1.0 / N)
    u = 1.0 * (1.0 + 0.4 * smooth_field(rng, r, 1, 0.5, 1.0)[:, 0])
    h = np.full(N, 1.2 / n)
    p = base_params(3, avisc="mon97", time_dependent_avisc="cd2010", alpha_visc_min=0.1, Nhydromax=8 * N)
    periodic_box(p, 3, [0, 0, 0], [1, 1, 1])
    d, _ = refio.run_harness(p, ic=dict(r=r, v=v, m=m, h=h, u=u), perturb_h=0.1)
    return p, d


def case_shock1d_gaussian():
    """1-D periodic, GAUSSIAN kernel (analytic), the kernel of tests/hydro_tests/adshock.dat and
    soundwave.dat: two states with a velocity jump."""
    rng = np.random.default_rng(909)
    n1, n2 = 320, 160
    x = np.concatenate([-2.0 + 2.0*(np.arange(n1) + 0.5)/n1, 2.0*(np.arange(n2) + 0.5)/n2])
    x += 0.05*rng.uniform(-1, 1, x.size)*np.concatenate([np.full(n1, 2.0/n1), np.full(n2, 2.0/n2)])
    N = x.size
    m = np.full(N, 2.0/n1)
    rho0 = np.where(x < 0, 1.0, 0.5)
    v = np.where(x < 0, 0.5, -0.5)[:, None]
    u = np.where(x < 0, 2.5, 2.0)
    h = 1.2*m/rho0
    p = base_params(1, kernel="gaussian", tabulated_kernel=0, gamma_eos=1.4, Nhydromax=12*N)
    periodic_box(p, 1, [-2.0], [2.0])
    d, _ = refio.run_harness(p, ic=dict(r=x[:, None], v=v, m=m, h=h, u=u), perturb_h=0.1)
    return p, d


def case_rti2d_mirror():
    """2-D, periodic in x, MIRROR walls in y (the boundary set of tests/hydro_tests/rti.dat), quintic
    tabulated: two layers of different density with a sheared, perturbed velocity field."""
    rng = np.random.default_rng(4242)
    nx, ny = 32, 64
    gx = (np.arange(nx) + 0.5) / nx * 0.5
    gy = (np.arange(ny) + 0.5) / ny
    r = np.stack(np.meshgrid(gx, gy, indexing="ij"), -1).reshape(-1, 2)
    r += 0.2 / ny * rng.uniform(-1, 1, r.shape)
    r[:, 0] = r[:, 0] % 0.5
    r[:, 1] = np.clip(r[:, 1], 1.0e-4, 1.0 - 1.0e-4)
    N = len(r)
    rho0 = np.where(r[:, 1] > 0.5, 2.0, 1.0)
    m = rho0 * 0.5 / N / 1.5 * 1.5
    v = np.stack([0.3 * np.sin(2 * np.pi * r[:, 1]), 0.1 * np.cos(4 * np.pi * r[:, 0]) * (1 + r[:, 1])], -1)
    u = 2.5 / rho0 / (1.4 - 1.0)
    h = 1.3 * np.sqrt(m / rho0)
    p = base_params(2, kernel="quintic", gamma_eos=1.4, Nhydromax=12 * N)
    periodic_box(p, 2, [0.0, 0.0], [0.5, 1.0])
    p["boundary_lhs[1]"] = "mirror"
    p["boundary_rhs[1]"] = "mirror"
    d, _ = refio.run_harness(p, ic=dict(r=r, v=v, m=m, h=h, u=u), perturb_h=0.1, dumps="setup,density,forces")
    return p, d


def case_box3d_mirror():
    """3-D: mirror walls on both x sides and the lower z side, open in y and above.  (Not periodic
    in y: with a mix of periodic and non-periodic dimensions the reference's force walk reads an
    uninitialised shift, GhostNeighbours.hpp:172-192 `dr_corr[k]` of the non-periodic dimensions, and
    can miss neighbours across the periodic boundary -- observed with this very set-up.)"""
    rng = np.random.default_rng(5151)
    n = 11
    g = (np.arange(n) + 0.5) / n
    r = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    r += 0.25 / n * rng.uniform(-1, 1, r.shape)
    r[:, 0] = np.clip(r[:, 0], 1.0e-4, 1.0 - 1.0e-4)
    r[:, 2] = np.maximum(r[:, 2], 1.0e-4)
    N = len(r)
    v = smooth_field(rng, r, 3, 0.6, 1.0)
    m = np.full(N, 1.0 / N)
    u = 1.0 * (1.0 + 0.3 * smooth_field(rng, r, 1, 0.5, 1.0)[:, 0])
    h = np.full(N, 1.2 / n)
    p = base_params(3, Nhydromax=12 * N)
    for k, (lo, hi) in enumerate((("mirror", "mirror"), ("open", "open"), ("mirror", "open"))):
        p["boxmin[%d]" % k] = 0.0
        p["boxmax[%d]" % k] = 1.0 if k < 2 else 2.0
        p["boundary_lhs[%d]" % k] = lo
        p["boundary_rhs[%d]" % k] = hi
    d, _ = refio.run_harness(p, ic=dict(r=r, v=v, m=m, h=h, u=u), perturb_h=0.1, dumps="setup,density,forces")
    return p, d


def sphere_ic(seed, N):
    rng = np.random.default_rng(seed)
    # centrally condensed sphere rho ~ 1/r: r = R*sqrt(U)
    rad = np.sqrt(rng.uniform(0, 1, N))
    d = rng.normal(size=(N, 3))
    d /= np.linalg.norm(d, axis=1)[:, None]
    r = d * rad[:, None]
    v = 0.3 * np.cross(np.array([0, 0, 1.0]), r) + 0.05 * rng.normal(size=(N, 3))
    m = np.full(N, 1.0 / N)
    u = np.full(N, 0.05) * (1 + 0.2 * rng.uniform(-1, 1, N))
    h = np.zeros(N)          # let the reference make its own initial guess
    return dict(r=r, v=v, m=m, h=h, u=u)


def case_sphere_grav_quad():
    ic = sphere_ic(2024, 3000)
    p = base_params(3, self_gravity=1, multipole="quadrupole", gravity_mac="geometric", thetamaxsqd=0.1)
    d, _ = refio.run_harness(p, ic=ic, perturb_h=0.1, dumps="setup,density,forces,neib,tree,gravlists")
    return p, d


def case_sphere_grav_fastmono():
    ic = sphere_ic(2025, 2500)
    p = base_params(3, self_gravity=1, multipole="fast_monopole", gravity_mac="geometric", thetamaxsqd=0.15,
                    gas_eos="barotropic", temp0=10.0, mu_bar=2.35, rho_bary=5.0, gamma_eos=1.4)
    d, _ = refio.run_harness(p, ic=ic, perturb_h=0.05, dumps="setup,density,forces,neib,tree,gravlists")
    return p, d


def case_sphere_grav_mono():
    ic = sphere_ic(2026, 2000)
    p = base_params(3, self_gravity=1, multipole="monopole", gravity_mac="geometric", thetamaxsqd=0.2,
                    gas_eos="isothermal", temp0=2.0, mu_bar=1.0)
    d, _ = refio.run_harness(p, ic=ic, perturb_h=0.05, dumps="setup,density,forces,neib,tree,gravlists")
    return p, d


def case_sphere_grav_bruteforce():
    """the reference's SECOND, tree-independent path (neib_search = bruteforce, GradhSphBruteForce): every pair is
    either a kernel neighbour (softened terms) or an exact direct sum -- no tree, no opening criterion"""
    ic = sphere_ic(2026, 2000)
    p = base_params(3, self_gravity=1, multipole="monopole", gravity_mac="geometric", thetamaxsqd=0.2,
                    gas_eos="isothermal", temp0=2.0, mu_bar=1.0, neib_search="bruteforce")
    d, _ = refio.run_harness(p, ic=ic, perturb_h=0.05, dumps="setup,density,forces")
    return p, d


STARS = np.array([[0.05, -0.02, 0.03, 0.4, 0.08],      # {r, m, h}: one star inside the cloud (its mean-h kernel
                  [1.3, 0.4, -0.2, 0.7, 0.05],         # overlaps many particles), one outside, one at the edge
                  [-0.55, 0.5, 0.45, 0.2, 0.3]])


def case_sphere_grav_stars():
    """Star particles: their terms in zeta (GradhSph.cpp:288-306, conservative_sph_star_gravity = 1 and
    nbody_softening = 1 are the defaults) and ComputeStarGravForces (GradhSph.cpp:699-743) in the force pass."""
    ic = sphere_ic(2029, 2000)
    p = base_params(3, self_gravity=1, multipole="quadrupole", gravity_mac="geometric", thetamaxsqd=0.15,
                    Nstar=len(STARS), nbody="lfkdk")
    d, _ = refio.run_harness(p, ic=ic, perturb_h=0.05, dumps="setup,density,forces,neib,tree,gravlists", stars=STARS)
    d["stars"] = STARS
    return p, d


def case_stargas_hermite4():
    """HydroTree::UpdateAllStarGasForces (HydroTree.cpp:552-654): gravity of the gas on three stars, with the
    Hermite-4 N-body class so that adot is filled too.  Only what that operator reads and writes is kept."""
    ic = sphere_ic(2031, 1500)
    p = base_params(3, self_gravity=1, multipole="quadrupole", gravity_mac="geometric", thetamaxsqd=0.15,
                    Nstar=len(STARS), nbody="hermite4")
    d, _ = refio.run_harness(p, ic=ic, perturb_h=0.05, dumps="forces,stargas", stars=STARS)
    keep = {k: v for k, v in d.items() if k.startswith("stargas.")}
    for f in ("r", "v", "m", "h", "iorig"):
        keep["force_in." + f] = d["force_in." + f]
    keep["stars"] = STARS
    return p, keep


def case_sphere_grav_gadget2():
    """Gadget-2 relative opening criterion (Tree.h:413-432): the cells' amin comes from the atree the
    reference's own setup force pass left in the particles (dens_in.atree)."""
    ic = sphere_ic(2027, 2200)
    p = base_params(3, self_gravity=1, multipole="quadrupole", gravity_mac="gadget2", macerror=5.0e-4,
                    thetamaxsqd=0.5)
    d, _ = refio.run_harness(p, ic=ic, perturb_h=0.05, dumps="setup,density,forces,neib,tree,gravlists")
    return p, d


def case_sphere_grav_eigen():
    """eigenvalue opening criterion: cell.mac from the quadrupole, macfactor from the previous gpot."""
    ic = sphere_ic(2028, 2100)
    p = base_params(3, self_gravity=1, multipole="fast_quadrupole", gravity_mac="eigenmac", macerror=1.0e-3,
                    thetamaxsqd=0.5)
    d, _ = refio.run_harness(p, ic=ic, perturb_h=0.05, dumps="setup,density,forces,neib,tree,gravlists")
    return p, d


def case_box3d_mm97():
    """analytic (untabulated) M4, mm97 time-dependent viscosity, price2008 conductivity."""
    rng = np.random.default_rng(4321)
    n = 12
    g = (np.arange(n) + 0.5) / n
    r = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    r = (r + 0.25 / n * rng.uniform(-1, 1, r.shape)) % 1.0
    N = r.shape[0]
    v = smooth_field(rng, r, 3, 0.8, 1.0)
    m = np.full(N, 1.0 / N)
    u = 1.0 * (1.0 + 0.4 * smooth_field(rng, r, 1, 0.5, 1.0)[:, 0])
    h = np.full(N, 1.2 / n)
    p = base_params(3, tabulated_kernel=0, avisc="mon97", time_dependent_avisc="mm97", acond="price2008",
                    alpha_visc_min=0.1, Nhydromax=8 * N)
    periodic_box(p, 3, [0, 0, 0], [1, 1, 1])
    d, _ = refio.run_harness(p, ic=dict(r=r, v=v, m=m, h=h, u=u), perturb_h=0.2)
    return p, d


CASES = {
    "adsod": case_adsod,
    "box3d_m4tab": case_box3d_m4tab,
    "khi2d_quintic": case_khi2d_quintic,
    "box3d_cd2010": case_box3d_cd2010,
    "shock1d_gaussian": case_shock1d_gaussian,
    "rti2d_mirror": case_rti2d_mirror,
    "box3d_mirror": case_box3d_mirror,
    "sphere_grav_quad": case_sphere_grav_quad,
    "sphere_grav_fastmono": case_sphere_grav_fastmono,
    "sphere_grav_mono": case_sphere_grav_mono,
    "brute_sphere_grav": case_sphere_grav_bruteforce,
    "sphere_grav_gadget2": case_sphere_grav_gadget2,
    "sphere_grav_stars": case_sphere_grav_stars,
    "stargas_hermite4": case_stargas_hermite4,
    "sphere_grav_eigen": case_sphere_grav_eigen,
    "box3d_mm97": case_box3d_mm97,
}


def pack(params, d, keep_direct=False):
    out = {}
    for k, v in params.items():
        out["param." + k] = np.array(str(v))
    for stage in ("dens_in", "dens_out", "force_in"):
        for f in KEEP_PARTICLE:
            key = "%s.%s" % (stage, f)
            if key in d:
                out[key] = d[key]
    if str(params.get("time_dependent_avisc", "none")) == "cd2010":
        for st in ("dens_in", "dens_out"):
            for f in ("a", "dalphadt"):          # the switch reads a (previous step) and writes alpha, dalphadt
                out["%s.%s" % (st, f)] = d["%s.%s" % (st, f)]
    if str(params.get("gravity_mac", "geometric")) != "geometric":
        for f in ("atree", "gpot"):          # what the cells' amin / macfactor were stocked from
            out["dens_in." + f] = d["dens_in." + f]
    if "stars" in d:
        out["stars"] = d["stars"]
    for f in KEEP_FORCE_OUT:
        key = "force_out.%s" % f
        if key in d:
            out[key] = d[key]
    for k, v in d.items():
        if k.startswith("stargas."):
            out[k] = v
        if ("neib_" in k or "tree_" in k or "direct_" in k or "smooth_" in k or "gravcell" in k
                or "leaf_of" in k):
            if k.endswith("direct_idx") and not keep_direct:
                continue          # 5 MB per case; the per-particle counts (direct_off) are kept
            if k.startswith("grav_in.neib_"):
                continue          # identical to force_in.neib_* (checked at generation time)
            out[k] = v
    return out


def main():
    names = sys.argv[1:] or list(CASES)
    for name in names:
        params, d = CASES[name]()
        if "grav_in.neib_idx" in d:
            assert np.array_equal(d["grav_in.neib_idx"], d["force_in.neib_idx"])
            assert np.array_equal(d["grav_in.neib_off"], d["force_in.neib_off"])
        out = pack(params, d, keep_direct=(name == "sphere_grav_mono"))
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print("%-24s N=%-6d %7.1f kB  (%d arrays)" % (name, len(d["dens_in.m"] if "dens_in.m" in d else d["force_in.m"]), os.path.getsize(path) / 1e3,
                                                      len(out)))


if __name__ == "__main__":
    main()
