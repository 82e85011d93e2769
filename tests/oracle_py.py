"""ctypes binding of the CPU oracle (oracle/gsph_oracle.c) -- TEST INFRASTRUCTURE ONLY.

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / reference legs may
import this module.  The product path (gandalf_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_ORACLE_DIR = os.path.join(_ROOT, "oracle")
_LIB_PATH = os.path.join(_ORACLE_DIR, "_build", "liboracle.so")


class OrcParams(C.Structure):
    _fields_ = [
        ("ndim", C.c_int), ("kernel", C.c_int), ("tabulated", C.c_int),
        ("h_fac", C.c_double), ("h_converge", C.c_double),
        ("avisc", C.c_int), ("acond", C.c_int), ("tdavisc", C.c_int),
        ("alpha_visc", C.c_double), ("beta_visc", C.c_double), ("alpha_visc_min", C.c_double),
        ("eos", C.c_int),
        ("gamma_eos", C.c_double), ("temp0", C.c_double), ("mu_bar", C.c_double),
        ("rho_bary", C.c_double),
        ("boxmin", C.c_double * 3), ("boxmax", C.c_double * 3), ("periodic", C.c_int * 3),
        ("self_gravity", C.c_int), ("thetamaxsqd", C.c_double), ("gravity_mac", C.c_int),
        ("macerror", C.c_double), ("multipole", C.c_int), ("Nleafmax", C.c_int),
        ("hydro_forces", C.c_int), ("ntarget", C.c_int),
        ("star_zeta", C.c_int), ("nbody_softening", C.c_int), ("nstar", C.c_int), ("stars", C.c_void_p),
    ]


def build(force=False):
    if force or not os.path.exists(_LIB_PATH) or (
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_ORACLE_DIR, "gsph_oracle.c"))):
        subprocess.check_call(["make", "-C", _ORACLE_DIR, "port"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_lp = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        P = C.POINTER(OrcParams)
        _lib.orc_kernel.restype = C.c_double
        _lib.orc_kernel.argtypes = [P, C.c_int, C.c_double]
        _lib.orc_kernel_s2.restype = C.c_double
        _lib.orc_kernel_s2.argtypes = [P, C.c_int, C.c_double]
        _lib.orc_kernel_table.restype = None
        _lib.orc_kernel_table.argtypes = [P, C.c_int, _dp]
        _lib.orc_neighbour_lists.restype = C.c_int64
        _lib.orc_neighbour_lists.argtypes = [P, C.c_int, _dp, _dp, _lp, C.c_void_p]
        _lib.orc_density.restype = C.c_int
        _lib.orc_density.argtypes = [P, C.c_int] + [_dp] * 11 + [_ip]
        _lib.orc_hydro_forces.restype = C.c_int
        _lib.orc_hydro_forces.argtypes = [P, C.c_int] + [_dp] * 12 + [_ip] + [_dp] * 4 + [_ip]
        _lib.orc_hydro_grav_forces_direct.restype = C.c_int
        _lib.orc_hydro_grav_forces_direct.argtypes = [P, C.c_int] + [_dp] * 13 + [_ip] + [_dp] * 6 + [_ip]
        _lib.orc_cd2010_switch.restype = C.c_int
        _lib.orc_cd2010_switch.argtypes = [P, C.c_int] + [_dp] * 10
        _lib.orc_star_grav_forces.restype = C.c_int
        _lib.orc_star_grav_forces.argtypes = [P, C.c_int, _dp, _dp, C.c_int, _dp, _dp, _dp, _dp]
        _lib.orc_tree_build.restype = C.c_void_p
        _lib.orc_tree_build.argtypes = [P, C.c_int, _dp, _dp, _dp, _dp]
        _lib.orc_tree_build_mac.restype = C.c_void_p
        _lib.orc_tree_build_mac.argtypes = [P, C.c_int, _dp, _dp, _dp, _dp, C.c_void_p, C.c_void_p]
        _lib.orc_tree_free.restype = None
        _lib.orc_tree_free.argtypes = [C.c_void_p]
        _lib.orc_tree_ncell.restype = C.c_int
        _lib.orc_tree_ncell.argtypes = [C.c_void_p]
        _lib.orc_tree_ltot.restype = C.c_int
        _lib.orc_tree_ltot.argtypes = [C.c_void_p]
        _lib.orc_tree_leaf_of.restype = None
        _lib.orc_tree_leaf_of.argtypes = [C.c_void_p, _ip]
        _lib.orc_tree_cells.restype = None
        _lib.orc_tree_cells.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, _ip]
        _lib.orc_hydro_grav_forces_tree.restype = C.c_int
        _lib.orc_hydro_grav_forces_tree.argtypes = ([P, C.c_void_p, C.c_int] + [_dp] * 13 + [_ip]
                                                    + [_dp] * 6 + [_ip, _ip, _ip])
    return _lib


_KERNELS = {"m4": 0, "quintic": 1, "gaussian": 2}
_AVISC = {"none": 0, "mon97": 1, "mon97mm97": 2, "mon97cd2010": 3}
_ACOND = {"none": 0, "wadsley2008": 1, "price2008": 2}
_TDAVISC = {"none": 0, "mm97": 1, "cd2010": 2}
_EOS = {"energy_eqn": 0, "constant_temp": 0, "isothermal": 1, "barotropic": 2, "barotropic2": 3}
_MULTIPOLE = {"monopole": 0, "quadrupole": 1, "fast_monopole": 2, "fast_quadrupole": 3}
_MAC = {"geometric": 0, "eigenmac": 1, "gadget2": 2}


def params_from_dict(d):
    """d uses the reference's .dat parameter names (src/Common/Parameters.cpp:166-536)."""
    p = OrcParams()
    p.ndim = int(d.get("ndim", 3))
    p.kernel = _KERNELS[d.get("kernel", "m4")]
    p.tabulated = int(d.get("tabulated_kernel", 1))
    p.h_fac = float(d.get("h_fac", 1.2))
    p.h_converge = float(d.get("h_converge", 0.01))
    # src/GradhSph/GradhSphSimulation.cpp:71-92: the pair (avisc, time_dependent_avisc) selects both enums
    av, td = d.get("avisc", "mon97"), d.get("time_dependent_avisc", "none")
    if av == "none":
        p.avisc, p.tdavisc = 0, 0
    elif av == "mon97" and td == "mm97":
        p.avisc, p.tdavisc = 2, 1
    elif av == "mon97" and td == "cd2010":
        p.avisc, p.tdavisc = 3, 2
    elif av == "mon97":
        p.avisc, p.tdavisc = 1, 0
    else:
        raise ValueError("Unrecognised parameter : avisc = %s" % av)
    p.acond = _ACOND[d.get("acond", "none")]
    p.alpha_visc = float(d.get("alpha_visc", 1.0))
    p.beta_visc = float(d.get("beta_visc", 2.0))
    p.alpha_visc_min = float(d.get("alpha_visc_min", 0.1))
    p.eos = _EOS[d.get("gas_eos", "energy_eqn")]
    p.gamma_eos = float(d.get("gamma_eos", 1.66666666666666))
    p.temp0 = float(d.get("temp0", 1.0))
    p.mu_bar = float(d.get("mu_bar", 1.0))
    p.rho_bary = float(d.get("rho_bary", 1.0e-14))
    for k in range(3):
        p.boxmin[k] = float(d.get("boxmin[%d]" % k, 0.0))
        p.boxmax[k] = float(d.get("boxmax[%d]" % k, 0.0))
        p.periodic[k] = 1 if d.get("boundary_lhs[%d]" % k, "open") == "periodic" else 0
    p.self_gravity = int(d.get("self_gravity", 0))
    p.thetamaxsqd = float(d.get("thetamaxsqd", 0.1))
    p.gravity_mac = _MAC[d.get("gravity_mac", "geometric")]
    p.macerror = float(d.get("macerror", 1e-4))
    p.multipole = _MULTIPOLE[d.get("multipole", "quadrupole")]
    p.Nleafmax = int(d.get("Nleafmax", 6))
    p.hydro_forces = int(d.get("hydro_forces", 1))
    p.ntarget = 0
    p.star_zeta = int(d.get("conservative_sph_star_gravity", 1))
    p.nbody_softening = int(d.get("nbody_softening", 1))
    p.nstar, p.stars = 0, None
    # mirror walls: (dimension, wall position) for boundary_lhs/rhs[k] = mirror.  The oracle treats them
    # the way the reference's density pass does: the mirror images are materialised as extra source
    # particles (Hydrodynamics.cpp:217-290: r_k -> 2*wall - r_k, v_k -> -v_k, incl. images of images)
    p.mirrors, p.mirror_sides = [], []
    for k in range(p.ndim):
        if d.get("boundary_lhs[%d]" % k, "open") == "mirror":
            p.mirrors.append((k, float(d.get("boxmin[%d]" % k, 0.0)))); p.mirror_sides.append("lhs")
        if d.get("boundary_rhs[%d]" % k, "open") == "mirror":
            p.mirrors.append((k, float(d.get("boxmax[%d]" % k, 0.0)))); p.mirror_sides.append("rhs")
    return p


def mirror_images(p, fields):
    """fields: dict name -> array with leading dimension N.  Returns (extended dict, N): the particles
    followed by every mirror image (one wall per dimension at most per image; images of images across
    dimensions), positions reflected as 2*wall - r_k and v_k flipped."""
    mirrors = getattr(p, "mirrors", [])
    N = len(fields["m"])
    if not mirrors:
        return fields, N
    nd = p.ndim
    combos = [[]]
    for k in range(nd):
        walls = [w for (kk, w) in mirrors if kk == k]
        combos = [c + [opt] for c in combos for opt in [None] + walls]
    ext = {k: [np.asarray(v)] for k, v in fields.items()}
    for c in combos:
        if all(w is None for w in c):
            continue
        r = np.array(fields["r"], dtype=np.float64).reshape(N, nd).copy()
        v = np.array(fields["v"], dtype=np.float64).reshape(N, nd).copy() if "v" in fields else None
        for k, w in enumerate(c):
            if w is not None:
                r[:, k] = 2*w - r[:, k]
                if v is not None:
                    v[:, k] = -v[:, k]
        for name in fields:
            if name == "r":
                ext[name].append(r)
            elif name == "v":
                ext[name].append(v)
            else:
                ext[name].append(np.asarray(fields[name]))
    return {k: np.concatenate([np.asarray(a).reshape((N,) + np.asarray(fields[k]).shape[1:]) for a in v], axis=0)
            for k, v in ext.items()}, N


def _c(a, dtype=np.float64):
    return np.ascontiguousarray(a, dtype=dtype)


def kernel_table(p, which):
    out = np.empty(1000)
    lib().orc_kernel_table(C.byref(p), which, out)
    return out


def neighbour_lists(p, r, hrangesqd):
    r = _c(r).reshape(len(hrangesqd), -1)
    N = r.shape[0]
    off = np.zeros(N + 1, dtype=np.int64)
    hr = _c(hrangesqd)
    tot = lib().orc_neighbour_lists(C.byref(p), N, r, hr, off, None)
    idx = np.empty(tot, dtype=np.int32)
    lib().orc_neighbour_lists(C.byref(p), N, r, hr, off, idx.ctypes.data_as(C.c_void_p))
    return off, idx


def density(p, r, m, h, u, stars=None):
    """stars = (Nstar, ndim + 2) rows {r, m, h}: star terms of zeta (conservative_sph_star_gravity)"""
    Nt = len(m)
    if stars is not None and len(stars):
        stars = _c(stars)
        p.nstar, p.stars = len(stars), stars.ctypes.data
    else:
        p.nstar, p.stars = 0, None
    ext, _ = mirror_images(p, dict(r=_c(r).reshape(Nt, -1), m=_c(m), h=_c(h), u=_c(u)))
    N = len(ext["m"])
    r = _c(ext["r"]).reshape(N, -1)
    out = {k: np.zeros(N) for k in ("rho", "invomega", "zeta", "hfactor", "hrangesqd", "sound", "pressure")}
    out["h"] = _c(ext["h"]).copy()
    out["u"] = _c(ext["u"]).copy()
    out["niter"] = np.zeros(N, dtype=np.int32)
    p.ntarget = Nt if N != Nt else 0
    rc = lib().orc_density(C.byref(p), N, r, _c(ext["m"]), out["h"], out["u"], out["rho"], out["invomega"],
                           out["zeta"], out["hfactor"], out["hrangesqd"], out["sound"], out["pressure"],
                           out["niter"])
    p.ntarget = 0
    p.nstar, p.stars = 0, None
    if rc != 0:
        raise RuntimeError("oracle: h-rho iteration did not converge")
    return {k: v[:Nt] for k, v in out.items()}


def cd2010_switch(p, r, v, a, m, dens, alpha):
    """Cullen & Dehnen switch at the end of ComputeH: dens = output of density(); returns (alpha, dalphadt)"""
    N = len(m)
    nd = p.ndim
    al = _c(alpha).copy()
    da = np.zeros(N)
    lib().orc_cd2010_switch(C.byref(p), N, _c(r).reshape(N, nd), _c(v).reshape(N, nd), _c(a).reshape(N, nd), _c(m),
                            _c(dens["h"]), _c(dens["rho"]), _c(dens["hfactor"]), _c(dens["sound"]), al, da)
    return al, da


def hydro_forces(p, s):
    """s: dict with r,v,m,h,rho,pressure,invomega,sound,u,alpha,hfactor,hrangesqd,level."""
    Nt = len(s["m"])
    nd = p.ndim
    names = ("r", "v", "m", "h", "rho", "pressure", "invomega", "sound", "u", "alpha", "hfactor", "hrangesqd", "level")
    base = {k: np.asarray(s[k]) for k in names}
    base["r"] = base["r"].reshape(Nt, nd); base["v"] = base["v"].reshape(Nt, nd)
    s, _ = mirror_images(p, base)
    N = len(s["m"])
    out = dict(a=np.zeros((N, nd)), dudt=np.zeros(N), div_v=np.zeros(N), dalphadt=np.zeros(N),
               levelneib=np.zeros(N, dtype=np.int32))
    p.ntarget = Nt if N != Nt else 0
    lib().orc_hydro_forces(C.byref(p), N, _c(s["r"]).reshape(N, nd), _c(s["v"]).reshape(N, nd), _c(s["m"]),
                           _c(s["h"]), _c(s["rho"]), _c(s["pressure"]), _c(s["invomega"]), _c(s["sound"]),
                           _c(s["u"]), _c(s["alpha"]), _c(s["hfactor"]), _c(s["hrangesqd"]),
                           _c(s["level"], np.int32), out["a"], out["dudt"], out["div_v"], out["dalphadt"],
                           out["levelneib"])
    p.ntarget = 0
    return {k: v[:Nt] for k, v in out.items()}


def _grav_args(p, s):
    N = len(s["m"])
    nd = p.ndim
    return [_c(s["r"]).reshape(N, nd), _c(s["v"]).reshape(N, nd), _c(s["m"]), _c(s["h"]), _c(s["rho"]),
            _c(s["pressure"]), _c(s["invomega"]), _c(s["zeta"]), _c(s["sound"]), _c(s["u"]), _c(s["alpha"]),
            _c(s["hfactor"]), _c(s["hrangesqd"]), _c(s["level"], np.int32)]


def _grav_out(N, nd):
    return dict(a=np.zeros((N, nd)), atree=np.zeros((N, nd)), gpot=np.zeros(N), dudt=np.zeros(N),
                div_v=np.zeros(N), dalphadt=np.zeros(N), levelneib=np.zeros(N, dtype=np.int32))


def hydro_grav_forces_direct(p, s):
    N = len(s["m"])
    out = _grav_out(N, p.ndim)
    lib().orc_hydro_grav_forces_direct(C.byref(p), N, *_grav_args(p, s), out["a"], out["atree"], out["gpot"],
                                       out["dudt"], out["div_v"], out["dalphadt"], out["levelneib"])
    return out


def star_grav_forces(p, r, h, rstar, mstar, hstar):
    """GradhSph::ComputeStarGravForces for every particle -> (a, atree, gpot) contributions"""
    N = len(h)
    nd = p.ndim
    stars = np.ascontiguousarray(np.concatenate([np.asarray(rstar, dtype=np.float64).reshape(len(mstar), nd),
                                                 np.asarray(mstar, dtype=np.float64)[:, None],
                                                 np.asarray(hstar, dtype=np.float64)[:, None]], axis=1))
    a = np.zeros((N, nd)); atree = np.zeros((N, nd)); gpot = np.zeros(N)
    lib().orc_star_grav_forces(C.byref(p), N, _c(r).reshape(N, nd), _c(h), len(mstar), stars, a, atree, gpot)
    return a, atree, gpot


class Tree:
    def __init__(self, p, r, v, m, h, atree=None, gpot=None):
        """atree / gpot: the previous step's values the gadget2 / eigenmac criteria read at stock time"""
        N = len(m)
        self.p, self.N = p, N
        # keep the arrays alive: the C tree stores pointers
        self._r = _c(r).reshape(N, p.ndim)
        self._v = _c(v).reshape(N, p.ndim)
        self._m, self._h = _c(m), _c(h)
        self._atree = None if atree is None else _c(atree).reshape(N, p.ndim)
        self._gpot = None if gpot is None else _c(gpot)
        self.ptr = lib().orc_tree_build_mac(C.byref(p), N, self._r, self._v, self._m, self._h,
                                            None if self._atree is None else self._atree.ctypes.data,
                                            None if self._gpot is None else self._gpot.ctypes.data)
        self.ncell = lib().orc_tree_ncell(self.ptr)
        self.ltot = lib().orc_tree_ltot(self.ptr)

    def __del__(self):
        if getattr(self, "ptr", None):
            lib().orc_tree_free(self.ptr)
            self.ptr = None

    def update_hmax(self, h):
        """KDTree::UpdateHmaxValues after a density pass (cdistsqd keeps the value of the stock)"""
        self._h = _c(h).copy()
        L = lib()
        L.orc_tree_update_hmax.restype = None
        L.orc_tree_update_hmax.argtypes = [C.c_void_p, _dp]
        L.orc_tree_update_hmax(self.ptr, self._h)

    def leaf_of(self):
        out = np.full(self.N, -1, dtype=np.int32)
        lib().orc_tree_leaf_of(self.ptr, out)
        return out

    def cells(self):
        nd, nc = self.p.ndim, self.ncell
        box = np.zeros((nc, 4 * nd)); com = np.zeros((nc, 2 * nd)); sc = np.zeros((nc, 5)); q = np.zeros((nc, 5))
        n = np.zeros(nc, dtype=np.int32)
        lib().orc_tree_cells(self.ptr, box, com, sc, q, n)
        return dict(box=box, com=com, scalars=sc, q=q, N=n)

    def star_gas_forces(self, rstar, vstar, hstar):
        """HydroTree::UpdateAllStarGasForces on this tree -> (a, adot, gpot, counts[Nstar][neib, direct, cells])"""
        nd = self.p.ndim
        ns = len(hstar)
        rows = np.ascontiguousarray(np.concatenate([np.asarray(rstar, dtype=np.float64).reshape(ns, nd),
                                                    np.asarray(vstar, dtype=np.float64).reshape(ns, nd),
                                                    np.asarray(hstar, dtype=np.float64)[:, None]], axis=1))
        a = np.zeros((ns, nd)); adot = np.zeros((ns, nd)); gpot = np.zeros(ns)
        counts = np.zeros((ns, 3), dtype=np.int32)
        L = lib()
        L.orc_star_gas_forces.restype = C.c_int
        L.orc_star_gas_forces.argtypes = [C.POINTER(OrcParams), C.c_void_p, C.c_int, _dp, _dp, _dp, _dp, _ip]
        L.orc_star_gas_forces(C.byref(self.p), self.ptr, ns, rows, a, adot, gpot, counts.reshape(-1))
        return a, adot, gpot, counts

    def hydro_grav_forces(self, s):
        N = self.N
        out = _grav_out(N, self.p.ndim)
        out["ndirect"] = np.zeros(N, dtype=np.int32)
        out["ncell"] = np.zeros(N, dtype=np.int32)
        rc = lib().orc_hydro_grav_forces_tree(C.byref(self.p), self.ptr, N, *_grav_args(self.p, s), out["a"],
                                              out["atree"], out["gpot"], out["dudt"], out["div_v"],
                                              out["dalphadt"], out["levelneib"], out["ndirect"], out["ncell"])
        if rc != 0:
            raise RuntimeError("oracle tree gravity: unsupported parameters")
        return out


# ---------------------------------------------------------------------------------------------------
# main loop (lfkdk, one global time step): restatement of SphSimulation::MainLoop around the operator
# functions above.  TEST INFRASTRUCTURE, the checker of gsph_integrator_*.
# ---------------------------------------------------------------------------------------------------
def global_timestep(p, s, courant_mult, accel_mult, energy_mult, energy_integration):
    """the minimum of SphIntegration::Timestep over the particles (Simulation::ComputeGlobalTimestep,
    Simulation.cpp:1690-1730)"""
    return float(particle_timesteps(p, s, courant_mult, accel_mult, energy_mult, energy_integration).min())


def particle_timesteps(p, s, courant_mult, accel_mult, energy_mult, energy_integration):
    """SphIntegration::Timestep per particle (SphIntegration.cpp:76-124)"""
    h, dv = s["h"], np.abs(s["div_v"])
    if p.hydro_forces == 1:
        ts = courant_mult*h/(s["sound"] + h*dv + 1.0e-50)
    else:
        ts = courant_mult*h/(h*dv + 1.0e-50)
    a = s["a"].reshape(len(h), -1)
    a2 = a[:, 0]*a[:, 0]
    for k in range(1, a.shape[1]):
        a2 = a2 + a[:, k]*a[:, k]
    ts = np.minimum(ts, accel_mult*np.sqrt(h/(np.sqrt(a2) + 1.0e-50)))
    if energy_integration:
        te = energy_mult*(s["u"]/(np.abs(s["dudt"]) + 1.0e-20))
        ts = np.where(s["ptype"] == 0, np.minimum(ts, te), ts)
    return ts


def star_star_forces(p, sr, sm, sh):
    """NbodyLeapfrogKDK::CalculateDirectSmoothedGravForces (NbodyLeapfrogKDK.cpp): softened gravity between
    the stars with the mean-h kernel -> (a, gpot)"""
    L = lib()
    ns, nd = sr.shape
    a = np.zeros((ns, nd)); gpot = np.zeros(ns)
    for i in range(ns):
        for j in range(ns):
            if i == j:
                continue
            dr = sr[j] - sr[i]
            drmag = np.sqrt(float(np.dot(dr, dr))) + 1.0e-20
            invdrmag = 1.0/drmag
            invhmean = 2.0/(sh[i] + sh[j])
            paux = sm[j]*invhmean*invhmean*L.orc_kernel(C.byref(p), 4, drmag*invhmean)*invdrmag
            gpot[i] += sm[j]*invhmean*L.orc_kernel(C.byref(p), 5, drmag*invhmean)
            a[i] += paux*dr
    return a, gpot


def star_star_forces_hermite(p, sr, sv, sm, sh):
    """NbodyHermite4::CalculateDirectSmoothedGravForces (NbodyHermite4.cpp:77-140): softened gravity between the stars
    with the mean-h kernel and its time derivative -> (a, adot, gpot).  (wmean carries the reference's single-precision
    powf.)"""
    L = lib()
    ns, nd = sr.shape
    a = np.zeros((ns, nd)); adot = np.zeros((ns, nd)); gpot = np.zeros(ns)
    twopi = 6.283185307179586
    for i in range(ns):
        for j in range(ns):
            if i == j:
                continue
            dr = sr[j] - sr[i]
            dv = sv[j] - sv[i]
            drsqd = float(np.dot(dr, dr))
            drmag = np.sqrt(drsqd)
            invdrmag = 1.0/np.sqrt(drsqd)
            invhmean = 2.0/(sh[i] + sh[j])
            drdt = float(np.dot(dv, dr))*invdrmag
            paux = sm[j]*invhmean*invhmean*L.orc_kernel(C.byref(p), 4, drmag*invhmean)*invdrmag
            wmean = L.orc_kernel(C.byref(p), 0, drmag*invhmean)*float(np.power(np.float32(invhmean), np.float32(nd)))
            gpot[i] += sm[j]*invhmean*L.orc_kernel(C.byref(p), 5, drmag*invhmean)
            a[i] += paux*dr
            adot[i] += paux*dv - 3.0*paux*drdt*invdrmag*dr + 2.0*twopi*sm[j]*drdt*wmean*invdrmag*dr
    return a, adot, gpot


def hermite4_timestep(nm, sa, sadot, sa2dot, sa3dot, sh):
    """NbodyHermite4::Timestep (NbodyHermite4.cpp:542-578): the Aarseth criterion, minimum over the stars"""
    small, big = 1.0e-50, 9.9e50                      # small_number_dp, big_number_dp (Constants.h:89-90)
    out = big
    for i in range(len(sh)):
        asqd, a1sqd = float(np.dot(sa[i], sa[i])), float(np.dot(sadot[i], sadot[i]))
        a2sqd, a3sqd = float(np.dot(sa2dot[i], sa2dot[i])), float(np.dot(sa3dot[i], sa3dot[i]))
        if a1sqd > small and a2sqd > small:
            ts = nm*np.sqrt((np.sqrt(asqd*a2sqd) + a1sqd)/(np.sqrt(a1sqd*a3sqd) + a2sqd))
        elif asqd > small and a2sqd > small:
            ts = nm*np.sqrt(asqd/(a2sqd + small))
        elif asqd > small:
            ts = np.sqrt(sh[i]/(np.sqrt(asqd) + small))
        else:
            ts = big
        out = min(out, float(ts))
    return out


def mainloop(params, setup, steps):
    """`steps` x SphSimulation::MainLoop (SphSimulation.cpp:574-880; SphLeapfrogKDK.cpp:76-125,226-270;
    Integration.cpp:11-75) from the reference's post-setup particle state; full tree rebuild every step.
    -> (state dict, t, next timestep)"""
    p = params_from_dict(params)
    nd = p.ndim
    cm, am, em = (float(params.get("courant_mult", 0.15)), float(params.get("accel_mult", 0.3)),
                  float(params.get("energy_mult", 0.4)))
    energy = params.get("gas_eos", "energy_eqn") == "energy_eqn"
    tdav = params.get("time_dependent_avisc", "none") != "none"
    s = {k: np.array(v, copy=True) for k, v in setup.items()}
    N = len(s["m"])
    for k in ("r", "v", "a", "atree"):
        s[k] = s[k].reshape(N, nd)
    r0, v0, a0 = s["r"].copy(), s["v"].copy(), s["a"].copy()
    u0, dudt0 = s["u"].copy(), s["dudt"].copy()
    timestep = global_timestep(p, s, cm, am, em, energy)
    # star particles: the reference's N-body leapfrog (NbodyLeapfrogKDK.cpp) around the same step
    stars = "star_r" in s
    if stars:
        nm = float(params.get("nbody_mult", 0.1))
        sr, sv, sa = (s[k].reshape(-1, nd).copy() for k in ("star_r", "star_v", "star_a"))
        sm, sh = s["star_m"], s["star_h"]
        sr0, sv0, sa0 = sr.copy(), sv.copy(), sa.copy()
        sgpot = np.zeros(len(sm))
        hermite = params.get("nbody", "hermite4") == "hermite4"      # the reference's default (Parameters.cpp:168)
        if hermite:
            sadot, sa2dot, sa3dot = (s[k].reshape(-1, nd).copy() for k in ("star_adot", "star_a2dot", "star_a3dot"))
            sadot0 = sadot.copy()

        def star_timestep():
            if hermite:
                return hermite4_timestep(nm, sa, sadot, sa2dot, sa3dot, sh)
            amag = np.sqrt((sa*sa).sum(1))
            return float((nm*np.sqrt(sh/(amag + 1.0e-50))).min())
        timestep = min(timestep, star_timestep())
    t = tlast = 0.0
    size = [p.boxmax[k] - p.boxmin[k] for k in range(nd)]
    for _ in range(steps):
        t = t + timestep
        dt = t - tlast
        s["r"] = r0 + v0*dt + 0.5*a0*dt*dt
        s["v"] = v0 + a0*dt
        if stars:
            sr = sr0 + sv0*dt + 0.5*sa0*dt*dt
            sv = sv0 + sa0*dt
            if hermite:                                 # NbodyHermite4::AdvanceParticles (NbodyHermite4.cpp:340-378)
                sr = sr0 + sv0*dt + 0.5*sa0*dt*dt + 0.16666666666666666666666*sadot0*dt*dt*dt
                sv = sv0 + sa0*dt + 0.5*sadot0*dt*dt
            rows = np.concatenate([sr, sm[:, None], sh[:, None]], axis=1)
        if tdav:
            s["alpha"] = s["alpha"] + s["dalphadt"]*timestep
        if energy:
            s["u"] = u0 + dudt0*dt
        for k in range(nd):                               # CheckBoundaries
            walls = dict((side, w) for (kk, w), side in zip(getattr(p, "mirrors", []),
                                                           getattr(p, "mirror_sides", [])) if kk == k)
            lo = s["r"][:, k] < p.boxmin[k]
            if p.periodic[k]:
                s["r"][lo, k] += size[k]; r0[lo, k] += size[k]
            if "lhs" in walls:
                for arr in (s["r"], r0):
                    arr[lo, k] = 2.0*p.boxmin[k] - arr[lo, k]
                for arr in (s["v"], v0, s["a"], a0):
                    arr[lo, k] = -arr[lo, k]
            hi = s["r"][:, k] > p.boxmax[k]
            if p.periodic[k]:
                s["r"][hi, k] -= size[k]; r0[hi, k] -= size[k]
            if "rhs" in walls:
                for arr in (s["r"], r0):
                    arr[hi, k] = 2.0*p.boxmax[k] - arr[hi, k]
                for arr in (s["v"], v0, s["a"], a0):
                    arr[hi, k] = -arr[hi, k]
        dens = density(p, s["r"], s["m"], s["h"], s["u"], stars=rows if stars else None)
        for k in ("h", "rho", "invomega", "zeta", "hfactor", "hrangesqd", "u", "sound", "pressure"):
            s[k] = dens[k]
        if tdav and params.get("time_dependent_avisc") == "cd2010":
            s["alpha"], s["dalphadt"] = cd2010_switch(p, s["r"], s["v"], s["a"], s["m"], dens, s["alpha"])
        if p.self_gravity:
            tr = Tree(p, s["r"], s["v"], s["m"], s["h"], s["atree"], s["gpot"])
            f = tr.hydro_grav_forces(s)
            if stars:                                   # ComputeStarGravForces, then the stars' own forces
                ga, gat, gg = star_grav_forces(p, s["r"], s["h"], sr, sm, sh)
                f["a"] = f["a"] + ga; f["atree"] = f["atree"] + gat; f["gpot"] = f["gpot"] + gg
                sa, sad, sgpot, _ = tr.star_gas_forces(sr, sv, sh)
                if hermite:
                    ssa, ssad, ssg = star_star_forces_hermite(p, sr, sv, sm, sh)
                    sa = sa + ssa; sadot = sad + ssad; sgpot = sgpot + ssg
                    # NbodyHermite4::CorrectionTerms (NbodyHermite4.cpp:388-440); dt^3 in single precision (powf)
                    dts = t - tlast
                    dt3 = float(np.power(np.float32(dts), np.float32(3.0)))
                    invdt = 1.0/dts
                    sa2dot = (-6.0*(sa0 - sa) - dts*(4.0*sadot0 + 2.0*sadot))*invdt*invdt
                    sa3dot = (12.0*(sa0 - sa) + 6.0*dts*(sadot0 + sadot))*invdt*invdt*invdt
                    sr = sr + sa2dot*dt3*dts/24.0 + sa3dot*dt3*dts*dts/120.0
                    sv = sv + sa2dot*dt3/6.0 + sa3dot*dt3*dts/24.0
                else:
                    ssa, ssg = star_star_forces(p, sr, sm, sh)
                    sa = sa + ssa; sgpot = sgpot + ssg
                    sv = sv + 0.5*(sa - sa0)*(t - tlast)    # NbodyLeapfrogKDK::CorrectionTerms
            s["atree"], s["gpot"] = f["atree"], f["gpot"]
        else:
            f = hydro_forces(p, s)
        s["a"], s["dudt"], s["div_v"] = f["a"], f["dudt"], f["div_v"]
        if params.get("time_dependent_avisc", "none") == "mm97":
            # the force pass ADDS its dalphadt to the particle (GradhSphTree.cpp:403) and MainLoop's
            # Sph::ZeroAccelerations (Sph.cpp:126-140) does not clear it: the reference accumulates it
            s["dalphadt"] = s["dalphadt"] + f["dalphadt"]
        nxt = global_timestep(p, s, cm, am, em, energy)
        if stars:
            nxt = min(nxt, star_timestep())
            sr0, sv0, sa0 = sr.copy(), sv.copy(), sa.copy()
            if hermite:
                sadot0 = sadot.copy()
        s["v"] = s["v"] + 0.5*timestep*(s["a"] - a0)          # EndTimestep, part.dt = the step just taken
        r0, v0, a0 = s["r"].copy(), s["v"].copy(), s["a"].copy()
        if energy:
            u = s["u"] + 0.5*(s["dudt"] - dudt0)*timestep
            u = np.where(u <= 0.0, u0 + dudt0*timestep, u)
            s["u"] = u
            u0, dudt0 = u.copy(), s["dudt"].copy()
        tlast = t
        timestep = nxt
    if stars:
        s["star_r"], s["star_v"], s["star_a"], s["star_gpot"] = sr, sv, sa, sgpot
        if hermite:
            s["star_adot"], s["star_a2dot"], s["star_a3dot"] = sadot, sa2dot, sa3dot
    return s, t, timestep



# ---------------------------------------------------------------------------------------------------
# block time steps (Nlevels > 1)
# ---------------------------------------------------------------------------------------------------
INVLOGETWO = 1.44269504088896          # src/Headers/Constants.h:81


def timestep_level(dt, dt_max):
    """ComputeTimestepLevel (InlineFuncs.h:550-558), elementwise"""
    return np.maximum((INVLOGETWO*np.log(dt_max/dt)).astype(np.int64) + 1, 0)


def thermal_all(p, rho, u):
    """GradhSph::ComputeThermalProperties for every particle (GradhSph.cpp:335-347; EOS classes as in
    oracle/gsph_oracle.c thermal()) -> (u, sound, pressure)"""
    g = p.gamma_eos; gm1 = g - 1.0
    u = np.array(u, copy=True)
    if p.eos == 1:
        u = np.full_like(u, p.temp0/gm1/p.mu_bar)
    elif p.eos == 2:
        u = p.temp0*(1.0 + np.power(rho*(1.0/p.rho_bary), gm1))/gm1/p.mu_bar
    elif p.eos == 3:
        u = np.where(rho < p.rho_bary, p.temp0/gm1/p.mu_bar, p.temp0*np.power(rho*(1.0/p.rho_bary), gm1)/gm1/p.mu_bar)
    sound = np.sqrt(g*gm1*u) if p.eos == 0 else np.sqrt(gm1*u)
    return u, sound, gm1*rho*u


def _pow2int(e):
    """(int) pow(2, e) as the reference evaluates it: 0 for negative exponents (a particle that has just moved below
    the deepest level gets nstep = 0 and dt_next = 0 for one step, Simulation.cpp:1990-1992 -- kept)"""
    e = np.asarray(e)
    return np.where(e >= 0, 2**np.clip(e, 0, 62), 0)


def mainloop_block(params, setup, sim, steps, trace=None):
    """`steps` x SphSimulation::MainLoop with hierarchical block time steps (Nlevels > 1): SphSimulation.cpp:574-880,
    SphLeapfrogKDK::AdvanceParticles / CheckTimesteps / EndTimestep (SphLeapfrogKDK.cpp:76-125, 226-339),
    Simulation::ComputeBlockTimesteps (Simulation.cpp:1764-2199), from the reference's post-setup state.
    setup: particle arrays incl. level, levelneib, nstep, nlast, dt, dt_next, tlast, r0, v0, a0, u0, dudt0;
    sim: dict n, nresync, level_max, level_step, dt_max, timestep, t (after the reference's setup).
    No stars, no sinks, lfkdk (integration_step = 1), sph_single_timestep = 0.  -> (state, sim)"""
    p = params_from_dict(params)
    nd = p.ndim
    cm, am, em = (float(params.get("courant_mult", 0.15)), float(params.get("accel_mult", 0.3)),
                  float(params.get("energy_mult", 0.4)))
    energy = params.get("gas_eos", "energy_eqn") == "energy_eqn"
    tdav = params.get("time_dependent_avisc", "none") != "none"
    Nlevels = int(params["Nlevels"])
    ldm = int(params.get("level_diff_max", 1))
    single = int(params.get("sph_single_timestep", 0)) == 1      # every gas particle on the deepest gas level
    s = {k: np.array(v, copy=True) for k, v in setup.items()}
    N = len(s["m"])
    for k in ("r", "v", "a", "atree", "r0", "v0", "a0"):
        s[k] = s[k].reshape(N, nd)
    for k in ("level", "levelneib", "nstep", "nlast"):
        s[k] = s[k].astype(np.int64)
    n, nresync, level_max, level_step = (int(sim[k]) for k in ("n", "nresync", "level_max", "level_step"))
    dt_max, timestep, t = float(sim["dt_max"]), float(sim["timestep"]), float(sim["t"])
    size = [p.boxmax[k] - p.boxmin[k] for k in range(nd)]
    end_ts = np.zeros(N, dtype=bool)
    # star particles (NbodyLeapfrogKDK / NbodyHermite4 with their own block time steps): at least on the deepest level
    # the gas occupies (Simulation.cpp:1858-1872, 2024-2068)
    stars = "star_r" in s
    rows = None
    if stars:
        nm = float(params.get("nbody_mult", 0.1))
        hermite = params.get("nbody", "hermite4") == "hermite4"
        S = {k: s["star_" + k].reshape(-1, nd).copy() for k in ("r", "v", "a")}
        S["m"], S["h"] = s["star_m"].copy(), s["star_h"].copy()
        ns = len(S["m"])
        for k in ("adot", "a2dot", "a3dot"):
            S[k] = s["star_" + k].reshape(-1, nd).copy() if hermite else np.zeros((ns, nd))
        S["r0"], S["v0"], S["a0"], S["adot0"] = S["r"].copy(), S["v"].copy(), S["a"].copy(), S["adot"].copy()
        S["gpot"] = np.zeros(ns)
        S["level"] = s["star_level"].astype(np.int64); S["nstep"] = s["star_nstep"].astype(np.int64)
        S["nlast"] = s["star_nlast"].astype(np.int64); S["tlast"] = s["star_tlast"].astype(np.float64)
        S["end"] = np.zeros(ns, dtype=bool)

        def star_dt(i):
            if hermite:
                return hermite4_timestep(nm, S["a"][i:i + 1], S["adot"][i:i + 1], S["a2dot"][i:i + 1], S["a3dot"][i:i + 1], S["h"][i:i + 1])
            return float(nm*np.sqrt(S["h"][i]/(np.sqrt(float(np.dot(S["a"][i], S["a"][i]))) + 1.0e-50)))

    def passes(active, tr):
        """UpdateAllSphProperties, ZeroAccelerations and the force pass for the active particles; tr = the tree
        BuildTree stocked with the smoothing lengths of the start of the step (cdistsqd!), refreshed here with the
        new ones like tree->UpdateAllHmaxValues (GradhSphTree.cpp:268)"""
        dens = density(p, s["r"], s["m"], s["h"], s["u"], stars=rows)
        for k in ("h", "rho", "invomega", "zeta", "hfactor", "hrangesqd", "u", "sound", "pressure"):
            s[k] = np.where(active, dens[k], s[k])
        if tr is not None:
            tr.update_hmax(s["h"])
        if params.get("time_dependent_avisc") == "cd2010":
            # the switch at the end of ComputeH, for the active particles; neighbours enter as they stand
            al, dal = cd2010_switch(p, s["r"], s["v"], s["a"], s["m"], s, s["alpha"])
            s["alpha"] = np.where(active, al, s["alpha"])
            s["dalphadt"] = np.where(active, dal, s["dalphadt"])
        s["levelneib"] = np.where(active, 0, s["levelneib"])
        for k in ("div_v", "dudt", "gpot"):
            s[k] = np.where(active, 0.0, s[k])
        # nradstep = 1 (the default): MainLoop refreshes the thermal properties of EVERY particle, active or not,
        # before the force pass (SphSimulation.cpp:669-679) -- inactive neighbours enter with the pressure of their
        # predicted u
        s["u"], s["sound"], s["pressure"] = thermal_all(p, s["rho"], s["u"])
        if p.self_gravity:
            f = tr.hydro_grav_forces(s)
            if stars:                                       # ComputeStarGravForces for the active particles
                ga, gat, gg = star_grav_forces(p, s["r"], s["h"], S["r"], S["m"], S["h"])
                f["a"] = f["a"] + ga; f["atree"] = f["atree"] + gat; f["gpot"] = f["gpot"] + gg
            s["atree"] = np.where(active[:, None], f["atree"], s["atree"])
            s["gpot"] = np.where(active, f["gpot"], s["gpot"])
        else:
            f = hydro_forces(p, s)
        s["a"] = np.where(active[:, None], f["a"], s["a"])
        s["dudt"] = np.where(active, f["dudt"], s["dudt"])
        s["div_v"] = np.where(active, f["div_v"], s["div_v"])
        if params.get("time_dependent_avisc", "none") == "mm97":
            s["dalphadt"] = np.where(active, s["dalphadt"] + f["dalphadt"], s["dalphadt"])
        # levelneib: every (active i, neighbour j) pair raises both (GradhSph.cpp:446, GradhSphTree.cpp:376-382)
        off, idx = neighbour_lists(p, s["r"], s["hrangesqd"])
        lev, ln = s["level"], s["levelneib"].copy()
        for i in np.nonzero(active)[0]:
            nb = idx[off[i]:off[i + 1]]
            if len(nb):
                ln[i] = max(ln[i], lev[nb].max())
                np.maximum.at(ln, nb, lev[i])
        s["levelneib"] = ln

    for _ in range(steps):
        n += 1
        t = t + timestep
        # ---- AdvanceParticles
        dn = n - s["nlast"]
        dtp = t - s["tlast"]
        s["r"] = s["r0"] + s["v0"]*dtp[:, None] + 0.5*s["a0"]*dtp[:, None]*dtp[:, None]
        s["v"] = s["v0"] + s["a0"]*dtp[:, None]
        if tdav:
            s["alpha"] = s["alpha"] + s["dalphadt"]*timestep
        if energy:
            s["u"] = s["u0"] + s["dudt0"]*dtp
        active = dn == s["nstep"]
        if stars:                                           # Nbody*::AdvanceParticles
            sdt = (t - S["tlast"])[:, None]
            S["r"] = S["r0"] + S["v0"]*sdt + 0.5*S["a0"]*sdt*sdt
            S["v"] = S["v0"] + S["a0"]*sdt
            if hermite:
                S["r"] = S["r0"] + S["v0"]*sdt + 0.5*S["a0"]*sdt*sdt + 0.16666666666666666666666*S["adot0"]*sdt*sdt*sdt
                S["v"] = S["v0"] + S["a0"]*sdt + 0.5*S["adot0"]*sdt*sdt
            sact = (n - S["nlast"]) == S["nstep"]
            rows = np.concatenate([S["r"], S["m"][:, None], S["h"][:, None]], axis=1)
        # ---- CheckBoundaries (Integration.cpp:11-75): periodic wrap, mirror reflection
        for k in range(nd):
            walls = dict((side, w) for (kk, w), side in zip(getattr(p, "mirrors", []),
                                                           getattr(p, "mirror_sides", [])) if kk == k)
            lo = s["r"][:, k] < p.boxmin[k]
            if p.periodic[k]:
                s["r"][lo, k] += size[k]; s["r0"][lo, k] += size[k]
            if "lhs" in walls:
                for arr in (s["r"], s["r0"]):
                    arr[lo, k] = 2.0*p.boxmin[k] - arr[lo, k]
                for arr in (s["v"], s["v0"], s["a"], s["a0"]):
                    arr[lo, k] = -arr[lo, k]
            hi = s["r"][:, k] > p.boxmax[k]
            if p.periodic[k]:
                s["r"][hi, k] -= size[k]; s["r0"][hi, k] -= size[k]
            if "rhs" in walls:
                for arr in (s["r"], s["r0"]):
                    arr[hi, k] = 2.0*p.boxmax[k] - arr[hi, k]
                for arr in (s["v"], s["v0"], s["a"], s["a0"]):
                    arr[hi, k] = -arr[hi, k]
        # ---- passes, repeated while CheckTimesteps activates more particles
        nredo = 0
        tr = Tree(p, s["r"], s["v"], s["m"], s["h"], s["atree"], s["gpot"]) if p.self_gravity else None
        while True:
            passes(active, tr)
            active = np.zeros(N, dtype=bool)
            dn = n - s["nlast"]
            cand = (dn != s["nstep"]) & (s["levelneib"] - s["level"] > ldm)
            level_new = s["levelneib"] - ldm
            nnewstep = _pow2int(level_step - level_new)
            ok = cand & (nnewstep > 0) & (dn % np.maximum(nnewstep, 1) == 0)
            s["nstep"] = np.where(ok & (dn > 0), dn, s["nstep"])
            s["level"] = np.where(ok, level_new, s["level"])
            active = ok
            if not ok.any():
                break
            nredo += 1
        # ---- the N-body block of MainLoop (SphSimulation.cpp:771-812): forces and correction of the ACTIVE stars
        if stars and sact.any():
            ga, gad, gg, _ = tr.star_gas_forces(S["r"], S["v"], S["h"])
            if hermite:
                ssa, ssad, ssg = star_star_forces_hermite(p, S["r"], S["v"], S["m"], S["h"])
            else:
                ssa, ssg = star_star_forces(p, S["r"], S["m"], S["h"]); ssad = np.zeros_like(ssa); gad = np.zeros_like(ga)
            S["a"] = np.where(sact[:, None], ga + ssa, S["a"]); S["gpot"] = np.where(sact, gg + ssg, S["gpot"])
            S["adot"] = np.where(sact[:, None], gad + ssad, S["adot"])
            for i in np.nonzero(sact)[0]:
                dts = t - S["tlast"][i]
                if hermite:
                    dt3 = float(np.power(np.float32(dts), np.float32(3.0))); invdt = 1.0/dts
                    S["a2dot"][i] = (-6.0*(S["a0"][i] - S["a"][i]) - dts*(4.0*S["adot0"][i] + 2.0*S["adot"][i]))*invdt*invdt
                    S["a3dot"][i] = (12.0*(S["a0"][i] - S["a"][i]) + 6.0*dts*(S["adot0"][i] + S["adot"][i]))*invdt*invdt*invdt
                    S["r"][i] = S["r"][i] + S["a2dot"][i]*dt3*dts/24.0 + S["a3dot"][i]*dt3*dts*dts/120.0
                    S["v"][i] = S["v"][i] + S["a2dot"][i]*dt3/6.0 + S["a3dot"][i]*dt3*dts/24.0
                else:
                    S["v"][i] = S["v"][i] + 0.5*(S["a"][i] - S["a0"][i])*dts
        # ---- ComputeBlockTimesteps
        ts = particle_timesteps(p, s, cm, am, em, energy)
        if n == nresync:
            n = 0
            s["dt_next"] = ts.copy()
            timestep = float(ts.min())
            if stars:
                sts = np.array([star_dt(i) for i in range(ns)])
                timestep = min(timestep, float(sts.min()))
            level_max = Nlevels - 1
            level_step = level_max
            dt_max = timestep*2.0**level_max
            if stars:
                lmh = min(int(timestep_level(np.array([float(ts.min())]), dt_max)[0]), level_max)
                S["level"] = np.maximum(np.minimum(timestep_level(sts, dt_max), level_max), lmh)
                S["nlast"][:] = n; S["nstep"] = _pow2int(level_step - S["level"]); S["tlast"][:] = t
                S["end"][:] = True
            lev = np.minimum(timestep_level(ts, dt_max), level_max)
            if single:                                      # Simulation.cpp:1889-1902
                lev = np.full(N, min(int(timestep_level(np.array([float(ts.min())]), dt_max)[0]), level_max), dtype=np.int64)
            s["level"] = lev; s["levelneib"] = lev.copy()
            s["nstep"] = _pow2int(level_step - lev)
            s["nlast"] = np.full(N, n, dtype=np.int64)
            s["dt_next"] = s["nstep"]*timestep
            end_ts[:] = True
            nresync = 2**level_step
            timestep = dt_max/float(nresync)
        else:
            level_max_old = level_max
            atend = (n - s["nlast"]) == s["nstep"]
            reduced = atend & (s["nstep"] != _pow2int(level_step - s["level"]))
            natural = atend & ~reduced
            lev_dt = timestep_level(ts, dt_max)
            level = np.maximum(lev_dt, s["levelneib"] - ldm)
            # Saitoh & Makino reduced steps
            newl = np.maximum(s["level"], level)
            s["level"] = np.where(reduced, newl, s["level"])
            s["levelneib"] = np.where(reduced, newl, s["levelneib"])
            # natural end of step: up one level if synchronised, down any number
            last = s["level"]
            up = natural & (level < last) & (last > 1) & (n % (2*s["nstep"]) == 0)
            down = natural & (level > last)
            s["level"] = np.where(up, last - 1, np.where(down, level, last))
            s["levelneib"] = np.where(natural, level, s["levelneib"])
            s["nlast"] = np.where(atend, n, s["nlast"])
            s["nstep"] = np.where(atend, _pow2int(level_step - s["level"]), s["nstep"])
            s["dt_next"] = np.where(atend, s["nstep"]*timestep, s["dt_next"])
            end_ts |= atend
            level_max = int(s["level"].max())
            if single:                                      # Simulation.cpp:2092-2099 (before the integer time is rescaled)
                s["level"] = np.where(s["nlast"] == n, level_max, s["level"])
            if stars:
                lmh = level_max                                 # level_max_hydro
                for i in range(ns):
                    if n - S["nlast"][i] == S["nstep"][i]:
                        nstep_i, last = int(S["nstep"][i]), int(S["level"][i])
                        lv = max(int(timestep_level(np.array([star_dt(i)]), dt_max)[0]), lmh)
                        if lv < last and lv > lmh and last > 1 and n % (2*nstep_i) == 0:
                            S["level"][i] = last - 1
                        elif lv > last:
                            S["level"][i] = lv
                        S["nlast"][i] = n; S["nstep"][i] = _pow2int(level_step - S["level"][i]); S["tlast"][i] = t
                        S["end"][i] = True
                level_max = max(level_max, int(S["level"].max()))
            istep = 2**(level_step - level_max_old + 1)
            if level_max > level_max_old:
                nfactor = 2**(level_max - level_max_old)
                n *= nfactor
                s["nstep"] = s["nstep"]*nfactor; s["nlast"] = s["nlast"]*nfactor
                if stars:
                    S["nstep"] = S["nstep"]*nfactor; S["nlast"] = S["nlast"]*nfactor
            elif level_max <= level_max_old - 1 and level_max_old > 1 and n % istep == 0:
                level_max = level_max_old - 1
                nfactor = 2**(level_max_old - level_max)
                n //= nfactor
                s["nlast"] = s["nlast"]//nfactor; s["nstep"] = s["nstep"]//nfactor
                if stars:
                    S["nstep"] = S["nstep"]//nfactor; S["nlast"] = S["nlast"]//nfactor
            else:
                level_max = level_max_old
            level_step = level_max
            nresync = 2**level_step
            timestep = dt_max/float(nresync)
            s["nstep"] = np.where(s["nlast"] == n, _pow2int(level_step - s["level"]), s["nstep"])
            if stars:
                S["nstep"] = np.where(S["nlast"] == n, _pow2int(level_step - S["level"]), S["nstep"])
        # ---- EndTimestep
        e = end_ts
        s["v"] = np.where(e[:, None], s["v"] + 0.5*s["dt"][:, None]*(s["a"] - s["a0"]), s["v"])
        for k0, k1 in (("r0", "r"), ("v0", "v"), ("a0", "a")):
            s[k0] = np.where(e[:, None], s[k1], s[k0])
        if energy:
            u = s["u"] + 0.5*(s["dudt"] - s["dudt0"])*s["dt"]
            u = np.where(u <= 0.0, s["u0"] + s["dudt0"]*s["dt"], u)
            s["u"] = np.where(e, u, s["u"])
            s["u0"] = np.where(e, s["u"], s["u0"])
            s["dudt0"] = np.where(e, s["dudt"], s["dudt0"])
        s["nlast"] = np.where(e, n, s["nlast"])
        s["tlast"] = np.where(e, t, s["tlast"])
        s["dt"] = np.where(e, s["dt_next"], s["dt"])
        s["dt_next"] = np.where(e, 0.0, s["dt_next"])
        end_ts[:] = False
        if stars:                                           # Nbody*::EndTimestep
            e = S["end"]
            for k0, k1 in (("r0", "r"), ("v0", "v"), ("a0", "a"), ("adot0", "adot")):
                S[k0] = np.where(e[:, None], S[k1], S[k0])
            S["nlast"] = np.where(e, n, S["nlast"]); S["tlast"] = np.where(e, t, S["tlast"])
            S["end"][:] = False
        if trace is not None:
            trace.append(dict(n=n, nresync=nresync, level_max=level_max, timestep=timestep, t=t, nredo=nredo))
    sim_out = dict(n=n, nresync=nresync, level_max=level_max, level_step=level_step, dt_max=dt_max, timestep=timestep, t=t)
    if stars:
        for k in ("r", "v", "a", "adot", "a2dot", "a3dot", "gpot", "level", "nstep", "nlast", "tlast"):
            s["star_" + k] = S[k]
    return s, sim_out
