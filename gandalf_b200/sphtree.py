"""Host-side mirror of the reference operator surface for the neighbour path.

``GpuSphTree`` carries the method names of ``SphNeighbourSearch<ndim>`` /
``GradhSphTree<ndim,GradhSphParticle>`` (src/Headers/SphNeighbourSearch.h:77-94,174-227,
src/Headers/NeighbourSearch.h:66-137) and forwards each to one libgsph C-ABI call, so the parity
tests read like calls on the reference's own ``sphneib`` pointer.  The C++ twin used for the real
drop-in is ``gandalf_b200/host/GpuSphTree.h``.

``ParticleSet`` stands in for the particle array owned by ``GradhSph::sphdata``: it can be laid out
as SoA (one contiguous array per field) or as an AoS record array with the field order of
``GradhSphParticle<ndim>`` (src/Headers/Particle.h:134-368) to exercise the strided-view upload.
"""
import ctypes as C

import numpy as np

from . import _capi
from .params import Parameters, config_from_params

_DBL = _capi.VEC_FIELDS + _capi.SCALAR_FIELDS
_INT = _capi.INT_FIELDS


class ParticleSet(object):
    """Host particle data.  ``data[name]`` is an (N,) or (N, ndim) numpy view; the views are
    strided when ``layout='aos'``."""

    def __init__(self, N, ndim, layout="soa", fields=None, alloc=None):
        """fields: restrict a SoA set to these names (absent ones are passed as NULL = reference
        constructor defaults); alloc(shape, dtype) -> ndarray lets the caller supply e.g. pinned memory."""
        self.N, self.ndim, self.layout = int(N), int(ndim), layout
        self.data = {}
        alloc = alloc or (lambda shape, dtype: np.zeros(shape, dtype=dtype))
        if layout == "aos":
            # field order follows Particle -> SphParticle -> GradhSphParticle (Particle.h:134-368)
            dt = [("flags", "<u4"), ("ptype", "<i4"), ("iorig", "<i4"), ("sinkid", "<i4"),
                  ("levelneib", "<i4"), ("nstep", "<i4"), ("nlast", "<i4"), ("level", "<i4"),
                  ("r", "<f8", (ndim,)), ("v", "<f8", (ndim,)), ("a", "<f8", (ndim,)),
                  ("atree", "<f8", (ndim,)), ("r0", "<f8", (ndim,)), ("v0", "<f8", (ndim,)),
                  ("a0", "<f8", (ndim,)), ("m", "<f8"), ("h", "<f8"), ("h_dust", "<f8"),
                  ("hrangesqd", "<f8"), ("hfactor", "<f8"), ("sound", "<f8"), ("rho", "<f8"),
                  ("pressure", "<f8"), ("u", "<f8"), ("u0", "<f8"), ("dudt0", "<f8"), ("dudt", "<f8"),
                  ("gpot", "<f8"), ("gpot_hydro", "<f8"), ("dt", "<f8"), ("dt_next", "<f8"),
                  ("tlast", "<f8"), ("ionfrac", "<f8"), ("Xion", "<f8"), ("ueq", "<f8"),
                  ("dt_therm", "<f8"), ("vsig_max", "<f8"), ("rad_pres", "<f8", (ndim,)),
                  ("ionstate", "<i4"), ("_pad", "<i4"), ("div_v", "<f8"), ("alpha", "<f8"),
                  ("dalphadt", "<f8"), ("invomega", "<f8"), ("zeta", "<f8")]
            self.records = np.zeros(self.N, dtype=np.dtype(dt))
            for n in _DBL + _INT:
                self.data[n] = self.records[n]
        else:
            self.records = None
            want = (lambda n: True) if fields is None else (lambda n: n in fields)
            for n in _capi.VEC_FIELDS:
                # component-major storage: every component is one contiguous array (stride 8)
                if want(n):
                    self.data[n] = alloc((ndim, self.N), np.float64).T
            for n in _capi.SCALAR_FIELDS:
                if want(n):
                    self.data[n] = alloc((self.N,), np.float64)
            for n in _INT:
                if want(n):
                    self.data[n] = alloc((self.N,), np.uint32 if n == "flags" else np.int32)
        for n, v in (("h", 1.0), ("invomega", 1.0), ("sinkid", -1), ("flags", _capi.FLAG_ACTIVE)):
            if n in self.data:
                self.data[n][...] = v
        if "iorig" in self.data:
            self.data["iorig"][...] = np.arange(self.N)

    def __getitem__(self, k):
        return self.data[k]

    def set(self, **fields):
        for k, v in fields.items():
            self.data[k][...] = v
        return self

    # ---- the reference's array names (Simulation::ImportArray, Simulation.cpp:1442-1612; SphSnapshot::ExtractArray,
    #      SphSnapshot.cpp:529-600): "x" "y" "z" "vx" "vy" "vz" "ax" "ay" "az" "m" "h" "rho" "u" "dudt" "gpot" "iorig"
    _ARRAY_NAMES = {"x": ("r", 0), "y": ("r", 1), "z": ("r", 2), "vx": ("v", 0), "vy": ("v", 1), "vz": ("v", 2),
                    "ax": ("a", 0), "ay": ("a", 1), "az": ("a", 2), "m": ("m", None), "h": ("h", None),
                    "rho": ("rho", None), "u": ("u", None), "dudt": ("dudt", None), "gpot": ("gpot", None),
                    "iorig": ("iorig", None)}

    def _array(self, name):
        if name not in self._ARRAY_NAMES:
            raise ValueError("Warning: the selected array: %s has not been recognized" % name)
        f, k = self._ARRAY_NAMES[name]
        if k is not None and k >= self.ndim:
            raise ValueError("Error: is not possible to use a dimension higher than %d for %s" % (self.ndim, name))
        if f not in self.data:
            raise KeyError("the particle set does not hold %s" % f)
        return self.data[f] if k is None else self.data[f][:, k]

    def ImportArray(self, values, name):
        """Simulation::ImportArray for the gas particles: one quantity by the reference's name"""
        a = self._array(name)
        values = np.asarray(values)
        if values.shape != (self.N,):
            raise ValueError("ImportArray: %d values for %d particles" % (values.size, self.N))
        a[...] = values

    def ExtractArray(self, name):
        """SphSnapshot::ExtractArray for the gas particles: a contiguous copy of one quantity"""
        return np.ascontiguousarray(self._array(name)).copy()

    def view(self, only=None):
        """gsph_particle_view over the current arrays (never cached: the reference may reallocate
        its array between calls, GradhSph.cpp:83-111)."""
        pv = _capi.ParticleView()
        for n in _capi.VEC_FIELDS:
            if (only is not None and n not in only) or n not in self.data:
                continue
            a = self.data[n]
            for k in range(self.ndim):
                f = getattr(pv, n)[k]
                f.base = a.ctypes.data + k * a.strides[1]
                f.stride = a.strides[0]
        for n in _capi.SCALAR_FIELDS + _INT:
            if (only is not None and n not in only) or n not in self.data:
                continue
            a = self.data[n]
            f = getattr(pv, n)
            f.base = a.ctypes.data
            f.stride = a.strides[0]
        return pv


class GpuSphTree(object):
    """neib_search = gpu : the B200 implementation behind the SphNeighbourSearch surface."""

    def __init__(self, params, device=0):
        self.params = Parameters(params)
        self._L = _capi.lib()
        self.cfg = config_from_params(self.params, device=device)
        self._h = C.c_void_p()
        rc = self._L.gsph_create(C.byref(self.cfg), C.byref(self._h))
        if rc:
            raise _capi.GsphError(rc, self._L.gsph_last_error(None).decode())
        self.neibcheck = False
        self.Nsteps_since_build = 0

    def close(self):
        if getattr(self, "_h", None):
            self._L.gsph_destroy(self._h)
            self._h = None

    __del__ = close

    def _check(self, rc):
        if rc:
            raise _capi.GsphError(rc, self._L.gsph_last_error(self._h).decode())

    # ---- data movement (what the C++ wrapper does around every virtual call) --------------------
    def Upload(self, parts, only=None):
        pv = parts.view(only)
        self._check(self._L.gsph_upload(self._h, C.byref(pv), parts.N))

    def Download(self, parts, mask=_capi.F_ALL, only=None, wait=True):
        """wait=False queues the device->host copies on the library's copy stream and returns
        (gsph_download_async); they overlap the kernels issued afterwards.  Synchronize() or the
        next blocking Download() completes them."""
        pv = parts.view(only)
        fn = self._L.gsph_download if wait else self._L.gsph_download_async
        self._check(fn(self._h, C.byref(pv), mask))

    def UploadRange(self, parts, lo, hi, only=None):
        """ship only the host-order particles [lo, hi) (gsph_upload_range): the share of this rank"""
        pv = parts.view(only)
        self._check(self._L.gsph_upload_range(self._h, C.byref(pv), parts.N, int(lo), int(hi)))

    def DownloadRange(self, parts, lo, hi, mask=_capi.F_ALL, only=None, wait=True):
        pv = parts.view(only)
        self._check(self._L.gsph_download_range(self._h, C.byref(pv), mask, int(lo), int(hi), 1 if wait else 0))

    def DeviceArray(self, name, staging=False):
        """-> (device pointer, capacity in elements, element size) of one particle array, tree order
        or host-order staging (gsph_device_array)"""
        ptr, cap, es = C.c_void_p(), C.c_int64(), C.c_int32()
        self._check(self._L.gsph_device_array(self._h, 1 if staging else 0, name.encode(), C.byref(ptr),
                                              C.byref(cap), C.byref(es)))
        return ptr.value, cap.value, es.value

    # ---- NeighbourSearch / SphNeighbourSearch surface -----------------------------------------------
    def BuildTree(self, rebuild_tree, n, ntreebuildstep, ntreestockstep, timestep, parts=None):
        """NeighbourSearch::BuildTree (HydroTree.cpp:310-372): rebuild | restock | extrapolate."""
        if parts is not None:
            self.Upload(parts)
        if n % ntreebuildstep == 0 or rebuild_tree:
            mode = 0
        elif n % ntreestockstep == 0:
            mode = 1
        else:
            mode = 2
        self._check(self._L.gsph_build_tree(self._h, mode, float(timestep)))

    def BuildTreeMode(self, mode, timestep=0.0):
        """gsph_build_tree with an explicit mode: 0 rebuild, 1 restock, 2 extrapolate, 3 restock keeping the MAC data"""
        self._check(self._L.gsph_build_tree(self._h, int(mode), float(timestep)))

    def BuildGhostTree(self, *a, **kw):
        """Periodic images are taken on the fly in the walks: nothing to build (HydroTree.cpp:382-441)."""

    def SearchBoundaryGhostParticles(self, *a, **kw):
        """No ghost particles are materialised (HydroTree.cpp:495-543 becomes a no-op)."""

    def UpdateActiveParticleCounters(self):
        self._check(self._L.gsph_update_active_counters(self._h))

    def UpdateAllSphProperties(self):
        self._check(self._L.gsph_update_properties(self._h))

    UpdateAllProperties = UpdateAllSphProperties

    def ZeroAccelerations(self):
        """Sph::ZeroAccelerations (Sph.cpp:126-140), called between the two passes by MainLoop."""
        self._check(self._L.gsph_zero_accelerations(self._h))

    def UpdateAllSphHydroForces(self):
        self._check(self._L.gsph_update_hydro_forces(self._h))

    def UpdateAllSphForces(self):
        self._check(self._L.gsph_update_all_forces(self._h, None, 0))

    def UpdateAllSphForcesBegin(self):
        """neighbour half of UpdateAllSphForces; afterwards dudt, div_v, dalphadt, levelneib are final and
        Download(..., wait=False) of them overlaps UpdateAllSphForcesFinish (tree gravity + stars)"""
        self._check(self._L.gsph_update_all_forces_begin(self._h))

    def UpdateAllSphForcesFinish(self):
        self._check(self._L.gsph_update_all_forces_finish(self._h))

    # ---- device-resident main loop (SphSimulation::MainLoop with lfkdk, Nlevels = 1) --------------
    def IntegratorBegin(self, parts=None, compute_forces=True, t=0.0):
        """tail of PostInitialConditionsSetup (SphSimulation.cpp:390-563) on the particles of the last
        Upload (or `parts`): optional first density + force pass, r0/v0/a0/u0, first global time step"""
        from .params import integrator_config_from_params
        if parts is not None:
            self.Upload(parts)
        ic = integrator_config_from_params(self.params)
        self._check(self._L.gsph_integrator_begin(self._h, C.byref(ic), int(compute_forces), float(t)))

    def IntegratorSetStars(self, r=None, v=None, m=None, h=None, a=None, adot=None, a2dot=None, a3dot=None):
        """star particles of the device-resident loop: the reference's N-body integrator -- nbody = hermite4 (its
        default, NbodyHermite4.cpp) or lfkdk -- runs on the host around the device passes; call before
        IntegratorBegin; r=None removes them.  a (and adot, a2dot, a3dot for hermite4) are only read by
        IntegratorBegin(compute_forces=False)"""
        if r is None or len(m) == 0:
            self._check(self._L.gsph_integrator_set_stars(self._h, None, 0, 0.1))
            self._nstar = 0
            return
        kind = self.params["nbody"]
        if kind not in ("lfkdk", "hermite4"):
            raise ValueError("nbody = %s is not on the accelerated path (lfkdk and hermite4 only)" % kind)
        nd, ns = self.cfg.ndim, len(m)
        z = np.zeros((ns, nd))
        vec = lambda x: z if x is None else np.asarray(x, dtype=np.float64).reshape(ns, nd)
        cols = [np.asarray(r, dtype=np.float64).reshape(ns, nd), vec(v), vec(a)]
        if kind == "hermite4":
            cols += [vec(adot), vec(a2dot), vec(a3dot)]
        rows = np.ascontiguousarray(np.concatenate(cols + [np.asarray(m, dtype=np.float64)[:, None],
                                                           np.asarray(h, dtype=np.float64)[:, None]], axis=1))
        fn = self._L.gsph_integrator_set_stars_hermite4 if kind == "hermite4" else self._L.gsph_integrator_set_stars
        self._check(fn(self._h, rows.ctypes.data, ns, float(self.params["nbody_mult"])))
        self._nstar = ns

    def IntegratorGetStars(self):
        """-> dict r, v, a, m, h, gpot of the stars after the last step (+ adot, a2dot, a3dot with nbody = hermite4)"""
        nd, ns = self.cfg.ndim, getattr(self, "_nstar", 0)
        rows = np.zeros((ns, 3*nd + 3))
        self._check(self._L.gsph_integrator_get_stars(self._h, rows.ctypes.data, ns))
        out = {"r": rows[:, :nd].copy(), "v": rows[:, nd:2*nd].copy(), "a": rows[:, 2*nd:3*nd].copy(),
               "m": rows[:, 3*nd].copy(), "h": rows[:, 3*nd + 1].copy(), "gpot": rows[:, 3*nd + 2].copy()}
        if ns > 0 and self.params["nbody"] == "hermite4":
            for k in ("adot", "a2dot", "a3dot"):
                out[k] = np.zeros((ns, nd))
            self._check(self._L.gsph_integrator_get_star_derivatives(self._h, out["adot"].ctypes.data,
                                                                     out["a2dot"].ctypes.data, out["a3dot"].ctypes.data))
        if ns > 0 and int(self.params["Nlevels"]) > 1:      # block time steps: the stars' own levels
            for k in ("level", "nstep", "nlast"):
                out[k] = np.zeros(ns, dtype=np.int64)
            out["tlast"] = np.zeros(ns)
            self._check(self._L.gsph_integrator_get_star_levels(self._h, out["level"].ctypes.data, out["nstep"].ctypes.data,
                                                                out["nlast"].ctypes.data, out["tlast"].ctypes.data))
        return out

    def IntegratorAdvance(self):
        """first half of a MainLoop -> (tree mode 0 rebuild / 1 restock / 2 extrapolate, time step)"""
        mode, ts = C.c_int(), C.c_double()
        self._check(self._L.gsph_integrator_advance(self._h, C.byref(mode), C.byref(ts)))
        return mode.value, ts.value

    def IntegratorFinish(self):
        self._check(self._L.gsph_integrator_finish(self._h))

    def IntegratorCheckTimesteps(self):
        """SphLeapfrogKDK::CheckTimesteps after the passes of a block-time-step MainLoop -> number of particles
        whose step was cut short (> 0: the passes have to run again for them)"""
        woken = C.c_int()
        self._check(self._L.gsph_integrator_check_timesteps(self._h, C.byref(woken)))
        return woken.value

    def UpdateThermalProperties(self):
        """ComputeThermalProperties of every particle (MainLoop with nradstep = 1, SphSimulation.cpp:669-679)"""
        self._check(self._L.gsph_update_thermal_properties(self._h))

    def MainLoop(self, nsteps=1):
        """nsteps x SphSimulation::MainLoop (SphSimulation.cpp:574-880) without leaving the device"""
        self._check(self._L.gsph_integrator_step(self._h, int(nsteps)))

    def IntegratorState(self):
        """-> (t, timestep of the next step, steps done)"""
        t, ts, n = C.c_double(), C.c_double(), C.c_int64()
        self._check(self._L.gsph_integrator_state(self._h, C.byref(t), C.byref(ts), C.byref(n)))
        return t.value, ts.value, n.value

    def IntegratorBlockState(self, arrays=False):
        """block time steps (Nlevels > 1): dict n, nresync, level_max, level_step, dt_max of the hierarchy
        (Simulation::ComputeBlockTimesteps) and, with arrays=True, nstep, nlast, dt, tlast of every particle"""
        st = _capi.BlockState()
        out = {}
        if arrays:
            N = self.Stats().Nhydro
            out = {"nstep": np.zeros(N, dtype=np.int32), "nlast": np.zeros(N, dtype=np.int32), "dt": np.zeros(N), "tlast": np.zeros(N)}
            self._check(self._L.gsph_integrator_block_state(self._h, C.byref(st), out["nstep"].ctypes.data, out["nlast"].ctypes.data,
                                                            out["dt"].ctypes.data, out["tlast"].ctypes.data))
        else:
            self._check(self._L.gsph_integrator_block_state(self._h, C.byref(st), None, None, None, None))
        out.update(n=st.n, nresync=st.nresync, level_max=st.level_max, level_step=st.level_step, dt_max=st.dt_max)
        return out

    def IntegratorStartState(self):
        """start-of-step state of every particle (r0, v0, a0, u0, dudt0 of SphLeapfrogKDK; dt_next with block steps)"""
        N, nd = self.Stats().Nhydro, self.cfg.ndim
        out = {"r0": np.zeros((nd, N)), "v0": np.zeros((nd, N)), "a0": np.zeros((nd, N)), "u0": np.zeros(N), "dudt0": np.zeros(N)}
        blk = int(self.params["Nlevels"]) > 1
        if blk:
            out["dt_next"] = np.zeros(N)
        self._check(self._L.gsph_integrator_start_state(self._h, out["r0"].ctypes.data, out["v0"].ctypes.data, out["a0"].ctypes.data,
                                                        out["u0"].ctypes.data, out["dudt0"].ctypes.data,
                                                        out["dt_next"].ctypes.data if blk else None))
        for k in ("r0", "v0", "a0"):
            out[k] = np.ascontiguousarray(out[k].T)
        return out

    # ---- snapshots in the reference's "su" wire format (gandalf_b200/snapshot.py) ------------------
    def WriteSnapshot(self, path, alloc=None, **header):
        """SEREN-unformatted snapshot of the particles on the device (SimulationIO.hpp:1655-2262): only
        the seven arrays of the format come home, as contiguous SoA copies.  In a device-resident run the
        header carries the loop's t and step count."""
        from . import snapshot
        N = self.Stats().Nhydro
        fields = {"r", "v", "m", "h", "rho", "u", "iorig", "ptype", "flags"}
        ps = ParticleSet(N, self.cfg.ndim, "soa", fields=fields, alloc=alloc)
        self.Download(ps, only=tuple(fields))
        hdr = dict(h_fac=self.cfg.h_fac)
        try:
            t, _, n = self.IntegratorState()
            hdr.update(t=t, Nsteps=n)
        except _capi.GsphError:
            pass
        if getattr(self, "_nstar", 0) > 0 and "stars" not in header:
            # the reference always writes nbody->stardata (WriteSerenUnformSnapshotFile, SimulationIO.hpp:2212-2250):
            # without them a restart from our own snapshot would lose the star particles
            st = self.IntegratorGetStars()
            hdr["stars"] = dict(r=st["r"], v=st["v"], m=st["m"], h=st["h"])
        hdr.update(header)
        snapshot.write_su_from_particle_set(path, ps, **hdr)
        return ps

    @staticmethod
    def ReadSnapshot(path, alloc=None):
        """-> (ParticleSet, header) from a SEREN-unformatted snapshot (the reference's or ours); star particles,
        if the file has any, are header["stars"] = dict r, v, m, h, radius (feed them to IntegratorSetStars)"""
        from . import snapshot
        return snapshot.particle_set_from_su(path, alloc=alloc)

    def UpdateAllStarGasForces(self, rstar, vstar, hstar):
        """NeighbourSearch::UpdateAllStarGasForces (HydroTree.cpp:552-654): gravity of the gas on the stars
        -> (a, adot, gpot, counts[Nstar][neighbours, direct, cells]); uses the tree / particles of the last pass"""
        nd = self.cfg.ndim
        ns = len(hstar)
        rows = np.ascontiguousarray(np.concatenate([np.asarray(rstar, dtype=np.float64).reshape(ns, nd),
                                                    np.asarray(vstar, dtype=np.float64).reshape(ns, nd),
                                                    np.asarray(hstar, dtype=np.float64)[:, None]], axis=1))
        a = np.zeros((ns, nd)); adot = np.zeros((ns, nd)); gpot = np.zeros(ns)
        counts = np.zeros((ns, 3), dtype=np.int32)
        self._check(self._L.gsph_update_star_gas_forces(self._h, rows.ctypes.data, ns, a.ctypes.data, adot.ctypes.data,
                                                        gpot.ctypes.data, counts.ctypes.data))
        return a, adot, gpot, counts

    def SetStars(self, r=None, m=None, h=None):
        """nbody->nbodydata for GradhSph::ComputeStarGravForces (GradhSph.cpp:699-743): star positions
        (Nstar, ndim), masses and smoothing lengths; applied by both force passes until cleared
        with SetStars()."""
        if r is None or len(m) == 0:
            self._check(self._L.gsph_set_stars(self._h, None, 0))
            return
        r = np.asarray(r, dtype=np.float64).reshape(len(m), -1)
        stars = np.ascontiguousarray(np.concatenate([r, np.asarray(m, dtype=np.float64)[:, None],
                                                     np.asarray(h, dtype=np.float64)[:, None]], axis=1))
        self._check(self._L.gsph_set_stars(self._h, stars.ctypes.data, len(m)))

    def GetMaximumSmoothingLength(self):
        v = C.c_double()
        self._check(self._L.gsph_get_hmax(self._h, C.byref(v)))
        return v.value

    def ToggleNeighbourCheck(self, do):
        self.neibcheck = bool(do)

    # ---- multi-GPU plumbing (include/gsph.h "multi-GPU plumbing") ------------------------------------
    def SetActiveSlice(self, lo, hi):
        self._check(self._L.gsph_set_active_slice(self._h, int(lo), int(hi)))

    def DensityPass(self):
        self._check(self._L.gsph_density_pass(self._h))

    def RefreshHmax(self):
        self._check(self._L.gsph_refresh_hmax(self._h))

    def TreeShardRange(self, rank, world):
        """tree positions [lo, hi) of the rank-th node of level log2(world) (gsph_tree_shard_range)"""
        lo, hi = C.c_int64(), C.c_int64()
        self._check(self._L.gsph_tree_shard_range(self._h, rank, world, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def DeriveDensityFields(self, lo, hi):
        self._check(self._L.gsph_derive_density_fields(self._h, int(lo), int(hi)))

    def BuildTreeShardBegin(self, rank, world):
        self._check(self._L.gsph_build_tree_shard_begin(self._h, rank, world))

    def BuildTreeShardFinish(self):
        self._check(self._L.gsph_build_tree_shard_finish(self._h))

    def DeviceField(self, name):
        """-> (device pointer, capacity in elements) of a tree-order array (float64; "perm": int32)."""
        ptr, cap = C.c_void_p(), C.c_int64()
        self._check(self._L.gsph_device_field(self._h, _capi.DEVICE_FIELDS.index(name), C.byref(ptr), C.byref(cap)))
        return ptr.value, cap.value

    def TimerStart(self):
        self._check(self._L.gsph_timer_start(self._h))

    def TimerStop(self):
        ms = C.c_float()
        self._check(self._L.gsph_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def Synchronize(self):
        self._check(self._L.gsph_synchronize(self._h))

    # ---- parity / inspection ---------------------------------------------------------------------
    def NeighbourLists(self):
        """-> (offsets int64 [N+1], idx int32) rows in host order, entries ascending host index."""
        N = self.Stats().Nhydro
        counts = np.zeros(N, dtype=np.int32)
        self._check(self._L.gsph_neighbour_counts(self._h, counts.ctypes.data))
        off = np.zeros(N + 1, dtype=np.int64)
        np.cumsum(counts, out=off[1:])
        idx = np.zeros(max(int(off[-1]), 1), dtype=np.int32)
        self._check(self._L.gsph_neighbour_lists(self._h, off.ctypes.data, idx.ctypes.data))
        return off, idx[:off[-1]]

    def TreeInfo(self):
        t = _capi.TreeInfo()
        self._check(self._L.gsph_tree_info_get(self._h, C.byref(t)))
        return t

    def TreeLeafOf(self):
        out = np.zeros(self.Stats().Nhydro, dtype=np.int32)
        self._check(self._L.gsph_tree_leaf_of(self._h, out.ctypes.data))
        return out

    def TreeCells(self):
        t = self.TreeInfo()
        nd, nc = self.cfg.ndim, t.Ncell
        out = dict(box=np.zeros((nc, 4 * nd)), com=np.zeros((nc, 2 * nd)), scalars=np.zeros((nc, 5)),
                   q=np.zeros((nc, 5)), N=np.zeros(nc, dtype=np.int32))
        self._check(self._L.gsph_tree_cells(self._h, out["box"].ctypes.data, out["com"].ctypes.data,
                                            out["scalars"].ctypes.data, out["q"].ctypes.data,
                                            out["N"].ctypes.data))
        return out

    def Stats(self):
        s = _capi.Stats()
        self._check(self._L.gsph_get_stats(self._h, C.byref(s)))
        return s


def kernel_table(ndim, kernel, which):
    out = np.zeros(1000)
    rc = _capi.lib().gsph_kernel_table(ndim, _capi.KERNEL[kernel], which, out.ctypes.data)
    if rc:
        raise _capi.GsphError(rc, "gsph_kernel_table")
    return out


def measure_fp_peak(device=0):
    a, b = C.c_double(), C.c_double()
    rc = _capi.lib().gsph_measure_fp_peak(device, C.byref(a), C.byref(b))
    if rc:
        raise _capi.GsphError(rc, "gsph_measure_fp_peak")
    return a.value, b.value
