"""ctypes binding of the libgsph C ABI (include/gsph.h).  Thin: structures, prototypes, error
translation.  Loading fails loudly when the CUDA library has not been built; there is no CPU path."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgsph.so")

GSPH_ABI_VERSION = 1
NTYPES = 5
KERNEL = {"m4": 0, "quintic": 1, "gaussian": 2}
AVISC = {"none": 0, "mon97": 1, "mon97mm97": 2, "mon97cd2010": 3}
ACOND = {"none": 0, "wadsley2008": 1, "price2008": 2}
TDAVISC = {"none": 0, "mm97": 1, "cd2010": 2}
EOS = {"energy_eqn": 0, "constant_temp": 0, "isothermal": 1, "barotropic": 2, "barotropic2": 3}
BOUNDARY = {"open": 0, "periodic": 1, "mirror": 2}
MAC = {"geometric": 0, "eigenmac": 1, "gadget2": 2}
MULTIPOLE = {"monopole": 0, "quadrupole": 1, "fast_monopole": 2, "fast_quadrupole": 3}
FLAG_DEAD, FLAG_ACTIVE, FLAG_POTMIN = 1, 2, 8
F_DENSITY, F_FORCES, F_ALL = 1, 2, 0xFFFFFFFF
DEVICE_FIELDS = ("h", "rho", "hrangesqd", "hfactor", "invomega", "zeta", "sound", "pressure", "u", "perm")
ERRORS = {1: "GSPH_EINVAL", 2: "GSPH_ENODEVICE", 3: "GSPH_ECUDA", 4: "GSPH_ENOMEM", 5: "GSPH_ESTATE",
          6: "GSPH_ECONVERGE"}


class Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("ndim", C.c_int32), ("kernel", C.c_int32),
        ("tabulated_kernel", C.c_int32), ("h_fac", C.c_double), ("h_converge", C.c_double),
        ("avisc", C.c_int32), ("acond", C.c_int32), ("tdavisc", C.c_int32),
        ("alpha_visc", C.c_double), ("beta_visc", C.c_double), ("alpha_visc_min", C.c_double),
        ("eos", C.c_int32), ("gamma_eos", C.c_double), ("temp0", C.c_double), ("mu_bar", C.c_double),
        ("rho_bary", C.c_double), ("boxmin", C.c_double * 3), ("boxmax", C.c_double * 3),
        ("boundary_lhs", C.c_int32 * 3), ("boundary_rhs", C.c_int32 * 3),
        ("hydro_forces", C.c_int32), ("self_gravity", C.c_int32), ("thetamaxsqd", C.c_double),
        ("gravity_mac", C.c_int32), ("macerror", C.c_double), ("multipole", C.c_int32),
        ("Nleafmax", C.c_int32), ("hmask", C.c_uint32 * NTYPES), ("hydromask", C.c_uint32 * NTYPES),
        ("gravmask", C.c_uint32), ("hydro_forces_type", C.c_uint32), ("self_gravity_type", C.c_uint32),
        ("create_sinks", C.c_int32), ("device", C.c_int32), ("conservative_sph_star_gravity", C.c_int32), ("nbody_softening", C.c_int32),
        ("reserved", C.c_int32 * 5),
    ]


class Field(C.Structure):
    _fields_ = [("base", C.c_void_p), ("stride", C.c_size_t)]


VEC_FIELDS = ("r", "v", "a", "atree")
SCALAR_FIELDS = ("m", "h", "u", "rho", "hrangesqd", "hfactor", "invomega", "zeta", "sound", "pressure",
                 "dudt", "div_v", "alpha", "dalphadt", "gpot", "gpot_hydro")
INT_FIELDS = ("flags", "ptype", "level", "levelneib", "iorig", "sinkid")


class ParticleView(C.Structure):
    _fields_ = ([(n, Field * 3) for n in VEC_FIELDS] + [(n, Field) for n in SCALAR_FIELDS]
                + [(n, Field) for n in INT_FIELDS])


class TreeInfo(C.Structure):
    _fields_ = [("ltot", C.c_int32), ("Ncell", C.c_int32), ("gtot", C.c_int32), ("Nleafmax", C.c_int32),
                ("Ntot", C.c_int64)]


class Stats(C.Structure):
    _fields_ = [("Nhydro", C.c_int64), ("n_active", C.c_int64), ("ngb_mean", C.c_double),
                ("ngb_max", C.c_int32), ("cand_mean", C.c_double), ("hiter_mean", C.c_double),
                ("hiter_max", C.c_int32), ("density_retries", C.c_int32), ("ndirect_mean", C.c_double),
                ("ncell_mean", C.c_double), ("kernel_launches", C.c_int64), ("ms_build_tree", C.c_float),
                ("ms_properties", C.c_float), ("ms_forces", C.c_float), ("ms_upload", C.c_float),
                ("ms_download", C.c_float), ("ms_kernel_density", C.c_float),
                ("ms_kernel_hydro", C.c_float), ("ms_kernel_gravity", C.c_float), ("ms_tree_sort", C.c_float),
                ("ms_tree_levels", C.c_float), ("ms_tree_stock", C.c_float)]


class IntegratorConfig(C.Structure):
    """gsph_integrator_config"""
    _fields_ = [("accel_mult", C.c_double), ("courant_mult", C.c_double), ("energy_mult", C.c_double),
                ("energy_integration", C.c_int32), ("ntreebuildstep", C.c_int32), ("ntreestockstep", C.c_int32),
                ("sink_particles", C.c_int32), ("Nlevels", C.c_int32), ("level_diff_max", C.c_int32),
                ("sph_single_timestep", C.c_int32), ("reserved", C.c_int32 * 1)]


class BlockState(C.Structure):
    """gsph_block_state"""
    _fields_ = [("n", C.c_int64), ("nresync", C.c_int64), ("level_max", C.c_int32), ("level_step", C.c_int32),
                ("dt_max", C.c_double)]


EXPORTS = [
    "gsph_default_config", "gsph_create", "gsph_destroy", "gsph_last_error", "gsph_upload", "gsph_download",
    "gsph_download_async", "gsph_set_stars", "gsph_tree_shard_range", "gsph_build_tree_shard_begin",
    "gsph_build_tree_shard_finish", "gsph_upload_range", "gsph_download_range", "gsph_device_array",
    "gsph_build_tree", "gsph_update_active_counters", "gsph_update_properties", "gsph_zero_accelerations",
    "gsph_update_hydro_forces", "gsph_update_all_forces", "gsph_get_hmax", "gsph_synchronize",
    "gsph_timer_start", "gsph_timer_stop", "gsph_set_active_slice", "gsph_density_pass", "gsph_refresh_hmax",
    "gsph_device_field",
    "gsph_neighbour_counts", "gsph_neighbour_lists", "gsph_tree_info_get", "gsph_tree_leaf_of",
    "gsph_tree_cells", "gsph_get_stats", "gsph_kernel_table", "gsph_measure_fp_peak",
    "gsph_integrator_begin", "gsph_integrator_step", "gsph_integrator_state", "gsph_integrator_block_state",
    "gsph_integrator_start_state",
    "gsph_update_all_forces_begin", "gsph_update_all_forces_finish", "gsph_integrator_advance",
    "gsph_integrator_finish", "gsph_integrator_check_timesteps", "gsph_update_thermal_properties",
    "gsph_update_star_gas_forces", "gsph_derive_density_fields",
    "gsph_integrator_set_stars", "gsph_integrator_get_stars", "gsph_integrator_set_stars_hermite4",
    "gsph_integrator_get_star_derivatives", "gsph_integrator_get_star_levels",
]

_lib = None


class GsphError(RuntimeError):
    def __init__(self, code, msg):
        RuntimeError.__init__(self, "%s: %s" % (ERRORS.get(code, "error %d" % code), msg))
        self.code = code


def lib():
    """Load gandalf_b200/libgsph.so.  Raises if it is missing: the product has no CPU fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("gandalf_b200/libgsph.so is not built (run `make -C gandalf_b200/csrc` or "
                          "__graft_entry__.build()); gandalf_b200 has no CPU fallback")
    L = C.CDLL(LIB_PATH)
    H = C.c_void_p
    L.gsph_default_config.argtypes = [C.POINTER(Config)]
    L.gsph_create.argtypes = [C.POINTER(Config), C.POINTER(H)]
    L.gsph_destroy.argtypes = [H]
    L.gsph_destroy.restype = None
    L.gsph_last_error.argtypes = [H]
    L.gsph_last_error.restype = C.c_char_p
    L.gsph_upload.argtypes = [H, C.POINTER(ParticleView), C.c_int64]
    L.gsph_download.argtypes = [H, C.POINTER(ParticleView), C.c_uint32]
    L.gsph_download_async.argtypes = [H, C.POINTER(ParticleView), C.c_uint32]
    L.gsph_set_stars.argtypes = [H, C.c_void_p, C.c_int]
    L.gsph_tree_shard_range.argtypes = [H, C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.gsph_build_tree_shard_begin.argtypes = [H, C.c_int, C.c_int]
    L.gsph_build_tree_shard_finish.argtypes = [H]
    L.gsph_upload_range.argtypes = [H, C.POINTER(ParticleView), C.c_int64, C.c_int64, C.c_int64]
    L.gsph_download_range.argtypes = [H, C.POINTER(ParticleView), C.c_uint32, C.c_int64, C.c_int64, C.c_int]
    L.gsph_device_array.argtypes = [H, C.c_int, C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64),
                                    C.POINTER(C.c_int32)]
    L.gsph_build_tree.argtypes = [H, C.c_int, C.c_double]
    L.gsph_integrator_begin.argtypes = [H, C.POINTER(IntegratorConfig), C.c_int, C.c_double]
    L.gsph_integrator_step.argtypes = [H, C.c_int]
    L.gsph_integrator_advance.argtypes = [H, C.POINTER(C.c_int), C.POINTER(C.c_double)]
    L.gsph_integrator_finish.argtypes = [H]
    L.gsph_integrator_check_timesteps.argtypes = [H, C.POINTER(C.c_int)]
    L.gsph_update_thermal_properties.argtypes = [H]
    L.gsph_integrator_set_stars.argtypes = [H, C.c_void_p, C.c_int, C.c_double]
    L.gsph_integrator_get_stars.argtypes = [H, C.c_void_p, C.c_int]
    L.gsph_integrator_set_stars_hermite4.argtypes = [H, C.c_void_p, C.c_int, C.c_double]
    L.gsph_integrator_get_star_derivatives.argtypes = [H, C.c_void_p, C.c_void_p, C.c_void_p]
    L.gsph_integrator_get_star_levels.argtypes = [H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.gsph_update_star_gas_forces.argtypes = [H, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.gsph_integrator_state.argtypes = [H, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    L.gsph_integrator_block_state.argtypes = [H, C.POINTER(BlockState), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.gsph_integrator_start_state.argtypes = [H] + [C.c_void_p]*6
    for n in ("gsph_update_active_counters", "gsph_update_properties", "gsph_zero_accelerations",
              "gsph_update_hydro_forces", "gsph_synchronize", "gsph_timer_start", "gsph_density_pass",
              "gsph_update_all_forces_begin", "gsph_update_all_forces_finish",
              "gsph_refresh_hmax"):
        getattr(L, n).argtypes = [H]
    L.gsph_update_all_forces.argtypes = [H, C.c_void_p, C.c_int]
    L.gsph_set_active_slice.argtypes = [H, C.c_int64, C.c_int64]
    L.gsph_derive_density_fields.argtypes = [H, C.c_int64, C.c_int64]
    L.gsph_device_field.argtypes = [H, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
    L.gsph_timer_stop.argtypes = [H, C.POINTER(C.c_float)]
    L.gsph_get_hmax.argtypes = [H, C.POINTER(C.c_double)]
    L.gsph_neighbour_counts.argtypes = [H, C.c_void_p]
    L.gsph_neighbour_lists.argtypes = [H, C.c_void_p, C.c_void_p]
    L.gsph_tree_info_get.argtypes = [H, C.POINTER(TreeInfo)]
    L.gsph_tree_leaf_of.argtypes = [H, C.c_void_p]
    L.gsph_tree_cells.argtypes = [H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.gsph_get_stats.argtypes = [H, C.POINTER(Stats)]
    L.gsph_kernel_table.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.gsph_measure_fp_peak.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    for n in EXPORTS:
        if n not in ("gsph_destroy", "gsph_last_error"):
            getattr(L, n).restype = C.c_int
    _lib = L
    return L
