"""Run-time parameters of the path: the reference's `.dat` files ("Comment : key = value") and
their defaults, restricted to the knobs the neighbour path reads.

Mirrors Parameters::ParseLine / SetDefaultValues (src/Common/Parameters.cpp:126-158, :166-536) and the
translation the reference does when it builds GradhSph / GradhSphTree
(src/GradhSph/GradhSphSimulation.cpp:54-333).  Unit scaling (SimUnits) stays with the host
program: temp0 and rho_bary are taken in code units, i.e. as for `dimensionless = 1`."""
import ctypes as C

from . import _capi

# defaults of the knobs on the path (Parameters.cpp:252-314, :373-384, :409)
DEFAULTS = {
    "ndim": 3, "sim": "sph", "neib_search": "kdtree", "kernel": "m4", "tabulated_kernel": 1,
    "h_fac": 1.2, "h_converge": 0.01, "hydro_forces": 1, "self_gravity": 0,
    "gas_eos": "energy_eqn", "gamma_eos": 1.66666666666666, "temp0": 1.0, "mu_bar": 1.0,
    "rho_bary": 1.0e-14, "avisc": "mon97", "acond": "none", "time_dependent_avisc": "none",
    "alpha_visc": 1.0, "alpha_visc_min": 0.1, "beta_visc": 2.0,
    "gravity_mac": "geometric", "multipole": "quadrupole", "thetamaxsqd": 0.1, "macerror": 0.0001,
    "Nleafmax": 6, "ntreebuildstep": 1, "ntreestockstep": 1, "create_sinks": 0, "conservative_sph_star_gravity": 1, "nbody_softening": 1, "sink_particles": 0,
    "Nlevels": 1, "sph_integration": "lfkdk", "accel_mult": 0.3, "courant_mult": 0.15, "energy_mult": 0.4,
    "nbody": "hermite4", "nbody_mult": 0.1, "level_diff_max": 1, "sph_single_timestep": 0,
}
for _k in range(3):
    DEFAULTS["boundary_lhs[%d]" % _k] = "open"
    DEFAULTS["boundary_rhs[%d]" % _k] = "open"
    DEFAULTS["boxmin[%d]" % _k] = -9.9e30
    DEFAULTS["boxmax[%d]" % _k] = 9.9e30


class Parameters(dict):
    """Parameter dictionary with the reference's parsing rules and defaults."""

    def __init__(self, *a, **kw):
        dict.__init__(self, DEFAULTS)
        self.update(dict(*a, **kw))

    @staticmethod
    def _convert(key, value):
        d = DEFAULTS.get(key)
        value = value.strip() if isinstance(value, str) else value
        if isinstance(d, int) and not isinstance(d, bool):
            return int(value)
        if isinstance(d, float):
            return float(value)
        return value

    def parse_line(self, line):
        """'Comment : variable = value' or 'variable = value'; '#' starts a comment line
        (Parameters::ParseLine, Parameters.cpp:126-158)."""
        line = "".join(line.split())            # TrimWhiteSpace removes all blanks
        if not line or line.startswith("#"):
            return
        colon, equal = line.find(":"), line.find("=")
        if equal < 0 or (colon >= 0 and colon > equal):
            return
        key = line[colon + 1:equal]
        self[key] = self._convert(key, line[equal + 1:])

    def read_params_file(self, path):
        with open(path) as f:
            for line in f:
                self.parse_line(line)
        return self


def integrator_config_from_params(p):
    """-> _capi.IntegratorConfig for gsph_integrator_begin (the device-resident main loop): lfkdk with one
    global time step or (Nlevels > 1) hierarchical block time steps; energy_integration as
    GradhSphSimulation.cpp:110-123 derives it from gas_eos."""
    q = Parameters(p)
    if q["sph_integration"] != "lfkdk":
        raise ValueError("sph_integration = %s is not on the accelerated path (lfkdk only)" % q["sph_integration"])
    if int(q["Nlevels"]) < 1:
        raise ValueError("Nlevels must be >= 1")
    if int(q["sink_particles"]) != 0 or int(q["create_sinks"]) != 0:
        # MainLoop would run SearchForNewSinkParticles / AccreteMassToSinks / UpdateStellarProperties
        # (SphSimulation.cpp:815-833): not on the accelerated path
        raise ValueError("sink_particles / create_sinks are not on the accelerated path of the device-resident "
                         "main loop (no sink search or accretion)")
    c = _capi.IntegratorConfig()
    c.accel_mult, c.courant_mult, c.energy_mult = (float(q["accel_mult"]), float(q["courant_mult"]),
                                                   float(q["energy_mult"]))
    c.energy_integration = 1 if q["gas_eos"] == "energy_eqn" else 0
    c.ntreebuildstep, c.ntreestockstep = int(q["ntreebuildstep"]), int(q["ntreestockstep"])
    c.sink_particles = int(q["sink_particles"])
    c.Nlevels = int(q["Nlevels"])
    c.level_diff_max = int(q["level_diff_max"])
    c.sph_single_timestep = 1 if int(q["sph_single_timestep"]) != 0 else 0
    return c


def config_from_params(p, device=0):
    """-> _capi.Config for gsph_create, from a Parameters/dict with the reference's names."""
    q = Parameters(p)
    L = _capi.lib()
    c = _capi.Config()
    rc = L.gsph_default_config(C.byref(c))
    assert rc == 0
    c.ndim = int(q["ndim"])
    c.kernel = _capi.KERNEL[q["kernel"]]
    c.tabulated_kernel = int(q["tabulated_kernel"])
    c.h_fac, c.h_converge = float(q["h_fac"]), float(q["h_converge"])
    c.avisc = _capi.AVISC[q["avisc"]]
    c.acond = _capi.ACOND[q["acond"]]
    td = q["time_dependent_avisc"]
    if td not in _capi.TDAVISC:
        raise ValueError("time_dependent_avisc = %s is not on the accelerated path" % td)
    c.tdavisc = _capi.TDAVISC[td]
    if c.tdavisc == 1 and c.avisc == 1:
        c.avisc = 2                               # mon97 + mm97 -> mon97mm97 (GradhSphSimulation.cpp:69-88)
    if c.tdavisc == 2 and c.avisc == 1:
        c.avisc = 3                               # mon97 + cd2010 -> mon97cd2010
    c.alpha_visc, c.beta_visc = float(q["alpha_visc"]), float(q["beta_visc"])
    c.alpha_visc_min = float(q["alpha_visc_min"])
    c.eos = _capi.EOS[q["gas_eos"]]
    c.gamma_eos, c.temp0 = float(q["gamma_eos"]), float(q["temp0"])
    c.mu_bar, c.rho_bary = float(q["mu_bar"]), float(q["rho_bary"])
    for k in range(3):
        c.boxmin[k] = float(q["boxmin[%d]" % k])
        c.boxmax[k] = float(q["boxmax[%d]" % k])
        c.boundary_lhs[k] = _capi.BOUNDARY[q["boundary_lhs[%d]" % k]] if k < c.ndim else 0
        c.boundary_rhs[k] = _capi.BOUNDARY[q["boundary_rhs[%d]" % k]] if k < c.ndim else 0
    c.hydro_forces, c.self_gravity = int(q["hydro_forces"]), int(q["self_gravity"])
    c.thetamaxsqd, c.macerror = float(q["thetamaxsqd"]), float(q["macerror"])
    c.gravity_mac = _capi.MAC[q["gravity_mac"]]
    c.multipole = _capi.MULTIPOLE[q["multipole"]]
    c.Nleafmax = int(q["Nleafmax"])
    c.create_sinks = int(q["create_sinks"])
    c.conservative_sph_star_gravity = int(q["conservative_sph_star_gravity"])
    c.nbody_softening = int(q["nbody_softening"])
    c.device = device
    return c
