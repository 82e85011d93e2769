"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol include/gsph.h
declares (no compute without a GPU), the .dat parameter mirror, the particle views, and the
multi-GPU slice/exchange logic under a world_size-2 gloo group."""
import ctypes as C
import os
import re
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from gandalf_b200 import _capi
    hdr = open(os.path.join(ROOT, "include", "gsph.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(gsph_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no prototypes found"
    L = _capi.lib()
    for name in sorted(declared):
        assert hasattr(L, name), "libgsph.so does not export %s" % name
    assert declared == set(_capi.EXPORTS)


def test_struct_layouts_match_the_header():
    """sizes computed by the C compiler from include/gsph.h equal the ctypes mirrors"""
    from gandalf_b200 import _capi
    src = '#include "gsph.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu %zu %zu\\n",sizeof(gsph_config),' \
          'sizeof(gsph_particle_view),sizeof(gsph_stats),sizeof(gsph_tree_info),sizeof(gsph_field));return 0;}\n'
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")])
        sizes = [int(x) for x in subprocess.check_output([os.path.join(d, "t")]).split()]
    assert sizes == [C.sizeof(_capi.Config), C.sizeof(_capi.ParticleView), C.sizeof(_capi.Stats),
                     C.sizeof(_capi.TreeInfo), C.sizeof(_capi.Field)]


def test_no_device_is_a_loud_error():
    import gandalf_b200 as g
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(g.GsphError) as e:
        g.GpuSphTree({"ndim": 3})
    assert e.value.code == 2 and "no CPU fallback" in str(e.value)


def test_kernel_tables_equal_oracle():
    """parity stage P0: the six tabulated-kernel tables, bit for bit (host side, no device needed)"""
    import oracle_py as orc
    from gandalf_b200 import sphtree
    for nd in (1, 2, 3):
        for kn in ("m4", "quintic"):
            p = orc.params_from_dict({"ndim": nd, "kernel": kn, "tabulated_kernel": 1})
            for w in range(6):
                assert np.array_equal(sphtree.kernel_table(nd, kn, w), orc.kernel_table(p, w)), (nd, kn, w)


def test_parameter_file_parsing(tmp_path):
    from gandalf_b200 import Parameters, config_from_params
    f = tmp_path / "x.dat"
    f.write_text(textwrap.dedent("""\
        # comment line
        Run id : run_id = TEST1
        Dimensions :  ndim = 2
        kernel = quintic
        Tabulate : tabulated_kernel = 0
        h_fac = 1.3
        this line has no equals sign
        : boundary_lhs[0] = periodic
        boundary_rhs[0] = periodic
        boxmin[0] = -0.5
        boxmax[0] = 0.5
        avisc = mon97
        time_dependent_avisc = mm97
        Nleafmax = 12
        """))
    p = Parameters().read_params_file(str(f))
    assert p["run_id"] == "TEST1" and p["ndim"] == 2 and p["kernel"] == "quintic" and p["h_fac"] == 1.3
    assert p["thetamaxsqd"] == 0.1 and p["multipole"] == "quadrupole"      # defaults (Parameters.cpp:309,313)
    c = config_from_params(p)
    assert (c.ndim, c.kernel, c.tabulated_kernel, c.Nleafmax) == (2, 1, 0, 12)
    assert c.avisc == 2 and c.tdavisc == 1
    assert c.boundary_lhs[0] == 1 and c.boundary_lhs[1] == 0 and c.boxmin[0] == -0.5


def test_particle_views():
    import gandalf_b200 as g
    a = g.ParticleSet(7, 3, "aos")
    assert a.records.dtype.itemsize == 448          # sizeof(GradhSphParticle<3>), SURVEY 8(a)
    v = a.view()
    assert v.m.stride == 448 and v.r[2].base - v.r[0].base == 16 and v.flags.stride == 448
    s = g.ParticleSet(7, 2, "soa")
    v = s.view()
    assert v.r[0].stride == 8 and v.r[1].base - v.r[0].base == 7*8 and not v.r[2].base
    s["r"][...] = np.arange(14).reshape(7, 2)
    assert s["r"][3, 1] == 7
    part = g.ParticleSet(7, 3, "soa", fields=("r", "m", "h"))
    assert not part.view().v[0].base and part.view().m.base


def test_slice_bounds_cover_everything():
    from gandalf_b200 import domain
    for N in (1, 7, 64, 1000003):
        for R in (1, 2, 4, 8):
            plan, padded = domain.exchange_plan(N, R)
            assert plan[0][0] == 0 and plan[-1][1] == N and padded >= N and padded - N < R
            for (a, b), (c, d) in zip(plan[:-1], plan[1:]):
                assert b == c and a <= b


def test_shard_ranges_are_the_kd_tree_top_levels():
    """the multi-GPU tree build shards by the nodes of level log2(R): sizes N -> (N//2, N - N//2)"""
    from gandalf_b200 import domain
    for N in (8, 1001, 64 * 2 ** 20, 16003008):
        for R in (1, 2, 4, 8):
            rg = domain.shard_ranges(N, R)
            assert len(rg) == R and rg[0][0] == 0 and rg[-1][1] == N
            for (a, b), (c, d) in zip(rg[:-1], rg[1:]):
                assert b == c
            sizes = [b - a for a, b in rg]
            assert max(sizes) - min(sizes) <= 1 or R == 1
    assert domain.shard_ranges(1001, 2) == [(0, 500), (500, 1001)]
    assert domain.shard_ranges(1001, 4) == [(0, 250), (250, 500), (500, 750), (750, 1001)]


_WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["GSPH_ROOT"])
from gandalf_b200 import domain
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
N = 1001
lo, hi, chunk = domain.slice_bounds(N, world, rank)
full = torch.full((chunk*world,), -1.0, dtype=torch.float64)
full[lo:hi] = torch.arange(lo, hi, dtype=torch.float64)          # "post-density field" of my slice
dist.all_gather_into_tensor(full, full[rank*chunk:(rank + 1)*chunk].clone())
assert torch.equal(full[:N], torch.arange(N, dtype=torch.float64)), full
t = torch.tensor([float(rank + 1)]); dist.all_reduce(t, op=dist.ReduceOp.MAX); assert t.item() == world
dist.barrier()
sys.stdout.write("rank %d ok\n" % rank); sys.stdout.flush()
"""


def test_exchange_step_world_size_2_gloo(tmp_path):
    """the one exchange step per pass (in-place all-gather of tree-order slices) with 2 CPU ranks"""
    w = tmp_path / "worker.py"
    w.write_text(_WORKER)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    env = dict(os.environ, GSPH_ROOT=ROOT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(w)],
                         env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=300)
    text = out.stdout.decode()
    assert out.returncode == 0, text[-2000:]
    assert text.count("ok") >= 2 and "rank 0" in text and "rank 1" in text, text[-500:]


def test_reference_perturbation_hash():
    """bench.py perturbs h with the same integer hash as oracle/ref_harness.cpp perturb_h"""
    from gandalf_b200 import workloads
    f = workloads.ref_perturb_factor(np.arange(5), 0.1)
    want = []
    for i in range(5):
        x = (i*2654435761) & 0xFFFFFFFF
        x ^= x >> 16; x = (x*2246822519) & 0xFFFFFFFF; x ^= x >> 13
        want.append(1.0 + 0.1*(2.0*((x & 0xFFFFFF)/float(0x1000000)) - 1.0))
    assert np.allclose(f, want, rtol=0, atol=0)


def test_integrator_config_follows_the_reference_parameters():
    """the device main loop's configuration from .dat parameters: multipliers (Parameters.cpp:239-263),
    energy integration from gas_eos (GradhSphSimulation.cpp:110-123), tree schedule; only lfkdk with one
    global time step is accepted"""
    import pytest
    from gandalf_b200.params import integrator_config_from_params
    c = integrator_config_from_params({"ndim": 3})
    assert (c.accel_mult, c.courant_mult, c.energy_mult) == (0.3, 0.15, 0.4)
    assert c.energy_integration == 1 and c.ntreebuildstep == 1 and c.ntreestockstep == 1
    c = integrator_config_from_params({"gas_eos": "isothermal", "courant_mult": 0.2, "ntreebuildstep": 8,
                                       "ntreestockstep": 2})
    assert c.energy_integration == 0 and c.courant_mult == 0.2 and (c.ntreebuildstep, c.ntreestockstep) == (8, 2)
    for eos in ("barotropic", "barotropic2", "constant_temp"):
        assert integrator_config_from_params({"gas_eos": eos}).energy_integration == 0
    c = integrator_config_from_params({"Nlevels": 5})
    assert c.Nlevels == 5 and c.level_diff_max == 1            # block time steps, Parameters.cpp:245
    with pytest.raises(ValueError):
        integrator_config_from_params({"sph_integration": "lfdkd"})


def test_c_abi_structs_match_the_header_sizes(tmp_path):
    """ctypes mirrors of the C structs: sizes as the C compiler lays out include/gsph.h"""
    import ctypes as C
    import os
    import subprocess
    from gandalf_b200 import _capi
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "gsph.h"\nint main(void) { printf("%zu %zu %zu %zu\\n", '
                   'sizeof(gsph_config), sizeof(gsph_integrator_config), sizeof(gsph_stats), sizeof(gsph_particle_view)); '
                   'return 0; }\n')
    exe = str(tmp_path / "sz")
    subprocess.check_call(["gcc", "-I", os.path.join(root, "include"), "-o", exe, str(src)])
    sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    assert sizes == [C.sizeof(_capi.Config), C.sizeof(_capi.IntegratorConfig), C.sizeof(_capi.Stats),
                     C.sizeof(_capi.ParticleView)]
    cfg = _capi.Config()
    assert _capi.lib().gsph_default_config(C.byref(cfg)) == 0
    assert cfg.conservative_sph_star_gravity == 1 and cfg.nbody_softening == 1 and list(cfg.reserved) == [0]*5


def test_domain_driver_main_loop_sequence():
    """host logic of the multi-rank device main loop (gandalf_b200/domain.py) with a recording stub in
    place of the device tree: per step advance -> (sharded) build -> slice -> density + exchange ->
    zero -> forces -> exchange of the force results -> finish; restock / extrapolate steps do not shard"""
    from gandalf_b200 import domain

    class Cfg:
        ndim, self_gravity, tdavisc, create_sinks = 3, 1, 0, 0

    class Stats:
        Nhydro = 1001

    class Stub:
        cfg = Cfg()

        def __init__(self, modes, Nlevels=1, woken=()):
            self.calls, self.modes = [], list(modes)
            self.params, self.woken = {"Nlevels": Nlevels}, list(woken)

        def IntegratorCheckTimesteps(self):
            self.calls.append("IntegratorCheckTimesteps")
            return self.woken.pop(0)

        def __getattr__(self, name):
            def f(*a, **k):
                self.calls.append(name)
            return f

        def Stats(self):
            return Stats()

        def IntegratorAdvance(self):
            self.calls.append("IntegratorAdvance")
            return self.modes.pop(0), 0.01

        def SetActiveSlice(self, lo, hi):
            self.calls.append(("SetActiveSlice", lo, hi))

        def BuildTreeMode(self, mode, ts):
            self.calls.append(("BuildTreeMode", mode))

        def DeriveDensityFields(self, lo, hi):
            self.calls.append(("DeriveDensityFields", lo, hi))

    class Driver(domain.DomainDriver):
        def _sharded_rebuild(self, parts):
            self.tree.calls.append("sharded_rebuild")

        def _wrap_fields(self):
            import numpy as np
            return [np.zeros(self.chunk*self.world) for _ in range(4)]      # h, rho, invomega, zeta

        def _gather_inplace(self, names, staging, n_valid):
            self.tree.calls.append(("gather", tuple(names), staging))

        def _reduce_max_inplace(self, name, n_valid):
            self.tree.calls.append(("reduce_max", name))

    import sys
    import types
    fake = types.ModuleType("torch")                      # UpdateAllSphProperties synchronises torch's stream
    fake.cuda = types.SimpleNamespace(current_stream=lambda: types.SimpleNamespace(synchronize=lambda: None))
    fake.distributed = types.ModuleType("torch.distributed")
    gathers = []
    fake.distributed.all_gather_into_tensor = lambda out, inp, group=None: gathers[-1].calls.append("all_gather")
    saved = {k: sys.modules.get(k) for k in ("torch", "torch.distributed")}
    sys.modules["torch"], sys.modules["torch.distributed"] = fake, fake.distributed
    try:
        stub = Stub([0, 1, 2])
        gathers.append(stub)
        drv = Driver(stub, rank=1, world_size=2)
        drv.MainLoop(3)
        # block time steps: thermal refresh of every particle before the forces, levelneib by a MAX reduction,
        # the passes repeated while CheckTimesteps wakes particles (here: once in the first step)
        bstub = Stub([0, 2], Nlevels=4, woken=[3, 0, 0])
        gathers.append(bstub)
        Driver(bstub, rank=1, world_size=2).MainLoop(2)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    lo, hi, _ = domain.slice_bounds(1001, 2, 1)
    dens = ["DensityPass", "Synchronize", "all_gather", "RefreshHmax", "all_gather", "all_gather", "all_gather",
            ("DeriveDensityFields", lo, hi)]
    step = dens + ["ZeroAccelerations", "UpdateAllSphForces", "Synchronize",
            ("gather", ("ax", "ay", "az", "atx", "aty", "atz", "gpot", "gpot_hydro", "dudt", "div_v", "dalphadt", "alpha", "levelneib"),
             False), "IntegratorFinish"]
    expect = []
    for build in ("sharded_rebuild", ("BuildTreeMode", 1), ("BuildTreeMode", 2)):
        expect += ["IntegratorAdvance", build, ("SetActiveSlice", lo, hi)] + step
    assert stub.calls == expect, stub.calls
    bpass = dens + ["ZeroAccelerations", "UpdateThermalProperties", "UpdateAllSphForces", "Synchronize",
             ("gather", ("ax", "ay", "az", "atx", "aty", "atz", "gpot", "gpot_hydro", "dudt", "div_v", "dalphadt", "alpha"), False),
             ("reduce_max", "levelneib")]
    bexpect = (["IntegratorAdvance", "sharded_rebuild", ("SetActiveSlice", lo, hi)] + bpass + ["IntegratorCheckTimesteps"] + bpass
               + ["IntegratorCheckTimesteps", "IntegratorFinish"]
               + ["IntegratorAdvance", ("BuildTreeMode", 2), ("SetActiveSlice", lo, hi)] + bpass
               + ["IntegratorCheckTimesteps", "IntegratorFinish"])
    assert bstub.calls == bexpect, bstub.calls


def test_defaults_equal_the_reference():
    """every key gandalf_b200.params.DEFAULTS mirrors has the value Parameters::SetDefaultValues assigns
    (Parameters.cpp:166-536; tests/golden/ref_defaults.json, generator make_defaults_golden.py) -- a
    .dat file that leaves a key out must mean the same run in both codes (e.g. nbody = hermite4)"""
    import json
    from gandalf_b200.params import DEFAULTS
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_defaults.json")))
    for k, v in DEFAULTS.items():
        assert k in ref, "%s is not a parameter of the reference" % k
        assert ref[k] == v and type(ref[k]) is type(v), "%s: %r, reference %r" % (k, v, ref[k])
    assert DEFAULTS["nbody"] == "hermite4"


def test_device_loop_rejects_what_it_does_not_run():
    """the device-resident main loop has no sink search / accretion and only the lfkdk gas integrator:
    parameter sets that ask for them are refused instead of silently running something else"""
    from gandalf_b200.params import integrator_config_from_params
    c = integrator_config_from_params({"ndim": 3})
    assert c.sink_particles == 0 and c.ntreebuildstep == 1
    assert integrator_config_from_params({"ndim": 3, "Nlevels": 5}).Nlevels == 5          # block time steps are on the path
    assert integrator_config_from_params({"Nlevels": 5, "sph_single_timestep": 1}).sph_single_timestep == 1
    for bad in ({"sink_particles": 1}, {"create_sinks": 1}, {"sph_integration": "lfdkd"}):
        with pytest.raises(ValueError):
            integrator_config_from_params(dict({"ndim": 3}, **bad))


def test_config_workloads_have_the_named_shapes():
    """the synthetic particle sets of BASELINE configurations 2 and 3 (gandalf_b200/workloads.py), generated on the CPU
    at a reduced size: KHI = two lattices of densities 1 and 2 with equal particle mass in a periodic unit square,
    bossbodenheimer = unit-mass sphere in solid-body rotation with an m = 2 mass perturbation"""
    from gandalf_b200 import workloads
    p, ic = workloads.khi(60000, device="cpu")
    r, m, v = ic["r"].numpy(), ic["m"].numpy(), ic["v"].numpy()
    assert p["ndim"] == 2 and p["kernel"] == "quintic" and p["Nleafmax"] == 12 and p["acond"] == "wadsley2008"
    assert r.shape[0] == 2 and np.abs(r).max() < 0.5 and abs(m.sum() - 1.5) < 1e-12          # 0.5*1 + 0.5*2
    inner = np.abs(r[1]) < 0.25
    assert abs(m[inner].sum() - 1.0) < 1e-12 and np.allclose(m[inner], m[inner][0]) and np.all(v[0][inner] == 0.5)
    assert np.all(v[0][~inner] == -0.5) and len(np.unique(np.round(r.T, 12), axis=0)) == r.shape[1]
    p, ic = workloads.bossbodenheimer(20000, device="cpu")
    r, m, v = ic["r"].numpy(), ic["m"].numpy(), ic["v"].numpy()
    assert p["gas_eos"] == "barotropic" and p["multipole"] == "fast_monopole" and p["Nleafmax"] == 6 and p["self_gravity"] == 1
    assert r.shape == (3, 20000) and (r*r).sum(0).max() < 1.0 and abs(m.sum() - 1.0) < 1e-12
    phi = np.arctan2(r[1], r[0])
    assert np.corrcoef(m, np.cos(2*phi))[0, 1] > 0.99                                       # the m = 2 mode
    omega = (r[0]*v[1] - r[1]*v[0])/(r[0]**2 + r[1]**2)
    assert np.allclose(omega, omega[0]) and np.all(v[2] == 0)                               # solid-body rotation about z


def test_particle_set_speaks_the_reference_array_names():
    """ParticleSet.ImportArray / ExtractArray take the reference's quantity names (Simulation::ImportArray,
    Simulation.cpp:1442-1612; SphSnapshot::ExtractArray, SphSnapshot.cpp:529-600) on SoA and AoS layouts"""
    import pytest
    import gandalf_b200 as g
    rng = np.random.default_rng(5)
    for layout in ("soa", "aos"):
        ps = g.ParticleSet(7, 2, layout)
        vals = {n: rng.normal(size=7) for n in ("x", "y", "vx", "vy", "m", "h", "rho", "u")}
        for n, v in vals.items():
            ps.ImportArray(v, n)
        assert np.array_equal(ps["r"][:, 1], vals["y"]) and np.array_equal(ps["v"][:, 0], vals["vx"])
        for n, v in vals.items():
            out = ps.ExtractArray(n)
            assert out.flags["C_CONTIGUOUS"] and np.array_equal(out, v)
        assert np.array_equal(ps.ExtractArray("iorig"), np.arange(7))
        with pytest.raises(ValueError):
            ps.ExtractArray("z")                      # no third dimension
        with pytest.raises(ValueError):
            ps.ImportArray(np.zeros(7), "temperature")
        with pytest.raises(ValueError):
            ps.ImportArray(np.zeros(6), "x")
