#!/usr/bin/env python
"""bench.py -- SPH particle-updates/s (density + hydro + grav force) of the grad-h neighbour path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload cloud|glass] [--particles NP]

One "step" = one pass of the hot path over the synthetic particle set, exactly the sequence the
reference's main loop issues (src/Hydrodynamics/SphSimulation.cpp:634-710):
    BuildTree(rebuild) -> UpdateAllSphProperties -> ZeroAccelerations -> UpdateAllSphForces
(UpdateAllSphHydroForces for the hydro-only glass box), every particle active.
`value`  : particle-updates/s, inputs resident in HBM (device staging) when the timed region starts.
`e2e`    : the same metric through the public API with HOST buffers: upload from pinned memory,
           the step, download of the results, all inside the timed region.
--impl reference : the reference's own OpenMP CPU implementation (oracle/_ref/ref_harness_omp,
           the UNMODIFIED reference sources) on the box's host cores, on a bounded sample.
The CPU oracle / reference harness is only ever the *baseline leg*, never part of the GPU path.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "SPH particle-updates/sec (density+hydro+grav force)"
UNIT = "particle-updates/s"
H_PERTURB = 0.03          # both arms restart the h-rho iteration from the converged h perturbed by +-3 %
DEFAULT_PARTICLES = {"cloud": 64 * 2 ** 20, "glass": 252 ** 3, "khi": 983552, "bossbodenheimer": 4000000}
CPU_SAMPLE = {"cloud": 400000, "glass": 400000, "khi": 245888, "bossbodenheimer": 400000}


# ---------------------------------------------------------------------------------------------------
def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        threading.Thread.__init__(self, daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], stdout=subprocess.PIPE,
                                     stderr=subprocess.DEVNULL, timeout=5).stdout.decode().strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows for k in range(4) if len(r) >= 7 and r[3 + k] == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def pinned_alloc():
    import torch

    def alloc(shape, dtype):
        tdt = {np.dtype(np.float64): torch.float64, np.dtype(np.int32): torch.int32,
               np.dtype(np.uint32): torch.int32}[np.dtype(dtype)]
        t = torch.zeros(tuple(shape), dtype=tdt, pin_memory=True)
        a = t.numpy()
        alloc.keep.append(t)
        return a.view(dtype) if np.dtype(dtype) == np.uint32 else a
    alloc.keep = []
    return alloc


IN_FIELDS = ("r", "v", "m", "h", "u")
OUT_FIELDS = ("h", "rho", "hrangesqd", "hfactor", "invomega", "zeta", "sound", "pressure", "u",
              "a", "atree", "gpot", "dudt", "div_v", "levelneib")


def field_bytes(names, N, ndim):
    b = 0
    for n in names:
        if n in ("r", "v", "a", "atree"):
            b += 8 * ndim * N
        elif n in ("levelneib", "flags"):
            b += 4 * N
        else:
            b += 8 * N
    return b


# ---------------------------------------------------------------------------------------------------
def run_reference_sample(workload, n_sample, steps, threads):
    """The unmodified reference (OpenMP, -O3 = its COMPILER_MODE=STANDARD; the -ffast-math build hangs
    in SetupSimulation with g++ 13, see oracle/Makefile) on a bounded sample of the same workload,
    generated by the same recipe on the host."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import refio                                            # test infrastructure: baseline leg only
    if refio.harness_path("omp") is None:
        return None
    from gandalf_b200 import workloads
    params, ic = workloads.WORKLOADS[workload](n_sample, device="cpu")
    ic = {k: v.numpy() for k, v in ic.items()}
    N = ic["m"].shape[0]
    dat = {"run_id": "BENCHREF", "sim": "gradhsph", "ic": "python", "dimensionless": 1, "Nlevels": 1,
           "Nstepsmax": 1, "tend": 1.0e-10, "dt_snap": 1.0e10, "noutputstep": 1000000, "ndiagstep": 1000000,
           "out_file_form": "su", "sink_particles": 0, "sph_integration": "lfkdk", "neib_search": "kdtree",
           "Nhydromax": 4 * N if workload in ("glass", "khi") else N + 1000}
    dat.update(params)
    t0 = time.time()
    _, bench = refio.run_harness(dat, ic=dict(r=ic["r"].T.copy(), v=ic["v"].T.copy(), m=ic["m"], h=ic["h"], u=ic["u"]),
                                 dumps="", perturb_h=H_PERTURB, kind="omp", bench_steps=steps, threads=threads,
                                 timeout=3000)
    wall = time.time() - t0
    if bench is None:
        return None
    bench["wall_s"] = wall
    bench["params"] = params
    return bench


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def main_reference(args, rank, world):
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    n_sample = args.cpu_particles or CPU_SAMPLE[args.workload]
    steps = max(1, args.steps)
    b = run_reference_sample(args.workload, n_sample, steps, threads)
    if b is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_harness_omp not built"}))
        return 0
    value = b["updates_per_s"]
    sample = "%d-particle %s, %d passes, %s" % (b["N"], args.workload, b["steps"], cpu_model())
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": b["steps"], "warmup": args.warmup, "ms_per_step": 1e3 * b["total"] / b["steps"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.workload, b["N"]), "sample_of": workload_name(
            args.workload, args.particles or DEFAULT_PARTICLES[args.workload]), "parallelism": "openmp%d" % threads},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "blocks_s": {k: b[k] for k in ("BUILD_TREE", "BUILD_GHOST_TREE", "SPH_PROPERTIES", "SPH_FORCES")},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def workload_name(w, N):
    if w == "khi":
        return "2D Kelvin-Helmholtz (khi.dat scaled), %d particles, quintic tabulated, Nleafmax 12, wadsley2008" % N
    if w == "bossbodenheimer":
        return ("rotating sphere with m=2 perturbation (bossbodenheimer.dat scaled), %d particles, barotropic, "
                "fast_monopole theta^2=0.15, M4 tabulated" % N)
    if w == "cloud":
        return "turbulent self-gravitating cloud, %d particles, fast_monopole theta^2=0.15, M4 tabulated" % N
    return "3D periodic jittered-lattice (glass) box, hydro only, %d particles, M4 tabulated" % N


# ---------------------------------------------------------------------------------------------------
def main_ours(args, rank, world, local_rank):
    import torch
    import gandalf_b200 as g
    from gandalf_b200 import workloads, domain, sphtree, _capi

    torch.cuda.set_device(local_rank)
    dev = "cuda:%d" % local_rank
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(dev))
    NP = args.particles or DEFAULT_PARTICLES[args.workload]
    params, ic = workloads.WORKLOADS[args.workload](NP, device=dev)
    N = int(ic["m"].shape[0])
    ndim = int(params["ndim"])
    grav = int(params["self_gravity"]) == 1

    alloc = pinned_alloc()
    ps = g.ParticleSet(N, ndim, "soa", fields=set(IN_FIELDS), alloc=alloc)       # inputs of a step (pinned)
    pso = g.ParticleSet(N, ndim, "soa", fields=set(OUT_FIELDS), alloc=alloc)     # its results (pinned)
    for k in IN_FIELDS:
        dst = ps[k]
        src = ic[k]
        if src.dim() == 2:
            for c in range(ndim):
                dst[:, c] = src[c].cpu().numpy()
        else:
            dst[...] = src.cpu().numpy()
    del ic
    torch.cuda.empty_cache()

    tree = g.GpuSphTree(params, device=local_rank)
    drv = domain.DomainDriver(tree, rank, world)

    def step():
        drv.BuildTree(True, 0, 1, 1, 0.0)
        drv.UpdateAllSphProperties()
        drv.ZeroAccelerations()
        if grav:
            drv.UpdateAllSphForces()
        else:
            drv.UpdateAllSphHydroForces()

    def barrier():
        tree.Synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # converge h once (untimed), then perturb it like the reference arm does
    tree.Upload(ps, only=IN_FIELDS)
    tree.BuildTree(True, 0, 1, 1, 0.0)
    tree.UpdateAllSphProperties()
    tree.Download(ps, mask=_capi.F_DENSITY, only=("h",))
    ps["h"][...] *= workloads.ref_perturb_factor(np.arange(N), H_PERTURB)
    tree.Upload(ps, only=IN_FIELDS)

    for _ in range(max(args.warmup, 0)):
        step()
    barrier()
    launches0 = tree.Stats().kernel_launches
    sampler = ClockSampler(local_rank)
    sampler.start()
    ksum = {"density": 0.0, "hydro": 0.0, "gravity": 0.0, "tree": 0.0, "tree_sort": 0.0, "tree_levels": 0.0,
            "tree_stock": 0.0, "properties": 0.0, "forces": 0.0}
    t_wall0 = time.perf_counter()
    tree.TimerStart()
    for _ in range(args.steps):
        step()
        st = tree.Stats()
        ksum["density"] += st.ms_kernel_density; ksum["hydro"] += st.ms_kernel_hydro
        ksum["gravity"] += st.ms_kernel_gravity; ksum["tree"] += st.ms_build_tree
        ksum["tree_sort"] += st.ms_tree_sort; ksum["tree_levels"] += st.ms_tree_levels
        ksum["tree_stock"] += st.ms_tree_stock
        ksum["properties"] += st.ms_properties; ksum["forces"] += st.ms_forces
    ms_dev = tree.TimerStop()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.summary()
    st = tree.Stats()
    launches = st.kernel_launches - launches0
    if world > 1:
        t = torch.tensor([ms_dev], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dev = float(t.item())
    ms_step = ms_dev / args.steps
    value = N / (ms_step * 1e-3)

    # ---- end to end: host buffers in, host buffers out, every step -----------------------------------
    DENS_OUT = tuple(k for k in OUT_FIELDS if k in ("h", "rho", "hrangesqd", "hfactor", "invomega", "zeta", "sound",
                                                    "pressure", "u"))
    FORCE_OUT = tuple(k for k in OUT_FIELDS if k not in DENS_OUT)
    FORCE_EARLY = tuple(k for k in FORCE_OUT if k in ("dudt", "div_v", "dalphadt", "levelneib"))
    FORCE_LATE = tuple(k for k in FORCE_OUT if k not in FORCE_EARLY)

    trace = os.environ.get("GSPH_E2E_TRACE") == "1"     # diagnostic: synchronise and time every phase
    marks = []

    def mark(name):
        if trace:
            tree.Synchronize(); torch.cuda.synchronize()
            marks.append((name, time.perf_counter()))

    def e2e_step():
        mark("start")
        # host buffers in, host buffers out; the copies ride the library's copy stream: the tree build
        # starts on r, h while v, m, u are in flight, the density results go home during the force pass
        # with several ranks each one ships / fetches only ITS host-order chunk over PCIe (what an MPI
        # rank of the reference owns); the other chunks travel GPU to GPU (NCCL all-gather)
        drv.Upload(ps, only=IN_FIELDS)
        mark("upload")
        drv.BuildTree(True, 0, 1, 1, 0.0)
        mark("tree")
        drv.UpdateAllSphProperties()
        mark("density")
        drv.Download(pso, mask=_capi.F_DENSITY, only=DENS_OUT, gather=False, wait=False)
        mark("dens_download")
        drv.ZeroAccelerations()
        if grav and world == 1:
            # the neighbour half finishes du/dt, div_v, levelneib: they travel during the tree-gravity half
            drv.UpdateAllSphForcesBegin()
            drv.Download(pso, mask=_capi.F_FORCES, only=FORCE_EARLY, gather=False, wait=False)
            drv.UpdateAllSphForcesFinish()
            drv.Download(pso, mask=_capi.F_FORCES, only=FORCE_LATE, gather=False)
        else:
            if grav:
                drv.UpdateAllSphForces()
            else:
                drv.UpdateAllSphHydroForces()
            mark("forces")
            drv.Download(pso, mask=_capi.F_FORCES, only=FORCE_OUT, gather=True)
        tree.Synchronize()
        mark("force_download")
        if trace and rank == 0:
            sys.stderr.write("e2e trace: " + " ".join("%s %.1f" % (marks[i][0], 1e3*(marks[i][1] - marks[i - 1][1]))
                                                      for i in range(1, len(marks))) + "\n")
        del marks[:]
    # (results land in their own host arrays, so every step starts from the same perturbed h guess)
    e2e_steps = max(1, min(args.steps, 3))
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    t_e2e = (time.perf_counter() - t0) / e2e_steps
    if world > 1:
        t = torch.tensor([t_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t.item())
    e2e = {"value": N / t_e2e, "unit": UNIT, "h2d_bytes_per_step": field_bytes(IN_FIELDS, N, ndim),
           "d2h_bytes_per_step": field_bytes(OUT_FIELDS, N, ndim), "ms_per_step": 1e3 * t_e2e}

    # ---- device-resident main loop: the same operators driven by the on-device lfkdk
    #      integrator, particles never leave HBM.  Each step starts from the previous step's h (no +-3 %
    #      perturbation), so the h-rho iteration does a little less work than in `value`. ---------------
    resident = None
    if not args.no_resident:
        try:
            tree.Upload(ps, only=IN_FIELDS)
            drv.IntegratorBegin(compute_forces=True)
            drv.MainLoop(1)
            barrier()
            k_res = max(1, min(args.steps, 3))
            tree.TimerStart()
            drv.MainLoop(k_res)
            ms_res = tree.TimerStop() / k_res
            if world > 1:
                t = torch.tensor([ms_res], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms_res = float(t.item())
            t_res, ts_res, n_res = tree.IntegratorState()
            resident = {"value": N / (ms_res * 1e-3), "unit": UNIT, "ms_per_step": ms_res, "steps": k_res,
                        "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "hiter_mean": tree.Stats().hiter_mean,
                        "integrator": "lfkdk, one global time step (gsph_integrator_step)", "timestep": ts_res}
        except _capi.GsphError as e:
            resident = {"value": None, "error": str(e)}

    if rank != 0:
        tree.close()
        dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (DESIGN.md "Kernels"): algorithmic bytes per particle ---------
    K = args.steps
    kms = {k: ksum[k] / K for k in ("density", "hydro", "gravity")}
    # density (d+2)s+8 in + 9s out; forces (2d+11)s+20 in + (d+3)s+4 out (+ (d+2)s with gravity); s = 8
    balg = {"density": (ndim + 2)*8.0 + 8 + 72, "hydro": (2*ndim + 11)*8.0 + 20 + (ndim + 3)*8 + 4 + ((ndim + 2)*8 if grav else 0),
            "gravity": 180.0}
    # dram__bytes_read+write per particle from the committed ncu --set full capture of the final code
    # (profiles/r2_ncu_full_cloud16m_final_batches.txt, 16 Mi-particle cloud), scaled to this launch.  Every entry is the
    # whole pass: the pre-walk kernel + the main kernel for density / hydro; for gravity the walk, cells and direct
    # kernels of the three batches of that capture (their interaction lists travel through HBM: 1.2 kB per particle).
    traffic_pp = {"gravity": ((0.343695 + 2.074097 + 0.362384 + 2.087211 + 0.345232 + 2.072511) +
                              (2.576733 + 0.304952 + 2.597583 + 0.305182 + 2.574433 + 0.304954) +
                              (0.890114 + 0.337520 + 0.908109 + 0.339224 + 0.892161 + 0.337667))*1e9/16777216.0,
                  "hydro": (8.880299e9 + 1.848491e9 + 0.657022e9 + 1.546330e9)/16777216.0,
                  "density": (3.672362e9 + 3.160585e9 + 0.375668e9 + 1.619185e9)/16777216.0}
    kernel_names = {"gravity": "gravity_walk_kernel + gravity_cells_kernel + gravity_direct_kernel",
                    "hydro": "hydro_walk_kernel + hydro_kernel", "density": "density_walk_kernel + density_kernel"}
    dom = max(kms, key=lambda k: kms[k])
    n_local = N / world
    hbm, src = peaks()
    ach = balg[dom] * n_local / (kms[dom] * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": kernel_names[dom], "achieved": ach, "peak": hbm, "unit": "GB/s",
                "frac": ach / hbm, "traffic": (traffic_pp[dom]*n_local/1e9 if (dom in traffic_pp and grav) else None),
                "traffic_unit": "GB per launch (ncu capture at 16 Mi particles, scaled by N)", "peak_source": src,
                "alg_bytes_per_particle": balg[dom], "ms_per_launch": kms[dom]}
    # FP64 pipe view (the path is FP64-bound): flops from the measured interaction counts
    ngb, nit = st.ngb_mean, st.hiter_mean
    ndir = max(st.ndirect_mean - ngb, 0.0)
    f_dens = ngb * 19.0 * nit
    f_hyd = ngb * (59.0 + (36.0 if grav else 0.0))
    nleaf = N / float(1 << tree.TreeInfo().ltot)
    f_grav = (ndir * 20.0 + st.ncell_mean * 30.0 / nleaf + 25.0) if grav else 0.0
    fp64_peak, fp32_peak = sphtree.measure_fp_peak(local_rank)
    t_kern = (kms["density"] + kms["hydro"] + kms["gravity"]) * 1e-3
    fp = {"bound": "fp64", "achieved": (f_dens + f_hyd + f_grav) * n_local / t_kern / 1e12, "peak": fp64_peak,
          "unit": "TFLOP/s", "peak_source": "gsph_measure_fp_peak (register-resident DFMA chains, this run)",
          "flops_per_particle": {"density": f_dens, "hydro": f_hyd, "gravity": f_grav}}
    fp["frac"] = fp["achieved"] / fp64_peak if fp64_peak else None

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.workload, N), "l2": "inputs_larger_than_l2",
                   "h_guess": "converged h perturbed by +-%g%% (both arms)" % (100 * H_PERTURB),
                   "parallelism": "tree-order slices x%d, replicated tree, NCCL all-gather after density" % world
                   if world > 1 else "single GPU"},
        "clocks": clocks, "e2e": e2e, "resident": resident, "gpu_launches": int(launches),
        "roofline": roofline, "roofline_fp64": fp,
        "blocks_ms": {"BUILD_TREE": ksum["tree"] / K, "tree_sort": ksum["tree_sort"] / K,
                      "tree_levels": ksum["tree_levels"] / K, "tree_stock": ksum["tree_stock"] / K,
                      "SPH_PROPERTIES": ksum["properties"] / K, "SPH_FORCES": ksum["forces"] / K,
                      "density_kernel": kms["density"], "hydro_kernel": kms["hydro"],
                      "gravity_kernel": kms["gravity"], "wall_ms_per_step": 1e3 * t_wall / K},
        "stats": {"N": N, "ngb_mean": st.ngb_mean, "ngb_max": st.ngb_max, "cand_per_ngb": st.cand_mean / max(ngb, 1e-9),
                  "hiter_mean": st.hiter_mean, "hiter_max": st.hiter_max, "ndirect_mean": st.ndirect_mean,
                  "ncell_mean": st.ncell_mean, "density_retries": st.density_retries,
                  "fp32_peak_tflops": fp32_peak},
    }
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        b = run_reference_sample(args.workload, args.cpu_particles or CPU_SAMPLE[args.workload], 2, threads)
        if b is not None:
            line["cpu_baseline"] = {"value": b["updates_per_s"], "unit": UNIT, "cores": threads, "kind": "reference",
                                    "sample": "%d-particle %s, 2 passes of the unmodified reference (OpenMP, -O3, "
                                              "IEEE arithmetic), %s" % (b["N"], args.workload, cpu_model())}
        else:
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": threads, "kind": "reference",
                                    "sample": "oracle/_ref/ref_harness_omp missing"}
    print(json.dumps(line))
    tree.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cloud", choices=["cloud", "glass", "khi", "bossbodenheimer"])
    ap.add_argument("--particles", type=int, default=0)
    ap.add_argument("--cpu-particles", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-resident", action="store_true", help="skip the device-resident main-loop measurement")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return main_reference(args, rank, world)
    return main_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    sys.exit(main())
