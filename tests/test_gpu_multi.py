"""Multi-GPU parity (stage P7): the tree-order-slice decomposition with the NCCL all-gather exchange
gives the SAME bits as the single-GPU run.  Needs >= 2 GPUs (skipped on the 1-GPU test box; run with
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["GSPH_ROOT"])
import gandalf_b200 as g
from gandalf_b200 import workloads, domain
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda:%d" % lr))
for workload, NP in (("cloud", 150000), ("glass", 64**3), ("glass+cd2010", 40**3)):
    params, ic = workloads.WORKLOADS[workload.split("+")[0]](NP, device="cuda:%d" % lr)
    if workload.endswith("cd2010"):                    # alpha is set in the density pass and read by the neighbours
        params = dict(params, time_dependent_avisc="cd2010", alpha_visc_min=0.1)
    N = int(ic["m"].shape[0])
    grav = int(params["self_gravity"]) == 1
    def fresh():
        ps = g.ParticleSet(N, 3)
        for k in ("r", "v"):
            ps[k][...] = ic[k].T.cpu().numpy()
        for k in ("m", "h", "u"):
            ps[k][...] = ic[k].cpu().numpy()
        return ps
    def run(world_, rank_):
        ps = fresh()
        tree = g.GpuSphTree(params, device=lr)
        drv = domain.DomainDriver(tree, rank_, world_)
        tree.Upload(ps)
        drv.BuildTree(True, 0, 1, 1, 0.0)
        drv.UpdateAllSphProperties(); drv.ZeroAccelerations()
        drv.UpdateAllSphForces() if grav else drv.UpdateAllSphHydroForces()
        tree.Download(ps)
        st = tree.Stats(); tree.close()
        return ps, st
    def run_chunked(world_, rank_):
        # multi-process host: each rank ships / fetches only its host-order chunk (gsph_upload_range,
        # gsph_download_range), the rest travels over NCCL
        ps = fresh()
        lo, hi, _ = domain.slice_bounds(N, world_, rank_)
        for k in ("r", "v", "m", "h", "u"):            # this rank only KNOWS its own chunk
            keep = ps[k][lo:hi].copy(); ps[k][...] = np.nan; ps[k][lo:hi] = keep
        tree = g.GpuSphTree(params, device=lr)
        drv = domain.DomainDriver(tree, rank_, world_)
        drv.Upload(ps, only=("r", "v", "m", "h", "u"))
        drv.BuildTree(True, 0, 1, 1, 0.0)
        drv.UpdateAllSphProperties()
        drv.Download(ps, mask=1, only=("h", "rho", "hrangesqd"), gather=False, wait=False)
        drv.ZeroAccelerations()
        drv.UpdateAllSphForces() if grav else drv.UpdateAllSphHydroForces()
        drv.Download(ps, mask=2, only=("a", "dudt", "div_v", "levelneib") + (("atree", "gpot") if grav else ()), gather=True)
        tree.Synchronize(); tree.close()
        return ps, lo, hi
    def run_loop(world_, rank_, steps, **extra):
        # device-resident main loop: replicated integrator state, sharded passes (DomainDriver.MainLoop)
        ps = fresh()
        tree = g.GpuSphTree(dict(params, **extra), device=lr)
        drv = domain.DomainDriver(tree, rank_, world_)
        drv.IntegratorBegin(ps, compute_forces=True)
        drv.MainLoop(steps)
        tree.Download(ps)
        state = tree.IntegratorState(); tree.close()
        return ps, state
    def run_block(world_, rank_, steps):
        # hierarchical block time steps across the ranks: replicated hierarchy, sharded passes repeated for the
        # particles CheckTimesteps wakes, levelneib combined by a MAX reduction
        ps = fresh()
        tree = g.GpuSphTree(dict(params, Nlevels=5), device=lr)
        drv = domain.DomainDriver(tree, rank_, world_)
        drv.IntegratorBegin(ps, compute_forces=True)
        drv.MainLoop(steps)
        tree.Download(ps)
        state, blk = tree.IntegratorState(), tree.IntegratorBlockState(arrays=True)
        nact = tree.Stats().n_active
        tree.close()
        return ps, state, blk
    def run_stars(world_, rank_, steps, nlevels):
        # star particles (the reference's default Hermite-4 integrator on the host of every rank, gas-on-star forces
        # from the replicated tree) with one global step or on block time steps of their own
        ps = fresh()
        tree = g.GpuSphTree(dict(params, nbody="hermite4", Nlevels=nlevels), device=lr)
        drv = domain.DomainDriver(tree, rank_, world_)
        tree.IntegratorSetStars(np.array([[0.1, 0.05, -0.08], [1.4, -0.3, 0.2]]), np.zeros((2, 3)),
                                np.array([0.05, 0.1]), np.array([0.05, 0.08]))
        drv.IntegratorBegin(ps, compute_forces=True)
        drv.MainLoop(steps)
        tree.Download(ps)
        st, state = tree.IntegratorGetStars(), tree.IntegratorState()
        tree.close()
        return ps, st, state
    if workload == "cloud":
        sres = [run_stars(world, rank, 4, nl) for nl in (1, 4)]
        # rebuild / restock / extrapolate schedule of HydroTree::BuildTree across the ranks (only rebuilds are sharded)
        mixloop, mixstate = run_loop(world, rank, 6, ntreebuildstep=3, ntreestockstep=2)
    if workload in ("cloud", "glass"):
        bloop, bstate, bblk = run_block(world, rank, 7)
    loop, lstate = run_loop(world, rank, 3)
    chunked, clo, chi = run_chunked(world, rank)
    multi, st = run(world, rank)
    # forces: every rank only wrote its own slice (other particles keep the uploaded zeros): sum over ranks
    tot = {}
    for f in ("a", "atree", "gpot", "dudt", "div_v"):
        t = torch.from_numpy(np.ascontiguousarray(multi[f])).cuda()
        dist.all_reduce(t)
        tot[f] = t.cpu().numpy()
    nact = torch.tensor([st.n_active], device="cuda"); dist.all_reduce(nact)
    assert int(nact.item()) == N, (int(nact.item()), N)
    if rank == 0:
        single, _ = run(1, 0)
        for f in ("h", "rho", "invomega", "zeta", "pressure", "sound", "hrangesqd", "alpha"):
            assert np.array_equal(multi[f], single[f]), (workload, f)
        for f in tot:
            assert np.array_equal(tot[f], single[f]), (workload, f, np.abs(tot[f] - single[f]).max())
        for f in ("h", "rho", "hrangesqd", "a", "dudt", "div_v") + (("atree", "gpot") if grav else ()):
            assert np.array_equal(chunked[f][clo:chi], single[f][clo:chi]), (workload, "chunked", f)
        sloop, sstate = run_loop(1, 0, 3)
        assert lstate == sstate, (lstate, sstate)
        for f in ("r", "v", "a", "h", "rho", "u", "dudt") + (("atree", "gpot") if grav else ()):
            assert np.array_equal(loop[f], sloop[f]), (workload, "mainloop", f, np.abs(loop[f] - sloop[f]).max())
        if workload in ("cloud", "glass"):
            sb, sbstate, sbblk = run_block(1, 0, 7)
            assert bstate == sbstate, (bstate, sbstate)
            for k in ("n", "nresync", "level_max", "level_step", "dt_max"):
                assert bblk[k] == sbblk[k], (workload, "block", k, bblk[k], sbblk[k])
            assert len(np.unique(sb["level"])) > 1, "the block-step case has one level only"
            for k in ("nstep", "nlast", "dt", "tlast"):
                assert np.array_equal(bblk[k], sbblk[k]), (workload, "block", k)
            for f in ("r", "v", "a", "h", "rho", "u", "dudt", "level", "levelneib") + (("atree", "gpot") if grav else ()):
                assert np.array_equal(bloop[f], sb[f]), (workload, "block", f, np.abs(bloop[f] - sb[f]).max())
            print("multi-GPU block time steps ok:", workload, "levels", np.unique(sb["level"]).tolist(), "n", sbblk["n"])
        if workload == "cloud":
            for (mp, mst, mstate), nl in zip(sres, (1, 4)):
                sp, sst, sstate = run_stars(1, 0, 4, nl)
                assert mstate == sstate, (nl, mstate, sstate)
                for k in sst:
                    assert np.array_equal(mst[k], sst[k]), ("stars", nl, k)
                for f in ("r", "v", "a", "h", "rho", "atree", "gpot"):
                    assert np.array_equal(mp[f], sp[f]), ("stars", nl, f)
            print("multi-GPU star particles ok (hermite4, Nlevels 1 and 4)")
            smix, smixstate = run_loop(1, 0, 6, ntreebuildstep=3, ntreestockstep=2)
            assert mixstate == smixstate, (mixstate, smixstate)
            for f in ("r", "v", "a", "h", "rho", "atree", "gpot"):
                assert np.array_equal(mixloop[f], smix[f]), ("tree schedule", f)
        print("multi-GPU parity ok:", workload, N, "particles, world", world)
dist.barrier()
"""


def test_two_gpu_results_identical_to_single_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    w = tmp_path / "worker.py"
    w.write_text(_WORKER)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    n = min(torch.cuda.device_count(), 4)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % n,
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(w)],
                         env=dict(os.environ, GSPH_ROOT=ROOT), stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                         timeout=1200)
    text = out.stdout.decode()
    assert out.returncode == 0, text[-3000:]
    assert text.count("multi-GPU parity ok") == 3, text[-3000:]
