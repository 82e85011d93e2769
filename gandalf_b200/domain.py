"""Multi-GPU driver of the path: one process per GPU, torch.distributed (NCCL over NVLink) for
the one exchange step per pass.

Replaces the MPI layer around the path (src/Mpi/MpiControl.cpp:1002-1157 ghost exchange,
src/Mpi/MpiKDTreeDecomposition.cpp:56-277 domain decomposition):

 * decomposition: rank r owns the tree-order slice [r*chunk, (r+1)*chunk) -- the leaves of the
   balanced KD-tree below the r-th node of level log2(R), i.e. the reference's median-split KD
   domains, with no separate decomposition step;
 * every rank holds the (replicated) tree and all particle positions, so the density pass needs
   no halo import at all; the force pass needs the post-density fields of the other slices:
   ONE in-place ``all_gather_into_tensor`` per field on the device SoA arrays (4 doubles per
   particle instead of the reference's 448-byte structs);
 * results are identical to the single-GPU run (same tree, same kernels, same summation order).
"""
import numpy as np


def slice_bounds(N, world_size, rank):
    """tree-order slice of `rank`; chunk = ceil(N/R) so that every slice but the last is full."""
    chunk = -(-int(N) // int(world_size))
    lo = min(rank * chunk, N)
    hi = min(lo + chunk, N)
    return lo, hi, chunk


def exchange_plan(N, world_size):
    """[(lo, hi)] per rank plus the padded length the in-place all-gather works on."""
    chunk = -(-int(N) // int(world_size))
    return [slice_bounds(N, world_size, r)[:2] for r in range(world_size)], chunk * world_size


class _CudaArray(object):
    """__cuda_array_interface__ holder so torch can wrap a libgsph device array without a copy."""

    def __init__(self, ptr, n, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def shard_ranges(N, world_size):
    """[(lo, hi)] of the nodes of level log2(world_size) of the balanced KD-tree (sizes N -> N//2,
    N - N//2, src/Tree/KDTree.cpp:545-560): the reference's median-split KD domains."""
    out = [(0, int(N))]
    while len(out) < world_size:
        nxt = []
        for lo, hi in out:
            left = (hi - lo) // 2
            nxt += [(lo, lo + left), (lo + left, hi)]
        out = nxt
    return out


class DomainDriver(object):
    """Runs BuildTree / UpdateAllSphProperties / UpdateAllSph(Hydro)Forces across the ranks of a
    torch.distributed process group (world_size 1 = plain single-GPU calls)."""

    def __init__(self, tree, rank=0, world_size=1, group=None):
        self.tree, self.rank, self.world, self.group = tree, rank, world_size, group
        self.shard_build = True      # split the tree build's re-ordering passes across the ranks
        self.N = 0
        self._fields = None

    def _sharded_rebuild(self, parts):
        """Full rebuild with the per-level re-ordering work split across the ranks: every rank builds
        the top log2(R) levels and its own subtree, then the slices of the tree order are
        all-gathered (4 bytes per particle) and every rank finishes the (identical) tree."""
        import torch
        import torch.distributed as dist
        if parts is not None:
            self.tree.Upload(parts)
        self.tree.BuildTreeShardBegin(self.rank, self.world)
        N = self.tree.Stats().Nhydro
        ranges = shard_ranges(N, self.world)
        maxc = max(hi - lo for lo, hi in ranges)
        ptr, cap = self.tree.DeviceField("perm")
        perm = torch.as_tensor(_CudaArray(ptr, N, "<i4"), device="cuda")
        lo, hi = ranges[self.rank]
        if all(b - a == maxc for a, b in ranges):
            # equal shards (N divisible by the rank count): gather straight into the tree order, in place
            dist.all_gather_into_tensor(perm, perm[lo:hi], group=self.group)
            torch.cuda.current_stream().synchronize()
            self.tree.BuildTreeShardFinish()
            return
        send = torch.zeros(maxc, dtype=torch.int32, device=perm.device)
        send[:hi - lo] = perm[lo:hi]
        recv = torch.empty(maxc * self.world, dtype=torch.int32, device=perm.device)
        dist.all_gather_into_tensor(recv, send, group=self.group)
        for r, (a, b) in enumerate(ranges):
            if r != self.rank:
                perm[a:b] = recv[r * maxc:r * maxc + (b - a)]
        torch.cuda.current_stream().synchronize()
        self.tree.BuildTreeShardFinish()

    def BuildTree(self, rebuild, n, ntreebuildstep, ntreestockstep, timestep, parts=None):
        full = rebuild or n % ntreebuildstep == 0
        self._build(0 if full else (1 if n % ntreestockstep == 0 else 2), timestep, parts)

    def _build(self, mode, timestep, parts=None):
        pow2 = self.world > 1 and (self.world & (self.world - 1)) == 0
        if mode == 0 and pow2 and self.shard_build:
            self._sharded_rebuild(parts)
        else:
            if parts is not None:
                self.tree.Upload(parts)
            self.tree.BuildTreeMode(mode, timestep)
        self.N = self.tree.Stats().Nhydro
        if self.world > 1:
            lo, hi, chunk = slice_bounds(self.N, self.world, self.rank)
            self.lo, self.hi, self.chunk = lo, hi, chunk
            self.tree.SetActiveSlice(lo, hi)
            self._fields = None

    def _wrap_fields(self):
        import torch
        if self._fields is None:
            out = []
            # hrangesqd, hfactor, u, sound, pressure follow from these on the receiving side
            names = ["h", "rho", "invomega", "zeta"]
            if self.tree.cfg.tdavisc == 2:
                names.append("alpha")        # the Cullen & Dehnen switch sets alpha in the density pass
            if self.tree.cfg.create_sinks:
                names.append("flags")        # potmin flag of the sink search
            for name in names:
                ptr, cap, es = self.tree.DeviceArray(name, staging=False)
                if cap < self.chunk * self.world:
                    raise RuntimeError("device arrays too small for the padded all-gather")
                out.append(torch.as_tensor(_CudaArray(ptr, self.chunk * self.world, "<f8" if es == 8 else "<i4"),
                                           device="cuda"))
            self._fields = out
        return self._fields

    def UpdateAllSphProperties(self):
        if self.world == 1:
            self.tree.UpdateAllSphProperties()
            return
        import torch.distributed as dist
        import torch
        self.tree.DensityPass()
        self.tree.Synchronize()
        fields = self._wrap_fields()           # h first
        sl = slice(self.rank * self.chunk, (self.rank + 1) * self.chunk)
        # in-place all-gathers (input = this rank's chunk of the output).  The smoothing lengths go first: the
        # hmax / h-box refresh of the tree (KDTree::UpdateAllHmaxValues) only needs them and runs on the library's
        # stream while rho, invomega, zeta are still crossing NVLink
        dist.all_gather_into_tensor(fields[0], fields[0][sl], group=self.group)
        torch.cuda.current_stream().synchronize()
        self.tree.RefreshHmax()
        for t in fields[1:]:
            dist.all_gather_into_tensor(t, t[sl], group=self.group)
        torch.cuda.current_stream().synchronize()
        self.tree.DeriveDensityFields(self.lo, self.hi)

    # ---- host <-> device for a multi-process host: every rank owns the host-order chunk ---------------
    #      [rank*ceil(N/R), ...) of the particle set (what an MPI rank owns in the reference); it ships
    #      and fetches only that chunk over PCIe, the rest travels GPU to GPU over NVLink.
    _COMPONENTS = {"r": ("rx", "ry", "rz"), "v": ("vx", "vy", "vz"), "a": ("ax", "ay", "az"),
                   "atree": ("atx", "aty", "atz")}

    def _array_names(self, fields, ndim):
        out = []
        for f in fields:
            out += list(self._COMPONENTS[f][:ndim]) if f in self._COMPONENTS else [f]
        return out

    def _gather_inplace(self, names, staging, n_valid):
        import torch
        import torch.distributed as dist
        chunk = -(-int(n_valid) // self.world)
        for name in names:
            ptr, cap, es = self.tree.DeviceArray(name, staging=staging)
            if cap < chunk * self.world:
                raise RuntimeError("device arrays too small for the padded all-gather")
            t = torch.as_tensor(_CudaArray(ptr, chunk * self.world, "<f8" if es == 8 else "<i4"), device="cuda")
            dist.all_gather_into_tensor(t, t[self.rank * chunk:(self.rank + 1) * chunk], group=self.group)
        torch.cuda.current_stream().synchronize()

    def Upload(self, parts, only):
        """inputs of one step: this rank's chunk over PCIe + one NVLink all-gather per array"""
        if self.world == 1:
            self.tree.Upload(parts, only=only)
            return
        lo, hi, _ = slice_bounds(parts.N, self.world, self.rank)
        self.tree.UploadRange(parts, lo, hi, only=only)
        self.tree.Synchronize()
        self._gather_inplace(self._array_names(only, parts.ndim), True, parts.N)

    def Download(self, parts, mask, only, gather=True, wait=True):
        """results of this rank's chunk; gather=True first all-gathers the tree-order result slices
        (the force fields; the density fields are already complete on every rank)"""
        if self.world == 1:
            self.tree.Download(parts, mask=mask, only=only, wait=wait)
            return
        if gather:
            self.tree.Synchronize()
            self._gather_inplace(self._array_names(only, parts.ndim), False, parts.N)
        lo, hi, _ = slice_bounds(parts.N, self.world, self.rank)
        self.tree.DownloadRange(parts, lo, hi, mask=mask, only=only, wait=wait)

    # ---- device-resident main loop across the ranks (gsph_integrator_*): the integrator state is
    #      replicated (streaming kernels over all particles, identical on every rank: the global time
    #      step is an order-independent minimum); between its two halves the ranks run the sharded
    #      tree build, their slices of the passes and the exchanges.  Nothing crosses PCIe.
    def _block(self):
        """hierarchical block time steps (Nlevels > 1, Simulation::ComputeBlockTimesteps)?"""
        return int(self.tree.params["Nlevels"]) > 1

    def _passes(self):
        block = self._block()
        self.UpdateAllSphProperties()
        self.ZeroAccelerations()
        if block:
            # nradstep = 1: thermal properties of every particle, active or not (SphSimulation.cpp:669-679)
            self.tree.UpdateThermalProperties()
        if self.tree.cfg.self_gravity:
            self.UpdateAllSphForces()
        else:
            self.UpdateAllSphHydroForces()
        if self.world > 1:
            # the force results of the other slices (tree order): one in-place all-gather per array
            fields = ["a", "atree", "gpot", "gpot_hydro", "dudt", "div_v", "dalphadt", "alpha"]
            if not block:
                fields.append("levelneib")
            self.tree.Synchronize()
            self._gather_inplace(self._array_names(fields, self.tree.cfg.ndim), False, self.N)
            if block:
                # with inactive particles the levelneib scatter of a rank's targets lands on every slice
                # (GradhSph.cpp:446, GradhSphTree.cpp:376-382): MAX over the ranks instead of a gather
                self._reduce_max_inplace("levelneib", self.N)

    def _reduce_max_inplace(self, name, n_valid):
        import torch
        import torch.distributed as dist
        ptr, cap, es = self.tree.DeviceArray(name, staging=False)
        t = torch.as_tensor(_CudaArray(ptr, int(n_valid), "<i4" if es == 4 else "<f8"), device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        torch.cuda.current_stream().synchronize()

    def IntegratorBegin(self, parts=None, compute_forces=True, t=0.0):
        if self.world == 1:
            self.tree.IntegratorBegin(parts, compute_forces, t)
            return
        if parts is not None:
            self.tree.Upload(parts)
        if compute_forces:
            self._build(0, 0.0)
            self._passes()
        self.tree.IntegratorBegin(compute_forces=2 if compute_forces else 0, t=t)

    def MainLoop(self, nsteps=1):
        if self.world == 1:
            self.tree.MainLoop(nsteps)
            return
        for _ in range(int(nsteps)):
            mode, ts = self.tree.IntegratorAdvance()
            self._build(mode, ts)
            self._passes()
            if self._block():
                # the passes are repeated for the particles whose step SphLeapfrogKDK::CheckTimesteps cuts
                # short (the do-while loop of MainLoop, SphSimulation.cpp:658-756); the check runs on the
                # replicated host-order state, so every rank sees the same count
                for it in range(66):
                    if self.tree.IntegratorCheckTimesteps() == 0:
                        break
                    if it == 65:
                        raise RuntimeError("the neighbour time-step limiter does not settle")
                    self._passes()
            self.tree.IntegratorFinish()

    def ZeroAccelerations(self):
        self.tree.ZeroAccelerations()

    def UpdateAllSphHydroForces(self):
        self.tree.UpdateAllSphHydroForces()

    def UpdateAllSphForces(self):
        self.tree.UpdateAllSphForces()

    def UpdateAllSphForcesBegin(self):
        self.tree.UpdateAllSphForcesBegin()

    def UpdateAllSphForcesFinish(self):
        self.tree.UpdateAllSphForcesFinish()
