/*=================================================================================================
 *  gravity.cu -- tree gravity of the hydro+gravity force pass (everything outside kernel overlap).
 *
 *  Replaces  Tree::ComputeGravityInteractionAndGhostList + open_cell_for_gravity
 *                                                  src/Tree/Tree.cpp:628-735, src/Headers/Tree.h:413-432
 *            GradhSphBase::ComputeDirectGravForces  src/GradhSph/GradhSph.cpp:657-690
 *            ComputeCellMonopoleForces / ComputeCellQuadrupoleForces / ComputeQuadropole
 *                                                  src/Headers/NeighbourSearch.h:349-475
 *            FastMultipoleForces, ComputeFast{Monopole,Quadrupole}Forces
 *                                                  src/Headers/NeighbourSearch.h:481-827
 *
 *  The reference decides the interaction list PER LEAF CELL (geometric MAC against the leaf's
 *  rcell/rmax/hmax) and, for fast multipoles, expands about the leaf's centre of mass; to give the
 *  same forces the GPU keeps that granularity: one warp per leaf.  The warp walks the tree
 *  cooperatively (32 nodes per step), writing accepted cells and direct-sum particles to its
 *  scratch lists; then the LANES run over the list entries, each lane accumulating the
 *  contributions to all (<= 8) particles of the leaf in registers, and warp-shuffle reductions
 *  produce the per-particle sums.
 *
 *  Split with hydro.cu: a particle j of an opened leaf is a kernel-overlap neighbour of target i
 *  iff the reference's predicate holds (NeighbourManager.h:521); those pairs get the softened
 *  terms in hydro.cu.  Here j contributes the unsoftened term iff the predicate fails -- the same
 *  partition NeighbourManager::TrimNeighbourLists makes (NeighbourManager.h:503-542).
 *=================================================================================================*/
#include "device_common.cuh"

#define GRAV_THREADS 128
#define GRAV_WARPS (GRAV_THREADS/32)
#define GRAV_NT 8                      /* max particles per leaf on this path */

namespace {

struct GravVisitor {
  const TreeDev& T; const DevParams& dp;
  double rc[3], rmax, hrangemax;
  int32_t* cells; int ncells, cap_cells;
  CandList& direct;
  DevStatus* st;

  __device__ __forceinline__ int test(int node, bool leaf) const
  {
    double dr[3];
#pragma unroll
    for (int k = 0; k < 3; k++) dr[k] = __dadd_rn(T.rcell[k][node], -rc[k]);
    const double drsqd = dot_exact<3>(dr);
    const double crmax = T.rmax[node], chmax = T.hmax[node];
    const double a = __dadd_rn(__dadd_rn(crmax, rmax), hrangemax);
    const double b = __dadd_rn(__dadd_rn(rmax, crmax), __dmul_rn(dp.kernrange, chmax));
    if (drsqd <= __dmul_rn(a, a) || drsqd <= __dmul_rn(b, b)) return 1;       /* overlap: open  */
    if (!(drsqd < T.cdistsqd[node])) return (leaf && T.count[node] == 1) ? 3 : 2;  /* MAC accepts */
    return 1;                                                                  /* MAC opens      */
  }
  __device__ __forceinline__ bool emit(int node, int code, bool leaf, int lane)
  {
    const unsigned cm = __ballot_sync(FULLMASK, code == 2);
    if (cm) {
      const int n = __popc(cm);
      if (ncells + n > cap_cells) {
        if (lane == 0) { atomicExch(&st->err_cand_overflow, 2); }
        return false;
      }
      if (code == 2) cells[ncells + __popc(cm & ((1u << lane) - 1u))] = node;
      ncells += n;
    }
    return direct.append_leaf(T, node, (code == 1 && leaf) || code == 3, lane);
  }
};

struct FastMp { double ac[3], dphi[3], q[6], pot; };

template <int MULTIPOLE>
__global__ void __launch_bounds__(GRAV_THREADS)
gravity_kernel(PartView P, TreeDev tree, DevParams dp, int* work_counter, int32_t* lists_all,
               int cap_cells, int cap_direct, DevStatus* st, uint8_t* done)
{
  extern __shared__ int smem_i[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int* stack = smem_i + warp*WALK_STACK_CAP;
  const int gwarp = blockIdx.x*GRAV_WARPS + warp;
  int32_t* cells = lists_all + (size_t) gwarp*(cap_cells + cap_direct);
  int32_t* dlist = cells + cap_cells;
  const int nleaf = 1 << tree.ltot;
  unsigned long long loc_direct = 0, loc_cells = 0, loc_leaf = 0;

  while (true) {
    int g = 0;
    if (lane == 0) g = atomicAdd(work_counter, 1);
    g = __shfl_sync(FULLMASK, g, 0);
    if (g >= nleaf) break;
    const int node = nleaf - 1 + g;
    const int first = tree.first[node], cnt = tree.count[node];
    if (cnt == 0 || tree.nactive[node] == 0 || done[g]) continue;

    CandList dl{dlist, cap_direct, 0, st};
    GravVisitor V{tree, dp, {tree.rcell[0][node], tree.rcell[1][node], tree.rcell[2][node]},
                  tree.rmax[node], __dmul_rn(dp.kernrange, tree.hmax[node]), cells, 0, cap_cells, dl, st};
    if (!warp_walk(tree, stack, V, st)) continue;
    const int ncells = V.ncells, ndirect = dl.n;
    __syncwarp();

    /* targets (uniform across the warp) */
    double x[GRAV_NT][3], hr2[GRAV_NT];
    double ax[GRAV_NT], ay[GRAV_NT], az[GRAV_NT], pot[GRAV_NT];
#pragma unroll
    for (int t = 0; t < GRAV_NT; t++) {
      const int i = first + (t < cnt ? t : 0);
      x[t][0] = P.f[F_RX][i]; x[t][1] = P.f[F_RY][i]; x[t][2] = P.f[F_RZ][i];
      hr2[t] = P.f[F_HR2][i];
      ax[t] = ay[t] = az[t] = pot[t] = 0.0;
    }

    /* ---- direct sum, ComputeDirectGravForces GradhSph.cpp:657-690 ------------------------------*/
    for (int k = lane; k < ndirect; k += 32) {
      const int j = dlist[k];
      if (((unsigned) P.i[I_FLAGS][j]) & GSPH_FLAG_DEAD) continue;
      const double xj = P.f[F_RX][j], yj = P.f[F_RY][j], zj = P.f[F_RZ][j];
      const double mj = P.f[F_M][j], hr2j = P.f[F_HR2][j];
#pragma unroll
      for (int t = 0; t < GRAV_NT; t++) {
        if (t >= cnt) break;
        double dr[3] = {xj - x[t][0], yj - x[t][1], zj - x[t][2]};
        const double d2 = dot_exact<3>(dr);
        if (d2 < hr2[t] || d2 < hr2j) continue;            /* kernel-overlap pair: hydro.cu */
        const double inv = rsqrt(d2 + GSPH_SMALL);
        const double f = mj*inv*inv*inv;
        ax[t] += f*dr[0]; ay[t] += f*dr[1]; az[t] += f*dr[2];
        pot[t] += mj*inv;
      }
    }

    /* ---- cells ------------------------------------------------------------------------------------*/
    if (MULTIPOLE == GSPH_MULTIPOLE_MONOPOLE || MULTIPOLE == GSPH_MULTIPOLE_QUADRUPOLE) {
      for (int k = lane; k < ncells; k += 32) {
        const int c = cells[k];
        const double cx = tree.com[0][c], cy = tree.com[1][c], cz = tree.com[2][c], cm = tree.m[c];
        double q0 = 0, q1 = 0, q2 = 0, q3 = 0, q4 = 0;
        if (MULTIPOLE == GSPH_MULTIPOLE_QUADRUPOLE) {
          q0 = tree.q[0][c]; q1 = tree.q[1][c]; q2 = tree.q[2][c]; q3 = tree.q[3][c]; q4 = tree.q[4][c];
        }
#pragma unroll
        for (int t = 0; t < GRAV_NT; t++) {
          if (t >= cnt) break;
          if (MULTIPOLE == GSPH_MULTIPOLE_MONOPOLE) {            /* NeighbourSearch.h:349-377 */
            const double dx = cx - x[t][0], dy = cy - x[t][1], dz = cz - x[t][2];
            const double d2 = dx*dx + dy*dy + dz*dz + GSPH_SMALL;
            const double inv = rsqrt(d2);
            const double f = cm*inv*inv*inv;
            pot[t] += cm*inv;
            ax[t] += f*dx; ay[t] += f*dy; az[t] += f*dz;
          }
          else {                                                  /* NeighbourSearch.h:384-475 */
            const double dx = x[t][0] - cx, dy = x[t][1] - cy, dz = x[t][2] - cz;
            const double d2 = dx*dx + dy*dy + dz*dz + GSPH_SMALL;
            const double inv = rsqrt(d2);
            const double inv2 = inv*inv;
            const double inv5 = inv2*inv2*inv;
            const double f = cm*inv2*inv;
            const double qs = q0*dx*dx + q2*dy*dy - (q0 + q2)*dz*dz + 2.0*(q1*dx*dy + q3*dx*dz + q4*dy*dz);
            const double qf = 2.5*qs*inv5*inv2;
            ax[t] += (q0*dx + q1*dy + q3*dz)*inv5 - qf*dx - f*dx;
            ay[t] += (q1*dx + q2*dy + q4*dz)*inv5 - qf*dy - f*dy;
            az[t] += (q3*dx + q4*dy - (q0 + q2)*dz)*inv5 - qf*dz - f*dz;
            pot[t] += cm*inv + 0.5*qs*inv5;
          }
        }
      }
    }
    else {
      /* fast multipoles: Taylor expansion about the leaf's centre of mass, HoT = false
         (NeighbourSearch.h:481-760); applied to each particle below (:768-827).           */
      const double rcx = tree.com[0][node], rcy = tree.com[1][node], rcz = tree.com[2][node];
      FastMp F;
      F.pot = 0.0;
#pragma unroll
      for (int k = 0; k < 3; k++) { F.ac[k] = 0.0; F.dphi[k] = 0.0; }
#pragma unroll
      for (int k = 0; k < 6; k++) F.q[k] = 0.0;
      for (int k = lane; k < ncells; k += 32) {
        const int c = cells[k];
        const double cm = tree.m[c];
        {
          const double dx = tree.com[0][c] - rcx, dy = tree.com[1][c] - rcy, dz = tree.com[2][c] - rcz;
          const double inv = rsqrt(dx*dx + dy*dy + dz*dz);
          const double inv2 = inv*inv;
          const double mc = cm*inv2*inv;
          F.pot += cm*inv;
          F.ac[0] += mc*dx; F.ac[1] += mc*dy; F.ac[2] += mc*dz;
          F.dphi[0] += mc*dx; F.dphi[1] += mc*dy; F.dphi[2] += mc*dz;
          F.q[0] += mc*(3.0*dx*dx*inv2 - 1.0);
          F.q[1] += mc*(3.0*dx*dy*inv2);
          F.q[2] += mc*(3.0*dy*dy*inv2 - 1.0);
          F.q[3] += mc*(3.0*dz*dx*inv2);
          F.q[4] += mc*(3.0*dz*dy*inv2);
          F.q[5] += mc*(3.0*dz*dz*inv2 - 1.0);
        }
        if (MULTIPOLE == GSPH_MULTIPOLE_FAST_QUADRUPOLE) {
          const double q0 = tree.q[0][c], q1 = tree.q[1][c], q2 = tree.q[2][c], q3 = tree.q[3][c],
                       q4 = tree.q[4][c];
          const double dx = rcx - tree.com[0][c], dy = rcy - tree.com[1][c], dz = rcz - tree.com[2][c];
          const double d2 = dx*dx + dy*dy + dz*dz + GSPH_SMALL;
          const double inv = rsqrt(d2);
          const double inv2 = inv*inv;
          const double inv5 = inv2*inv2*inv;
          const double qs = q0*dx*dx + q2*dy*dy - (q0 + q2)*dz*dz + 2.0*(q1*dx*dy + q3*dx*dz + q4*dy*dz);
          const double qf = 2.5*qs*inv5*inv2;
          double qx0 = (q0*dx + q1*dy + q3*dz)*inv5;
          double qx1 = (q1*dx + q2*dy + q4*dz)*inv5;
          double qx2 = (q3*dx + q4*dy - (q0 + q2)*dz)*inv5;
          F.pot += 0.5*qs*inv5;
          F.ac[0] += qx0 - qf*dx; F.ac[1] += qx1 - qf*dy; F.ac[2] += qx2 - qf*dz;
          F.dphi[0] += qx0 - qf*dx; F.dphi[1] += qx1 - qf*dy; F.dphi[2] += qx2 - qf*dz;
          qx0 *= 5.0*inv2; qx1 *= 5.0*inv2; qx2 *= 5.0*inv2;
          F.q[0] += qf*(7.0*dx*dx*inv2 - 1.0);
          F.q[1] += qf*(7.0*dx*dy*inv2);
          F.q[2] += qf*(7.0*dy*dy*inv2 - 1.0);
          F.q[3] += qf*(7.0*dx*dz*inv2);
          F.q[4] += qf*(7.0*dy*dz*inv2);
          F.q[5] += qf*(7.0*dz*dz*inv2 - 1.0);
          F.q[0] -= qx0*dx + qx0*dx - q0*inv5;
          F.q[1] -= qx0*dy + qx1*dx - q1*inv5;
          F.q[2] -= qx1*dy + qx1*dy - q2*inv5;
          F.q[3] -= qx0*dz + qx2*dx - q3*inv5;
          F.q[4] -= qx1*dz + qx2*dy - q4*inv5;
          F.q[5] -= qx2*dz + qx2*dz + (q0 + q2)*inv5;
        }
      }
      F.pot = warp_sum(F.pot);
#pragma unroll
      for (int k = 0; k < 3; k++) { F.ac[k] = warp_sum(F.ac[k]); F.dphi[k] = warp_sum(F.dphi[k]); }
#pragma unroll
      for (int k = 0; k < 6; k++) F.q[k] = warp_sum(F.q[k]);
      if (lane == 0) {
#pragma unroll
        for (int t = 0; t < GRAV_NT; t++) {
          const double dx = x[t][0] - rcx, dy = x[t][1] - rcy, dz = x[t][2] - rcz;
          ax[t] += F.ac[0] + F.q[0]*dx + F.q[1]*dy + F.q[3]*dz;
          ay[t] += F.ac[1] + F.q[1]*dx + F.q[2]*dy + F.q[4]*dz;
          az[t] += F.ac[2] + F.q[3]*dx + F.q[4]*dy + F.q[5]*dz;
          pot[t] += F.pot + F.dphi[0]*dx + F.dphi[1]*dy + F.dphi[2]*dz;
        }
      }
    }

    /* ---- reduce over lanes and add to the particles (GradhSphTree.cpp:610-620) -------------------*/
    double outx = 0, outy = 0, outz = 0, outp = 0;
#pragma unroll
    for (int t = 0; t < GRAV_NT; t++) {
      const double sx = warp_sum(ax[t]), sy = warp_sum(ay[t]), sz = warp_sum(az[t]), sp = warp_sum(pot[t]);
      if (lane == t) { outx = sx; outy = sy; outz = sz; outp = sp; }
    }
    if (lane < cnt) {
      const int i = first + lane;
      const unsigned fl = (unsigned) P.i[I_FLAGS][i];
      if ((fl & GSPH_FLAG_ACTIVE) && !(fl & GSPH_FLAG_DEAD)) {
        P.f[F_AX][i] += outx; P.f[F_AY][i] += outy; P.f[F_AZ][i] += outz;
        P.f[F_ATX][i] += outx; P.f[F_ATY][i] += outy; P.f[F_ATZ][i] += outz;
        P.f[F_GPOT][i] += outp;
      }
    }
    if (lane == 0) done[g] = 1;
    loc_direct += (unsigned long long) ndirect*cnt; loc_cells += (unsigned long long) ncells*cnt; loc_leaf += cnt;
  }
  if (lane == 0) {
    atomicAdd(&st->ndirect_sum, loc_direct); atomicAdd(&st->ncell_sum, loc_cells);
    atomicAdd(&st->nleaf_grav, loc_leaf);
  }
}

template <int MULTIPOLE>
int launch_gravity(gsph_handle* h)
{
  const size_t smem = GRAV_WARPS*WALK_STACK_CAP*sizeof(int);
  auto kern = gravity_kernel<MULTIPOLE>;
  int bps = 0;
  CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, GRAV_THREADS, smem));
  if (bps < 1) bps = 1;
  int blocks = h->sm_count*bps;
  const int nleaf = 1 << h->tree.ltot;
  const int need_blocks = (nleaf + GRAV_WARPS - 1)/GRAV_WARPS;
  if (blocks > need_blocks) blocks = need_blocks;
  const int warps = blocks*GRAV_WARPS;
  if (h->grav_cap_cells == 0) { h->grav_cap_cells = 8192; h->grav_cap_direct = 16384; }
  if (warps > h->grav_warps || !h->gravlist) {
    if (h->gravlist) cudaFree(h->gravlist);
    h->gravlist = nullptr;
    CUDA_TRY(h, cudaMalloc(&h->gravlist, (size_t) warps*(h->grav_cap_cells + h->grav_cap_direct)*sizeof(int32_t)));
    h->grav_warps = warps;
  }
  CUDA_TRY(h, cudaMemsetAsync(h->work_counter, 0, sizeof(int), h->stream));
  PartView P;
  for (int k = 0; k < NF_DOUBLE; k++) P.f[k] = h->P[k];
  for (int k = 0; k < NF_INT; k++) P.i[k] = h->PI[k];
  KernelTimer kt(h, 2);
  kern<<<blocks, GRAV_THREADS, smem, h->stream>>>(P, h->tree, h->dp, h->work_counter, h->gravlist,
                                                  (int) h->grav_cap_cells, (int) h->grav_cap_direct, h->status, h->done);
  h->stats.kernel_launches++;
  CUDA_TRY(h, cudaGetLastError());
  return GSPH_OK;
}

}  // namespace

int gsph_run_gravity(gsph_handle* h)
{
  if (h->dp.ndim != 3) { h->err = "self-gravity on the accelerated path is 3-D only"; return GSPH_EINVAL; }
  if (h->cfg.Nleafmax > GRAV_NT) { h->err = "self-gravity on the accelerated path needs Nleafmax <= 8"; return GSPH_EINVAL; }
  switch (h->dp.multipole) {
  case GSPH_MULTIPOLE_MONOPOLE: return launch_gravity<GSPH_MULTIPOLE_MONOPOLE>(h);
  case GSPH_MULTIPOLE_QUADRUPOLE: return launch_gravity<GSPH_MULTIPOLE_QUADRUPOLE>(h);
  case GSPH_MULTIPOLE_FAST_MONOPOLE: return launch_gravity<GSPH_MULTIPOLE_FAST_MONOPOLE>(h);
  default: return launch_gravity<GSPH_MULTIPOLE_FAST_QUADRUPOLE>(h);
  }
}
