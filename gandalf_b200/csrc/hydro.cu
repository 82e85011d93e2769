/*=================================================================================================
 *  hydro.cu -- the SPH force pass over kernel-overlap neighbours.
 *
 *  Replaces  GradhSphTree::UpdateAllSphHydroForces   src/GradhSph/GradhSphTree.cpp:280-435
 *            (and the neighbour part of ::UpdateAllSphForces, :444-657)
 *            Tree::ComputeNeighbourAndGhostList       src/Tree/Tree.cpp:562-617
 *            NeighbourManager::_EndSearch / TrimNeighbourLists   src/Headers/NeighbourManager.h:368-543
 *            GradhSph::ComputeSphHydroForces          src/GradhSph/GradhSph.cpp:361-460
 *            GradhSph::ComputeSphHydroGravForces      src/GradhSph/GradhSph.cpp:474-585
 *            GradhSph::ComputeSphGravForces           src/GradhSph/GradhSph.cpp:598-647
 *
 *  One warp per cohort (KD-tree node of <= 32 particles), one target per lane.  The walk opens a
 *  node when cohort.bb overlaps node.hbox or cohort.hbox overlaps node.bb (scatter-gather); the
 *  candidate slabs are streamed through a shared-memory tile of packed (x,y,z,hrangesqd) records and
 *  compacted into per-lane neighbour lists, over which the pair terms are then evaluated from packed
 *  128-byte force records (1/h, 1/rho and P*invomega/rho^2 are derived once per particle by the
 *  pack kernel).  Neighbour membership is the reference's predicate
 *      NOT (drsqd >= hrangesqd_i && drsqd >= hrangesqd_j)          NeighbourManager.h:521
 *  evaluated with the reference's rounding (no FMA contraction).
 *=================================================================================================*/
#include "device_common.cuh"

#define HYD_THREADS 256
#define HYD_WARPS (HYD_THREADS/32)

namespace {


template <int NDIM>
struct ScatterGatherVisitor {
  const TreeDev& T; const DevParams& dp; CandList& cl;
  double bmin[3], bmax[3], hmin[3], hmax[3];
  bool needwrap;
  __device__ __forceinline__ bool overlap(const double (&amin)[3], const double (&amax)[3],
                                          const double (&cmin)[3], const double (&cmax)[3], bool leaf)
  {
    bool wrapped = false;
#pragma unroll
    for (int k = 0; k < NDIM; k++) {
      const double ca = 0.5*(amin[k] + amax[k]), cb = 0.5*(cmin[k] + cmax[k]);
      const double hw = 0.5*(amax[k] - amin[k]) + 0.5*(cmax[k] - cmin[k]);
      double d = fabs(ca - cb);
      if (dp.periodic[k]) {
        if (d + hw > dp.boxhalf[k]) wrapped = true;
        d = fmin(d, dp.boxsize[k] - d);
      }
      if (d > hw*(1.0 + 1.0e-12) + 1.0e-300) return false;
    }
    if (leaf && wrapped) needwrap = true;
    return true;
  }
  __device__ __forceinline__ int test(int node, bool leaf)
  {
    double nb0[3], nb1[3], nh0[3], nh1[3];
#pragma unroll
    for (int k = 0; k < NDIM; k++) {
      nb0[k] = T.bbmin[k][node]; nb1[k] = T.bbmax[k][node];
      nh0[k] = T.hbmin[k][node]; nh1[k] = T.hbmax[k][node];
    }
    if (overlap(bmin, bmax, nh0, nh1, leaf)) return 1;
    if (overlap(hmin, hmax, nb0, nb1, leaf)) return 1;
    return 0;
  }
  __device__ __forceinline__ bool emit(int node, int code, bool leaf, int lane)
  {
    return cl.append_leaf(T, node, code == 1 && leaf, lane);
  }
};

/* scatter-gather walk on the outward-rounded float boxes (superset of
 * Tree::ComputeNeighbourAndGhostList): open when cohort.bb meets node.hbox or cohort.hbox meets node.bb */
template <int NDIM>
struct ScatterGatherVisitorF {
  const float4* __restrict__ nodeb; const float4* __restrict__ nodeh; const WalkBox& wb; CandList& cl;
  float bmin[3], bmax[3], hmin[3], hmax[3];
  bool needwrap;
  int first, count;
  int tag;                          /* image code << IMG_SHIFT of the walk (mirror boundaries) */
  __device__ __forceinline__ void set_boxes(int n)
  {
    const float4 lo = nodeb[2*(size_t) n], hi = nodeb[2*(size_t) n + 1];
    const float4 hlo = nodeh[2*(size_t) n], hhi = nodeh[2*(size_t) n + 1];
    bmin[0] = lo.x; bmin[1] = lo.y; bmin[2] = lo.z; bmax[0] = hi.x; bmax[1] = hi.y; bmax[2] = hi.z;
    hmin[0] = hlo.x; hmin[1] = hlo.y; hmin[2] = hlo.z; hmax[0] = hhi.x; hmax[1] = hhi.y; hmax[2] = hhi.z;
  }
  /* the cohort's boxes seen through a mirror: "image of a node meets the cohort" = "node meets the
     image of the cohort" (outward rounding of the reflected float bounds) */
  __device__ __forceinline__ void reflect_boxes(const ImageOp& op)
  {
#pragma unroll
    for (int k = 0; k < 3; k++) {
      if (!op.refl[k]) continue;
      const float lo = __double2float_rd(op.wall2[k] - (double) bmax[k]), hi = __double2float_ru(op.wall2[k] - (double) bmin[k]);
      const float hlo = __double2float_rd(op.wall2[k] - (double) hmax[k]), hhi = __double2float_ru(op.wall2[k] - (double) hmin[k]);
      bmin[k] = lo; bmax[k] = hi; hmin[k] = hlo; hmax[k] = hhi;
    }
  }
  __device__ __forceinline__ int test(int nd, bool leaf)
  {
    const float4 lo = nodeb[2*(size_t) nd], hi = nodeb[2*(size_t) nd + 1];
    first = __float_as_int(lo.w); count = __float_as_int(hi.w);
    if (count <= 0) return 0;
    bool wrapped = false;
    bool open = fboxes_overlap<NDIM>(wb, hmin, hmax, lo, hi, wrapped);
    if (!open) {
      const float4 hlo = nodeh[2*(size_t) nd], hhi = nodeh[2*(size_t) nd + 1];
      open = fboxes_overlap<NDIM>(wb, bmin, bmax, hlo, hhi, wrapped);
    }
    if (open && leaf && wrapped) needwrap = true;
    return open ? 1 : 0;
  }
  __device__ __forceinline__ bool emit(bool take, int lane) { return cand_append_range(cl, first, count, take, lane, tag); }
};

struct Target {
  double x[3], v[3];
  double invh, hfac, invrho, pfac, sound, u, alpha, press, zeta, hr2, m, rho, invomega;
  int level;
  /* accumulators */
  double a[3], at[3], gpot, dudt, divv;
  int levelneib, nngb;
};

struct Neib {            /* one gathered force record (frec, 128 B) */
  double v[3], m, invh, hfac, invrho, pfac, sound, u, alpha, zeta, press;
  int level;
};

/* frec layout (16 doubles = 128 B, eight 16-byte chunks):
 *   0 x 1 y | 2 z 3 hrangesqd | 4 vx 5 vy | 6 vz 7 m | 8 1/h 9 hfactor | 10 1/rho 11 P*invomega/rho^2 |
 *   12 sound 13 zeta | 14 u 15 aux      aux = alpha for the time-dependent viscosity switches, else level.
 * Chunks 0-1 double as the geometry record of the candidate filter.  Dead particles get x = BIG and
 * hrangesqd = -1 so that no predicate ever accepts them.                                            */
template <int NDIM>
__device__ __forceinline__ void pack_force_one(int p, const PartView& P, int aux_is_alpha, double* __restrict__ frec,
                                               double* __restrict__ geoh);

/* only the particles of the cohort nodes flagged by needed_cohorts_kernel: one warp per cohort node */
template <int NDIM>
__global__ void pack_force_needed_kernel(TreeDev tree, int cohort_level, const uint8_t* __restrict__ needed, PartView P,
                                         int aux_is_alpha, double* __restrict__ frec, double* __restrict__ geoh)
{
  const int c = (blockIdx.x*blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int ncohort = 1 << cohort_level;
  if (c >= ncohort || !needed[c]) return;
  const int node = ncohort - 1 + c;
  const int first = tree.first[node], count = tree.count[node];
  for (int o = lane; o < count; o += 32) pack_force_one<NDIM>(first + o, P, aux_is_alpha, frec, geoh);
}

template <int NDIM>
__global__ void pack_force_kernel(int N, PartView P, int aux_is_alpha, double* __restrict__ frec,
                                  double* __restrict__ geoh)
{
  const int p = blockIdx.x*blockDim.x + threadIdx.x;
  if (p >= N) return;
  pack_force_one<NDIM>(p, P, aux_is_alpha, frec, geoh);
}

template <int NDIM>
__device__ __forceinline__ void pack_force_one(int p, const PartView& P, int aux_is_alpha, double* __restrict__ frec,
                                               double* __restrict__ geoh)
{
  const bool dead = ((unsigned) P.i[I_FLAGS][p]) & GSPH_FLAG_DEAD;
  const double x = P.f[F_RX][p], y = NDIM > 1 ? P.f[F_RY][p] : 0.0, z = NDIM > 2 ? P.f[F_RZ][p] : 0.0;
  const double h = P.f[F_H][p], rho = P.f[F_RHO][p], pr = P.f[F_PRESS][p];
  double2* r = reinterpret_cast<double2*>(frec) + 8*(size_t) p;
  r[0] = make_double2(dead ? GSPH_BIG : x, y);
  r[1] = make_double2(z, dead ? -1.0 : P.f[F_HR2][p]);
  double2* g = reinterpret_cast<double2*>(geoh) + 2*(size_t) p;   /* compact copy for the candidate filter */
  g[0] = r[0]; g[1] = r[1];
  r[2] = make_double2(P.f[F_VX][p], NDIM > 1 ? P.f[F_VY][p] : 0.0);
  r[3] = make_double2(NDIM > 2 ? P.f[F_VZ][p] : 0.0, P.f[F_M][p]);
  r[4] = make_double2(1.0/h, P.f[F_HFAC][p]);
  r[5] = make_double2(1.0/rho, (pr*P.f[F_INVOMEGA][p])/(rho*rho));
  r[6] = make_double2(P.f[F_SOUND][p], P.f[F_ZETA][p]);
  r[7] = make_double2(P.f[F_U][p], aux_is_alpha ? P.f[F_ALPHA][p] : (double) P.i[I_LEVEL][p]);
}

/* Multi-GPU (gsph_set_active_slice): a device only evaluates the targets of its tree-order slice, so it only
 * needs the force records of the cohort nodes that can hold a neighbour of the slice.  The slice is covered by
 * a few tree nodes K (host side, from the data-independent topology of the balanced tree); a cohort node c is
 * NEEDED if it intersects the slice or  c.hbox meets K.bb  or  c.bb meets K.hbox  for some K -- the walk's own
 * scatter-gather test on the same outward-rounded float boxes, applied to super-sets of the boxes it compares,
 * hence a super-set of every candidate the slice's cohorts can find.                                        */
#define SLICE_COVER_MAX 64
struct SliceCover { int n; int node[SLICE_COVER_MAX]; };

template <int NDIM>
__global__ void needed_cohorts_kernel(TreeDev tree, DevParams dp, int cohort_level, SliceCover cov, uint8_t* __restrict__ needed)
{
  const int c = blockIdx.x*blockDim.x + threadIdx.x;
  const int ncohort = 1 << cohort_level;
  if (c >= ncohort) return;
  const int node = ncohort - 1 + c;
  const float4 blo = tree.nodeb[2*(size_t) node], bhi = tree.nodeb[2*(size_t) node + 1];
  const int first = __float_as_int(blo.w), count = __float_as_int(bhi.w);
  bool need = false;
  if (count > 0) {
    need = !GSPH_RANGE_OUTSIDE_SLICE(dp, first, count);
    if (!need) {
      const WalkBox wb(dp);
      const float4 hlo = tree.nodeh[2*(size_t) node], hhi = tree.nodeh[2*(size_t) node + 1];
      const float cb0[3] = {blo.x, blo.y, blo.z}, cb1[3] = {bhi.x, bhi.y, bhi.z};
      const float ch0[3] = {hlo.x, hlo.y, hlo.z}, ch1[3] = {hhi.x, hhi.y, hhi.z};
      for (int k = 0; k < cov.n && !need; k++) {
        const int kn = cov.node[k];
        bool wrapped = false;
        need = fboxes_overlap<NDIM>(wb, ch0, ch1, tree.nodeb[2*(size_t) kn], tree.nodeb[2*(size_t) kn + 1], wrapped) ||
               fboxes_overlap<NDIM>(wb, cb0, cb1, tree.nodeh[2*(size_t) kn], tree.nodeh[2*(size_t) kn + 1], wrapped);
      }
    }
  }
  needed[c] = need ? 1 : 0;
}

static void slice_cover(int node, int64_t a, int64_t b, int64_t lo, int64_t hi, int level, int stop_level, SliceCover& out, bool& overflow)
{
  if (a >= hi || b <= lo || a == b) return;
  if ((lo <= a && b <= hi) || level == stop_level) {
    if (out.n < SLICE_COVER_MAX) out.node[out.n++] = node; else overflow = true;
    return;
  }
  const int64_t left = (b - a)/2;                       /* node sizes N -> (N/2, N - N/2), KDTree.cpp:545-560 */
  slice_cover(2*node + 1, a, a + left, lo, hi, level + 1, stop_level, out, overflow);
  slice_cover(2*node + 2, a + left, b, lo, hi, level + 1, stop_level, out, overflow);
}

#define REC_STRIDE 9      /* double2 units per staged record: 144 B, conflict-free for 16-byte accesses */

/* One (target, neighbour) pair.  GRAV selects ComputeSphHydroGravForces (GradhSph.cpp:498-572),
 * otherwise ComputeSphHydroForces (:384-448); HYDRO = false gives ComputeSphGravForces (:610-645). */
template <int NDIM, bool GRAV, bool HYDRO, class KE>
__device__ __forceinline__ void pair_terms(const DevParams& dp, const KE& K, Target& t, double (&dr)[3],
                                           double drsqd, const Neib& nb)
{
  const double mj = nb.m, invh_j = nb.invh, hfac_j = nb.hfac;
  double drmag, dvdr = 0.0, invdr = 0.0;
  /* drmag feeds the (truncating) table index and keeps the reference's correctly rounded sqrt; the
     unit vector only needs 1/drmag to rounding, so rsqrt replaces the division */
  if (!GRAV) {
    drmag = sqrt(drsqd);
    if (drmag > 0.0) {
      invdr = rsqrt(drsqd);
#pragma unroll
      for (int k = 0; k < NDIM; k++) dr[k] *= invdr;
    }
  }
  else {
    drmag = sqrt(drsqd + GSPH_SMALL);
    invdr = rsqrt(drsqd + GSPH_SMALL);
#pragma unroll
    for (int k = 0; k < NDIM; k++) dr[k] *= invdr;
  }
  const double w1i = K.w1(drmag*t.invh), w1j = K.w1(drmag*invh_j);
  const double wkerni = t.hfac*w1i, wkernj = hfac_j*w1j;
  if (HYDRO) {
#pragma unroll
    for (int k = 0; k < NDIM; k++) dvdr += (nb.v[k] - t.v[k])*dr[k];
    const double invrho_j = nb.invrho;
    t.divv -= mj*dvdr*wkerni;
    double paux = t.pfac*wkerni + nb.pfac*wkernj;
    if (dvdr < 0.0) {
      const double winvrho = 0.25*(wkerni + wkernj)*(t.invrho + invrho_j);
      if (dp.avisc == GSPH_AVISC_MON97) {
        const double vsignal = t.sound + nb.sound - dp.beta_visc*dp.alpha_visc*dvdr;
        paux   -= dp.alpha_visc*vsignal*dvdr*winvrho;
        t.dudt -= 0.5*mj*dp.alpha_visc*vsignal*dvdr*dvdr*winvrho;
      }
      else if (dp.avisc == GSPH_AVISC_MON97MM97 || dp.avisc == GSPH_AVISC_MON97CD2010) {
        const double alpha_mean = 0.5*(t.alpha + nb.alpha);
        const double vsignal = t.sound + nb.sound - dp.beta_visc*alpha_mean*dvdr;
        paux   -= alpha_mean*vsignal*dvdr*winvrho;
        t.dudt -= 0.5*mj*alpha_mean*vsignal*dvdr*dvdr*winvrho;
      }
      if (dp.acond == GSPH_ACOND_WADSLEY2008)
        t.dudt += mj*dvdr*(nb.u - t.u)*(t.invrho*wkerni + invrho_j*wkernj);
      else if (dp.acond == GSPH_ACOND_PRICE2008)
        t.dudt += 0.5*mj*(t.u - nb.u)*winvrho*(t.invrho + invrho_j)*sqrt(fabs(t.press - nb.press));
    }
#pragma unroll
    for (int k = 0; k < NDIM; k++) t.a[k] += mj*dr[k]*paux;
  }
  if (GRAV) {
    double gi, pi, gj, pj;
    K.grav_terms(drmag*t.invh, t.invh, invdr, gi, pi);
    K.grav_terms(drmag*invh_j, invh_j, invdr, gj, pj);
    const double gaux = 0.5*(gi + t.zeta*wkerni + gj + nb.zeta*wkernj);
#pragma unroll
    for (int k = 0; k < NDIM; k++) t.at[k] += mj*dr[k]*gaux;
    t.gpot += 0.5*mj*(pi + pj);
  }
  if (HYDRO || !GRAV) t.levelneib = max(t.levelneib, nb.level);
  t.nngb++;
}

/*------------------------------------------------------------------------------------------------
 * Pre-walk: the scatter-gather walks of the cohorts, in a kernel of their own.  A walk is a chain of dependent
 * node loads -- latency bound -- and needs few registers; inside hydro_kernel (128 registers per thread, 16 warps
 * per SM) it cannot hide that latency.  Here every SM holds 48+ walking warps; the candidate ids go to HBM
 * (pre_cand, a few KB per cohort) and hydro_kernel starts with its FP32 filter.  A cohort whose list outgrows
 * its slot gets count -1 and is walked by hydro_kernel itself, as before.
 *-----------------------------------------------------------------------------------------------*/
template <int NDIM, bool MIRROR>
__global__ void __launch_bounds__(HYD_THREADS)
hydro_walk_kernel(TreeDev tree, DevParams dp, int cohort_level, int c0, int nc, int* work_counter,
                  int32_t* __restrict__ pre_cand, int pre_cap, int* __restrict__ pre_count, const uint8_t* __restrict__ done)
{
  __shared__ int stacks[HYD_WARPS][NB_WALK_STACK];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int* stack = stacks[warp];
  const WalkBox wbx(dp);
  const int ncohort = 1 << cohort_level;
  while (true) {
    int b = 0;
    if (lane == 0) b = atomicAdd(work_counter, 1);
    b = __shfl_sync(FULLMASK, b, 0);
    if (b >= nc) break;
    const int c = c0 + b;
    const int node = ncohort - 1 + c;
    const int cfirst = tree.first[node], ccount = tree.count[node];
    if (ccount == 0 || tree.nactive[node] == 0 || done[c] || GSPH_RANGE_OUTSIDE_SLICE(dp, cfirst, ccount)) {
      if (lane == 0) pre_count[b] = -1;
      continue;
    }
    CandList cl{pre_cand + (size_t) b*pre_cap, pre_cap, 0, nullptr, true};
    ScatterGatherVisitorF<NDIM> V{tree.nodeb, tree.nodeh, wbx, cl, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f},
                                  {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, false, 0, 0, 0};
    V.set_boxes(node);
    __syncwarp();
    bool walk_ok = warp_walk_f<NB_WALK_STACK>(tree.ltot, stack, V, (DevStatus*) nullptr);
    for (int code = 1; code < (MIRROR ? dp.nimg : 1) && walk_ok; code++) {
      V.set_boxes(node);
      V.reflect_boxes(image_op<NDIM>(dp, code));
      V.tag = code << IMG_SHIFT;
      __syncwarp();
      walk_ok = warp_walk_f<NB_WALK_STACK>(tree.ltot, stack, V, (DevStatus*) nullptr);
    }
    const bool wrapped = __any_sync(FULLMASK, V.needwrap);
    __syncwarp();
    if (lane == 0) pre_count[b] = walk_ok ? (cl.n | (wrapped ? (1 << 30) : 0)) : -1;
  }
}

/*------------------------------------------------------------------------------------------------
 * One warp per cohort (KD-tree node of <= 32 particles), one target per lane.
 *  phase 1  candidate filter on the FP32 pipe: the candidate slabs are streamed through a shared
 *           tile of float4 records (position relative to the cohort centre, scaled by the largest
 *           kernel range in the cohort, and hrangesqd_j) and every lane appends the ids inside
 *           max(hrangesqd_i, hrangesqd_j)*(1 + 1e-3) to its own list ([k][lane] rows in global
 *           scratch).  The list is a SUPERSET of the neighbours; the margin is far above the FP32
 *           error (lanes whose range is tiny against the cohort size simply keep every candidate).
 *  phase 2  per list slot the warp gathers the 32 force records (128 B each) cooperatively -- eight
 *           lanes per record, 16-byte cp.async copies into a padded shared-memory stage, double
 *           buffered so that the gather of slot k+1 overlaps the arithmetic of slot k -- then every
 *           lane reads its own record, applies the reference's predicate (NeighbourManager.h:521)
 *           in exact FP64 arithmetic and evaluates the pair terms.
 *-----------------------------------------------------------------------------------------------*/
template <int NDIM, int KERN, bool TAB, bool GRAV, bool HYDRO, bool MIRROR, bool SCATTER>
__global__ void __launch_bounds__(HYD_THREADS, 2)
hydro_kernel(PartView P, TreeDev tree, DevParams dp, const double* __restrict__ gtables,
             const double2* __restrict__ frec, const double2* __restrict__ geoh, int aux_is_alpha,
             int cohort_level, int* work_counter, int32_t* cand_all, int cand_cap,
             int32_t* nlist_all, int nlist_cap, DevStatus* st, uint8_t* done,
             const int32_t* __restrict__ pre_cand, int pre_cap, const int* __restrict__ pre_count, int pre_c0, int pre_n)
{
  extern __shared__ double smem[];
  constexpr int NTAB = TAB ? (GRAV ? 3 : 1) : 0;
  constexpr int WARP_D2 = 2*32*REC_STRIDE;              /* double2 per warp: two record stages */
  double* tab = smem;                                   /* w1 [, wgrav, wpot] */
  double2* stages = reinterpret_cast<double2*>(smem + NTAB*GSPH_TABLE_RES);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (TAB) {
    for (int i = threadIdx.x; i < NTAB*GSPH_TABLE_RES; i += blockDim.x) tab[i] = gtables[3*GSPH_TABLE_RES + i];
    __syncthreads();
  }
  /* the per-warp region is time-shared: walk stack | float4 tile + ids (phase 1) | record stages (phase 2) */
  double2* stage = stages + warp*WARP_D2;
  int* stack = reinterpret_cast<int*>(stage);
  float* ftile = reinterpret_cast<float*>(stage);       /* x[32] | y[32] | z[32] | w[32] */
  int* tj = reinterpret_cast<int*>(ftile + 128);
  const WalkBox wbx(dp);
  const int gwarp = blockIdx.x*HYD_WARPS + warp;
  int32_t* const mycand = cand_all + (size_t) gwarp*cand_cap;
  int32_t* nlist = nlist_all + (size_t) gwarp*nlist_cap*32;
  /* evaluator indexes tables as 3=w1, 4=wgrav, 5=wpot: rebase so that slot 3 is tab[0] */
  const KernelEval<KERN, NDIM, TAB> K(dp, tab - 3*GSPH_TABLE_RES);
  const int ncohort = 1 << cohort_level;
  const bool price = dp.acond == GSPH_ACOND_PRICE2008;
  const double rootR = dp.kernrange*tree.hmax[0];
  unsigned long long loc_ngb = 0, loc_cand = 0, loc_targets = 0; int loc_ngbmax = 0;

  while (true) {
    int c = 0;
    if (lane == 0) c = atomicAdd(work_counter, 1);
    c = __shfl_sync(FULLMASK, c, 0);
    if (c >= ncohort) break;
    const int node = ncohort - 1 + c;
    const int cfirst = tree.first[node], ccount = tree.count[node];
    if (ccount == 0 || tree.nactive[node] == 0 || done[c] || GSPH_RANGE_OUTSIDE_SLICE(dp, cfirst, ccount)) continue;

    /* ---- candidates: from the pre-walk kernel, or (no pre-walk / its slot overflowed) walked here ---------*/
    double ctr[3] = {0, 0, 0}, diag2 = 0.0;
    bool wide = false;                /* periodic images not unique about the cohort centre: keep all */
#pragma unroll
    for (int k = 0; k < NDIM; k++) {
      const double b0 = tree.bbmin[k][node], b1 = tree.bbmax[k][node];
      ctr[k] = 0.5*(b0 + b1);
      const double hw = 0.5*(b1 - b0);
      diag2 += hw*hw;
      if (dp.periodic[k] && hw + rootR >= 0.49*dp.boxsize[k]) wide = true;
    }
    const int32_t* cand = mycand;
    int ncand = 0; bool wrap = false;
    const int pre = (c >= pre_c0 && c < pre_c0 + pre_n) ? pre_count[c - pre_c0] : -1;
    if (pre >= 0) {
      ncand = pre & 0x3fffffff; wrap = dp.any_periodic && (pre >> 30);
      cand = pre_cand + (size_t)(c - pre_c0)*pre_cap;
    }
    else {
      CandList cl{mycand, cand_cap, 0, st};
      ScatterGatherVisitorF<NDIM> V{tree.nodeb, tree.nodeh, wbx, cl, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f},
                                    {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, false, 0, 0, 0};
      V.set_boxes(node);
      __syncwarp();
      bool walk_ok = warp_walk_f<NB_WALK_STACK>(tree.ltot, stack, V, st);
      /* mirror walls: one more walk per image code with the cohort's boxes reflected */
      for (int code = 1; code < (MIRROR ? dp.nimg : 1) && walk_ok; code++) {
        V.set_boxes(node);
        V.reflect_boxes(image_op<NDIM>(dp, code));
        V.tag = code << IMG_SHIFT;
        __syncwarp();
        walk_ok = warp_walk_f<NB_WALK_STACK>(tree.ltot, stack, V, st);
      }
      if (!walk_ok) continue;
      ncand = cl.n;
      wrap = dp.any_periodic && __any_sync(FULLMASK, V.needwrap);
    }
    bool overflow = false;

    for (int t0 = 0; t0 < ccount; t0 += 32) {
      Target g;
      const int ti = cfirst + t0 + lane;
      bool live = t0 + lane < ccount && GSPH_IN_SLICE(dp, ti);
      if (live) {
        const unsigned fl = (unsigned) P.i[I_FLAGS][ti];
        live = (fl & GSPH_FLAG_ACTIVE) && !(fl & GSPH_FLAG_DEAD);
      }
      g.x[0] = g.x[1] = g.x[2] = GSPH_BIG; g.v[0] = g.v[1] = g.v[2] = 0.0;
      g.invh = g.hfac = g.invrho = g.pfac = g.sound = g.u = g.alpha = g.press = g.zeta = 0.0;
      g.hr2 = -1.0; g.m = 0.0; g.rho = 1.0; g.invomega = 1.0; g.level = 0;
      if (live) {
#pragma unroll
        for (int k = 0; k < NDIM; k++) { g.x[k] = P.f[F_RX + k][ti]; g.v[k] = P.f[F_VX + k][ti]; }
#pragma unroll
        for (int k = NDIM; k < 3; k++) g.x[k] = 0.0;
        const double hi = P.f[F_H][ti];
        g.rho = P.f[F_RHO][ti]; g.press = P.f[F_PRESS][ti]; g.invomega = P.f[F_INVOMEGA][ti];
        g.invh = 1.0/hi; g.invrho = 1.0/g.rho; g.hfac = P.f[F_HFAC][ti];
        g.pfac = (g.press*g.invomega)/(g.rho*g.rho);
        g.sound = P.f[F_SOUND][ti]; g.u = P.f[F_U][ti]; g.alpha = P.f[F_ALPHA][ti];
        g.zeta = P.f[F_ZETA][ti]; g.hr2 = P.f[F_HR2][ti]; g.m = P.f[F_M][ti];
        g.level = P.i[I_LEVEL][ti];
      }
      g.a[0] = g.a[1] = g.a[2] = 0.0; g.at[0] = g.at[1] = g.at[2] = 0.0;
      g.gpot = 0.0; g.dudt = 0.0; g.divv = 0.0; g.levelneib = 0; g.nngb = 0;
      if (GRAV && live) g.gpot = (g.m*g.invh)*K.wpot(0.0);          /* GradhSphTree.cpp:512 */

      /* ---- phase 1: FP32 candidate filter (superset of the neighbour list) --------------------------*/
      double hr2max = live ? g.hr2 : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) hr2max = fmax(hr2max, __shfl_xor_sync(FULLMASK, hr2max, o));
      if (!(hr2max > 0.0)) continue;                                  /* no live target in this block */
      const double sc = rsqrt(hr2max), sc2 = sc*sc*1.001;
      float xf[3] = {0.f, 0.f, 0.f};
      float thr = -1.0f;
      if (live) {
#pragma unroll
        for (int k = 0; k < NDIM; k++) xf[k] = (float)((g.x[k] - ctr[k])*sc);
        thr = __double2float_ru(g.hr2*sc2);
        if (wide || diag2 > 1.0e6*g.hr2) thr = __int_as_float(0x7f800000);   /* keep every candidate */
      }
      int n = 0; bool full = false;
      /* the candidate ids are fetched one tile ahead: the record load of a tile never waits for its address */
      int jpre = lane < ncand ? cand[lane] : -1;
      for (int base = 0; base < ncand; base += 32) {
        const int jl = base + lane;
        const int jcur = jpre;
        jpre = (jl + 32 < ncand) ? cand[jl + 32] : -1;
        __syncwarp();
        {
          int j = -1;
          float fx = __int_as_float(0x7f800000), fy = 0.f, fz = 0.f, fw = -1.0f;
          if (jl < ncand) {
            j = jcur;
            const int jp = MIRROR ? (j & IMG_MASK) : j;
            const double2 a = geoh[2*(size_t) jp], b = geoh[2*(size_t) jp + 1];
            double xs[3] = {a.x, a.y, b.x};
            if (MIRROR && (j >> IMG_SHIFT)) {
              const ImageOp op = image_op<NDIM>(dp, j >> IMG_SHIFT);
#pragma unroll
              for (int k = 0; k < NDIM; k++) xs[k] = reflect_coord(op, k, xs[k]);
            }
            double dr[3] = {0, 0, 0};
            dr[0] = xs[0] - ctr[0];
            if (NDIM > 1) dr[1] = xs[1] - ctr[1];
            if (NDIM > 2) dr[2] = xs[2] - ctr[2];
            if (dp.any_periodic) wrap_dr<NDIM>(dp, dr);
            if (b.y > 0.0) { fx = (float)(dr[0]*sc); fw = __double2float_ru(b.y*sc2); }
            fy = (float)(dr[1]*sc); fz = (float)(dr[2]*sc);
          }
          ftile[lane] = fx; ftile[32 + lane] = fy; ftile[64 + lane] = fz; ftile[96 + lane] = fw;
          tj[lane] = j;
        }
        __syncwarp();
        unsigned pass = live ? filter_tile<true>(ftile, xf, thr) : 0u;
        if (n + __popc(pass) > nlist_cap) { full = true; pass = 0; }
        while (pass) {
          const int jj = __ffs(pass) - 1;
          pass &= pass - 1u;
          nlist[(size_t) n*32 + lane] = tj[jj];
          n++;
        }
      }
      if (__any_sync(FULLMASK, full)) {
        if (lane == 0) atomicExch(&st->err_nlist_overflow, 1);
        overflow = true; break;
      }
      int nmax = n;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) nmax = max(nmax, __shfl_xor_sync(FULLMASK, nmax, o));
      __syncwarp();                                                   /* tile region becomes the stages */

      /* ---- phase 2: cooperative double-buffered gather + pair terms -----------------------------------*/
      auto gather = [&](int jv, double2* buf) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
          const int src = r*4 + (lane >> 3);
          const int jj = __shfl_sync(FULLMASK, jv, src);
          if (jj >= 0) cp_async16_cg(buf + src*REC_STRIDE + (lane & 7), frec + 8*(size_t)(MIRROR ? (jj & IMG_MASK) : jj) + (lane & 7));
        }
        cp_async_commit();
      };
      int jc = (0 < n) ? nlist[lane] : -1;
      int jn = (1 < n) ? nlist[32 + lane] : -1;
      if (nmax > 0) gather(jc, stage);
      for (int k = 0; k < nmax; k++) {
        double2* buf = stage + (k & 1)*32*REC_STRIDE;
        const bool more = k + 1 < nmax;
        if (more) gather(jn, stage + ((k + 1) & 1)*32*REC_STRIDE);
        const int jnn = (k + 2 < n) ? nlist[(size_t)(k + 2)*32 + lane] : -1;
        if (more) cp_async_wait<1>(); else cp_async_wait<0>();
        __syncwarp();
        if (jc >= 0) {
          const double2* r = buf + lane*REC_STRIDE;
          const double2 r0 = r[0], r1 = r[1];
          double xs[3] = {r0.x, r0.y, r1.x};
          ImageOp op;
          const bool img = MIRROR && (jc >> IMG_SHIFT);
          if (img) {                                          /* a mirror image: r_k -> 2*wall - r_k */
            op = image_op<NDIM>(dp, jc >> IMG_SHIFT);
#pragma unroll
            for (int k = 0; k < NDIM; k++) xs[k] = reflect_coord(op, k, xs[k]);
          }
          double dr[3] = {0, 0, 0};
          dr[0] = xs[0] - g.x[0];
          if (NDIM > 1) dr[1] = xs[1] - g.x[1];
          if (NDIM > 2) dr[2] = xs[2] - g.x[2];
          if (wrap) wrap_dr<NDIM>(dp, dr);
          const double drsqd = dot_exact<NDIM>(dr);
          if (drsqd < g.hr2 || drsqd < r1.y) {                  /* NeighbourManager.h:521 */
            const double2 r2 = r[2], r3 = r[3], r4 = r[4], r5 = r[5], r6 = r[6], r7 = r[7];
            Neib nb;
            nb.v[0] = r2.x; nb.v[1] = r2.y; nb.v[2] = r3.x;
            if (img) {                                        /* ... and v_k -> -v_k (Particle.h:603-607) */
#pragma unroll
              for (int k = 0; k < NDIM; k++) if (op.refl[k]) nb.v[k] = -nb.v[k];
            }
            nb.m = r3.y; nb.invh = r4.x; nb.hfac = r4.y; nb.invrho = r5.x; nb.pfac = r5.y;
            nb.sound = r6.x; nb.zeta = r6.y; nb.u = r7.x;
            nb.alpha = 0.0; nb.press = 0.0;
            const int jpart = MIRROR ? (jc & IMG_MASK) : jc;
            if (aux_is_alpha) { nb.alpha = r7.y; nb.level = P.i[I_LEVEL][jpart]; }
            else nb.level = (int) r7.y;
            if (price) nb.press = P.f[F_PRESS][jpart];
            pair_terms<NDIM, GRAV, HYDRO>(dp, K, g, dr, drsqd, nb);
            /* levelneib scatter half (GradhSph.cpp:446, GradhSphTree.cpp:376-382; a mirror image hands it to its
               original): only instantiated for steps with inactive particles -- with every particle active the
               relation is symmetric and the gather above already produces the maximum */
            if (SCATTER && (HYDRO || !GRAV) && P.i[I_LEVELNEIB][jpart] < g.level) atomicMax(&P.i[I_LEVELNEIB][jpart], g.level);
          }
        }
        __syncwarp();
        jc = jn; jn = jnn;
      }

      /* ---- end-of-loop terms and write-back (GradhSph.cpp:451-457 / :576-582;
       *      GradhSphTree.cpp:396-406 / :610-620) ---------------------------------------------------*/
      if (live) {
        const int i = ti;
        if (HYDRO) {
          g.divv *= g.invrho;
          const double press = dp.gammam1*P.f[F_RHO][i]*g.u;            /* eos->Pressure(parti) */
          g.dudt -= press*g.divv*g.invrho*P.f[F_INVOMEGA][i];
#pragma unroll
          for (int k = 0; k < NDIM; k++) P.f[F_AX + k][i] += g.a[k];
          P.f[F_DUDT][i] += g.dudt;
          P.f[F_DIVV][i] += g.divv;
          if (!GRAV && dp.tdavisc == GSPH_TDAVISC_MM97)
            P.f[F_DALPHADT][i] += 0.1*g.sound*(dp.alpha_visc_min - g.alpha)*g.invh +
              fmax(-g.divv, 0.0)*(dp.alpha_visc - g.alpha);
        }
        if (GRAV) {
#pragma unroll
          for (int k = 0; k < NDIM; k++) { P.f[F_AX + k][i] += g.at[k]; P.f[F_ATX + k][i] += g.at[k]; }
          P.f[F_GPOT][i] += g.gpot;
        }
        if (g.levelneib > P.i[I_LEVELNEIB][i]) atomicMax(&P.i[I_LEVELNEIB][i], g.levelneib);
        loc_ngb += g.nngb; loc_targets++; loc_ngbmax = max(loc_ngbmax, g.nngb);
        loc_cand += (unsigned long long) ncand;
      }
    }
    __syncwarp();
    if (!overflow && lane == 0) done[c] = 1;
  }
  loc_ngb = (unsigned long long) warp_sum((double) loc_ngb);
  loc_targets = (unsigned long long) warp_sum((double) loc_targets);
  loc_cand = (unsigned long long) warp_sum((double) loc_cand);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) loc_ngbmax = max(loc_ngbmax, __shfl_xor_sync(FULLMASK, loc_ngbmax, o));
  if (lane == 0) {
    atomicAdd(&st->ngb_sum, loc_ngb); atomicAdd(&st->ntarget, loc_targets);
    atomicAdd(&st->cand_sum, loc_cand); atomicMax(&st->ngb_max, loc_ngbmax);
  }
}

/* The parity API: neighbour counts / lists of every particle with the reference's predicate.  (The levelneib scatter
 * half of steps with inactive particles used to be a third mode of this kernel; it now lives in
 * hydro_kernel<..., SCATTER>.)                                                                             */
template <int NDIM, int MODE /*0 count, 1 fill*/>
__global__ void __launch_bounds__(HYD_THREADS)
neib_kernel(PartView P, TreeDev tree, DevParams dp, int cohort_level, int* work_counter,
            int32_t* cand_all, int cand_cap, DevStatus* st, const int32_t* perm, int32_t* counts,
            const int64_t* offsets, int32_t* idx)
{
  extern __shared__ double smem[];
  double* tiles = smem;
  int* stacks = (int*)(tiles + HYD_WARPS*32*4);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* geom = tiles + warp*32*4;
  int* stack = stacks + warp*NB_WALK_STACK;
  __shared__ int jidx[HYD_WARPS][32];
  const int gwarp = blockIdx.x*HYD_WARPS + warp;
  int32_t* cand = cand_all + (size_t) gwarp*cand_cap;
  const int ncohort = 1 << cohort_level;
  while (true) {
    int c = 0;
    if (lane == 0) c = atomicAdd(work_counter, 1);
    c = __shfl_sync(FULLMASK, c, 0);
    if (c >= ncohort) break;
    const int node = ncohort - 1 + c;
    const int cfirst = tree.first[node], ccount = tree.count[node];
    if (ccount == 0) continue;
    CandList cl{cand, cand_cap, 0, st};
    ScatterGatherVisitor<NDIM> V{tree, dp, cl, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, false};
#pragma unroll
    for (int k = 0; k < NDIM; k++) {
      V.bmin[k] = tree.bbmin[k][node]; V.bmax[k] = tree.bbmax[k][node];
      V.hmin[k] = tree.hbmin[k][node]; V.hmax[k] = tree.hbmax[k][node];
    }
    if (!warp_walk<NB_WALK_STACK>(tree, stack, V, st)) continue;
    const int ncand = cl.n;
    const bool wrap = dp.any_periodic && __any_sync(FULLMASK, V.needwrap);
    for (int t0 = 0; t0 < ccount; t0 += 32) {
      const int o = t0 + lane;
      const int i = cfirst + o;
      bool live = o < ccount;
      double x[3] = {GSPH_BIG, 0, 0}, hr2 = -1.0; int lev = 0;
      if (live) {
        const unsigned fl = (unsigned) P.i[I_FLAGS][i];
        if (fl & GSPH_FLAG_DEAD) live = false;
      }
      if (live) {
#pragma unroll
        for (int k = 0; k < NDIM; k++) x[k] = P.f[F_RX + k][i];
        hr2 = P.f[F_HR2][i]; lev = P.i[I_LEVEL][i];
      }
      int n = 0;
      int64_t off = 0;
      if (MODE == 1 && live) off = offsets[perm[i]];
      for (int base = 0; base < ncand; base += 32) {
        const int jl = base + lane;
        __syncwarp();
        if (jl < ncand) {
          const int j = cand[jl];
          const bool dead = ((unsigned) P.i[I_FLAGS][j]) & GSPH_FLAG_DEAD;
          geom[lane*4 + 0] = dead ? GSPH_BIG : P.f[F_RX][j];
          geom[lane*4 + 1] = NDIM > 1 ? P.f[F_RY][j] : 0.0;
          geom[lane*4 + 2] = NDIM > 2 ? P.f[F_RZ][j] : 0.0;
          geom[lane*4 + 3] = dead ? -1.0 : P.f[F_HR2][j];
          jidx[warp][lane] = j;
        }
        __syncwarp();
        const int nt = min(32, ncand - base);
        for (int jj = 0; jj < nt; jj++) {
          double dr[3] = {0, 0, 0};
          dr[0] = geom[jj*4] - x[0];
          if (NDIM > 1) dr[1] = geom[jj*4 + 1] - x[1];
          if (NDIM > 2) dr[2] = geom[jj*4 + 2] - x[2];
          if (wrap) wrap_dr<NDIM>(dp, dr);
          const double drsqd = dot_exact<NDIM>(dr);
          if (live && (drsqd < hr2 || drsqd < geom[jj*4 + 3])) {
            if (MODE == 1) idx[off + n] = perm[jidx[warp][jj]];
            n++;
          }
        }
      }
      if (MODE == 0 && live) counts[perm[i]] = n;
    }
  }
}

template <int NDIM, int KERN, bool TAB, bool GRAV, bool HYDRO, bool MIRROR, bool SCATTER = false>
int launch_hydro_m(gsph_handle* h)
{
  if constexpr (!SCATTER) {
    if (!h->all_active) return launch_hydro_m<NDIM, KERN, TAB, GRAV, HYDRO, MIRROR, true>(h);
  }
  const int Lc = cohort_level_for(h->N, h->tree.ltot, 32);
  constexpr int NTAB = TAB ? (GRAV ? 3 : 1) : 0;
  static_assert(2*32*REC_STRIDE*sizeof(double2) >= NB_WALK_STACK*sizeof(int), "walk stack must fit the stage region");
  const size_t smem = NTAB*GSPH_TABLE_RES*sizeof(double) + HYD_WARPS*2*32*REC_STRIDE*sizeof(double2);
  auto kern = hydro_kernel<NDIM, KERN, TAB, GRAV, HYDRO, MIRROR, SCATTER>;
  CUDA_TRY(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  int bps = 0;
  CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, HYD_THREADS, smem));
  if (bps < 1) bps = 1;
  int blocks = h->sm_count*bps;
  const int ncohort = 1 << Lc;
  const int need_blocks = (ncohort + HYD_WARPS - 1)/HYD_WARPS;
  if (blocks > need_blocks) blocks = need_blocks;
  int rc = gsph_ensure_cand(h, blocks*HYD_WARPS, 0);
  if (rc) return rc;
  rc = gsph_ensure_nlist(h, blocks*HYD_WARPS, 0);
  if (rc) return rc;
  CUDA_TRY(h, cudaMemsetAsync(h->work_counter, 0, sizeof(int), h->stream));
  PartView P;
  for (int k = 0; k < NF_DOUBLE; k++) P.f[k] = h->P[k];
  for (int k = 0; k < NF_INT; k++) P.i[k] = h->PI[k];
  const int N = (int) h->N;
  const int aux_is_alpha = (h->dp.avisc == GSPH_AVISC_MON97MM97 || h->dp.avisc == GSPH_AVISC_MON97CD2010) ? 1 : 0;
  bool packed = false;
  if (h->slice_mode && !h->dp.any_mirror) {
    /* multi-GPU: records only where a neighbour of this device's slice can live */
    SliceCover cov; cov.n = 0; bool overflow = false;
    slice_cover(0, 0, h->N, h->dp.slice_lo, h->dp.slice_hi, 0, Lc, cov, overflow);
    if (!overflow && cov.n > 0) {
      if ((size_t) ncohort > h->needed_cap) {
        if (h->needed) cudaFree(h->needed);
        h->needed = nullptr; h->needed_cap = 0;
        CUDA_TRY(h, cudaMalloc(&h->needed, (size_t) ncohort));
        h->needed_cap = (size_t) ncohort;
      }
      needed_cohorts_kernel<NDIM><<<(ncohort + 255)/256, 256, 0, h->stream>>>(h->tree, h->dp, Lc, cov, h->needed);
      pack_force_needed_kernel<NDIM><<<(int)(((size_t) ncohort*32 + 255)/256), 256, 0, h->stream>>>(h->tree, Lc, h->needed, P, aux_is_alpha,
                                                                                               h->frec, h->geoh);
      h->stats.kernel_launches++;
      packed = true;
    }
  }
  if (!packed) pack_force_kernel<NDIM><<<(N + 255)/256, 256, 0, h->stream>>>(N, P, aux_is_alpha, h->frec, h->geoh);
  KernelTimer kt(h, 1);
  /* the walks first, in their own lean kernel (hydro_walk_kernel) */
  rc = gsph_prewalk_prepare(h, Lc);
  if (rc) return rc;
  if (h->pre_n > 0) {
    CUDA_TRY(h, cudaMemsetAsync(h->work_counter + 1, 0, sizeof(int), h->stream));
    auto wk = hydro_walk_kernel<NDIM, MIRROR>;
    int wbps = 0;
    CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&wbps, wk, HYD_THREADS, 0));
    const int wblocks = std::min(h->sm_count*std::max(wbps, 1), (h->pre_n + HYD_WARPS - 1)/HYD_WARPS);
    wk<<<wblocks, HYD_THREADS, 0, h->stream>>>(h->tree, h->dp, Lc, h->pre_c0, h->pre_n, h->work_counter + 1, h->pre_cand,
                                               h->pre_cap, h->pre_count, h->done);
    h->stats.kernel_launches++;
  }
  kern<<<blocks, HYD_THREADS, smem, h->stream>>>(P, h->tree, h->dp, h->tables, (const double2*) h->frec,
                                                 (const double2*) h->geoh, aux_is_alpha, Lc, h->work_counter, h->cand,
                                                 (int) h->cand_cap_per_warp, h->nlist, (int) h->nlist_cap,
                                                 h->status, h->done, h->pre_cand, h->pre_cap, h->pre_count, h->pre_c0, h->pre_n);
  h->stats.kernel_launches += 2;
  CUDA_TRY(h, cudaGetLastError());
  return GSPH_OK;
}

template <int NDIM, int KERN, bool TAB, bool GRAV, bool HYDRO>
int launch_hydro(gsph_handle* h)
{
  /* mirror walls (never with self-gravity, Tree.cpp:636) are a separate instantiation */
  if (!GRAV && h->dp.any_mirror) return launch_hydro_m<NDIM, KERN, TAB, false, true, true>(h);
  return launch_hydro_m<NDIM, KERN, TAB, GRAV, HYDRO, false>(h);
}

template <int NDIM, int MODE>
int launch_neib(gsph_handle* h, int32_t* counts, const int64_t* offsets, int32_t* idx)
{
  const int Lc = cohort_level_for(h->N, h->tree.ltot, 32);
  const size_t smem = HYD_WARPS*32*4*sizeof(double) + HYD_WARPS*NB_WALK_STACK*sizeof(int);
  auto kern = neib_kernel<NDIM, MODE>;
  int bps = 0;
  CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, HYD_THREADS, smem));
  if (bps < 1) bps = 1;
  int blocks = h->sm_count*bps;
  const int ncohort = 1 << Lc;
  const int need_blocks = (ncohort + HYD_WARPS - 1)/HYD_WARPS;
  if (blocks > need_blocks) blocks = need_blocks;
  int rc = gsph_ensure_cand(h, blocks*HYD_WARPS, 0);
  if (rc) return rc;
  CUDA_TRY(h, cudaMemsetAsync(h->work_counter, 0, sizeof(int), h->stream));
  PartView P;
  for (int k = 0; k < NF_DOUBLE; k++) P.f[k] = h->P[k];
  for (int k = 0; k < NF_INT; k++) P.i[k] = h->PI[k];
  kern<<<blocks, HYD_THREADS, smem, h->stream>>>(P, h->tree, h->dp, Lc, h->work_counter, h->cand,
                                                 (int) h->cand_cap_per_warp, h->status, h->perm,
                                                 counts, offsets, idx);
  h->stats.kernel_launches++;
  CUDA_TRY(h, cudaGetLastError());
  return GSPH_OK;
}

template <int NDIM, int KERN, bool TAB>
int dispatch_mode(gsph_handle* h, int with_gravity)
{
  if (!with_gravity) return launch_hydro<NDIM, KERN, TAB, false, true>(h);
  if (h->dp.hydro_forces) return launch_hydro<NDIM, KERN, TAB, true, true>(h);
  return launch_hydro<NDIM, KERN, TAB, true, false>(h);
}

template <int NDIM>
int dispatch_hydro(gsph_handle* h, int with_gravity)
{
  const bool tab = h->dp.tabulated != 0;
  int rc;
  if (h->dp.kernel == GSPH_KERNEL_M4)
    rc = tab ? dispatch_mode<NDIM, GSPH_KERNEL_M4, true>(h, with_gravity)
             : dispatch_mode<NDIM, GSPH_KERNEL_M4, false>(h, with_gravity);
  else if (h->dp.kernel == GSPH_KERNEL_GAUSSIAN) {
    /* no self-gravity with this kernel (validated at create): hydro-only instantiations */
    if (with_gravity) { h->err = "gaussian kernel with gravity"; return GSPH_EINVAL; }
    rc = tab ? launch_hydro<NDIM, GSPH_KERNEL_GAUSSIAN, true, false, true>(h)
             : launch_hydro<NDIM, GSPH_KERNEL_GAUSSIAN, false, false, true>(h);
  }
  else
    rc = tab ? dispatch_mode<NDIM, GSPH_KERNEL_QUINTIC, true>(h, with_gravity)
             : dispatch_mode<NDIM, GSPH_KERNEL_QUINTIC, false>(h, with_gravity);
  if (rc) return rc;
  /* (the levelneib scatter half of steps with inactive particles happens inside hydro_kernel<..., SCATTER>; with an
     active slice the scatter of this device's targets lands on particles of every slice: the host layer combines
     the devices' levelneib arrays with a MAX reduction instead of gathering slices) */
  return rc;
}

/* GradhSph::ComputeStarGravForces (GradhSph.cpp:699-743) for every active particle: softened gravity
 * of the stars with the mean-h kernel into atree, a (both force drivers add atree to a,
 * GradhSphTree.cpp:396-399 / :611-614) and gpot.  Nstar is small: one thread per particle.        */
template <int NDIM, int KERN, bool TAB>
__global__ void star_forces_kernel(int N, PartView P, DevParams dp, const double* __restrict__ gtables,
                                   const double* __restrict__ stars, int Nstar)
{
  const int i = blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= N || !GSPH_IN_SLICE(dp, i)) return;
  const unsigned fl = (unsigned) P.i[I_FLAGS][i];
  if (!(fl & GSPH_FLAG_ACTIVE) || (fl & GSPH_FLAG_DEAD)) return;
  const KernelEval<KERN, NDIM, TAB> K(dp, gtables);
  double x[3] = {0, 0, 0}, at[3] = {0, 0, 0}, pot = 0.0;
#pragma unroll
  for (int k = 0; k < NDIM; k++) x[k] = P.f[F_RX + k][i];
  const double hi = P.f[F_H][i];
  for (int j = 0; j < Nstar; j++) {
    const double* s = stars + (size_t) j*(NDIM + 2);
    double dr[3] = {0, 0, 0};
#pragma unroll
    for (int k = 0; k < NDIM; k++) dr[k] = s[k] - x[k];
    const double drsqd = __dadd_rn(dot_exact<NDIM>(dr), GSPH_SMALL);
    const double drmag = sqrt(drsqd), invdrmag = 1.0/drmag;
    const double invhmean = 2.0/(hi + s[NDIM + 1]);
    const double ms = s[NDIM];
    const double paux = ms*invhmean*invhmean*K.wgrav(drmag*invhmean)*invdrmag;
#pragma unroll
    for (int k = 0; k < NDIM; k++) at[k] += paux*dr[k];
    pot += ms*invhmean*K.wpot(drmag*invhmean);
  }
#pragma unroll
  for (int k = 0; k < NDIM; k++) { P.f[F_ATX + k][i] += at[k]; P.f[F_AX + k][i] += at[k]; }
  P.f[F_GPOT][i] += pot;
}

template <int NDIM, int KERN, bool TAB>
int launch_stars(gsph_handle* h)
{
  PartView P;
  for (int k = 0; k < NF_DOUBLE; k++) P.f[k] = h->P[k];
  for (int k = 0; k < NF_INT; k++) P.i[k] = h->PI[k];
  const int N = (int) h->N;
  star_forces_kernel<NDIM, KERN, TAB><<<(N + 127)/128, 128, 0, h->stream>>>(N, P, h->dp, h->tables, h->stars, h->Nstar);
  h->stats.kernel_launches++;
  CUDA_TRY(h, cudaGetLastError());
  return GSPH_OK;
}

template <int NDIM>
int dispatch_stars(gsph_handle* h)
{
  const bool tab = h->dp.tabulated != 0;
  if (h->dp.kernel == GSPH_KERNEL_M4) return tab ? launch_stars<NDIM, GSPH_KERNEL_M4, true>(h) : launch_stars<NDIM, GSPH_KERNEL_M4, false>(h);
  if (h->dp.kernel == GSPH_KERNEL_GAUSSIAN) return tab ? launch_stars<NDIM, GSPH_KERNEL_GAUSSIAN, true>(h) : launch_stars<NDIM, GSPH_KERNEL_GAUSSIAN, false>(h);
  return tab ? launch_stars<NDIM, GSPH_KERNEL_QUINTIC, true>(h) : launch_stars<NDIM, GSPH_KERNEL_QUINTIC, false>(h);
}

/* HydroTree::UpdateAllStarGasForces (HydroTree.cpp:552-654): gravity of the gas on the star particles.
 * One warp per star walks the tree from the root, 32 nodes per round (Tree::ComputeStarGravityInteractionList,
 * Tree.cpp:748-885 -- the geometric opening test whatever gravity_mac says): leaves whose kernels may overlap
 * the star's give smoothed pair terms with the mean h, accepted cells their monopole / quadrupole at the
 * star's position, the remaining leaves unsoftened pair terms (NbodyHermite4::CalculateDirectHydroForces,
 * NbodyHermite4.cpp:151-247; adot is what the Hermite schemes use, the leapfrog classes ignore it).
 * stars: rows {r[NDIM], v[NDIM], h}; out: rows {a[NDIM], adot[NDIM], gpot} (sums, not added);
 * counts: per star neighbours, direct particles, cells.                                            */
#define STARGAS_STACK 4096
template <int NDIM, int KERN, bool TAB>
__global__ void __launch_bounds__(32)
star_gas_kernel(TreeDev T, PartView P, DevParams dp, const double* __restrict__ gtables,
                const double* __restrict__ stars, double* __restrict__ out, int* __restrict__ counts, int* __restrict__ err)
{
  __shared__ int stack[STARGAS_STACK];
  const int lane = threadIdx.x, s = blockIdx.x;
  const KernelEval<KERN, NDIM, TAB> K(dp, gtables);
  const double* sp = stars + (size_t) s*(2*NDIM + 1);
  double rs[3] = {0, 0, 0}, vs[3] = {0, 0, 0};
#pragma unroll
  for (int k = 0; k < NDIM; k++) { rs[k] = sp[k]; vs[k] = sp[NDIM + k]; }
  const double hs = sp[2*NDIM];
  const double halfrange = __dmul_rn(0.5, __dmul_rn(dp.kernrange, hs));
  const double twopi = 6.28318530717959;                  /* src/Headers/Constants.h:62 */
  const int leaf0 = (1 << T.ltot) - 1;
  double acc[3] = {0, 0, 0}, ad[3] = {0, 0, 0}, pot = 0.0;
  int nneib = 0, ndir = 0, ncell = 0;
  int top = 0;
  if (lane == 0) stack[0] = 0;
  top = 1;
  __syncwarp();
  while (top > 0) {
    const int take = min(top, 32);
    const int n = lane < take ? stack[top - 1 - lane] : -1;
    top -= take;
    __syncwarp();
    int mode = 0;                                         /* 1 smoothed leaf, 2 direct leaf, 3 cell, 4 open */
    if (n >= 0 && T.count[n] > 0) {
      double dr[3] = {0, 0, 0};
#pragma unroll
      for (int k = 0; k < NDIM; k++) dr[k] = T.rcell[k][n] - rs[k];
      const double drsqd = dot_exact<NDIM>(dr);
      const double thr = __dadd_rn(__dadd_rn(halfrange, T.rmax[n]), __dmul_rn(__dmul_rn(0.5, dp.kernrange), T.hmax[n]));
      const bool leaf = n >= leaf0;
      if (drsqd < __dmul_rn(thr, thr)) mode = leaf ? 1 : 4;
      else if (drsqd > T.cdistsqd[n]) mode = (leaf && T.count[n] == 1) ? 2 : 3;
      else mode = leaf ? 2 : 4;
    }
    /* children of the opened nodes */
    const unsigned open = __ballot_sync(FULLMASK, mode == 4);
    if (open) {
      const int before = __popc(open & ((1u << lane) - 1u));
      const int nopen = __popc(open);
      if (top + 2*nopen > STARGAS_STACK) { if (lane == 0) atomicExch(err, 1); return; }
      if (mode == 4) { stack[top + 2*before] = 2*n + 2; stack[top + 2*before + 1] = 2*n + 1; }
      top += 2*nopen;
    }
    __syncwarp();
    if (mode == 3) {
      ncell++;
      if (dp.multipole == GSPH_MULTIPOLE_MONOPOLE || dp.multipole == GSPH_MULTIPOLE_FAST_MONOPOLE) {
        /* ComputeCellMonopoleForces, NeighbourSearch.h:349-380 */
        double dr[3] = {0, 0, 0};
#pragma unroll
        for (int k = 0; k < NDIM; k++) dr[k] = T.com[k][n] - rs[k];
        const double drsqd = dot_exact<NDIM>(dr) + GSPH_SMALL;
        const double invdrsqd = 1.0/drsqd, invdrmag = sqrt(invdrsqd), invdr3 = invdrsqd*invdrmag, mc = T.m[n];
        pot += mc*invdrmag;
#pragma unroll
        for (int k = 0; k < NDIM; k++) acc[k] += mc*dr[k]*invdr3;
      }
      else {
        /* ComputeCellQuadrupoleForces, NeighbourSearch.h:390-470 */
        double dr[3] = {0, 0, 0};
#pragma unroll
        for (int k = 0; k < NDIM; k++) dr[k] = rs[k] - T.com[k][n];
        const double drsqd = dot_exact<NDIM>(dr) + GSPH_SMALL;
        const double invdrsqd = 1.0/drsqd, invdrmag = sqrt(invdrsqd), invdr5 = invdrsqd*invdrsqd*invdrmag, mc = T.m[n];
#pragma unroll
        for (int k = 0; k < NDIM; k++) acc[k] -= mc*dr[k]*invdrsqd*invdrmag;
        double q[5] = {0, 0, 0, 0, 0};
        for (int k = 0; k < 5; k++) q[k] = T.q[k][n];
        double qscalar, qfactor;
        if (NDIM == 3) {
          qscalar = q[0]*dr[0]*dr[0] + q[2]*dr[1]*dr[1] - (q[0] + q[2])*dr[2]*dr[2] +
                    2.0*(q[1]*dr[0]*dr[1] + q[3]*dr[0]*dr[2] + q[4]*dr[1]*dr[2]);
          qfactor = 2.5*qscalar*invdr5*invdrsqd;
          acc[0] += (q[0]*dr[0] + q[1]*dr[1] + q[3]*dr[2])*invdr5 - qfactor*dr[0];
          acc[1] += (q[1]*dr[0] + q[2]*dr[1] + q[4]*dr[2])*invdr5 - qfactor*dr[1];
          acc[2] += (q[3]*dr[0] + q[4]*dr[1] - (q[0] + q[2])*dr[2])*invdr5 - qfactor*dr[2];
        }
        else if (NDIM == 2) {
          qscalar = q[0]*dr[0]*dr[0] + q[2]*dr[1]*dr[1] + 2.0*q[1]*dr[0]*dr[1];
          qfactor = 2.5*qscalar*invdr5*invdrsqd;
          acc[0] += (q[0]*dr[0] + q[1]*dr[1])*invdr5 - qfactor*dr[0];
          acc[1] += (q[1]*dr[0] + q[2]*dr[1])*invdr5 - qfactor*dr[1];
        }
        else {
          qscalar = q[0]*dr[0]*dr[0];
          qfactor = 2.5*qscalar*invdr5*invdrsqd;
          acc[0] += (q[0]*dr[0])*invdr5 - qfactor*dr[0];
        }
        pot += mc*invdrmag + 0.5*qscalar*invdr5;
      }
    }
    else if (mode == 1 || mode == 2) {
      const int f = T.first[n], c = T.count[n];
      for (int j = f; j < f + c; j++) {
        if (((unsigned) P.i[I_FLAGS][j]) & GSPH_FLAG_DEAD) continue;
        double dr[3] = {0, 0, 0}, dv[3] = {0, 0, 0};
#pragma unroll
        for (int k = 0; k < NDIM; k++) { dr[k] = P.f[F_RX + k][j] - rs[k]; dv[k] = P.f[F_VX + k][j] - vs[k]; }
        const double drsqd = dot_exact<NDIM>(dr), drmag = sqrt(drsqd), invdrmag = 1.0/drmag;
        double dvdr = 0.0;
#pragma unroll
        for (int k = 0; k < NDIM; k++) dvdr += dv[k]*dr[k];
        const double drdt = dvdr*invdrmag, mj = P.f[F_M][j];
        if (mode == 1) {
          nneib++;
          const double invhmean = 2.0/(hs + P.f[F_H][j]);
          const double sj = drmag*invhmean;
          const double paux = mj*invhmean*invhmean*K.wgrav(sj)*invdrmag;
          double w0;
          if (TAB) w0 = sj >= K.kr ? 0.0 : kern_analytic<KERN, NDIM>(0, (K.kr/GSPH_TABLE_RES)*(int)(sj*K.res_invkr), K.kn);
          else w0 = kern_analytic<KERN, NDIM>(0, sj, K.kn);
          const double wkern = w0*(double) powf((float) invhmean, (float) NDIM);     /* powf: NbodyHermite4.cpp:194 */
#pragma unroll
          for (int k = 0; k < NDIM; k++) {
            acc[k] += paux*dr[k];
            ad[k] += paux*dv[k] - 3.0*paux*drdt*invdrmag*dr[k] + 2.0*twopi*mj*drdt*wkern*invdrmag*dr[k];
          }
          pot += mj*invhmean*K.wpot(sj);
        }
        else {
          ndir++;
          const double invdr3 = invdrmag*invdrmag*invdrmag;
#pragma unroll
          for (int k = 0; k < NDIM; k++) {
            acc[k] += mj*dr[k]*invdr3;
            ad[k] += mj*invdr3*(dv[k] - 3.0*drdt*invdrmag*dr[k]);
          }
          pot += mj*invdrmag;
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < NDIM; k++) { acc[k] = warp_sum(acc[k]); ad[k] = warp_sum(ad[k]); }
  pot = warp_sum(pot);
  nneib = (int) warp_sum((double) nneib); ndir = (int) warp_sum((double) ndir); ncell = (int) warp_sum((double) ncell);
  if (lane == 0) {
    double* o = out + (size_t) s*(2*NDIM + 1);
#pragma unroll
    for (int k = 0; k < NDIM; k++) { o[k] = acc[k]; o[NDIM + k] = ad[k]; }
    o[2*NDIM] = pot;
    counts[3*s] = nneib; counts[3*s + 1] = ndir; counts[3*s + 2] = ncell;
  }
}

template <int NDIM, int KERN, bool TAB>
int launch_star_gas(gsph_handle* h, const double* dstars, int Nstar, double* dout, int* dcounts, int* derr)
{
  PartView P;
  for (int k = 0; k < NF_DOUBLE; k++) P.f[k] = h->P[k];
  for (int k = 0; k < NF_INT; k++) P.i[k] = h->PI[k];
  star_gas_kernel<NDIM, KERN, TAB><<<Nstar, 32, 0, h->stream>>>(h->tree, P, h->dp, h->tables, dstars, dout, dcounts, derr);
  h->stats.kernel_launches++;
  CUDA_TRY(h, cudaGetLastError());
  return GSPH_OK;
}

template <int NDIM>
int dispatch_star_gas(gsph_handle* h, const double* dstars, int Nstar, double* dout, int* dcounts, int* derr)
{
  const bool tab = h->dp.tabulated != 0;
  if (h->dp.kernel == GSPH_KERNEL_M4)
    return tab ? launch_star_gas<NDIM, GSPH_KERNEL_M4, true>(h, dstars, Nstar, dout, dcounts, derr)
               : launch_star_gas<NDIM, GSPH_KERNEL_M4, false>(h, dstars, Nstar, dout, dcounts, derr);
  return tab ? launch_star_gas<NDIM, GSPH_KERNEL_QUINTIC, true>(h, dstars, Nstar, dout, dcounts, derr)
             : launch_star_gas<NDIM, GSPH_KERNEL_QUINTIC, false>(h, dstars, Nstar, dout, dcounts, derr);
}

}  // namespace

int gsph_run_star_gas(gsph_handle* h, const double* dstars, int Nstar, double* dout, int* dcounts, int* derr)
{
  switch (h->dp.ndim) {
  case 1: return dispatch_star_gas<1>(h, dstars, Nstar, dout, dcounts, derr);
  case 2: return dispatch_star_gas<2>(h, dstars, Nstar, dout, dcounts, derr);
  default: return dispatch_star_gas<3>(h, dstars, Nstar, dout, dcounts, derr);
  }
}

int gsph_run_star_forces(gsph_handle* h)
{
  if (h->Nstar <= 0) return GSPH_OK;
  switch (h->dp.ndim) {
  case 1: return dispatch_stars<1>(h);
  case 2: return dispatch_stars<2>(h);
  default: return dispatch_stars<3>(h);
  }
}

int gsph_run_hydro(gsph_handle* h, int with_gravity)
{
  switch (h->dp.ndim) {
  case 1: return dispatch_hydro<1>(h, with_gravity);
  case 2: return dispatch_hydro<2>(h, with_gravity);
  default: return dispatch_hydro<3>(h, with_gravity);
  }
}

int gsph_run_neiblist(gsph_handle* h, int32_t* counts, const int64_t* offsets, int32_t* idx)
{
  if (h->dp.any_mirror) { h->err = "neighbour-list inspection is not available with mirror boundaries"; return GSPH_EINVAL; }
  const bool fill = idx != nullptr;
  switch (h->dp.ndim) {
  case 1: return fill ? launch_neib<1, 1>(h, counts, offsets, idx) : launch_neib<1, 0>(h, counts, offsets, idx);
  case 2: return fill ? launch_neib<2, 1>(h, counts, offsets, idx) : launch_neib<2, 0>(h, counts, offsets, idx);
  default: return fill ? launch_neib<3, 1>(h, counts, offsets, idx) : launch_neib<3, 0>(h, counts, offsets, idx);
  }
}
