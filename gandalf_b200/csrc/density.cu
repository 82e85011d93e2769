/*=================================================================================================
 *  density.cu -- the h-rho pass.
 *
 *  Replaces  GradhSphTree::UpdateAllSphProperties   src/GradhSph/GradhSphTree.cpp:83-271
 *            Tree::ComputeGatherNeighbourList       src/Tree/Tree.cpp:291-381
 *            GradhSph::ComputeH                     src/GradhSph/GradhSph.cpp:142-326
 *            GradhSph::ComputeThermalProperties     src/GradhSph/GradhSph.cpp:335-347 (+ EOS classes)
 *
 *  Work decomposition: one warp per COHORT = one KD-tree node of <= 32 particles (a contiguous
 *  slab in tree order), one target particle per lane.  The warp walks the tree once for the
 *  cohort's box grown by kernrange*hmax*1.05^n, streams the candidate slabs through a 32-entry
 *  shared-memory tile (coalesced loads of packed (x,y,z,m) records, broadcast reads) and lets every
 *  lane compact the candidates inside its own search sphere into a private neighbour list; the
 *  fixed-point iteration of ComputeH then runs per lane over that list only, with exactly the
 *  reference's control flow.  A lane whose h outgrows the search radius makes the cohort redo the
 *  walk with a larger radius (the reference's `hmax *= 1.05` loop).
 *=================================================================================================*/
#include "device_common.cuh"

#define DENS_THREADS 256
#define DENS_WARPS (DENS_THREADS/32)
/* staged (x,y,z,m) records: two 16-byte chunks per record, chunk c of record s at unit 2s + (c ^ ((s>>2)&1)):
   the cooperative cp.async writes (two lanes per record) and the per-lane reads are both conflict-free */
#define DENS_UNIT(s, c) (2*(s) + ((c) ^ (((s) >> 2) & 1)))
#define DENS_WARP_D2 (2*32*2)                     /* double2 per warp: two record stages (2 KiB = walk stack) */

namespace {

/* gather walk on the outward-rounded float boxes (superset of Tree::ComputeGatherNeighbourList) */
template <int NDIM>
struct GatherVisitorF {
  const float4* __restrict__ nodeb; const WalkBox& wb; CandList& cl;
  float smin[3], smax[3];
  int first, count;
  int tag;                          /* image code << IMG_SHIFT of the walk (mirror boundaries) */
  __device__ __forceinline__ int test(int nd, bool)
  {
    const float4 lo = nodeb[2*(size_t) nd], hi = nodeb[2*(size_t) nd + 1];
    first = __float_as_int(lo.w); count = __float_as_int(hi.w);
    if (count <= 0) return 0;
    bool wrapped = false;
    return fboxes_overlap<NDIM>(wb, smin, smax, lo, hi, wrapped) ? 1 : 0;
  }
  __device__ __forceinline__ bool emit(bool take, int lane) { return cand_append_range(cl, first, count, take, lane, tag); }
};

/* the gather walk(s) of one cohort: its box grown by R, once per image code (mirror walls: the search box is
   reflected instead of the tree); candidates are appended to cl */
template <int NDIM, bool MIRROR>
__device__ __forceinline__ bool density_cohort_walk(const TreeDev& tree, const DevParams& dp, const WalkBox& wbx, CandList& cl,
                                                    int* stack, const double (&cmin)[3], const double (&cmax)[3], double R,
                                                    DevStatus* st)
{
  GatherVisitorF<NDIM> V{tree.nodeb, wbx, cl, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, 0, 0, 0};
  bool walk_ok = true;
  for (int code = 0; code < (MIRROR ? dp.nimg : 1) && walk_ok; code++) {
    const ImageOp op = image_op<NDIM>(dp, code);
    bool reach = true;
#pragma unroll
    for (int k = 0; k < NDIM; k++) {
      double lo = cmin[k] - R, hi = cmax[k] + R;
      if (op.refl[k]) {
        /* the image lies beyond the wall: only needed if the search box crosses that wall */
        const double w = 0.5*op.wall2[k];
        if (!(lo < w && hi > w) && !(lo <= w && hi >= w)) reach = false;
        const double l2 = op.wall2[k] - hi, h2 = op.wall2[k] - lo;
        lo = l2; hi = h2;
      }
      V.smin[k] = __double2float_rd(lo); V.smax[k] = __double2float_ru(hi);
    }
    if (!reach) continue;
    V.tag = code << IMG_SHIFT;
    __syncwarp();
    walk_ok = warp_walk_f<NB_WALK_STACK>(tree.ltot, stack, V, st);
  }
  return walk_ok;
}

/* Pre-walk (see hydro_walk_kernel): the gather walks of the first attempt of every cohort in a lean kernel of
   their own.  Search radius = kernrange * 1.05 * hmax of the cohort node, which is the largest first-attempt
   radius of its leaves (GradhSphTree.cpp:154); a cohort with inactive particles may get a few candidates more
   than density_kernel would collect -- the list only has to be a superset.                                   */
template <int NDIM, bool MIRROR>
__global__ void __launch_bounds__(DENS_THREADS)
density_walk_kernel(TreeDev tree, DevParams dp, int cohort_level, int c0, int nc, int* work_counter,
                    int32_t* __restrict__ pre_cand, int pre_cap, int* __restrict__ pre_count, const uint8_t* __restrict__ done)
{
  __shared__ int stacks[DENS_WARPS][NB_WALK_STACK];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
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
    if (ccount == 0 || ccount > 32 || tree.nactive[node] == 0 || done[c] || GSPH_RANGE_OUTSIDE_SLICE(dp, cfirst, ccount)) {
      if (lane == 0) pre_count[b] = -1;
      continue;
    }
    double cmin[3] = {0, 0, 0}, cmax[3] = {0, 0, 0};
#pragma unroll
    for (int k = 0; k < NDIM; k++) { cmin[k] = tree.bbmin[k][node]; cmax[k] = tree.bbmax[k][node]; }
    const double hmax = 1.05*tree.hmax[node];
    const double R = dp.kernrange*hmax;
    CandList cl{pre_cand + (size_t) b*pre_cap, pre_cap, 0, nullptr, true};
    const bool ok = density_cohort_walk<NDIM, MIRROR>(tree, dp, wbx, cl, stacks[warp], cmin, cmax, R, nullptr);
    __syncwarp();
    if (lane == 0) pre_count[b] = ok ? cl.n : -1;
  }
}

/* EOS: src/Headers/EOS.h:156, src/Thermal/{Adiabatic,Isothermal,Barotropic}EOS.cpp */
__device__ __forceinline__ void thermal(const DevParams& dp, double rho, double& u, double& sound,
                                        double& press)
{
  if (dp.eos == GSPH_EOS_ISOTHERMAL) u = dp.temp0/dp.gammam1/dp.mu_bar;
  else if (dp.eos == GSPH_EOS_BAROTROPIC)
    u = dp.temp0*(1.0 + pow(rho*(1.0/dp.rho_bary), dp.gammam1))/dp.gammam1/dp.mu_bar;
  else if (dp.eos == GSPH_EOS_BAROTROPIC2)                   /* Barotropic2EOS.cpp:78-83 */
    u = rho < dp.rho_bary ? dp.temp0/dp.gammam1/dp.mu_bar
                          : dp.temp0*pow(rho*(1.0/dp.rho_bary), dp.gammam1)/dp.gammam1/dp.mu_bar;
  if (dp.eos == GSPH_EOS_ADIABATIC) sound = sqrt(dp.gamma_eos*dp.gammam1*u);
  else sound = sqrt(dp.gammam1*u);
  press = dp.gammam1*rho*u;
}

template <int NDIM> __device__ __forceinline__ double root_ndim(double x)
{
  if (NDIM == 1) return x;
  if (NDIM == 2) return sqrt(x);
  return cbrt(x);
}
template <int NDIM> __device__ __forceinline__ double pow_ndim(double x)
{
  if (NDIM == 1) return x;
  if (NDIM == 2) return x*x;
  return x*x*x;
}

/*------------------------------------------------------------------------------------------------
 * Cullen & Dehnen (2010) viscosity switch, Sph<ndim>::ComputeCullenAndDehnenViscosity (Sph.h:364-456),
 * called at the end of ComputeH (GradhSph.cpp:319-321): velocity- and acceleration-gradient matrices
 * from the gather neighbours (their v and the PREVIOUS step's a), condition-number guard, Balsara-like
 * limiter, alpha = max(alpha, alpha_loc) and dalphadt.  Rarely used and register hungry (27 sums):
 * kept out of line, reading the lane's list and the records straight from global memory.
 *-----------------------------------------------------------------------------------------------*/
template <int NDIM, int KERN, bool TAB, bool MIRROR>
__device__ __forceinline__ void cd2010_switch(const PartView& P, const DevParams& dp, const double* __restrict__ gtables,
                                           const double2* __restrict__ geo4, const int32_t* nlist, int n, int lane,
                                           int ti, bool crosses, double hf, double rho, double hfactor, double sound)
{
  const KernelEval<KERN, NDIM, TAB> K(dp, gtables);       /* all six tables, in global memory */
  double rr[3][3], dv[3][3], da[3][3];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) rr[i][j] = dv[i][j] = da[i][j] = 0.0;
  double xi[3] = {0, 0, 0}, vi[3] = {0, 0, 0}, ai[3] = {0, 0, 0};
#pragma unroll
  for (int k = 0; k < NDIM; k++) { xi[k] = P.f[F_RX + k][ti]; vi[k] = P.f[F_VX + k][ti]; ai[k] = P.f[F_AX + k][ti]; }
  const double invh = 1.0/hf;
  const double hfac = invh*hfactor/rho;
  for (int kk = 0; kk < n; kk++) {
    const int e = nlist[(size_t) kk*32 + lane];
    const int j = MIRROR ? (e & IMG_MASK) : e;
    const double2 g0 = geo4[2*(size_t) j], g1 = geo4[2*(size_t) j + 1];
    if (!(g1.y > 0.0)) continue;                           /* dead */
    double xs[3] = {g0.x, g0.y, g1.x}, vj[3] = {0, 0, 0}, aj[3] = {0, 0, 0};
#pragma unroll
    for (int k = 0; k < NDIM; k++) { vj[k] = P.f[F_VX + k][j]; aj[k] = P.f[F_AX + k][j]; }
    if (MIRROR && (e >> IMG_SHIFT)) {
      const ImageOp op = image_op<NDIM>(dp, e >> IMG_SHIFT);
#pragma unroll
      for (int k = 0; k < NDIM; k++) {
        xs[k] = reflect_coord(op, k, xs[k]);
        if (op.refl[k]) { vj[k] = -vj[k]; aj[k] = -aj[k]; }
      }
    }
    double dr[3] = {0, 0, 0};
#pragma unroll
    for (int k = 0; k < NDIM; k++) dr[k] = xs[k] - xi[k];
    if (crosses) wrap_dr<NDIM>(dp, dr);
    const double w = g1.y*hfac*K.w1(invh*sqrt(dot_exact<NDIM>(dr)));
    if (w == 0.0) continue;
#pragma unroll
    for (int a = 0; a < NDIM; a++)
#pragma unroll
      for (int b = 0; b < NDIM; b++) {
        rr[a][b] += w*dr[a]*dr[b];
        dv[a][b] += w*dr[a]*(vj[b] - vi[b]);
        da[a][b] += w*dr[a]*(aj[b] - ai[b]);
      }
  }
  /* InvertMatrix, InlineFuncs.h:573-612 */
  double T[3][3];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) T[i][j] = 0.0;
  if (NDIM == 1) T[0][0] = 1.0/rr[0][0];
  else if (NDIM == 2) {
    const double invdet = 1.0/(rr[0][0]*rr[1][1] - rr[0][1]*rr[1][0]);
    T[0][0] = invdet*rr[1][1]; T[0][1] = -invdet*rr[0][1]; T[1][0] = -invdet*rr[1][0]; T[1][1] = invdet*rr[0][0];
  }
  else {
    const double (&A)[3][3] = rr;
    const double invdet = 1.0/(A[0][0]*(A[1][1]*A[2][2] - A[2][1]*A[1][2]) - A[0][1]*(A[1][0]*A[2][2] - A[1][2]*A[2][0]) +
                               A[0][2]*(A[1][0]*A[2][1] - A[1][1]*A[2][0]));
    T[0][0] = (A[1][1]*A[2][2] - A[2][1]*A[1][2])*invdet; T[0][1] = (A[0][2]*A[2][1] - A[0][1]*A[2][2])*invdet;
    T[0][2] = (A[0][1]*A[1][2] - A[0][2]*A[1][1])*invdet; T[1][0] = (A[1][2]*A[2][0] - A[1][0]*A[2][2])*invdet;
    T[1][1] = (A[0][0]*A[2][2] - A[0][2]*A[2][0])*invdet; T[1][2] = (A[1][0]*A[0][2] - A[0][0]*A[1][2])*invdet;
    T[2][0] = (A[1][0]*A[2][1] - A[2][0]*A[1][1])*invdet; T[2][1] = (A[2][0]*A[0][1] - A[0][0]*A[2][1])*invdet;
    T[2][2] = (A[0][0]*A[1][1] - A[1][0]*A[0][1])*invdet;
  }
  double modR = 0.0, modT = 0.0;
#pragma unroll
  for (int i = 0; i < NDIM; i++)
#pragma unroll
    for (int j = 0; j < NDIM; j++) { modR += rr[i][j]*rr[i][j]; modT += T[i][j]*T[i][j]; }
  const double sqd_condition_number = modR*modT/(NDIM*NDIM);
  double alpha_loc = 0.0;
  if (sqd_condition_number > 1.0e4) alpha_loc = dp.alpha_visc;
  else {
    double dvdx[3][3], dadx[3][3];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) dvdx[i][j] = dadx[i][j] = 0.0;
#pragma unroll
    for (int i = 0; i < NDIM; i++)
#pragma unroll
      for (int j = 0; j < NDIM; j++)
#pragma unroll
        for (int k = 0; k < NDIM; k++) { dvdx[i][j] += T[j][k]*dv[k][i]; dadx[i][j] += T[j][k]*da[k][i]; }
    double ddivdt = 0.0, divv2 = 0.0;
#pragma unroll
    for (int i = 0; i < NDIM; i++) {
      ddivdt += dadx[i][i];
#pragma unroll
      for (int j = 0; j < NDIM; j++) ddivdt -= dvdx[i][j]*dvdx[j][i];
      divv2 += dvdx[i][i];
    }
    divv2 *= divv2;
    double curlv2 = 0.0;                                   /* CurlVelSqd, Sph.h:340-357 */
    if (NDIM == 2) { const double c = dvdx[1][0] - dvdx[0][1]; curlv2 = c*c; }
    if (NDIM == 3) {
      const double c0 = dvdx[1][2] - dvdx[2][1], c1 = dvdx[2][0] - dvdx[0][2], c2 = dvdx[0][1] - dvdx[1][0];
      curlv2 = c0*c0 + c1*c1 + c2*c2;
    }
    double f_balsara = 1.0;
    if (curlv2 > 0.0) f_balsara = divv2/(divv2 + curlv2);
    if (ddivdt < 0.0) {
      alpha_loc = (10.0*hf*hf/(sound*sound))*f_balsara*(-ddivdt);
      alpha_loc = fmin(alpha_loc, dp.alpha_visc);
    }
  }
  double alpha = P.f[F_ALPHA][ti];
  if (alpha_loc > alpha) alpha = alpha_loc;
  P.f[F_ALPHA][ti] = alpha;
  P.f[F_DALPHADT][ti] = 0.1*sound*(fmax(dp.alpha_visc_min, alpha_loc) - alpha)*invh;
}

/*------------------------------------------------------------------------------------------------
 * One warp per cohort (KD-tree node of <= 32 particles), one target per lane.
 *  phase 1  the candidate slabs found by the walk are streamed once through a 32-entry shared
 *           tile of packed (x,y,z,m) records; every lane appends the candidates inside its search
 *           sphere kernrange*hmax to its own neighbour-id list (global scratch, [k][lane] layout:
 *           coalesced rows).
 *  phase 2  the ComputeH fixed-point iteration runs over the lane's own list only (gathering the
 *           32-byte records through L1/L2), so every lane does useful work in every instruction.
 *-----------------------------------------------------------------------------------------------*/
template <int NDIM, int KERN, bool TAB, bool MIRROR, bool CD2010>
__global__ void __launch_bounds__(DENS_THREADS, CD2010 ? 2 : 3)
density_kernel(PartView P, TreeDev tree, DevParams dp, const double* __restrict__ gtables,
               const double2* __restrict__ geo4, int cohort_level, int* work_counter,
               int32_t* cand_all, int cand_cap, int32_t* nlist_all, int nlist_cap, DevStatus* st,
               uint8_t* done, double* __restrict__ hnew,
               const int32_t* __restrict__ pre_cand, int pre_cap, const int* __restrict__ pre_count, int pre_c0, int pre_n)
{
  extern __shared__ double smem[];
  double* tab = smem;                                          /* 3 x 1000 (tabulated only) */
  /* per-warp region, time-shared: walk stack | float tile x|y|z|w + ids (phase 1) | two record stages (phase 2) */
  double2* stage = reinterpret_cast<double2*>(smem + (TAB ? 3*GSPH_TABLE_RES : 0)) + (threadIdx.x >> 5)*DENS_WARP_D2;
  const int lane = threadIdx.x & 31;
  if (TAB) {
    for (int i = threadIdx.x; i < 3*GSPH_TABLE_RES; i += blockDim.x) tab[i] = gtables[i];
    __syncthreads();
  }
  int* stack = reinterpret_cast<int*>(stage);
  float* ftile = reinterpret_cast<float*>(stage);
  int* tj = reinterpret_cast<int*>(ftile + 128);
  const int warp = threadIdx.x >> 5;
  const WalkBox wbx(dp);
  const int gwarp = blockIdx.x*DENS_WARPS + warp;
  int32_t* const mycand = cand_all + (size_t) gwarp*cand_cap;
  int32_t* nlist = nlist_all + (size_t) gwarp*nlist_cap*32;
  const KernelEval<KERN, NDIM, TAB> K(dp, tab);
  const int ncohort = 1 << cohort_level;
  const double krs = dp.kernrangesqd;
  const double invndim = 1.0/NDIM;

  unsigned long long loc_iter = 0, loc_targets = 0, loc_cand = 0; int loc_itmax = 0, loc_retry = 0;

  while (true) {
    int c = 0;
    if (lane == 0) c = atomicAdd(work_counter, 1);
    c = __shfl_sync(FULLMASK, c, 0);
    if (c >= ncohort) break;
    const int node = ncohort - 1 + c;
    const int cfirst = tree.first[node], ccount = tree.count[node];
    if (ccount == 0 || tree.nactive[node] == 0 || done[c] || GSPH_RANGE_OUTSIDE_SLICE(dp, cfirst, ccount)) continue;
    bool overflow = false;
    double cmin[3], cmax[3];
#pragma unroll
    for (int k = 0; k < NDIM; k++) { cmin[k] = tree.bbmin[k][node]; cmax[k] = tree.bbmax[k][node]; }

    for (int t0 = 0; t0 < ccount && !overflow; t0 += 32) {
      /* ---- the target of this lane ---------------------------------------------------------------*/
      const int ti = cfirst + t0 + lane;
      bool live = t0 + lane < ccount && GSPH_IN_SLICE(dp, ti);
      if (live) {
        const unsigned fl = (unsigned) P.i[I_FLAGS][ti];
        live = (fl & GSPH_FLAG_ACTIVE) && !(fl & GSPH_FLAG_DEAD);
      }
      double x[3] = {0, 0, 0}, mi = 1.0, h0 = 1.0;
      if (live) {
#pragma unroll
        for (int k = 0; k < NDIM; k++) x[k] = P.f[F_RX + k][ti];
        mi = P.f[F_M][ti]; h0 = P.f[F_H][ti];
      }

      /* The reference iterates per LEAF CELL with hmax = 1.05^n * cell.hmax as search radius and
         as upper bisection bound, and redoes the whole leaf when one of its particles outgrows it
         (GradhSphTree.cpp:144-240).  The same tree gives the same leaves here, so every lane
         carries its own leaf's hmax and retry count; the walk uses the largest radius in the warp. */
      int leafnode = -1 - lane;
      double hmax_l = 0.0;
      if (t0 + lane < ccount) {
        int nd = node;
        const int leaf0 = (1 << tree.ltot) - 1;
        while (nd < leaf0) nd = (ti < tree.first[2*nd + 2]) ? 2*nd + 1 : 2*nd + 2;
        leafnode = nd;
        hmax_l = tree.hmax[nd];
      }
      bool fin = !live;
      for (int attempt = 0; ; attempt++) {
        if (!fin) hmax_l = 1.05*hmax_l;                        /* GradhSphTree.cpp:154 */
        double hmax = fin ? 0.0 : hmax_l;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) hmax = fmax(hmax, __shfl_xor_sync(FULLMASK, hmax, o));
        if (!(hmax > 0.0)) break;                              /* every lane is finished */
        /* ---- candidates whose leaf overlaps the grown cohort box: first attempt from the pre-walk kernel,
         *      otherwise (no pre-walk, its slot overflowed, or a larger radius is needed) walked here -------*/
        const double R = dp.kernrange*hmax;
        bool crosses = false, wide = false;
#pragma unroll
        for (int k = 0; k < NDIM; k++) {
          const double lo = cmin[k] - R, hi = cmax[k] + R;
          if (dp.periodic[k] && (lo < dp.boxmin[k] || hi > dp.boxmax[k] || hi - lo > dp.boxhalf[k])) crosses = true;
          if (dp.periodic[k] && 0.5*(hi - lo) >= 0.49*dp.boxsize[k]) wide = true;
        }
        const int32_t* cand = mycand;
        int ncand = 0;
        const int pre = (attempt == 0 && t0 == 0 && c >= pre_c0 && c < pre_c0 + pre_n) ? pre_count[c - pre_c0] : -1;
        if (pre >= 0) { ncand = pre; cand = pre_cand + (size_t)(c - pre_c0)*pre_cap; }
        else {
          CandList cl{mycand, cand_cap, 0, st};
          if (!density_cohort_walk<NDIM, MIRROR>(tree, dp, wbx, cl, stack, cmin, cmax, R, st)) { overflow = true; break; }
          ncand = cl.n;
        }

        /* ---- phase 1: per-lane candidate lists on the FP32 pipe.  The list only has to be a SUPERSET
         *      of the particles inside kernrange*h for every h <= hmax_l (membership is decided in
         *      phase 2 by the exact ssqd < kernrangesqd test), so positions relative to the cohort
         *      centre, scaled by the search radius, are compared in single precision with a 1e-3
         *      margin; lanes whose radius is tiny against the cohort size keep every candidate. -------*/
        const double sc = 1.0/R;
        float xf[3] = {0.f, 0.f, 0.f};
        float thr = -1.0f;
        if (!fin) {
          double diag2 = 0.0;
#pragma unroll
          for (int k = 0; k < NDIM; k++) {
            xf[k] = (float)((x[k] - 0.5*(cmin[k] + cmax[k]))*sc);
            diag2 += 0.25*(cmax[k] - cmin[k])*(cmax[k] - cmin[k]);
          }
          const double q = hmax_l/hmax;
          thr = __double2float_ru(q*q*1.001);
          if (wide || diag2 > 1.0e6*krs*hmax_l*hmax_l) thr = __int_as_float(0x7f800000);
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
            float fx = __int_as_float(0x7f800000), fy = 0.f, fz = 0.f;
            if (jl < ncand) {
              j = jcur;
              const int jp = MIRROR ? (j & IMG_MASK) : j;
              const double2 a = geo4[2*(size_t) jp], b = geo4[2*(size_t) jp + 1];
              double xs[3] = {a.x, a.y, b.x};
              if (MIRROR && (j >> IMG_SHIFT)) {
                const ImageOp op = image_op<NDIM>(dp, j >> IMG_SHIFT);
#pragma unroll
                for (int k = 0; k < NDIM; k++) xs[k] = reflect_coord(op, k, xs[k]);
              }
              double dr[3] = {0, 0, 0};
              dr[0] = xs[0] - 0.5*(cmin[0] + cmax[0]);
              if (NDIM > 1) dr[1] = xs[1] - 0.5*(cmin[1] + cmax[1]);
              if (NDIM > 2) dr[2] = xs[2] - 0.5*(cmin[2] + cmax[2]);
              if (dp.any_periodic) wrap_dr<NDIM>(dp, dr);
              if (b.y > 0.0) fx = (float)(dr[0]*sc);
              fy = (float)(dr[1]*sc); fz = (float)(dr[2]*sc);
            }
            ftile[lane] = fx; ftile[32 + lane] = fy; ftile[64 + lane] = fz;
            tj[lane] = j;
          }
          __syncwarp();
          unsigned pass = filter_tile<false>(ftile, xf, thr);
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
        loc_cand += (unsigned long long) ncand;
        __syncwarp();                                          /* tile region becomes the record stages */

        /* ---- phase 2: ComputeH (GradhSph.cpp:142-326) over the lane's own list.  Per list slot the
         *      warp gathers the 32 (x,y,z,m) records cooperatively (two lanes per record, 16-byte
         *      cp.async copies into a padded stage, double buffered). -----------------------------------*/
        auto gather = [&](int jv, double2* buf) {
#pragma unroll
          for (int r = 0; r < 2; r++) {
            const int src = r*16 + (lane >> 1);
            const int jj = __shfl_sync(FULLMASK, jv, src);
            if (jj >= 0) cp_async16_cg(buf + DENS_UNIT(src, lane & 1), geo4 + 2*(size_t)(MIRROR ? (jj & IMG_MASK) : jj) + (lane & 1));
          }
          cp_async_commit();
        };
        double h = h0, hlo = 0.0, hup = hmax_l, rho = 0.0, invom = 0.0, zeta = 0.0;
        int iter = 0, state = fin ? 1 : 0;     /* 0 iterating, 1 left the loop, 2 need bigger list, 3 failed */
        while (__any_sync(FULLMASK, state == 0)) {
          const double invh = 1.0/h, invhsqd = invh*invh;
          double s_rho = 0.0, s_om = 0.0, s_ze = 0.0;
          const int kend = state == 0 ? n : 0;
          int nmax = kend;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) nmax = max(nmax, __shfl_xor_sync(FULLMASK, nmax, o));
          int jc = (0 < kend) ? nlist[lane] : -1;
          int jn = (1 < kend) ? nlist[32 + lane] : -1;
          gather(jc, stage);
          for (int k = 0; k < nmax; k++) {
            const double2* buf = stage + (k & 1)*64;
            const bool more = k + 1 < nmax;
            if (more) gather(jn, stage + ((k + 1) & 1)*64);
            const int jnn = (k + 2 < kend) ? nlist[(size_t)(k + 2)*32 + lane] : -1;
            if (more) cp_async_wait<1>(); else cp_async_wait<0>();
            __syncwarp();
            if (jc >= 0) {
              const double2 a = buf[DENS_UNIT(lane, 0)], b = buf[DENS_UNIT(lane, 1)];
              double xs[3] = {a.x, a.y, b.x};
              if (MIRROR && (jc >> IMG_SHIFT)) {          /* a mirror image: 2*wall - r_k */
                const ImageOp op = image_op<NDIM>(dp, jc >> IMG_SHIFT);
#pragma unroll
                for (int k = 0; k < NDIM; k++) xs[k] = reflect_coord(op, k, xs[k]);
              }
              double dr[3] = {0, 0, 0};
              dr[0] = xs[0] - x[0];
              if (NDIM > 1) dr[1] = xs[1] - x[1];
              if (NDIM > 2) dr[2] = xs[2] - x[2];
              if (crosses) wrap_dr<NDIM>(dp, dr);
              const double ssqd = __dmul_rn(invhsqd, dot_exact<NDIM>(dr));
              if (ssqd < krs) {
                double w0, wom, wze;
                K.density_terms(ssqd, w0, wom, wze);
                s_rho += b.y*w0;
                s_om += b.y*invh*wom;
                s_ze += b.y*wze;
              }
            }
            __syncwarp();
            jc = jn; jn = jnn;
          }
          if (state == 0) {
            /* fixed-point / bisection update (GradhSph.cpp:206-257) */
            iter++;
            const double hfac = pow_ndim<NDIM>(invh);
            rho = s_rho*hfac; invom = s_om*hfac; zeta = s_ze*invhsqd;
            const double hnew = dp.h_fac*root_ndim<NDIM>(mi/rho);
            if (rho > 0.0 && h > hlo && fabs(h - hnew)*invh < dp.h_converge) state = 1;
            else {
              if (iter < 30) h = hnew;
              else if (iter == 30) h = 0.5*(hlo + hup);
              else if (iter < 150) {
                if (rho < GSPH_SMALL || h > hnew) hup = h; else hlo = h;
                h = 0.5*(hlo + hup);
              }
              else state = 3;
              if (state == 0) {
                if (h > hmax_l) state = 2;
                else if (!(h > hlo && h < hup)) state = 1;
              }
            }
          }
        }
        if (__any_sync(FULLMASK, state == 3)) {
          if (lane == 0) atomicExch(&st->err_converge, 1);
          return;
        }
        /* a leaf is redone as a whole when one of its particles needs a larger list */
        const unsigned peers = __match_any_sync(FULLMASK, leafnode);
        const unsigned need = __ballot_sync(FULLMASK, state == 2);
        const bool redo = (need & peers) != 0u;
        if (need) {
          loc_retry++;
          if (attempt > 60) { if (lane == 0) atomicExch(&st->err_hsearch, 1); return; }
        }
        /* ---- normalise and store (GradhSph.cpp:260-317) ------------------------------------------*/
        if (!fin && !redo) {
          const double hf = fmax(dp.h_fac*root_ndim<NDIM>(mi/rho), hlo);
          const double ih = 1.0/hf;
          hnew[ti] = hf;                                        /* committed by the host after the pass */
          P.f[F_RHO][ti] = rho;
          P.f[F_HFAC][ti] = pow_ndim<NDIM>(ih)*ih;
          P.f[F_HR2][ti] = __dmul_rn(__dmul_rn(krs, hf), hf);
          P.f[F_DIVV][ti] = 0.0;
          double io = 1.0, ze = 0.0;
          if (P.i[I_SINKID][ti] == -1) {
            const double deriv = -invndim*hf/rho;                /* h_rho_deriv, Sph.h:264-268 */
            io = 1.0/(1.0 - deriv*invom);
            double zs = zeta;
            if (dp.star_zeta && dp.nstar > 0) {
              /* SPH-star conservative-gravity terms of zeta (Hubber et al. 2013, GradhSph.cpp:288-306):
                 every star, with the mean of the two smoothing lengths or the particle's own          */
              for (int j = 0; j < dp.nstar; j++) {
                const double* s = dp.stars + (size_t) j*(NDIM + 2);
                double ihm = dp.nbody_softening == 1 ? 2.0/(hf + s[NDIM + 1]) : 2.0*ih;
                const double ihs2 = dp.nbody_softening == 1 ? ihm*ihm : 4.0*ih*ih;
                double dr[3] = {0, 0, 0};
#pragma unroll
                for (int k = 0; k < NDIM; k++) dr[k] = s[k] - x[k];
                const double ssqd = dot_exact<NDIM>(dr)*ihs2;
                if (ssqd < krs) {
                  double w0, wom, wze;
                  K.density_terms(ssqd, w0, wom, wze);
                  zs += s[NDIM]*ihs2*wze;
                }
              }
            }
            ze = deriv*zs*io;
          }
          P.f[F_INVOMEGA][ti] = io;
          P.f[F_ZETA][ti] = ze;
          double u = P.f[F_U][ti], cs, pr;
          thermal(dp, rho, u, cs, pr);
          P.f[F_U][ti] = u; P.f[F_SOUND][ti] = cs; P.f[F_PRESS][ti] = pr;
          if (CD2010)
            cd2010_switch<NDIM, KERN, TAB, MIRROR>(P, dp, gtables, geo4, nlist, n, lane, ti, crosses, hf, rho,
                                                   pow_ndim<NDIM>(ih)*ih, cs);
          if (dp.create_sinks) {
            /* potential-minimum flag for the sink search (GradhSph.cpp:270-280): no neighbour inside the
               kernel (of the last iterate of h) has a larger gpot.  The reference evaluates the distance
               of the PREVIOUS neighbour of its list here, which makes its flag depend
               on the list order; this is the intended test, with each neighbour's own distance.     */
            const double gp = P.f[F_GPOT][ti], ihs = 1.0/(h*h);
            bool pm = true;
            for (int k = 0; k < n; k++) {
              const int e = nlist[(size_t) k*32 + lane];
              const int j = MIRROR ? (e & IMG_MASK) : e;
              const double2 a = geo4[2*(size_t) j], b = geo4[2*(size_t) j + 1];
              double xs[3] = {a.x, a.y, b.x};
              if (MIRROR && (e >> IMG_SHIFT)) {
                const ImageOp op = image_op<NDIM>(dp, e >> IMG_SHIFT);
#pragma unroll
                for (int kk = 0; kk < NDIM; kk++) xs[kk] = reflect_coord(op, kk, xs[kk]);
              }
              double dr[3] = {0, 0, 0};
              dr[0] = xs[0] - x[0];
              if (NDIM > 1) dr[1] = xs[1] - x[1];
              if (NDIM > 2) dr[2] = xs[2] - x[2];
              if (crosses) wrap_dr<NDIM>(dp, dr);
              if (P.f[F_GPOT][j] > 1.000000001*gp && dot_exact<NDIM>(dr)*ihs < krs) pm = false;
            }
            const unsigned fl = (unsigned) P.i[I_FLAGS][ti];
            P.i[I_FLAGS][ti] = (int32_t)(pm ? (fl | GSPH_FLAG_POTMIN) : (fl & ~(unsigned) GSPH_FLAG_POTMIN));
          }
          loc_iter += iter; loc_targets++; loc_itmax = max(loc_itmax, iter);
          fin = true;
        }
        if (!need) break;
      }
    }
    __syncwarp();
    if (!overflow && lane == 0) done[c] = 1;
  }
  /* statistics */
  loc_iter = (unsigned long long) warp_sum((double) loc_iter);
  loc_targets = (unsigned long long) warp_sum((double) loc_targets);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) loc_itmax = max(loc_itmax, __shfl_xor_sync(FULLMASK, loc_itmax, o));
  if (lane == 0) {
    atomicAdd(&st->hiter_sum, loc_iter); atomicAdd(&st->ntarget, loc_targets);
    atomicAdd(&st->cand_sum, loc_cand);
    atomicMax(&st->hiter_max, loc_itmax);
    if (loc_retry) atomicAdd(&st->density_retries, loc_retry);
  }
}

template <int NDIM, int KERN, bool TAB, bool MIRROR, bool CD2010>
int launch_density_m(gsph_handle* h)
{
  const int Lc = cohort_level_for(h->N, h->tree.ltot, 32);
  static_assert(DENS_WARP_D2*sizeof(double2) >= NB_WALK_STACK*sizeof(int), "walk stack must fit the stage region");
  const size_t smem = (TAB ? 3*GSPH_TABLE_RES*sizeof(double) : 0) + DENS_WARPS*DENS_WARP_D2*sizeof(double2);
  auto kern = density_kernel<NDIM, KERN, TAB, MIRROR, CD2010>;
  CUDA_TRY(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  int bps = 0;
  CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, DENS_THREADS, smem));
  if (bps < 1) bps = 1;
  int blocks = h->sm_count*bps;
  const int ncohort = 1 << Lc;
  const int need_blocks = (ncohort + DENS_WARPS - 1)/DENS_WARPS;
  if (blocks > need_blocks) blocks = need_blocks;
  int rc = gsph_ensure_cand(h, blocks*DENS_WARPS, 0);
  if (rc) return rc;
  rc = gsph_ensure_nlist(h, blocks*DENS_WARPS, 0);
  if (rc) return rc;
  CUDA_TRY(h, cudaMemsetAsync(h->work_counter, 0, sizeof(int), h->stream));
  PartView P;
  for (int k = 0; k < NF_DOUBLE; k++) P.f[k] = h->P[k];
  for (int k = 0; k < NF_INT; k++) P.i[k] = h->PI[k];
  KernelTimer kt(h, 0);
  /* the first-attempt walks in their own lean kernel (density_walk_kernel) */
  rc = gsph_prewalk_prepare(h, Lc);
  if (rc) return rc;
  if (h->pre_n > 0) {
    CUDA_TRY(h, cudaMemsetAsync(h->work_counter + 1, 0, sizeof(int), h->stream));
    auto wk = density_walk_kernel<NDIM, MIRROR>;
    int wbps = 0;
    CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&wbps, wk, DENS_THREADS, 0));
    const int wblocks = std::min(h->sm_count*std::max(wbps, 1), (h->pre_n + DENS_WARPS - 1)/DENS_WARPS);
    wk<<<wblocks, DENS_THREADS, 0, h->stream>>>(h->tree, h->dp, Lc, h->pre_c0, h->pre_n, h->work_counter + 1, h->pre_cand,
                                                h->pre_cap, h->pre_count, h->done);
    h->stats.kernel_launches++;
  }
  kern<<<blocks, DENS_THREADS, smem, h->stream>>>(P, h->tree, h->dp, h->tables, (const double2*) h->geo4, Lc,
                                                  h->work_counter, h->cand, (int) h->cand_cap_per_warp,
                                                  h->nlist, (int) h->nlist_cap, h->status, h->done, h->hnew,
                                                  h->pre_cand, h->pre_cap, h->pre_count, h->pre_c0, h->pre_n);
  h->stats.kernel_launches++;
  CUDA_TRY(h, cudaGetLastError());
  return GSPH_OK;
}

template <int NDIM, int KERN, bool TAB>
int launch_density(gsph_handle* h)
{
  /* mirror walls are a separate instantiation: boxes without them pay nothing for the image codes */
  if (h->dp.tdavisc == GSPH_TDAVISC_CD2010) {          /* the Cullen & Dehnen switch: its own (register hungry) instantiation */
    if (h->dp.any_mirror) { h->err = "time_dependent_avisc = cd2010 with mirror boundaries is not on the accelerated path"; return GSPH_EINVAL; }
    return launch_density_m<NDIM, KERN, TAB, false, true>(h);
  }
  return h->dp.any_mirror ? launch_density_m<NDIM, KERN, TAB, true, false>(h) : launch_density_m<NDIM, KERN, TAB, false, false>(h);
}

template <int NDIM>
int dispatch_density(gsph_handle* h)
{
  const bool tab = h->dp.tabulated != 0;
  if (h->dp.kernel == GSPH_KERNEL_M4)
    return tab ? launch_density<NDIM, GSPH_KERNEL_M4, true>(h) : launch_density<NDIM, GSPH_KERNEL_M4, false>(h);
  if (h->dp.kernel == GSPH_KERNEL_GAUSSIAN)
    return tab ? launch_density<NDIM, GSPH_KERNEL_GAUSSIAN, true>(h) : launch_density<NDIM, GSPH_KERNEL_GAUSSIAN, false>(h);
  return tab ? launch_density<NDIM, GSPH_KERNEL_QUINTIC, true>(h)
             : launch_density<NDIM, GSPH_KERNEL_QUINTIC, false>(h);
}

/* Multi-GPU: hrangesqd, hfactor, u, sound and pressure of the particles of the OTHER slices follow from
 * their exchanged h and rho exactly as the density kernel's store section computes them, so only h, rho,
 * invomega and zeta have to cross NVLink.                                                            */
template <int NDIM>
__global__ void derive_density_fields_kernel(int N, int lo, int hi, PartView P, DevParams dp)
{
  const int ti = blockIdx.x*blockDim.x + threadIdx.x;
  if (ti >= N || (ti >= lo && ti < hi)) return;
  const unsigned fl = (unsigned) P.i[I_FLAGS][ti];
  if ((fl & GSPH_FLAG_DEAD) || !(fl & GSPH_FLAG_ACTIVE)) return;     /* (the density pass left the inactive ones alone) */
  const double hf = P.f[F_H][ti], rho = P.f[F_RHO][ti];
  const double ih = 1.0/hf;
  P.f[F_HFAC][ti] = pow_ndim<NDIM>(ih)*ih;
  P.f[F_HR2][ti] = __dmul_rn(__dmul_rn(dp.kernrangesqd, hf), hf);
  double u = P.f[F_U][ti], cs, pr;
  thermal(dp, rho, u, cs, pr);
  P.f[F_U][ti] = u; P.f[F_SOUND][ti] = cs; P.f[F_PRESS][ti] = pr;
}

/* GradhSph::ComputeThermalProperties (GradhSph.cpp:335-347) for EVERY particle, active or not: what MainLoop does
 * before the force pass when nradstep = 1 (SphSimulation.cpp:669-679).  With block time steps the inactive
 * neighbours then enter the force sums with the sound speed and pressure of their PREDICTED u.            */
__global__ void thermal_all_kernel(int N, PartView P, DevParams dp)
{
  const int ti = blockIdx.x*blockDim.x + threadIdx.x;
  if (ti >= N) return;
  if (((unsigned) P.i[I_FLAGS][ti]) & GSPH_FLAG_DEAD) return;
  double u = P.f[F_U][ti], cs, pr;
  thermal(dp, P.f[F_RHO][ti], u, cs, pr);
  P.f[F_U][ti] = u; P.f[F_SOUND][ti] = cs; P.f[F_PRESS][ti] = pr;
}

}  // namespace

int gsph_thermal_all_impl(gsph_handle* h)
{
  const int N = (int) h->N;
  PartView P;
  for (int k = 0; k < NF_DOUBLE; k++) P.f[k] = h->P[k];
  for (int k = 0; k < NF_INT; k++) P.i[k] = h->PI[k];
  thermal_all_kernel<<<(N + 255)/256, 256, 0, h->stream>>>(N, P, h->dp);
  h->stats.kernel_launches++;
  CUDA_TRY(h, cudaGetLastError());
  return GSPH_OK;
}

int gsph_derive_density_fields_impl(gsph_handle* h, int lo, int hi)
{
  const int N = (int) h->N;
  PartView P;
  for (int k = 0; k < NF_DOUBLE; k++) P.f[k] = h->P[k];
  for (int k = 0; k < NF_INT; k++) P.i[k] = h->PI[k];
  switch (h->dp.ndim) {
  case 1: derive_density_fields_kernel<1><<<(N + 255)/256, 256, 0, h->stream>>>(N, lo, hi, P, h->dp); break;
  case 2: derive_density_fields_kernel<2><<<(N + 255)/256, 256, 0, h->stream>>>(N, lo, hi, P, h->dp); break;
  default: derive_density_fields_kernel<3><<<(N + 255)/256, 256, 0, h->stream>>>(N, lo, hi, P, h->dp); break;
  }
  h->stats.kernel_launches++;
  CUDA_TRY(h, cudaGetLastError());
  return GSPH_OK;
}

int gsph_run_density(gsph_handle* h)
{
  switch (h->dp.ndim) {
  case 1: return dispatch_density<1>(h);
  case 2: return dispatch_density<2>(h);
  default: return dispatch_density<3>(h);
  }
}
