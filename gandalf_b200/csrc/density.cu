/*=================================================================================================
 *  density.cu -- the h-rho pass.
 *
 *  Replaces  GradhSphTree::UpdateAllSphProperties   src/GradhSph/GradhSphTree.cpp:83-271
 *            Tree::ComputeGatherNeighbourList       src/Tree/Tree.cpp:291-381
 *            GradhSph::ComputeH                     src/GradhSph/GradhSph.cpp:142-326
 *            GradhSph::ComputeThermalProperties     src/GradhSph/GradhSph.cpp:335-347 (+ EOS classes)
 *
 *  Work decomposition: one warp per COHORT = one KD-tree node of <= 32*T particles (a contiguous
 *  slab in tree order).  The warp walks the tree once for the cohort's box grown by
 *  kernrange*hmax*1.05^n, writes the candidate slab indices to its scratch list, then every lane
 *  owns T target particles and streams the candidates through a 32-entry shared-memory tile
 *  (coalesced SoA loads, broadcast reads).  The fixed-point iteration of ComputeH runs per lane
 *  with exactly the reference's control flow; a lane whose h outgrows the search radius makes the
 *  cohort redo the walk with a larger radius (the reference's `hmax *= 1.05` loop).
 *=================================================================================================*/
#include "device_common.cuh"

#define DENS_THREADS 256
#define DENS_WARPS (DENS_THREADS/32)

namespace {

template <int NDIM>
struct GatherVisitor {
  const TreeDev& T; const DevParams& dp; CandList& cl;
  double smin[3], smax[3];
  __device__ __forceinline__ int test(int node, bool) const
  {
    double bmin[3], bmax[3];
#pragma unroll
    for (int k = 0; k < NDIM; k++) { bmin[k] = T.bbmin[k][node]; bmax[k] = T.bbmax[k][node]; }
    return boxes_overlap<NDIM>(dp, smin, smax, bmin, bmax) ? 1 : 0;
  }
  __device__ __forceinline__ bool emit(int node, int code, bool leaf, int lane)
  {
    return cl.append_leaf(T, node, code == 1 && leaf, lane);
  }
};

/* EOS: src/Headers/EOS.h:156, src/Thermal/{Adiabatic,Isothermal,Barotropic}EOS.cpp */
__device__ __forceinline__ void thermal(const DevParams& dp, double rho, double& u, double& sound,
                                        double& press)
{
  if (dp.eos == GSPH_EOS_ISOTHERMAL) u = dp.temp0/dp.gammam1/dp.mu_bar;
  else if (dp.eos == GSPH_EOS_BAROTROPIC)
    u = dp.temp0*(1.0 + pow(rho*(1.0/dp.rho_bary), dp.gammam1))/dp.gammam1/dp.mu_bar;
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

template <int NDIM, int KERN, bool TAB, int T>
__global__ void __launch_bounds__(DENS_THREADS)
density_kernel(PartView P, TreeDev tree, DevParams dp, const double* __restrict__ gtables,
               int cohort_level, int* work_counter, int32_t* cand_all, int cand_cap, DevStatus* st,
               uint8_t* done)
{
  extern __shared__ double smem[];
  double* tab = smem;                                          /* 3 x 1000 (tabulated only) */
  double* tiles = smem + (TAB ? 3*GSPH_TABLE_RES : 0);         /* per warp: 32 x 4 doubles  */
  int* stacks = (int*)(tiles + DENS_WARPS*32*4);               /* per warp: WALK_STACK_CAP  */
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (TAB) {
    for (int i = threadIdx.x; i < 3*GSPH_TABLE_RES; i += blockDim.x) tab[i] = gtables[i];
    __syncthreads();
  }
  double* tile = tiles + warp*32*4;
  int* stack = stacks + warp*WALK_STACK_CAP;
  const int gwarp = blockIdx.x*DENS_WARPS + warp;
  int32_t* cand = cand_all + (size_t) gwarp*cand_cap;
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
    if (ccount == 0 || tree.nactive[node] == 0 || done[c]) continue;
    bool overflow = false;
    double cmin[3], cmax[3];
#pragma unroll
    for (int k = 0; k < NDIM; k++) { cmin[k] = tree.bbmin[k][node]; cmax[k] = tree.bbmax[k][node]; }
    const double hmax_cell = tree.hmax[node];

    for (int t0 = 0; t0 < ccount && !overflow; t0 += 32*T) {
      /* ---- targets of this lane -----------------------------------------------------------------*/
      int ti[T]; bool live[T];
      double x[T][3], mi[T], h0[T];
#pragma unroll
      for (int t = 0; t < T; t++) {
        const int o = t0 + t*32 + lane;
        ti[t] = cfirst + o;
        live[t] = o < ccount;
        if (live[t]) {
          const unsigned fl = (unsigned) P.i[I_FLAGS][ti[t]];
          live[t] = (fl & GSPH_FLAG_ACTIVE) && !(fl & GSPH_FLAG_DEAD);
        }
        x[t][0] = x[t][1] = x[t][2] = 0.0; mi[t] = 1.0; h0[t] = 1.0;
        if (live[t]) {
#pragma unroll
          for (int k = 0; k < NDIM; k++) x[t][k] = P.f[F_RX + k][ti[t]];
          mi[t] = P.f[F_M][ti[t]]; h0[t] = P.f[F_H][ti[t]];
        }
      }

      double hmax = hmax_cell;
      for (int attempt = 0; ; attempt++) {
        hmax = 1.05*hmax;                                      /* GradhSphTree.cpp:154 */
        /* ---- walk: candidates whose leaf overlaps the grown cohort box --------------------------*/
        CandList cl{cand, cand_cap, 0, st};
        GatherVisitor<NDIM> V{tree, dp, cl, {0, 0, 0}, {0, 0, 0}};
        const double R = dp.kernrange*hmax;
        bool crosses = false;
#pragma unroll
        for (int k = 0; k < NDIM; k++) {
          V.smin[k] = cmin[k] - R; V.smax[k] = cmax[k] + R;
          if (dp.periodic[k] && (V.smin[k] < dp.boxmin[k] || V.smax[k] > dp.boxmax[k] ||
                                 V.smax[k] - V.smin[k] > dp.boxhalf[k])) crosses = true;
        }
        if (!warp_walk(tree, stack, V, st)) { overflow = true; break; }
        const int ncand = cl.n;
        loc_cand += (unsigned long long) ncand;

        /* ---- ComputeH state per target ----------------------------------------------------------*/
        double h[T], hlo[T], hup[T], rho[T], invom[T], zeta[T];
        int iter[T]; int state[T];            /* 0 iterating, 1 converged/left loop, 2 need bigger list, 3 failed */
#pragma unroll
        for (int t = 0; t < T; t++) {
          h[t] = h0[t]; hlo[t] = 0.0; hup[t] = hmax; iter[t] = 0; state[t] = live[t] ? 0 : 1;
          rho[t] = invom[t] = zeta[t] = 0.0;
        }
        while (true) {
          bool any = false;
#pragma unroll
          for (int t = 0; t < T; t++) any |= (state[t] == 0);
          if (!__any_sync(FULLMASK, any)) break;

          double invh[T], invhsqd[T];
#pragma unroll
          for (int t = 0; t < T; t++) {
            invh[t] = 1.0/h[t]; invhsqd[t] = invh[t]*invh[t];
            if (state[t] == 0) { rho[t] = 0.0; invom[t] = 0.0; zeta[t] = 0.0; }
          }
          /* ---- neighbour sums (GradhSph.cpp:197-204) over the candidate stream -----------------*/
          for (int base = 0; base < ncand; base += 32) {
            const int jl = base + lane;
            double cx = 0.0, cy = 0.0, cz = 0.0, cm = 0.0;
            if (jl < ncand) {
              const int j = cand[jl];
              const bool dead = ((unsigned) P.i[I_FLAGS][j]) & GSPH_FLAG_DEAD;
              cx = P.f[F_RX][j];
              if (NDIM > 1) cy = P.f[F_RY][j];
              if (NDIM > 2) cz = P.f[F_RZ][j];
              cm = dead ? 0.0 : P.f[F_M][j];
            }
            __syncwarp();
            tile[lane*4 + 0] = cx; tile[lane*4 + 1] = cy; tile[lane*4 + 2] = cz; tile[lane*4 + 3] = cm;
            __syncwarp();
            const int nt = min(32, ncand - base);
            for (int jj = 0; jj < nt; jj++) {
              const double2 a = *reinterpret_cast<const double2*>(tile + jj*4);
              const double2 b = *reinterpret_cast<const double2*>(tile + jj*4 + 2);
#pragma unroll
              for (int t = 0; t < T; t++) {
                double dr[3] = {0, 0, 0};
                dr[0] = a.x - x[t][0];
                if (NDIM > 1) dr[1] = a.y - x[t][1];
                if (NDIM > 2) dr[2] = b.x - x[t][2];
                if (crosses) wrap_dr<NDIM>(dp, dr);
                const double ssqd = __dmul_rn(invhsqd[t], dot_exact<NDIM>(dr));
                if (ssqd < krs && state[t] == 0) {
                  double w0, wom, wze;
                  K.density_terms(ssqd, w0, wom, wze);
                  rho[t] += b.y*w0;
                  invom[t] += b.y*invh[t]*wom;
                  zeta[t] += b.y*wze;
                }
              }
            }
          }
          /* ---- fixed-point / bisection update (GradhSph.cpp:206-257) ----------------------------*/
#pragma unroll
          for (int t = 0; t < T; t++) {
            if (state[t] != 0) continue;
            iter[t]++;
            const double hfac = pow_ndim<NDIM>(invh[t]);
            rho[t] *= hfac; invom[t] *= hfac; zeta[t] *= invhsqd[t];
            const double hnew = dp.h_fac*root_ndim<NDIM>(mi[t]/rho[t]);
            if (rho[t] > 0.0 && h[t] > hlo[t] && fabs(h[t] - hnew)*invh[t] < dp.h_converge) { state[t] = 1; continue; }
            if (iter[t] < 30) h[t] = hnew;
            else if (iter[t] == 30) h[t] = 0.5*(hlo[t] + hup[t]);
            else if (iter[t] < 150) {
              if (rho[t] < GSPH_SMALL || h[t] > hnew) hup[t] = h[t]; else hlo[t] = h[t];
              h[t] = 0.5*(hlo[t] + hup[t]);
            }
            else { state[t] = 3; continue; }
            if (h[t] > hmax) { state[t] = 2; continue; }
            if (!(h[t] > hlo[t] && h[t] < hup[t])) state[t] = 1;
          }
        }
        bool retry = false, failed = false;
#pragma unroll
        for (int t = 0; t < T; t++) { retry |= (state[t] == 2); failed |= (state[t] == 3); }
        if (__any_sync(FULLMASK, failed)) {
          if (lane == 0) atomicExch(&st->err_converge, 1);
          return;
        }
        if (__any_sync(FULLMASK, retry)) {
          loc_retry++;
          if (attempt > 60) { if (lane == 0) atomicExch(&st->err_hsearch, 1); return; }
          continue;
        }
        /* ---- normalise and store (GradhSph.cpp:260-317) ------------------------------------------*/
#pragma unroll
        for (int t = 0; t < T; t++) {
          if (!live[t]) continue;
          const double hf = fmax(dp.h_fac*root_ndim<NDIM>(mi[t]/rho[t]), hlo[t]);
          const double ih = 1.0/hf;
          const int i = ti[t];
          P.f[F_H][i] = hf;
          P.f[F_RHO][i] = rho[t];
          P.f[F_HFAC][i] = pow_ndim<NDIM>(ih)*ih;
          P.f[F_HR2][i] = __dmul_rn(__dmul_rn(krs, hf), hf);
          P.f[F_DIVV][i] = 0.0;
          double io = 1.0, ze = 0.0;
          if (P.i[I_SINKID][i] == -1) {
            const double deriv = -invndim*hf/rho[t];                /* h_rho_deriv, Sph.h:264-268 */
            io = 1.0/(1.0 - deriv*invom[t]);
            ze = deriv*zeta[t]*io;
          }
          P.f[F_INVOMEGA][i] = io;
          P.f[F_ZETA][i] = ze;
          double u = P.f[F_U][i], cs, pr;
          thermal(dp, rho[t], u, cs, pr);
          P.f[F_U][i] = u; P.f[F_SOUND][i] = cs; P.f[F_PRESS][i] = pr;
          loc_iter += iter[t]; loc_targets++; loc_itmax = max(loc_itmax, iter[t]);
        }
        break;
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

template <int NDIM, int KERN, bool TAB>
int launch_density(gsph_handle* h)
{
  constexpr int T = 2;
  const int Lc = cohort_level_for(h->N, h->tree.ltot, 32*T);
  const size_t smem = (TAB ? 3*GSPH_TABLE_RES*sizeof(double) : 0) + DENS_WARPS*32*4*sizeof(double)
                      + DENS_WARPS*WALK_STACK_CAP*sizeof(int);
  auto kern = density_kernel<NDIM, KERN, TAB, T>;
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
  CUDA_TRY(h, cudaMemsetAsync(h->work_counter, 0, sizeof(int), h->stream));
  PartView P;
  for (int k = 0; k < NF_DOUBLE; k++) P.f[k] = h->P[k];
  for (int k = 0; k < NF_INT; k++) P.i[k] = h->PI[k];
  KernelTimer kt(h, 0);
  kern<<<blocks, DENS_THREADS, smem, h->stream>>>(P, h->tree, h->dp, h->tables, Lc, h->work_counter,
                                                  h->cand, (int) h->cand_cap_per_warp, h->status, h->done);
  h->stats.kernel_launches++;
  CUDA_TRY(h, cudaGetLastError());
  return GSPH_OK;
}

template <int NDIM>
int dispatch_density(gsph_handle* h)
{
  const bool tab = h->dp.tabulated != 0;
  if (h->dp.kernel == GSPH_KERNEL_M4)
    return tab ? launch_density<NDIM, GSPH_KERNEL_M4, true>(h) : launch_density<NDIM, GSPH_KERNEL_M4, false>(h);
  return tab ? launch_density<NDIM, GSPH_KERNEL_QUINTIC, true>(h)
             : launch_density<NDIM, GSPH_KERNEL_QUINTIC, false>(h);
}

}  // namespace

int gsph_run_density(gsph_handle* h)
{
  switch (h->dp.ndim) {
  case 1: return dispatch_density<1>(h);
  case 2: return dispatch_density<2>(h);
  default: return dispatch_density<3>(h);
  }
}
