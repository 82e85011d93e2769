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
 *  One warp per cohort (KD-tree node of <= 32*T particles).  The walk opens a node when
 *  cohort.bb overlaps node.hbox or cohort.hbox overlaps node.bb (scatter-gather), candidates are
 *  streamed through a shared-memory tile; the loading lane derives 1/h, 1/rho and
 *  P*invomega/rho^2 once per candidate.  Neighbour membership is the reference's predicate
 *      NOT (drsqd >= hrangesqd_i && drsqd >= hrangesqd_j)          NeighbourManager.h:521
 *  evaluated with the reference's rounding (no FMA contraction).
 *=================================================================================================*/
#include "device_common.cuh"

#define HYD_THREADS 128
#define HYD_WARPS (HYD_THREADS/32)

namespace {

/* tile layout per warp: geom[32][4] = x,y,z,hr2 ; then SoA fields [NTF][32] */
enum TileField { TF_VX, TF_VY, TF_VZ, TF_M, TF_INVH, TF_HFAC, TF_INVRHO, TF_PFAC, TF_SOUND, TF_U,
                 TF_ALPHA, TF_PRESS, TF_ZETA, TF_LEVEL, NTF };
#define TILE_DOUBLES (32*4 + NTF*32)

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

struct Target {
  double x[3], v[3];
  double invh, hfac, invrho, pfac, sound, u, alpha, press, zeta, hr2, m, rho, invomega;
  int level;
  /* accumulators */
  double a[3], at[3], gpot, dudt, divv;
  int levelneib, nngb;
};

/* One (target, neighbour) pair.  GRAV selects ComputeSphHydroGravForces (GradhSph.cpp:498-572),
 * otherwise ComputeSphHydroForces (:384-448); HYDRO = false gives ComputeSphGravForces (:610-645). */
template <int NDIM, bool GRAV, bool HYDRO, class KE>
__device__ __forceinline__ void pair_terms(const DevParams& dp, const KE& K, Target& t, double (&dr)[3],
                                           double drsqd, const double* __restrict__ tf, int jj)
{
  const double mj = tf[TF_M*32 + jj], invh_j = tf[TF_INVH*32 + jj], hfac_j = tf[TF_HFAC*32 + jj];
  double drmag, dvdr = 0.0;
  if (!GRAV) {
    drmag = sqrt(drsqd);
    if (drmag > 0.0) {
      const double inv = 1.0/drmag;
#pragma unroll
      for (int k = 0; k < NDIM; k++) dr[k] *= inv;
    }
  }
  else {
    drmag = sqrt(drsqd + GSPH_SMALL);
    const double inv = 1.0/drmag;
#pragma unroll
    for (int k = 0; k < NDIM; k++) dr[k] *= inv;
  }
  const double w1i = K.w1(drmag*t.invh), w1j = K.w1(drmag*invh_j);
  const double wkerni = t.hfac*w1i, wkernj = hfac_j*w1j;
  if (HYDRO) {
#pragma unroll
    for (int k = 0; k < NDIM; k++) dvdr += (tf[(TF_VX + k)*32 + jj] - t.v[k])*dr[k];
    const double invrho_j = tf[TF_INVRHO*32 + jj];
    t.divv -= mj*dvdr*wkerni;
    double paux = t.pfac*wkerni + tf[TF_PFAC*32 + jj]*wkernj;
    if (dvdr < 0.0) {
      const double winvrho = 0.25*(wkerni + wkernj)*(t.invrho + invrho_j);
      if (dp.avisc == GSPH_AVISC_MON97) {
        const double vsignal = t.sound + tf[TF_SOUND*32 + jj] - dp.beta_visc*dp.alpha_visc*dvdr;
        paux   -= dp.alpha_visc*vsignal*dvdr*winvrho;
        t.dudt -= 0.5*mj*dp.alpha_visc*vsignal*dvdr*dvdr*winvrho;
      }
      else if (dp.avisc == GSPH_AVISC_MON97MM97 || dp.avisc == GSPH_AVISC_MON97CD2010) {
        const double alpha_mean = 0.5*(t.alpha + tf[TF_ALPHA*32 + jj]);
        const double vsignal = t.sound + tf[TF_SOUND*32 + jj] - dp.beta_visc*alpha_mean*dvdr;
        paux   -= alpha_mean*vsignal*dvdr*winvrho;
        t.dudt -= 0.5*mj*alpha_mean*vsignal*dvdr*dvdr*winvrho;
      }
      if (dp.acond == GSPH_ACOND_WADSLEY2008)
        t.dudt += mj*dvdr*(tf[TF_U*32 + jj] - t.u)*(t.invrho*wkerni + invrho_j*wkernj);
      else if (dp.acond == GSPH_ACOND_PRICE2008)
        t.dudt += 0.5*mj*(t.u - tf[TF_U*32 + jj])*winvrho*(t.invrho + invrho_j)*
          sqrt(fabs(t.press - tf[TF_PRESS*32 + jj]));
    }
#pragma unroll
    for (int k = 0; k < NDIM; k++) t.a[k] += mj*dr[k]*paux;
  }
  if (GRAV) {
    const double si = drmag*t.invh, sj = drmag*invh_j;
    const double gaux = 0.5*(t.invh*t.invh*K.wgrav(si) + t.zeta*wkerni +
                             invh_j*invh_j*K.wgrav(sj) + tf[TF_ZETA*32 + jj]*wkernj);
#pragma unroll
    for (int k = 0; k < NDIM; k++) t.at[k] += mj*dr[k]*gaux;
    t.gpot += 0.5*mj*(t.invh*K.wpot(si) + invh_j*K.wpot(sj));
  }
  if (HYDRO || !GRAV) {
    const int lj = (int) tf[TF_LEVEL*32 + jj];
    t.levelneib = max(t.levelneib, lj);
  }
  t.nngb++;
}

template <int NDIM, int KERN, bool TAB, bool GRAV, bool HYDRO, int T>
__global__ void __launch_bounds__(HYD_THREADS)
hydro_kernel(PartView P, TreeDev tree, DevParams dp, const double* __restrict__ gtables,
             int cohort_level, int* work_counter, int32_t* cand_all, int cand_cap, DevStatus* st,
             uint8_t* done)
{
  extern __shared__ double smem[];
  constexpr int NTAB = TAB ? (GRAV ? 3 : 1) : 0;
  double* tab = smem;                                   /* w1 [, wgrav, wpot] */
  double* tiles = smem + NTAB*GSPH_TABLE_RES;
  int* stacks = (int*)(tiles + HYD_WARPS*TILE_DOUBLES);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (TAB) {
    for (int i = threadIdx.x; i < NTAB*GSPH_TABLE_RES; i += blockDim.x) tab[i] = gtables[3*GSPH_TABLE_RES + i];
    __syncthreads();
  }
  double* geom = tiles + warp*TILE_DOUBLES;
  double* tf = geom + 32*4;
  int* stack = stacks + warp*WALK_STACK_CAP;
  const int gwarp = blockIdx.x*HYD_WARPS + warp;
  int32_t* cand = cand_all + (size_t) gwarp*cand_cap;
  /* evaluator indexes tables as 3=w1, 4=wgrav, 5=wpot: rebase so that slot 3 is tab[0] */
  const KernelEval<KERN, NDIM, TAB> K(dp, tab - 3*GSPH_TABLE_RES);
  const int ncohort = 1 << cohort_level;
  unsigned long long loc_ngb = 0, loc_cand = 0, loc_targets = 0; int loc_ngbmax = 0;

  while (true) {
    int c = 0;
    if (lane == 0) c = atomicAdd(work_counter, 1);
    c = __shfl_sync(FULLMASK, c, 0);
    if (c >= ncohort) break;
    const int node = ncohort - 1 + c;
    const int cfirst = tree.first[node], ccount = tree.count[node];
    if (ccount == 0 || tree.nactive[node] == 0 || done[c]) continue;

    /* ---- walk ---------------------------------------------------------------------------------*/
    CandList cl{cand, cand_cap, 0, st};
    ScatterGatherVisitor<NDIM> V{tree, dp, cl, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, false};
#pragma unroll
    for (int k = 0; k < NDIM; k++) {
      V.bmin[k] = tree.bbmin[k][node]; V.bmax[k] = tree.bbmax[k][node];
      V.hmin[k] = tree.hbmin[k][node]; V.hmax[k] = tree.hbmax[k][node];
    }
    if (!warp_walk(tree, stack, V, st)) continue;
    const int ncand = cl.n;
    const bool wrap = dp.any_periodic && __any_sync(FULLMASK, V.needwrap);

    for (int t0 = 0; t0 < ccount; t0 += 32*T) {
      Target tg[T]; bool live[T]; int ti[T];
#pragma unroll
      for (int t = 0; t < T; t++) {
        const int o = t0 + t*32 + lane;
        ti[t] = cfirst + o;
        live[t] = o < ccount;
        if (live[t]) {
          const unsigned fl = (unsigned) P.i[I_FLAGS][ti[t]];
          live[t] = (fl & GSPH_FLAG_ACTIVE) && !(fl & GSPH_FLAG_DEAD);
        }
        Target& g = tg[t];
        g.x[0] = g.x[1] = g.x[2] = GSPH_BIG; g.v[0] = g.v[1] = g.v[2] = 0.0;
        g.invh = g.hfac = g.invrho = g.pfac = g.sound = g.u = g.alpha = g.press = g.zeta = 0.0;
        g.hr2 = -1.0; g.m = 0.0; g.rho = 1.0; g.invomega = 1.0; g.level = 0;
        if (live[t]) {
          const int i = ti[t];
#pragma unroll
          for (int k = 0; k < NDIM; k++) { g.x[k] = P.f[F_RX + k][i]; g.v[k] = P.f[F_VX + k][i]; }
#pragma unroll
          for (int k = NDIM; k < 3; k++) g.x[k] = 0.0;
          const double hi = P.f[F_H][i];
          g.rho = P.f[F_RHO][i]; g.press = P.f[F_PRESS][i]; g.invomega = P.f[F_INVOMEGA][i];
          g.invh = 1.0/hi; g.invrho = 1.0/g.rho; g.hfac = P.f[F_HFAC][i];
          g.pfac = (g.press*g.invomega)/(g.rho*g.rho);
          g.sound = P.f[F_SOUND][i]; g.u = P.f[F_U][i]; g.alpha = P.f[F_ALPHA][i];
          g.zeta = P.f[F_ZETA][i]; g.hr2 = P.f[F_HR2][i]; g.m = P.f[F_M][i];
          g.level = P.i[I_LEVEL][i];
        }
        g.a[0] = g.a[1] = g.a[2] = 0.0; g.at[0] = g.at[1] = g.at[2] = 0.0;
        g.gpot = 0.0; g.dudt = 0.0; g.divv = 0.0; g.levelneib = 0; g.nngb = 0;
        if (GRAV && live[t]) g.gpot = (g.m*g.invh)*K.wpot(0.0);          /* GradhSphTree.cpp:512 */
      }

      for (int base = 0; base < ncand; base += 32) {
        const int jl = base + lane;
        __syncwarp();
        if (jl < ncand) {
          const int j = cand[jl];
          const bool dead = ((unsigned) P.i[I_FLAGS][j]) & GSPH_FLAG_DEAD;
          const double hj = P.f[F_H][j], rhoj = P.f[F_RHO][j];
          geom[lane*4 + 0] = dead ? GSPH_BIG : P.f[F_RX][j];
          geom[lane*4 + 1] = NDIM > 1 ? P.f[F_RY][j] : 0.0;
          geom[lane*4 + 2] = NDIM > 2 ? P.f[F_RZ][j] : 0.0;
          geom[lane*4 + 3] = dead ? -1.0 : P.f[F_HR2][j];
#pragma unroll
          for (int k = 0; k < NDIM; k++) tf[(TF_VX + k)*32 + lane] = P.f[F_VX + k][j];
          tf[TF_M*32 + lane] = P.f[F_M][j];
          tf[TF_INVH*32 + lane] = 1.0/hj;
          tf[TF_HFAC*32 + lane] = P.f[F_HFAC][j];
          tf[TF_INVRHO*32 + lane] = 1.0/rhoj;
          const double pj = P.f[F_PRESS][j];
          tf[TF_PFAC*32 + lane] = (pj*P.f[F_INVOMEGA][j])/(rhoj*rhoj);
          tf[TF_SOUND*32 + lane] = P.f[F_SOUND][j];
          tf[TF_U*32 + lane] = P.f[F_U][j];
          tf[TF_ALPHA*32 + lane] = P.f[F_ALPHA][j];
          tf[TF_PRESS*32 + lane] = pj;
          tf[TF_ZETA*32 + lane] = P.f[F_ZETA][j];
          tf[TF_LEVEL*32 + lane] = (double) P.i[I_LEVEL][j];
        }
        __syncwarp();
        const int nt = min(32, ncand - base);
        for (int jj = 0; jj < nt; jj++) {
          const double2 ga = *reinterpret_cast<const double2*>(geom + jj*4);
          const double2 gb = *reinterpret_cast<const double2*>(geom + jj*4 + 2);
#pragma unroll
          for (int t = 0; t < T; t++) {
            double dr[3] = {0, 0, 0};
            dr[0] = ga.x - tg[t].x[0];
            if (NDIM > 1) dr[1] = ga.y - tg[t].x[1];
            if (NDIM > 2) dr[2] = gb.x - tg[t].x[2];
            if (wrap) wrap_dr<NDIM>(dp, dr);
            const double drsqd = dot_exact<NDIM>(dr);
            if (drsqd < tg[t].hr2 || drsqd < gb.y)
              pair_terms<NDIM, GRAV, HYDRO>(dp, K, tg[t], dr, drsqd, tf, jj);
          }
        }
      }

      /* ---- end-of-loop terms and write-back (GradhSph.cpp:451-457 / :576-582;
       *      GradhSphTree.cpp:396-406 / :610-620) ---------------------------------------------------*/
#pragma unroll
      for (int t = 0; t < T; t++) {
        if (!live[t]) continue;
        Target& g = tg[t];
        const int i = ti[t];
        if (HYDRO) {
          g.divv *= g.invrho;
          const double press = dp.gammam1*g.rho*g.u;                    /* eos->Pressure(parti) */
          g.dudt -= press*g.divv*g.invrho*g.invomega;
#pragma unroll
          for (int k = 0; k < NDIM; k++) P.f[F_AX + k][i] += g.a[k];
          P.f[F_DUDT][i] += g.dudt;
          P.f[F_DIVV][i] += g.divv;
          if (!GRAV && dp.tdavisc == GSPH_TDAVISC_MM97) {
            const double hi_inv = g.invh;
            P.f[F_DALPHADT][i] += 0.1*g.sound*(dp.alpha_visc_min - g.alpha)*hi_inv +
              fmax(-g.divv, 0.0)*(dp.alpha_visc - g.alpha);
          }
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
    if (lane == 0) done[c] = 1;
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

/* levelneib scatter half (GradhSph.cpp:446, GradhSphTree.cpp:376-382): only needed when some
 * particles are inactive -- with every particle active the neighbour relation is symmetric and the
 * gather above already produces the maximum.  Also serves the parity API (neighbour counts / lists). */
template <int NDIM, int MODE /*0 count, 1 fill, 2 levelneib scatter*/>
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
  int* stack = stacks + warp*WALK_STACK_CAP;
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
    if (MODE == 2 && tree.nactive[node] == 0) continue;
    CandList cl{cand, cand_cap, 0, st};
    ScatterGatherVisitor<NDIM> V{tree, dp, cl, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, false};
#pragma unroll
    for (int k = 0; k < NDIM; k++) {
      V.bmin[k] = tree.bbmin[k][node]; V.bmax[k] = tree.bbmax[k][node];
      V.hmin[k] = tree.hbmin[k][node]; V.hmax[k] = tree.hbmax[k][node];
    }
    if (!warp_walk(tree, stack, V, st)) continue;
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
        if (MODE == 2 && !(fl & GSPH_FLAG_ACTIVE)) live = false;
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
            if (MODE == 2) atomicMax(&P.i[I_LEVELNEIB][jidx[warp][jj]], lev);
            n++;
          }
        }
      }
      if (MODE == 0 && live) counts[perm[i]] = n;
    }
  }
}

template <int NDIM, int KERN, bool TAB, bool GRAV, bool HYDRO>
int launch_hydro(gsph_handle* h)
{
  constexpr int T = 1;
  const int Lc = cohort_level_for(h->N, h->tree.ltot, 32*T);
  constexpr int NTAB = TAB ? (GRAV ? 3 : 1) : 0;
  const size_t smem = NTAB*GSPH_TABLE_RES*sizeof(double) + HYD_WARPS*TILE_DOUBLES*sizeof(double)
                      + HYD_WARPS*WALK_STACK_CAP*sizeof(int);
  auto kern = hydro_kernel<NDIM, KERN, TAB, GRAV, HYDRO, T>;
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
  CUDA_TRY(h, cudaMemsetAsync(h->work_counter, 0, sizeof(int), h->stream));
  PartView P;
  for (int k = 0; k < NF_DOUBLE; k++) P.f[k] = h->P[k];
  for (int k = 0; k < NF_INT; k++) P.i[k] = h->PI[k];
  KernelTimer kt(h, 1);
  kern<<<blocks, HYD_THREADS, smem, h->stream>>>(P, h->tree, h->dp, h->tables, Lc, h->work_counter,
                                                 h->cand, (int) h->cand_cap_per_warp, h->status, h->done);
  h->stats.kernel_launches++;
  CUDA_TRY(h, cudaGetLastError());
  return GSPH_OK;
}

template <int NDIM, int MODE>
int launch_neib(gsph_handle* h, int32_t* counts, const int64_t* offsets, int32_t* idx)
{
  const int Lc = cohort_level_for(h->N, h->tree.ltot, 32);
  const size_t smem = HYD_WARPS*32*4*sizeof(double) + HYD_WARPS*WALK_STACK_CAP*sizeof(int);
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
  else
    rc = tab ? dispatch_mode<NDIM, GSPH_KERNEL_QUINTIC, true>(h, with_gravity)
             : dispatch_mode<NDIM, GSPH_KERNEL_QUINTIC, false>(h, with_gravity);
  if (rc) return rc;
  if (!h->all_active && !h->slice_mode && (h->dp.hydro_forces || !with_gravity))
    rc = launch_neib<NDIM, 2>(h, nullptr, nullptr, nullptr);
  return rc;
}

}  // namespace

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
  const bool fill = idx != nullptr;
  switch (h->dp.ndim) {
  case 1: return fill ? launch_neib<1, 1>(h, counts, offsets, idx) : launch_neib<1, 0>(h, counts, offsets, idx);
  case 2: return fill ? launch_neib<2, 1>(h, counts, offsets, idx) : launch_neib<2, 0>(h, counts, offsets, idx);
  default: return fill ? launch_neib<3, 1>(h, counts, offsets, idx) : launch_neib<3, 0>(h, counts, offsets, idx);
  }
}
