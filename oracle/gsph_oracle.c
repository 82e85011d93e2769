/*=================================================================================================
 *  oracle/gsph_oracle.c -- TEST INFRASTRUCTURE ONLY (see gsph_oracle.h).
 *
 *  Plain-C restatement of the reference (GANDALF v0.4.0) grad-h SPH neighbour path.
 *  Every function cites the reference file:line it follows (paths relative to /root/reference).
 *  Compiled with -ffp-contract=off: the reference's x86-64 build has no FMA contraction, and the
 *  neighbour predicate must reproduce its rounding (src/Headers/InlineFuncs.h:46-54).
 *
 *  PARITY PINNED against golden vectors from the unmodified reference (tests/golden/).
 *=================================================================================================*/
#include "gsph_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#define SMALL_NUMBER 1.0e-20          /* src/Headers/Constants.h:73 */
#define BIG_NUMBER   9.9e20           /* src/Headers/Constants.h:72 */
#define INVPI        0.31830988618379 /* src/Headers/Constants.h:63 */
#define TWOTHIRDS    0.66666666666666666666666
#define ONETHIRD     0.33333333333333333333333
#define TABLE_RES    1000

/*------------------------------------------------------------------------------------------------
 * Smoothing kernels.  M4: src/Headers/SmoothingKernel.h:131-240, norms src/Hydrodynamics/
 * M4Kernel.cpp:39-53.  Quintic: SmoothingKernel.h:281-408, norms QuinticKernel.cpp:39-53.
 *-----------------------------------------------------------------------------------------------*/
static double kern_range(const orc_params* p) { return p->kernel == 0 ? 2.0 : 3.0; }   /* m4 | quintic, gaussian */

static double kern_norm(const orc_params* p)
{
  if (p->kernel == 0) {
    if (p->ndim == 1) return TWOTHIRDS;
    if (p->ndim == 2) return INVPI*(10.0/7.0);
    return INVPI;
  }
  if (p->kernel == 2) {                          /* GaussianKernel.cpp:39-58 */
    if (p->ndim == 1) return sqrt(INVPI);
    if (p->ndim == 2) return INVPI;
    return INVPI*sqrt(INVPI);
  }
  if (p->ndim == 1) return 1.0/120.0;
  if (p->ndim == 2) return INVPI*(7.0/478.0);
  return INVPI*(1/120.);
}

static double m4_eval(int which, double s, double kernnorm, int ndim)
{
  switch (which) {
  case 0:
    if (s < 1.0) return kernnorm*(1.0 - 1.5*s*s + 0.75*s*s*s);
    else if (s < 2.0) return 0.25*kernnorm*pow(2.0 - s, 3);
    else return 0.0;
  case 1:
    if (s < 1.0) return kernnorm*(-3.0*s + 2.25*s*s);
    else if (s < 2.0) return -0.75*kernnorm*(2.0 - s)*(2.0 - s);
    else return 0.0;
  case 2:
    if (s < 1.0)
      return kernnorm*(-(ndim) + 1.5*((ndim) + 2.0)*s*s - 0.75*((ndim) + 3.0)*pow(s,3));
    else if (s < 2.0)
      return kernnorm*(-2.0*(ndim) + 3.0*((ndim) + 1.0)*s - 1.50*((ndim) + 2.0)*s*s +
                       0.25*((ndim) + 3.0)*pow(s,3));
    else return 0.0;
  case 3:
    if (s < 1.0) return 1.4 - 2.0*s*s + 1.5*pow(s,4) - 0.6*pow(s,5);
    else if (s < 2.0) return 1.6 - 4.0*s*s + 4.0*pow(s,3) - 1.5*pow(s,4) + 0.2*pow(s,5);
    else return 0.0;
  case 4:
    if (s < 1.0) return 1.333333333333333333333*s - 1.2*pow(s,3) + 0.5*pow(s,4);
    else if (s < 2.0)
      return 2.6666666666666666667*s - 3.0*s*s + 1.2*pow(s,3) -
        0.166666666666666666667*pow(s,4) - 0.06666666666666666667/(s*s);
    else return 1.0/(s*s);
  default:
    if (s < 1.0)
      return 1.4 - 0.666666666666666666666666*s*s + 0.3*pow(s,4) - 0.1*pow(s,5);
    else if (s < 2.0)
      return -1.0/(15.0*s) + 1.6 - 1.33333333333333333333333333*s*s + pow(s,3) - 0.3*pow(s,4) +
        (1.0/30.0)*pow(s,5);
    else return 1.0/s;
  }
}

static double quintic_eval(int which, double s, double kernnorm, int ndim)
{
  switch (which) {
  case 0:
    if (s < 1.0) return kernnorm*(66.0 - 60.0*s*s + 30.0*pow(s,4) - 10.0*pow(s,5));
    else if (s < 2.0)
      return kernnorm*(51.0 + 75.0*s - 210.0*s*s + 150.0*pow(s,3) - 45.0*pow(s,4) + 5.0*pow(s,5));
    else if (s < 3.0)
      return kernnorm*(243.0 - 405*s + 270.0*s*s - 90.0*pow(s,3) + 15.0*pow(s,4) - pow(s,5));
    else return 0.0;
  case 1:
    if (s < 1.0) return kernnorm*(-120.0*s + 120.0*pow(s,3) - 50.0*pow(s,4));
    else if (s < 2.0)
      return kernnorm*(75.0 - 420.0*s + 450.0*s*s - 180.0*pow(s,3) + 25.0*pow(s,4));
    else if (s < 3.0)
      return kernnorm*(-405.0 + 540.0*s - 270.0*s*s + 60.0*pow(s,3) - 5.0*pow(s,4));
    else return 0.0;
  case 2:
    if (s < 1.0)
      return kernnorm*(-66.0*(ndim) + 60.0*((ndim) + 2.0)*s*s - 30.0*((ndim) + 4.0)*pow(s,4) +
                       10.0*((ndim) + 5.0)*pow(s,5));
    else if (s < 2.0)
      return kernnorm*(-51.0*(ndim) - 75.0*((ndim) + 1.0)*s + 210.0*((ndim) + 2.0)*s*s -
                       150.0*((ndim) + 3.0)*pow(s,3) + 45.0*((ndim) + 4.0)*pow(s,4) -
                       5.0*((ndim) + 5.0)*pow(s,5));
    else if (s < 3.0)
      return kernnorm*(-243.0*(ndim) + 405.0*((ndim) + 1.0)*s - 270.0*((ndim) + 2.0)*s*s +
                       90.0*((ndim) + 3.0)*pow(s,3) - 15.0*((ndim) + 4.0)*pow(s,4) +
                       ((ndim) + 5.0)*pow(s,5));
    else return 0.0;
  case 3:
    if (s < 1.0)
      return 33.0*s*s - 15.0*pow(s,4) + 5.0*pow(s,6) - 1.42857142857*pow(s,7) - 34.14285714;
    else if (s < 2.0)
      return 25.5*s*s + 25.0*pow(s,3) - 52.5*pow(s,4) + 30.0*pow(s,5) - 7.5*pow(s,6) +
        0.7142857143*pow(s,7) - 33.785714286;
    else if (s < 3.0)
      return 121.5*s*s - 135.0*pow(s,3) + 67.5*pow(s,4) - 18.0*pow(s,5) + 2.5*pow(s,6) -
        0.142857143*pow(s,7) - 52.07142857;
    else return 0.0;
  case 4:
    if (s < 1.0)
      return (12.0/359.0)*(22.0*s - 12.0*pow(s,3) + (30.0/7.0)*pow(s,5) - (5.0/4.0)*pow(s,6));
    else if (s < 2.0)
      return (12.0/359.0)*(17.0*s + (75.0/4.0)*s*s - 42.0*pow(s,3) + 25.0*pow(s,4) -
                           (45.0/7.0)*pow(s,5) + (5.0/8.0)*pow(s,6) + (5.0/56.0)/(s*s));
    else if (s < 3.0)
      return (12.0/359.0)*(81.0*s - (405.0/4.0)*pow(s,2) + 54.0*pow(s,3) - 15.0*pow(s,4) +
                           (15.0/7.0)*pow(s,5) - (1.0/8.0)*pow(s,6) - (507.0/56.0)/(s*s));
    else return 1.0/(s*s);
  default:
    if (s < 1.0)
      return (12.0/359.0)*(-11.0*s*s + 3.0*pow(s,4) - (5.0/7.0)*pow(s,6) + (5.0/28.0)*pow(s,7) +
                           (478.0/14.0));
    else if (s < 2.0)
      return (12.0/359.0)*(-(17.0/2.0)*s*s - (25.0/4.0)*pow(s,3) + (21.0/2.0)*pow(s,4) -
                           5.0*pow(s,5) + (15.0/14.0)*pow(s,6) - (5.0/56.0)*pow(s,7) +
                           (473.0/14.0) + (5.0/56.0)/s);
    else if (s < 3.0)
      return (12.0/359.0)*(-(81.0/2.0)*s*s + (135.0/4.0)*pow(s,3) - (27.0/2.0)*pow(s,4) +
                           3.0*pow(s,5) - (5.0/14.0)*pow(s,6) + (1.0/56.0)*pow(s,7) +
                           (729.0/14.0) - (507.0/56.0)/s);
    else return 1.0/s;
  }
}

/* GaussianKernel, src/Headers/SmoothingKernel.h:447-521 (wzeta = wgrav = wpot = 0) */
static double gaussian_eval(int which, double s, double kn, int nd)
{
  if (!(s < 3.0)) return 0.0;
  switch (which) {
  case 0: return kn*exp(-s*s);
  case 1: return -2.0*kn*s*exp(-s*s);
  case 2: return kn*(2.0*s*exp(-s*s) - nd*exp(-s*s));
  default: return 0.0;
  }
}

static double analytic(const orc_params* p, int which, double s)
{
  if (p->kernel == 2) return gaussian_eval(which, s, kern_norm(p), p->ndim);
  return p->kernel == 0 ? m4_eval(which, s, kern_norm(p), p->ndim)
                        : quintic_eval(which, s, kern_norm(p), p->ndim);
}

/* Tabulated kernel: tables built as src/Headers/SmoothingKernel.h:585-605 (initializeTable /
 * initializeTableSqd), looked up as :619-661 (truncating index, no interpolation).             */
typedef struct { int ready; int ndim, kernel; double t[6][TABLE_RES]; } orc_tables;
static orc_tables g_tab[9];   /* cache keyed by (kernel, ndim): 3 kernels x 3 dimensions */

static const orc_tables* get_tables(const orc_params* p)
{
  orc_tables* T = &g_tab[p->kernel*3 + (p->ndim - 1)];
  if (T->ready) return T;
  #pragma omp critical(orc_tables_init)
  {
    if (!T->ready) {
      const double kr = kern_range(p), krs = kr*kr;
      const double step = kr/TABLE_RES, step2 = krs/TABLE_RES;
      for (int i = 0; i < TABLE_RES; i++) {
        T->t[0][i] = analytic(p, 0, sqrt(step2*i));    /* w0_s2     */
        T->t[1][i] = analytic(p, 2, sqrt(step2*i));    /* womega_s2 */
        T->t[2][i] = analytic(p, 3, sqrt(step2*i));    /* wzeta_s2  */
        T->t[3][i] = analytic(p, 1, step*i);           /* w1        */
        T->t[4][i] = analytic(p, 4, step*i);           /* wgrav     */
        T->t[5][i] = analytic(p, 5, step*i);           /* wpot      */
      }
      T->ndim = p->ndim; T->kernel = p->kernel;
      T->ready = 1;
    }
  }
  return T;
}

void orc_kernel_table(const orc_params* p, int which, double* out)
{
  const orc_tables* T = get_tables(p);
  memcpy(out, T->t[which], sizeof(double)*TABLE_RES);
}

double orc_kernel(const orc_params* p, int which, double s)
{
  if (!p->tabulated) return analytic(p, which, s);
  const orc_tables* T = get_tables(p);
  const double kr = kern_range(p);
  const double resinvkernrange = TABLE_RES/kr;
  int ti;
  switch (which) {                         /* only the functions the hot path evaluates */
  case 1: ti = 3; break;
  case 4: ti = 4; break;
  case 5: ti = 5; break;
  default: {
      /* w0/womega/wzeta(s) tables exist in the reference too but the path only uses the s2 forms;
         build on demand from the analytic form with the same truncation */
      if (s >= kr) return 0.0;
      int idx = (int)(s*resinvkernrange);
      return analytic(p, which, (kr/TABLE_RES)*idx);
    }
  }
  if (s >= kr) {
    if (which == 4) return 1.0/(s*s);      /* GravTableLookup    SmoothingKernel.h:645-650 */
    if (which == 5) return 1.0/s;          /* GravPotTableLookup SmoothingKernel.h:657-662 */
    return 0.0;
  }
  int idx = (int)(s*resinvkernrange);
  return T->t[ti][idx];
}

double orc_kernel_s2(const orc_params* p, int which, double s2)
{
  static const int map[3] = {0, 2, 3};
  if (!p->tabulated) return analytic(p, map[which], sqrt(s2));   /* SmoothingKernel.h:77-79 */
  const orc_tables* T = get_tables(p);
  const double krs = kern_range(p)*kern_range(p);
  if (s2 >= krs) return 0.0;
  const double resinvkernrangesqd = TABLE_RES/krs;
  int idx = (int)(s2*resinvkernrangesqd);
  return T->t[which][idx];
}

/*------------------------------------------------------------------------------------------------
 * geometry helpers
 *-----------------------------------------------------------------------------------------------*/
static inline double dot(const double* a, const double* b, int ndim)
{
  /* src/Headers/InlineFuncs.h:46-54: left-to-right, no FMA */
  if (ndim == 1) return a[0]*b[0];
  if (ndim == 2) return a[0]*b[0] + a[1]*b[1];
  return a[0]*b[0] + a[1]*b[1] + a[2]*b[2];
}

/* NearestPeriodicVector / ApplyPeriodicDistanceCorrection, src/Headers/GhostNeighbours.hpp:119-166 */
static inline void wrap(const orc_params* p, double* dr)
{
  for (int k = 0; k < p->ndim; k++) {
    if (p->periodic[k]) {
      const double size = p->boxmax[k] - p->boxmin[k];
      const double half = 0.5*size;
      if (dr[k] > half) dr[k] -= size;
      else if (dr[k] < -half) dr[k] += size;
    }
  }
}

/*------------------------------------------------------------------------------------------------
 * Neighbour lists: NeighbourManager::TrimNeighbourLists, src/Headers/NeighbourManager.h:503-542
 *   draux = r_j - r_i (+ periodic correction); drsqd = DotProduct;
 *   NOT a neighbour iff (drsqd >= hrangesqd_i && drsqd >= hrangesqd_j).   Self stays in the list.
 *-----------------------------------------------------------------------------------------------*/
static inline int is_neighbour(const orc_params* p, const double* ri, const double* rj,
                               double hr2i, double hr2j, double* dr_out, double* drsqd_out)
{
  double dr[3];
  for (int k = 0; k < p->ndim; k++) dr[k] = rj[k] - ri[k];
  wrap(p, dr);
  const double drsqd = dot(dr, dr, p->ndim);
  if (dr_out) for (int k = 0; k < p->ndim; k++) dr_out[k] = dr[k];
  if (drsqd_out) *drsqd_out = drsqd;
  return !(drsqd >= hr2i && drsqd >= hr2j);
}

int64_t orc_neighbour_lists(const orc_params* p, int N, const double* r, const double* hrangesqd,
                            int64_t* offsets, int32_t* idx)
{
  const int nd = p->ndim;
  const int Nt = p->ntarget > 0 ? p->ntarget : N;     /* targets = the first Nt particles (the rest: mirror images) */
  if (!idx) {
    #pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < Nt; i++) {
      int64_t c = 0;
      for (int j = 0; j < N; j++)
        c += is_neighbour(p, r + (size_t)i*nd, r + (size_t)j*nd, hrangesqd[i], hrangesqd[j], 0, 0);
      offsets[i + 1] = c;
    }
    offsets[0] = 0;
    for (int i = 0; i < Nt; i++) offsets[i + 1] += offsets[i];
    return offsets[Nt];
  }
  #pragma omp parallel for schedule(dynamic, 64)
  for (int i = 0; i < Nt; i++) {
    int64_t o = offsets[i];
    for (int j = 0; j < N; j++)
      if (is_neighbour(p, r + (size_t)i*nd, r + (size_t)j*nd, hrangesqd[i], hrangesqd[j], 0, 0))
        idx[o++] = j;
  }
  return offsets[Nt];
}

/*------------------------------------------------------------------------------------------------
 * EOS: src/Headers/EOS.h:156 (Pressure), src/Thermal/AdiabaticEOS.cpp:69-84,
 * IsothermalEOS.cpp:76-91, BarotropicEOS.cpp:76-91; GradhSph::ComputeThermalProperties
 * src/GradhSph/GradhSph.cpp:335-347 (u, then sound, then pressure).
 *-----------------------------------------------------------------------------------------------*/
static void thermal(const orc_params* p, double rho, double* u, double* sound, double* pressure)
{
  const double gamma = p->gamma_eos, gammam1 = p->gamma_eos - 1.0;
  if (p->eos == 1) *u = p->temp0/gammam1/p->mu_bar;
  else if (p->eos == 2) *u = p->temp0*(1.0 + pow(rho*(1.0/p->rho_bary), gammam1))/gammam1/p->mu_bar;
  else if (p->eos == 3)                      /* Barotropic2, src/Thermal/Barotropic2EOS.cpp:78-83 */
    *u = rho < p->rho_bary ? p->temp0/gammam1/p->mu_bar
                           : p->temp0*pow(rho*(1.0/p->rho_bary), gammam1)/gammam1/p->mu_bar;
  if (p->eos == 0) *sound = sqrt(gamma*gammam1*(*u));
  else *sound = sqrt(gammam1*(*u));
  *pressure = gammam1*rho*(*u);
}

/*------------------------------------------------------------------------------------------------
 * GradhSph::ComputeH, src/GradhSph/GradhSph.cpp:142-326 for one particle, given the gathered
 * neighbour list (positions already periodic-corrected relative to i, masses).
 * Returns 1 ok, 0 "h > hmax: need bigger list" (:255), -2 not converged (:249-250).
 *-----------------------------------------------------------------------------------------------*/
typedef struct { double h, rho, invomega, zeta, hfactor, hrangesqd; int niter; } hres;

static int compute_h(const orc_params* p, const double* ri, double mi, double h0, double hmax, int Nneib,
                     const double* drsqd, const double* mj, hres* out)
{
  const int ndim = p->ndim;
  const double invndim = 1.0/ndim;
  const double krs = kern_range(p)*kern_range(p);
  int iteration = 0;
  const int iteration_max = 30;
  double h_lower_bound = 0.0, h_upper_bound = hmax;
  double h = h0, rho = 0, invomega = 0, zeta = 0, hfactor = 0, invh, invhsqd;

  do {
    iteration++;
    invh     = 1.0/h;
    rho      = 0.0; invomega = 0.0; zeta = 0.0;
    hfactor  = pow(invh, ndim);
    invhsqd  = invh*invh;
    for (int j = 0; j < Nneib; j++) {
      const double ssqd = invhsqd*drsqd[j];
      rho      += mj[j]*orc_kernel_s2(p, 0, ssqd);
      invomega += mj[j]*invh*orc_kernel_s2(p, 1, ssqd);
      zeta     += mj[j]*orc_kernel_s2(p, 2, ssqd);
    }
    rho      *= hfactor;
    invomega *= hfactor;
    zeta     *= invhsqd;

    const double h_new = p->h_fac*pow(mi/rho, invndim);         /* h_rho_func Sph.h:259-263 */
    if (rho > 0.0 && h > h_lower_bound && fabs(h - h_new)*invh < p->h_converge) break;

    if (iteration < iteration_max) h = h_new;
    else if (iteration == iteration_max) h = 0.5*(h_lower_bound + h_upper_bound);
    else if (iteration < 5*iteration_max) {
      if (rho < SMALL_NUMBER || h > h_new) h_upper_bound = h;
      else h_lower_bound = h;
      h = 0.5*(h_lower_bound + h_upper_bound);
    }
    else return -2;

    if (h > hmax) return 0;
  } while (h > h_lower_bound && h < h_upper_bound);

  h = fmax(p->h_fac*pow(mi/rho, invndim), h_lower_bound);
  invh = 1/h;
  out->h = h;
  out->rho = rho;
  out->hfactor = pow(invh, ndim + 1);
  out->hrangesqd = krs*h*h;
  /* sinkid == -1 branch (:284-309); h_rho_deriv = -invndim*h/rho (Sph.h:264-268) */
  const double deriv = -invndim*h/rho;
  invomega = 1.0 - deriv*invomega;
  invomega = 1.0/invomega;
  out->invomega = invomega;
  if (p->star_zeta == 1 && p->nstar > 0) {          /* GradhSph.cpp:288-306 */
    for (int j = 0; j < p->nstar; j++) {
      const double* s = p->stars + (size_t) j*(ndim + 2);
      double dr[3];
      if (p->nbody_softening == 1) invhsqd = pow(2.0/(h + s[ndim + 1]), 2);
      else invhsqd = 4.0*invh*invh;
      for (int k = 0; k < ndim; k++) dr[k] = s[k] - ri[k];
      const double ssqd = dot(dr, dr, ndim)*invhsqd;
      zeta += s[ndim]*invhsqd*orc_kernel_s2(p, 2, ssqd);
    }
  }
  out->zeta = deriv*zeta*invomega;
  out->niter = iteration;
  return 1;
}

/* Driver: GradhSphTree::UpdateAllSphProperties, src/GradhSph/GradhSphTree.cpp:144-240.
 * The reference gathers per leaf cell with the cell's hmax*1.05^k; because every kernel has
 * compact support and the fixed-point iterates do not depend on the list (only on it being
 * complete), the converged result is the same for any complete list -- here the list is built per
 * particle with hmax = h_i*1.05^k.  Gather filter: drsqd + small_number <= kernrangesqd*hmax^2
 * (:212-215).  Summation order = ascending particle index.                                        */
int orc_density(const orc_params* p, int N, const double* r, const double* m, double* h,
                double* u, double* rho, double* invomega, double* zeta, double* hfactor,
                double* hrangesqd, double* sound, double* pressure, int32_t* niter)
{
  const int nd = p->ndim;
  const double krs = kern_range(p)*kern_range(p);
  const int Nt = p->ntarget > 0 ? p->ntarget : N;     /* targets = the first Nt particles (the rest: mirror images) */
  int fail = 0;
  double* hnew = (double*) malloc(sizeof(double)*N);

  #pragma omp parallel
  {
    int cap = 256;
    double* d2 = (double*) malloc(sizeof(double)*cap);
    double* mj = (double*) malloc(sizeof(double)*cap);
    #pragma omp for schedule(dynamic, 32)
    for (int i = 0; i < Nt; i++) {
      double hmax = h[i];
      hres res; int ok;
      do {
        hmax = 1.05*hmax;
        const double hrangesqd_g = krs*hmax*hmax;
        int Nneib = 0;
        for (int j = 0; j < N; j++) {
          double dr[3];
          for (int k = 0; k < nd; k++) dr[k] = r[(size_t)j*nd + k] - r[(size_t)i*nd + k];
          wrap(p, dr);
          const double drsqd = dot(dr, dr, nd);
          if (drsqd + SMALL_NUMBER <= hrangesqd_g) {
            if (Nneib == cap) {
              cap *= 2;
              d2 = (double*) realloc(d2, sizeof(double)*cap);
              mj = (double*) realloc(mj, sizeof(double)*cap);
            }
            d2[Nneib] = drsqd; mj[Nneib] = m[j]; Nneib++;
          }
        }
        ok = compute_h(p, r + (size_t) i*nd, m[i], h[i], hmax, Nneib, d2, mj, &res);
      } while (ok == 0);
      if (ok < 0) { fail = 1; continue; }
      hnew[i] = res.h; rho[i] = res.rho; invomega[i] = res.invomega; zeta[i] = res.zeta;
      hfactor[i] = res.hfactor; hrangesqd[i] = res.hrangesqd;
      if (niter) niter[i] = res.niter;
      thermal(p, rho[i], &u[i], &sound[i], &pressure[i]);
    }
    free(d2); free(mj);
  }
  if (!fail) memcpy(h, hnew, sizeof(double)*Nt);
  free(hnew);
  return fail ? -1 : 0;
}

/*------------------------------------------------------------------------------------------------
 * pair terms shared by ComputeSphHydroForces (GradhSph.cpp:384-448) and
 * ComputeSphHydroGravForces (:498-572).  `grav` selects the second form (dr normalised with
 * sqrt(d2 + small_number), gravity terms, sign quirk of dalphadt handled by the callers).
 *-----------------------------------------------------------------------------------------------*/
typedef struct {
  const double *r, *v, *m, *h, *rho, *pressure, *invomega, *zeta, *sound, *u, *alpha, *hfactor,
               *hrangesqd;
  const int32_t* level;
} pview;

typedef struct { double a[3], atree[3], gpot, dudt, div_v; int levelneib; } acc_t;

static void hydro_pair(const orc_params* p, const pview* P, int i, int j, const double* dr_in,
                       double drsqd, int grav, acc_t* A)
{
  const int nd = p->ndim;
  double draux[3], dv[3];
  const double invh_i = 1/P->h[i], invrho_i = 1/P->rho[i];
  const double invh_j = 1/P->h[j], invrho_j = 1/P->rho[j];
  double drmag, dvdr;
  for (int k = 0; k < nd; k++) draux[k] = dr_in[k];
  if (!grav) {
    drmag = sqrt(drsqd);
    if (drmag > 0) for (int k = 0; k < nd; k++) draux[k] /= drmag;
    dvdr = dot(P->v + (size_t)j*nd, draux, nd);
    dvdr -= dot(P->v + (size_t)i*nd, draux, nd);
  }
  else {
    for (int k = 0; k < nd; k++) dv[k] = P->v[(size_t)j*nd + k] - P->v[(size_t)i*nd + k];
    drmag = sqrt(drsqd + SMALL_NUMBER);
    const double invdrmag = 1.0/drmag;
    for (int k = 0; k < nd; k++) draux[k] *= invdrmag;
    dvdr = dot(dv, draux, nd);
  }
  const double wkerni = P->hfactor[i]*orc_kernel(p, 1, drmag*invh_i);
  const double wkernj = P->hfactor[j]*orc_kernel(p, 1, drmag*invh_j);

  if (!grav) A->div_v -= P->m[j]*dvdr*wkerni;

  double paux = ((P->pressure[i]*P->invomega[i])/(P->rho[i]*P->rho[i]))*wkerni +
                ((P->pressure[j]*P->invomega[j])/(P->rho[j]*P->rho[j]))*wkernj;

  if (dvdr < 0.0) {
    const double winvrho = 0.25*(wkerni + wkernj)*(invrho_i + invrho_j);
    if (p->avisc == 1) {
      const double vsignal = P->sound[i] + P->sound[j] - p->beta_visc*p->alpha_visc*dvdr;
      paux    -= p->alpha_visc*vsignal*dvdr*winvrho;
      A->dudt -= 0.5*P->m[j]*p->alpha_visc*vsignal*dvdr*dvdr*winvrho;
    }
    else if (p->avisc == 2 || p->avisc == 3) {
      const double alpha_mean = 0.5*(P->alpha[i] + P->alpha[j]);
      const double vsignal = P->sound[i] + P->sound[j] - p->beta_visc*alpha_mean*dvdr;
      paux    -= alpha_mean*vsignal*dvdr*winvrho;
      A->dudt -= 0.5*P->m[j]*alpha_mean*vsignal*dvdr*dvdr*winvrho;
    }
    if (p->acond == 1) {
      A->dudt += P->m[j]*dvdr*(P->u[j] - P->u[i])*(invrho_i*wkerni + invrho_j*wkernj);
    }
    else if (p->acond == 2) {
      A->dudt += 0.5*P->m[j]*(P->u[i] - P->u[j])*winvrho*(invrho_i + invrho_j)*
        sqrt(fabs(P->pressure[i] - P->pressure[j]));
    }
  }
  for (int k = 0; k < nd; k++) A->a[k] += P->m[j]*draux[k]*paux;

  if (grav) {
    const double invhsqdi = invh_i*invh_i;
    paux = 0.5*(invhsqdi*orc_kernel(p, 4, drmag*invh_i) + P->zeta[i]*wkerni +
                invh_j*invh_j*orc_kernel(p, 4, drmag*invh_j) + P->zeta[j]*wkernj);
    for (int k = 0; k < nd; k++) A->atree[k] += P->m[j]*draux[k]*paux;
    A->gpot += 0.5*P->m[j]*(invh_i*orc_kernel(p, 5, drmag*invh_i) +
                            invh_j*orc_kernel(p, 5, drmag*invh_j));
    A->div_v -= P->m[j]*dvdr*wkerni;
  }
  if (P->level[j] > A->levelneib) A->levelneib = P->level[j];
}

/* ComputeSphGravForces GradhSph.cpp:598-647 (softened gravity between overlapping particles that
 * do not interact hydrodynamically, e.g. hydro_forces = 0) */
static void smoothgrav_pair(const orc_params* p, const pview* P, int i, int j, const double* dr_in,
                            double drsqd, acc_t* A)
{
  const int nd = p->ndim;
  double dr[3];
  const double invh_i = 1/P->h[i], invh_j = 1/P->h[j];
  const double drmag = sqrt(drsqd + SMALL_NUMBER), invdrmag = 1.0/drmag;
  for (int k = 0; k < nd; k++) dr[k] = dr_in[k]*invdrmag;
  const double paux = 0.5*(invh_i*invh_i*orc_kernel(p, 4, drmag*invh_i) +
                           P->zeta[i]*P->hfactor[i]*orc_kernel(p, 1, drmag*invh_i) +
                           invh_j*invh_j*orc_kernel(p, 4, drmag*invh_j) +
                           P->zeta[j]*P->hfactor[j]*orc_kernel(p, 1, drmag*invh_j));
  const double gaux = 0.5*(invh_i*orc_kernel(p, 5, drmag*invh_i) + invh_j*orc_kernel(p, 5, drmag*invh_j));
  for (int k = 0; k < nd; k++) A->atree[k] += P->m[j]*dr[k]*paux;
  A->gpot += P->m[j]*gaux;
}

/* ComputeDirectGravForces GradhSph.cpp:657-690 */
static inline void direct_pair(int nd, const double* dr, double mj, acc_t* A)
{
  const double drsqd = dot(dr, dr, nd) + SMALL_NUMBER;
  const double invdrmag = 1.0/sqrt(drsqd);
  const double invdr3 = invdrmag*invdrmag*invdrmag;
  for (int k = 0; k < nd; k++) A->atree[k] += mj*dr[k]*invdr3;
  A->gpot += mj*invdrmag;
}

/* end-of-loop terms: GradhSph.cpp:451-457 (hydro) / :576-582 (hydro+grav; note max(+div_v,0)) */
static void finish_particle(const orc_params* p, const pview* P, int i, int grav, acc_t* A,
                            double* dalphadt_out)
{
  const double invrho_i = 1/P->rho[i], invh_i = 1/P->h[i];
  A->div_v *= invrho_i;
  const double press = (p->gamma_eos - 1.0)*P->rho[i]*P->u[i];     /* eos->Pressure(parti) */
  A->dudt -= press*A->div_v*invrho_i*P->invomega[i];
  *dalphadt_out = 0.0;
  if (p->tdavisc == 1) {
    const double dv = grav ? A->div_v : -A->div_v;
    *dalphadt_out = 0.1*P->sound[i]*(p->alpha_visc_min - P->alpha[i])*invh_i +
      fmax(dv, 0.0)*(p->alpha_visc - P->alpha[i]);
  }
}

/* InvertMatrix, src/Headers/InlineFuncs.h:573-612 */
static void invert_matrix(int nd, double A[3][3], double B[3][3])
{
  if (nd == 1) B[0][0] = 1/A[0][0];
  else if (nd == 2) {
    const double invdet = 1.0/(A[0][0]*A[1][1] - A[0][1]*A[1][0]);
    B[0][0] = invdet*A[1][1]; B[0][1] = -invdet*A[0][1]; B[1][0] = -invdet*A[1][0]; B[1][1] = invdet*A[0][0];
  }
  else {
    const double invdet = 1.0/(A[0][0]*(A[1][1]*A[2][2] - A[2][1]*A[1][2]) - A[0][1]*(A[1][0]*A[2][2] - A[1][2]*A[2][0]) +
                               A[0][2]*(A[1][0]*A[2][1] - A[1][1]*A[2][0]));
    B[0][0] = (A[1][1]*A[2][2] - A[2][1]*A[1][2])*invdet; B[0][1] = (A[0][2]*A[2][1] - A[0][1]*A[2][2])*invdet;
    B[0][2] = (A[0][1]*A[1][2] - A[0][2]*A[1][1])*invdet; B[1][0] = (A[1][2]*A[2][0] - A[1][0]*A[2][2])*invdet;
    B[1][1] = (A[0][0]*A[2][2] - A[0][2]*A[2][0])*invdet; B[1][2] = (A[1][0]*A[0][2] - A[0][0]*A[1][2])*invdet;
    B[2][0] = (A[1][0]*A[2][1] - A[2][0]*A[1][1])*invdet; B[2][1] = (A[2][0]*A[0][1] - A[0][0]*A[2][1])*invdet;
    B[2][2] = (A[0][0]*A[1][1] - A[1][0]*A[0][1])*invdet;
  }
}

/* Sph<ndim>::ComputeCullenAndDehnenViscosity, src/Headers/Sph.h:364-456.  The reference sums over the
 * whole gather list of the cell; w1 vanishes outside the kernel of the final h, so any complete list
 * gives the same sums (here: every particle, ascending index). */
int orc_cd2010_switch(const orc_params* p, int N, const double* r, const double* v, const double* a,
                      const double* m, const double* h, const double* rho, const double* hfactor,
                      const double* sound, double* alpha, double* dalphadt)
{
  const int nd = p->ndim;
  const int Nt = p->ntarget > 0 ? p->ntarget : N;
  #pragma omp parallel for schedule(dynamic, 32)
  for (int i = 0; i < Nt; i++) {
    double rr[3][3] = {{0}}, dv[3][3] = {{0}}, da[3][3] = {{0}}, T[3][3] = {{0}}, dvdx[3][3] = {{0}}, dadx[3][3] = {{0}};
    const double invh = 1/h[i];
    const double hfac = invh*hfactor[i]/rho[i];
    for (int j = 0; j < N; j++) {
      double dr[3] = {0, 0, 0};
      for (int k = 0; k < nd; k++) dr[k] = r[(size_t)j*nd + k] - r[(size_t)i*nd + k];
      wrap(p, dr);
      const double w = m[j]*hfac*orc_kernel(p, 1, invh*sqrt(dot(dr, dr, nd)));
      if (w == 0.0) continue;
      for (int jj = 0; jj < nd; jj++)
        for (int k = 0; k < nd; k++) {
          rr[jj][k] += w*dr[jj]*dr[k];
          dv[jj][k] += w*dr[jj]*(v[(size_t)j*nd + k] - v[(size_t)i*nd + k]);
          da[jj][k] += w*dr[jj]*(a[(size_t)j*nd + k] - a[(size_t)i*nd + k]);
        }
    }
    invert_matrix(nd, rr, T);
    double modR = 0, modT = 0;
    for (int ii = 0; ii < nd; ii++)
      for (int j = 0; j < nd; j++) { modR += rr[ii][j]*rr[ii][j]; modT += T[ii][j]*T[ii][j]; }
    const double sqd_condition_number = modR*modT/(nd*nd);
    double alpha_loc = 0;
    if (sqd_condition_number > 1e4) alpha_loc = p->alpha_visc;
    else {
      for (int ii = 0; ii < nd; ii++)
        for (int j = 0; j < nd; j++)
          for (int k = 0; k < nd; k++) { dvdx[ii][j] += T[j][k]*dv[k][ii]; dadx[ii][j] += T[j][k]*da[k][ii]; }
      double ddivdt = 0, divv2 = 0;
      for (int ii = 0; ii < nd; ii++) {
        ddivdt += dadx[ii][ii];
        for (int j = 0; j < nd; j++) ddivdt -= dvdx[ii][j]*dvdx[j][ii];
        divv2 += dvdx[ii][ii];
      }
      divv2 *= divv2;
      double curlv2 = 0;
      if (nd == 2) { const double c = dvdx[1][0] - dvdx[0][1]; curlv2 = c*c; }
      if (nd == 3) {
        const double c[3] = { dvdx[1][2] - dvdx[2][1], dvdx[2][0] - dvdx[0][2], dvdx[0][1] - dvdx[1][0] };
        curlv2 = dot(c, c, 3);
      }
      double f_balsara = 1;
      if (curlv2 > 0) f_balsara = divv2/(divv2 + curlv2);
      if (ddivdt < 0) {
        alpha_loc = (10*h[i]*h[i]/(sound[i]*sound[i]))*f_balsara*(-ddivdt);
        alpha_loc = fmin(alpha_loc, p->alpha_visc);
      }
    }
    if (alpha_loc > alpha[i]) alpha[i] = alpha_loc;
    dalphadt[i] = 0.1*sound[i]*(fmax(p->alpha_visc_min, alpha_loc) - alpha[i])*invh;
  }
  return 0;
}

/* GradhSph::ComputeStarGravForces, src/GradhSph/GradhSph.cpp:699-743 (fixed_sink_mass = 0) */
int orc_star_grav_forces(const orc_params* p, int N, const double* r, const double* h, int Nstar,
                         const double* stars, double* a, double* atree, double* gpot)
{
  const int nd = p->ndim;
  for (int i = 0; i < N; i++) {
    double at[3] = {0, 0, 0}, pot = 0.0;
    for (int j = 0; j < Nstar; j++) {
      const double* s = stars + (size_t)j*(nd + 2);
      const double ms = s[nd], hs = s[nd + 1];
      double dr[3] = {0, 0, 0}, drsqd = 0.0;
      for (int k = 0; k < nd; k++) { dr[k] = s[k] - r[(size_t)i*nd + k]; drsqd += dr[k]*dr[k]; }
      drsqd += 1.0e-20;
      const double drmag = sqrt(drsqd), invdrmag = 1.0/drmag;
      const double invhmean = 2.0/(h[i] + hs);
      const double paux = ms*invhmean*invhmean*orc_kernel(p, 4, drmag*invhmean)*invdrmag;
      for (int k = 0; k < nd; k++) at[k] += paux*dr[k];
      pot += ms*invhmean*orc_kernel(p, 5, drmag*invhmean);
    }
    for (int k = 0; k < nd; k++) { atree[(size_t)i*nd + k] += at[k]; a[(size_t)i*nd + k] += at[k]; }
    gpot[i] += pot;
  }
  return 0;
}

/* GradhSphTree::UpdateAllSphHydroForces, src/GradhSph/GradhSphTree.cpp:327-406 */
int orc_hydro_forces(const orc_params* p, int N, const double* r, const double* v, const double* m,
                     const double* h, const double* rho, const double* pressure,
                     const double* invomega, const double* sound, const double* u,
                     const double* alpha, const double* hfactor, const double* hrangesqd,
                     const int32_t* level,
                     double* a, double* dudt, double* div_v, double* dalphadt, int32_t* levelneib)
{
  const int nd = p->ndim;
  pview P = { r, v, m, h, rho, pressure, invomega, 0, sound, u, alpha, hfactor, hrangesqd, level };
  const int Nt = p->ntarget > 0 ? p->ntarget : N;     /* targets = the first Nt particles (the rest: mirror images) */
  #pragma omp parallel for schedule(dynamic, 32)
  for (int i = 0; i < Nt; i++) {
    acc_t A; memset(&A, 0, sizeof A);
    for (int j = 0; j < N; j++) {
      double dr[3], drsqd;
      if (!is_neighbour(p, r + (size_t)i*nd, r + (size_t)j*nd, hrangesqd[i], hrangesqd[j], dr, &drsqd))
        continue;
      hydro_pair(p, &P, i, j, dr, drsqd, 0, &A);
    }
    double dadt;
    finish_particle(p, &P, i, 0, &A, &dadt);
    for (int k = 0; k < nd; k++) a[(size_t)i*nd + k] += A.a[k];
    dudt[i] += A.dudt; div_v[i] += A.div_v; dalphadt[i] += dadt;
    if (A.levelneib > levelneib[i]) levelneib[i] = A.levelneib;
  }
  /* scatter half of the levelneib update (GradhSph.cpp:446, GradhSphTree.cpp:376-382) */
  for (int i = 0; i < Nt; i++) {
    if (level[i] == 0) continue;
    for (int j = 0; j < Nt; j++)
      if (is_neighbour(p, r + (size_t)i*nd, r + (size_t)j*nd, hrangesqd[i], hrangesqd[j], 0, 0))
        if (level[i] > levelneib[j]) levelneib[j] = level[i];
  }
  return 0;
}

/* neib_search = bruteforce semantics: every pair is either an SPH neighbour or a direct-sum
 * neighbour (GradhSphTree.cpp:498-620 with an interaction list that contains every particle). */
int orc_hydro_grav_forces_direct(const orc_params* p, int N, const double* r, const double* v,
                     const double* m, const double* h, const double* rho, const double* pressure,
                     const double* invomega, const double* zeta, const double* sound,
                     const double* u, const double* alpha, const double* hfactor,
                     const double* hrangesqd, const int32_t* level,
                     double* a, double* atree, double* gpot, double* dudt, double* div_v,
                     double* dalphadt, int32_t* levelneib)
{
  const int nd = p->ndim;
  pview P = { r, v, m, h, rho, pressure, invomega, zeta, sound, u, alpha, hfactor, hrangesqd, level };
  (void) dalphadt;   /* the hydro+gravity driver never writes dalphadt back (GradhSphTree.cpp:610-620) */
  #pragma omp parallel for schedule(dynamic, 32)
  for (int i = 0; i < N; i++) {
    acc_t A; memset(&A, 0, sizeof A);
    A.gpot = (m[i]/h[i])*orc_kernel(p, 5, 0.0);                    /* GradhSphTree.cpp:512 */
    for (int j = 0; j < N; j++) {
      double dr[3], drsqd;
      if (is_neighbour(p, r + (size_t)i*nd, r + (size_t)j*nd, hrangesqd[i], hrangesqd[j], dr, &drsqd)) {
        if (p->hydro_forces) hydro_pair(p, &P, i, j, dr, drsqd, 1, &A);
        else smoothgrav_pair(p, &P, i, j, dr, drsqd, &A);
      }
      else direct_pair(nd, dr, m[j], &A);
    }
    double dadt;
    finish_particle(p, &P, i, 1, &A, &dadt);
    for (int k = 0; k < nd; k++) {
      a[(size_t)i*nd + k] += A.a[k];
      a[(size_t)i*nd + k] += A.atree[k];
      atree[(size_t)i*nd + k] += A.atree[k];
    }
    gpot[i] += A.gpot; dudt[i] += A.dudt; div_v[i] += A.div_v;
    if (A.levelneib > levelneib[i]) levelneib[i] = A.levelneib;
  }
  return 0;
}

/*================================================================================================
 * KD-tree: src/Tree/KDTree.cpp.  ComputeTreeSize :322-352, CreateTreeStructure :362-433,
 * BuildTree :220-313, DivideTreeCell :442-595, QuickSelect :682-750, StockCellProperties :808-1083.
 *===============================================================================================*/
typedef struct {
  int level, copen, cnext, c1, c2, ifirst, ilast, N;
  double bbmin[3], bbmax[3], hbmin[3], hbmax[3], rcell[3], r[3], v[3];
  double m, rmax, hmax, cdistsqd, mac, q[5];
  double amin, macfactor;          /* gadget2 / eigenmac auxiliaries, KDTree.cpp:838-841,899-903,967-984 */
} kdcell;

struct orc_tree {
  orc_params p;
  int N, ltot, gtot, Ncell;
  kdcell* cell;
  int *ids, *inext;
  const double *r, *v, *m, *h;
  const double *atree, *gpot;      /* previous-step values read by the gadget2 / eigenmac criteria (may be NULL) */
};

static double quickselect(orc_tree* T, int left, int right, int jpivot, int k)
{
  int* ids = T->ids; const int nd = T->p.ndim; const double* r = T->r;
  int j, jguess, jtemp; double rpivot;
  do {
    jguess = (left + right)/2;
    rpivot = r[(size_t)ids[jguess]*nd + k];
    jtemp = ids[jguess]; ids[jguess] = ids[right]; ids[right] = jtemp;
    jguess = left;
    for (j = left; j < right; j++) {
      if (r[(size_t)ids[j]*nd + k] <= rpivot) {
        jtemp = ids[j]; ids[j] = ids[jguess]; ids[jguess] = jtemp;
        jguess++;
      }
    }
    jtemp = ids[right]; ids[right] = ids[jguess]; ids[jguess] = jtemp;
    if (jguess < jpivot) left = jguess + 1;
    else if (jguess > jpivot) right = jguess - 1;
  } while (jguess != jpivot);
  return rpivot;
}

static void stock_cell(orc_tree* T, kdcell* c)
{
  const orc_params* p = &T->p; const int nd = p->ndim;
  const double kr = kern_range(p);
  const int need_q = (p->multipole == 1 || p->multipole == 3 || p->gravity_mac == 1);
  double dr[3], drsqd, mi;
  int k, i;
  c->N = 0; c->m = 0; c->hmax = 0; c->rmax = 0; c->mac = 0; c->cdistsqd = BIG_NUMBER;
  c->amin = BIG_NUMBER; c->macfactor = 0;
  for (k = 0; k < 5; k++) c->q[k] = 0;
  for (k = 0; k < nd; k++) {
    c->r[k] = 0; c->v[k] = 0; c->rcell[k] = 0;
    c->bbmin[k] = BIG_NUMBER; c->bbmax[k] = -BIG_NUMBER;
    c->hbmin[k] = BIG_NUMBER; c->hbmax[k] = -BIG_NUMBER;
  }
  if (c->level == T->ltot) {
    i = c->ifirst;
    while (i != -1) {
      c->N++;
      c->hmax = fmax(c->hmax, T->h[i]);
      c->m += T->m[i];
      for (k = 0; k < nd; k++) c->r[k] += T->m[i]*T->r[(size_t)i*nd + k];
      for (k = 0; k < nd; k++) c->v[k] += T->m[i]*T->v[(size_t)i*nd + k];
      for (k = 0; k < nd; k++) {
        const double x = T->r[(size_t)i*nd + k];
        if (x < c->bbmin[k]) c->bbmin[k] = x;
        if (x > c->bbmax[k]) c->bbmax[k] = x;
        if (x - kr*T->h[i] < c->hbmin[k]) c->hbmin[k] = x - kr*T->h[i];
        if (x + kr*T->h[i] > c->hbmax[k]) c->hbmax[k] = x + kr*T->h[i];
      }
      if (p->gravity_mac == 2 && T->atree)                                  /* KDTree.cpp:899-901 */
        c->amin = fmin(c->amin, sqrt(dot(T->atree + (size_t)i*nd, T->atree + (size_t)i*nd, nd)));
      else if (p->gravity_mac == 1 && T->gpot)                              /* KDTree.cpp:902-903 */
        c->macfactor = fmax(c->macfactor, pow(T->gpot[i], -0.66666666666666666666666));
      if (i == c->ilast) break;
      i = T->inext[i];
    }
    if (c->m > 0) {
      for (k = 0; k < nd; k++) c->r[k] /= c->m;
      for (k = 0; k < nd; k++) c->v[k] /= c->m;
    }
    if (c->N > 0) {
      for (k = 0; k < nd; k++) c->rcell[k] = 0.5*(c->bbmin[k] + c->bbmax[k]);
      for (k = 0; k < nd; k++) dr[k] = 0.5*(c->bbmax[k] - c->bbmin[k]);
      c->cdistsqd = fmax(dot(dr, dr, nd), c->hmax*c->hmax)/p->thetamaxsqd;
      c->rmax = sqrt(dot(dr, dr, nd));
    }
    if (need_q) {
      i = c->ifirst;
      while (i != -1) {
        mi = T->m[i];
        for (k = 0; k < nd; k++) dr[k] = T->r[(size_t)i*nd + k] - c->r[k];
        drsqd = dot(dr, dr, nd);
        if (nd == 3) {
          c->q[0] += mi*(3.0*dr[0]*dr[0] - drsqd);
          c->q[1] += mi*3.0*dr[0]*dr[1];
          c->q[2] += mi*(3.0*dr[1]*dr[1] - drsqd);
          c->q[3] += mi*3.0*dr[2]*dr[0];
          c->q[4] += mi*3.0*dr[2]*dr[1];
        }
        else if (nd == 2) {
          c->q[0] += mi*(3.0*dr[0]*dr[0] - drsqd);
          c->q[1] += mi*3.0*dr[0]*dr[1];
          c->q[2] += mi*(3.0*dr[1]*dr[1] - drsqd);
        }
        else c->q[0] += mi*(3.0*dr[0]*dr[0] - drsqd);
        if (i == c->ilast) break;
        i = T->inext[i];
      }
    }
  }
  else {
    kdcell* ch[2] = { &T->cell[c->c1], &T->cell[c->c2] };
    for (int s = 0; s < 2; s++) {
      if (ch[s]->N > 0) {
        for (k = 0; k < nd; k++) {
          c->bbmin[k] = fmin(ch[s]->bbmin[k], c->bbmin[k]);
          c->bbmax[k] = fmax(ch[s]->bbmax[k], c->bbmax[k]);
          c->hbmin[k] = fmin(ch[s]->hbmin[k], c->hbmin[k]);
          c->hbmax[k] = fmax(ch[s]->hbmax[k], c->hbmax[k]);
        }
        c->hmax = fmax(c->hmax, ch[s]->hmax);
        if (p->gravity_mac == 2) c->amin = fmin(c->amin, ch[s]->amin);
        else if (p->gravity_mac == 1) c->macfactor = fmax(c->macfactor, ch[s]->macfactor);
      }
    }
    c->N = ch[0]->N + ch[1]->N;
    c->m = ch[0]->m + ch[1]->m;
    if (c->m > 0) {
      for (k = 0; k < nd; k++) c->r[k] = (ch[0]->m*ch[0]->r[k] + ch[1]->m*ch[1]->r[k])/c->m;
      for (k = 0; k < nd; k++) c->v[k] = (ch[0]->m*ch[0]->v[k] + ch[1]->m*ch[1]->v[k])/c->m;
    }
    if (c->N > 0) {
      for (k = 0; k < nd; k++) c->rcell[k] = 0.5*(c->bbmin[k] + c->bbmax[k]);
      for (k = 0; k < nd; k++) dr[k] = 0.5*(c->bbmax[k] - c->bbmin[k]);
      c->cdistsqd = fmax(dot(dr, dr, nd), c->hmax*c->hmax)/p->thetamaxsqd;
      c->rmax = sqrt(dot(dr, dr, nd));
    }
    for (int s = 0; s < 2; s++) {
      if (need_q && ch[s]->m > 0) {
        mi = ch[s]->m;
        for (k = 0; k < nd; k++) dr[k] = ch[s]->r[k] - c->r[k];
        drsqd = dot(dr, dr, nd);
        if (nd == 3) {
          for (k = 0; k < 5; k++) c->q[k] += ch[s]->q[k];
          c->q[0] += mi*(3.0*dr[0]*dr[0] - drsqd);
          c->q[1] += mi*3.0*dr[0]*dr[1];
          c->q[2] += mi*(3.0*dr[1]*dr[1] - drsqd);
          c->q[3] += mi*3.0*dr[2]*dr[0];
          c->q[4] += mi*3.0*dr[2]*dr[1];
        }
        else if (nd == 2) {
          for (k = 0; k < 3; k++) c->q[k] += ch[s]->q[k];
          c->q[0] += mi*(3.0*dr[0]*dr[0] - drsqd);
          c->q[1] += mi*3.0*dr[0]*dr[1];
          c->q[2] += mi*(3.0*dr[1]*dr[1] - drsqd);
        }
        else {
          c->q[0] += ch[s]->q[0];
          c->q[0] += mi*(3.0*dr[0]*dr[0] - drsqd);
        }
      }
    }
  }
  /* eigenvalue MAC, KDTree.cpp:1058-1079 */
  if (p->gravity_mac == 1) {
    double lambda, pp;
    if (nd == 3) {
      pp = c->q[0]*c->q[2] - (c->q[0] + c->q[2])*(c->q[0] + c->q[2]) - c->q[1]*c->q[1] - c->q[3]*c->q[3] - c->q[4]*c->q[4];
      lambda = pp >= 0.0 ? 0.0 : 2.0*sqrt(-pp/3.0);
    }
    else if (nd == 2) {
      pp = (c->q[0] - c->q[2])*(c->q[0] - c->q[2]) + 4*c->q[1]*c->q[1];
      lambda = 0.5*fmax(c->q[0] + c->q[2] + sqrt(pp), 0.);
    }
    else lambda = fabs(c->q[0]);
    c->mac = pow(0.5*lambda/p->macerror, 0.66666666666666);
  }
}

static void divide_cell(orc_tree* T, int ifirst, int ilast, kdcell* cell)
{
  const int nd = T->p.ndim;
  int k, j, k_divide = 0;
  double rkmax = 0.0, rdivide;
  if (cell->level == T->ltot) {
    if (cell->N > 0) {
      for (j = cell->ifirst; j < cell->ilast; j++) T->inext[T->ids[j]] = T->ids[j + 1];
      cell->ifirst = T->ids[cell->ifirst];
      cell->ilast  = T->ids[cell->ilast];
    }
    else { cell->ifirst = -1; cell->ilast = -1; }
    stock_cell(T, cell);
    return;
  }
  for (k = 0; k < nd; k++) {
    if (cell->bbmax[k] - cell->bbmin[k] > rkmax) {
      rkmax = cell->bbmax[k] - cell->bbmin[k];
      k_divide = k;
    }
  }
  rdivide = quickselect(T, cell->ifirst, cell->ilast, cell->ifirst + cell->N/2, k_divide);
  kdcell* c1 = &T->cell[cell->c1]; kdcell* c2 = &T->cell[cell->c2];
  for (k = 0; k < nd; k++) { c1->bbmin[k] = cell->bbmin[k]; c1->bbmax[k] = cell->bbmax[k]; }
  c1->bbmax[k_divide] = rdivide;
  c1->N = cell->N/2;
  if (c1->N != 0) { c1->ifirst = ifirst; c1->ilast = ifirst + cell->N/2 - 1; }
  for (k = 0; k < nd; k++) { c2->bbmin[k] = cell->bbmin[k]; c2->bbmax[k] = cell->bbmax[k]; }
  c2->bbmin[k_divide] = rdivide;
  c2->N = cell->N - c1->N;
  if (c2->N != 0) { c2->ifirst = ifirst + cell->N/2; c2->ilast = ilast; }

  divide_cell(T, ifirst, ifirst + cell->N/2 - 1, c1);
  divide_cell(T, ifirst + cell->N/2, ilast, c2);

  if (c1->N > 0) { cell->ifirst = c1->ifirst; T->inext[c1->ilast] = c2->ifirst; }
  else cell->ifirst = c2->ifirst;
  cell->ilast = c2->ilast;
  stock_cell(T, cell);
}

/* KDTree::UpdateHmaxValues (KDTree.cpp:1128-1208): after the density pass only hmax and hbox of the cells follow
 * the new smoothing lengths; cdistsqd keeps the value of the last stock.  h replaces the tree's h array. */
static void update_hmax_cell(orc_tree* T, kdcell* c)
{
  const int nd = T->p.ndim; const double kr = kern_range(&T->p);
  int k;
  if (c->level != T->ltot) { update_hmax_cell(T, &T->cell[c->c1]); update_hmax_cell(T, &T->cell[c->c2]); }
  c->hmax = 0.0;
  for (k = 0; k < nd; k++) { c->hbmin[k] = BIG_NUMBER; c->hbmax[k] = -BIG_NUMBER; }
  if (c->level == T->ltot) {
    int i = c->ifirst;
    while (i != -1) {
      c->hmax = fmax(c->hmax, T->h[i]);
      for (k = 0; k < nd; k++) {
        const double x = T->r[(size_t)i*nd + k];
        if (x - kr*T->h[i] < c->hbmin[k]) c->hbmin[k] = x - kr*T->h[i];
        if (x + kr*T->h[i] > c->hbmax[k]) c->hbmax[k] = x + kr*T->h[i];
      }
      if (i == c->ilast) break;
      i = T->inext[i];
    }
  }
  else {
    kdcell* ch[2] = { &T->cell[c->c1], &T->cell[c->c2] };
    for (int s = 0; s < 2; s++) {
      if (ch[s]->N > 0) {
        c->hmax = fmax(c->hmax, ch[s]->hmax);
        for (k = 0; k < nd; k++) { c->hbmin[k] = fmin(ch[s]->hbmin[k], c->hbmin[k]); c->hbmax[k] = fmax(ch[s]->hbmax[k], c->hbmax[k]); }
      }
    }
  }
}

void orc_tree_update_hmax(orc_tree* T, const double* h)
{
  T->h = h;
  update_hmax_cell(T, &T->cell[0]);
}

orc_tree* orc_tree_build(const orc_params* p, int N, const double* r, const double* v,
                         const double* m, const double* h)
{
  return orc_tree_build_mac(p, N, r, v, m, h, 0, 0);
}

orc_tree* orc_tree_build_mac(const orc_params* p, int N, const double* r, const double* v,
                             const double* m, const double* h, const double* atree, const double* gpot)
{
  orc_tree* T = (orc_tree*) calloc(1, sizeof(orc_tree));
  const int nd = p->ndim; const double kr = kern_range(p);
  T->p = *p; T->N = N; T->r = r; T->v = v; T->m = m; T->h = h; T->atree = atree; T->gpot = gpot;
  T->ltot = 0;
  while (p->Nleafmax*pow(2, T->ltot) < N) T->ltot++;
  T->gtot = (int) pow(2, T->ltot);
  T->Ncell = 2*T->gtot - 1;
  T->cell = (kdcell*) calloc(T->Ncell, sizeof(kdcell));
  T->ids = (int*) malloc(sizeof(int)*N);
  T->inext = (int*) malloc(sizeof(int)*N);
  /* CreateTreeStructure */
  for (int c = 0; c < T->Ncell; c++) {
    kdcell* cc = &T->cell[c];
    cc->copen = cc->cnext = cc->c1 = cc->c2 = cc->ifirst = cc->ilast = -1; cc->N = 0;
  }
  T->cell[0].level = 0;
  for (int c = 0; c < T->Ncell; c++) {
    kdcell* cc = &T->cell[c];
    if (cc->level == T->ltot) cc->cnext = c + 1;
    else {
      const int c2L = (int) pow(2, T->ltot - cc->level);
      T->cell[c + 1].level = cc->level + 1;
      cc->copen = c + 1; cc->c1 = c + 1; cc->c2 = c + c2L;
      T->cell[cc->c2].level = cc->level + 1;
      cc->cnext = c + 2*c2L - 1;
    }
  }
  /* BuildTree: root box = particle extent inflated by kernrange*h (:272-279) */
  double bbmin[3], bbmax[3];
  for (int k = 0; k < nd; k++) { bbmin[k] = BIG_NUMBER; bbmax[k] = -BIG_NUMBER; }
  for (int i = 0; i < N; i++) {
    for (int k = 0; k < nd; k++) {
      bbmax[k] = fmax(bbmax[k], r[(size_t)i*nd + k] + kr*h[i]);
      bbmin[k] = fmin(bbmin[k], r[(size_t)i*nd + k] - kr*h[i]);
    }
    T->ids[i] = i; T->inext[i] = (i < N - 1) ? i + 1 : -1;
  }
  T->cell[0].N = N; T->cell[0].ifirst = 0; T->cell[0].ilast = N - 1;
  for (int k = 0; k < nd; k++) { T->cell[0].bbmin[k] = bbmin[k]; T->cell[0].bbmax[k] = bbmax[k]; }
  if (N > 0) divide_cell(T, 0, N - 1, &T->cell[0]);
  return T;
}

void orc_tree_free(orc_tree* t) { if (t) { free(t->cell); free(t->ids); free(t->inext); free(t); } }
int orc_tree_ncell(const orc_tree* t) { return t->Ncell; }
int orc_tree_ltot(const orc_tree* t) { return t->ltot; }

void orc_tree_leaf_of(const orc_tree* t, int32_t* leaf_of)
{
  for (int c = 0; c < t->Ncell; c++) {
    const kdcell* cc = &t->cell[c];
    if (cc->copen != -1 || cc->N == 0) continue;
    int i = cc->ifirst;
    while (i != -1) { leaf_of[i] = c; if (i == cc->ilast) break; i = t->inext[i]; }
  }
}

void orc_tree_cells(const orc_tree* t, double* box, double* com, double* scalars, double* q, int32_t* N)
{
  const int nd = t->p.ndim;
  for (int c = 0; c < t->Ncell; c++) {
    const kdcell* cc = &t->cell[c];
    for (int k = 0; k < nd; k++) {
      if (box) {
        box[((size_t)c*4 + 0)*nd + k] = cc->bbmin[k]; box[((size_t)c*4 + 1)*nd + k] = cc->bbmax[k];
        box[((size_t)c*4 + 2)*nd + k] = cc->hbmin[k]; box[((size_t)c*4 + 3)*nd + k] = cc->hbmax[k];
      }
      if (com) { com[((size_t)c*2 + 0)*nd + k] = cc->rcell[k]; com[((size_t)c*2 + 1)*nd + k] = cc->r[k]; }
    }
    if (scalars) {
      scalars[c*5 + 0] = cc->m; scalars[c*5 + 1] = cc->rmax; scalars[c*5 + 2] = cc->hmax;
      scalars[c*5 + 3] = cc->cdistsqd; scalars[c*5 + 4] = cc->mac;
    }
    if (q) for (int k = 0; k < 5; k++) q[c*5 + k] = cc->q[k];
    if (N) N[c] = cc->N;
  }
}

/*------------------------------------------------------------------------------------------------
 * Cell multipole forces: ComputeCellMonopoleForces / ComputeQuadropole,
 * src/Headers/NeighbourSearch.h:349-377, :384-455.
 *-----------------------------------------------------------------------------------------------*/
static void cell_monopole(int nd, const kdcell* c, const double* rp, acc_t* A)
{
  double dr[3];
  for (int k = 0; k < nd; k++) dr[k] = c->r[k] - rp[k];
  const double drsqd = dot(dr, dr, nd) + SMALL_NUMBER;
  const double invdrsqd = 1.0/drsqd, invdrmag = sqrt(invdrsqd), invdr3 = invdrsqd*invdrmag;
  A->gpot += c->m*invdrmag;
  for (int k = 0; k < nd; k++) A->atree[k] += c->m*dr[k]*invdr3;
}

static void cell_quadrupole(int nd, const kdcell* c, const double* rp, acc_t* A)
{
  double dr[3] = {0, 0, 0};
  for (int k = 0; k < nd; k++) dr[k] = rp[k] - c->r[k];
  const double drsqd = dot(dr, dr, nd) + SMALL_NUMBER;
  const double invdrsqd = 1.0/drsqd, invdrmag = sqrt(invdrsqd);
  const double invdr5 = invdrsqd*invdrsqd*invdrmag;
  for (int k = 0; k < nd; k++) A->atree[k] -= c->m*dr[k]*invdrsqd*invdrmag;
  double qscalar, qfactor;
  if (nd == 3) {
    qscalar = c->q[0]*dr[0]*dr[0] + c->q[2]*dr[1]*dr[1] - (c->q[0] + c->q[2])*dr[2]*dr[2] +
      2.0*(c->q[1]*dr[0]*dr[1] + c->q[3]*dr[0]*dr[2] + c->q[4]*dr[1]*dr[2]);
    qfactor = 2.5*qscalar*invdr5*invdrsqd;
    A->atree[0] += (c->q[0]*dr[0] + c->q[1]*dr[1] + c->q[3]*dr[2])*invdr5 - qfactor*dr[0];
    A->atree[1] += (c->q[1]*dr[0] + c->q[2]*dr[1] + c->q[4]*dr[2])*invdr5 - qfactor*dr[1];
    A->atree[2] += (c->q[3]*dr[0] + c->q[4]*dr[1] - (c->q[0] + c->q[2])*dr[2])*invdr5 - qfactor*dr[2];
  }
  else if (nd == 2) {
    qscalar = c->q[0]*dr[0]*dr[0] + c->q[2]*dr[1]*dr[1] + 2.0*c->q[1]*dr[0]*dr[1];
    qfactor = 2.5*qscalar*invdr5*invdrsqd;
    A->atree[0] += (c->q[0]*dr[0] + c->q[1]*dr[1])*invdr5 - qfactor*dr[0];
    A->atree[1] += (c->q[1]*dr[0] + c->q[2]*dr[1])*invdr5 - qfactor*dr[1];
  }
  else {
    qscalar = c->q[0]*dr[0]*dr[0];
    qfactor = 2.5*qscalar*invdr5*invdrsqd;
    A->atree[0] += (c->q[0]*dr[0])*invdr5 - qfactor*dr[0];
  }
  A->gpot += c->m*invdrmag + 0.5*qscalar*invdr5;
}

/* FastMultipoleForces (3-D and lower), src/Headers/NeighbourSearch.h:481-760, HoT = false */
typedef struct { double rc[3], ac[3], dphi[3], q[6], pot; } fastmp;

static void fast_add_monopole(int nd, fastmp* F, const kdcell* c)
{
  double dr[3] = {0, 0, 0};
  for (int k = 0; k < nd; k++) dr[k] = c->r[k] - F->rc[k];
  const double invdrmag = sqrt(1.0/dot(dr, dr, nd));
  const double invdrsqd = invdrmag*invdrmag, invdr3 = invdrsqd*invdrmag;
  double mc = c->m;
  F->pot += mc*invdrmag;
  mc *= invdr3;
  for (int k = 0; k < nd; k++) F->ac[k] += mc*dr[k];
  for (int k = 0; k < nd; k++) F->dphi[k] += mc*dr[k];
  F->q[0] += mc*(3.0*dr[0]*dr[0]*invdrsqd - 1);
  if (nd >= 2) {
    F->q[1] += mc*(3.0*dr[0]*dr[1]*invdrsqd);
    F->q[2] += mc*(3.0*dr[1]*dr[1]*invdrsqd - 1);
  }
  if (nd == 3) {
    F->q[3] += mc*(3.0*dr[2]*dr[0]*invdrsqd);
    F->q[4] += mc*(3.0*dr[2]*dr[1]*invdrsqd);
    F->q[5] += mc*(3.0*dr[2]*dr[2]*invdrsqd - 1);
  }
}

static void fast_add_quadrupole(int nd, fastmp* F, const kdcell* c)
{
  fast_add_monopole(nd, F, c);
  if (nd != 3) return;           /* only the 3-D form is restated (bossbodenheimer/cloud configs) */
  double dr[3];
  for (int k = 0; k < 3; k++) dr[k] = F->rc[k] - c->r[k];
  const double drsqd = dot(dr, dr, 3) + SMALL_NUMBER;
  const double invdrsqd = 1.0/drsqd, invdrmag = sqrt(invdrsqd);
  const double invdr5 = invdrsqd*invdrsqd*invdrmag;
  const double qscalar = c->q[0]*dr[0]*dr[0] + c->q[2]*dr[1]*dr[1] - (c->q[0] + c->q[2])*dr[2]*dr[2] +
    2.0*(c->q[1]*dr[0]*dr[1] + c->q[3]*dr[0]*dr[2] + c->q[4]*dr[1]*dr[2]);
  const double qfactor = 2.5*qscalar*invdr5*invdrsqd;
  double qx[3];
  qx[0] = (c->q[0]*dr[0] + c->q[1]*dr[1] + c->q[3]*dr[2])*invdr5;
  qx[1] = (c->q[1]*dr[0] + c->q[2]*dr[1] + c->q[4]*dr[2])*invdr5;
  qx[2] = (c->q[3]*dr[0] + c->q[4]*dr[1] - (c->q[0] + c->q[2])*dr[2])*invdr5;
  F->pot += 0.5*qscalar*invdr5;
  for (int k = 0; k < 3; k++) F->ac[k]   += qx[k] - qfactor*dr[k];
  for (int k = 0; k < 3; k++) F->dphi[k] += qx[k] - qfactor*dr[k];
  for (int k = 0; k < 3; k++) qx[k] *= 5*invdrsqd;
  F->q[0] += qfactor*(7*dr[0]*dr[0]*invdrsqd - 1);
  F->q[1] += qfactor*(7*dr[0]*dr[1]*invdrsqd);
  F->q[2] += qfactor*(7*dr[1]*dr[1]*invdrsqd - 1);
  F->q[3] += qfactor*(7*dr[0]*dr[2]*invdrsqd);
  F->q[4] += qfactor*(7*dr[1]*dr[2]*invdrsqd);
  F->q[5] += qfactor*(7*dr[2]*dr[2]*invdrsqd - 1);
  F->q[0] -= qx[0]*dr[0] + qx[0]*dr[0] - c->q[0]*invdr5;
  F->q[1] -= qx[0]*dr[1] + qx[1]*dr[0] - c->q[1]*invdr5;
  F->q[2] -= qx[1]*dr[1] + qx[1]*dr[1] - c->q[2]*invdr5;
  F->q[3] -= qx[0]*dr[2] + qx[2]*dr[0] - c->q[3]*invdr5;
  F->q[4] -= qx[1]*dr[2] + qx[2]*dr[1] - c->q[4]*invdr5;
  F->q[5] -= qx[2]*dr[2] + qx[2]*dr[2] + (c->q[0] + c->q[2])*invdr5;
}

static void fast_apply(int nd, const fastmp* F, const double* r, acc_t* A)
{
  double dr[3] = {0, 0, 0};
  for (int k = 0; k < nd; k++) dr[k] = r[k] - F->rc[k];
  if (nd == 3) {
    A->atree[0] += F->ac[0] + F->q[0]*dr[0] + F->q[1]*dr[1] + F->q[3]*dr[2];
    A->atree[1] += F->ac[1] + F->q[1]*dr[0] + F->q[2]*dr[1] + F->q[4]*dr[2];
    A->atree[2] += F->ac[2] + F->q[3]*dr[0] + F->q[4]*dr[1] + F->q[5]*dr[2];
    A->gpot += F->pot + F->dphi[0]*dr[0] + F->dphi[1]*dr[1] + F->dphi[2]*dr[2];
  }
  else if (nd == 2) {
    A->atree[0] += F->ac[0] + F->q[0]*dr[0] + F->q[1]*dr[1];
    A->atree[1] += F->ac[1] + F->q[1]*dr[0] + F->q[2]*dr[1];
    A->gpot += F->pot + F->dphi[0]*dr[0] + F->dphi[1]*dr[1];
  }
  else {
    A->atree[0] += F->ac[0] + F->q[0]*dr[0];
    A->gpot += F->pot + F->dphi[0]*dr[0];
  }
}

/*------------------------------------------------------------------------------------------------
 * GradhSphTree::UpdateAllSphForces, src/GradhSph/GradhSphTree.cpp:498-620 per leaf cell:
 *  walk  Tree::ComputeGravityInteractionAndGhostList  src/Tree/Tree.cpp:648-731 (geometric MAC,
 *        open_cell_for_gravity Tree.h:413-432; open boundaries)
 *  cull  NeighbourManager::_EndSearch                 NeighbourManager.h:396-470
 *  trim  NeighbourManager::TrimNeighbourLists         NeighbourManager.h:503-542
 *-----------------------------------------------------------------------------------------------*/
/* open_cell_for_gravity, src/Headers/Tree.h:413-432 */
static int open_cell_for_gravity(const orc_params* p, const kdcell* c, double drsqd, double macfactor, double amag)
{
  if (drsqd < c->cdistsqd) return 1;
  if (p->gravity_mac == 1) return drsqd < c->mac*macfactor;                               /* eigenmac */
  if (p->gravity_mac == 2) return drsqd*drsqd*amag*p->macerror < c->rmax*c->rmax*c->m;   /* gadget2  */
  return 0;
}

int orc_hydro_grav_forces_tree(const orc_params* p, const orc_tree* T, int N, const double* r,
                     const double* v, const double* m, const double* h, const double* rho,
                     const double* pressure, const double* invomega, const double* zeta,
                     const double* sound, const double* u, const double* alpha,
                     const double* hfactor, const double* hrangesqd, const int32_t* level,
                     double* a, double* atree, double* gpot, double* dudt, double* div_v,
                     double* dalphadt, int32_t* levelneib, int32_t* ndirect_out, int32_t* ncell_out)
{
  const int nd = p->ndim; const double kr = kern_range(p);
  pview P = { r, v, m, h, rho, pressure, invomega, zeta, sound, u, alpha, hfactor, hrangesqd, level };
  (void) dalphadt; (void) N;
  if (p->gravity_mac != 0 && (!T->atree && !T->gpot)) return -1;

  #pragma omp parallel
  {
    int capc = 1024, capp = 4096;
    int* gravcell = (int*) malloc(sizeof(int)*capc);
    int* cand = (int*) malloc(sizeof(int)*capp);      /* tempperneib  */
    int* tdir = (int*) malloc(sizeof(int)*capp);      /* tempdirectneib */
    int* neibl = (int*) malloc(sizeof(int)*capp);
    int* dirl = (int*) malloc(sizeof(int)*capp);
    #pragma omp for schedule(dynamic, 8)
    for (int lc = 0; lc < T->Ncell; lc++) {
      const kdcell* cell = &T->cell[lc];
      if (cell->copen != -1 || cell->N == 0) continue;
      int Ncell_l = 0, Ncand = 0, Ntdir = 0;
      const double hrangemax = kr*cell->hmax, rmax = cell->rmax;
      int cc = 0;
      while (cc < T->Ncell) {
        const kdcell* c = &T->cell[cc];
        double dr[3];
        for (int k = 0; k < nd; k++) dr[k] = c->rcell[k] - cell->rcell[k];
        const double drsqd = dot(dr, dr, nd);
        if (drsqd <= pow(c->rmax + rmax + hrangemax, 2) ||
            drsqd <= pow(rmax + c->rmax + kr*c->hmax, 2)) {
          if (c->copen != -1) cc = c->copen;
          else if (c->N == 0) cc = c->cnext;
          else {
            int i = c->ifirst;
            while (i != -1) {
              if (Ncand == capp || Ntdir == capp) {
                capp *= 2;
                cand = (int*) realloc(cand, sizeof(int)*capp); tdir = (int*) realloc(tdir, sizeof(int)*capp);
                neibl = (int*) realloc(neibl, sizeof(int)*capp); dirl = (int*) realloc(dirl, sizeof(int)*capp);
              }
              cand[Ncand++] = i;
              if (i == c->ilast) break;
              i = T->inext[i];
            }
            cc = c->cnext;
          }
        }
        else if (c->N == 0) cc = c->cnext;
        else if (!open_cell_for_gravity(p, c, drsqd, cell->macfactor, cell->amin)) {
          if (c->copen == -1 && c->N == 1) {
            if (Ntdir == capp) { capp *= 2;
              cand = (int*) realloc(cand, sizeof(int)*capp); tdir = (int*) realloc(tdir, sizeof(int)*capp);
              neibl = (int*) realloc(neibl, sizeof(int)*capp); dirl = (int*) realloc(dirl, sizeof(int)*capp); }
            tdir[Ntdir++] = c->ifirst;
          }
          else {
            if (Ncell_l == capc) { capc *= 2; gravcell = (int*) realloc(gravcell, sizeof(int)*capc); }
            gravcell[Ncell_l++] = cc;
          }
          cc = c->cnext;
        }
        else {
          if (c->copen != -1) cc = c->copen;
          else {
            int i = c->ifirst;
            while (i != -1) {
              if (Ntdir == capp) { capp *= 2;
                cand = (int*) realloc(cand, sizeof(int)*capp); tdir = (int*) realloc(tdir, sizeof(int)*capp);
                neibl = (int*) realloc(neibl, sizeof(int)*capp); dirl = (int*) realloc(dirl, sizeof(int)*capp); }
              tdir[Ntdir++] = i;
              if (i == c->ilast) break;
              i = T->inext[i];
            }
            cc = c->cnext;
          }
        }
      }
      /* _EndSearch: cell-level cull of the hydro candidates */
      int Nneibl = 0, Ndirl = 0;
      for (int ii = 0; ii < Ntdir; ii++) dirl[Ndirl++] = tdir[ii];
      const double hrangemaxsqd = pow(cell->rmax + kr*cell->hmax, 2);
      for (int ii = 0; ii < Ncand; ii++) {
        const int i = cand[ii];
        double dr[3];
        for (int k = 0; k < nd; k++) dr[k] = r[(size_t)i*nd + k] - cell->rcell[k];
        const double drsqd = dot(dr, dr, nd);
        const double h2 = rmax + kr*h[i];
        if (drsqd < hrangemaxsqd || drsqd < h2*h2) neibl[Nneibl++] = i;
        else dirl[Ndirl++] = i;
      }
      const int NCellDirect = Ndirl;

      fastmp F; memset(&F, 0, sizeof F);
      if (p->multipole >= 2) {
        for (int k = 0; k < nd; k++) F.rc[k] = cell->r[k];
        for (int c = 0; c < Ncell_l; c++) {
          if (p->multipole == 2) fast_add_monopole(nd, &F, &T->cell[gravcell[c]]);
          else fast_add_quadrupole(nd, &F, &T->cell[gravcell[c]]);
        }
      }

      /* loop over the (all active) particles of the leaf */
      int i = cell->ifirst;
      while (i != -1) {
        acc_t A; memset(&A, 0, sizeof A);
        A.gpot = (m[i]/h[i])*orc_kernel(p, 5, 0.0);
        int Nd = NCellDirect;
        /* trim + hydro/softened-gravity neighbours */
        for (int jj = 0; jj < Nneibl; jj++) {
          const int j = neibl[jj];
          double dr[3], drsqd;
          if (is_neighbour(p, r + (size_t)i*nd, r + (size_t)j*nd, hrangesqd[i], hrangesqd[j], dr, &drsqd)) {
            if (p->hydro_forces) hydro_pair(p, &P, i, j, dr, drsqd, 1, &A);
            else smoothgrav_pair(p, &P, i, j, dr, drsqd, &A);
          }
          else dirl[Nd++] = j;
        }
        for (int jj = 0; jj < Nd; jj++) {
          const int j = dirl[jj];
          double dr[3];
          for (int k = 0; k < nd; k++) dr[k] = r[(size_t)j*nd + k] - r[(size_t)i*nd + k];
          direct_pair(nd, dr, m[j], &A);
        }
        if (p->multipole == 0)
          for (int c = 0; c < Ncell_l; c++) cell_monopole(nd, &T->cell[gravcell[c]], r + (size_t)i*nd, &A);
        else if (p->multipole == 1)
          for (int c = 0; c < Ncell_l; c++) cell_quadrupole(nd, &T->cell[gravcell[c]], r + (size_t)i*nd, &A);
        double dadt;
        finish_particle(p, &P, i, 1, &A, &dadt);
        if (p->multipole >= 2) fast_apply(nd, &F, r + (size_t)i*nd, &A);
        for (int k = 0; k < nd; k++) {
          a[(size_t)i*nd + k] += A.a[k];
          a[(size_t)i*nd + k] += A.atree[k];
          atree[(size_t)i*nd + k] += A.atree[k];
        }
        gpot[i] += A.gpot; dudt[i] += A.dudt; div_v[i] += A.div_v;
        if (A.levelneib > levelneib[i]) levelneib[i] = A.levelneib;
        if (ndirect_out) ndirect_out[i] = Nd;
        if (ncell_out) ncell_out[i] = Ncell_l;
        if (i == cell->ilast) break;
        i = T->inext[i];
      }
    }
    free(gravcell); free(cand); free(tdir); free(neibl); free(dirl);
  }
  return 0;
}


/*================================================================================================
 * Gravity of the gas on the star particles: HydroTree::UpdateAllStarGasForces
 * (src/Tree/HydroTree.cpp:552-654) = per star Tree::ComputeStarGravityInteractionList
 * (src/Tree/Tree.cpp:748-885; the opening test is the GEOMETRIC one whatever gravity_mac says),
 * then NbodyHermite4::CalculateDirectHydroForces (src/Nbody/NbodyHermite4.cpp:151-247; a and gpot are
 * the same in the leapfrog classes, adot exists only for the Hermite schemes -- note the
 * single-precision powf in wkern, :194) and ComputeCellMonopoleForces / ComputeCellQuadrupoleForces
 * at the star's position.  Open boundaries only (no Ewald).  stars: rows {r[nd], v[nd], h};
 * a, adot, gpot are ADDED to.
 *===============================================================================================*/
int orc_star_gas_forces(const orc_params* p, const orc_tree* T, int Nstar, const double* stars,
                        double* a, double* adot, double* gpot, int32_t* counts /* [Nstar][3] neib, direct, cells or NULL */)
{
  const int nd = p->ndim; const double kr = kern_range(p);
  const double twopi = 6.28318530717959;                  /* as spelled in src/Headers/Constants.h:62 */
  for (int s = 0; s < Nstar; s++) {
    const double* rs = stars + (size_t) s*(2*nd + 1);
    const double* vs = rs + nd;
    const double hs = rs[2*nd];
    const double hrangemax = kr*hs;
    acc_t A; memset(&A, 0, sizeof A);
    double ad[3] = {0, 0, 0};
    int Nneib = 0, Ndirect = 0, Ngravcell = 0;
    int cc = 0;
    while (cc < T->Ncell) {
      const kdcell* c = &T->cell[cc];
      double dr[3];
      for (int k = 0; k < nd; k++) dr[k] = c->rcell[k] - rs[k];
      const double drsqd = dot(dr, dr, nd);
      int mode = 0;                                    /* 1: smoothed neighbours, 2: direct particles */
      if (drsqd < pow(0.5*hrangemax + c->rmax + 0.5*kr*c->hmax, 2)) {
        if (c->copen != -1) { cc = c->copen; continue; }
        mode = 1;
      }
      else if (drsqd > c->cdistsqd && c->N > 0) {
        if (c->copen == -1 && c->N == 1) mode = 2;
        else { if (p->multipole == 0 || p->multipole == 2) cell_monopole(nd, c, rs, &A); else cell_quadrupole(nd, c, rs, &A);
               Ngravcell++; cc = c->cnext; continue; }
      }
      else if (drsqd <= c->cdistsqd && c->N > 0) {
        if (c->copen != -1) { cc = c->copen; continue; }
        mode = 2;
      }
      else { cc = c->cnext; continue; }
      int i = c->ifirst;
      while (i != -1 && c->N > 0) {
        double d[3], dv[3];
        for (int k = 0; k < nd; k++) { d[k] = T->r[(size_t) i*nd + k] - rs[k]; dv[k] = T->v[(size_t) i*nd + k] - vs[k]; }
        const double d2 = dot(d, d, nd), drmag = sqrt(d2), invdrmag = 1.0/drmag;
        const double drdt = dot(dv, d, nd)*invdrmag;
        const double mj = T->m[i];
        if (mode == 1) {
          const double invhmean = 2.0/(hs + T->h[i]);
          const double paux = mj*invhmean*invhmean*orc_kernel(p, 4, drmag*invhmean)*invdrmag;
          const double wkern = orc_kernel(p, 0, drmag*invhmean)*(double) powf((float) invhmean, (float) nd);
          for (int k = 0; k < nd; k++) A.atree[k] += paux*d[k];
          for (int k = 0; k < nd; k++)
            ad[k] += paux*dv[k] - 3.0*paux*drdt*invdrmag*d[k] + 2.0*twopi*mj*drdt*wkern*invdrmag*d[k];
          A.gpot += mj*invhmean*orc_kernel(p, 5, drmag*invhmean);
          Nneib++;
        }
        else {
          for (int k = 0; k < nd; k++) A.atree[k] += mj*d[k]*pow(invdrmag, 3);
          for (int k = 0; k < nd; k++) ad[k] += mj*pow(invdrmag, 3)*(dv[k] - 3.0*drdt*invdrmag*d[k]);
          A.gpot += mj*invdrmag;
          Ndirect++;
        }
        if (i == c->ilast) break;
        i = T->inext[i];
      }
      cc = c->cnext;
    }
    for (int k = 0; k < nd; k++) { a[(size_t) s*nd + k] += A.atree[k]; adot[(size_t) s*nd + k] += ad[k]; }
    gpot[s] += A.gpot;
    if (counts) { counts[3*s] = Nneib; counts[3*s + 1] = Ndirect; counts[3*s + 2] = Ngravcell; }
  }
  return 0;
}
