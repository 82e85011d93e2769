/*=================================================================================================
 *  tables.cpp -- host construction of the tabulated smoothing-kernel look-up tables.
 *
 *  Follows TabulatedKernel (src/Hydrodynamics/TabulatedKernel.cpp:58-105) and
 *  initializeTable / initializeTableSqd (src/Headers/SmoothingKernel.h:585-605): 1000 samples of
 *  the analytic kernel at s = i*kernrange/1000 (w1, wgrav, wpot) or at s = sqrt(i*kernrangesqd/1000)
 *  (w0_s2, womega_s2, wzeta_s2).  Analytic forms: M4 SmoothingKernel.h:131-240, quintic :281-408,
 *  norms {M4,Quintic}Kernel.cpp:39-53.  pow() is used where the reference uses it so that the table
 *  entries carry the same rounding.
 *=================================================================================================*/
#include <cmath>
#include "../../include/gsph.h"

#define RES 1000
static const double invpi = 0.31830988618379;
static const double twothirds = 0.66666666666666666666666;

double gsph_kernnorm(int ndim, int kernel)
{
  if (kernel == GSPH_KERNEL_M4) return ndim == 1 ? twothirds : (ndim == 2 ? invpi*(10.0/7.0) : invpi);
  if (kernel == GSPH_KERNEL_GAUSSIAN)                  /* GaussianKernel.cpp:39-58 */
    return ndim == 1 ? sqrt(invpi) : (ndim == 2 ? invpi : invpi*sqrt(invpi));
  return ndim == 1 ? 1.0/120.0 : (ndim == 2 ? invpi*(7.0/478.0) : invpi*(1/120.));
}

static double m4(int which, double s, double kn, int nd)
{
  switch (which) {
  case 0:
    if (s < 1.0) return kn*(1.0 - 1.5*s*s + 0.75*s*s*s);
    if (s < 2.0) return 0.25*kn*pow(2.0 - s, 3);
    return 0.0;
  case 1:
    if (s < 1.0) return kn*(-3.0*s + 2.25*s*s);
    if (s < 2.0) return -0.75*kn*(2.0 - s)*(2.0 - s);
    return 0.0;
  case 2:
    if (s < 1.0) return kn*(-nd + 1.5*(nd + 2.0)*s*s - 0.75*(nd + 3.0)*pow(s, 3));
    if (s < 2.0) return kn*(-2.0*nd + 3.0*(nd + 1.0)*s - 1.50*(nd + 2.0)*s*s + 0.25*(nd + 3.0)*pow(s, 3));
    return 0.0;
  case 3:
    if (s < 1.0) return 1.4 - 2.0*s*s + 1.5*pow(s, 4) - 0.6*pow(s, 5);
    if (s < 2.0) return 1.6 - 4.0*s*s + 4.0*pow(s, 3) - 1.5*pow(s, 4) + 0.2*pow(s, 5);
    return 0.0;
  case 4:
    if (s < 1.0) return 1.333333333333333333333*s - 1.2*pow(s, 3) + 0.5*pow(s, 4);
    if (s < 2.0)
      return 2.6666666666666666667*s - 3.0*s*s + 1.2*pow(s, 3) - 0.166666666666666666667*pow(s, 4) -
        0.06666666666666666667/(s*s);
    return 1.0/(s*s);
  default:
    if (s < 1.0) return 1.4 - 0.666666666666666666666666*s*s + 0.3*pow(s, 4) - 0.1*pow(s, 5);
    if (s < 2.0)
      return -1.0/(15.0*s) + 1.6 - 1.33333333333333333333333333*s*s + pow(s, 3) - 0.3*pow(s, 4) +
        (1.0/30.0)*pow(s, 5);
    return 1.0/s;
  }
}

static double quintic(int which, double s, double kn, int nd)
{
  switch (which) {
  case 0:
    if (s < 1.0) return kn*(66.0 - 60.0*s*s + 30.0*pow(s, 4) - 10.0*pow(s, 5));
    if (s < 2.0) return kn*(51.0 + 75.0*s - 210.0*s*s + 150.0*pow(s, 3) - 45.0*pow(s, 4) + 5.0*pow(s, 5));
    if (s < 3.0) return kn*(243.0 - 405*s + 270.0*s*s - 90.0*pow(s, 3) + 15.0*pow(s, 4) - pow(s, 5));
    return 0.0;
  case 1:
    if (s < 1.0) return kn*(-120.0*s + 120.0*pow(s, 3) - 50.0*pow(s, 4));
    if (s < 2.0) return kn*(75.0 - 420.0*s + 450.0*s*s - 180.0*pow(s, 3) + 25.0*pow(s, 4));
    if (s < 3.0) return kn*(-405.0 + 540.0*s - 270.0*s*s + 60.0*pow(s, 3) - 5.0*pow(s, 4));
    return 0.0;
  case 2:
    if (s < 1.0)
      return kn*(-66.0*nd + 60.0*(nd + 2.0)*s*s - 30.0*(nd + 4.0)*pow(s, 4) + 10.0*(nd + 5.0)*pow(s, 5));
    if (s < 2.0)
      return kn*(-51.0*nd - 75.0*(nd + 1.0)*s + 210.0*(nd + 2.0)*s*s - 150.0*(nd + 3.0)*pow(s, 3) +
                 45.0*(nd + 4.0)*pow(s, 4) - 5.0*(nd + 5.0)*pow(s, 5));
    if (s < 3.0)
      return kn*(-243.0*nd + 405.0*(nd + 1.0)*s - 270.0*(nd + 2.0)*s*s + 90.0*(nd + 3.0)*pow(s, 3) -
                 15.0*(nd + 4.0)*pow(s, 4) + (nd + 5.0)*pow(s, 5));
    return 0.0;
  case 3:
    if (s < 1.0) return 33.0*s*s - 15.0*pow(s, 4) + 5.0*pow(s, 6) - 1.42857142857*pow(s, 7) - 34.14285714;
    if (s < 2.0)
      return 25.5*s*s + 25.0*pow(s, 3) - 52.5*pow(s, 4) + 30.0*pow(s, 5) - 7.5*pow(s, 6) +
        0.7142857143*pow(s, 7) - 33.785714286;
    if (s < 3.0)
      return 121.5*s*s - 135.0*pow(s, 3) + 67.5*pow(s, 4) - 18.0*pow(s, 5) + 2.5*pow(s, 6) -
        0.142857143*pow(s, 7) - 52.07142857;
    return 0.0;
  case 4:
    if (s < 1.0) return (12.0/359.0)*(22.0*s - 12.0*pow(s, 3) + (30.0/7.0)*pow(s, 5) - (5.0/4.0)*pow(s, 6));
    if (s < 2.0)
      return (12.0/359.0)*(17.0*s + (75.0/4.0)*s*s - 42.0*pow(s, 3) + 25.0*pow(s, 4) - (45.0/7.0)*pow(s, 5) +
                           (5.0/8.0)*pow(s, 6) + (5.0/56.0)/(s*s));
    if (s < 3.0)
      return (12.0/359.0)*(81.0*s - (405.0/4.0)*pow(s, 2) + 54.0*pow(s, 3) - 15.0*pow(s, 4) +
                           (15.0/7.0)*pow(s, 5) - (1.0/8.0)*pow(s, 6) - (507.0/56.0)/(s*s));
    return 1.0/(s*s);
  default:
    if (s < 1.0)
      return (12.0/359.0)*(-11.0*s*s + 3.0*pow(s, 4) - (5.0/7.0)*pow(s, 6) + (5.0/28.0)*pow(s, 7) + (478.0/14.0));
    if (s < 2.0)
      return (12.0/359.0)*(-(17.0/2.0)*s*s - (25.0/4.0)*pow(s, 3) + (21.0/2.0)*pow(s, 4) - 5.0*pow(s, 5) +
                           (15.0/14.0)*pow(s, 6) - (5.0/56.0)*pow(s, 7) + (473.0/14.0) + (5.0/56.0)/s);
    if (s < 3.0)
      return (12.0/359.0)*(-(81.0/2.0)*s*s + (135.0/4.0)*pow(s, 3) - (27.0/2.0)*pow(s, 4) + 3.0*pow(s, 5) -
                           (5.0/14.0)*pow(s, 6) + (1.0/56.0)*pow(s, 7) + (729.0/14.0) - (507.0/56.0)/s);
    return 1.0/s;
  }
}

/* SmoothingKernel.h:447-521: the Gaussian has no gravity softening (wzeta = wgrav = wpot = 0) */
static double gaussian(int which, double s, double kn, int nd)
{
  if (!(s < 3.0)) return 0.0;
  switch (which) {
  case 0: return kn*exp(-s*s);
  case 1: return -2.0*kn*s*exp(-s*s);
  case 2: return kn*(2.0*s*exp(-s*s) - nd*exp(-s*s));
  default: return 0.0;
  }
}

void gsph_make_tables(int ndim, int kernel, double* out)
{
  const double kr = kernel == GSPH_KERNEL_M4 ? 2.0 : 3.0, krs = kr*kr;
  const double kn = gsph_kernnorm(ndim, kernel);
  const double step = kr/RES, step2 = krs/RES;
  double (*f)(int, double, double, int) = kernel == GSPH_KERNEL_M4 ? m4 : (kernel == GSPH_KERNEL_GAUSSIAN ? gaussian : quintic);
  for (int i = 0; i < RES; i++) {
    out[0*RES + i] = f(0, sqrt(step2*i), kn, ndim);    /* w0_s2     */
    out[1*RES + i] = f(2, sqrt(step2*i), kn, ndim);    /* womega_s2 */
    out[2*RES + i] = f(3, sqrt(step2*i), kn, ndim);    /* wzeta_s2  */
    out[3*RES + i] = f(1, step*i, kn, ndim);           /* w1        */
    out[4*RES + i] = f(4, step*i, kn, ndim);           /* wgrav     */
    out[5*RES + i] = f(5, step*i, kn, ndim);           /* wpot      */
  }
}

/* host-side wgrav (which = 4) / wpot (which = 5) with the reference's look-up semantics
   (TabulatedKernel GravTableLookup / GravPotTableLookup, SmoothingKernel.h:645-662) or the analytic form:
   used by the N-body side of the device-resident main loop, where the stars are a handful */
double gsph_kernel_grav_host(int ndim, int kernel, int tabulated, int which, double s)
{
  const double kr = kernel == GSPH_KERNEL_M4 ? 2.0 : 3.0;
  const double kn = gsph_kernnorm(ndim, kernel);
  double (*f)(int, double, double, int) = kernel == GSPH_KERNEL_M4 ? m4 : (kernel == GSPH_KERNEL_GAUSSIAN ? gaussian : quintic);
  if (!tabulated) return f(which, s, kn, ndim);
  if (s >= kr) return which == 4 ? 1.0/(s*s) : (which == 5 ? 1.0/s : 0.0);
  const int idx = (int)(s*((double) RES/kr));
  return f(which, (kr/RES)*idx, kn, ndim);
}
