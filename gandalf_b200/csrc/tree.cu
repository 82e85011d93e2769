/*=================================================================================================
 *  tree.cu -- GPU build and stocking of the reference's balanced KD-tree.
 *
 *  Replaces  KDTree::BuildTree / DivideTreeCell / QuickSelect   src/Tree/KDTree.cpp:220-313,442-595,682-750
 *            KDTree::StockTree / StockCellProperties            src/Tree/KDTree.cpp:760-798,808-1083
 *            KDTree::UpdateHmaxValues                           src/Tree/KDTree.cpp:1128-1208
 *            KDTree::UpdateActiveParticleCounters               src/Tree/KDTree.cpp:1217-1254
 *            Tree::ExtrapolateCellProperties                    src/Tree/Tree.cpp:172-198
 *
 *  The reference recursion (median QuickSelect along the longest axis of the INHERITED box, cell
 *  sizes N -> N/2, N-N/2) is turned into a level-synchronous build from three pre-sorted index
 *  lists: at every level the median of a node along its split axis is simply the element at
 *  first + count/2 of that axis' list, and the other lists are brought into the children's order
 *  by a stable segmented partition (one exclusive scan + one scatter per list and level).
 *  For particle sets without coordinate ties this yields exactly the reference's leaf membership
 *  and therefore bit-identical bb/rcell/rmax/hmax/cdistsqd (the quantities the MAC reads); ties
 *  are broken by host index instead of QuickSelect's visiting order (documented in DESIGN.md).
 *=================================================================================================*/
#include "device_common.cuh"
#include <cub/cub.cuh>

namespace {

struct Bump {
  char* base; size_t off, cap;
  template <class T> T* get(size_t n)
  {
    off = (off + 255) & ~(size_t)255;
    T* p = (T*)(base + off);
    off += n*sizeof(T);
    return p;
  }
};

/* ---- root box: extent of r -+ kernrange*h, KDTree.cpp:272-279 ---------------------------------*/
template <int NDIM>
__global__ void root_box_kernel(int N, PartView Hh, double kr, double* part /*[blocks][2*NDIM]*/)
{
  double mn[3] = {GSPH_BIG, GSPH_BIG, GSPH_BIG}, mx[3] = {-GSPH_BIG, -GSPH_BIG, -GSPH_BIG};
  for (int i = blockIdx.x*blockDim.x + threadIdx.x; i < N; i += gridDim.x*blockDim.x) {
    const double hr = __dmul_rn(kr, Hh.f[F_H][i]);
#pragma unroll
    for (int k = 0; k < NDIM; k++) {
      const double x = Hh.f[F_RX + k][i];
      mx[k] = fmax(mx[k], __dadd_rn(x, hr));
      mn[k] = fmin(mn[k], __dadd_rn(x, -hr));
    }
  }
  __shared__ double sm[256];
  for (int k = 0; k < 2*NDIM; k++) {
    double v = k < NDIM ? mn[k] : mx[k - NDIM];
    sm[threadIdx.x] = v;
    __syncthreads();
    for (int s = blockDim.x/2; s > 0; s >>= 1) {
      if (threadIdx.x < s)
        sm[threadIdx.x] = k < NDIM ? fmin(sm[threadIdx.x], sm[threadIdx.x + s])
                                   : fmax(sm[threadIdx.x], sm[threadIdx.x + s]);
      __syncthreads();
    }
    if (threadIdx.x == 0) part[blockIdx.x*2*NDIM + k] = sm[0];
    __syncthreads();
  }
}

template <int NDIM>
__global__ void root_box_final(int nblocks, const double* part, double* ibmin0, double* ibmin1,
                               double* ibmin2, double* ibmax0, double* ibmax1, double* ibmax2)
{
  double* mnp[3] = {ibmin0, ibmin1, ibmin2};
  double* mxp[3] = {ibmax0, ibmax1, ibmax2};
  if (threadIdx.x < 2*NDIM) {
    const int k = threadIdx.x;
    double v = part[k];
    for (int b = 1; b < nblocks; b++)
      v = k < NDIM ? fmin(v, part[b*2*NDIM + k]) : fmax(v, part[b*2*NDIM + k]);
    if (k < NDIM) mnp[k][0] = v; else mxp[k - NDIM][0] = v;
  }
}

__global__ void iota_kernel(int N, int* a)
{
  const int i = blockIdx.x*blockDim.x + threadIdx.x;
  if (i < N) a[i] = i;
}

/* ---- permute particle fields from host order into tree order: the fields of dmask / imask are
 *      permuted, those of dfill / ifill (absent on the host: all particles hold the default) are set --*/
struct FillVals { double d[NF_DOUBLE]; int i[NF_INT]; };
__global__ void gather_kernel(int N, const int32_t* perm, PartView Hh, PartView Pp, unsigned long long dmask,
                              unsigned imask, unsigned long long dfill, unsigned ifill, FillVals fv,
                              double* __restrict__ geo4, int ndim)
{
  const int p = blockIdx.x*blockDim.x + threadIdx.x;
  if (p >= N) return;
  const int i = perm[p];
  if (geo4) {
    /* packed (x, y, z, m) record of the density gather phase and the gravity sums; dead particles get m = 0 */
    const bool dead = ((imask >> I_FLAGS) & 1u) ? ((((unsigned) Hh.i[I_FLAGS][i]) & GSPH_FLAG_DEAD) != 0) : false;
    double2 a, b;
    a.x = Hh.f[F_RX][i]; a.y = ndim > 1 ? Hh.f[F_RY][i] : 0.0;
    b.x = ndim > 2 ? Hh.f[F_RZ][i] : 0.0; b.y = dead ? 0.0 : Hh.f[F_M][i];
    reinterpret_cast<double2*>(geo4)[2*(size_t) p] = a;
    reinterpret_cast<double2*>(geo4)[2*(size_t) p + 1] = b;
  }
#pragma unroll 4
  for (int f = 0; f < NF_DOUBLE; f++) {
    if ((dmask >> f) & 1ull) Pp.f[f][p] = Hh.f[f][i];
    else if ((dfill >> f) & 1ull) Pp.f[f][p] = fv.d[f];
  }
#pragma unroll
  for (int f = 0; f < NF_INT; f++) {
    if ((imask >> f) & 1u) Pp.i[f][p] = Hh.i[f][i];
    else if ((ifill >> f) & 1u) Pp.i[f][p] = fv.i[f];
  }
}

/* The packed walk records of one node (TreeDev::nodeb / nodeh / gnode / mnode), written by the stocking kernels
 * straight from their registers (pack_nodes_kernel below does the same from the SoA arrays, for the extrapolation
 * mode).  full = 0 (hmax refresh after the density pass): only hmax / hbox changed.                          */
template <int NDIM>
__device__ __forceinline__ void pack_node_from_regs(const TreeDev& T, int n, int cnt, int full, const DevParams& dp,
                                                    const double (&bmin)[3], const double (&bmax)[3],
                                                    const double (&hbmin)[3], const double (&hbmax)[3],
                                                    const double (&rc)[3], const double (&com)[3], double m, double rmax,
                                                    double hmax, double cdist, const double (&q)[5])
{
  if (T.gnode) {
    if (full) {
      T.gnode[4*(size_t) n] = make_double2(rc[0], NDIM > 1 ? rc[1] : 0.0);
      T.gnode[4*(size_t) n + 1] = make_double2(NDIM > 2 ? rc[2] : 0.0, rmax);
      double mac = 0.0;
      if (dp.gravity_mac == GSPH_MAC_GADGET2) mac = __dmul_rn(__dmul_rn(rmax, rmax), m);
      else if (dp.gravity_mac == GSPH_MAC_EIGEN) {                   /* KDTree.cpp:1058-1079 */
        double lambda, pp;
        if (NDIM == 3) {
          pp = q[0]*q[2] - (q[0] + q[2])*(q[0] + q[2]) - q[1]*q[1] - q[3]*q[3] - q[4]*q[4];
          lambda = pp >= 0.0 ? 0.0 : 2.0*sqrt(-pp/3.0);
        }
        else if (NDIM == 2) { pp = (q[0] - q[2])*(q[0] - q[2]) + 4*q[1]*q[1]; lambda = 0.5*fmax(q[0] + q[2] + sqrt(pp), 0.); }
        else lambda = fabs(q[0]);
        mac = pow(0.5*lambda/dp.macerror, 0.66666666666666);
      }
      T.gnode[4*(size_t) n + 3] = make_double2(__hiloint2double(cnt, T.first[n]), mac);
      T.mnode[2*(size_t) n] = make_double2(com[0], NDIM > 1 ? com[1] : 0.0);
      T.mnode[2*(size_t) n + 1] = make_double2(NDIM > 2 ? com[2] : 0.0, m);
    }
    T.gnode[4*(size_t) n + 2] = make_double2(cdist, __dmul_rn(dp.kernrange, hmax));
  }
  float lo[3] = {0.f, 0.f, 0.f}, hi[3] = {0.f, 0.f, 0.f};
  if (full) {
#pragma unroll
    for (int k = 0; k < NDIM; k++) {
      lo[k] = cnt > 0 ? __double2float_rd(bmin[k]) : 3.0e38f;
      hi[k] = cnt > 0 ? __double2float_ru(bmax[k]) : -3.0e38f;
    }
    T.nodeb[2*(size_t) n] = make_float4(lo[0], lo[1], lo[2], __int_as_float(T.first[n]));
    T.nodeb[2*(size_t) n + 1] = make_float4(hi[0], hi[1], hi[2], __int_as_float(cnt));
  }
#pragma unroll
  for (int k = 0; k < NDIM; k++) {
    lo[k] = cnt > 0 ? __double2float_rd(hbmin[k]) : 3.0e38f;
    hi[k] = cnt > 0 ? __double2float_ru(hbmax[k]) : -3.0e38f;
  }
  T.nodeh[2*(size_t) n] = make_float4(lo[0], lo[1], lo[2], 0.f);
  T.nodeh[2*(size_t) n + 1] = make_float4(hi[0], hi[1], hi[2], 0.f);
}

/* ---- StockCellProperties for the leaves, KDTree.cpp:841-1010 -----------------------------------*/
template <int NDIM>
__global__ void stock_leaves_kernel(TreeDev T, PartView P, DevParams dp, int full_in)
{
  const int full = full_in & 3;
  const bool with_v = !(full_in & 4);              /* 4: the velocities are not in tree order yet (stock_vcom later) */
  const int g = blockIdx.x*blockDim.x + threadIdx.x;
  const int nleaf = 1 << T.ltot;
  if (g >= nleaf) return;
  const int n = nleaf - 1 + g;
  const int fst = T.first[n], cnt = T.count[n];
  const double kr = dp.kernrange;
  double hmax = 0.0, hbmin[3], hbmax[3];
  double m = 0.0, com[3] = {0, 0, 0}, vc[3] = {0, 0, 0}, bmin[3], bmax[3];
  int nact = 0;
#pragma unroll
  for (int k = 0; k < 3; k++) { bmin[k] = hbmin[k] = GSPH_BIG; bmax[k] = hbmax[k] = -GSPH_BIG; }
  for (int i = fst; i < fst + cnt; i++) {
    const unsigned fl = (unsigned) P.i[I_FLAGS][i];
    if (fl & GSPH_FLAG_DEAD) continue;
    if (fl & GSPH_FLAG_ACTIVE) nact++;
    const double hi = P.f[F_H][i], mi = P.f[F_M][i];
    hmax = fmax(hmax, hi);
    m += mi;
    const double hr = __dmul_rn(kr, hi);
#pragma unroll
    for (int k = 0; k < NDIM; k++) {
      const double x = P.f[F_RX + k][i];
      com[k] += mi*x;
      if (with_v) vc[k] += mi*P.f[F_VX + k][i];
      bmin[k] = fmin(bmin[k], x); bmax[k] = fmax(bmax[k], x);
      hbmin[k] = fmin(hbmin[k], __dadd_rn(x, -hr)); hbmax[k] = fmax(hbmax[k], __dadd_rn(x, hr));
    }
  }
  T.nactive[n] = nact;
  T.hmax[n] = hmax;
#pragma unroll
  for (int k = 0; k < NDIM; k++) { T.hbmin[k][n] = hbmin[k]; T.hbmax[k][n] = hbmax[k]; }
  if (!full) {
    const double z3[3] = {0, 0, 0}, z5[5] = {0, 0, 0, 0, 0};
    pack_node_from_regs<NDIM>(T, n, cnt, 0, dp, z3, z3, hbmin, hbmax, z3, z3, 0.0, 0.0, hmax, T.cdistsqd[n], z5);
    return;
  }
  if (T.macaux && full == 1) {                     /* KDTree.cpp:838-841, 899-903 */
    double aux = dp.gravity_mac == GSPH_MAC_GADGET2 ? GSPH_BIG : 0.0;
    for (int i = fst; i < fst + cnt; i++) {
      if (((unsigned) P.i[I_FLAGS][i]) & GSPH_FLAG_DEAD) continue;
      if (dp.gravity_mac == GSPH_MAC_GADGET2) {
        double at[3] = {P.f[F_ATX][i], NDIM > 1 ? P.f[F_ATY][i] : 0.0, NDIM > 2 ? P.f[F_ATZ][i] : 0.0};
        aux = fmin(aux, __dsqrt_rn(dot_exact<NDIM>(at)));
      }
      else aux = fmax(aux, pow(P.f[F_GPOT][i], -0.66666666666666666666666));
    }
    T.macaux[n] = aux;
  }
  if (m > 0.0) {
#pragma unroll
    for (int k = 0; k < NDIM; k++) { com[k] = com[k]/m; vc[k] = vc[k]/m; }
  }
  double rc[3] = {0, 0, 0}, rmax = 0.0, cdist = GSPH_BIG;
  if (cnt > 0) {
    double dr[3] = {0, 0, 0};
#pragma unroll
    for (int k = 0; k < NDIM; k++) {
      rc[k] = __dmul_rn(0.5, __dadd_rn(bmin[k], bmax[k]));
      dr[k] = __dmul_rn(0.5, __dadd_rn(bmax[k], -bmin[k]));
    }
    const double d2 = dot_exact<NDIM>(dr);
    cdist = __ddiv_rn(fmax(d2, __dmul_rn(hmax, hmax)), dp.thetamaxsqd);
    rmax = __dsqrt_rn(d2);
  }
  T.m[n] = m; T.rmax[n] = rmax; T.cdistsqd[n] = cdist;
#pragma unroll
  for (int k = 0; k < NDIM; k++) {
    T.bbmin[k][n] = bmin[k]; T.bbmax[k][n] = bmax[k];
    T.rcell[k][n] = rc[k]; T.com[k][n] = com[k];
    if (with_v) T.vcom[k][n] = vc[k];
  }
  /* quadrupole about the centre of mass, KDTree.cpp:937-1000 */
  double q[5] = {0, 0, 0, 0, 0};
  for (int i = fst; i < fst + cnt; i++) {
    if (((unsigned) P.i[I_FLAGS][i]) & GSPH_FLAG_DEAD) continue;
    const double mi = P.f[F_M][i];
    double dr[3] = {0, 0, 0};
#pragma unroll
    for (int k = 0; k < NDIM; k++) dr[k] = P.f[F_RX + k][i] - com[k];
    const double d2 = dr[0]*dr[0] + dr[1]*dr[1] + dr[2]*dr[2];
    q[0] += mi*(3.0*dr[0]*dr[0] - d2);
    if (NDIM > 1) { q[1] += mi*3.0*dr[0]*dr[1]; q[2] += mi*(3.0*dr[1]*dr[1] - d2); }
    if (NDIM > 2) { q[3] += mi*3.0*dr[2]*dr[0]; q[4] += mi*3.0*dr[2]*dr[1]; }
  }
#pragma unroll
  for (int k = 0; k < 5; k++) T.q[k][n] = q[k];
  pack_node_from_regs<NDIM>(T, n, cnt, 1, dp, bmin, bmax, hbmin, hbmax, rc, com, m, rmax, hmax, cdist, q);
}

/* ---- StockCellProperties for one internal level, KDTree.cpp:1013-1080 ---------------------------*/
template <int NDIM>
__global__ void stock_level_kernel(TreeDev T, DevParams dp, int level, int full_in)
{
  const int full = full_in & 3;
  const bool with_v = !(full_in & 4);
  const int k = blockIdx.x*blockDim.x + threadIdx.x;
  if (k >= (1 << level)) return;
  const int n = (1 << level) - 1 + k;
  const int c[2] = {2*n + 1, 2*n + 2};
  double hmax = 0.0, hbmin[3], hbmax[3], bmin[3], bmax[3];
#pragma unroll
  for (int d = 0; d < 3; d++) { bmin[d] = hbmin[d] = GSPH_BIG; bmax[d] = hbmax[d] = -GSPH_BIG; }
  const int cnt = T.count[n];
#pragma unroll
  for (int s = 0; s < 2; s++) {
    if (T.count[c[s]] > 0) {
#pragma unroll
      for (int d = 0; d < NDIM; d++) {
        hbmin[d] = fmin(T.hbmin[d][c[s]], hbmin[d]); hbmax[d] = fmax(T.hbmax[d][c[s]], hbmax[d]);
        if (full) { bmin[d] = fmin(T.bbmin[d][c[s]], bmin[d]); bmax[d] = fmax(T.bbmax[d][c[s]], bmax[d]); }
      }
      hmax = fmax(hmax, T.hmax[c[s]]);
    }
  }
  T.nactive[n] = T.nactive[c[0]] + T.nactive[c[1]];
  T.hmax[n] = hmax;
#pragma unroll
  for (int d = 0; d < NDIM; d++) { T.hbmin[d][n] = hbmin[d]; T.hbmax[d][n] = hbmax[d]; }
  if (!full) {
    const double z3[3] = {0, 0, 0}, z5[5] = {0, 0, 0, 0, 0};
    pack_node_from_regs<NDIM>(T, n, cnt, 0, dp, z3, z3, hbmin, hbmax, z3, z3, 0.0, 0.0, hmax, T.cdistsqd[n], z5);
    return;
  }
  if (T.macaux && full == 1) {                     /* KDTree.cpp:967-984 */
    const bool gad = dp.gravity_mac == GSPH_MAC_GADGET2;
    double aux = gad ? GSPH_BIG : 0.0;
#pragma unroll
    for (int s = 0; s < 2; s++)
      if (T.count[c[s]] > 0) aux = gad ? fmin(aux, T.macaux[c[s]]) : fmax(aux, T.macaux[c[s]]);
    T.macaux[n] = aux;
  }
  const double m0 = T.m[c[0]], m1 = T.m[c[1]];
  const double m = m0 + m1;
  double com[3] = {0, 0, 0}, vc[3] = {0, 0, 0};
  if (m > 0.0) {
#pragma unroll
    for (int d = 0; d < NDIM; d++) {
      com[d] = (m0*T.com[d][c[0]] + m1*T.com[d][c[1]])/m;
      if (with_v) vc[d] = (m0*T.vcom[d][c[0]] + m1*T.vcom[d][c[1]])/m;
    }
  }
  double rc[3] = {0, 0, 0}, rmax = 0.0, cdist = GSPH_BIG;
  if (cnt > 0) {
    double dr[3] = {0, 0, 0};
#pragma unroll
    for (int d = 0; d < NDIM; d++) {
      rc[d] = __dmul_rn(0.5, __dadd_rn(bmin[d], bmax[d]));
      dr[d] = __dmul_rn(0.5, __dadd_rn(bmax[d], -bmin[d]));
    }
    const double d2 = dot_exact<NDIM>(dr);
    cdist = __ddiv_rn(fmax(d2, __dmul_rn(hmax, hmax)), dp.thetamaxsqd);
    rmax = __dsqrt_rn(d2);
  }
  T.m[n] = m; T.rmax[n] = rmax; T.cdistsqd[n] = cdist;
#pragma unroll
  for (int d = 0; d < NDIM; d++) {
    T.bbmin[d][n] = bmin[d]; T.bbmax[d][n] = bmax[d];
    T.rcell[d][n] = rc[d]; T.com[d][n] = com[d];
    if (with_v) T.vcom[d][n] = vc[d];
  }
  double q[5] = {0, 0, 0, 0, 0};
#pragma unroll
  for (int s = 0; s < 2; s++) {
    const double mi = s ? m1 : m0;
    if (mi > 0.0) {
      double dr[3] = {0, 0, 0};
#pragma unroll
      for (int d = 0; d < NDIM; d++) dr[d] = T.com[d][c[s]] - com[d];
      const double d2 = dr[0]*dr[0] + dr[1]*dr[1] + dr[2]*dr[2];
#pragma unroll
      for (int j = 0; j < 5; j++) q[j] += T.q[j][c[s]];
      q[0] += mi*(3.0*dr[0]*dr[0] - d2);
      if (NDIM > 1) { q[1] += mi*3.0*dr[0]*dr[1]; q[2] += mi*(3.0*dr[1]*dr[1] - d2); }
      if (NDIM > 2) { q[3] += mi*3.0*dr[2]*dr[0]; q[4] += mi*3.0*dr[2]*dr[1]; }
    }
  }
#pragma unroll
  for (int j = 0; j < 5; j++) T.q[j][n] = q[j];
  pack_node_from_regs<NDIM>(T, n, cnt, 1, dp, bmin, bmax, hbmin, hbmax, rc, com, m, rmax, hmax, cdist, q);
}

/* Tree::ExtrapolateCellProperties, Tree.cpp:172-198 */
template <int NDIM>
__global__ void extrapolate_kernel(TreeDev T, double dt)
{
  const int n = blockIdx.x*blockDim.x + threadIdx.x;
  if (n >= T.ncell) return;
#pragma unroll
  for (int d = 0; d < NDIM; d++) {
    const double s = T.vcom[d][n]*dt;
    T.com[d][n] += s; T.rcell[d][n] += s;
    T.bbmin[d][n] += s; T.bbmax[d][n] += s; T.hbmin[d][n] += s; T.hbmax[d][n] += s;
  }
}

/* the cells' centre-of-mass velocities alone (same sums as the stocking kernels above): run when the velocities
   reached tree order after the tree was stocked (deferred permutation) and an extrapolation asks for them */
template <int NDIM>
__global__ void stock_vcom_leaves_kernel(TreeDev T, PartView P)
{
  const int g = blockIdx.x*blockDim.x + threadIdx.x;
  const int nleaf = 1 << T.ltot;
  if (g >= nleaf) return;
  const int n = nleaf - 1 + g;
  const int fst = T.first[n], cnt = T.count[n];
  double m = 0.0, vc[3] = {0, 0, 0};
  for (int i = fst; i < fst + cnt; i++) {
    if (((unsigned) P.i[I_FLAGS][i]) & GSPH_FLAG_DEAD) continue;
    const double mi = P.f[F_M][i];
    m += mi;
#pragma unroll
    for (int k = 0; k < NDIM; k++) vc[k] += mi*P.f[F_VX + k][i];
  }
#pragma unroll
  for (int k = 0; k < NDIM; k++) T.vcom[k][n] = m > 0.0 ? vc[k]/m : vc[k];
}
template <int NDIM>
__global__ void stock_vcom_level_kernel(TreeDev T, int level)
{
  const int k = blockIdx.x*blockDim.x + threadIdx.x;
  if (k >= (1 << level)) return;
  const int n = (1 << level) - 1 + k;
  const int c0 = 2*n + 1, c1 = 2*n + 2;
  const double m0 = T.m[c0], m1 = T.m[c1], m = m0 + m1;
#pragma unroll
  for (int d = 0; d < NDIM; d++) T.vcom[d][n] = m > 0.0 ? (m0*T.vcom[d][c0] + m1*T.vcom[d][c1])/m : 0.0;
}

/* outward-rounded single-precision node records for the neighbour walks (TreeDev::nodeb / nodeh) */
template <int NDIM>
__global__ void pack_nodes_kernel(TreeDev T, int with_bb, double kernrange, int gravity_mac, double macerror)
{
  const int n = blockIdx.x*blockDim.x + threadIdx.x;
  if (n >= T.ncell) return;
  const int cnt = T.count[n];
  if (T.gnode) {                                   /* gravity walk records (exact doubles) */
    if (with_bb) {
      T.gnode[4*(size_t) n] = make_double2(T.rcell[0][n], NDIM > 1 ? T.rcell[1][n] : 0.0);
      T.gnode[4*(size_t) n + 1] = make_double2(NDIM > 2 ? T.rcell[2][n] : 0.0, T.rmax[n]);
      double mac = 0.0;
      if (gravity_mac == GSPH_MAC_GADGET2) mac = __dmul_rn(__dmul_rn(T.rmax[n], T.rmax[n]), T.m[n]);
      else if (gravity_mac == GSPH_MAC_EIGEN) {                      /* KDTree.cpp:1058-1079 */
        double lambda, pp;
        const double q0 = T.q[0][n], q1 = T.q[1][n], q2 = T.q[2][n], q3 = T.q[3][n], q4 = T.q[4][n];
        if (NDIM == 3) {
          pp = q0*q2 - (q0 + q2)*(q0 + q2) - q1*q1 - q3*q3 - q4*q4;
          lambda = pp >= 0.0 ? 0.0 : 2.0*sqrt(-pp/3.0);
        }
        else if (NDIM == 2) { pp = (q0 - q2)*(q0 - q2) + 4*q1*q1; lambda = 0.5*fmax(q0 + q2 + sqrt(pp), 0.); }
        else lambda = fabs(q0);
        mac = pow(0.5*lambda/macerror, 0.66666666666666);
      }
      /* (first, count) travel as the two 32-bit halves of one double: never used in arithmetic */
      T.gnode[4*(size_t) n + 3] = make_double2(__hiloint2double(cnt, T.first[n]), mac);
      T.mnode[2*(size_t) n] = make_double2(T.com[0][n], NDIM > 1 ? T.com[1][n] : 0.0);
      T.mnode[2*(size_t) n + 1] = make_double2(NDIM > 2 ? T.com[2][n] : 0.0, T.m[n]);
    }
    /* cdistsqd keeps the value of the last stock, hmax follows the density pass (KDTree.cpp:1163-1205) */
    T.gnode[4*(size_t) n + 2] = make_double2(T.cdistsqd[n], __dmul_rn(kernrange, T.hmax[n]));
  }
  float lo[3] = {0.f, 0.f, 0.f}, hi[3] = {0.f, 0.f, 0.f};
  if (with_bb) {
#pragma unroll
    for (int k = 0; k < NDIM; k++) {
      lo[k] = cnt > 0 ? __double2float_rd(T.bbmin[k][n]) : 3.0e38f;
      hi[k] = cnt > 0 ? __double2float_ru(T.bbmax[k][n]) : -3.0e38f;
    }
    T.nodeb[2*(size_t) n] = make_float4(lo[0], lo[1], lo[2], __int_as_float(T.first[n]));
    T.nodeb[2*(size_t) n + 1] = make_float4(hi[0], hi[1], hi[2], __int_as_float(cnt));
  }
#pragma unroll
  for (int k = 0; k < NDIM; k++) {
    lo[k] = cnt > 0 ? __double2float_rd(T.hbmin[k][n]) : 3.0e38f;
    hi[k] = cnt > 0 ? __double2float_ru(T.hbmax[k][n]) : -3.0e38f;
  }
  T.nodeh[2*(size_t) n] = make_float4(lo[0], lo[1], lo[2], 0.f);
  T.nodeh[2*(size_t) n + 1] = make_float4(hi[0], hi[1], hi[2], 0.f);
}

int alloc_tree(gsph_handle* h, int ltot)
{
  TreeDev& T = h->tree;
  const int ncell = (2 << ltot) - 1;
  T.ltot = ltot; T.ncell = ncell;
  if (ncell <= T.ncell_cap) return GSPH_OK;
  auto fre = [](void* p) { if (p) cudaFree(p); };
  fre(T.first); fre(T.count); fre(T.nactive); fre(T.m); fre(T.rmax); fre(T.hmax); fre(T.cdistsqd);
  fre(T.nodeb); fre(T.nodeh); T.nodeb = T.nodeh = nullptr;
  fre(T.gnode); fre(T.mnode); T.gnode = T.mnode = nullptr;
  fre(T.macaux); T.macaux = nullptr;
  for (int k = 0; k < 3; k++) {
    fre(T.bbmin[k]); fre(T.bbmax[k]); fre(T.hbmin[k]); fre(T.hbmax[k]);
    fre(T.rcell[k]); fre(T.com[k]); fre(T.vcom[k]);
  }
  for (int k = 0; k < 5; k++) fre(T.q[k]);
  const size_t cap = (size_t) ncell + 1;
  const int nd = h->dp.ndim;
  CUDA_TRY(h, cudaMalloc(&T.first, cap*sizeof(int)));
  CUDA_TRY(h, cudaMalloc(&T.count, cap*sizeof(int)));
  CUDA_TRY(h, cudaMalloc(&T.nactive, cap*sizeof(int)));
  CUDA_TRY(h, cudaMalloc(&T.m, cap*sizeof(double)));
  CUDA_TRY(h, cudaMalloc(&T.rmax, cap*sizeof(double)));
  CUDA_TRY(h, cudaMalloc(&T.hmax, cap*sizeof(double)));
  CUDA_TRY(h, cudaMalloc(&T.cdistsqd, cap*sizeof(double)));
  CUDA_TRY(h, cudaMalloc(&T.nodeb, 2*cap*sizeof(float4)));
  CUDA_TRY(h, cudaMalloc(&T.nodeh, 2*cap*sizeof(float4)));
  if (h->dp.self_gravity && h->dp.gravity_mac != GSPH_MAC_GEOMETRIC) CUDA_TRY(h, cudaMalloc(&T.macaux, cap*sizeof(double)));
  if (h->dp.self_gravity) {
    CUDA_TRY(h, cudaMalloc(&T.gnode, 4*cap*sizeof(double2)));
    CUDA_TRY(h, cudaMalloc(&T.mnode, 2*cap*sizeof(double2)));
  }
  for (int k = 0; k < 3; k++) {
    T.bbmin[k] = T.bbmax[k] = T.hbmin[k] = T.hbmax[k] = T.rcell[k] = T.com[k] = T.vcom[k] = nullptr;
    if (k >= nd) continue;
    CUDA_TRY(h, cudaMalloc(&T.bbmin[k], cap*sizeof(double)));
    CUDA_TRY(h, cudaMalloc(&T.bbmax[k], cap*sizeof(double)));
    CUDA_TRY(h, cudaMalloc(&T.hbmin[k], cap*sizeof(double)));
    CUDA_TRY(h, cudaMalloc(&T.hbmax[k], cap*sizeof(double)));
    CUDA_TRY(h, cudaMalloc(&T.rcell[k], cap*sizeof(double)));
    CUDA_TRY(h, cudaMalloc(&T.com[k], cap*sizeof(double)));
    CUDA_TRY(h, cudaMalloc(&T.vcom[k], cap*sizeof(double)));
  }
  for (int k = 0; k < 5; k++) CUDA_TRY(h, cudaMalloc(&T.q[k], cap*sizeof(double)));
  T.ncell_cap = ncell;
  return GSPH_OK;
}

PartView view_of(double* const* f, int32_t* const* i)
{
  PartView v;
  for (int k = 0; k < NF_DOUBLE; k++) v.f[k] = f[k];
  for (int k = 0; k < NF_INT; k++) v.i[k] = i[k];
  return v;
}

template <int NDIM>
int stock_tree(gsph_handle* h, int full)
{
  TreeDev& T = h->tree;
  const PartView P = view_of(h->P, h->PI);
  const int nleaf = 1 << T.ltot;
  const int fl = full | ((full && h->v_pending) ? 4 : 0);
  if (full) h->vcom_valid = !h->v_pending;
  stock_leaves_kernel<NDIM><<<(nleaf + 127)/128, 128, 0, h->stream>>>(T, P, h->dp, fl);
  for (int l = T.ltot - 1; l >= 0; l--) {
    const int nn = 1 << l;
    stock_level_kernel<NDIM><<<(nn + 127)/128, 128, 0, h->stream>>>(T, h->dp, l, fl);
  }
  /* (the packed walk records are written by the stocking kernels themselves, pack_node_from_regs) */
  h->stats.kernel_launches += 1 + T.ltot;
  CUDA_TRY(h, cudaGetLastError());
  return GSPH_OK;
}

template <int NDIM>
int build_tree_v2(gsph_handle* h, int phase, int shard_rank, int shard_world);

template <int NDIM>
int build_tree_nd(gsph_handle* h, int mode, double dt)
{
  const int N = (int) h->N;
  const int nd = NDIM;
  cudaStream_t s = h->stream;
  const PartView Hh = view_of(h->H, h->HI);
  const PartView Pp = view_of(h->P, h->PI);
  if (mode != 0) cudaStreamWaitEvent(s, h->ev_geom, 0);   /* r, h and the integer fields of the last upload have landed
                                                             (a rebuild waits per coordinate, see build_tree_v2) */

  if (mode == 2) {                                   /* extrapolate: Tree.cpp:172-198 */
    if (!h->have_tree) { h->err = "gsph_build_tree(extrapolate) before the first rebuild"; return GSPH_ESTATE; }
    /* (gsph_upload has flushed a deferred velocity permutation) the cells move with the vcom of the last stock */
    { const int rv = gsph_ensure_vcom(h); if (rv) return rv; }
    if (!h->sorted_valid) {
      cudaStreamWaitEvent(s, h->ev_all, 0);      /* every uploaded field has landed */
      { const int rg = gsph_gather_fields(h); if (rg) return rg; }
    }
    extrapolate_kernel<NDIM><<<(h->tree.ncell + 255)/256, 256, 0, s>>>(h->tree, dt);
    pack_nodes_kernel<NDIM><<<(h->tree.ncell + 255)/256, 256, 0, s>>>(h->tree, 1, h->dp.kernrange, h->dp.gravity_mac, h->dp.macerror);
    h->stats.kernel_launches += 2;
    CUDA_TRY(h, cudaGetLastError());
    return GSPH_OK;
  }
  if (mode == 1 || mode == 3) {                      /* restock: KDTree::StockTree (3: keep amin / macfactor) */
    if (!h->have_tree) { h->err = "gsph_build_tree(restock) before the first rebuild"; return GSPH_ESTATE; }
    cudaStreamWaitEvent(s, h->ev_all, 0);      /* every uploaded field has landed */
    { const int rg = gsph_gather_fields(h); if (rg) return rg; }
    return stock_tree<NDIM>(h, mode == 3 ? 2 : 1);
  }

  /* ---- full rebuild ----------------------------------------------------------------------------*/
  return build_tree_v2<NDIM>(h, 0, 0, 1);
}

/*================================================================================================
 * Second-generation level-synchronous build.  Every per-axis list carries, next to the particle id,
 * the particle's GLOBAL RANK along the other axes (its position in their sorted lists).  Inside a
 * node the order of an axis list is the order of its global ranks, so "does this entry go to the
 * right child of a node split along kd" is simply  rank_kd(entry) >= rank_kd(median entry)  --
 * no per-particle side array, no random access inside the level loop.  Per level: one kernel per
 * node (split axis, median, child boxes), one pass counting right-going entries per 1024-entry
 * tile, one scan of the tile counts, one pass that scans inside the tile and scatters the
 * (id, ranks) entries of all axis lists to the children's order (stable segmented partition).
 *================================================================================================*/
#define TB_THREADS 256
#define TB_ITEMS 4
#define TB_TILE (TB_THREADS*TB_ITEMS)

template <int NDIM>
struct ListPtrs {
  int* ids[3][2];            /* [axis][buffer]                                                      */
  int* rk[3][2][2];          /* [axis][q][buffer]: rank along axis (axis+1+q) % NDIM               */
  const int* rank[3];        /* [axis][host index]: global rank                                     */
  const double* r[3];
  double* ib[2][2][3];       /* [buffer][min|max][axis] inherited boxes of a level                 */
  int* first; int* count;
  int8_t* splitdim; int* medrank; int* noderb;
  int* tilecnt;              /* [NDIM][ntile] right-going entries per tile, then its exclusive scan */
  int ntile;
};

__global__ void rank_scatter_kernel(int N, const int* __restrict__ list, int* __restrict__ rank)
{
  const int p = blockIdx.x*blockDim.x + threadIdx.x;
  if (p < N) rank[list[p]] = p;
}

template <int NDIM>
__global__ void fill_ranks_kernel(int N, ListPtrs<NDIM> L)
{
  const int p = blockIdx.x*blockDim.x + threadIdx.x;
  if (p >= N) return;
#pragma unroll
  for (int d = 0; d < NDIM; d++) {
    const int id = L.ids[d][0][p];
#pragma unroll
    for (int q = 0; q < NDIM - 1; q++) L.rk[d][q][0][p] = L.rank[(d + 1 + q) % NDIM][id];
  }
}

/* DivideTreeCell for every node of one level, KDTree.cpp:492-560 */
template <int NDIM>
__global__ void split_nodes_v2_kernel(int level, int cur, int klo, int khi, ListPtrs<NDIM> L)
{
  const int k = blockIdx.x*blockDim.x + threadIdx.x;
  if (k >= (1 << level)) return;
  const int n = (1 << level) - 1 + k;
  const int a = level & 1;
  const int cnt = L.count[n], fst = L.first[n];
  /* node sizes of the balanced tree do not depend on the data: every rank fills first/count (and the
     right-going counts) of ALL nodes; split axis, median and boxes only exist for the nodes of its
     own shard [klo, khi) */
  if (k < klo || k >= khi) {
    const int c1o = 2*n + 1, c2o = 2*n + 2, lefto = cnt/2;
    L.first[c1o] = fst;         L.count[c1o] = lefto;
    L.first[c2o] = fst + lefto; L.count[c2o] = cnt - lefto;
    L.noderb[k] = cnt - lefto;
    return;
  }
  int kd = 0; double rkmax = 0.0;
#pragma unroll
  for (int d = 0; d < NDIM; d++) {
    const double w = __dadd_rn(L.ib[a][1][d][k], -L.ib[a][0][d][k]);
    if (w > rkmax) { rkmax = w; kd = d; }
  }
  const int left = cnt/2;
  double rdiv = 0.0; int mr = 0x7fffffff;
  if (cnt > 0) {
    const int idm = L.ids[kd][cur][fst + left];
    rdiv = L.r[kd][idm];
    mr = L.rank[kd][idm];
  }
  const int c1 = 2*n + 1, c2 = 2*n + 2;
  L.first[c1] = fst;        L.count[c1] = left;
  L.first[c2] = fst + left; L.count[c2] = cnt - left;
#pragma unroll
  for (int d = 0; d < NDIM; d++) {
    const double mn = L.ib[a][0][d][k], mx = L.ib[a][1][d][k];
    L.ib[a ^ 1][0][d][2*k] = mn;                       L.ib[a ^ 1][1][d][2*k] = (d == kd) ? rdiv : mx;
    L.ib[a ^ 1][0][d][2*k + 1] = (d == kd) ? rdiv : mn; L.ib[a ^ 1][1][d][2*k + 1] = mx;
  }
  L.splitdim[k] = (int8_t) kd;
  L.medrank[k] = mr;
  L.noderb[k] = cnt - left;          /* exclusive-scanned over the level afterwards */
}

/* node (index inside the level) that owns list position p: the balanced tree keeps first[] within a
 * few entries of the uniform estimate, so a short correction walk from p*2^level/N finds it        */
__device__ __forceinline__ int node_of_position(int p, double nodes_per_pos, int level, const int* __restrict__ first)
{
  const int nn = 1 << level, base = nn - 1;
  int k = (int)((double) p*nodes_per_pos);
  if (k > nn - 1) k = nn - 1;
  while (k > 0 && p < first[base + k]) k--;
  while (k + 1 < nn && p >= first[base + k + 1]) k++;
  return k;
}

/* Per-thread cursor over TB_ITEMS consecutive list positions: the owning node is searched once and
 * then advanced incrementally; the node's split data are re-read only when the node changes.      */
template <int NDIM>
struct SideCursor {
  const ListPtrs<NDIM>& L; int d, cur, level, nn, base;
  int k, fst, mid, end, kd, mr;
  __device__ __forceinline__ SideCursor(const ListPtrs<NDIM>& L_, int d_, int cur_, int level_, int N, int p)
    : L(L_), d(d_), cur(cur_), level(level_), nn(1 << level_), base((1 << level_) - 1)
  {
    k = node_of_position(p, (double) nn/(double) N, level, L.first);
    load();
  }
  __device__ __forceinline__ void load()
  {
    const int n = base + k;
    fst = L.first[n]; const int cnt = L.count[n];
    mid = fst + cnt/2; end = fst + cnt;
    kd = L.splitdim[k]; mr = L.medrank[k];
  }
  __device__ __forceinline__ int goes_right(int p, int rk0, int rk1)
  {
    while (p >= end && k + 1 < nn) { k++; load(); }
    if (d == kd) return p >= mid;
    const int q = (kd - d - 1 + NDIM) % NDIM;      /* which of the carried ranks belongs to axis kd */
    return (q == 0 ? rk0 : rk1) >= mr;
  }
};

/* Blocked tile layout: every thread owns TB_ITEMS = 4 consecutive list positions, read as one
 * 16-byte vector per array (the arrays are padded to whole tiles).                                */
template <int NDIM>
__global__ void __launch_bounds__(TB_THREADS)
count_right_kernel(int N, int level, int cur, int plo, int phi, ListPtrs<NDIM> L)
{
  const int d = blockIdx.y;
  const int p0 = blockIdx.x*TB_TILE + threadIdx.x*TB_ITEMS;
  if ((int)(blockIdx.x + 1)*TB_TILE <= plo || (int) blockIdx.x*TB_TILE >= phi) {     /* tile of another shard */
    if (threadIdx.x == 0) L.tilecnt[d*L.ntile + blockIdx.x] = 0;
    return;
  }
  int c = 0;
  if (p0 < phi && p0 + TB_ITEMS > plo) {
    SideCursor<NDIM> sc(L, d, cur, level, N, p0);
    int4 r0 = make_int4(0, 0, 0, 0), r1 = make_int4(0, 0, 0, 0);
    if (NDIM > 1) r0 = *reinterpret_cast<const int4*>(L.rk[d][0][cur] + p0);
    if (NDIM > 2) r1 = *reinterpret_cast<const int4*>(L.rk[d][1][cur] + p0);
    const int a0[4] = {r0.x, r0.y, r0.z, r0.w}, a1[4] = {r1.x, r1.y, r1.z, r1.w};
#pragma unroll
    for (int i = 0; i < TB_ITEMS; i++) {
      const int p = p0 + i;
      if (p >= plo && p < phi) c += sc.goes_right(p, a0[i], a1[i]);
    }
  }
  typedef cub::BlockReduce<int, TB_THREADS> BR;
  __shared__ typename BR::TempStorage tmp;
  const int tot = BR(tmp).Sum(c);
  if (threadIdx.x == 0) L.tilecnt[d*L.ntile + blockIdx.x] = tot;
}

template <int NDIM>
__global__ void __launch_bounds__(TB_THREADS)
partition_v2_kernel(int N, int level, int cur, int plo, int phi, int klo, ListPtrs<NDIM> L)
{
  const int d = blockIdx.y;
  const int p0 = blockIdx.x*TB_TILE + threadIdx.x*TB_ITEMS;
  if ((int)(blockIdx.x + 1)*TB_TILE <= plo || (int) blockIdx.x*TB_TILE >= phi) return;   /* tile of another shard */
  int f[TB_ITEMS], kk[TB_ITEMS], md[TB_ITEMS];
  int idv[4] = {0, 0, 0, 0}, a0[4] = {0, 0, 0, 0}, a1[4] = {0, 0, 0, 0};
  int c = 0;
#pragma unroll
  for (int i = 0; i < TB_ITEMS; i++) { f[i] = 0; kk[i] = 0; md[i] = 0; }
  if (p0 < phi && p0 + TB_ITEMS > plo) {
    SideCursor<NDIM> sc(L, d, cur, level, N, p0);
    const int4 iv = *reinterpret_cast<const int4*>(L.ids[d][cur] + p0);
    idv[0] = iv.x; idv[1] = iv.y; idv[2] = iv.z; idv[3] = iv.w;
    if (NDIM > 1) { const int4 r = *reinterpret_cast<const int4*>(L.rk[d][0][cur] + p0); a0[0] = r.x; a0[1] = r.y; a0[2] = r.z; a0[3] = r.w; }
    if (NDIM > 2) { const int4 r = *reinterpret_cast<const int4*>(L.rk[d][1][cur] + p0); a1[0] = r.x; a1[1] = r.y; a1[2] = r.z; a1[3] = r.w; }
#pragma unroll
    for (int i = 0; i < TB_ITEMS; i++) {
      const int p = p0 + i;
      if (p >= plo && p < phi) { f[i] = sc.goes_right(p, a0[i], a1[i]); kk[i] = sc.k; md[i] = sc.mid; }
      c += f[i];
    }
  }
  typedef cub::BlockScan<int, TB_THREADS> BS;
  __shared__ typename BS::TempStorage tmp;
  int pre;
  BS(tmp).ExclusiveSum(c, pre);
  /* right-going entries of THIS list before the tile: scanned tile counts minus the list's base;
     the counts only cover this shard, whose first node klo starts with noderb[klo] entries before it */
  int G = L.tilecnt[d*L.ntile + blockIdx.x] - L.tilecnt[d*L.ntile] + pre + L.noderb[klo];
#pragma unroll
  for (int i = 0; i < TB_ITEMS; i++) {
    const int p = p0 + i;
    if (p >= plo && p < phi) {
      const int rb = G - L.noderb[kk[i]];              /* right-going entries before p inside the node */
      const int dest = f[i] ? md[i] + rb : p - rb;
      L.ids[d][cur ^ 1][dest] = idv[i];
      if (NDIM > 1) L.rk[d][0][cur ^ 1][dest] = a0[i];
      if (NDIM > 2) L.rk[d][1][cur ^ 1][dest] = a1[i];
      G += f[i];
    }
  }
}

/*================================================================================================
 * Bottom of the build: once a node holds at most BT_CAP particles, its whole subtree is finished by ONE thread
 * block in shared memory -- the same per-level algorithm (split axis from the inherited box, median entry of the
 * split-axis list, stable partition of the lists by global rank), but the lists are 16-bit local indices and no
 * level touches HBM.  Replaces the last 8-9 level passes (each re-reads and re-writes all three lists).
 * The result (order of list 0 below the node) is bit-identical to the level passes.
 *================================================================================================*/
#define BT_CAP 1024
#define BT_THREADS 256
#define BT_ITEMS (BT_CAP/BT_THREADS)
#define BT_MAXSPLIT 256                 /* nodes that still split at the deepest in-block level: nlev <= 9 */

template <int NDIM>
struct BottomShared {
  int id[BT_CAP];
  int rank[NDIM][BT_CAP];               /* global rank of local particle j along every axis */
  unsigned short lst[2][NDIM][BT_CAP];  /* the lists as local indices, double buffered */
  unsigned short scan[NDIM][BT_CAP];
  unsigned short nodeof[BT_CAP];
  int nfirst[2][2*BT_MAXSPLIT], ncount[2][2*BT_MAXSPLIT];
  int mr[2*BT_MAXSPLIT], mid[2*BT_MAXSPLIT];
  signed char kd[2*BT_MAXSPLIT];
  double box[2][2][NDIM][BT_MAXSPLIT];  /* [buffer][min|max][axis][node]: inherited boxes of the nodes that split */
  typename cub::BlockScan<unsigned long long, BT_THREADS>::TempStorage scan_tmp;
};

template <int NDIM>
__global__ void __launch_bounds__(BT_THREADS)
bottom_levels_kernel(int level0, int nlev, int cur, int klo, ListPtrs<NDIM> L)
{
  extern __shared__ __align__(16) unsigned char bt_smem[];
  BottomShared<NDIM>& S = *reinterpret_cast<BottomShared<NDIM>*>(bt_smem);
  typedef cub::BlockScan<unsigned long long, BT_THREADS> BS;
  const int tid = threadIdx.x;
  const int k = klo + blockIdx.x;
  const int n0 = (1 << level0) - 1 + k;
  const int fst = L.first[n0], cnt = L.count[n0];
  if (cnt <= 0) return;
  /* ---- local numbering = order of list 0; ranks along the other axes travel with the entries, the rank along
   *      axis 0 comes from the global table; lists 1, 2 find their particles by that rank (binary search) -------*/
  for (int j = tid; j < cnt; j += BT_THREADS) {
    const int id = L.ids[0][cur][fst + j];
    S.id[j] = id;
    S.rank[0][j] = L.rank[0][id];
#pragma unroll
    for (int q = 0; q < NDIM - 1; q++) S.rank[1 + q][j] = L.rk[0][q][cur][fst + j];     /* axis (0 + 1 + q) */
    S.lst[0][0][j] = (unsigned short) j;
    S.nodeof[j] = 0;
  }
  if (tid == 0) {
    S.nfirst[0][0] = 0; S.ncount[0][0] = cnt;
#pragma unroll
    for (int d = 0; d < NDIM; d++) { S.box[0][0][d][0] = L.ib[level0 & 1][0][d][k]; S.box[0][1][d][0] = L.ib[level0 & 1][1][d][k]; }
  }
  __syncthreads();
#pragma unroll
  for (int d = 1; d < NDIM; d++) {
    const int q = NDIM - 1 - d;                              /* (d + 1 + q) % NDIM == 0: the carried rank along axis 0 */
    for (int p = tid; p < cnt; p += BT_THREADS) {
      const int rx = L.rk[d][q][cur][fst + p];
      int lo = 0, hi = cnt;
      while (lo < hi) { const int m = (lo + hi) >> 1; if (S.rank[0][m] < rx) lo = m + 1; else hi = m; }
      S.lst[0][d][p] = (unsigned short) lo;
    }
  }
  __syncthreads();
  int cb = 0, bb = 0;                                       /* list buffer, node / box buffer */
  for (int s = 0; s < nlev; s++) {
    const int nn = 1 << s;
    /* ---- DivideTreeCell for every node of the level (split_nodes_v2_kernel) ---------------------------------*/
    for (int i = tid; i < nn; i += BT_THREADS) {
      const int f = S.nfirst[bb][i], c = S.ncount[bb][i];
      int kd = 0; double rkmax = 0.0;
#pragma unroll
      for (int d = 0; d < NDIM; d++) {
        const double w = __dadd_rn(S.box[bb][1][d][i], -S.box[bb][0][d][i]);
        if (w > rkmax) { rkmax = w; kd = d; }
      }
      const int left = c/2;
      double rdiv = 0.0; int mr = 0x7fffffff;
      if (c > 0) {
        const int jm = S.lst[cb][kd][f + left];
        rdiv = L.r[kd][S.id[jm]];
        mr = S.rank[kd][jm];
      }
      S.nfirst[bb ^ 1][2*i] = f;            S.ncount[bb ^ 1][2*i] = left;
      S.nfirst[bb ^ 1][2*i + 1] = f + left; S.ncount[bb ^ 1][2*i + 1] = c - left;
      if (s + 1 < nlev) {                                   /* the children split again: their inherited boxes */
#pragma unroll
        for (int d = 0; d < NDIM; d++) {
          const double mn = S.box[bb][0][d][i], mx = S.box[bb][1][d][i];
          S.box[bb ^ 1][0][d][2*i] = mn;                       S.box[bb ^ 1][1][d][2*i] = (d == kd) ? rdiv : mx;
          S.box[bb ^ 1][0][d][2*i + 1] = (d == kd) ? rdiv : mn; S.box[bb ^ 1][1][d][2*i + 1] = mx;
        }
      }
      S.kd[i] = (signed char) kd; S.mr[i] = mr; S.mid[i] = f + left;
    }
    __syncthreads();
    /* ---- stable partition of every list inside every node (count_right + partition_v2 kernels): the right-going
     *      counts of the NDIM lists ride in one 64-bit scan, 21 bits each ------------------------------------------*/
    {
      int jj[NDIM][BT_ITEMS], nd[BT_ITEMS];
      unsigned fl = 0u;
      unsigned long long c = 0ull;
#pragma unroll
      for (int i = 0; i < BT_ITEMS; i++) {
        const int p = tid*BT_ITEMS + i;
        nd[i] = 0;
        if (p < cnt) {
          nd[i] = S.nodeof[p];
          const int kd = S.kd[nd[i]], mid = S.mid[nd[i]], mr = S.mr[nd[i]];
#pragma unroll
          for (int d = 0; d < NDIM; d++) {
            jj[d][i] = S.lst[cb][d][p];
            const bool f = (d == kd) ? (p >= mid) : (S.rank[kd][jj[d][i]] >= mr);
            if (f) { fl |= 1u << (d*BT_ITEMS + i); c += 1ull << (21*d); }
          }
        }
      }
      unsigned long long pre;
      BS(S.scan_tmp).ExclusiveSum(c, pre);
#pragma unroll
      for (int i = 0; i < BT_ITEMS; i++) {
        const int p = tid*BT_ITEMS + i;
        if (p < cnt) {
#pragma unroll
          for (int d = 0; d < NDIM; d++) {
            S.scan[d][p] = (unsigned short)((pre >> (21*d)) & 0x1fffffull);
            if ((fl >> (d*BT_ITEMS + i)) & 1u) pre += 1ull << (21*d);
          }
        }
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < BT_ITEMS; i++) {
        const int p = tid*BT_ITEMS + i;
        if (p < cnt) {
          const int nf = S.nfirst[bb][nd[i]], mid = S.mid[nd[i]];
#pragma unroll
          for (int d = 0; d < NDIM; d++) {
            const int rb = (int) S.scan[d][p] - (int) S.scan[d][nf];          /* right-going entries before p in its node */
            const int dest = ((fl >> (d*BT_ITEMS + i)) & 1u) ? mid + rb : p - rb;
            S.lst[cb ^ 1][d][dest] = (unsigned short) jj[d][i];
          }
          S.nodeof[p] = (unsigned short)(2*nd[i] + (p >= mid ? 1 : 0));
        }
      }
    }
    cb ^= 1; bb ^= 1;
    __syncthreads();
  }
  /* the tree order below this node: list 0 */
  for (int p = tid; p < cnt; p += BT_THREADS) L.ids[0][cur ^ 1][fst + p] = S.id[S.lst[cb][0][p]];
}

/* level from which the bottom kernel takes over: the shallowest one whose nodes hold <= BT_CAP particles, but not
   more than 9 levels above the leaves; ltot (= no bottom kernel) for 1-D and for trees that are too small */
static int bottom_start_level(int64_t N, int ltot, int ndim)
{
  if (ndim == 1) return ltot;
  const char* e = getenv("GSPH_TREE_BOTTOM");
  if (e && atoi(e) == 0) return ltot;
  int l = 0;
  while (l < ltot && ((N + ((int64_t) 1 << l) - 1) >> l) > BT_CAP) l++;
  if (l < ltot - 9) l = ltot - 9;
  if (ltot - l < 2) return ltot;                            /* not worth a launch */
  return l;
}

template <int NDIM>
static int launch_bottom_levels(gsph_handle* h, int level0, int ltot, int cur, int klo, int nblocks, const ListPtrs<NDIM>& L, cudaStream_t s)
{
  auto kern = bottom_levels_kernel<NDIM>;
  const size_t smem = sizeof(BottomShared<NDIM>);
  CUDA_TRY(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  kern<<<nblocks, BT_THREADS, smem, s>>>(level0, ltot - level0, cur, klo, L);
  h->stats.kernel_launches++;
  CUDA_TRY(h, cudaGetLastError());
  return GSPH_OK;
}

/*================================================================================================
 * Multi-GPU build, phase 1, for NDIM > 1: a rank does not sort all particles to find its shard.
 * The top log2(world) splits on the path to its node are found by SELECTION: the median entry of a
 * node along its split axis is the (cnt/2)-th smallest (coordinate, id) pair -- a radix select over
 * the 12 bytes of (order-preserving key of the double, particle id), one histogram pass per byte,
 * ending early when one candidate is left.  (key, id) is exactly the order of the stable radix sort
 * of iota ids that the single-GPU build uses, so the shard -- and therefore the tree -- is the
 * same.  The rank then compacts its half, descends, and finally sorts and builds only its own
 * N/world particles.
 *================================================================================================*/
__device__ __forceinline__ unsigned long long sort_key_of(double x)
{
  if (x == 0.0) x = 0.0;                                   /* -0.0 sorts with +0.0 */
  const unsigned long long b = (unsigned long long) __double_as_longlong(x);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

/* first nb bytes (most significant first) of the 12-byte composite (key, id) equal to (pk, pid)? */
__device__ __forceinline__ bool prefix_match(unsigned long long k, unsigned id, int nb, unsigned long long pk, unsigned pid)
{
  if (nb == 0) return true;
  if (nb <= 8) return (k >> (64 - 8*nb)) == (pk >> (64 - 8*nb));
  if (k != pk) return false;
  if (nb >= 12) return id == pid;
  return (id >> (32 - 8*(nb - 8))) == (pid >> (32 - 8*(nb - 8)));
}

/* histogram of byte nb of the composite over the candidates that match the prefix; cand == nullptr: ids 0..cnt-1 */
__global__ void select_hist_kernel(int cnt, const int* __restrict__ cand, const double* __restrict__ r, int nb,
                                   unsigned long long pk, unsigned pid, unsigned* __restrict__ hist)
{
  __shared__ unsigned sh[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  for (int i = blockIdx.x*blockDim.x + threadIdx.x; i < cnt; i += gridDim.x*blockDim.x) {
    const unsigned id = cand ? (unsigned) cand[i] : (unsigned) i;
    const unsigned long long k = sort_key_of(r[id]);
    if (prefix_match(k, id, nb, pk, pid)) {
      const unsigned dig = nb < 8 ? (unsigned)((k >> (56 - 8*nb)) & 255ull) : ((id >> (24 - 8*(nb - 8))) & 255u);
      const unsigned peers = __match_any_sync(__activemask(), dig);
      if ((int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&sh[dig], (unsigned) __popc(peers));
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 256; i += blockDim.x) if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

/* the (single) candidate that matches the first nb bytes */
__global__ void select_find_kernel(int cnt, const int* __restrict__ cand, const double* __restrict__ r, int nb,
                                   unsigned long long pk, unsigned pid, int* __restrict__ out_id, double* __restrict__ out_r)
{
  for (int i = blockIdx.x*blockDim.x + threadIdx.x; i < cnt; i += gridDim.x*blockDim.x) {
    const unsigned id = cand ? (unsigned) cand[i] : (unsigned) i;
    const double x = r[id];
    if (prefix_match(sort_key_of(x), id, nb, pk, pid)) { *out_id = (int) id; *out_r = x; }
  }
}

/* which side of the median entry (ks, ids) a particle lies on */
struct SideOfMedian {
  const double* r; unsigned long long ks; unsigned ids; bool right;
  __device__ __forceinline__ bool operator()(const int& id) const
  {
    const unsigned long long k = sort_key_of(r[id]);
    const bool ge = k > ks || (k == ks && (unsigned) id >= ids);
    return ge == right;
  }
};

__global__ void gather_keys_kernel(int n, const int* __restrict__ ids, const double* __restrict__ r, double* __restrict__ out)
{
  const int i = blockIdx.x*blockDim.x + threadIdx.x;
  if (i < n) out[i] = r[ids[i]];
}

/* first / count of EVERY node of the balanced tree: sizes N -> (N/2, N - N/2) do not depend on the data */
__global__ void fill_topology_kernel(int N, int ltot, int* __restrict__ first, int* __restrict__ count)
{
  const int n = blockIdx.x*blockDim.x + threadIdx.x;
  if (n >= (2 << ltot) - 1) return;
  int l = 31 - __clz(n + 1);
  const int k = n + 1 - (1 << l);
  int a = 0, b = N;
  for (int bit = l - 1; bit >= 0; bit--) {
    const int left = (b - a)/2;
    if ((k >> bit) & 1) a += left; else b = a + left;
  }
  first[n] = a; count[n] = b - a;
}

/* phase 0: whole build on this GPU.  phase 1 / 2: multi-GPU build -- every rank sorts and splits the
 * top log2(world) levels (they need all particles), then re-orders only the lists of ITS subtree
 * (1/world of the per-level work); phase 1 ends with this rank's slice of the tree order in
 * h->perm[lo, hi).  The host layer all-gathers the slices (NCCL) and phase 2 permutes the particle
 * fields and stocks the tree.  The shard of rank r is the r-th node of level log2(world): the
 * reference's median-split KD domains (MpiKDTreeDecomposition.cpp:56-277).                         */
template <int NDIM>
int build_tree_v2(gsph_handle* h, int phase, int shard_rank, int shard_world)
{
  const int N = (int) h->N;
  if (phase == 2) {
    cudaStream_t s2 = h->stream;
    cudaEventRecord(h->evk[4][0], s2);
    cudaStreamWaitEvent(s2, h->ev_all, 0);      /* every uploaded field has landed */
    { const int rg = gsph_gather_fields(h); if (rg) return rg; }
    h->have_tree = true;
    cudaEventRecord(h->evk[4][1], s2); h->evk_used[4] = true;
    cudaEventRecord(h->evk[5][0], s2);
    const int rc2 = stock_tree<NDIM>(h, 1);
    cudaEventRecord(h->evk[5][1], s2); h->evk_used[5] = true;
    return rc2;
  }
  constexpr int nd = NDIM;
  constexpr int NR = NDIM - 1;
  cudaStream_t s = h->stream;
  const PartView Hh = view_of(h->H, h->HI);
  const PartView Pp = view_of(h->P, h->PI);

  int ltot = 0;                                       /* ComputeTreeSize, KDTree.cpp:322-352 */
  while ((int64_t) h->cfg.Nleafmax*((int64_t)1 << ltot) < (int64_t) N) ltot++;
  int rc = alloc_tree(h, ltot);
  if (rc) return rc;
  TreeDev& T = h->tree;
  const int nleaf = 1 << ltot;
  const int ntile = (N + TB_TILE - 1)/TB_TILE;
  const size_t Npad = (size_t) ntile*TB_TILE;

  /* scratch layout */
  size_t cub_sort = 0, cub_scan = 0, cub_scan2 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cub_sort, (const double*) nullptr, (double*) nullptr,
                                  (const int*) nullptr, (int*) nullptr, N, 0, 64, s);
  cub::DeviceScan::ExclusiveSum(nullptr, cub_scan, (const int*) nullptr, (int*) nullptr, nd*ntile, s);
  cub::DeviceScan::ExclusiveSum(nullptr, cub_scan2, (const int*) nullptr, (int*) nullptr, nleaf, s);
  size_t cub_bytes = cub_sort > cub_scan ? cub_sort : cub_scan;
  if (cub_scan2 > cub_bytes) cub_bytes = cub_scan2;
  const bool by_selection = phase == 1 && nd > 1;     /* multi-GPU shard found by selection, see above */
  size_t extra = 0;
  if (by_selection) {
    size_t cub_sel = 0;
    cub::DeviceSelect::If(nullptr, cub_sel, cub::CountingInputIterator<int>(0), (int*) nullptr, (int*) nullptr, N,
                          SideOfMedian{nullptr, 0ull, 0u, false}, s);
    if (cub_sel > cub_bytes) cub_bytes = cub_sel;
    extra = 2*Npad*sizeof(int) + Npad*sizeof(double) + (size_t) 4*nleaf*sizeof(int) + 8192;
  }
  const size_t need = Npad*sizeof(int)*(size_t)(nd*2*(1 + NR) + nd) + Npad*sizeof(double)
                      + (size_t) nleaf*(sizeof(double)*4*nd + 1 + 2*sizeof(int)) + (size_t) nd*ntile*sizeof(int)
                      + cub_bytes + 256*64 + 4096*2*nd*sizeof(double) + extra;
  if (need > h->scratch_bytes) {
    if (h->scratch) cudaFree(h->scratch);
    h->scratch = nullptr; h->scratch_bytes = 0;
    CUDA_TRY(h, cudaMalloc(&h->scratch, need + need/16));
    h->scratch_bytes = need + need/16;
  }
  Bump B{(char*) h->scratch, 0, h->scratch_bytes};
  ListPtrs<NDIM> L;
  int* rank[3] = {nullptr, nullptr, nullptr};
  for (int d = 0; d < 3; d++) {
    for (int b = 0; b < 2; b++) {
      L.ids[d][b] = d < nd ? B.get<int>(Npad) : nullptr;
      for (int q = 0; q < 2; q++) L.rk[d][q][b] = (d < nd && q < NR) ? B.get<int>(Npad) : nullptr;
    }
    rank[d] = d < nd ? B.get<int>(Npad) : nullptr;
    L.rank[d] = rank[d];
    L.r[d] = d < nd ? h->H[F_RX + d] : nullptr;
  }
  double* keys_out = B.get<double>(Npad);
  for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++) for (int d = 0; d < 3; d++)
    L.ib[a][b][d] = d < nd ? B.get<double>(nleaf) : nullptr;
  L.splitdim = B.get<int8_t>(nleaf);
  L.medrank = B.get<int>(nleaf);
  L.noderb = B.get<int>(nleaf);
  L.tilecnt = B.get<int>((size_t) nd*ntile);
  L.ntile = ntile;
  L.first = T.first; L.count = T.count;
  const int rb_blocks = 1024;
  double* rbpart = B.get<double>((size_t) rb_blocks*2*nd);
  void* cubtmp = B.get<char>(cub_bytes);
  int* candbuf[2] = {nullptr, nullptr}; double* keys_in = nullptr; int* lfirst = nullptr; int* lcount = nullptr;
  unsigned* hist = nullptr; int* sel_i = nullptr; double* sel_r = nullptr;
  if (by_selection) {
    candbuf[0] = B.get<int>(Npad); candbuf[1] = B.get<int>(Npad);
    keys_in = B.get<double>(Npad);
    lfirst = B.get<int>((size_t) 2*nleaf); lcount = B.get<int>((size_t) 2*nleaf);
    hist = B.get<unsigned>(256); sel_i = B.get<int>(2); sel_r = B.get<double>(1);
  }
  if (B.off > B.cap) { h->err = "internal: tree scratch layout overflow"; return GSPH_ENOMEM; }

  if (by_selection) {
    int ltop = 0;
    while ((1 << ltop) < shard_world) ltop++;
    cudaEventRecord(h->evk[3][0], s);
    for (int d = 0; d < nd; d++) cudaStreamWaitEvent(s, h->ev_axis[d], 0);
    cudaStreamWaitEvent(s, h->ev_geom, 0);
    root_box_kernel<NDIM><<<rb_blocks, 256, 0, s>>>(N, Hh, h->dp.kernrange, rbpart);
    root_box_final<NDIM><<<1, 32, 0, s>>>(rb_blocks, rbpart, L.ib[0][0][0], L.ib[0][0][1], L.ib[0][0][2],
                                          L.ib[0][1][0], L.ib[0][1][1], L.ib[0][1][2]);
    h->stats.kernel_launches += 2;
    double box[2][3] = {{0, 0, 0}, {0, 0, 0}};
    for (int b = 0; b < 2; b++) for (int d = 0; d < nd; d++)
      CUDA_TRY(h, cudaMemcpyAsync(&box[b][d], L.ib[0][b][d], sizeof(double), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(h, cudaStreamSynchronize(s));
    /* descend the path to this rank's node of level ltop */
    const int* cand = nullptr;                         /* nullptr: all particles, ids 0..N-1 */
    int cnt = N, slo = 0, which = 0;
    for (int l = 0; l < ltop; l++) {
      int kd = 0; double rkmax = 0.0;                   /* split axis: widest side of the inherited box, KDTree.cpp:503-512 */
      for (int d = 0; d < nd; d++) { const double w = box[1][d] - box[0][d]; if (w > rkmax) { rkmax = w; kd = d; } }
      const int left = cnt/2;
      const double* rk = h->H[F_RX + kd];
      unsigned long long pk = 0ull; unsigned pid = 0u; int nb = 0; unsigned target = (unsigned) left;
      bool single = false;
      while (nb < 12 && !single) {
        CUDA_TRY(h, cudaMemsetAsync(hist, 0, 256*sizeof(unsigned), s));
        select_hist_kernel<<<h->sm_count*8, 256, 0, s>>>(cnt, cand, rk, nb, pk, pid, hist);
        unsigned hh[256];
        CUDA_TRY(h, cudaMemcpyAsync(hh, hist, sizeof hh, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(h, cudaStreamSynchronize(s));
        h->stats.kernel_launches++;
        int bin = 0;
        while (bin < 255 && target >= hh[bin]) { target -= hh[bin]; bin++; }
        if (target >= hh[bin]) { h->err = "internal: median selection lost its target"; return GSPH_ECUDA; }
        if (nb < 8) pk |= (unsigned long long) bin << (56 - 8*nb);
        else pid |= (unsigned) bin << (24 - 8*(nb - 8));
        nb++;
        single = hh[bin] == 1;
      }
      select_find_kernel<<<h->sm_count*8, 256, 0, s>>>(cnt, cand, rk, nb, pk, pid, sel_i, sel_r);
      int idm = 0; double rdiv = 0.0;
      CUDA_TRY(h, cudaMemcpyAsync(&idm, sel_i, sizeof(int), cudaMemcpyDeviceToHost, s));
      CUDA_TRY(h, cudaMemcpyAsync(&rdiv, sel_r, sizeof(double), cudaMemcpyDeviceToHost, s));
      CUDA_TRY(h, cudaStreamSynchronize(s));
      unsigned long long ks; { double x = rdiv == 0.0 ? 0.0 : rdiv; unsigned long long b; memcpy(&b, &x, 8); ks = (b >> 63) ? ~b : (b | 0x8000000000000000ull); }
      const bool right = (shard_rank >> (ltop - 1 - l)) & 1;
      int* out = candbuf[which];
      size_t tb = cub_bytes;
      const SideOfMedian op{rk, ks, (unsigned) idm, right};
      if (cand) cub::DeviceSelect::If(cubtmp, tb, cand, out, sel_i + 1, cnt, op, s);
      else cub::DeviceSelect::If(cubtmp, tb, cub::CountingInputIterator<int>(0), out, sel_i + 1, cnt, op, s);
      int got = 0;
      CUDA_TRY(h, cudaMemcpyAsync(&got, sel_i + 1, sizeof(int), cudaMemcpyDeviceToHost, s));
      CUDA_TRY(h, cudaStreamSynchronize(s));
      h->stats.kernel_launches += 3;
      const int want = right ? cnt - left : left;
      if (got != want) { h->err = "internal: shard selection size mismatch"; return GSPH_ECUDA; }
      if (right) { slo += left; box[0][kd] = rdiv; } else box[1][kd] = rdiv;
      cand = out; cnt = got; which ^= 1;
    }
    /* the local build: this shard's particles only, as a tree of ltot - ltop levels */
    const int Nloc = cnt, lloc = ltot - ltop;
    for (int b = 0; b < 2; b++) for (int d = 0; d < nd; d++)
      CUDA_TRY(h, cudaMemcpyAsync(L.ib[0][b][d], &box[b][d], sizeof(double), cudaMemcpyHostToDevice, s));
    for (int d = 0; d < nd; d++) {
      gather_keys_kernel<<<(Nloc + 255)/256, 256, 0, s>>>(Nloc, cand, h->H[F_RX + d], keys_in);
      size_t tb = cub_bytes;
      cub::DeviceRadixSort::SortPairs(cubtmp, tb, (const double*) keys_in, keys_out, cand, L.ids[d][0], Nloc, 0, 64, s);
      rank_scatter_kernel<<<(Nloc + 255)/256, 256, 0, s>>>(Nloc, L.ids[d][0], rank[d]);
      h->stats.kernel_launches += 10;
    }
    fill_ranks_kernel<NDIM><<<(Nloc + 255)/256, 256, 0, s>>>(Nloc, L);
    h->stats.kernel_launches++;
    cudaEventRecord(h->evk[3][1], s); h->evk_used[3] = true;
    cudaEventRecord(h->evk[4][0], s);
    L.first = lfirst; L.count = lcount;
    const int ntile_loc = (Nloc + TB_TILE - 1)/TB_TILE;
    L.ntile = ntile_loc;
    {
      const int one[2] = {0, Nloc};
      CUDA_TRY(h, cudaMemcpyAsync(lfirst, &one[0], sizeof(int), cudaMemcpyHostToDevice, s));
      CUDA_TRY(h, cudaMemcpyAsync(lcount, &one[1], sizeof(int), cudaMemcpyHostToDevice, s));
    }
    int cur = 0;
    const int lbot = bottom_start_level(Nloc, lloc, nd);   /* levels >= lbot: one block per node, in shared memory */
    for (int l = 0; l < lloc; l++) {
      if (l == lbot) {
        const int rb = launch_bottom_levels<NDIM>(h, lbot, lloc, cur, 0, 1 << lbot, L, s); if (rb) return rb;
        cur ^= 1;
        break;
      }
      const int nn = 1 << l;
      split_nodes_v2_kernel<NDIM><<<(nn + 127)/128, 128, 0, s>>>(l, cur, 0, nn, L);
      size_t tb = cub_bytes;
      cub::DeviceScan::ExclusiveSum(cubtmp, tb, L.noderb, L.noderb, nn, s);
      const dim3 grid(ntile_loc, nd);
      count_right_kernel<NDIM><<<grid, TB_THREADS, 0, s>>>(Nloc, l, cur, 0, Nloc, L);
      tb = cub_bytes;
      cub::DeviceScan::ExclusiveSum(cubtmp, tb, L.tilecnt, L.tilecnt, nd*ntile_loc, s);
      partition_v2_kernel<NDIM><<<grid, TB_THREADS, 0, s>>>(Nloc, l, cur, 0, Nloc, 0, L);
      cur ^= 1;
      h->stats.kernel_launches += 7;
    }
    fill_topology_kernel<<<(T.ncell + 255)/256, 256, 0, s>>>(N, ltot, T.first, T.count);
    h->stats.kernel_launches++;
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, cudaMemcpyAsync(h->perm + slo, L.ids[0][cur], sizeof(int)*(size_t) Nloc, cudaMemcpyDeviceToDevice, s));
    cudaEventRecord(h->evk[4][1], s); h->evk_used[4] = true;
    h->have_tree = false; h->sorted_valid = false;
    return GSPH_OK;
  }

  /* the presorted lists: the sort of an axis starts as soon as that coordinate of the upload has landed */
  cudaEventRecord(h->evk[3][0], s);
  for (int d = 0; d < nd; d++) {
    cudaStreamWaitEvent(s, h->ev_axis[d], 0);
    iota_kernel<<<(N + 255)/256, 256, 0, s>>>(N, L.ids[d][1]);
    size_t tb = cub_bytes;
    cub::DeviceRadixSort::SortPairs(cubtmp, tb, (const double*) h->H[F_RX + d], keys_out,
                                    (const int*) L.ids[d][1], L.ids[d][0], N, 0, 64, s);
    h->stats.kernel_launches += 8;
    if (nd > 1) { rank_scatter_kernel<<<(N + 255)/256, 256, 0, s>>>(N, L.ids[d][0], rank[d]); h->stats.kernel_launches++; }
  }
  /* root box = extent of r -+ kernrange*h: needs h (and every coordinate) */
  cudaStreamWaitEvent(s, h->ev_geom, 0);
  root_box_kernel<NDIM><<<rb_blocks, 256, 0, s>>>(N, Hh, h->dp.kernrange, rbpart);
  root_box_final<NDIM><<<1, 32, 0, s>>>(rb_blocks, rbpart, L.ib[0][0][0], L.ib[0][0][1], L.ib[0][0][2],
                                        L.ib[0][1][0], L.ib[0][1][1], L.ib[0][1][2]);
  h->stats.kernel_launches += 2;
  if (nd > 1) { fill_ranks_kernel<NDIM><<<(N + 255)/256, 256, 0, s>>>(N, L); h->stats.kernel_launches++; }
  cudaEventRecord(h->evk[3][1], s); h->evk_used[3] = true;
  cudaEventRecord(h->evk[4][0], s);
  {
    const int one[2] = {0, N};
    CUDA_TRY(h, cudaMemcpyAsync(T.first, &one[0], sizeof(int), cudaMemcpyHostToDevice, s));
    CUDA_TRY(h, cudaMemcpyAsync(T.count, &one[1], sizeof(int), cudaMemcpyHostToDevice, s));
  }
  /* the shard of this rank below level ltop: node `shard_rank` of that level */
  int ltop = 0;
  while ((1 << ltop) < shard_world) ltop++;
  int slo = 0, shi = N;
  for (int b = ltop - 1; b >= 0; b--) {                /* node sizes N -> (N/2, N - N/2), KDTree.cpp:545-560 */
    const int cnt = shi - slo, left = cnt/2;
    if ((shard_rank >> b) & 1) slo += left; else shi = slo + left;
  }
  int cur = 0;
  const int lbot = bottom_start_level(N, ltot, nd);        /* levels >= lbot: one block per node, in shared memory */
  for (int l = 0; l < ltot; l++) {
    const int nn = 1 << l;
    const bool mine_only = phase == 1 && l >= ltop;
    const int klo = mine_only ? shard_rank << (l - ltop) : 0;
    const int khi = mine_only ? (shard_rank + 1) << (l - ltop) : nn;
    if (l == lbot) {
      const int rb = launch_bottom_levels<NDIM>(h, lbot, ltot, cur, klo, khi - klo, L, s); if (rb) return rb;
      cur ^= 1;
      fill_topology_kernel<<<(T.ncell + 255)/256, 256, 0, s>>>(N, ltot, T.first, T.count);   /* first / count of the levels below */
      h->stats.kernel_launches++;
      break;
    }
    const int plo = mine_only ? slo : 0, phi = mine_only ? shi : N;
    split_nodes_v2_kernel<NDIM><<<(nn + 127)/128, 128, 0, s>>>(l, cur, klo, khi, L);
    h->stats.kernel_launches++;
    if (nd == 1) continue;                             /* one sorted list: nothing to re-order */
    size_t tb = cub_bytes;
    cub::DeviceScan::ExclusiveSum(cubtmp, tb, L.noderb, L.noderb, nn, s);
    const dim3 grid(ntile, nd);
    count_right_kernel<NDIM><<<grid, TB_THREADS, 0, s>>>(N, l, cur, plo, phi, L);
    tb = cub_bytes;
    cub::DeviceScan::ExclusiveSum(cubtmp, tb, L.tilecnt, L.tilecnt, nd*ntile, s);
    partition_v2_kernel<NDIM><<<grid, TB_THREADS, 0, s>>>(N, l, cur, plo, phi, klo, L);
    cur ^= 1;
    h->stats.kernel_launches += 6;
  }
  CUDA_TRY(h, cudaGetLastError());
  if (phase == 1) {
    /* this rank's slice of the tree order; the rest arrives through the host layer's all-gather */
    CUDA_TRY(h, cudaMemcpyAsync(h->perm + slo, L.ids[0][cur] + slo, sizeof(int)*(size_t)(shi - slo), cudaMemcpyDeviceToDevice, s));
    cudaEventRecord(h->evk[4][1], s); h->evk_used[4] = true;
    h->have_tree = false; h->sorted_valid = false;
    return GSPH_OK;
  }
  CUDA_TRY(h, cudaMemcpyAsync(h->perm, L.ids[0][cur], sizeof(int)*(size_t) N, cudaMemcpyDeviceToDevice, s));
  /* single-device rebuild straight after an upload: the velocities (uploaded last) are permuted when the force
     pass asks for them -- the build and the density pass do not read them (the cd2010 switch does) */
  const bool defer_v = phase == 0 && h->upload_pending && !h->integ.on && !h->integ.starting && h->dp.tdavisc != GSPH_TDAVISC_CD2010;
  cudaStreamWaitEvent(s, defer_v ? h->ev_early : h->ev_all, 0);      /* the uploaded fields have landed */
  { const int rg = gsph_gather_fields(h, defer_v); if (rg) return rg; }
  h->have_tree = true;
  cudaEventRecord(h->evk[4][1], s); h->evk_used[4] = true;
  cudaEventRecord(h->evk[5][0], s);
  const int rcs = stock_tree<NDIM>(h, 1);
  cudaEventRecord(h->evk[5][1], s); h->evk_used[5] = true;
  return rcs;
}

}  // namespace

/* Host order -> tree order.  When every particle is active and alive in this step (no flags from the
 * host, or the device-resident loop with one global time step) the passes overwrite their outputs for
 * every particle, so only the INPUTS of the step are permuted: r, v, m, h, u, alpha, level, sinkid, the
 * MAC / switch inputs of the previous step where the configuration reads them, and accumulators the
 * host supplied.  Otherwise (block time steps, dead particles) every field is permuted, because the
 * old values of inactive particles are inputs.                                                       */
int gsph_gather_fields(gsph_handle* h, bool defer_v)
{
  const int N = (int) h->N;
  const int nd = h->dp.ndim;
  h->v_pending = false;
  /* inside a step of the resident loop every particle was just made active (one global time step); a
     BuildTree between steps sees the flags EndTimestep left (inactive): everything is permuted then */
  const bool integ = h->integ.on && h->integ.mid_step && h->integ.cfg.Nlevels <= 1;
  const unsigned long long ALLD = (1ull << NF_DOUBLE) - 1ull; const unsigned ALLI = (1u << NF_INT) - 1u;
  unsigned long long dmask, dfill = 0ull; unsigned imask, ifill = 0u;
  h->lazy_fields = integ || (!h->integ.on && !h->integ.starting && !((h->present_i >> I_FLAGS) & 1u));
  if (!h->lazy_fields) { dmask = ALLD; imask = ALLI; }
  else {
    unsigned long long core = GSPH_DMASK(F_M) | GSPH_DMASK(F_H) | GSPH_DMASK(F_U) | GSPH_DMASK(F_ALPHA);
    for (int k = 0; k < nd; k++) core |= GSPH_DMASK(F_RX + k) | GSPH_DMASK(F_VX + k);
    if (h->dp.tdavisc == GSPH_TDAVISC_MM97) core |= GSPH_DMASK(F_DALPHADT);          /* accumulated across steps */
    if (h->dp.tdavisc == GSPH_TDAVISC_CD2010) for (int k = 0; k < nd; k++) core |= GSPH_DMASK(F_AX + k);
    if (h->dp.self_gravity && h->dp.gravity_mac != GSPH_MAC_GEOMETRIC) {              /* KDTree.cpp:899-903 */
      core |= GSPH_DMASK(F_GPOT);
      for (int k = 0; k < nd; k++) core |= GSPH_DMASK(F_ATX + k);
    }
    if (h->dp.create_sinks) core |= GSPH_DMASK(F_GPOT);                               /* potmin test, GradhSph.cpp:270-280 */
    unsigned icore = (1u << I_FLAGS) | (1u << I_LEVEL) | (1u << I_SINKID) | (1u << I_PTYPE);
    /* accumulators: the force passes add to them; what the host supplied counts (the resident loop always
       zeroes them, gsph_zero_accelerations) */
    unsigned long long acc = 0ull; unsigned iacc = 0u;
    if (!integ) {
      acc = GSPH_OUT_FORCES & h->present_d; iacc = (1u << I_LEVELNEIB) & h->present_i;
    }
    const unsigned long long pres_d = integ ? ALLD : h->present_d; const unsigned pres_i = integ ? ALLI : h->present_i;
    dmask = (core | acc) & pres_d; imask = (icore | iacc) & pres_i;
    dfill = core & ~pres_d; ifill = icore & ~pres_i;
  }
  if (defer_v && h->lazy_fields) {
    unsigned long long vmask = 0ull;
    for (int k = 0; k < nd; k++) vmask |= GSPH_DMASK(F_VX + k);
    if ((dmask & vmask) == vmask) { dmask &= ~vmask; h->v_pending = true; }
  }
  FillVals fv;
  for (int f = 0; f < NF_DOUBLE; f++) fv.d[f] = (f == F_H || f == F_INVOMEGA) ? 1.0 : 0.0;
  for (int f = 0; f < NF_INT; f++) fv.i[f] = f == I_FLAGS ? (int) GSPH_FLAG_ACTIVE : ((f == I_SINKID || f == I_IORIG) ? -1 : 0);
  gather_kernel<<<(N + 255)/256, 256, 0, h->stream>>>(N, h->perm, view_of(h->H, h->HI), view_of(h->P, h->PI),
                                                     dmask, imask, dfill, ifill, fv, h->geo4, nd);
  h->stats.kernel_launches++;
  CUDA_TRY(h, cudaGetLastError());
  h->pvalid_d = dmask | dfill; h->pvalid_i = imask | ifill;
  h->sorted_valid = true; h->acc_zeroed = false;
  return GSPH_OK;
}

/* the deferred part of the permutation: the velocities of the last upload into tree order */
int gsph_ensure_velocities(gsph_handle* h)
{
  if (!h->v_pending) return GSPH_OK;
  const int N = (int) h->N;
  unsigned long long vmask = 0ull;
  for (int k = 0; k < h->dp.ndim; k++) vmask |= GSPH_DMASK(F_VX + k);
  CUDA_TRY(h, cudaStreamWaitEvent(h->stream, h->ev_all, 0));
  FillVals fv;
  for (int f = 0; f < NF_DOUBLE; f++) fv.d[f] = 0.0;
  for (int f = 0; f < NF_INT; f++) fv.i[f] = 0;
  gather_kernel<<<(N + 255)/256, 256, 0, h->stream>>>(N, h->perm, view_of(h->H, h->HI), view_of(h->P, h->PI),
                                                     vmask, 0u, 0ull, 0u, fv, (double*) nullptr, h->dp.ndim);
  h->stats.kernel_launches++;
  CUDA_TRY(h, cudaGetLastError());
  h->pvalid_d |= vmask;
  h->v_pending = false;
  return GSPH_OK;
}

/* the cells' centre-of-mass velocities, if the last stock ran before the velocities were in tree order */
int gsph_ensure_vcom(gsph_handle* h)
{
  int rc = gsph_ensure_velocities(h); if (rc) return rc;
  if (h->vcom_valid || !h->have_tree) return GSPH_OK;
  const PartView Pv = view_of(h->P, h->PI);
  const int nleaf = 1 << h->tree.ltot;
  cudaStream_t s = h->stream;
  switch (h->dp.ndim) {
  case 1: stock_vcom_leaves_kernel<1><<<(nleaf + 127)/128, 128, 0, s>>>(h->tree, Pv); break;
  case 2: stock_vcom_leaves_kernel<2><<<(nleaf + 127)/128, 128, 0, s>>>(h->tree, Pv); break;
  default: stock_vcom_leaves_kernel<3><<<(nleaf + 127)/128, 128, 0, s>>>(h->tree, Pv); break;
  }
  for (int l = h->tree.ltot - 1; l >= 0; l--) {
    const int nb = ((1 << l) + 127)/128;
    switch (h->dp.ndim) {
    case 1: stock_vcom_level_kernel<1><<<nb, 128, 0, s>>>(h->tree, l); break;
    case 2: stock_vcom_level_kernel<2><<<nb, 128, 0, s>>>(h->tree, l); break;
    default: stock_vcom_level_kernel<3><<<nb, 128, 0, s>>>(h->tree, l); break;
    }
  }
  h->stats.kernel_launches += 1 + h->tree.ltot;
  CUDA_TRY(h, cudaGetLastError());
  h->vcom_valid = true;
  return GSPH_OK;
}

int gsph_tree_build_sharded(gsph_handle* h, int phase, int rank, int world)
{
  int rc;
  switch (h->dp.ndim) {
  case 1: rc = build_tree_v2<1>(h, phase, rank, world); break;
  case 2: rc = build_tree_v2<2>(h, phase, rank, world); break;
  default: rc = build_tree_v2<3>(h, phase, rank, world); break;
  }
  return rc;                         /* (geo4 is packed by the permutation kernel, gsph_gather_fields) */
}

int gsph_tree_build(gsph_handle* h, int mode, double dt)
{
  int rc;
  switch (h->dp.ndim) {
  case 1: rc = build_tree_nd<1>(h, mode, dt); break;
  case 2: rc = build_tree_nd<2>(h, mode, dt); break;
  default: rc = build_tree_nd<3>(h, mode, dt); break;
  }
  return rc;                         /* (geo4 is packed by the permutation kernel, gsph_gather_fields) */
}

/* KDTree::UpdateAllHmaxValues after the density pass: only hmax and hbox change (cdistsqd keeps
 * the value of the last stock, exactly like the reference, KDTree.cpp:1163-1205).                 */
int gsph_tree_update_hmax(gsph_handle* h)
{
  switch (h->dp.ndim) {
  case 1: return stock_tree<1>(h, 0);
  case 2: return stock_tree<2>(h, 0);
  default: return stock_tree<3>(h, 0);
  }
}

int gsph_tree_update_active(gsph_handle* h) { return gsph_tree_update_hmax(h); }
