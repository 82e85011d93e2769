/*=================================================================================================
 *  device_common.cuh -- device helpers shared by the density / force / gravity kernels:
 *  smoothing-kernel evaluation, exact (non-contracted) distance arithmetic, periodic helpers and
 *  the warp-cooperative KD-tree walk.
 *=================================================================================================*/
#pragma once
#include "gsph_internal.cuh"

#define FULLMASK 0xffffffffu
#define NB_WALK_STACK 512             /* ints per warp: neighbour walks (frontiers of a few hundred nodes) */

struct PartView {
  double* f[NF_DOUBLE];
  int32_t* i[NF_INT];
};

/*------------------------------------------------------------------------------------------------
 * Exact arithmetic for anything that feeds a comparison or a truncation in the reference:
 * DotProduct is evaluated left to right without FMA (src/Headers/InlineFuncs.h:46-54); nvcc would
 * contract a*b+c, so these use the explicit round-to-nearest intrinsics.
 *-----------------------------------------------------------------------------------------------*/
template <int NDIM>
__device__ __forceinline__ double dot_exact(const double (&a)[3])
{
  double s = __dmul_rn(a[0], a[0]);
  if (NDIM > 1) s = __dadd_rn(s, __dmul_rn(a[1], a[1]));
  if (NDIM > 2) s = __dadd_rn(s, __dmul_rn(a[2], a[2]));
  return s;
}

/* NearestPeriodicVector / ApplyPeriodicDistanceCorrection, src/Headers/GhostNeighbours.hpp:119-166 */
template <int NDIM>
__device__ __forceinline__ void wrap_dr(const DevParams& dp, double (&dr)[3])
{
#pragma unroll
  for (int k = 0; k < NDIM; k++) {
    if (dp.periodic[k]) {
      if (dr[k] > dp.boxhalf[k]) dr[k] -= dp.boxsize[k];
      else if (dr[k] < -dp.boxhalf[k]) dr[k] += dp.boxsize[k];
    }
  }
}

/*------------------------------------------------------------------------------------------------
 * Smoothing kernels.  Analytic forms: M4 src/Headers/SmoothingKernel.h:131-240, quintic :281-408;
 * norms src/Hydrodynamics/{M4,Quintic}Kernel.cpp:39-53.  Tabulated forms: 1000-bin truncating
 * look-up without interpolation, SmoothingKernel.h:585-661.
 * `which`: 0 w0, 1 w1, 2 womega, 3 wzeta, 4 wgrav, 5 wpot.
 *-----------------------------------------------------------------------------------------------*/
template <int KERN, int NDIM>
__host__ __device__ __forceinline__ double kern_analytic(int which, double s, double kn)
{
  const double nd = (double) NDIM;
  const double s2 = s*s, s3 = s2*s, s4 = s2*s2, s5 = s4*s;
  if (KERN == GSPH_KERNEL_M4) {
    switch (which) {
    case 0:
      if (s < 1.0) return kn*(1.0 - 1.5*s2 + 0.75*s3);
      else if (s < 2.0) { const double t = 2.0 - s; return 0.25*kn*(t*t*t); }
      return 0.0;
    case 1:
      if (s < 1.0) return kn*(-3.0*s + 2.25*s2);
      else if (s < 2.0) return -0.75*kn*(2.0 - s)*(2.0 - s);
      return 0.0;
    case 2:
      if (s < 1.0) return kn*(-nd + 1.5*(nd + 2.0)*s2 - 0.75*(nd + 3.0)*s3);
      else if (s < 2.0)
        return kn*(-2.0*nd + 3.0*(nd + 1.0)*s - 1.50*(nd + 2.0)*s2 + 0.25*(nd + 3.0)*s3);
      return 0.0;
    case 3:
      if (s < 1.0) return 1.4 - 2.0*s2 + 1.5*s4 - 0.6*s5;
      else if (s < 2.0) return 1.6 - 4.0*s2 + 4.0*s3 - 1.5*s4 + 0.2*s5;
      return 0.0;
    case 4:
      if (s < 1.0) return 1.333333333333333333333*s - 1.2*s3 + 0.5*s4;
      else if (s < 2.0)
        return 2.6666666666666666667*s - 3.0*s2 + 1.2*s3 - 0.166666666666666666667*s4 -
          0.06666666666666666667/s2;
      return 1.0/s2;
    default:
      if (s < 1.0) return 1.4 - 0.666666666666666666666666*s2 + 0.3*s4 - 0.1*s5;
      else if (s < 2.0)
        return -1.0/(15.0*s) + 1.6 - 1.33333333333333333333333333*s2 + s3 - 0.3*s4 +
          (1.0/30.0)*s5;
      return 1.0/s;
    }
  }
  else if (KERN == GSPH_KERNEL_GAUSSIAN) {             /* SmoothingKernel.h:447-521 */
    if (!(s < 3.0)) return 0.0;
    switch (which) {
    case 0: return kn*exp(-s*s);
    case 1: return -2.0*kn*s*exp(-s*s);
    case 2: return kn*(2.0*s*exp(-s*s) - nd*exp(-s*s));
    default: return 0.0;
    }
  }
  else {
    const double s6 = s3*s3, s7 = s6*s;
    switch (which) {
    case 0:
      if (s < 1.0) return kn*(66.0 - 60.0*s2 + 30.0*s4 - 10.0*s5);
      else if (s < 2.0) return kn*(51.0 + 75.0*s - 210.0*s2 + 150.0*s3 - 45.0*s4 + 5.0*s5);
      else if (s < 3.0) return kn*(243.0 - 405*s + 270.0*s2 - 90.0*s3 + 15.0*s4 - s5);
      return 0.0;
    case 1:
      if (s < 1.0) return kn*(-120.0*s + 120.0*s3 - 50.0*s4);
      else if (s < 2.0) return kn*(75.0 - 420.0*s + 450.0*s2 - 180.0*s3 + 25.0*s4);
      else if (s < 3.0) return kn*(-405.0 + 540.0*s - 270.0*s2 + 60.0*s3 - 5.0*s4);
      return 0.0;
    case 2:
      if (s < 1.0)
        return kn*(-66.0*nd + 60.0*(nd + 2.0)*s2 - 30.0*(nd + 4.0)*s4 + 10.0*(nd + 5.0)*s5);
      else if (s < 2.0)
        return kn*(-51.0*nd - 75.0*(nd + 1.0)*s + 210.0*(nd + 2.0)*s2 - 150.0*(nd + 3.0)*s3 +
                   45.0*(nd + 4.0)*s4 - 5.0*(nd + 5.0)*s5);
      else if (s < 3.0)
        return kn*(-243.0*nd + 405.0*(nd + 1.0)*s - 270.0*(nd + 2.0)*s2 + 90.0*(nd + 3.0)*s3 -
                   15.0*(nd + 4.0)*s4 + (nd + 5.0)*s5);
      return 0.0;
    case 3:
      if (s < 1.0) return 33.0*s2 - 15.0*s4 + 5.0*s6 - 1.42857142857*s7 - 34.14285714;
      else if (s < 2.0)
        return 25.5*s2 + 25.0*s3 - 52.5*s4 + 30.0*s5 - 7.5*s6 + 0.7142857143*s7 - 33.785714286;
      else if (s < 3.0)
        return 121.5*s2 - 135.0*s3 + 67.5*s4 - 18.0*s5 + 2.5*s6 - 0.142857143*s7 - 52.07142857;
      return 0.0;
    case 4:
      if (s < 1.0)
        return (12.0/359.0)*(22.0*s - 12.0*s3 + (30.0/7.0)*s5 - (5.0/4.0)*s6);
      else if (s < 2.0)
        return (12.0/359.0)*(17.0*s + (75.0/4.0)*s2 - 42.0*s3 + 25.0*s4 - (45.0/7.0)*s5 +
                             (5.0/8.0)*s6 + (5.0/56.0)/s2);
      else if (s < 3.0)
        return (12.0/359.0)*(81.0*s - (405.0/4.0)*s2 + 54.0*s3 - 15.0*s4 + (15.0/7.0)*s5 -
                             (1.0/8.0)*s6 - (507.0/56.0)/s2);
      return 1.0/s2;
    default:
      if (s < 1.0)
        return (12.0/359.0)*(-11.0*s2 + 3.0*s4 - (5.0/7.0)*s6 + (5.0/28.0)*s7 + (478.0/14.0));
      else if (s < 2.0)
        return (12.0/359.0)*(-(17.0/2.0)*s2 - (25.0/4.0)*s3 + (21.0/2.0)*s4 - 5.0*s5 +
                             (15.0/14.0)*s6 - (5.0/56.0)*s7 + (473.0/14.0) + (5.0/56.0)/s);
      else if (s < 3.0)
        return (12.0/359.0)*(-(81.0/2.0)*s2 + (135.0/4.0)*s3 - (27.0/2.0)*s4 + 3.0*s5 -
                             (5.0/14.0)*s6 + (1.0/56.0)*s7 + (729.0/14.0) - (507.0/56.0)/s);
      return 1.0/s;
    }
  }
}

/* Kernel evaluator bound to one (kernel, ndim, tabulated) combination.  `tab` points at the six
 * tables (shared memory): 0 w0_s2, 1 womega_s2, 2 wzeta_s2, 3 w1, 4 wgrav, 5 wpot.               */
template <int KERN, int NDIM, bool TAB>
struct KernelEval {
  const double* tab;
  double kn, kr, krs, res_invkr, res_invkrs;
  __device__ __forceinline__ KernelEval(const DevParams& dp, const double* t) : tab(t)
  {
    kn = dp.kernnorm; kr = dp.kernrange; krs = dp.kernrangesqd;
    res_invkr = (double) GSPH_TABLE_RES/kr; res_invkrs = (double) GSPH_TABLE_RES/krs;
  }
  /* s^2-argument family used by ComputeH (which: 0 w0, 1 womega, 2 wzeta) */
  __device__ __forceinline__ void density_terms(double ssqd, double& w0, double& wom, double& wze) const
  {
    if (TAB) {
      const int idx = (int)(ssqd*res_invkrs);
      w0 = tab[idx]; wom = tab[GSPH_TABLE_RES + idx]; wze = tab[2*GSPH_TABLE_RES + idx];
    }
    else {
      const double s = sqrt(ssqd);
      w0 = kern_analytic<KERN, NDIM>(0, s, kn);
      wom = kern_analytic<KERN, NDIM>(2, s, kn);
      wze = kern_analytic<KERN, NDIM>(3, s, kn);
    }
  }
  __device__ __forceinline__ double w1(double s) const
  {
    if (TAB) { if (s >= kr) return 0.0; return tab[3*GSPH_TABLE_RES + (int)(s*res_invkr)]; }
    return kern_analytic<KERN, NDIM>(1, s, kn);
  }
  __device__ __forceinline__ double wgrav(double s) const
  {
    if (TAB) { if (s >= kr) return 1.0/(s*s); return tab[4*GSPH_TABLE_RES + (int)(s*res_invkr)]; }
    return kern_analytic<KERN, NDIM>(4, s, kn);
  }
  __device__ __forceinline__ double wpot(double s) const
  {
    if (TAB) { if (s >= kr) return 1.0/s; return tab[5*GSPH_TABLE_RES + (int)(s*res_invkr)]; }
    return kern_analytic<KERN, NDIM>(5, s, kn);
  }
  /* invh^2 * wgrav(s) and invh * wpot(s) for s = drmag*invh.  Outside the kernel both reduce to the
     Newtonian 1/drmag^2 and 1/drmag, taken from the caller's 1/drmag instead of two more divisions
     (the reference evaluates invh^2 * (1/s^2); the two agree to rounding).                        */
  __device__ __forceinline__ void grav_terms(double s, double invh, double invdr, double& g, double& p) const
  {
    if (s >= kr) { g = invdr*invdr; p = invdr; return; }
    if (TAB) {
      const int idx = (int)(s*res_invkr);
      g = invh*invh*tab[4*GSPH_TABLE_RES + idx]; p = invh*tab[5*GSPH_TABLE_RES + idx];
    }
    else { g = invh*invh*kern_analytic<KERN, NDIM>(4, s, kn); p = invh*kern_analytic<KERN, NDIM>(5, s, kn); }
  }
};

/*------------------------------------------------------------------------------------------------
 * warp helpers
 *-----------------------------------------------------------------------------------------------*/
__device__ __forceinline__ int warp_excl_scan(int v, int lane, int& total)
{
  int x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(FULLMASK, x, o);
    if (lane >= o) x += y;
  }
  total = __shfl_sync(FULLMASK, x, 31);
  return x - v;
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULLMASK, v, o);
  return v;
}

/*------------------------------------------------------------------------------------------------
 * Conservative periodic-aware interval-overlap test used by the neighbour walks.  The candidate
 * set only has to be a SUPERSET of the true neighbours (membership is decided afterwards by the
 * exact predicate), so a small relative slack is added.
 *-----------------------------------------------------------------------------------------------*/
template <int NDIM>
__device__ __forceinline__ bool boxes_overlap(const DevParams& dp, const double (&amin)[3],
                                              const double (&amax)[3], const double (&bmin)[3],
                                              const double (&bmax)[3])
{
#pragma unroll
  for (int k = 0; k < NDIM; k++) {
    const double ca = 0.5*(amin[k] + amax[k]), cb = 0.5*(bmin[k] + bmax[k]);
    const double hw = 0.5*(amax[k] - amin[k]) + 0.5*(bmax[k] - bmin[k]);
    double d = fabs(ca - cb);
    if (dp.periodic[k]) d = fmin(d, dp.boxsize[k] - d);
    if (d > hw*(1.0 + 1.0e-12) + 1.0e-300) return false;
  }
  return true;
}

/*------------------------------------------------------------------------------------------------
 * Warp-cooperative KD-tree walk.  The tree is a perfect binary tree in heap order (children of n
 * are 2n+1, 2n+2; leaves are the nodes >= leaf0).  Up to 32 nodes are popped from a per-warp
 * shared-memory stack and tested by one lane each; opened internal nodes push their children.
 * Visitor interface:
 *     int  test(int node, bool leaf)        per lane : 0 skip, 1 open, >=2 visitor-defined
 *     bool emit(int node, int code, bool leaf, int lane)   warp collective; false = overflow
 * This replaces the serial copen/cnext pointer chase of Tree::ComputeGatherNeighbourList,
 * ComputeNeighbourAndGhostList and ComputeGravityInteractionAndGhostList
 * (src/Tree/Tree.cpp:291-381, :562-617, :628-735).
 *-----------------------------------------------------------------------------------------------*/
template <int STACKCAP, class Visitor>
__device__ __forceinline__ bool warp_walk(const TreeDev& T, int* stack, Visitor& V, DevStatus* st)
{
  const int lane = threadIdx.x & 31;
  const int leaf0 = (1 << T.ltot) - 1;
  int sp = 1;
  if (lane == 0) stack[0] = 0;
  __syncwarp();
  while (sp > 0) {
    const int take = sp < 32 ? sp : 32;
    int node = -1;
    if (lane < take) node = stack[sp - 1 - lane];
    sp -= take;
    __syncwarp();
    const bool leaf = node >= leaf0;
    int code = 0;
    if (node >= 0 && T.count[node] > 0) code = V.test(node, leaf);
    const bool push = (code == 1) && !leaf;
    const unsigned pm = __ballot_sync(FULLMASK, push);
    const int npush = __popc(pm);
    if (sp + 2*npush > STACKCAP) {
      if (lane == 0) atomicExch(&st->err_stack_overflow, 1);
      return false;
    }
    if (push) {
      const int off = __popc(pm & ((1u << lane) - 1u));
      stack[sp + 2*off]     = 2*node + 2;
      stack[sp + 2*off + 1] = 2*node + 1;
    }
    sp += 2*npush;
    if (!V.emit(node, code, leaf, lane)) return false;
    __syncwarp();
  }
  return true;
}

/* Visitor pieces shared by the neighbour walks: append the particles of accepted leaves to a
 * per-warp candidate list in global scratch.                                                     */
struct CandList {
  int32_t* list; int cap; int n;
  DevStatus* st;
  bool quiet = false;        /* overflow is not an error (pre-walk kernels: the cohort is simply left to the main kernel) */
  __device__ __forceinline__ bool append_leaf(const TreeDev& T, int node, bool take, int lane)
  {
    int cnt = 0, first = 0;
    if (take) { cnt = T.count[node]; first = T.first[node]; }
    int total;
    const int off = warp_excl_scan(cnt, lane, total);
    if (total == 0) return true;
    if (n + total > cap) {
      if (lane == 0) { atomicExch(&st->err_cand_overflow, 1); atomicMax(&st->cand_need, (unsigned long long)(n + total)); }
      return false;
    }
    for (int i = 0; i < cnt; i++) list[n + off + i] = first + i;
    n += total;
    return true;
  }
};

/*------------------------------------------------------------------------------------------------
 * Single-precision walk machinery for the neighbour searches (density gather walk, force
 * scatter-gather walk).  The walks only have to produce a SUPERSET of the true neighbours -- the
 * exact predicate is applied afterwards in double precision -- so the node boxes are kept as
 * outward-rounded floats (TreeDev::nodeb / nodeh, 32 B each incl. first/count) and compared on the
 * FP32 pipe.  Periodic axes use the centre/half-width form with an explicit slack that covers the
 * float rounding of the box coordinates.
 *-----------------------------------------------------------------------------------------------*/
struct WalkBox {            /* per-kernel constants of the float box tests */
  float size[3], half[3], slack[3];
  int periodic[3];
  __device__ __forceinline__ WalkBox(const DevParams& dp)
  {
#pragma unroll
    for (int k = 0; k < 3; k++) {
      periodic[k] = dp.periodic[k];
      size[k] = (float) dp.boxsize[k]; half[k] = (float) dp.boxhalf[k];
      slack[k] = 2.0e-6f*(float)(fabs(dp.boxmin[k]) + fabs(dp.boxmax[k]) + dp.boxsize[k]);
    }
  }
};

template <int NDIM>
__device__ __forceinline__ bool fboxes_overlap(const WalkBox& wb, const float (&amin)[3], const float (&amax)[3],
                                               const float4& bmin, const float4& bmax, bool& wrapped)
{
  const float bmn[3] = {bmin.x, bmin.y, bmin.z}, bmx[3] = {bmax.x, bmax.y, bmax.z};
#pragma unroll
  for (int k = 0; k < NDIM; k++) {
    if (!wb.periodic[k]) { if (bmx[k] < amin[k] || bmn[k] > amax[k]) return false; }
    else {
      const float hw = 0.5f*(amax[k] - amin[k]) + 0.5f*(bmx[k] - bmn[k]) + wb.slack[k];
      float d = fabsf(0.5f*(amin[k] + amax[k]) - 0.5f*(bmn[k] + bmx[k]));
      if (d + hw > wb.half[k]) wrapped = true;
      d = fminf(d, wb.size[k] - d);
      if (d > hw) return false;
    }
  }
  return true;
}

/* Same traversal as warp_walk, for visitors that read the packed float node records themselves
 * (test() returns 0 for empty nodes and remembers node/first/count for emit()).
 *   leaf_level : nodes of this level are treated as leaves (normally ltot);
 *   init/ninit : start from this node list instead of the top of the tree.                         */
template <int STACKCAP, class Visitor>
__device__ __forceinline__ bool warp_walk_f(int leaf_level, int* stack, Visitor& V, DevStatus* st,
                                            const int* init = nullptr, int ninit = 0)
{
  const int lane = threadIdx.x & 31;
  const int leaf0 = (1 << leaf_level) - 1;
  int sp;
  if (init) {
    for (int i = lane; i < ninit; i += 32) stack[i] = init[i];
    sp = ninit;
  }
  /* Without a frontier the walk starts with the 32 nodes of level 5 (one round of independent loads
     instead of six dependent ones): a node that meets the search box has ancestors that meet it
     too, so skipping the ancestors never loses a leaf.                                            */
  else if (leaf_level > 5) { stack[lane] = 31 + lane; sp = 32; }
  else { if (lane == 0) stack[0] = 0; sp = 1; }
  __syncwarp();
  while (sp > 0) {
    const int take = sp < 32 ? sp : 32;
    int node = -1;
    if (lane < take) node = stack[sp - 1 - lane];
    sp -= take;
    __syncwarp();
    const bool leaf = node >= leaf0;
    int code = 0;
    if (node >= 0) code = V.test(node, leaf);
    const bool push = (code == 1) && !leaf;
    const unsigned pm = __ballot_sync(FULLMASK, push);
    const int npush = __popc(pm);
    if (sp + 2*npush > STACKCAP) {
      if (lane == 0 && st) atomicExch(&st->err_stack_overflow, 1);
      return false;
    }
    if (push) {
      const int off = __popc(pm & ((1u << lane) - 1u));
      stack[sp + 2*off]     = 2*node + 2;
      stack[sp + 2*off + 1] = 2*node + 1;
    }
    sp += 2*npush;
    if (!V.emit(code == 1 && leaf, lane)) return false;
    __syncwarp();
  }
  return true;
}

/*------------------------------------------------------------------------------------------------
 * Mirror boundaries.  The reference materialises mirror images as ghost particles for the density
 * pass (Hydrodynamics.cpp:217-290) and builds them on the fly in the force pass
 * (GhostNeighbourFinder::ConstructGhostsScatterGather, GhostNeighbours.hpp:252-268,411-450):
 * r_k -> 2*wall - r_k, v_k -> -v_k, images of images across dimensions.  Here an image is a
 * candidate id tagged with an IMAGE CODE (mixed-radix digit per dimension: 0 none, then lhs, rhs as
 * far as that wall is a mirror) in bits 26..30; code 0 is the real particle.  The walks are repeated
 * with the reflected search boxes (a reflection is an isometry, so "image of the node meets the box"
 * = "node meets the image of the box").
 *-----------------------------------------------------------------------------------------------*/
#define IMG_SHIFT 26
#define IMG_MASK ((1 << IMG_SHIFT) - 1)

struct ImageOp { int refl[3]; double wall2[3]; };       /* reflected in dimension k; 2*wall */

template <int NDIM>
__device__ __forceinline__ ImageOp image_op(const DevParams& dp, int code)
{
  ImageOp op;
#pragma unroll
  for (int k = 0; k < 3; k++) { op.refl[k] = 0; op.wall2[k] = 0.0; }
#pragma unroll
  for (int k = 0; k < NDIM; k++) {
    const int radix = 1 + dp.mirror[k][0] + dp.mirror[k][1];
    const int dgt = code % radix;
    code /= radix;
    if (dgt) {
      op.refl[k] = 1;
      op.wall2[k] = 2.0*((dgt == 1 && dp.mirror[k][0]) ? dp.boxmin[k] : dp.boxmax[k]);
    }
  }
  return op;
}

/* 2*x_mirror - r, Particle.h:603-607 (one rounding, like the reference) */
__device__ __forceinline__ double reflect_coord(const ImageOp& op, int k, double x)
{
  return op.refl[k] ? __dadd_rn(op.wall2[k], -x) : x;
}

/* append the slab [first, first+cnt) of every lane with take = true to the candidate list */
__device__ __forceinline__ bool cand_append_range(CandList& cl, int first, int cnt, bool take, int lane, int tag = 0)
{
  if (!take) cnt = 0;
  int total;
  const int off = warp_excl_scan(cnt, lane, total);
  if (total == 0) return true;
  if (cl.n + total > cl.cap) {
    if (lane == 0 && !cl.quiet) { atomicExch(&cl.st->err_cand_overflow, 1); atomicMax(&cl.st->cand_need, (unsigned long long)(cl.n + total)); }
    return false;
  }
  for (int i = 0; i < cnt; i++) cl.list[cl.n + off + i] = (first + i) | tag;
  cl.n += total;
  return true;
}

/*------------------------------------------------------------------------------------------------
 * Packed-FP32 candidate filter (Blackwell FADD2 / FMUL2 / FFMA2: two candidates per instruction).
 * The tile holds 32 candidates as SoA floats (x[32] y[32] z[32] w[32]); a lane tests all of them
 * against its own position and returns the bitmap of candidates with d^2 < max(thr, w_j) (force
 * pass) or d^2 < thr (density pass, USEW = false).  Unused tile slots must hold x = +inf.
 *-----------------------------------------------------------------------------------------------*/
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t f2_pack(float lo, float hi)
{ return ((f32x2_t) __float_as_uint(hi) << 32) | (f32x2_t) __float_as_uint(lo); }
__device__ __forceinline__ f32x2_t f2_sub(f32x2_t a, f32x2_t b) { f32x2_t d; asm("sub.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2_t f2_mul(f32x2_t a, f32x2_t b) { f32x2_t d; asm("mul.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2_t f2_fma(f32x2_t a, f32x2_t b, f32x2_t c) { f32x2_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

template <bool USEW>
__device__ __forceinline__ unsigned filter_tile(const float* tile /* x|y|z|w, 32 each */, const float (&xf)[3], float thr)
{
  const f32x2_t X = f2_pack(xf[0], xf[0]), Y = f2_pack(xf[1], xf[1]), Z = f2_pack(xf[2], xf[2]);
  const ulonglong2* tx = reinterpret_cast<const ulonglong2*>(tile);
  const ulonglong2* ty = reinterpret_cast<const ulonglong2*>(tile + 32);
  const ulonglong2* tz = reinterpret_cast<const ulonglong2*>(tile + 64);
  const float4* tw = reinterpret_cast<const float4*>(tile + 96);
  unsigned pass = 0;
#pragma unroll
  for (int q = 0; q < 8; q++) {                      /* 4 candidates per step */
    const ulonglong2 ax = tx[q], ay = ty[q], az = tz[q];
    const f32x2_t dx0 = f2_sub(ax.x, X), dx1 = f2_sub(ax.y, X);
    const f32x2_t dy0 = f2_sub(ay.x, Y), dy1 = f2_sub(ay.y, Y);
    const f32x2_t dz0 = f2_sub(az.x, Z), dz1 = f2_sub(az.y, Z);
    const f32x2_t s0 = f2_fma(dz0, dz0, f2_fma(dy0, dy0, f2_mul(dx0, dx0)));
    const f32x2_t s1 = f2_fma(dz1, dz1, f2_fma(dy1, dy1, f2_mul(dx1, dx1)));
    const float d0 = __uint_as_float((unsigned) s0), d1 = __uint_as_float((unsigned)(s0 >> 32));
    const float d2 = __uint_as_float((unsigned) s1), d3 = __uint_as_float((unsigned)(s1 >> 32));
    float t0 = thr, t1 = thr, t2 = thr, t3 = thr;
    if (USEW) { const float4 w = tw[q]; t0 = fmaxf(thr, w.x); t1 = fmaxf(thr, w.y); t2 = fmaxf(thr, w.z); t3 = fmaxf(thr, w.w); }
    if (d0 < t0) pass |= 1u << (4*q);
    if (d1 < t1) pass |= 2u << (4*q);
    if (d2 < t2) pass |= 4u << (4*q);
    if (d3 < t3) pass |= 8u << (4*q);
  }
  return pass;
}

/* 16-byte asynchronous global->shared copies (LDGSTS): the gather of the next neighbour slot is in
 * flight while the terms of the current one are evaluated.                                         */
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src)
{
  const unsigned s = (unsigned) __cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" :: "r"(s), "l"(gmem_src) : "memory");
}
/* the same through L2 only, without allocating in L1: the record gathers of the pair kernels (every record is
   used once per list slot; measured 2-3 ms per pass better than .ca at 64 Mi particles) */
__device__ __forceinline__ void cp_async16_cg(void* smem_dst, const void* gmem_src)
{
  const unsigned s = (unsigned) __cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int NPENDING>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" :: "n"(NPENDING) : "memory"); }
