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
 *  rcell/rmax/hmax) and, for fast multipoles, expands about the leaf's centre of mass; forces only
 *  match if that granularity is kept.  A walk per leaf is latency bound on a GPU, so one warp takes
 *  a COHORT of 8 neighbouring leaves (one KD-tree node 3 levels above the leaves) and walks the
 *  tree ONCE for all of them: every stack entry carries an 8-bit mask of the leaves for which the
 *  node is still undecided, the lane that tests a node evaluates the reference's opening criterion
 *  separately (and exactly) for every leaf in the mask, and children inherit the mask of the
 *  leaves that opened their parent.  Accepted cells and direct-sum leaves are queued in small
 *  shared-memory buffers as (node, leaf-mask) and consumed as soon as a tile is full -- no
 *  interaction list is ever written to HBM.  In the consumption phase the lanes are the TARGET
 *  PARTICLES (4 lanes per leaf, T particles per lane): a lane streams the staged tile (broadcast
 *  reads) and accumulates the entries whose mask contains its leaf.  For fast multipoles the four
 *  lanes of a leaf share the cell entries and combine their partial Taylor coefficients with two
 *  warp-shuffle steps.
 *
 *  Split with hydro.cu: a particle j of an opened leaf is a kernel-overlap neighbour of target i
 *  iff the reference's predicate holds (NeighbourManager.h:521); those pairs get the softened
 *  terms in hydro.cu.  Here j contributes the unsoftened term iff the predicate fails -- the same
 *  partition NeighbourManager::TrimNeighbourLists makes (NeighbourManager.h:503-542).
 *=================================================================================================*/
#include "device_common.cuh"

#define GRAV_THREADS 256
#define GRAV_WARPS (GRAV_THREADS/32)
#define GRAV_NT 8                      /* max particles per leaf on this path */
#define GRAV_STACK 512                 /* (node, mask) pairs per warp */
#define CELL_TILE 32
#define CELL_BUF 64
#define DIR_PART 64                    /* staged source particles per direct tile */
#define DIR_TILE_MAX 16                /* direct leaves per tile: DIR_PART/(4*T) */
#define DIR_BUF (DIR_TILE_MAX + 32)
#define GRAV_LEAVES 8                  /* leaves per cohort: 4 lanes each */
#ifndef GRAV_MINBLOCKS
#define GRAV_MINBLOCKS 2               /* resident blocks per SM the register allocation is tuned for */
#endif

namespace {

/* 1/sqrt(x) for finite normal x > 0: hardware seed (MUFU.RSQ64H, ~2^-22) + one cubic Newton step,
 * branch-free so that several independent evaluations can be in flight per thread (the library
 * rsqrt() carries a special-case branch that blocks that interleaving).  Error <= 2 ulp.         */
__device__ __forceinline__ double rsqrt_nr(double x)
{
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-x*y, y, 1.0);
  return fma(y*e, fma(0.375, e, 0.5), y);
}

template <bool QUAD>
struct WarpSharedT {
  int2 stack[GRAV_STACK];
  double leafgeom[GRAV_LEAVES][6];     /* rcell xyz, rmax, kernrange*hmax, MAC auxiliary (amin | macfactor) of the cohort's leaves */
  int cell_node[CELL_BUF]; unsigned cell_mask[CELL_BUF];
  int dir_node[DIR_BUF]; unsigned dir_mask[DIR_BUF];
  double cell_data[CELL_TILE][QUAD ? 9 : 4];   /* com xyz, m [, q0..q4] */
  double dir_data[DIR_PART][5];        /* x, y, z, m, hrangesqd of the staged sources: 4*T slots per direct
                                          leaf, unused slots hold m = 0 (they contribute exactly nothing) */
  int dir_count[DIR_TILE_MAX];
};

struct FastMp { double ac[3], q[6], pot; };   /* dphi == ac for HoT = false */

/* (d2 < ha || d2 < hb) ? 0 : m  with two predicate instructions: written as C the compiler turns the condition
   into d2 < fmax(ha, hb), and FP64 max is an eight-instruction emulation on this part */
__device__ __forceinline__ double mass_unless_overlap(double d2, double ha, double hb, double m)
{
  double r;
  asm("{ .reg .pred p, q; setp.lt.f64 p, %1, %2; setp.lt.or.f64 q, %1, %3, p; selp.f64 %0, 0d0000000000000000, %4, q; }"
      : "=d"(r) : "d"(d2), "d"(ha), "d"(hb), "d"(m));
  return r;
}

template <int MULTIPOLE, int T, bool GEOMAC>
__global__ void __launch_bounds__(GRAV_THREADS, (T == 1 && MULTIPOLE != GSPH_MULTIPOLE_FAST_QUADRUPOLE) ? GRAV_MINBLOCKS : 2)
gravity_kernel(PartView P, TreeDev tree, DevParams dp, int cohort_level, int* work_counter,
               DevStatus* st, uint8_t* done, int2* gstack, int gstack_cap, int smem_stack_cap)
{
  extern __shared__ char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  typedef WarpSharedT<(MULTIPOLE == GSPH_MULTIPOLE_QUADRUPOLE || MULTIPOLE == GSPH_MULTIPOLE_FAST_QUADRUPOLE)> WarpShared;
  WarpShared& S = reinterpret_cast<WarpShared*>(smem_raw)[warp];
  const int leaf0 = (1 << tree.ltot) - 1;
  const int ncohort = 1 << cohort_level;
  const int nl = 1 << (tree.ltot - cohort_level);          /* leaves per cohort (<= 8) */
  /* walk stack: shared memory; a cohort whose walk outgrows it is redone by a second launch whose
     stacks live in HBM (gstack) -- deep, strongly clustered trees never fail, they only get slower */
  int2* const stk = gstack ? gstack + (size_t)(blockIdx.x*GRAV_WARPS + warp)*gstack_cap : S.stack;
  const int stkcap = gstack ? gstack_cap : smem_stack_cap;
  const int myl = lane >> 2, sub = lane & 3;               /* my leaf inside the cohort, my slot */
  constexpr int DIR_TILE = DIR_PART/(4*T);               /* direct leaves per tile (<= 4*T particles each) */
  static_assert(DIR_TILE <= DIR_TILE_MAX, "direct tile");
  constexpr bool FAST = (MULTIPOLE == GSPH_MULTIPOLE_FAST_MONOPOLE || MULTIPOLE == GSPH_MULTIPOLE_FAST_QUADRUPOLE);
  constexpr bool QUAD = (MULTIPOLE == GSPH_MULTIPOLE_QUADRUPOLE || MULTIPOLE == GSPH_MULTIPOLE_FAST_QUADRUPOLE);
  unsigned long long loc_direct = 0, loc_cells = 0, loc_part = 0;

  while (true) {
    int c = 0;
    if (lane == 0) c = atomicAdd(work_counter, 1);
    c = __shfl_sync(FULLMASK, c, 0);
    if (c >= ncohort) break;
    const int cnode = ncohort - 1 + c;
    if (tree.count[cnode] == 0 || tree.nactive[cnode] == 0 || done[c] ||
        GSPH_RANGE_OUTSIDE_SLICE(dp, tree.first[cnode], tree.count[cnode])) continue;

    /* ---- this lane's leaf and target particles ------------------------------------------------------*/
    const int myleaf = ((cnode + 1) << (tree.ltot - cohort_level)) - 1 + myl;    /* heap id */
    int first = 0, cnt = 0;
    bool leaf_live = false;
    if (myl < nl) {
      first = tree.first[myleaf]; cnt = tree.count[myleaf];
      leaf_live = cnt > 0 && tree.nactive[myleaf] > 0 && !GSPH_RANGE_OUTSIDE_SLICE(dp, first, cnt);
    }
    double x[T][3], hr2[T], ax[T], ay[T], az[T], pot[T];
    bool tlive[T];
#pragma unroll
    for (int t = 0; t < T; t++) {
      const int o = sub + 4*t;
      tlive[t] = leaf_live && o < cnt && GSPH_IN_SLICE(dp, first + o);
      x[t][0] = x[t][1] = x[t][2] = 0.0; hr2[t] = 0.0; ax[t] = ay[t] = az[t] = pot[t] = 0.0;
      if (tlive[t]) {
        x[t][0] = P.f[F_RX][first + o]; x[t][1] = P.f[F_RY][first + o]; x[t][2] = P.f[F_RZ][first + o];
        hr2[t] = P.f[F_HR2][first + o];
      }
    }
    double rcom[3] = {0, 0, 0};
    FastMp F;
    F.pot = 0.0;
#pragma unroll
    for (int k = 0; k < 3; k++) F.ac[k] = 0.0;
#pragma unroll
    for (int k = 0; k < 6; k++) F.q[k] = 0.0;
    __syncwarp();
    if (sub == 0 && myl < nl) {
#pragma unroll
      for (int k = 0; k < 3; k++) S.leafgeom[myl][k] = leaf_live ? tree.rcell[k][myleaf] : 0.0;
      S.leafgeom[myl][3] = leaf_live ? tree.rmax[myleaf] : 0.0;
      S.leafgeom[myl][4] = leaf_live ? __dmul_rn(dp.kernrange, tree.hmax[myleaf]) : 0.0;
      if (!GEOMAC) S.leafgeom[myl][5] = (leaf_live && tree.macaux) ? tree.macaux[myleaf] : 0.0;
    }
    if (FAST && leaf_live) { rcom[0] = tree.com[0][myleaf]; rcom[1] = tree.com[1][myleaf]; rcom[2] = tree.com[2][myleaf]; }
    const unsigned lb = __ballot_sync(FULLMASK, leaf_live && sub == 0);
    unsigned rootmask = 0;
#pragma unroll
    for (int l = 0; l < GRAV_LEAVES; l++) rootmask |= ((lb >> (4*l)) & 1u) << l;
    int ncell_mine = 0, ndir_mine = 0;

    /* ---- consumption of queued entries ----------------------------------------------------------------*/
    auto consume_cells = [&](int n) {
      __syncwarp();
      if (lane < n) {
        const int cc = S.cell_node[lane];
        const double2 m0 = tree.mnode[2*(size_t) cc], m1 = tree.mnode[2*(size_t) cc + 1];
        S.cell_data[lane][0] = m0.x; S.cell_data[lane][1] = m0.y;
        S.cell_data[lane][2] = m1.x; S.cell_data[lane][3] = m1.y;
        if (QUAD) {
#pragma unroll
          for (int k = 0; k < 5; k++) S.cell_data[lane][4 + k] = tree.q[k][cc];
        }
      }
      __syncwarp();
      /* every leaf (4 lanes) visits only the entries whose mask names it: one ballot per leaf gives
         the leaf's entry bitmap; with fast multipoles the 4 lanes split the set bits round-robin,
         otherwise each lane (a different target particle) takes them all */
      unsigned ebits = 0;
      {
        const unsigned mm = lane < n ? S.cell_mask[lane] : 0u;
#pragma unroll
        for (int l = 0; l < GRAV_LEAVES; l++) {
          const unsigned bl = __ballot_sync(FULLMASK, (mm >> l) & 1u);
          if (l == myl) ebits = bl;
        }
      }
      if (FAST) { for (int q = 0; q < sub; q++) ebits &= ebits - 1u; }
      if (FAST && !QUAD) {
        /* fast monopole: two independent entries per iteration (instruction-level parallelism hides
           the latency of the dependent rsqrt chains); a missing second entry gets mass 0 */
        while (ebits) {
          const int e0 = __ffs(ebits) - 1;
          ebits &= ebits - 1u; ebits &= ebits - 1u; ebits &= ebits - 1u; ebits &= ebits - 1u;
          const bool two = ebits != 0u;
          const int e1 = two ? __ffs(ebits) - 1 : e0;
          if (two) { ebits &= ebits - 1u; ebits &= ebits - 1u; ebits &= ebits - 1u; ebits &= ebits - 1u; }
          ncell_mine += two ? 2 : 1;
          const double ax0 = S.cell_data[e0][0] - rcom[0], ay0 = S.cell_data[e0][1] - rcom[1], az0 = S.cell_data[e0][2] - rcom[2];
          const double bx0 = S.cell_data[e1][0] - rcom[0], by0 = S.cell_data[e1][1] - rcom[1], bz0 = S.cell_data[e1][2] - rcom[2];
          const double ma = S.cell_data[e0][3], mb = two ? S.cell_data[e1][3] : 0.0;
          const double ia = rsqrt_nr(ax0*ax0 + ay0*ay0 + az0*az0), ib = rsqrt_nr(bx0*bx0 + by0*by0 + bz0*bz0);
          const double ia2 = ia*ia, ib2 = ib*ib;
          const double ca = ma*ia2*ia, cb = mb*ib2*ib;
          F.pot += ma*ia; F.pot += mb*ib;
          F.ac[0] += ca*ax0; F.ac[1] += ca*ay0; F.ac[2] += ca*az0;
          F.ac[0] += cb*bx0; F.ac[1] += cb*by0; F.ac[2] += cb*bz0;
          F.q[0] += ca*(3.0*ax0*ax0*ia2 - 1.0); F.q[1] += ca*(3.0*ax0*ay0*ia2); F.q[2] += ca*(3.0*ay0*ay0*ia2 - 1.0);
          F.q[3] += ca*(3.0*az0*ax0*ia2);       F.q[4] += ca*(3.0*az0*ay0*ia2); F.q[5] += ca*(3.0*az0*az0*ia2 - 1.0);
          F.q[0] += cb*(3.0*bx0*bx0*ib2 - 1.0); F.q[1] += cb*(3.0*bx0*by0*ib2); F.q[2] += cb*(3.0*by0*by0*ib2 - 1.0);
          F.q[3] += cb*(3.0*bz0*bx0*ib2);       F.q[4] += cb*(3.0*bz0*by0*ib2); F.q[5] += cb*(3.0*bz0*bz0*ib2 - 1.0);
        }
      }
      while (ebits) {
        const int e = __ffs(ebits) - 1;
        ebits &= ebits - 1u;
        if (FAST) { ebits &= ebits - 1u; ebits &= ebits - 1u; ebits &= ebits - 1u; }
        ncell_mine++;
        const double cx = S.cell_data[e][0], cy = S.cell_data[e][1], cz = S.cell_data[e][2], cm = S.cell_data[e][3];
        double q0 = 0, q1 = 0, q2 = 0, q3 = 0, q4 = 0;
        if (QUAD) { q0 = S.cell_data[e][4]; q1 = S.cell_data[e][5]; q2 = S.cell_data[e][6]; q3 = S.cell_data[e][7]; q4 = S.cell_data[e][8]; }
        if (FAST) {
          /* Taylor coefficients about the leaf's centre of mass (NeighbourSearch.h:481-760) */
          {
            const double dx = cx - rcom[0], dy = cy - rcom[1], dz = cz - rcom[2];
            const double inv = rsqrt_nr(dx*dx + dy*dy + dz*dz);
            const double inv2 = inv*inv;
            const double mc = cm*inv2*inv;
            F.pot += cm*inv;
            F.ac[0] += mc*dx; F.ac[1] += mc*dy; F.ac[2] += mc*dz;
            F.q[0] += mc*(3.0*dx*dx*inv2 - 1.0);
            F.q[1] += mc*(3.0*dx*dy*inv2);
            F.q[2] += mc*(3.0*dy*dy*inv2 - 1.0);
            F.q[3] += mc*(3.0*dz*dx*inv2);
            F.q[4] += mc*(3.0*dz*dy*inv2);
            F.q[5] += mc*(3.0*dz*dz*inv2 - 1.0);
          }
          if (QUAD) {
            const double dx = rcom[0] - cx, dy = rcom[1] - cy, dz = rcom[2] - cz;
            const double inv = rsqrt_nr(dx*dx + dy*dy + dz*dz + GSPH_SMALL);
            const double inv2 = inv*inv;
            const double inv5 = inv2*inv2*inv;
            const double qs = q0*dx*dx + q2*dy*dy - (q0 + q2)*dz*dz + 2.0*(q1*dx*dy + q3*dx*dz + q4*dy*dz);
            const double qf = 2.5*qs*inv5*inv2;
            double qx0 = (q0*dx + q1*dy + q3*dz)*inv5;
            double qx1 = (q1*dx + q2*dy + q4*dz)*inv5;
            double qx2 = (q3*dx + q4*dy - (q0 + q2)*dz)*inv5;
            F.pot += 0.5*qs*inv5;
            F.ac[0] += qx0 - qf*dx; F.ac[1] += qx1 - qf*dy; F.ac[2] += qx2 - qf*dz;
            qx0 *= 5.0*inv2; qx1 *= 5.0*inv2; qx2 *= 5.0*inv2;
            F.q[0] += qf*(7.0*dx*dx*inv2 - 1.0) - (qx0*dx + qx0*dx - q0*inv5);
            F.q[1] += qf*(7.0*dx*dy*inv2) - (qx0*dy + qx1*dx - q1*inv5);
            F.q[2] += qf*(7.0*dy*dy*inv2 - 1.0) - (qx1*dy + qx1*dy - q2*inv5);
            F.q[3] += qf*(7.0*dx*dz*inv2) - (qx0*dz + qx2*dx - q3*inv5);
            F.q[4] += qf*(7.0*dy*dz*inv2) - (qx1*dz + qx2*dy - q4*inv5);
            F.q[5] += qf*(7.0*dz*dz*inv2 - 1.0) - (qx2*dz + qx2*dz + (q0 + q2)*inv5);
          }
        }
        else {
#pragma unroll
          for (int t = 0; t < T; t++) {
            if (!tlive[t]) continue;
            const double dx = x[t][0] - cx, dy = x[t][1] - cy, dz = x[t][2] - cz;
            const double inv = rsqrt_nr(dx*dx + dy*dy + dz*dz + GSPH_SMALL);
            const double inv2 = inv*inv;
            const double f = cm*inv2*inv;
            if (!QUAD) {                                       /* NeighbourSearch.h:349-377 */
              pot[t] += cm*inv;
              ax[t] -= f*dx; ay[t] -= f*dy; az[t] -= f*dz;
            }
            else {                                             /* NeighbourSearch.h:384-475 */
              const double inv5 = inv2*inv2*inv;
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
      __syncwarp();
    };

    auto consume_direct = [&](int n0, int n) {
      /* stage the particles of direct leaves [n0, n0+n), n <= DIR_TILE: 4*T slots per leaf */
      constexpr int SLOTS = 4*T;
      __syncwarp();
      unsigned ebits = 0;
      {
        const unsigned mm = lane < n ? S.dir_mask[n0 + lane] : 0u;
#pragma unroll
        for (int l = 0; l < GRAV_LEAVES; l++) {
          const unsigned bl = __ballot_sync(FULLMASK, (mm >> l) & 1u);
          if (l == myl) ebits = bl;
        }
      }
      for (int s = lane; s < n*SLOTS; s += 32) {
        const int e = s/SLOTS, k = s - e*SLOTS;
        const int ln = S.dir_node[n0 + e];
        const int lf = tree.first[ln], lc = tree.count[ln];
        if (k == 0) S.dir_count[e] = lc;
        double sx = 0.0, sy = 0.0, sz = 0.0, sm = 0.0, sh = -1.0;
        if (k < lc) {
          const int j = lf + k;
          const bool dead = ((unsigned) P.i[I_FLAGS][j]) & GSPH_FLAG_DEAD;
          sx = P.f[F_RX][j]; sy = P.f[F_RY][j]; sz = P.f[F_RZ][j];
          sm = dead ? 0.0 : P.f[F_M][j]; sh = dead ? -1.0 : P.f[F_HR2][j];
        }
        S.dir_data[s][0] = sx; S.dir_data[s][1] = sy; S.dir_data[s][2] = sz; S.dir_data[s][3] = sm; S.dir_data[s][4] = sh;
      }
      __syncwarp();
      /* the 4 lanes of a leaf (different target particles) visit the leaf's own entries only; the
         SLOTS sources of an entry are evaluated branch-free (a kernel-overlap pair, which belongs to
         hydro.cu, or an empty slot simply gets mass 0) so that their rsqrt chains overlap */
      while (ebits) {
        const int e = __ffs(ebits) - 1;
        ebits &= ebits - 1u;
        if (sub == 0) ndir_mine += S.dir_count[e];
#pragma unroll
        for (int k = 0; k < SLOTS; k++) {
          const double* pj = S.dir_data[e*SLOTS + k];
          const double xj = pj[0], yj = pj[1], zj = pj[2], mj = pj[3], hr2j = pj[4];
#pragma unroll
          for (int t = 0; t < T; t++) {
            double dr[3] = {xj - x[t][0], yj - x[t][1], zj - x[t][2]};
            const double d2 = dot_exact<3>(dr);
            const double mm = (!tlive[t] || d2 < hr2[t] || d2 < hr2j) ? 0.0 : mj;
            const double inv = rsqrt_nr(d2 + GSPH_SMALL);           /* ComputeDirectGravForces */
            const double f = mm*inv*inv*inv;
            ax[t] += f*dr[0]; ay[t] += f*dr[1]; az[t] += f*dr[2];
            pot[t] += mm*inv;
          }
        }
      }
      __syncwarp();
    };

    /* ---- masked cohort walk ------------------------------------------------------------------------*/
    int sp = 0, ncb = 0, ndb = 0;
    bool failed = false;
    if (rootmask) {
      if (lane == 0) stk[0] = make_int2(0, (int) rootmask);
      sp = 1;
    }
    __syncwarp();
    while (sp > 0) {
      const int take = sp < 32 ? sp : 32;
      int node = -1; unsigned mask = 0;
      if (lane < take) { const int2 e = stk[sp - 1 - lane]; node = e.x; mask = (unsigned) e.y; }
      sp -= take;
      __syncwarp();
      unsigned open_mask = 0, acc_mask = 0;
      bool isleaf = false; int ncount = 0;
      if (node >= 0) {
        const double2 g0 = tree.gnode[4*(size_t) node], g1 = tree.gnode[4*(size_t) node + 1];
        const double2 g2 = tree.gnode[4*(size_t) node + 2], g3 = tree.gnode[4*(size_t) node + 3];
        ncount = __double2hiint(g3.x);
        isleaf = node >= leaf0;
        if (ncount > 0) {
          const double rx = g0.x, ry = g0.y, rz = g1.x;
          const double crmax = g1.y, cdist = g2.x;
          const double chr = g2.y;                         /* kernrange*hmax of the node */
          const double cmac = g3.y;                        /* gadget2: rmax^2*m; eigenmac: TreeCell::mac */
          const int mac = dp.gravity_mac; const double macerror = dp.macerror;
          /* only the leaves still undecided for this node (set bits of the mask) are evaluated, two
             per iteration so that their dependent arithmetic chains overlap */
          auto opens = [&](int l) -> bool {
            const double* lg = S.leafgeom[l];
            double dr[3];
            dr[0] = __dadd_rn(rx, -lg[0]); dr[1] = __dadd_rn(ry, -lg[1]); dr[2] = __dadd_rn(rz, -lg[2]);
            const double drsqd = dot_exact<3>(dr);
            const double rmax = lg[3];
            const double a = __dadd_rn(__dadd_rn(crmax, rmax), lg[4]);
            const double b = __dadd_rn(__dadd_rn(rmax, crmax), chr);
            /* Tree.cpp:660-731: overlap -> open; else open_cell_for_gravity (Tree.h:413-432):
               geometric part drsqd < cdistsqd, then the eigenvalue or the Gadget-2 criterion with the
               target leaf's macfactor / amin, evaluated left to right without contraction */
            if (drsqd <= __dmul_rn(a, a) || drsqd <= __dmul_rn(b, b) || drsqd < cdist) return true;
            if (GEOMAC) return false;
            if (mac == GSPH_MAC_EIGEN) return drsqd < __dmul_rn(cmac, lg[5]);
            if (mac == GSPH_MAC_GADGET2) return __dmul_rn(__dmul_rn(__dmul_rn(drsqd, drsqd), lg[5]), macerror) < cmac;
            return false;
          };
          unsigned todo = mask;
          while (todo) {
            const int l0 = __ffs(todo) - 1;
            todo &= todo - 1u;
            const bool two = todo != 0u;
            const int l1 = two ? __ffs(todo) - 1 : l0;
            todo &= todo - 1u;
            const bool o0 = opens(l0), o1 = opens(l1);
            if (o0) open_mask |= 1u << l0; else acc_mask |= 1u << l0;
            if (two) { if (o1) open_mask |= 1u << l1; else acc_mask |= 1u << l1; }
          }
        }
      }
      /* children of opened internal nodes */
      const bool push = open_mask && !isleaf;
      const unsigned pm = __ballot_sync(FULLMASK, push);
      const int npush = __popc(pm);
      if (sp + 2*npush > stkcap) { failed = true; break; }
      if (push) {
        const int off = __popc(pm & ((1u << lane) - 1u));
        stk[sp + 2*off] = make_int2(2*node + 2, (int) open_mask);
        stk[sp + 2*off + 1] = make_int2(2*node + 1, (int) open_mask);
      }
      sp += 2*npush;
      /* queue accepted cells / direct leaves */
      unsigned dmask = isleaf ? open_mask : 0u;
      unsigned cmask = acc_mask;
      if (isleaf && ncount == 1) { dmask |= acc_mask; cmask = 0u; }     /* Tree.cpp:698-703 */
      const unsigned cb = __ballot_sync(FULLMASK, cmask != 0u);
      const unsigned db = __ballot_sync(FULLMASK, dmask != 0u);
      if (cmask) { const int o = ncb + __popc(cb & ((1u << lane) - 1u)); S.cell_node[o] = node; S.cell_mask[o] = cmask; }
      if (dmask) { const int o = ndb + __popc(db & ((1u << lane) - 1u)); S.dir_node[o] = node; S.dir_mask[o] = dmask; }
      ncb += __popc(cb); ndb += __popc(db);
      __syncwarp();
      if (ncb >= CELL_TILE) {
        consume_cells(CELL_TILE);
        const int rest = ncb - CELL_TILE;
        int tn = 0; unsigned tm = 0;
        if (lane < rest) { tn = S.cell_node[CELL_TILE + lane]; tm = S.cell_mask[CELL_TILE + lane]; }
        __syncwarp();
        if (lane < rest) { S.cell_node[lane] = tn; S.cell_mask[lane] = tm; }
        ncb = rest;
        __syncwarp();
      }
      while (ndb >= DIR_TILE) {
        consume_direct(ndb - DIR_TILE, DIR_TILE);
        ndb -= DIR_TILE;
      }
    }
    if (failed) {
      if (lane == 0) atomicExch(&st->err_stack_overflow, 1);
      continue;
    }
    if (ncb > 0) consume_cells(ncb);
    if (ndb > 0) consume_direct(0, ndb);

    /* ---- combine the 4 partial expansions of a leaf, apply, add to the particles
     *      (GradhSphTree.cpp:588-620) ---------------------------------------------------------------------*/
    if (FAST) {
#pragma unroll
      for (int o = 1; o <= 2; o <<= 1) {
        F.pot += __shfl_xor_sync(FULLMASK, F.pot, o);
#pragma unroll
        for (int k = 0; k < 3; k++) F.ac[k] += __shfl_xor_sync(FULLMASK, F.ac[k], o);
#pragma unroll
        for (int k = 0; k < 6; k++) F.q[k] += __shfl_xor_sync(FULLMASK, F.q[k], o);
        ncell_mine += __shfl_xor_sync(FULLMASK, ncell_mine, o);
      }
    }
#pragma unroll
    for (int t = 0; t < T; t++) {
      if (!tlive[t]) continue;
      if (FAST) {
        const double dx = x[t][0] - rcom[0], dy = x[t][1] - rcom[1], dz = x[t][2] - rcom[2];
        ax[t] += F.ac[0] + F.q[0]*dx + F.q[1]*dy + F.q[3]*dz;
        ay[t] += F.ac[1] + F.q[1]*dx + F.q[2]*dy + F.q[4]*dz;
        az[t] += F.ac[2] + F.q[3]*dx + F.q[4]*dy + F.q[5]*dz;
        pot[t] += F.pot + F.ac[0]*dx + F.ac[1]*dy + F.ac[2]*dz;
      }
      const int i = first + sub + 4*t;
      const unsigned fl = (unsigned) P.i[I_FLAGS][i];
      if ((fl & GSPH_FLAG_ACTIVE) && !(fl & GSPH_FLAG_DEAD)) {
        P.f[F_AX][i] += ax[t]; P.f[F_AY][i] += ay[t]; P.f[F_AZ][i] += az[t];
        P.f[F_ATX][i] += ax[t]; P.f[F_ATY][i] += ay[t]; P.f[F_ATZ][i] += az[t];
        const double gp = P.f[F_GPOT][i] + pot[t];
        P.f[F_GPOT][i] = gp;
        P.f[F_GPOTHYDRO][i] = gp;        /* gpot before the star terms (GradhSphTree.cpp:595-597, :616) */
      }
    }
    if (leaf_live && sub == 0) {
      loc_direct += (unsigned long long) ndir_mine*cnt; loc_cells += (unsigned long long) ncell_mine*cnt;
      loc_part += cnt;
    }
    __syncwarp();
    if (lane == 0) done[c] = 1;
  }
  loc_direct = (unsigned long long) warp_sum((double) loc_direct);
  loc_cells = (unsigned long long) warp_sum((double) loc_cells);
  loc_part = (unsigned long long) warp_sum((double) loc_part);
  if (lane == 0) {
    atomicAdd(&st->ndirect_sum, loc_direct); atomicAdd(&st->ncell_sum, loc_cells);
    atomicAdd(&st->nleaf_grav, loc_part);
  }
}

/*=================================================================================================
 *  Decoupled form of the same algorithm (the default path; the fused kernel above finishes whatever this one
 *  leaves, e.g. cohorts whose lists outgrow their slots).  The masked cohort walk is latency bound (dependent
 *  loads of node records), the consumption of the interaction lists is FP64-pipe bound; fused in one kernel
 *  at 128 registers per thread neither gets enough warps.  Here
 *     gravity_walk_kernel     walks (few registers, 4+ blocks per SM) and writes the cohort's lists to HBM:
 *                             cells  (node, leaf mask)            direct leaves  (first, leaf mask | count << 8)
 *     gravity_consume_kernel  streams the lists back in tiles (next tile prefetched in registers) and does the
 *                             arithmetic of consume_cells / consume_direct above,
 *  batch by batch on two streams so that the walk of batch b+1 shares the SMs with the consumption of batch b.
 *=================================================================================================*/
#define GW_THREADS 256
#define GW_WARPS (GW_THREADS/32)
#define GW_STACK 448
/* staged node records: 64 bytes = four 16-byte chunks, chunk c of record r at unit 4r + (c ^ ((r >> 1) & 3)):
   both the cooperative writes (four lanes per record, two records per quarter warp) and the per-lane reads of one
   chunk (eight records per quarter warp) touch every bank once */
#define GW_NODE_UNIT(r, c) (4*(r) + ((c) ^ (((r) >> 1) & 3)))

struct WalkShared {
  int2 stack[GW_STACK];
  double2 leafgeom[GRAV_LEAVES][3];    /* (x, y) (z, rmax) (kernrange*hmax, MAC auxiliary): three broadcast 16-byte reads per test */
  double2 nodes[32*4];                 /* the 64-byte records of the 32 nodes of a round, gathered cooperatively */
};

/* counts[2b], counts[2b+1] of batch slot b: cells, direct leaves; -1 = nothing to do here, -2 = left to the fused kernel */
template <bool GEOMAC>
__global__ void __launch_bounds__(GW_THREADS, 4)
gravity_walk_kernel(TreeDev tree, DevParams dp, int cohort_level, int c0, int nc, int* work_counter,
                    int2* __restrict__ cells, int cell_cap, int2* __restrict__ dirs, int dir_cap,
                    int* __restrict__ counts, DevStatus* st, const uint8_t* __restrict__ done, int stack_cap)
{
  extern __shared__ char walk_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  WalkShared& S = reinterpret_cast<WalkShared*>(walk_smem)[warp];
  const int leaf0 = (1 << tree.ltot) - 1;
  const int ncohort = 1 << cohort_level;
  const int nl = 1 << (tree.ltot - cohort_level);
  while (true) {
    int b = 0;
    if (lane == 0) b = atomicAdd(work_counter, 1);
    b = __shfl_sync(FULLMASK, b, 0);
    if (b >= nc) break;
    const int c = c0 + b;
    const int cnode = ncohort - 1 + c;
    int2* const mycells = cells + (size_t) b*cell_cap;
    int2* const mydirs = dirs + (size_t) b*dir_cap;
    if (tree.count[cnode] == 0 || tree.nactive[cnode] == 0 || done[c] ||
        GSPH_RANGE_OUTSIDE_SLICE(dp, tree.first[cnode], tree.count[cnode])) {
      if (lane == 0) { counts[2*b] = -1; counts[2*b + 1] = -1; }
      continue;
    }
    /* geometry of the cohort's leaves: lanes 0 .. nl-1 */
    bool leaf_live = false;
    __syncwarp();
    if (lane < GRAV_LEAVES) {
      const int myleaf = ((cnode + 1) << (tree.ltot - cohort_level)) - 1 + lane;
      if (lane < nl) {
        const int first = tree.first[myleaf], cnt = tree.count[myleaf];
        leaf_live = cnt > 0 && tree.nactive[myleaf] > 0 && !GSPH_RANGE_OUTSIDE_SLICE(dp, first, cnt);
      }
      S.leafgeom[lane][0] = leaf_live ? make_double2(tree.rcell[0][myleaf], tree.rcell[1][myleaf]) : make_double2(0.0, 0.0);
      S.leafgeom[lane][1] = leaf_live ? make_double2(tree.rcell[2][myleaf], tree.rmax[myleaf]) : make_double2(0.0, 0.0);
      S.leafgeom[lane][2] = make_double2(leaf_live ? __dmul_rn(dp.kernrange, tree.hmax[myleaf]) : 0.0,
                                         (!GEOMAC && leaf_live && tree.macaux) ? tree.macaux[myleaf] : 0.0);
    }
    const unsigned rootmask = __ballot_sync(FULLMASK, leaf_live) & 0xffu;
    int sp = 0, ncb = 0, ndb = 0, spmax = 0;
    bool failed = false;
    if (rootmask) {
      if (lane == 0) S.stack[0] = make_int2(0, (int) rootmask);
      sp = 1;
    }
    __syncwarp();
    while (sp > 0) {
      const int take = sp < 32 ? sp : 32;
      int node = -1; unsigned mask = 0;
      if (lane < take) { const int2 e = S.stack[sp - 1 - lane]; node = e.x; mask = (unsigned) e.y; }
      sp -= take;
      __syncwarp();
      /* the node records of the round: four lanes copy the four 16-byte chunks of one record (sibling nodes are
         adjacent in the heap order, so a 128-byte line serves two of them) instead of every lane fetching its
         own 64 bytes with four scattered requests.  (Keeping a second round in flight was measured: the extra
         stage costs a resident block per SM and loses more than the overlap gains.)                         */
#pragma unroll
      for (int r = 0; r < 4; r++) {
        const int src = r*8 + (lane >> 2);
        const int nsrc = __shfl_sync(FULLMASK, node, src);
        if (nsrc >= 0) cp_async16_cg(&S.nodes[GW_NODE_UNIT(src, lane & 3)], tree.gnode + 4*(size_t) nsrc + (lane & 3));
      }
      cp_async_commit();
      cp_async_wait<0>();
      __syncwarp();
      unsigned open_mask = 0, acc_mask = 0, live_mask = 0;
      bool isleaf = false; int ncount = 0, nfirst = 0;
      double rx = 0.0, ry = 0.0, rz = 0.0, crmax = 0.0, cdist = 0.0, chr = 0.0, cmac = 0.0;
      const int mac = dp.gravity_mac; const double macerror = dp.macerror;
      /* the reference's opening test per leaf (Tree.cpp:660-731, Tree.h:413-432), as in gravity_kernel */
      auto opens = [&](int l) -> bool {
        const double2 la = S.leafgeom[l][0], lb = S.leafgeom[l][1], lc = S.leafgeom[l][2];
        double dr[3];
        dr[0] = __dadd_rn(rx, -la.x); dr[1] = __dadd_rn(ry, -la.y); dr[2] = __dadd_rn(rz, -lb.x);
        const double drsqd = dot_exact<3>(dr);
        const double rmax = lb.y;
        const double a = __dadd_rn(__dadd_rn(crmax, rmax), lc.x);
        const double bb = __dadd_rn(__dadd_rn(rmax, crmax), chr);
        if (drsqd <= __dmul_rn(a, a) || drsqd <= __dmul_rn(bb, bb) || drsqd < cdist) return true;
        if (GEOMAC) return false;
        if (mac == GSPH_MAC_EIGEN) return drsqd < __dmul_rn(cmac, lc.y);
        if (mac == GSPH_MAC_GADGET2) return __dmul_rn(__dmul_rn(__dmul_rn(drsqd, drsqd), lc.y), macerror) < cmac;
        return false;
      };
      if (node >= 0) {
        const double2 g0 = S.nodes[GW_NODE_UNIT(lane, 0)], g1 = S.nodes[GW_NODE_UNIT(lane, 1)],
                      g2 = S.nodes[GW_NODE_UNIT(lane, 2)], g3 = S.nodes[GW_NODE_UNIT(lane, 3)];
        ncount = __double2hiint(g3.x); nfirst = __double2loint(g3.x);
        isleaf = node >= leaf0;
        if (ncount > 0) {
          rx = g0.x; ry = g0.y; rz = g1.x; crmax = g1.y; cdist = g2.x; chr = g2.y; cmac = g3.y;
          live_mask = mask;
        }
      }
      {
        /* leaf by leaf in a warp-uniform loop: every lane reads the SAME leaf record (a shared-memory broadcast:
           one wavefront instead of one per distinct leaf), two leaves per iteration for independent chains;
           lanes whose mask does not name the leaf compute along and discard */
        const unsigned want = __reduce_or_sync(FULLMASK, live_mask);
#pragma unroll
        for (int l0 = 0; l0 < GRAV_LEAVES; l0 += 2) {
          if (!((want >> l0) & 3u)) continue;
          const bool o0 = opens(l0), o1 = opens(l0 + 1);
          const unsigned b0 = (live_mask >> l0) & 1u, b1 = (live_mask >> (l0 + 1)) & 1u;
          open_mask |= ((o0 ? b0 : 0u) << l0) | ((o1 ? b1 : 0u) << (l0 + 1));
          acc_mask |= ((o0 ? 0u : b0) << l0) | ((o1 ? 0u : b1) << (l0 + 1));
        }
      }
      const bool push = open_mask && !isleaf;
      const unsigned pm = __ballot_sync(FULLMASK, push);
      const int npush = __popc(pm);
      if (sp + 2*npush > stack_cap) { failed = true; break; }
      if (push) {
        const int off = __popc(pm & ((1u << lane) - 1u));
        S.stack[sp + 2*off] = make_int2(2*node + 2, (int) open_mask);
        S.stack[sp + 2*off + 1] = make_int2(2*node + 1, (int) open_mask);
      }
      sp += 2*npush;
      spmax = max(spmax, sp);
      unsigned dmask = isleaf ? open_mask : 0u;
      unsigned cmask = acc_mask;
      if (isleaf && ncount == 1) { dmask |= acc_mask; cmask = 0u; }     /* Tree.cpp:698-703 */
      const unsigned cb = __ballot_sync(FULLMASK, cmask != 0u);
      const unsigned db = __ballot_sync(FULLMASK, dmask != 0u);
      if (ncb + __popc(cb) > cell_cap || ndb + __popc(db) > dir_cap) { failed = true; break; }
      if (cmask) mycells[ncb + __popc(cb & ((1u << lane) - 1u))] = make_int2(node, (int) cmask);
      if (dmask) mydirs[ndb + __popc(db & ((1u << lane) - 1u))] = make_int2(nfirst, (int)(dmask | ((unsigned) ncount << 8)));
      ncb += __popc(cb); ndb += __popc(db);
      __syncwarp();
    }
    if (lane == 0) {
      counts[2*b] = failed ? -2 : ncb; counts[2*b + 1] = failed ? -2 : ndb;
      if (failed) atomicAdd(&st->grav_left, 1);           /* the fused kernel finishes this cohort */
      else {
        if (ncb > st->grav_cell_max) atomicMax(&st->grav_cell_max, ncb);
        if (ndb > st->grav_dir_max) atomicMax(&st->grav_dir_max, ndb);
        if (spmax > st->grav_stack_max) atomicMax(&st->grav_stack_max, spmax);
      }
    }
    __syncwarp();
  }
}

/* ---- consumption of the lists: one kernel for the cells, one for the direct leaves (each fits 3 blocks per
 *      SM; together they would need 126 registers).  Both stream the cohort's list in tiles through a two-stage
 *      register pipeline: the list entries are fetched two tiles ahead, the records they point to one tile ahead,
 *      so that no load waits for the address another load delivers.                                        */
struct CohortTargets {                 /* this lane's leaf inside the cohort (4 lanes per leaf) */
  int first, cnt; bool leaf_live; int myleaf;
};

__device__ __forceinline__ CohortTargets cohort_targets(const TreeDev& tree, const DevParams& dp, int cohort_level, int cnode, int myl)
{
  CohortTargets ct;
  const int nl = 1 << (tree.ltot - cohort_level);
  ct.myleaf = ((cnode + 1) << (tree.ltot - cohort_level)) - 1 + myl;
  ct.first = 0; ct.cnt = 0; ct.leaf_live = false;
  if (myl < nl) {
    ct.first = tree.first[ct.myleaf]; ct.cnt = tree.count[ct.myleaf];
    ct.leaf_live = ct.cnt > 0 && tree.nactive[ct.myleaf] > 0 && !GSPH_RANGE_OUTSIDE_SLICE(dp, ct.first, ct.cnt);
  }
  return ct;
}

template <int MULTIPOLE, int T>
__global__ void __launch_bounds__(GRAV_THREADS, 3)
gravity_cells_kernel(PartView P, TreeDev tree, DevParams dp, const double2* __restrict__ geo4, int cohort_level,
                     int c0, int nc, int* work_counter, const int2* __restrict__ cells, int cell_cap,
                     const int* __restrict__ counts, DevStatus* st)
{
  constexpr bool FAST = (MULTIPOLE == GSPH_MULTIPOLE_FAST_MONOPOLE || MULTIPOLE == GSPH_MULTIPOLE_FAST_QUADRUPOLE);
  constexpr bool QUAD = (MULTIPOLE == GSPH_MULTIPOLE_QUADRUPOLE || MULTIPOLE == GSPH_MULTIPOLE_FAST_QUADRUPOLE);
  constexpr int NQ = QUAD ? 9 : 4;
  __shared__ double tile[GRAV_WARPS][NQ][CELL_TILE];       /* SoA: com x | y | z | m [| q0..q4], conflict-free columns */
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double (&S)[NQ][CELL_TILE] = tile[warp];
  const int ncohort = 1 << cohort_level;
  const int myl = lane >> 2, sub = lane & 3;
  unsigned long long loc_cells = 0;

  while (true) {
    int b = 0;
    if (lane == 0) b = atomicAdd(work_counter, 1);
    b = __shfl_sync(FULLMASK, b, 0);
    if (b >= nc) break;
    const int ncell = counts[2*b];
    if (ncell <= 0) continue;                              /* nothing to do, or left to the fused kernel */
    const int cnode = ncohort - 1 + c0 + b;
    const int2* const mycells = cells + (size_t) b*cell_cap;
    const CohortTargets ct = cohort_targets(tree, dp, cohort_level, cnode, myl);

    double x[T][3], ax[T], ay[T], az[T], pot[T];
    bool tlive[T];
#pragma unroll
    for (int t = 0; t < T; t++) {
      const int o = sub + 4*t;
      tlive[t] = ct.leaf_live && o < ct.cnt && GSPH_IN_SLICE(dp, ct.first + o);
      x[t][0] = x[t][1] = x[t][2] = 0.0; ax[t] = ay[t] = az[t] = pot[t] = 0.0;
      if (!FAST && tlive[t]) {
        const double2 a = geo4[2*(size_t)(ct.first + o)], bb = geo4[2*(size_t)(ct.first + o) + 1];
        x[t][0] = a.x; x[t][1] = a.y; x[t][2] = bb.x;
      }
    }
    double rcom[3] = {0, 0, 0};
    FastMp F;
    F.pot = 0.0;
#pragma unroll
    for (int k = 0; k < 3; k++) F.ac[k] = 0.0;
#pragma unroll
    for (int k = 0; k < 6; k++) F.q[k] = 0.0;
    if (FAST && ct.leaf_live) { rcom[0] = tree.com[0][ct.myleaf]; rcom[1] = tree.com[1][ct.myleaf]; rcom[2] = tree.com[2][ct.myleaf]; }
    int ncell_mine = 0;

    /* pipeline registers: e2 = entries two tiles ahead, e1 + records = one tile ahead */
    int2 e1 = make_int2(0, 0), e2 = make_int2(0, 0);
    double2 m0 = make_double2(0, 0), m1 = make_double2(0, 0);
    double qn[5] = {0, 0, 0, 0, 0};
    auto fetch_records = [&]() {
      m0 = tree.mnode[2*(size_t) e1.x]; m1 = tree.mnode[2*(size_t) e1.x + 1];      /* e1.x = 0 (the root) when idle */
      if (QUAD) {
#pragma unroll
        for (int k = 0; k < 5; k++) qn[k] = tree.q[k][e1.x];
      }
    };
    if (lane < ncell) e1 = mycells[lane];
    if (CELL_TILE + lane < ncell) e2 = mycells[CELL_TILE + lane];
    fetch_records();
    for (int base = 0; base < ncell; base += CELL_TILE) {
      const int n = min(CELL_TILE, ncell - base);
      __syncwarp();
      S[0][lane] = m0.x; S[1][lane] = m0.y; S[2][lane] = m1.x; S[3][lane] = m1.y;
      if (QUAD) {
#pragma unroll
        for (int k = 0; k < 5; k++) S[4 + k][lane] = qn[k];
      }
      const unsigned mm = lane < n ? (unsigned) e1.y : 0u;
      e1 = e2; e2 = make_int2(0, 0);
      if (base + 2*CELL_TILE + lane < ncell) e2 = mycells[base + 2*CELL_TILE + lane];
      if (base + CELL_TILE < ncell) fetch_records();
      __syncwarp();
      /* every leaf (4 lanes) visits only the entries whose mask names it; with fast multipoles the 4 lanes split
         them round-robin, otherwise each lane (a different target particle) takes them all */
      unsigned ebits = 0;
#pragma unroll
      for (int l = 0; l < GRAV_LEAVES; l++) {
        const unsigned bl = __ballot_sync(FULLMASK, (mm >> l) & 1u);
        if (l == myl) ebits = bl;
      }
      if (FAST) { for (int q = 0; q < sub; q++) ebits &= ebits - 1u; }
      if (FAST && !QUAD) {
        /* fast monopole: two independent entries per iteration; a missing second entry gets mass 0 */
        while (ebits) {
          const int e0 = __ffs(ebits) - 1;
          ebits &= ebits - 1u; ebits &= ebits - 1u; ebits &= ebits - 1u; ebits &= ebits - 1u;
          const bool two = ebits != 0u;
          const int e1i = two ? __ffs(ebits) - 1 : e0;
          if (two) { ebits &= ebits - 1u; ebits &= ebits - 1u; ebits &= ebits - 1u; ebits &= ebits - 1u; }
          ncell_mine += two ? 2 : 1;
          const double ax0 = S[0][e0] - rcom[0], ay0 = S[1][e0] - rcom[1], az0 = S[2][e0] - rcom[2];
          const double bx0 = S[0][e1i] - rcom[0], by0 = S[1][e1i] - rcom[1], bz0 = S[2][e1i] - rcom[2];
          const double ma = S[3][e0], mb = two ? S[3][e1i] : 0.0;
          const double ia = rsqrt_nr(ax0*ax0 + ay0*ay0 + az0*az0), ib = rsqrt_nr(bx0*bx0 + by0*by0 + bz0*bz0);
          const double ia2 = ia*ia, ib2 = ib*ib;
          const double ca = ma*ia2*ia, cb = mb*ib2*ib;
          F.pot += ma*ia; F.pot += mb*ib;
          F.ac[0] += ca*ax0; F.ac[1] += ca*ay0; F.ac[2] += ca*az0;
          F.ac[0] += cb*bx0; F.ac[1] += cb*by0; F.ac[2] += cb*bz0;
          F.q[0] += ca*(3.0*ax0*ax0*ia2 - 1.0); F.q[1] += ca*(3.0*ax0*ay0*ia2); F.q[2] += ca*(3.0*ay0*ay0*ia2 - 1.0);
          F.q[3] += ca*(3.0*az0*ax0*ia2);       F.q[4] += ca*(3.0*az0*ay0*ia2); F.q[5] += ca*(3.0*az0*az0*ia2 - 1.0);
          F.q[0] += cb*(3.0*bx0*bx0*ib2 - 1.0); F.q[1] += cb*(3.0*bx0*by0*ib2); F.q[2] += cb*(3.0*by0*by0*ib2 - 1.0);
          F.q[3] += cb*(3.0*bz0*bx0*ib2);       F.q[4] += cb*(3.0*bz0*by0*ib2); F.q[5] += cb*(3.0*bz0*bz0*ib2 - 1.0);
        }
      }
      while (ebits) {
        const int e = __ffs(ebits) - 1;
        ebits &= ebits - 1u;
        if (FAST) { ebits &= ebits - 1u; ebits &= ebits - 1u; ebits &= ebits - 1u; }
        ncell_mine++;
        const double cx = S[0][e], cy = S[1][e], cz = S[2][e], cm = S[3][e];
        double q0 = 0, q1 = 0, q2 = 0, q3 = 0, q4 = 0;
        if (QUAD) { q0 = S[4][e]; q1 = S[5][e]; q2 = S[6][e]; q3 = S[7][e]; q4 = S[8][e]; }
        if (FAST) {
          /* Taylor coefficients about the leaf's centre of mass (NeighbourSearch.h:481-760) */
          {
            const double dx = cx - rcom[0], dy = cy - rcom[1], dz = cz - rcom[2];
            const double inv = rsqrt_nr(dx*dx + dy*dy + dz*dz);
            const double inv2 = inv*inv;
            const double mc = cm*inv2*inv;
            F.pot += cm*inv;
            F.ac[0] += mc*dx; F.ac[1] += mc*dy; F.ac[2] += mc*dz;
            F.q[0] += mc*(3.0*dx*dx*inv2 - 1.0);
            F.q[1] += mc*(3.0*dx*dy*inv2);
            F.q[2] += mc*(3.0*dy*dy*inv2 - 1.0);
            F.q[3] += mc*(3.0*dz*dx*inv2);
            F.q[4] += mc*(3.0*dz*dy*inv2);
            F.q[5] += mc*(3.0*dz*dz*inv2 - 1.0);
          }
          if (QUAD) {
            const double dx = rcom[0] - cx, dy = rcom[1] - cy, dz = rcom[2] - cz;
            const double inv = rsqrt_nr(dx*dx + dy*dy + dz*dz + GSPH_SMALL);
            const double inv2 = inv*inv;
            const double inv5 = inv2*inv2*inv;
            const double qs = q0*dx*dx + q2*dy*dy - (q0 + q2)*dz*dz + 2.0*(q1*dx*dy + q3*dx*dz + q4*dy*dz);
            const double qf = 2.5*qs*inv5*inv2;
            double qx0 = (q0*dx + q1*dy + q3*dz)*inv5;
            double qx1 = (q1*dx + q2*dy + q4*dz)*inv5;
            double qx2 = (q3*dx + q4*dy - (q0 + q2)*dz)*inv5;
            F.pot += 0.5*qs*inv5;
            F.ac[0] += qx0 - qf*dx; F.ac[1] += qx1 - qf*dy; F.ac[2] += qx2 - qf*dz;
            qx0 *= 5.0*inv2; qx1 *= 5.0*inv2; qx2 *= 5.0*inv2;
            F.q[0] += qf*(7.0*dx*dx*inv2 - 1.0) - (qx0*dx + qx0*dx - q0*inv5);
            F.q[1] += qf*(7.0*dx*dy*inv2) - (qx0*dy + qx1*dx - q1*inv5);
            F.q[2] += qf*(7.0*dy*dy*inv2 - 1.0) - (qx1*dy + qx1*dy - q2*inv5);
            F.q[3] += qf*(7.0*dx*dz*inv2) - (qx0*dz + qx2*dx - q3*inv5);
            F.q[4] += qf*(7.0*dy*dz*inv2) - (qx1*dz + qx2*dy - q4*inv5);
            F.q[5] += qf*(7.0*dz*dz*inv2 - 1.0) - (qx2*dz + qx2*dz + (q0 + q2)*inv5);
          }
        }
        else {
#pragma unroll
          for (int t = 0; t < T; t++) {
            if (!tlive[t]) continue;
            const double dx = x[t][0] - cx, dy = x[t][1] - cy, dz = x[t][2] - cz;
            const double inv = rsqrt_nr(dx*dx + dy*dy + dz*dz + GSPH_SMALL);
            const double inv2 = inv*inv;
            const double f = cm*inv2*inv;
            if (!QUAD) {                                       /* NeighbourSearch.h:349-377 */
              pot[t] += cm*inv;
              ax[t] -= f*dx; ay[t] -= f*dy; az[t] -= f*dz;
            }
            else {                                             /* NeighbourSearch.h:384-475 */
              const double inv5 = inv2*inv2*inv;
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
    }

    /* combine the 4 partial expansions of a leaf, apply (GradhSphTree.cpp:588-620) and add to the particles */
    if (FAST) {
#pragma unroll
      for (int o = 1; o <= 2; o <<= 1) {
        F.pot += __shfl_xor_sync(FULLMASK, F.pot, o);
#pragma unroll
        for (int k = 0; k < 3; k++) F.ac[k] += __shfl_xor_sync(FULLMASK, F.ac[k], o);
#pragma unroll
        for (int k = 0; k < 6; k++) F.q[k] += __shfl_xor_sync(FULLMASK, F.q[k], o);
        ncell_mine += __shfl_xor_sync(FULLMASK, ncell_mine, o);
      }
    }
#pragma unroll
    for (int t = 0; t < T; t++) {
      if (!tlive[t]) continue;
      const int i = ct.first + sub + 4*t;
      if (FAST) {
        const double2 a = geo4[2*(size_t) i], bb = geo4[2*(size_t) i + 1];
        const double dx = a.x - rcom[0], dy = a.y - rcom[1], dz = bb.x - rcom[2];
        ax[t] = F.ac[0] + F.q[0]*dx + F.q[1]*dy + F.q[3]*dz;
        ay[t] = F.ac[1] + F.q[1]*dx + F.q[2]*dy + F.q[4]*dz;
        az[t] = F.ac[2] + F.q[3]*dx + F.q[4]*dy + F.q[5]*dz;
        pot[t] = F.pot + F.ac[0]*dx + F.ac[1]*dy + F.ac[2]*dz;
      }
      const unsigned fl = (unsigned) P.i[I_FLAGS][i];
      if ((fl & GSPH_FLAG_ACTIVE) && !(fl & GSPH_FLAG_DEAD)) {
        P.f[F_AX][i] += ax[t]; P.f[F_AY][i] += ay[t]; P.f[F_AZ][i] += az[t];
        P.f[F_ATX][i] += ax[t]; P.f[F_ATY][i] += ay[t]; P.f[F_ATZ][i] += az[t];
        P.f[F_GPOT][i] += pot[t];
      }
    }
    if (ct.leaf_live && sub == 0) loc_cells += (unsigned long long) ncell_mine*ct.cnt;
    __syncwarp();
  }
  loc_cells = (unsigned long long) warp_sum((double) loc_cells);
  if (lane == 0) atomicAdd(&st->ncell_sum, loc_cells);
}

/* The direct leaves; runs after gravity_cells_kernel of the same batch (same stream) and closes the cohort
 * (gpot_hydro, done flag, statistics).  Target leaf by target leaf: the LANES ARE THE SOURCE PARTICLES.  The entries
 * of the cohort's list that name the target leaf are compacted (one ballot per chunk of 32 entries) into a ring of
 * (first particle, count); a round takes the next 32 / W of them, W lanes per source leaf (W = particles of the
 * largest leaf: consecutive lanes read consecutive particles, coalesced), and each lane evaluates its source against
 * the (up to) four target particles of the leaf, four independent chains; the per-lane partial sums are added up
 * over the warp at the end of the target leaf.  Every lane is busy whatever the leaf masks of the entries look
 * like -- with lanes as targets the warp waits for the leaf with most entries in every tile.  (Expanding the entries
 * into a ring of particle ids first cost a scan and a divergent fill loop per chunk and leaf, a quarter of the
 * kernel's instructions; one lane per source LEAF instead was measured too: its loads do not coalesce, +13 ms.) */

/* sum of 16 per-lane values over the warp in 16 exchanges instead of 80: every step halves the number of values a
   lane carries (it keeps one half and hands the other to its partner).  Afterwards a[0] of lane L is the warp
   total of value  v(L) = 8*bit4(L) + 4*bit3(L) + 2*bit2(L) + bit1(L)  (both lanes of a pair hold it).      */
__device__ __forceinline__ double warp_reduce16(double (&a)[16], int lane)
{
#pragma unroll
  for (int n = 8, o = 16; n >= 1; n >>= 1, o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < n; i++) {
      const double keep = up ? a[i + n] : a[i], send = up ? a[i] : a[i + n];
      a[i] = keep + __shfl_xor_sync(FULLMASK, send, o);
    }
  }
  return a[0] + __shfl_xor_sync(FULLMASK, a[0], 1);
}

template <int NTG>                       /* target groups of 4 particles per leaf: 1 (leaves <= 4) or 2 (<= 8) */
__global__ void __launch_bounds__(GRAV_THREADS, 3)
gravity_direct_kernel(PartView P, TreeDev tree, DevParams dp, const double2* __restrict__ geo4, int cohort_level,
                      int c0, int nc, int* work_counter, const int2* __restrict__ dirs, int dir_cap,
                      const int* __restrict__ counts, DevStatus* st, uint8_t* done, int W)
{
  constexpr int NP = 32*NTG;                                /* particles of a cohort: 8 leaves x 4 NTG */
  __shared__ int2 queue_s[GRAV_WARPS][64];                  /* compacted (first, count) entries of the current leaf */
  __shared__ double2 tgt_s[GRAV_WARPS][NP][2];              /* (x, y) (z, hrangesqd) of the cohort's particles: broadcast reads */
  __shared__ double res_s[GRAV_WARPS][NP][4];               /* their sums ax, ay, az, pot */
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int2* const queue = queue_s[warp];
  /* W = particles of the largest leaf: a round serves EPR = 32 / W source leaves with W lanes each */
  const int EPR = 32/W, myslot = lane/W, myk = lane - myslot*W;
  double2 (&tgt)[NP][2] = tgt_s[warp];
  double (&res)[NP][4] = res_s[warp];
  const int ncohort = 1 << cohort_level;
  const int nl = 1 << (tree.ltot - cohort_level);
  const double* __restrict__ hr2a = P.f[F_HR2];
  unsigned long long loc_direct = 0, loc_part = 0;

  while (true) {
    int b = 0;
    if (lane == 0) b = atomicAdd(work_counter, 1);
    b = __shfl_sync(FULLMASK, b, 0);
    if (b >= nc) break;
    const int ndir = counts[2*b + 1];
    if (ndir < 0) continue;                                /* nothing to do, or left to the fused kernel */
    const int c = c0 + b;
    const int cnode = ncohort - 1 + c;
    const int2* const mydirs = dirs + (size_t) b*dir_cap;

    /* ---- everything the cohort needs from memory, requested at once: the leaves' ranges (lanes 0..nl-1), the
     *      particles as targets, their flags.  An absent, foreign or inactive target gets an infinite kernel
     *      range: every pair then counts as "kernel overlap" and contributes nothing here.                  */
    const int cfirst = tree.first[cnode], ccount = tree.count[cnode];
    int lfirst = 0, lcnt = 0;
    if (lane < nl) {
      const int lf = ((cnode + 1) << (tree.ltot - cohort_level)) - 1 + lane;
      lfirst = tree.first[lf]; lcnt = tree.count[lf];
      if (tree.nactive[lf] <= 0 || GSPH_RANGE_OUTSIDE_SLICE(dp, lfirst, lcnt)) lcnt = 0;
    }
    __syncwarp();
    bool mine[NTG];
#pragma unroll
    for (int r = 0; r < NTG; r++) {
      const int o = lane + 32*r, i = cfirst + o;
      double2 a = make_double2(0.0, 0.0), bb = make_double2(0.0, GSPH_BIG*GSPH_BIG);
      mine[r] = false;
      if (o < ccount && GSPH_IN_SLICE(dp, i)) {
        const unsigned fl = (unsigned) P.i[I_FLAGS][i];
        mine[r] = (fl & GSPH_FLAG_ACTIVE) && !(fl & GSPH_FLAG_DEAD);
        a = geo4[2*(size_t) i]; bb = geo4[2*(size_t) i + 1]; bb.y = hr2a[i];
      }
      tgt[o][0] = a; tgt[o][1] = bb;
      res[o][0] = res[o][1] = res[o][2] = res[o][3] = 0.0;
    }
    int2 enext = make_int2(0, 0);
    if (lane < ndir) enext = mydirs[lane];                  /* first chunk of the list: the same for every leaf */
    const int2 efirst = enext;
    __syncwarp();

    for (int l = 0; l < nl; l++) {
      const int first = __shfl_sync(FULLMASK, lfirst, l), cnt = __shfl_sync(FULLMASK, lcnt, l);
      if (cnt <= 0) continue;
#pragma unroll 1
      for (int tg = 0; tg < NTG; tg++) {
        if (4*tg >= cnt) break;
        const int tb0 = first - cfirst + 4*tg;              /* the four targets of this group inside the cohort */
        double acc[16];                                     /* [target][ax, ay, az, pot] */
#pragma unroll
        for (int k = 0; k < 16; k++) acc[k] = 0.0;
        int qhead = 0, qn = 0, nsrc = 0;
        /* software pipeline: the sources of the next round are loaded while the current ones are evaluated */
        double px = 0.0, py = 0.0, pz = 0.0, pm = 0.0, ph = -1.0;
        bool pend = false;
        auto evaluate = [&]() {
#pragma unroll
          for (int t = 0; t < 4; t++) {
            const double2 ta = tgt[tb0 + t][0], tb = tgt[tb0 + t][1];
            double dr[3] = {px - ta.x, py - ta.y, pz - tb.x};
            const double d2 = dot_exact<3>(dr);
            const double mmj = mass_unless_overlap(d2, tb.y, ph, pm);   /* kernel-overlap pairs belong to hydro.cu */
            const double inv = rsqrt_nr(d2 + GSPH_SMALL);               /* ComputeDirectGravForces */
            const double f = mmj*inv*inv*inv;
            acc[4*t] += f*dr[0]; acc[4*t + 1] += f*dr[1]; acc[4*t + 2] += f*dr[2];
            acc[4*t + 3] += mmj*inv;
          }
        };
        /* one round: the next (up to) EPR queued source leaves, W lanes each -- lane -> (leaf myslot, particle myk):
           consecutive lanes read consecutive particles, coalesced like a slab; a leaf with fewer than W particles
           leaves lanes idle (none in a tree with full leaves) */
        auto round = [&](int n) {
          double xj = 0.0, yj = 0.0, zj = 0.0, mj = 0.0, hj = -1.0;
          if (myslot < n) {
            const int2 q = queue[(qhead + myslot) & 63];
            if (myk < q.y) {
              const int j = q.x + myk;
              const double2 a = geo4[2*(size_t) j], bb = geo4[2*(size_t) j + 1];      /* m = 0 for dead particles */
              xj = a.x; yj = a.y; zj = bb.x; mj = bb.y; hj = hr2a[j];
            }
          }
          qhead = (qhead + n) & 63; qn -= n;
          if (pend) evaluate();
          px = xj; py = yj; pz = zj; pm = mj; ph = hj; pend = true;
        };
        enext = efirst;
        for (int base = 0; base < ndir; base += 32) {
          const int2 e = enext;
          enext = make_int2(0, 0);
          if (base + 32 + lane < ndir) enext = mydirs[base + 32 + lane];        /* next chunk of entries on its way */
          const bool take = (e.y >> l) & 1;                                       /* (zero beyond the list) */
          const unsigned tbal = __ballot_sync(FULLMASK, take);
          if (take) {
            queue[(qhead + qn + __popc(tbal & ((1u << lane) - 1u))) & 63] = make_int2(e.x, (e.y >> 8) & 0xff);
            nsrc += (e.y >> 8) & 0xff;
          }
          qn += __popc(tbal);
          __syncwarp();
          while (qn >= EPR) round(EPR);
          __syncwarp();
        }
        if (qn > 0) round(qn);
        __syncwarp();                                       /* the ring is rewritten by the next leaf */
        if (pend) evaluate();
        nsrc = __reduce_add_sync(FULLMASK, nsrc);
        /* sum over the lanes; the even lane L holds value v(L) = (target, component) */
        const double tot = warp_reduce16(acc, lane);
        const int v = ((lane >> 4) & 1)*8 + ((lane >> 3) & 1)*4 + ((lane >> 2) & 1)*2 + ((lane >> 1) & 1);
        if (!(lane & 1) && 4*tg + (v >> 2) < cnt) res[tb0 + (v >> 2)][v & 3] = tot;
        if (tg == 0 && lane == 0) { loc_direct += (unsigned long long) nsrc*cnt; loc_part += cnt; }
      }
    }
    __syncwarp();
    /* ---- add to the particles: one round of read-modify-writes per cohort ---------------------------------*/
#pragma unroll
    for (int r = 0; r < NTG; r++) {
      if (!mine[r]) continue;
      const int o = lane + 32*r, i = cfirst + o;
      const double sx = res[o][0], sy = res[o][1], sz = res[o][2], sp = res[o][3];
      P.f[F_AX][i] += sx; P.f[F_AY][i] += sy; P.f[F_AZ][i] += sz;
      P.f[F_ATX][i] += sx; P.f[F_ATY][i] += sy; P.f[F_ATZ][i] += sz;
      const double gp = P.f[F_GPOT][i] + sp;
      P.f[F_GPOT][i] = gp;
      P.f[F_GPOTHYDRO][i] = gp;          /* gpot before the star terms (GradhSphTree.cpp:595-597, :616) */
    }
    __syncwarp();
    if (lane == 0) done[c] = 1;
  }
  loc_direct = (unsigned long long) warp_sum((double) loc_direct);
  loc_part = (unsigned long long) warp_sum((double) loc_part);
  if (lane == 0) { atomicAdd(&st->ndirect_sum, loc_direct); atomicAdd(&st->nleaf_grav, loc_part); }
}

/* tree-order range [a, b) of node k of `level` in the balanced tree (sizes N -> N/2, N - N/2) */
static void node_range(int64_t N, int level, int k, int64_t& a, int64_t& b)
{
  a = 0; b = N;
  for (int bit = level - 1; bit >= 0; bit--) {
    const int64_t left = (b - a)/2;
    if ((k >> bit) & 1) a += left; else b = a + left;
  }
}

template <int MULTIPOLE, int T, bool GEOMAC>
int launch_gravity_decoupled(gsph_handle* h)
{
  const int Lg = h->tree.ltot > 3 ? h->tree.ltot - 3 : 0;
  const int ncohort = 1 << Lg;
  /* cohorts that can hold targets of this device (multi-GPU: those meeting the slice) */
  int c_lo = 0, c_hi = ncohort;
  if (h->slice_mode) {
    auto first_of = [&](int k) { int64_t a, b; node_range(h->N, Lg, k, a, b); return a; };
    int lo = 0, hi = ncohort;                            /* first cohort whose range ends after slice_lo */
    while (lo < hi) { const int mid = (lo + hi)/2; int64_t a, b; node_range(h->N, Lg, mid, a, b); if (b <= h->dp.slice_lo) lo = mid + 1; else hi = mid; }
    c_lo = lo;
    lo = c_lo; hi = ncohort;                             /* first cohort that starts at or after slice_hi */
    while (lo < hi) { const int mid = (lo + hi)/2; if (first_of(mid) < h->dp.slice_hi) lo = mid + 1; else hi = mid; }
    c_hi = lo;
  }
  const int ncs = c_hi - c_lo;
  if (ncs <= 0) return GSPH_OK;
  if (!h->aux_stream) {
    CUDA_TRY(h, cudaStreamCreateWithFlags(&h->aux_stream, cudaStreamNonBlocking));
    for (int k = 0; k < 5; k++) CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_grav[k], cudaEventDisableTiming));
  }
  const char* e1 = getenv("GSPH_GRAV_CELL_CAP"); const char* e2 = getenv("GSPH_GRAV_DIR_CAP"); const char* e3 = getenv("GSPH_GRAV_LIST_MB");
  const int cell_cap = e1 ? std::max(32, atoi(e1)) : 4096, dir_cap = e2 ? std::max(16, atoi(e2)) : 1536;
  const size_t per = ((size_t) cell_cap + dir_cap)*sizeof(int2) + 2*sizeof(int);
  /* list budget: 16 GB (fewer, larger batches: 64 Mi particles 195 -> 191 ms against 6 GB), but never more than a
     quarter of what is free on the device when the lists are first allocated */
  size_t budget = (e3 ? (size_t) atol(e3) : (size_t) 16384) << 20;
  if (!e3) {
    if (h->gravlist_budget == 0) {
      size_t fr = 0, tot = 0;
      h->gravlist_budget = budget;
      if (cudaMemGetInfo(&fr, &tot) == cudaSuccess) h->gravlist_budget = std::min(budget, std::max((size_t) 2048 << 20, (fr + h->gravlist_bytes)/4));
    }
    budget = h->gravlist_budget;
  }
  int Bc = (int) std::min<size_t>((size_t) ncs, std::max<size_t>(budget/2/per, 64));
  const int nbatch = (ncs + Bc - 1)/Bc;
  Bc = (ncs + nbatch - 1)/nbatch;                          /* even batches */
  const int nbuf = nbatch > 1 ? 2 : 1;
  const size_t need = (size_t) nbuf*Bc*per;
  if (need > h->gravlist_bytes) {
    if (h->gravlist) cudaFree(h->gravlist);
    h->gravlist = nullptr; h->gravlist_bytes = 0;
    CUDA_TRY(h, cudaMalloc(&h->gravlist, need));
    h->gravlist_bytes = need;
  }
  const char* tcap = getenv("GSPH_TEST_GRAV_STACK");
  const int stack_cap = tcap ? std::max(40, std::min(GW_STACK, atoi(tcap))) : GW_STACK;
  PartView P;
  for (int k = 0; k < NF_DOUBLE; k++) P.f[k] = h->P[k];
  for (int k = 0; k < NF_INT; k++) P.i[k] = h->PI[k];
  auto walk = gravity_walk_kernel<GEOMAC>;
  auto kcells = gravity_cells_kernel<MULTIPOLE, T>;
  auto kdirect = gravity_direct_kernel<T>;     /* T = groups of 4 targets per leaf */
  int wbps = 0, cbps = 0, dbps = 0;
  const size_t wsmem = GW_WARPS*sizeof(WalkShared);
  CUDA_TRY(h, cudaFuncSetAttribute(walk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) wsmem));
  CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&wbps, walk, GW_THREADS, wsmem));
  CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cbps, kcells, GRAV_THREADS, 0));
  CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&dbps, kdirect, GRAV_THREADS, 0));
  wbps = std::max(wbps, 1); cbps = std::max(cbps, 1); dbps = std::max(dbps, 1);
  /* the two kernels of neighbouring batches share the SMs: resident blocks per SM of each (tuning knobs) */
  { const char* ew = getenv("GSPH_GRAV_WALK_BPS"); const char* ec = getenv("GSPH_GRAV_CONS_BPS");
    if (ew) wbps = std::max(1, std::min(wbps, atoi(ew)));
    if (ec) { cbps = std::max(1, std::min(cbps, atoi(ec))); dbps = std::max(1, std::min(dbps, atoi(ec))); } }
  KernelTimer kt(h, 2);
  cudaStream_t sc = h->stream, sw = h->aux_stream;
  CUDA_TRY(h, cudaEventRecord(h->ev_grav[4], sc));           /* everything before (done flags, hydro half) */
  CUDA_TRY(h, cudaStreamWaitEvent(sw, h->ev_grav[4], 0));
  for (int b = 0; b < nbatch; b++) {
    const int buf = b & (nbuf - 1);
    const int c0 = c_lo + b*Bc, nc = std::min(Bc, c_hi - c0);
    char* base = (char*) h->gravlist + (size_t) buf*Bc*per;
    int2* cells = (int2*) base;
    int2* dirs = cells + (size_t) Bc*cell_cap;
    int* counts = (int*)(dirs + (size_t) Bc*dir_cap);
    if (b >= nbuf) CUDA_TRY(h, cudaStreamWaitEvent(sw, h->ev_grav[2 + buf], 0));      /* lists of batch b - 2 consumed */
    CUDA_TRY(h, cudaMemsetAsync(h->work_counter + 1, 0, sizeof(int), sw));
    const int wblocks = std::min(h->sm_count*wbps, (nc + GW_WARPS - 1)/GW_WARPS);
    walk<<<wblocks, GW_THREADS, wsmem, sw>>>(h->tree, h->dp, Lg, c0, nc, h->work_counter + 1, cells, cell_cap, dirs, dir_cap,
                                         counts, h->status, h->done, stack_cap);
    CUDA_TRY(h, cudaEventRecord(h->ev_grav[buf], sw));
    CUDA_TRY(h, cudaStreamWaitEvent(sc, h->ev_grav[buf], 0));
    CUDA_TRY(h, cudaMemsetAsync(h->work_counter + 2, 0, 2*sizeof(int), sc));
    const int cblocks = std::min(h->sm_count*cbps, (nc + GRAV_WARPS - 1)/GRAV_WARPS);
    kcells<<<cblocks, GRAV_THREADS, 0, sc>>>(P, h->tree, h->dp, (const double2*) h->geo4, Lg, c0, nc, h->work_counter + 2,
                                             cells, cell_cap, counts, h->status);
    const int dblocks = std::min(h->sm_count*dbps, (nc + GRAV_WARPS - 1)/GRAV_WARPS);
    kdirect<<<dblocks, GRAV_THREADS, 0, sc>>>(P, h->tree, h->dp, (const double2*) h->geo4, Lg, c0, nc, h->work_counter + 3,
                                              dirs, dir_cap, counts, h->status, h->done,
                                              (int) std::max<int64_t>(1, (h->N + ((int64_t) 1 << h->tree.ltot) - 1) >> h->tree.ltot));
    CUDA_TRY(h, cudaEventRecord(h->ev_grav[2 + buf], sc));
    h->stats.kernel_launches += 3;
  }
  CUDA_TRY(h, cudaGetLastError());
  return GSPH_OK;
}

template <int MULTIPOLE, int T, bool GEOMAC>
int launch_gravity_t(gsph_handle* h)
{
  const size_t smem = GRAV_WARPS*sizeof(WarpSharedT<(MULTIPOLE == GSPH_MULTIPOLE_QUADRUPOLE || MULTIPOLE == GSPH_MULTIPOLE_FAST_QUADRUPOLE)>);
  auto kern = gravity_kernel<MULTIPOLE, T, GEOMAC>;
  CUDA_TRY(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  int bps = 0;
  CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, GRAV_THREADS, smem));
  if (bps < 1) bps = 1;
  int blocks = h->sm_count*bps;
  const int Lg = h->tree.ltot > 3 ? h->tree.ltot - 3 : 0;
  const int ncohort = 1 << Lg;
  const int need_blocks = (ncohort + GRAV_WARPS - 1)/GRAV_WARPS;
  if (blocks > need_blocks) blocks = need_blocks;
  CUDA_TRY(h, cudaMemsetAsync(h->work_counter, 0, sizeof(int), h->stream));
  PartView P;
  for (int k = 0; k < NF_DOUBLE; k++) P.f[k] = h->P[k];
  for (int k = 0; k < NF_INT; k++) P.i[k] = h->PI[k];
  KernelTimer kt(h, 2);
  int2* gstack = nullptr;
  /* GSPH_TEST_GRAV_STACK=n shrinks the shared-memory stack so that the tests can exercise the retry */
  const char* tcap = getenv("GSPH_TEST_GRAV_STACK");
  const int smem_stack_cap = tcap ? std::max(40, std::min(GRAV_STACK, atoi(tcap))) : GRAV_STACK;
  if (h->gstack_on) {                                  /* retry after a stack overflow: stacks in HBM */
    if (h->gstack_warps < blocks*GRAV_WARPS) {
      if (h->gstack) cudaFree(h->gstack);
      h->gstack = nullptr; h->gstack_warps = 0;
      CUDA_TRY(h, cudaMalloc(&h->gstack, (size_t) blocks*GRAV_WARPS*h->gstack_cap*sizeof(int2)));
      h->gstack_warps = blocks*GRAV_WARPS;
    }
    gstack = h->gstack;
  }
  kern<<<blocks, GRAV_THREADS, smem, h->stream>>>(P, h->tree, h->dp, Lg, h->work_counter, h->status, h->done,
                                                  gstack, (int) h->gstack_cap, smem_stack_cap);
  h->stats.kernel_launches++;
  CUDA_TRY(h, cudaGetLastError());
  return GSPH_OK;
}

template <int MULTIPOLE>
int launch_gravity(gsph_handle* h, int decoupled)
{
  /* largest leaf of the balanced tree: ceil(N / 2^ltot); 4 lanes per leaf -> T particles per lane */
  const int64_t maxleaf = (h->N + ((int64_t) 1 << h->tree.ltot) - 1) >> h->tree.ltot;
  if (decoupled) {
    if (h->dp.gravity_mac == GSPH_MAC_GEOMETRIC)
      return maxleaf <= 4 ? launch_gravity_decoupled<MULTIPOLE, 1, true>(h) : launch_gravity_decoupled<MULTIPOLE, 2, true>(h);
    return maxleaf <= 4 ? launch_gravity_decoupled<MULTIPOLE, 1, false>(h) : launch_gravity_decoupled<MULTIPOLE, 2, false>(h);
  }
  if (h->dp.gravity_mac == GSPH_MAC_GEOMETRIC)
    return maxleaf <= 4 ? launch_gravity_t<MULTIPOLE, 1, true>(h) : launch_gravity_t<MULTIPOLE, 2, true>(h);
  return maxleaf <= 4 ? launch_gravity_t<MULTIPOLE, 1, false>(h) : launch_gravity_t<MULTIPOLE, 2, false>(h);
}

}  // namespace

/* decoupled = 1: walk + consume kernels over batches of cohorts (the default first pass); 0: the fused kernel,
   which skips every cohort already done and finishes the rest (list or stack overflow of the first pass) */
int gsph_run_gravity(gsph_handle* h, int decoupled)
{
  if (h->dp.ndim != 3) { h->err = "self-gravity on the accelerated path is 3-D only"; return GSPH_EINVAL; }
  if (h->cfg.Nleafmax > GRAV_NT) { h->err = "self-gravity on the accelerated path needs Nleafmax <= 8"; return GSPH_EINVAL; }
  switch (h->dp.multipole) {
  case GSPH_MULTIPOLE_MONOPOLE: return launch_gravity<GSPH_MULTIPOLE_MONOPOLE>(h, decoupled);
  case GSPH_MULTIPOLE_QUADRUPOLE: return launch_gravity<GSPH_MULTIPOLE_QUADRUPOLE>(h, decoupled);
  case GSPH_MULTIPOLE_FAST_MONOPOLE: return launch_gravity<GSPH_MULTIPOLE_FAST_MONOPOLE>(h, decoupled);
  default: return launch_gravity<GSPH_MULTIPOLE_FAST_QUADRUPOLE>(h, decoupled);
  }
}
