/*=================================================================================================
 *  gsph_internal.cuh -- shared declarations of libgsph (B200 / sm_100a grad-h SPH neighbour path).
 *
 *  Data layout in HBM
 *  ------------------
 *  Particles live twice on the device:
 *    - `host order`  staging arrays (H*): exactly the order of the host array handed to gsph_upload;
 *    - `tree order`  SoA arrays     (P*): particles permuted so that every KD-tree cell of every
 *      level is ONE contiguous slab [first, first+count).  All kernels work on tree order; `perm`
 *      maps tree position -> host index.
 *  The KD-tree is the reference's balanced tree (KDTree.cpp:220-595): a perfect binary tree of depth
 *  ltot whose node sizes follow N -> (N/2, N - N/2).  Nodes are stored in HEAP order
 *  (node = 2^level - 1 + k); the reference's depth-first ids are only produced for the inspection
 *  API (gsph_tree_cells).
 *=================================================================================================*/
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include "../../include/gsph.h"

#define GSPH_TABLE_RES 1000
#define GSPH_SMALL 1.0e-20           /* small_number, src/Headers/Constants.h:73 */
#define GSPH_BIG   9.9e20            /* big_number,   src/Headers/Constants.h:72 */
#define GSPH_INVPI 0.31830988618379  /* invpi as spelled in src/Headers/Constants.h:63 */

/* particle fields, double precision, tree order and host order */
enum PField {
  F_RX, F_RY, F_RZ, F_VX, F_VY, F_VZ, F_AX, F_AY, F_AZ, F_ATX, F_ATY, F_ATZ,
  F_M, F_H, F_U, F_RHO, F_HR2, F_HFAC, F_INVOMEGA, F_ZETA, F_SOUND, F_PRESS,
  F_DUDT, F_DIVV, F_ALPHA, F_DALPHADT, F_GPOT, F_GPOTHYDRO,
  NF_DOUBLE
};
enum IField { I_FLAGS, I_PTYPE, I_LEVEL, I_LEVELNEIB, I_IORIG, I_SINKID, NF_INT };

/* per-node fields of the tree (heap order) */
struct TreeDev {
  int ltot = 0;            /* leaf level */
  int ncell = 0;           /* 2^(ltot+1) - 1 */
  int ncell_cap = 0;
  int* first = nullptr;    /* first tree-order particle of the node */
  int* count = nullptr;    /* number of particles */
  int* nactive = nullptr;  /* active particles below the node */
  double* bbmin[3] = {0, 0, 0};   /* tight particle box                  (TreeCell::bb)    */
  double* bbmax[3] = {0, 0, 0};
  double* hbmin[3] = {0, 0, 0};   /* box of r -+ kernrange*h             (TreeCell::hbox)  */
  double* hbmax[3] = {0, 0, 0};
  double* rcell[3] = {0, 0, 0};   /* geometric centre of bb              (TreeCell::rcell) */
  double* com[3] = {0, 0, 0};     /* centre of mass                      (TreeCell::r)     */
  double* vcom[3] = {0, 0, 0};    /* mass-weighted velocity              (TreeCell::v)     */
  double* m = nullptr;
  double* rmax = nullptr;         /* half diagonal of bb                                   */
  double* hmax = nullptr;
  double* cdistsqd = nullptr;     /* geometric MAC distance^2, KDTree.cpp:917              */
  double* q[5] = {0, 0, 0, 0, 0}; /* traceless quadrupole, KDTree.cpp:937-1000             */
  double* macaux = nullptr;       /* gadget2: min |atree| below the node (TreeCell::amin); eigenmac:
                                     max gpot^(-2/3) (TreeCell::macfactor); KDTree.cpp:899-903,967-984;
                                     allocated only for those criteria                          */
  /* single-precision node records for the neighbour walks (candidate sets only have to be
     supersets, so boxes rounded OUTWARD to float are enough and cost a third of the loads):
       nodeb[2n]   = (bbmin.xyz, first as int bits)   nodeb[2n+1] = (bbmax.xyz, count as int bits)
       nodeh[2n]   = (hbmin.xyz, 0)                   nodeh[2n+1] = (hbmax.xyz, 0)                 */
  float4* nodeb = nullptr;
  float4* nodeh = nullptr;
  /* exact double-precision node records of the gravity walk (one 64-byte line / one 32-byte sector
     per node instead of seven / four separate arrays):
       gnode[4n..4n+3] = (rcell.x, rcell.y) (rcell.z, rmax) (cdistsqd, kernrange*hmax) (count:first, mac)
                         count:first = the two int32 halves of one double (high: count, low: first)
                         mac = rmax^2*m for gadget2, TreeCell::mac for eigenmac (Tree.h:413-432)
       mnode[2n..2n+1] = (com.x, com.y) (com.z, m)                                               */
  double2* gnode = nullptr;
  double2* mnode = nullptr;
};

struct DevParams {
  int ndim, kernel, tabulated;
  double kernrange, kernrangesqd, kernnorm;
  double h_fac, h_converge;
  int avisc, acond, tdavisc;
  double alpha_visc, beta_visc, alpha_visc_min;
  int eos;
  double gamma_eos, gammam1, temp0, mu_bar, rho_bary;
  double boxmin[3], boxmax[3], boxsize[3], boxhalf[3];
  int periodic[3];
  int any_periodic;
  int mirror[3][2];          /* mirror wall on the lhs / rhs of dimension k                           */
  int any_mirror, nimg;      /* nimg = prod_k (1 + mirror[k][0] + mirror[k][1]): image codes, 0 = the real particle */
  int hydro_forces, self_gravity, multipole, gravity_mac;
  double thetamaxsqd, invthetamaxsqd, macerror;
  int Nleafmax;
  int create_sinks;
  int star_zeta, nbody_softening;   /* conservative_sph_star_gravity: star terms of zeta (GradhSph.cpp:288-306) */
  const double* stars; int nstar;   /* device rows {r[ndim], m, h} of gsph_set_stars                    */
  int slice_lo, slice_hi;           /* multi-GPU: only the tree-order particles [slice_lo, slice_hi) are targets of
                                       this device (gsph_set_active_slice); [0, INT_MAX) otherwise             */
};
#define GSPH_IN_SLICE(dp, p) ((p) >= (dp).slice_lo && (p) < (dp).slice_hi)
#define GSPH_RANGE_OUTSIDE_SLICE(dp, first, count) ((first) >= (dp).slice_hi || (first) + (count) <= (dp).slice_lo)

/* device-side error / statistics block */
struct DevStatus {
  int err_cand_overflow;     /* candidate scratch too small                                    */
  int err_stack_overflow;    /* walk stack too small                                           */
  int err_converge;          /* h-rho iteration did not converge (GradhSph.cpp:249-250)        */
  int err_hsearch;           /* search radius had to grow too often                            */
  int err_nlist_overflow;    /* per-lane neighbour list too small                              */
  int density_retries;
  int hiter_max;
  int ngb_max;
  unsigned long long hiter_sum, ngb_sum, cand_sum, ntarget, ndirect_sum, ncell_sum, nleaf_grav;
  unsigned long long cand_need;   /* largest candidate count requested */
  int grav_left;             /* cohorts the decoupled gravity pass left to the fused kernel */
  int grav_cell_max, grav_dir_max, grav_stack_max;   /* longest interaction lists / deepest walk stack of a cohort (decoupled gravity pass) */
};

/* device-resident main loop (gsph_integrator_*): start-of-step copies in host order */
struct IntegState {
  bool on = false;
  gsph_integrator_config cfg;
  double* s0[11] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};   /* r0[3], v0[3], a0[3], u0, dudt0              */
  int64_t cap = 0;
  unsigned long long* red = nullptr;   /* [0] min dt (bits of a positive double), [1] bad-timestep count */
  double t = 0.0, tlast = 0.0, timestep = 0.0;
  int64_t Nsteps = 0;
  bool rebuild = true;
  bool mid_step = false;     /* between gsph_integrator_advance and gsph_integrator_finish */
  bool starting = false;     /* inside gsph_integrator_begin's first force pass */
  /* hierarchical block time steps (cfg.Nlevels > 1): per-particle integer step / start of step (host order) and the
     scalars of Simulation::ComputeBlockTimesteps */
  int32_t* nstep = nullptr; int32_t* nlast = nullptr;
  double* pdt = nullptr; double* dt_next = nullptr; double* ptlast = nullptr;     /* Particle::dt, dt_next, tlast */
  int64_t blk_cap = 0;
  long long n = 0, nresync = 0;
  int level_max = 0, level_step = 0;
  double dt_max = 0.0;
  /* star particles: the N-body leapfrog (NbodyLeapfrogKDK.cpp) runs on the host around the device passes */
  int nstar = 0;
  double nbody_mult = 0.1;
  std::vector<double> sr, sv, sa, sr0, sv0, sa0, sm, sh, sgpot;
  bool hermite4 = false;     /* the stars move with NbodyHermite4 (the reference's default) instead of the leapfrog */
  std::vector<double> sadot, sadot0, sa2dot, sa3dot;
  /* block time steps of the stars (Simulation::ComputeBlockTimesteps, N-body parts): level, integer step and start,
     start time, end-of-step flag, and who is active in the current step */
  std::vector<long long> slevel, snstep, snlast;
  std::vector<double> stlast;
  std::vector<char> send, sact;
};

struct gsph_handle {
  gsph_config cfg;
  IntegState integ;
  bool forces_half_done = false;       /* gsph_update_all_forces_begin done, _finish pending */
  DevParams dp;
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  /* host<->device copies run on their own stream so that they overlap the kernels:
       ev_geom   : r, h and the integer fields of the last upload have landed (tree build may start)
       ev_all    : every field of the last upload has landed (permutation into tree order may start)
       ev_compute: the compute stream reached the point where a copy may touch the staging arrays  */
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_geom = nullptr, ev_all = nullptr, ev_compute = nullptr;
  cudaEvent_t ev_early = nullptr;   /* everything but the velocities has landed (they are uploaded last) */
  /* the velocities are not needed before the force pass: a full rebuild leaves their permutation (and the cells'
     vcom) for later, so that their upload overlaps the tree build and the density pass */
  bool v_pending = false, vcom_valid = true;
  cudaEvent_t ev_axis[3] = {nullptr, nullptr, nullptr};   /* coordinate k of the last upload has landed (its sort may start) */
  bool upload_pending = false;   /* ev_geom / ev_all still have to be waited for by the compute stream */
  std::string err;

  int64_t N = 0;             /* particles currently uploaded */
  int64_t Ncap = 0;          /* capacity of particle arrays  */
  double* H[NF_DOUBLE] = {0};   /* host-order staging */
  int32_t* HI[NF_INT] = {0};
  double* P[NF_DOUBLE] = {0};   /* tree order */
  int32_t* PI[NF_INT] = {0};
  int32_t* perm = nullptr;      /* tree position -> host index */
  /* packed tree-order records for the gather phases of the density / force kernels */
  double* hnew = nullptr;       /* [N] h written by the density pass, committed when the pass succeeded */
  double* geo4 = nullptr;       /* [N][4]  x, y, z, m          (density; m = 0 for dead particles)  */
  double* geoh = nullptr;       /* [N][4]  x, y, z, hrangesqd  (force-pass predicate; dead: x = BIG) */
  double* frec = nullptr;       /* [N][16] force record, 128 B: x y z vx | vy vz m 1/h | hfactor 1/rho
                                   P*invomega/rho^2 sound | u alpha zeta (pressure|level)           */
  bool uploaded = false, have_tree = false, sorted_valid = false, all_active = true;
  bool density_done = false;
  bool slice_mode = false;   /* gsph_set_active_slice: inactive particles belong to other ranks */
  /* device-resident runs keep the particle state in HOST order (H*) and the passes write tree order (P*):
       results_newer : a pass wrote tree-order results that have not been scattered back to host order yet
       p_stale       : the integrator changed host-order inputs (v, u at the end of a step) after the last
                       BuildTree, so an operator call would work on outdated tree-order copies          */
  bool results_newer = false, p_stale = false;
  /* which particle arrays carry information (bit = PField / IField index):
       present_* : the host handed the field to the last gsph_upload (absent ones hold the defaults)
       hdef_*    : the host-order staging array currently holds the default fill (no need to refill)
       pvalid_*  : the tree-order array is up to date (permuted by BuildTree or written by a pass); a
                   download takes the other fields straight from host order
     lazy_fields : every particle is active and alive in this step, so pure outputs need not be permuted
     acc_zeroed  : gsph_zero_accelerations ran since the last BuildTree                                   */
  unsigned long long present_d = 0, hdef_d = 0, pvalid_d = 0;
  unsigned present_i = 0, hdef_i = 0, pvalid_i = 0;
  bool lazy_fields = false, acc_zeroed = false;

  double* stars = nullptr; int Nstar = 0, stars_cap = 0;   /* device copy of {r[ndim], m, h} per star (gsph_set_stars) */

  TreeDev tree;
  double* tables = nullptr;     /* 6 x 1000 doubles: w0_s2, womega_s2, wzeta_s2, w1, wgrav, wpot */

  /* scratch */
  void* scratch = nullptr; size_t scratch_bytes = 0;          /* tree build temporaries          */
  int32_t* cand = nullptr; size_t cand_cap_per_warp = 0; int cand_warps = 0;
  int32_t* nlist = nullptr; size_t nlist_cap = 0; int nlist_warps = 0;   /* per warp: [cap][32] neighbour ids */
  void* gravlist = nullptr; size_t gravlist_bytes = 0;   /* interaction lists of the decoupled gravity pass (two batches) */
  size_t gravlist_budget = 0;   /* bytes the gravity lists may take (decided once: cudaMemGetInfo is not free) */
  cudaStream_t aux_stream = nullptr;                      /* the walk kernels of the decoupled gravity pass */
  cudaEvent_t ev_grav[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   /* walked[2], consumed[2], fork */
  bool gstack_on = false;        /* this launch is a retry with the stacks in HBM */
  int2* gstack = nullptr; size_t gstack_cap = 0; int gstack_warps = 0;   /* gravity walk stacks in HBM: only after a
                                                                            shared-memory stack overflowed */
  DevStatus* status = nullptr;  /* device */
  DevStatus* status_host = nullptr; /* pinned */
  int* work_counter = nullptr;
  uint8_t* done = nullptr; size_t done_cap = 0;   /* per cohort / leaf: finished in this pass */
  /* candidate lists of the neighbour passes, produced by lean pre-walk kernels (density_walk_kernel,
     hydro_walk_kernel) for the cohorts [pre_c0, pre_c0 + pre_n): pre_count[b] = candidates of cohort pre_c0 + b
     (bit 30: a periodic wrap was seen), -1 = the main kernel walks itself; ids at pre_cand + b*pre_cap */
  int32_t* pre_cand = nullptr; int* pre_count = nullptr; size_t pre_bytes = 0, pre_count_cap = 0;
  int pre_c0 = 0, pre_n = 0, pre_cap = 0;
  uint8_t* needed = nullptr; size_t needed_cap = 0;   /* per cohort node: may hold a neighbour of this device's slice */
  void* pinned = nullptr; size_t pinned_bytes = 0;             /* host staging for strided views  */

  gsph_stats stats;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t evt[2] = {nullptr, nullptr};          /* gsph_timer_start/stop */
  cudaEvent_t evk[6][2] = {};                        /* per-kernel stopwatches: 0 density 1 hydro 2 gravity 3 sort 4 levels 5 stock */
  bool evk_used[6] = {false, false, false, false, false, false};
};

#define CUDA_TRY(h, call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { \
  (h)->err = std::string(#call) + ": " + cudaGetErrorString(e__); return GSPH_ECUDA; } } while (0)

/* ---- module entry points (host side) -----------------------------------------------------------*/
void gsph_make_tables(int ndim, int kernel, double* out6x1000);
double gsph_kernel_grav_host(int ndim, int kernel, int tabulated, int which, double s);   /* tables.cpp */       /* tables.cu  */
double gsph_kernnorm(int ndim, int kernel);
int gsph_tree_build(gsph_handle* h, int mode, double dt);             /* tree.cu    */
int gsph_tree_build_sharded(gsph_handle* h, int phase, int rank, int world);
int gsph_tree_update_hmax(gsph_handle* h);
int gsph_tree_update_active(gsph_handle* h);
int gsph_derive_density_fields_impl(gsph_handle* h, int lo, int hi);   /* density.cu */
int gsph_thermal_all_impl(gsph_handle* h);                            /* density.cu: ComputeThermalProperties of every particle */
int gsph_run_density(gsph_handle* h);                                 /* density.cu */
int gsph_run_hydro(gsph_handle* h, int with_gravity);                 /* hydro.cu   */
int gsph_run_neiblist(gsph_handle* h, int32_t* counts_dev, const int64_t* offsets_dev, int32_t* idx_dev);
int gsph_run_gravity(gsph_handle* h, int decoupled);                  /* gravity.cu */
int gsph_run_star_forces(gsph_handle* h);
int gsph_run_star_gas(gsph_handle* h, const double* dstars, int Nstar, double* dout, int* dcounts, int* derr);   /* hydro.cu */                             /* hydro.cu   */
int gsph_check_status(gsph_handle* h, const char* where);             /* capi.cu    */
int gsph_ensure_cand(gsph_handle* h, int warps, size_t cap_per_warp);
int gsph_ensure_nlist(gsph_handle* h, int warps, size_t cap);
int gsph_gather_fields(gsph_handle* h, bool defer_v = false);         /* tree.cu: host order -> tree order */
int gsph_ensure_velocities(gsph_handle* h);                           /* tree.cu: the deferred part of it */
int gsph_ensure_vcom(gsph_handle* h);                                 /* tree.cu: ... and the cells' vcom */
int gsph_prewalk_prepare(gsph_handle* h, int cohort_level);           /* capi.cu: sizes pre_cand / pre_count; pre_n = 0: disabled */
/* fields written by the passes (for pvalid_*) */
#define GSPH_DMASK(f) (1ull << (f))
static const unsigned long long GSPH_OUT_DENSITY =
  GSPH_DMASK(F_H) | GSPH_DMASK(F_RHO) | GSPH_DMASK(F_HR2) | GSPH_DMASK(F_HFAC) | GSPH_DMASK(F_INVOMEGA) | GSPH_DMASK(F_ZETA) |
  GSPH_DMASK(F_SOUND) | GSPH_DMASK(F_PRESS) | GSPH_DMASK(F_U) | GSPH_DMASK(F_DIVV) | GSPH_DMASK(F_ALPHA) | GSPH_DMASK(F_DALPHADT);
static const unsigned long long GSPH_OUT_FORCES =
  GSPH_DMASK(F_AX) | GSPH_DMASK(F_AY) | GSPH_DMASK(F_AZ) | GSPH_DMASK(F_ATX) | GSPH_DMASK(F_ATY) | GSPH_DMASK(F_ATZ) |
  GSPH_DMASK(F_GPOT) | GSPH_DMASK(F_GPOTHYDRO) | GSPH_DMASK(F_DUDT) | GSPH_DMASK(F_DIVV) | GSPH_DMASK(F_DALPHADT);

struct KernelTimer {          /* brackets one launch (or group) with events; read back later */
  gsph_handle* h; int k;
  KernelTimer(gsph_handle* h_, int k_) : h(h_), k(k_) { cudaEventRecord(h->evk[k][0], h->stream); }
  ~KernelTimer() { cudaEventRecord(h->evk[k][1], h->stream); h->evk_used[k] = true; }
};

static inline int cohort_level_for(int64_t N, int ltot, int maxsize)
{
  /* shallowest level whose largest node has <= maxsize particles (node sizes are
     ceil/floor of N / 2^level in the balanced tree) */
  int l = 0;
  while (l < ltot && ((N + ((int64_t)1 << l) - 1) >> l) > maxsize) l++;
  return l;
}
