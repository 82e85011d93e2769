/*=================================================================================================
 *  gsph.h -- C ABI of the B200-native grad-h SPH neighbour path ("libgsph").
 *
 *  This is the drop-in boundary for ONE hot path of GANDALF v0.4.0: tree build -> h-rho
 *  iteration -> hydro / hydro+gravity force sums.  In the reference that path sits behind the
 *  C++ pure-virtual surface
 *      NeighbourSearch<ndim>        src/Headers/NeighbourSearch.h:66-137
 *      SphNeighbourSearch<ndim>     src/Headers/SphNeighbourSearch.h:77-94
 *  whose concrete class GradhSphTree<ndim,GradhSphParticle> is created in
 *      GradhSphSimulation<ndim>::ProcessSphParameters   src/GradhSph/GradhSphSimulation.cpp:232-246
 *  from the .dat parameter `neib_search`.  A host-side subclass (host/GpuSphTree.h in this
 *  repository; binding shown in INTEGRATION.md) forwards each virtual to one entry point below.
 *
 *  Conventions
 *    - plain C, no C++/torch types; every function returns 0 on success, a GSPH_E* code
 *      otherwise, never throws (the reference library is built -fno-exceptions, src/Makefile:57);
 *      gsph_last_error() gives the message the host wrapper hands to
 *      ExceptionHandler::raise (src/Common/Exception.cpp:47-69).
 *    - all entry points are called from one host thread (the reference calls the L3 surface from
 *      the main thread only, src/Hydrodynamics/SphSimulation.cpp:616-710).
 *    - FLOAT = double (the reference's default PRECISION=DOUBLE, Makefile:27).
 *    - there is NO CPU fallback: without a CUDA device gsph_create fails with GSPH_ENODEVICE.
 *=================================================================================================*/
#ifndef GSPH_H
#define GSPH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSPH_ABI_VERSION 1

/* ---- return codes ---------------------------------------------------------------------------*/
enum {
  GSPH_OK          = 0,
  GSPH_EINVAL      = 1,   /* bad argument / unsupported parameter combination                  */
  GSPH_ENODEVICE   = 2,   /* no CUDA device (there is no CPU fallback)                          */
  GSPH_ECUDA       = 3,   /* CUDA runtime error, text in gsph_last_error                        */
  GSPH_ENOMEM      = 4,
  GSPH_ESTATE      = 5,   /* call out of order (e.g. forces before tree build)                  */
  GSPH_ECONVERGE   = 6    /* h-rho iteration failed: reference raises "Problem with convergence
                             of h-rho iteration" (src/GradhSph/GradhSph.cpp:249-250)             */
};

/* ---- enumerations mirroring the reference's run-time parameters -------------------------------*/
/* kernel = m4 | quintic            (src/Common/Parameters.cpp:253; SmoothingKernel.h:131,281)    */
enum { GSPH_KERNEL_M4 = 0, GSPH_KERNEL_QUINTIC = 1, GSPH_KERNEL_GAUSSIAN = 2 /* SmoothingKernel.h:410-545 */ };
/* avisc  = none | mon97 | mon97mm97 | mon97cd2010   (src/Headers/Hydrodynamics.h aviscenum)      */
enum { GSPH_AVISC_NONE = 0, GSPH_AVISC_MON97 = 1, GSPH_AVISC_MON97MM97 = 2, GSPH_AVISC_MON97CD2010 = 3 };
/* acond  = none | wadsley2008 | price2008                                                         */
enum { GSPH_ACOND_NONE = 0, GSPH_ACOND_WADSLEY2008 = 1, GSPH_ACOND_PRICE2008 = 2 };
/* time_dependent_avisc = none | mm97 | cd2010 (cd2010 is not on the accelerated path)             */
enum { GSPH_TDAVISC_NONE = 0, GSPH_TDAVISC_MM97 = 1, GSPH_TDAVISC_CD2010 = 2 /* Sph.h:364-456 */ };
/* gas_eos = energy_eqn|constant_temp (Adiabatic) | isothermal | barotropic
 *           (src/Hydrodynamics/Hydrodynamics.cpp:91-108; src/Thermal/{Adiabatic,Isothermal,Barotropic}EOS.cpp) */
enum { GSPH_EOS_ADIABATIC = 0, GSPH_EOS_ISOTHERMAL = 1, GSPH_EOS_BAROTROPIC = 2,
       GSPH_EOS_BAROTROPIC2 = 3 /* src/Thermal/Barotropic2EOS.cpp:69-91 */ };
/* boundary_lhs/rhs = open | periodic | mirror (src/Headers/DomainBox.h:49-92)                     */
enum { GSPH_BOUNDARY_OPEN = 0, GSPH_BOUNDARY_PERIODIC = 1, GSPH_BOUNDARY_MIRROR = 2 };
/* gravity_mac = geometric | eigenmac | gadget2 (src/Headers/Tree.h:413-432)                        */
enum { GSPH_MAC_GEOMETRIC = 0, GSPH_MAC_EIGEN = 1, GSPH_MAC_GADGET2 = 2 };
/* multipole = monopole | quadrupole | fast_monopole | fast_quadrupole
 *           (src/GradhSph/GradhSphTree.cpp:549-556,588-593)                                       */
enum { GSPH_MULTIPOLE_MONOPOLE = 0, GSPH_MULTIPOLE_QUADRUPOLE = 1,
       GSPH_MULTIPOLE_FAST_MONOPOLE = 2, GSPH_MULTIPOLE_FAST_QUADRUPOLE = 3 };
/* particle types (src/Headers/Particle.h:44)                                                      */
enum { GSPH_TYPE_GAS = 0, GSPH_TYPE_ICM = 1, GSPH_TYPE_CDM = 2, GSPH_TYPE_DUST = 3,
       GSPH_TYPE_BOUNDARY = 4, GSPH_NTYPES = 5 };
/* flag bits of Particle::flags used by the path (src/Headers/Flags.h:93-136)                       */
enum { GSPH_FLAG_DEAD = 1u << 0, GSPH_FLAG_ACTIVE = 1u << 1, GSPH_FLAG_POTMIN = 1u << 3,
       GSPH_FLAG_END_TIMESTEP = 1u << 4   /* device-resident loop: the particle ends its step now (Particle.h end_timestep) */ };

/* ---- configuration ----------------------------------------------------------------------------
 * Everything the reference passes to GradhSph / GradhSphTree constructors
 * (src/GradhSph/GradhSphSimulation.cpp:131-161,243-246) plus the DomainBox.                        */
typedef struct gsph_config {
  int32_t abi_version;        /* = GSPH_ABI_VERSION                                               */
  int32_t ndim;               /* 1, 2 or 3                                                        */
  int32_t kernel;             /* GSPH_KERNEL_*                                                    */
  int32_t tabulated_kernel;   /* 1: TabulatedKernel semantics (1000 bins, truncating look-up,
                                 SmoothingKernel.h:619-637), 0: analytic polynomials              */
  double  h_fac;              /* h = h_fac (m/rho)^(1/ndim)          (Sph.h:259-263)               */
  double  h_converge;         /* |h - h(rho)|/h tolerance            (GradhSph.cpp:219-220)        */
  int32_t avisc, acond, tdavisc;
  double  alpha_visc, beta_visc, alpha_visc_min;
  int32_t eos;                /* GSPH_EOS_*                                                       */
  double  gamma_eos;          /* ratio of specific heats                                          */
  double  temp0;              /* isothermal/barotropic temperature, CODE units (temp0/temp.outscale) */
  double  mu_bar;             /* mean molecular weight                                            */
  double  rho_bary;           /* barotropic critical density, CODE units                          */
  double  boxmin[3], boxmax[3];
  int32_t boundary_lhs[3], boundary_rhs[3];      /* GSPH_BOUNDARY_*                               */
  int32_t hydro_forces;       /* 1: hydro terms on                                                */
  int32_t self_gravity;       /* 1: UpdateAllSphForces path available                             */
  double  thetamaxsqd;        /* geometric MAC (KDTree.cpp:917)                                   */
  int32_t gravity_mac;        /* GSPH_MAC_*                                                       */
  double  macerror;
  int32_t multipole;          /* GSPH_MULTIPOLE_*                                                 */
  int32_t Nleafmax;           /* max particles per KD-tree leaf (Parameters.cpp:306)              */
  /* ParticleTypeRegister (src/Common/Particle.cpp:31-71): bit j of mask[i] <=> mask[i][j]         */
  uint32_t hmask[GSPH_NTYPES], hydromask[GSPH_NTYPES];
  uint32_t gravmask;          /* bit j: type j contributes to gravity                             */
  uint32_t hydro_forces_type; /* bit i: type i experiences hydro forces                           */
  uint32_t self_gravity_type; /* bit i: type i experiences gravity                                */
  int32_t create_sinks;       /* 1: compute the potmin flag in the density pass (GradhSph.cpp:271) */
  int32_t device;             /* CUDA device ordinal                                              */
  int32_t conservative_sph_star_gravity;  /* 1 (default, Parameters.cpp:253): the density pass adds the
                                 stars of gsph_set_stars to zeta (GradhSph.cpp:288-306)             */
  int32_t nbody_softening;    /* 1 (default, Parameters.cpp:321): mean-h softening in those terms   */
  int32_t reserved[5];
} gsph_config;

/* Fill *cfg with the reference's defaults (Parameters::SetDefaultValues, Parameters.cpp:166-536):
 * ndim 3, m4, tabulated, h_fac 1.2, h_converge 0.01, mon97 alpha 1 beta 2, no conductivity,
 * energy_eqn gamma 5/3, open boundaries, Nleafmax 6, theta^2 0.1, geometric MAC, quadrupole;
 * gas-only type masks.                                                                             */
int gsph_default_config(gsph_config* cfg);

/* ---- particle views ---------------------------------------------------------------------------
 * A view describes where each per-particle field lives in HOST memory as (base pointer, byte
 * stride), so the same call serves the reference's AoS array (stride = sizeof(GradhSphParticle<ndim>)
 * = 448 B in 3-D double, src/Headers/Particle.h:286-368; base = &sphdata[0].field) and plain SoA
 * arrays (stride = 8).  Vector fields are given per component.  A NULL base means "absent":
 * on upload the field takes the reference's constructor default (Particle.h:180-221); on download
 * it is skipped.  The array may be reallocated by the host between calls (Sph::AllocateMemory,
 * GradhSph.cpp:83-111), so views are never cached.                                                 */
typedef struct gsph_field { void* base; size_t stride; } gsph_field;

typedef struct gsph_particle_view {
  gsph_field r[3], v[3], a[3], atree[3];
  gsph_field m, h, u, rho, hrangesqd, hfactor, invomega, zeta, sound, pressure;
  gsph_field dudt, div_v, alpha, dalphadt, gpot, gpot_hydro;
  gsph_field flags;      /* uint32  Particle::flags                                               */
  gsph_field ptype;      /* int32                                                                  */
  gsph_field level;      /* int32                                                                  */
  gsph_field levelneib;  /* int32                                                                  */
  gsph_field iorig;      /* int32                                                                  */
  gsph_field sinkid;     /* int32 (-1: not inside a sink)                                          */
} gsph_particle_view;

/* field_mask bits for gsph_download                                                               */
enum {
  GSPH_F_DENSITY = 1u << 0,   /* h, rho, hrangesqd, hfactor, invomega, zeta, u, sound, pressure,
                                 div_v(=0), flags (potmin)  -- what ComputeH writes                */
  GSPH_F_FORCES  = 1u << 1,   /* a, dudt, div_v, dalphadt, levelneib (+ atree, gpot, gpot_hydro)    */
  GSPH_F_ALL     = 0xffffffffu
};

typedef struct gsph_handle gsph_handle;

/* ---- lifecycle ---------------------------------------------------------------------------------*/
int  gsph_create(const gsph_config* cfg, gsph_handle** out);       /* replaces `new GradhSphTree`,
                                                                      GradhSphSimulation.cpp:243-246 */
void gsph_destroy(gsph_handle* h);
const char* gsph_last_error(const gsph_handle* h);                  /* h may be NULL: create errors  */

/* ---- data movement (host <-> HBM) --------------------------------------------------------------
 * gsph_upload : Nhydro real particles [0,Nhydro) of the host array -> device SoA.
 *               (Boundary ghosts are never materialised on the device: periodic images are taken
 *               on the fly in the walks, as the reference's force pass does,
 *               NeighbourManager.h:417-444 / GhostNeighbours.hpp:146-166.)
 * gsph_download: results back into the host view, in the host's particle order.                    */
int gsph_upload(gsph_handle* h, const gsph_particle_view* view, int64_t Nhydro);
int gsph_download(gsph_handle* h, const gsph_particle_view* view, uint32_t field_mask);
/* Copies run on the library's own copy stream.  gsph_upload returns as soon as the copies are
 * queued (host memory must stay valid until the next synchronising call; pinned memory makes them
 * truly asynchronous): r, h and the integer fields go first and gsph_build_tree starts on them while
 * v, m, u, ... are still in flight.  gsph_download_async queues the device->host copies behind the
 * kernels issued so far and returns; later kernels (e.g. the force pass after a density download)
 * overlap them.  gsph_synchronize (or a plain gsph_download) waits for everything.               */
int gsph_download_async(gsph_handle* h, const gsph_particle_view* view, uint32_t field_mask);
/* Multi-process hosts (one process per GPU, each owning the host-order particles [lo, hi) of the
 * N-particle set -- what an MPI rank owns in the reference, src/Mpi/MpiControl.cpp): ship / fetch only
 * that share over PCIe; the other shares travel GPU-to-GPU (NCCL all-gather issued by the host layer
 * on the arrays of gsph_device_array).  Contiguous (SoA) fields only.                               */
int gsph_upload_range(gsph_handle* h, const gsph_particle_view* view, int64_t N, int64_t lo, int64_t hi);
int gsph_download_range(gsph_handle* h, const gsph_particle_view* view, uint32_t field_mask,
                        int64_t lo, int64_t hi, int wait);
enum { GSPH_ARRAY_TREE_ORDER = 0, GSPH_ARRAY_STAGING = 1 /* host order */ };
int gsph_device_array(gsph_handle* h, int which, const char* name, void** ptr, int64_t* capacity,
                      int32_t* elem_size);

/* ---- the operator surface ----------------------------------------------------------------------
 * gsph_build_tree           NeighbourSearch::BuildTree          HydroTree.cpp:310-372
 *                           (+ BuildGhostTree/SearchBoundaryGhostParticles, :382-441,:495-543,
 *                            which become no-ops: see gsph_upload)
 *     mode 0 = rebuild (KDTree::BuildTree, KDTree.cpp:220-313), 1 = restock (StockTree, :760-798),
 *     2 = extrapolate cell boxes by v*dt (Tree.cpp:172-198), 3 = restock but keep the cells'
 *     gadget2 / eigenmac auxiliaries (amin, macfactor: the reference stocks them from the PREVIOUS
 *     step's atree / gpot in BuildTree, before Sph::ZeroAccelerations; a host that re-uploads after
 *     zeroing uses mode 3 so that they survive).
 * gsph_update_active_counters  KDTree::UpdateActiveParticleCounters   KDTree.cpp:1217-1254
 * gsph_update_properties    SphNeighbourSearch::UpdateAllSphProperties  GradhSphTree.cpp:83-271
 *                           (ComputeH GradhSph.cpp:142-326 + UpdateAllHmaxValues KDTree.cpp:1128-1208)
 * gsph_zero_accelerations   Sph::ZeroAccelerations               Sph.cpp:126-140
 * gsph_update_hydro_forces  ::UpdateAllSphHydroForces            GradhSphTree.cpp:280-435
 * gsph_update_all_forces    ::UpdateAllSphForces                 GradhSphTree.cpp:444-657
 *                           stars = Nstar x {r[3], m, h} doubles (ComputeStarGravForces,
 *                           GradhSph.cpp:699-743); may be NULL/0.
 * gsph_get_hmax             NeighbourSearch::GetMaximumSmoothingLength (SphSimulation.cpp:762-771)  */
int gsph_build_tree(gsph_handle* h, int mode, double dt);
int gsph_update_active_counters(gsph_handle* h);
int gsph_update_properties(gsph_handle* h);
int gsph_zero_accelerations(gsph_handle* h);
int gsph_update_hydro_forces(gsph_handle* h);
int gsph_update_all_forces(gsph_handle* h, const double* stars, int Nstar);
/* The same pass in two halves (stars: gsph_set_stars).  After _begin the neighbour half is done:
 * du/dt, div_v, dalpha/dt and levelneib are final and may be fetched (gsph_download_async) while
 * _finish runs the tree-gravity half and the stars; a, atree and gpot are final after _finish.     */
int gsph_update_all_forces_begin(gsph_handle* h);
int gsph_update_all_forces_finish(gsph_handle* h);
/* Star particles for GradhSph::ComputeStarGravForces (GradhSph.cpp:699-743; called for every active
 * particle by both force drivers when nbody->Nnbody > 0, GradhSphTree.cpp:385-392, :600-607):
 * stars = Nstar x {r[ndim], m, h} doubles.  The list stays on the device and is applied by
 * gsph_update_hydro_forces and gsph_update_all_forces until replaced; Nstar = 0 clears it.         */
int gsph_set_stars(gsph_handle* h, const double* stars, int Nstar);

/* Gravity of the gas on the star particles -- NeighbourSearch::UpdateAllStarGasForces
 * (src/Tree/HydroTree.cpp:552-654: Tree::ComputeStarGravityInteractionList Tree.cpp:748-885, then
 * Nbody::CalculateDirectHydroForces and the cell multipoles at the star's position).  Uses the tree and
 * the tree-order particles of the last passes.  stars = Nstar x {r[ndim], v[ndim], h}; a, adot
 * (Nstar x ndim; adot may be NULL -- only the Hermite N-body schemes use it) and gpot (Nstar) are ADDED
 * to, like the reference adds to star->a, star->adot, star->gpot; counts (may be NULL) = per star the
 * number of smoothed neighbours, direct particles and cells.  Open boundaries only.                  */
int gsph_update_star_gas_forces(gsph_handle* h, const double* stars, int Nstar, double* a, double* adot,
                                double* gpot, int32_t* counts);
int gsph_get_hmax(gsph_handle* h, double* hmax);
int gsph_synchronize(gsph_handle* h);
/* CUDA-event stopwatch on the handle's stream (torch.cuda.Event only sees torch's streams):
 * start records an event, stop records a second one, synchronises and returns the elapsed ms.     */
int gsph_timer_start(gsph_handle* h);
int gsph_timer_stop(gsph_handle* h, float* ms);

/* ---- multi-GPU plumbing (one process per GPU) ----------------------------------------
 * The reference shards the path by spatial domain decomposition over MPI ranks
 * (src/Mpi/MpiKDTreeDecomposition.cpp:56-277) and exchanges ghost particles as whole 448-byte
 * structs (src/Mpi/MpiControl.cpp:1002-1157).  Here every rank holds the tree and works on ONE
 * contiguous slice of the tree order -- which IS the reference's KD domain decomposition, because
 * the top log2(R) levels of the balanced KD-tree are exactly its median splits -- and the one
 * exchange step per pass (post-density fields of the other slices) is an NCCL all-gather issued by
 * the host layer on the device arrays exposed below.
 *   gsph_set_active_slice : particles at tree positions [lo, hi) become the targets of this device, all
 *                           others stay sources only (replaces the rank-ownership test).  No flag is
 *                           rewritten: the kernels test the range.  Result arrays are written inside
 *                           [lo, hi) only; outside they hold what the last exchange (or allocation: zeros)
 *                           left there.  The force records of the force pass are packed only for the cohort
 *                           nodes that can hold a neighbour of the slice.  Cleared by the next gsph_upload.
 *   gsph_derive_density_fields : after the exchange of h, rho, invomega, zeta: hrangesqd, hfactor, u,
 *                           sound, pressure of the particles OUTSIDE [lo, hi) are recomputed from them
 *                           exactly as the density pass stores them (4 instead of 9 arrays cross NVLink).
 *   gsph_density_pass     : gsph_update_properties without the tree refresh (h of the other
 *                           slices is not known yet);  gsph_refresh_hmax does that refresh
 *                           (KDTree::UpdateAllHmaxValues) after the exchange.
 *   gsph_device_field     : device pointer of one tree-order SoA array (field = GSPH_DF_*),
 *                           valid until the next gsph_upload that grows the arrays.              */
enum { GSPH_DF_H = 0, GSPH_DF_RHO, GSPH_DF_HRANGESQD, GSPH_DF_HFACTOR, GSPH_DF_INVOMEGA, GSPH_DF_ZETA,
       GSPH_DF_SOUND, GSPH_DF_PRESSURE, GSPH_DF_U,
       GSPH_DF_PERM,   /* int32 tree position -> host index (the tree order itself) */
       GSPH_DF_COUNT };
/* Sharded tree build.  The balanced KD-tree's top log2(world) levels ARE the reference's domain
 * decomposition (median splits, MpiKDTreeDecomposition.cpp:56-277): rank r owns the r-th node of that
 * level = tree positions [lo, hi) of gsph_tree_shard_range.
 *   gsph_build_tree_shard_begin  : finds this rank's shard by selecting the medians of the top levels
 *                                  (radix select over (coordinate, id); no sort of all particles), then
 *                                  sorts and builds only this rank's subtree; leaves this rank's slice
 *                                  of GSPH_DF_PERM filled (1-D: sorts everything, splits its part);
 *   (host layer)                 : all-gathers the GSPH_DF_PERM slices over NCCL;
 *   gsph_build_tree_shard_finish : permutes the particle fields into tree order and stocks the cells
 *                                  (KDTree::StockTree) -- the tree is then identical on every rank
 *                                  and identical to the one gsph_build_tree(mode 0) builds.         */
int gsph_tree_shard_range(gsph_handle* h, int rank, int world, int64_t* lo, int64_t* hi);
int gsph_build_tree_shard_begin(gsph_handle* h, int rank, int world);
int gsph_build_tree_shard_finish(gsph_handle* h);
int gsph_set_active_slice(gsph_handle* h, int64_t lo, int64_t hi);
int gsph_density_pass(gsph_handle* h);
int gsph_refresh_hmax(gsph_handle* h);
int gsph_derive_density_fields(gsph_handle* h, int64_t lo, int64_t hi);
int gsph_device_field(gsph_handle* h, int field, void** ptr, int64_t* capacity);

/* ---- parity / debug access ---------------------------------------------------------------------
 * Neighbour lists as defined by NeighbourManager::TrimNeighbourLists (NeighbourManager.h:521):
 *   j is a neighbour of i  <=>  NOT( drsqd >= hrangesqd_i && drsqd >= hrangesqd_j )  and
 *   hydromask[ptype_i][ptype_j]; self included.  Indices are HOST particle indices (upload order).
 * Call gsph_neighbour_counts first (counts[Nhydro]), build offsets, then gsph_neighbour_lists
 * (idx sorted ascending within each row).                                                          */
int gsph_neighbour_counts(gsph_handle* h, int32_t* counts);
int gsph_neighbour_lists(gsph_handle* h, const int64_t* offsets, int32_t* idx);

/* KD-tree inspection.  Cells are reported with the reference's depth-first ids
 * (KDTree::CreateTreeStructure, KDTree.cpp:362-433).                                              */
typedef struct gsph_tree_info {
  int32_t ltot;        /* leaf level                                                               */
  int32_t Ncell;       /* 2^(ltot+1) - 1                                                           */
  int32_t gtot;        /* number of leaf cells 2^ltot                                              */
  int32_t Nleafmax;
  int64_t Ntot;
} gsph_tree_info;
int gsph_tree_info_get(gsph_handle* h, gsph_tree_info* info);
/* leaf_of[i] = depth-first id of the leaf cell holding host particle i                            */
int gsph_tree_leaf_of(gsph_handle* h, int32_t* leaf_of);
/* per cell (depth-first id c): box[c][4*ndim] = bb.min, bb.max, hbox.min, hbox.max;
 * com[c][2*ndim] = rcell, r; scalars[c][5] = m, rmax, hmax, cdistsqd, mac; q[c][5]; N[c].
 * Any pointer may be NULL.                                                                         */
int gsph_tree_cells(gsph_handle* h, double* box, double* com, double* scalars, double* q, int32_t* N);

/* run statistics of the most recent calls (SURVEY 8d "value distributions")                        */
typedef struct gsph_stats {
  int64_t Nhydro;
  int64_t n_active;            /* particles updated by the last pass                               */
  double  ngb_mean; int32_t ngb_max;      /* hydro neighbours per particle, last force pass        */
  double  cand_mean;                      /* candidates tested per target, last force pass         */
  double  hiter_mean; int32_t hiter_max;  /* h-rho iterations, last density pass                   */
  int32_t density_retries;                /* cohorts that needed a larger search radius            */
  double  ndirect_mean, ncell_mean;       /* gravity: direct-sum particles / cells per particle    */
  int64_t kernel_launches;     /* CUDA kernels launched by this handle since create               */
  float   ms_build_tree, ms_properties, ms_forces;   /* device time of the last call of each       */
  float   ms_upload, ms_download;
  /* device time of the most recent launch of each hot kernel (CUDA events on the handle's stream),
     the denominators of bench.py's roofline line                                                  */
  float   ms_kernel_density, ms_kernel_hydro, ms_kernel_gravity;
  float   ms_tree_sort, ms_tree_levels, ms_tree_stock;   /* parts of the last rebuild              */
} gsph_stats;
int gsph_get_stats(gsph_handle* h, gsph_stats* st);

/*--------------------------------------------------------------------------------------------------
 * Device-resident main loop: leapfrog kick-drift-kick with one global time step
 * (Nlevels = 1) or hierarchical block time steps, so that particles stay in HBM between steps instead
 * of travelling over PCIe around every operator call.  Replaces, for sph_integration = lfkdk,
 *   SphSimulation::MainLoop                      src/Hydrodynamics/SphSimulation.cpp:574-880
 *   SphLeapfrogKDK::AdvanceParticles/EndTimestep src/Hydrodynamics/SphLeapfrogKDK.cpp:76-125,226-270
 *   SphIntegration::Timestep                     src/Hydrodynamics/SphIntegration.cpp:76-124
 *   Simulation::ComputeGlobalTimestep            src/Common/Simulation.cpp:1669-1745
 *   TimeIntegration::CheckBoundaries             src/Common/Integration.cpp:11-75
 * around the operators above (BuildTree, UpdateAllSphProperties, ZeroAccelerations, force pass).
 * Block time steps (gsph_integrator_config::Nlevels > 1; several devices through gsph_integrator_advance /
 * _check_timesteps / _finish): additionally
 *   Simulation::ComputeBlockTimesteps            src/Common/Simulation.cpp:1764-2199
 *   SphLeapfrogKDK::CheckTimesteps               src/Hydrodynamics/SphLeapfrogKDK.cpp:284-339 (neighbour-level limiter; the
 *                                                passes are repeated for the particles it wakes, SphSimulation.cpp:658-756)
 *   the thermal refresh of every particle before the force pass (nradstep = 1, SphSimulation.cpp:669-679).
 * gsph_integrator_begin then sets up the hierarchy like PostInitialConditionsSetup (a resynchronisation at n = 0).
 * Star particles: gsph_integrator_set_stars / _set_stars_hermite4 below, with one global time step or on block time
 * steps of their own.  Not covered (GSPH_EINVAL): sinks, external potentials, dust.                                                                              */
typedef struct gsph_integrator_config {
  double  accel_mult, courant_mult, energy_mult;   /* Parameters.cpp:239,240,263 (0.3, 0.15, 0.4)      */
  int32_t energy_integration;   /* 1: u is integrated explicitly (gas_eos = energy_eqn,
                                   GradhSphSimulation.cpp:110-123); 0: u follows from the EOS         */
  int32_t ntreebuildstep;       /* full rebuild every this many steps (HydroTree.cpp BuildTree)       */
  int32_t ntreestockstep;       /* otherwise restock every this many steps, else extrapolate          */
  int32_t sink_particles;       /* Parameters.cpp "sink_particles": must be 0 -- MainLoop's sink search, accretion
                                   and stellar-property update (SphSimulation.cpp:815-833) are not on this path   */
  int32_t Nlevels;              /* 0 or 1: one global time step; > 1: hierarchical block time steps
                                   (Simulation::ComputeBlockTimesteps, Simulation.cpp:1764-2199)                   */
  int32_t level_diff_max;       /* largest level difference between neighbours (Parameters.cpp:245, 1);
                                   0 is read as the default 1                                                     */
  int32_t sph_single_timestep;  /* Parameters.cpp:246: with Nlevels > 1 every gas particle runs on the deepest level the gas
                                   needs (star particles may sit below it); Simulation.cpp:1889-1902, 2092-2099            */
  int32_t reserved[1];
} gsph_integrator_config;

/* Start from the particles of the last gsph_upload (host order is kept).
 *   compute_forces = 1 : first run BuildTree, UpdateAllSphProperties, ZeroAccelerations and the
 *                        force pass once (the tail of PostInitialConditionsSetup,
 *                        SphSimulation.cpp:390-465);
 *   compute_forces = 0 : take the uploaded h, sound, div_v, a, du/dt, alpha ... as they are
 *                        (restart from a complete state, e.g. the reference's own setup);
 *   compute_forces = 2 : the caller has just run those passes itself (tree-order results present).
 * Then r0, v0, a0, u0, dudt0 are set and the first global time step is computed
 * (SphSimulation.cpp:480-486, 544-563).                                                            */
int gsph_integrator_begin(gsph_handle* h, const gsph_integrator_config* cfg, int compute_forces, double t0);
/* nsteps x MainLoop.  Afterwards gsph_download(..., GSPH_F_ALL) returns the particle state as the
 * reference holds it at the end of MainLoop (corrected v, u; r0 = r ...).                          */
int gsph_integrator_step(gsph_handle* h, int nsteps);
/* One MainLoop in two halves, for hosts that run the operators in between themselves (several
 * ranks: sharded tree build, slice passes and the NCCL exchanges, see gandalf_b200/domain.py):
 *   advance : t += timestep, prediction, boundaries; the host-order arrays are a fresh "upload";
 *             *tree_mode = 0 rebuild / 1 restock / 2 extrapolate for BuildTree, *timestep = the step;
 *   finish  : after BuildTree, UpdateAllSphProperties, ZeroAccelerations and the force pass, with the
 *             tree-order result arrays complete on this device: results back to host order, next
 *             global time step, end-of-step terms.
 * gsph_integrator_begin(compute_forces = 2) is the matching start: the caller has just run the passes.
 * gsph_integrator_step(h, 1) == advance + gsph_build_tree + the passes + finish.                     */
/* Star particles in the device-resident loop.  Their N-body side is the reference's leapfrog
 * (NbodyLeapfrogKDK.cpp: AdvanceParticles, CalculateDirectSmoothedGravForces, CorrectionTerms, Timestep,
 * EndTimestep; nbody = lfkdk, nbody_softening = 1, Npec = 1, open boundaries), run on the host -- the stars
 * are a handful -- around gsph_update_star_gas_forces; the gas feels them through the star terms of the
 * density and force passes, and their time step enters the global minimum.
 *   set_stars : rows = Nstar x {r[ndim], v[ndim], a[ndim], m, h}, before gsph_integrator_begin (a is read
 *               only with compute_forces = 0); nbody_mult as Parameters.cpp:241 (0.1)
 *   get_stars : rows = Nstar x {r[ndim], v[ndim], a[ndim], m, h, gpot} after the last step          */
int gsph_integrator_set_stars(gsph_handle* h, const double* rows, int Nstar, double nbody_mult);
int gsph_integrator_get_stars(gsph_handle* h, double* rows, int Nstar);
/* nbody = hermite4, the reference's DEFAULT N-body integrator (NbodyHermite4.cpp: AdvanceParticles third-order
 * predictor, CalculateDirectSmoothedGravForces with the jerk, CalculateAllStartupQuantities, CorrectionTerms Hermite
 * corrector, the Aarseth Timestep, EndTimestep), again on the host around gsph_update_star_gas_forces (which returns
 * the jerk of the gas on the stars).  rows = Nstar x {r, v, a, adot, a2dot, a3dot (ndim each), m, h}; a .. a3dot are
 * read only by gsph_integrator_begin(compute_forces = 0).  get_stars returns r, v, a, m, h, gpot as above,
 * get_star_derivatives adot, a2dot, a3dot (Nstar x ndim each, NULL = not wanted).                               */
int gsph_integrator_set_stars_hermite4(gsph_handle* h, const double* rows, int Nstar, double nbody_mult);
int gsph_integrator_get_star_derivatives(gsph_handle* h, double* adot, double* a2dot, double* a3dot);
/* with block time steps (Nlevels > 1) the stars have levels of their own, at least the deepest one the gas occupies
 * (Simulation::ComputeBlockTimesteps, Simulation.cpp:1858-1872, 2024-2068, 2146-2152): only the stars at the end of
 * their step get new forces, the correction and EndTimestep.  level, nstep, nlast: int64[Nstar], tlast: double[Nstar]. */
int gsph_integrator_get_star_levels(gsph_handle* h, int64_t* level, int64_t* nstep, int64_t* nlast, double* tlast);
int gsph_integrator_advance(gsph_handle* h, int* tree_mode, double* timestep);
int gsph_integrator_finish(gsph_handle* h);
/* Block time steps (Nlevels > 1) in a host-driven loop (several devices): between the passes and
 * gsph_integrator_finish the host layer calls gsph_integrator_check_timesteps (SphLeapfrogKDK::CheckTimesteps,
 * SphLeapfrogKDK.cpp:283-348): *woken = particles whose step the neighbour limiter cuts short.  While it is > 0
 * their active flags and levels are in place and the passes run again (the do-while loop of MainLoop,
 * SphSimulation.cpp:658-756); gsph_integrator_finish refuses to run before it has reported 0.
 * gsph_update_thermal_properties = GradhSph::ComputeThermalProperties of EVERY particle (active or not), which
 * MainLoop does between the density and the force pass with nradstep = 1 (SphSimulation.cpp:669-679); only the
 * passes of a block-time-step run need it (with one global step the density pass has just done the same).
 * With an active slice and inactive particles the levelneib scatter of the force pass reaches particles of
 * every slice: combine the devices' levelneib arrays with a MAX reduction, not with a gather of slices.      */
int gsph_integrator_check_timesteps(gsph_handle* h, int* woken);
int gsph_update_thermal_properties(gsph_handle* h);
/* simulation time, the time step the NEXT MainLoop will take, steps done since begin             */
int gsph_integrator_state(gsph_handle* h, double* t, double* timestep, int64_t* nsteps);
/* block time steps: the scalars of the hierarchy (Simulation::n, nresync, level_max, level_step, dt_max) and, where
 * the pointers are not NULL, the integrator state of every particle in host order: nstep, nlast (int32[N]), dt, tlast
 * (double[N]).  level and levelneib come with gsph_download.                                          */
typedef struct gsph_block_state { int64_t n, nresync; int32_t level_max, level_step; double dt_max; } gsph_block_state;
int gsph_integrator_block_state(gsph_handle* h, gsph_block_state* st, int32_t* nstep, int32_t* nlast, double* dt, double* tlast);
/* the start-of-step state SphLeapfrogKDK keeps in every particle (r0, v0, a0, u0, dudt0, Particle.h:150-165; with
 * block time steps also dt_next), host order; r0, v0, a0 = ndim arrays of N doubles one after the other.  Together with
 * gsph_download and gsph_integrator_block_state this is everything the reference's own MainLoop needs to carry a
 * device-resident run on (GpuSphTree::ResidentFetch).  NULL = not wanted.                                       */
int gsph_integrator_start_state(gsph_handle* h, double* r0, double* v0, double* a0, double* u0, double* dudt0, double* dt_next);

/* Kernel tables exactly as TabulatedKernel builds them (TabulatedKernel.cpp:58-105,
 * SmoothingKernel.h:585-605): which = 0 w0_s2, 1 womega_s2, 2 wzeta_s2, 3 w1, 4 wgrav, 5 wpot;
 * out[1000].  Host-side, no device needed (parity stage P0).                                      */
int gsph_kernel_table(int ndim, int kernel, int which, double* out);

/* Device peak probe used by bench.py for the FP64 roofline ("measure FP64 and FP32 FMA
 * peak with a register-resident micro-kernel").  Returns TFLOP/s (2 flops per FMA).               */
int gsph_measure_fp_peak(int device, double* fp64_tflops, double* fp32_tflops);

#ifdef __cplusplus
}
#endif
#endif /* GSPH_H */
