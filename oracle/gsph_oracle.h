/*=================================================================================================
 *  oracle/gsph_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 *  Plain-C CPU restatement of the reference algorithm for the grad-h SPH neighbour path
 *  (GANDALF v0.4.0).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference
 *  legs may load this; the product library (gandalf_b200/libgsph.so) never does.
 *
 *  PARITY PINNED: checked against golden vectors dumped from the unmodified reference built from
 *  /root/reference (oracle/ref_harness.cpp -> tests/golden/*.npz, tests/test_oracle_golden.py).
 *
 *  Arrays: vectors are [N][ndim] row-major doubles; all particles are treated as active.
 *=================================================================================================*/
#ifndef GSPH_ORACLE_H
#define GSPH_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_params {
  int ndim;
  int kernel;            /* 0 m4, 1 quintic                                                       */
  int tabulated;         /* 1: 1000-bin truncating tables (SmoothingKernel.h:585-637)              */
  double h_fac, h_converge;
  int avisc, acond, tdavisc;        /* Sph.h:50-52 enum order                                      */
  double alpha_visc, beta_visc, alpha_visc_min;
  int eos;               /* 0 adiabatic (energy_eqn), 1 isothermal, 2 barotropic                    */
  double gamma_eos, temp0, mu_bar, rho_bary;
  double boxmin[3], boxmax[3];
  int periodic[3];       /* 1: periodic in that dimension                                          */
  int self_gravity;
  double thetamaxsqd;
  int gravity_mac;       /* 0 geometric, 1 eigenmac, 2 gadget2 (Tree.h:413-432)                    */
  double macerror;
  int multipole;         /* 0 monopole 1 quadrupole 2 fast_monopole 3 fast_quadrupole              */
  int Nleafmax;
  int hydro_forces;
  int ntarget;           /* > 0: only the first ntarget particles are targets; the others are sources only
                            (materialised mirror-boundary images, Hydrodynamics.cpp:217-290)            */
  /* stars for the zeta terms of the density pass (GradhSph.cpp:288-306): nstar rows {r[ndim], m, h};
     star_zeta = conservative_sph_star_gravity, nbody_softening as in Parameters.cpp:321             */
  int star_zeta, nbody_softening, nstar;
  const double* stars;
} orc_params;

/* kernel functions, which: 0 w0, 1 w1, 2 womega, 3 wzeta, 4 wgrav, 5 wpot; s = r/h              */
double orc_kernel(const orc_params* p, int which, double s);
/* s2-argument variants which: 0 w0_s2, 1 womega_s2, 2 wzeta_s2                                    */
double orc_kernel_s2(const orc_params* p, int which, double s2);
/* tables as TabulatedKernel builds them; which 0 w0_s2 1 womega_s2 2 wzeta_s2 3 w1 4 wgrav 5 wpot  */
void orc_kernel_table(const orc_params* p, int which, double* out1000);

/* neighbour lists by the force-pass predicate (NeighbourManager.h:521), self included,
 * ascending index.  Pass idx = NULL to get counts only.  Returns total number of entries.         */
int64_t orc_neighbour_lists(const orc_params* p, int N, const double* r, const double* hrangesqd,
                            int64_t* offsets /*[N+1]*/, int32_t* idx);

/* h-rho iteration + thermal properties for every particle (GradhSph::ComputeH,
 * GradhSph.cpp:142-326, driven as in GradhSphTree.cpp:144-240).
 * in : r, m, h (guess), u ;  out: h, rho, invomega, zeta, hfactor, hrangesqd, u, sound, pressure,
 * niter (may be NULL).  Returns 0, or -1 if an iteration failed to converge.                      */
int orc_density(const orc_params* p, int N, const double* r, const double* m, double* h,
                double* u, double* rho, double* invomega, double* zeta, double* hfactor,
                double* hrangesqd, double* sound, double* pressure, int32_t* niter);

/* hydro-only force pass (GradhSph::ComputeSphHydroForces, GradhSph.cpp:361-460, driven as in
 * GradhSphTree.cpp:327-406).  Outputs are ADDED to a, dudt, div_v, dalphadt like the reference
 * does; levelneib is max-combined.                                                                 */
int orc_hydro_forces(const orc_params* p, int N, const double* r, const double* v, const double* m,
                     const double* h, const double* rho, const double* pressure,
                     const double* invomega, const double* sound, const double* u,
                     const double* alpha, const double* hfactor, const double* hrangesqd,
                     const int32_t* level,
                     double* a, double* dudt, double* div_v, double* dalphadt, int32_t* levelneib);

/* Sph<ndim>::ComputeCullenAndDehnenViscosity (src/Headers/Sph.h:364-456), called at the end of ComputeH
 * when time_dependent_avisc = cd2010 (GradhSph.cpp:319-321): for every target, from the gather
 * neighbours' v and a (the previous step's accelerations) and the FINAL h, rho, hfactor, sound of the
 * density pass; alpha is updated in place (max with the local value), dalphadt is written.          */
int orc_cd2010_switch(const orc_params* p, int N, const double* r, const double* v, const double* a,
                      const double* m, const double* h, const double* rho, const double* hfactor,
                      const double* sound, double* alpha, double* dalphadt);

/* GradhSph::ComputeStarGravForces, src/GradhSph/GradhSph.cpp:699-743, for every particle: softened
 * gravity of Nstar stars {r[ndim], m, h} (stride ndim+2) with the mean-h kernel; adds to atree and,
 * like both force drivers do (GradhSphTree.cpp:396-399, :611-614), to a, and to gpot.
 * (pinned by tests/golden/sphere_grav_stars.npz)                                                    */
int orc_star_grav_forces(const orc_params* p, int N, const double* r, const double* h, int Nstar,
                         const double* stars, double* a, double* atree, double* gpot);

/* hydro+gravity with EXACT direct summation for everything outside the kernel overlap
 * (what the reference computes with neib_search = bruteforce: every other particle is either a
 * hydro neighbour -> ComputeSphHydroGravForces GradhSph.cpp:474-585, or an unsoftened direct
 * neighbour -> ComputeDirectGravForces :657-690).  a gets hydro + gravity, atree gravity only,
 * gpot includes the self term m/h wpot(0) (GradhSphTree.cpp:512).                                  */
int orc_hydro_grav_forces_direct(const orc_params* p, int N, const double* r, const double* v,
                     const double* m, const double* h, const double* rho, const double* pressure,
                     const double* invomega, const double* zeta, const double* sound,
                     const double* u, const double* alpha, const double* hfactor,
                     const double* hrangesqd, const int32_t* level,
                     double* a, double* atree, double* gpot, double* dudt, double* div_v,
                     double* dalphadt, int32_t* levelneib);

/* ---- KD-tree restatement (KDTree.cpp:220-1083) ---------------------------------------------------
 * Builds the reference's balanced KD-tree (same QuickSelect, same DFS cell numbering) and stocks
 * the cells.  Cell arrays are indexed by the reference's depth-first id.                           */
typedef struct orc_tree orc_tree;
orc_tree* orc_tree_build(const orc_params* p, int N, const double* r, const double* v,
                         const double* m, const double* h);
/* same with the previous step's atree [N][ndim] and gpot [N], which the gadget2 / eigenmac opening
 * criteria read when the cells are stocked (KDTree.cpp:899-903); either may be NULL */
orc_tree* orc_tree_build_mac(const orc_params* p, int N, const double* r, const double* v,
                             const double* m, const double* h, const double* atree, const double* gpot);
/* KDTree::UpdateHmaxValues (KDTree.cpp:1128-1208): hmax / hbox of the cells from new smoothing lengths h (the array
 * must stay alive); cdistsqd keeps the value of the last stock, as in the reference */
void orc_tree_update_hmax(orc_tree* t, const double* h);
void orc_tree_free(orc_tree* t);
int  orc_tree_ncell(const orc_tree* t);
int  orc_tree_ltot(const orc_tree* t);
/* leaf_of[i] = DFS id of the leaf containing particle i                                            */
void orc_tree_leaf_of(const orc_tree* t, int32_t* leaf_of);
/* same layout as gsph_tree_cells                                                                   */
void orc_tree_cells(const orc_tree* t, double* box, double* com, double* scalars, double* q, int32_t* N);

/* hydro+gravity through the tree walk (Tree::ComputeGravityInteractionAndGhostList Tree.cpp:628-735,
 * NeighbourManager::_EndSearch/TrimNeighbourLists, ComputeCell{Monopole,Quadrupole}Forces
 * NeighbourSearch.h:349-475, fast multipoles :481-827), open boundaries only.
 * Optionally returns per-particle counts of direct-sum particles and cells (may be NULL).          */
int orc_hydro_grav_forces_tree(const orc_params* p, const orc_tree* t, int N, const double* r,
                     const double* v, const double* m, const double* h, const double* rho,
                     const double* pressure, const double* invomega, const double* zeta,
                     const double* sound, const double* u, const double* alpha,
                     const double* hfactor, const double* hrangesqd, const int32_t* level,
                     double* a, double* atree, double* gpot, double* dudt, double* div_v,
                     double* dalphadt, int32_t* levelneib, int32_t* ndirect, int32_t* ncell);

/* gravity of the gas on star particles (HydroTree::UpdateAllStarGasForces, HydroTree.cpp:552-654):
 * stars = Nstar rows {r[ndim], v[ndim], h}; a, adot (Hermite schemes), gpot are ADDED to;
 * counts (may be NULL) = per star the sizes of its neighbour, direct and cell lists.              */
int orc_star_gas_forces(const orc_params* p, const orc_tree* T, int Nstar, const double* stars,
                        double* a, double* adot, double* gpot, int32_t* counts);

#ifdef __cplusplus
}
#endif
#endif
