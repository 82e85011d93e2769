//=================================================================================================
//  GpuSphTree.h -- host-side C++ twin of the reference's neighbour-search object:
//  `neib_search = gpu`.
//
//  GpuSphTree<ndim> IS-A GradhSphTree<ndim,GradhSphParticle> (src/Headers/SphNeighbourSearch.h:
//  174-227) and overrides exactly the virtuals of the hot path, forwarding each one to the C ABI of
//  libgsph (include/gsph.h):
//
//      BuildTree                 HydroTree.cpp:310-372      -> gsph_upload + gsph_build_tree
//      UpdateAllSphProperties    GradhSphTree.cpp:83-271    -> gsph_update_properties
//      UpdateAllSphHydroForces   GradhSphTree.cpp:280-435   -> gsph_update_hydro_forces
//      UpdateAllSphForces        GradhSphTree.cpp:444-657   -> gsph_update_all_forces
//      GetMaximumSmoothingLength                            -> gsph_get_hmax
//
//  Everything that is NOT on the accelerated path keeps the inherited CPU implementation
//  (GetGatherNeighbourList for sinks/feedback, UpdateAllStarGasForces, ghost creation for the
//  integrator, GetTree for dust ...).  This header is compiled INTO the reference (it needs the
//  reference headers); it contains no reference code.  The particle array is owned by
//  GradhSph::sphdata and may be reallocated between calls (GradhSph.cpp:83-111), so the view is
//  rebuilt from hydro->GetParticleArray() on every call.
//
//  Error convention: a non-zero C-ABI status is handed to ExceptionHandler::raise
//  (src/Common/Exception.cpp:47-69) with gsph_last_error() as the message.
//=================================================================================================
#ifndef _GPU_SPH_TREE_H_
#define _GPU_SPH_TREE_H_

#include <string>
#include <cstring>
#include <vector>
#include "gsph.h"
#include "Exception.h"
#include "Parameters.h"
#include "SimUnits.h"
#include "Sph.h"
#include "Particle.h"
#include "SphNeighbourSearch.h"
#include "CodeTiming.h"
#include "Nbody.h"

template <int ndim>
class GpuSphTree : public GradhSphTree<ndim, GradhSphParticle>
{
  typedef GradhSphTree<ndim, GradhSphParticle> Base;
  typedef GradhSphParticle<ndim> Part;

  gsph_handle* handle;
  bool keep_host_tree;        ///< also maintain the CPU tree for the non-accelerated consumers
  bool staged;                ///< device copy matches the host array of this step
  int resident_levels = 1;    ///< Nlevels of the device-resident loop started by ResidentBegin
  bool resident_energy = true; ///< ... and whether it integrates u (SphLeapfrogKDK touches u0, dudt0 only then)
  int resident_stars = 0;     ///< star particles handed to the device loop
  bool resident_h4 = false;   ///< ... integrated with NbodyHermite4

  static int lookup(const std::string& v, const char* const* names, int n, const char* what) {
    for (int i = 0; i < n; i++) if (v == names[i]) return i;
    ExceptionHandler::getIstance().raise(std::string("GpuSphTree: unsupported ") + what + " = " + v);
    return 0;
  }

  void check(int rc) {
    if (rc != GSPH_OK) ExceptionHandler::getIstance().raise(std::string("libgsph: ") + gsph_last_error(handle));
  }

  static gsph_field field(void* base) { gsph_field f; f.base = base; f.stride = sizeof(Part); return f; }

  /// (base pointer, stride) of every field of the reference's AoS particle array
  static gsph_particle_view make_view(Part* p) {
    gsph_particle_view v;
    memset(&v, 0, sizeof v);
    for (int k = 0; k < ndim; k++) {
      v.r[k] = field(&p[0].r[k]); v.v[k] = field(&p[0].v[k]);
      v.a[k] = field(&p[0].a[k]); v.atree[k] = field(&p[0].atree[k]);
    }
    v.m = field(&p[0].m); v.h = field(&p[0].h); v.u = field(&p[0].u); v.rho = field(&p[0].rho);
    v.hrangesqd = field(&p[0].hrangesqd); v.hfactor = field(&p[0].hfactor);
    v.invomega = field(&p[0].invomega); v.zeta = field(&p[0].zeta); v.sound = field(&p[0].sound);
    v.pressure = field(&p[0].pressure); v.dudt = field(&p[0].dudt); v.div_v = field(&p[0].div_v);
    v.alpha = field(&p[0].alpha); v.dalphadt = field(&p[0].dalphadt); v.gpot = field(&p[0].gpot);
    v.gpot_hydro = field(&p[0].gpot_hydro);
    v.flags = field(&p[0].flags);            // type_flag holds exactly one unsigned int (Flags.h:93-96)
    v.ptype = field(&p[0].ptype); v.level = field(&p[0].level); v.levelneib = field(&p[0].levelneib);
    v.iorig = field(&p[0].iorig); v.sinkid = field(&p[0].sinkid);
    return v;
  }

  void upload(Hydrodynamics<ndim>* hydro) {
    Part* p = hydro->template GetParticleArray<GradhSphParticle>();
    gsph_particle_view v = make_view(p);
    check(gsph_upload(handle, &v, hydro->Nhydro));
  }
  void download(Hydrodynamics<ndim>* hydro, unsigned mask) {
    Part* p = hydro->template GetParticleArray<GradhSphParticle>();
    gsph_particle_view v = make_view(p);
    check(gsph_download(handle, &v, mask));
  }

 public:

  /// Same argument list as GradhSphTree (GradhSphSimulation.cpp:243-246) plus the objects the
  /// configuration is read from.
  GpuSphTree(int Nleafmax, int Nmpi, int pruning_level_min, int pruning_level_max, FLOAT thetamaxsqd,
             FLOAT kernrange, FLOAT macerror, std::string gravity_mac, std::string multipole,
             DomainBox<ndim>* box, SmoothingKernel<ndim>* kern, CodeTiming* timing,
             ParticleTypeRegister& types, Sph<ndim>* sph, Parameters* params, SimUnits& units,
             bool keep_host_tree_aux = true)
    : Base("kdtree", Nleafmax, Nmpi, pruning_level_min, pruning_level_max, thetamaxsqd, kernrange,
           macerror, gravity_mac, multipole, box, kern, timing, types),
      handle(NULL), keep_host_tree(keep_host_tree_aux), staged(false)
  {
    std::map<std::string, int>& ip = params->intparams;
    std::map<std::string, double>& fp = params->floatparams;
    std::map<std::string, std::string>& sp = params->stringparams;
    gsph_config c;
    gsph_default_config(&c);
    c.ndim = ndim;
    static const char* const kernels[] = {"m4", "quintic", "gaussian"};
    c.kernel = lookup(sp["kernel"], kernels, 3, "kernel");
    c.tabulated_kernel = ip["tabulated_kernel"];
    c.h_fac = sph->h_fac; c.h_converge = sph->h_converge;
    c.avisc = (int) sph->avisc; c.acond = (int) sph->acond;           // same enum order, Sph.h:50-52
    c.tdavisc = (int) sph->tdavisc;
    c.alpha_visc = sph->alpha_visc; c.beta_visc = sph->beta_visc; c.alpha_visc_min = sph->alpha_visc_min;
    static const char* const eos[] = {"energy_eqn", "isothermal", "barotropic", "barotropic2"};
    c.eos = sp["gas_eos"] == "constant_temp" ? GSPH_EOS_ADIABATIC : lookup(sp["gas_eos"], eos, 4, "gas_eos");
    c.gamma_eos = fp["gamma_eos"];
    c.temp0 = fp["temp0"]/units.temp.outscale;                       // code units, as the EOS classes scale them
    c.mu_bar = fp["mu_bar"];
    c.rho_bary = fp["rho_bary"]/units.rho.outscale/units.rho.outcgs;
    static const char* const bnd[] = {"open", "periodic", "mirror"};
    for (int k = 0; k < ndim; k++) {
      c.boxmin[k] = box->min[k]; c.boxmax[k] = box->max[k];
      c.boundary_lhs[k] = box->boundary_lhs[k] == periodicBoundary ? 1 : (box->boundary_lhs[k] == mirrorBoundary ? 2 : 0);
      c.boundary_rhs[k] = box->boundary_rhs[k] == periodicBoundary ? 1 : (box->boundary_rhs[k] == mirrorBoundary ? 2 : 0);
    }
    (void) bnd;
    c.hydro_forces = sph->hydro_forces; c.self_gravity = sph->self_gravity;
    c.thetamaxsqd = thetamaxsqd; c.macerror = macerror;
    static const char* const macs[] = {"geometric", "eigenmac", "gadget2"};
    c.gravity_mac = lookup(gravity_mac, macs, 3, "gravity_mac");
    static const char* const mps[] = {"monopole", "quadrupole", "fast_monopole", "fast_quadrupole"};
    c.multipole = lookup(multipole, mps, 4, "multipole");
    c.Nleafmax = Nleafmax;
    c.create_sinks = sph->create_sinks;
    c.conservative_sph_star_gravity = sph->conservative_sph_star_gravity;
    c.nbody_softening = params->intparams["nbody_softening"];
    int rc = gsph_create(&c, &handle);
    if (rc != GSPH_OK) ExceptionHandler::getIstance().raise(std::string("libgsph: ") + gsph_last_error(NULL));
  }

  virtual ~GpuSphTree() { gsph_destroy(handle); }

  //-----------------------------------------------------------------------------------------------
  virtual void BuildTree(const bool rebuild_tree, const int n, const int ntreebuildstep,
                         const int ntreestockstep, const FLOAT timestep, Hydrodynamics<ndim>* hydro)
  {
    if (keep_host_tree) Base::BuildTree(rebuild_tree, n, ntreebuildstep, ntreestockstep, timestep, hydro);
    CodeTiming::BlockTimer timer = this->timing->StartNewTimer("GPU_BUILD_TREE");
    upload(hydro);
    const int mode = (n%ntreebuildstep == 0 || rebuild_tree) ? 0 : (n%ntreestockstep == 0 ? 1 : 2);
    check(gsph_build_tree(handle, mode, timestep));
    staged = true;
  }

  /// nbody->nbodydata -> {r[ndim], m, h} rows for gsph_set_stars (the fields
  /// GradhSph::ComputeStarGravForces reads, GradhSph.cpp:722-731)
  void set_stars(Sph<ndim>* sph, Nbody<ndim>* nbody)
  {
    const int Nstar = nbody ? nbody->Nnbody : 0;
    if (Nstar > 0 && sph->fixed_sink_mass)
      ExceptionHandler::getIstance().raise("GpuSphTree: fixed_sink_mass is not on the accelerated path");
    std::vector<double> rows((size_t) Nstar*(ndim + 2));
    for (int j = 0; j < Nstar; j++) {
      for (int k = 0; k < ndim; k++) rows[(size_t) j*(ndim + 2) + k] = nbody->nbodydata[j]->r[k];
      rows[(size_t) j*(ndim + 2) + ndim] = nbody->nbodydata[j]->m;
      rows[(size_t) j*(ndim + 2) + ndim + 1] = nbody->nbodydata[j]->h;
    }
    check(gsph_set_stars(handle, Nstar > 0 ? rows.data() : NULL, Nstar));
  }

  /// nbody->stardata -> the same rows for the star terms of zeta in the density pass
  /// (GradhSph::ComputeH loops over stardata[0..Nstar), GradhSph.cpp:288-306)
  void set_stars_for_zeta(Nbody<ndim>* nbody)
  {
    const int Nstar = nbody ? nbody->Nstar : 0;
    std::vector<double> rows((size_t) Nstar*(ndim + 2));
    for (int j = 0; j < Nstar; j++) {
      for (int k = 0; k < ndim; k++) rows[(size_t) j*(ndim + 2) + k] = nbody->stardata[j].r[k];
      rows[(size_t) j*(ndim + 2) + ndim] = nbody->stardata[j].m;
      rows[(size_t) j*(ndim + 2) + ndim + 1] = nbody->stardata[j].h;
    }
    check(gsph_set_stars(handle, Nstar > 0 ? rows.data() : NULL, Nstar));
  }

  virtual void UpdateAllSphProperties(Sph<ndim>* sph, Nbody<ndim>* nbody)
  {
    CodeTiming::BlockTimer timer = this->timing->StartNewTimer("SPH_PROPERTIES");
    if (!staged) { upload(sph); check(gsph_build_tree(handle, 1, 0.0)); }
    set_stars_for_zeta(nbody);
    check(gsph_update_properties(handle));
    download(sph, GSPH_F_DENSITY);
    staged = false;                       // the host zeroes accelerations / moves data after this call
  }

  virtual void UpdateAllSphHydroForces(Sph<ndim>* sph, Nbody<ndim>* nbody, DomainBox<ndim>& simbox)
  {
    CodeTiming::BlockTimer timer = this->timing->StartNewTimer("SPH_HYDRO_FORCES");
    upload(sph);                          // picks up Sph::ZeroAccelerations and the active flags
    check(gsph_build_tree(handle, 3, 0.0));
    set_stars(sph, nbody);                // ComputeStarGravForces, GradhSphTree.cpp:385-392
    check(gsph_update_hydro_forces(handle));
    download(sph, GSPH_F_FORCES);
    staged = false;
    (void) simbox;
  }

  virtual void UpdateAllSphForces(Sph<ndim>* sph, Nbody<ndim>* nbody, DomainBox<ndim>& simbox,
                                  Ewald<ndim>* ewald)
  {
    CodeTiming::BlockTimer timer = this->timing->StartNewTimer("SPH_ALL_FORCES");
    upload(sph);
    check(gsph_build_tree(handle, 3, 0.0));     // restock; amin / macfactor keep their BuildTree values
    set_stars(sph, nbody);                // ComputeStarGravForces, GradhSphTree.cpp:600-607
    check(gsph_update_all_forces(handle, NULL, 0));
    download(sph, GSPH_F_FORCES);
    staged = false;
    (void) simbox; (void) ewald;
  }

  /// Gravity of the gas on the active stars (HydroTree.cpp:552-654) from the device copy of the tree and
  /// the particles of the force pass that MainLoop has just run (SphSimulation.cpp:786-792).
  virtual void UpdateAllStarGasForces(Hydrodynamics<ndim>* hydro, Nbody<ndim>* nbody, DomainBox<ndim>& simbox,
                                      Ewald<ndim>* ewald)
  {
    CodeTiming::BlockTimer timer = this->timing->StartNewTimer("STAR_GAS_GRAV_FORCES");
    std::vector<int> act;
    for (int i = 0; i < nbody->Nstar; i++) if (nbody->nbodydata[i]->flags.check(active)) act.push_back(i);
    const int ns = (int) act.size();
    if (ns == 0) return;
    const int row = 2*ndim + 1;
    std::vector<double> rows((size_t) ns*row), a((size_t) ns*ndim, 0.0), adot((size_t) ns*ndim, 0.0), gpot(ns, 0.0);
    for (int j = 0; j < ns; j++) {
      NbodyParticle<ndim>* st = nbody->nbodydata[act[j]];
      for (int k = 0; k < ndim; k++) { rows[(size_t) j*row + k] = st->r[k]; rows[(size_t) j*row + ndim + k] = st->v[k]; }
      rows[(size_t) j*row + 2*ndim] = st->h;
    }
    check(gsph_update_star_gas_forces(handle, rows.data(), ns, a.data(), adot.data(), gpot.data(), NULL));
    for (int j = 0; j < ns; j++) {
      NbodyParticle<ndim>* st = nbody->nbodydata[act[j]];
      for (int k = 0; k < ndim; k++) { st->a[k] += a[(size_t) j*ndim + k]; st->adot[k] += adot[(size_t) j*ndim + k]; }
      st->gpot += gpot[j];
    }
    (void) hydro; (void) simbox; (void) ewald;
  }

  virtual double GetMaximumSmoothingLength() const
  {
    double hmax = 0.0;
    if (gsph_get_hmax(handle, &hmax) != GSPH_OK) return Base::GetMaximumSmoothingLength();
    return hmax;
  }

  //-----------------------------------------------------------------------------------------------
  // Device-resident main loop (include/gsph.h, gsph_integrator_*): for sph_integration = lfkdk (one global
  // time step or hierarchical block time steps) and no star particles, SphSimulation::MainLoop itself
  // (SphSimulation.cpp:574-880) runs on the device; the particle array crosses PCIe once in each direction.
  //-----------------------------------------------------------------------------------------------
  /// Call after PostInitialConditionsSetup (or any completed MainLoop): ships the complete particle
  /// state and starts the device loop at simulation time t.
  void ResidentBegin(Sph<ndim>* sph, Parameters* params, const double t) { ResidentBegin(sph, NULL, params, t); }

  /// ... with the star particles of `nbody` (nbody = lfkdk or the default hermite4, nbody_softening = 1, open
  /// boundaries): their integrator runs inside the device loop (host side of libgsph, around the gas-on-star pass).
  void ResidentBegin(Sph<ndim>* sph, Nbody<ndim>* nbody, Parameters* params, const double t)
  {
    std::map<std::string, int>& ip = params->intparams;
    std::map<std::string, double>& fp = params->floatparams;
    std::map<std::string, std::string>& sp = params->stringparams;
    if (sp["sph_integration"] != "lfkdk" || ip["Nlevels"] < 1)
      ExceptionHandler::getIstance().raise("GpuSphTree: the device-resident main loop needs sph_integration = lfkdk");
    // hierarchical block time steps (Nlevels > 1): the device loop starts with a resynchronisation of the hierarchy like
    // PostInitialConditionsSetup, so begin there (or at a step with n == nresync)

    // MainLoop's sink search / accretion / stellar-property update (SphSimulation.cpp:815-833) is not on this path
    if (ip["sink_particles"] != 0 || ip["create_sinks"] != 0)
      ExceptionHandler::getIstance().raise("GpuSphTree: sink_particles / create_sinks are not on the accelerated path of the device-resident main loop");
    gsph_integrator_config ic;
    memset(&ic, 0, sizeof ic);
    ic.accel_mult = fp["accel_mult"]; ic.courant_mult = fp["courant_mult"]; ic.energy_mult = fp["energy_mult"];
    ic.energy_integration = sp["gas_eos"] == "energy_eqn" ? 1 : 0;      // GradhSphSimulation.cpp:110-123
    ic.ntreebuildstep = ip["ntreebuildstep"]; ic.ntreestockstep = ip["ntreestockstep"];
    ic.sink_particles = ip["sink_particles"];
    ic.Nlevels = ip["Nlevels"]; ic.level_diff_max = ip["level_diff_max"];
    ic.sph_single_timestep = ip["sph_single_timestep"] != 0 ? 1 : 0;
    resident_levels = ic.Nlevels; resident_energy = ic.energy_integration != 0;
    check(gsph_set_stars(handle, NULL, 0));
    resident_stars = 0;
    if (nbody != NULL && nbody->Nstar > 0) {
      const bool h4 = sp["nbody"] == "hermite4";
      if (!h4 && sp["nbody"] != "lfkdk")
        ExceptionHandler::getIstance().raise("GpuSphTree: nbody = " + sp["nbody"] + " is not on the accelerated path (lfkdk, hermite4)");
      if (ic.Nlevels > 1)
        ExceptionHandler::getIstance().raise("GpuSphTree: the hand-over of star particles on block time steps is not built (start such a run with gsph_integrator_begin)");
      const int ns = nbody->Nstar, row = (h4 ? 6 : 3)*ndim + 2;
      std::vector<double> rows((size_t) ns*row);
      for (int s = 0; s < ns; s++) {
        const StarParticle<ndim>& st = nbody->stardata[s];
        double* q = &rows[(size_t) s*row];
        for (int k = 0; k < ndim; k++) {
          q[k] = st.r[k]; q[ndim + k] = st.v[k]; q[2*ndim + k] = st.a[k];
          if (h4) { q[3*ndim + k] = st.adot[k]; q[4*ndim + k] = st.a2dot[k]; q[5*ndim + k] = st.a3dot[k]; }
        }
        q[row - 2] = st.m; q[row - 1] = st.h;
      }
      check(h4 ? gsph_integrator_set_stars_hermite4(handle, rows.data(), ns, fp["nbody_mult"])
               : gsph_integrator_set_stars(handle, rows.data(), ns, fp["nbody_mult"]));
      resident_stars = ns; resident_h4 = h4;
    }
    upload(sph);
    check(gsph_integrator_begin(handle, &ic, 0, t));
    staged = false;
  }

  void ResidentMainLoop(const int nsteps) { check(gsph_integrator_step(handle, nsteps)); }

  /// Brings the particles home in the state the reference holds at the end of MainLoop (r0 = r, v0 = v,
  /// a0 = a, u0 = u, dudt0 = dudt, tlast = t, dt = the next step; SphLeapfrogKDK::EndTimestep,
  /// Simulation::ComputeGlobalTimestep), so that its own MainLoop can carry on from there.  With block time
  /// steps every particle sits somewhere inside its own step: its level, nstep, nlast, dt, dt_next, tlast and the
  /// start-of-step state r0, v0, a0, u0, dudt0 come from the device as they are, and *blk receives the scalars of
  /// the hierarchy (Simulation::n, nresync, level_max, level_step, dt_max).
  /// ... and the star particles as Nbody*::EndTimestep leaves them (r0 = r, v0 = v, a0 = a, adot0 = adot, tlast = t).
  void ResidentFetch(Sph<ndim>* sph, Nbody<ndim>* nbody, double* t_out, double* timestep_out, long long* nsteps_out)
  {
    double t = 0.0;
    ResidentFetch(sph, &t, timestep_out, nsteps_out, NULL);
    if (t_out) *t_out = t;
    if (nbody == NULL || resident_stars == 0) return;
    const int ns = resident_stars;
    std::vector<double> rows((size_t) ns*(3*ndim + 3)), ad((size_t) ns*ndim), a2((size_t) ns*ndim), a3((size_t) ns*ndim);
    check(gsph_integrator_get_stars(handle, rows.data(), ns));
    if (resident_h4) check(gsph_integrator_get_star_derivatives(handle, ad.data(), a2.data(), a3.data()));
    for (int s = 0; s < ns; s++) {
      StarParticle<ndim>& st = nbody->stardata[s];
      const double* q = &rows[(size_t) s*(3*ndim + 3)];
      for (int k = 0; k < ndim; k++) {
        st.r[k] = st.r0[k] = q[k]; st.v[k] = st.v0[k] = q[ndim + k]; st.a[k] = st.a0[k] = q[2*ndim + k];
        if (resident_h4) { st.adot[k] = st.adot0[k] = ad[(size_t) s*ndim + k]; st.a2dot[k] = a2[(size_t) s*ndim + k]; st.a3dot[k] = a3[(size_t) s*ndim + k]; }
      }
      st.gpot = q[3*ndim + 2];
      st.level = 0; st.nstep = 1; st.nlast = 0; st.tlast = t; st.dt_next = 0;
      st.flags.unset(active); st.flags.unset(end_timestep);
    }
  }

  void ResidentFetch(Sph<ndim>* sph, double* t_out, double* timestep_out, long long* nsteps_out, gsph_block_state* blk = NULL)
  {
    download(sph, GSPH_F_ALL);
    double t = 0.0, ts = 0.0; int64_t n = 0;
    check(gsph_integrator_state(handle, &t, &ts, &n));
    Part* p = sph->template GetParticleArray<GradhSphParticle>();
    const size_t N = (size_t) sph->Nhydro;
    if (resident_levels > 1) {
      std::vector<int32_t> nstep(N), nlast(N);
      std::vector<double> dt(N), tlast(N), dtn(N), u0(N), dudt0(N), r0(ndim*N), v0(ndim*N), a0(ndim*N);
      gsph_block_state st;
      check(gsph_integrator_block_state(handle, &st, nstep.data(), nlast.data(), dt.data(), tlast.data()));
      check(gsph_integrator_start_state(handle, r0.data(), v0.data(), a0.data(), u0.data(), dudt0.data(), dtn.data()));
      for (size_t i = 0; i < N; i++) {
        if (p[i].flags.is_dead()) continue;
        for (int k = 0; k < ndim; k++) { p[i].r0[k] = r0[k*N + i]; p[i].v0[k] = v0[k*N + i]; p[i].a0[k] = a0[k*N + i]; }
        if (resident_energy) { p[i].u0 = u0[i]; p[i].dudt0 = dudt0[i]; }
        p[i].nstep = nstep[i]; p[i].nlast = nlast[i];            // (level, levelneib came with the download)
        p[i].tlast = tlast[i]; p[i].dt = dt[i]; p[i].dt_next = dtn[i];
        p[i].flags.unset(active); p[i].flags.unset(end_timestep);
      }
      if (blk) *blk = st;
    }
    else {
      for (size_t i = 0; i < N; i++) {
        if (p[i].flags.is_dead()) continue;
        for (int k = 0; k < ndim; k++) { p[i].r0[k] = p[i].r[k]; p[i].v0[k] = p[i].v[k]; p[i].a0[k] = p[i].a[k]; }
        p[i].u0 = p[i].u; p[i].dudt0 = p[i].dudt;
        p[i].level = 0; p[i].levelneib = 0; p[i].nstep = 1; p[i].nlast = 0;
        p[i].tlast = t; p[i].dt = ts; p[i].dt_next = 0;
        p[i].flags.unset(active); p[i].flags.unset(end_timestep);
      }
      if (blk) { blk->n = 0; blk->nresync = 0; blk->level_max = 0; blk->level_step = 0; blk->dt_max = ts; }
    }
    if (t_out) *t_out = t;
    if (timestep_out) *timestep_out = ts;
    if (nsteps_out) *nsteps_out = (long long) n;
  }

  gsph_handle* GetHandle() { return handle; }
};

#endif
