//=================================================================================================
//  oracle/ref_harness.cpp  --  TEST INFRASTRUCTURE ONLY (never part of the product path).
//
//  A small driver of OUR OWN that links against the unmodified reference objects
//  (oracle/_ref/{std,omp}/libgandalf_ref.a, built by oracle/Makefile from /root/reference/src)
//  and runs the reference's own implementation of the grad-h SPH neighbour path so that its
//  results can be dumped at full double precision (the reference's snapshot writers print only
//  6 significant digits and omit a, du/dt: src/Common/SimulationIO.hpp:193-230,2073-2106).
//
//  It calls exactly the operator surface that the product replaces:
//      SphNeighbourSearch<ndim>::BuildTree / SearchBoundaryGhostParticles / BuildGhostTree
//      ::UpdateAllSphProperties   (src/GradhSph/GradhSphTree.cpp:83-271)
//      ::UpdateAllSphHydroForces  (src/GradhSph/GradhSphTree.cpp:280-435)
//      ::UpdateAllSphForces       (src/GradhSph/GradhSphTree.cpp:444-657)
//  in the order SphSimulation::PostInitialConditionsSetup uses them
//  (src/Hydrodynamics/SphSimulation.cpp:390-465), and -- for neighbour-list parity -- the
//  reference's own list machinery (Tree::ComputeNeighbourAndGhostList, Tree.cpp:562-617;
//  NeighbourManager::EndSearch/GetParticleNeib, NeighbourManager.h:265-327,483-543).
//
//  usage:
//    ref_harness PARAMS.dat OUTDIR [--in IC.bin] [--stars STARS.bin] [--set key=value]... [--perturb-h AMP]
//                [--dump setup,density,forces,neib,tree,gravlists] [--bench STEPS] [--quiet]
//                [--mainloop N [--gpu]]
//  --mainloop N : run N steps of the reference's own SphSimulation::MainLoop and dump the particles
//                 ("loop.*"); with --gpu (ref_harness_gpu build only) sim->sphneib is first replaced by
//                 gandalf_b200/host/GpuSphTree.h, i.e. the reference drives the B200 path.
//
//  IC.bin  : "GSPHIC1\0", int32 ndim, int32 pad, int64 N, then N doubles each of
//            x[,y[,z]], vx[,vy[,vz]], m, h, u   (h<=0 everywhere => let the reference guess h)
//  output  : OUTDIR/<stage>.<field>.npy   (numpy .npy v1.0, little endian)
//  --bench : time STEPS passes of {BuildTree, ghosts, UpdateAllSphProperties, forces} with all
//            particles active (with --perturb-h AMP the converged h is perturbed before each pass
//            so that the h-rho iteration has real work, as after a time step);
//            prints one JSON line with the wall-clock per block.
//            particles active and print one JSON line with the wall-clock per block.
//=================================================================================================
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <string>
#include <vector>
#include <map>
#include <algorithm>
#include <chrono>
#include <iostream>
#include <sstream>
#include <unistd.h>
#include <sys/stat.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "Exception.h"
#include "Parameters.h"
#include "Simulation.h"
#include "Sph.h"
#include "Particle.h"
#include "Nbody.h"
#include "Ghosts.h"
#include "NeighbourSearch.h"
#include "SphNeighbourSearch.h"
#include "NeighbourManager.h"
#include "KDTree.h"
#include "TreeCell.h"
#ifdef WITH_GPU
#include "GpuSphTree.h"      // gandalf_b200/host: the drop-in subclass under test (links libgsph.so)
#endif

using namespace std;

static string g_outdir;
static bool g_quiet = false;
static int g_mainloop = 0;         // --mainloop N : run N steps of the reference's own MainLoop
static string g_starfile;         // --stars FILE : int64 Nstar, then Nstar doubles each of x[,y[,z]], m, h
static string g_snapshot;          // --snapshot FILE : write the post-setup state with the reference's own "su" writer
static string g_snapshot_sf;       // --snapshot-sf FILE : the same in the formatted "sf" text format
static string g_snapshot_col;      // --snapshot-column FILE : the same in the "column" text format
static int g_resident = 0;         // --resident K : (with --gpu) the first K of the --mainloop steps run in the device-resident loop
static bool g_gpu = false;         // --gpu : swap sim->sphneib for GpuSphTree before doing so
static FILE* g_out = NULL;          // the caller's stdout (the reference's chatter goes elsewhere)

//-------------------------------------------------------------------------------------------------
// minimal .npy writer
//-------------------------------------------------------------------------------------------------
static void write_npy(const string& name, const char* descr, size_t itemsize,
                      const void* data, size_t n0, size_t n1 /*0 => 1-D*/)
{
  string path = g_outdir + "/" + name + ".npy";
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) { fprintf(stderr, "cannot open %s\n", path.c_str()); exit(2); }
  char shape[64];
  if (n1 == 0) snprintf(shape, sizeof shape, "(%zu,)", n0);
  else snprintf(shape, sizeof shape, "(%zu, %zu)", n0, n1);
  char hdr[256];
  int len = snprintf(hdr, sizeof hdr, "{'descr': '%s', 'fortran_order': False, 'shape': %s, }",
                     descr, shape);
  size_t total = 10 + len + 1;
  size_t pad = (64 - total % 64) % 64;
  uint16_t hlen = (uint16_t)(len + pad + 1);
  fwrite("\x93NUMPY\x01\x00", 1, 8, f);
  fwrite(&hlen, 2, 1, f);
  fwrite(hdr, 1, len, f);
  for (size_t i = 0; i < pad; i++) fputc(' ', f);
  fputc('\n', f);
  size_t count = n0 * (n1 ? n1 : 1);
  if (count) fwrite(data, itemsize, count, f);
  fclose(f);
}
static void dump_f64(const string& n, const vector<double>& v, size_t n0, size_t n1 = 0)
{ write_npy(n, "<f8", 8, v.data(), n0, n1); }
static void dump_i32(const string& n, const vector<int>& v, size_t n0, size_t n1 = 0)
{ write_npy(n, "<i4", 4, v.data(), n0, n1); }
static void dump_i64(const string& n, const vector<long long>& v, size_t n0)
{ write_npy(n, "<i8", 8, v.data(), n0, 0); }

static double now()
{
  return chrono::duration<double>(chrono::steady_clock::now().time_since_epoch()).count();
}

//-------------------------------------------------------------------------------------------------
template <int ndim>
struct Harness
{
  typedef GradhSphParticle<ndim> Part;
  typedef typename Part::HydroForcesParticle HydroParticle;

  SphSimulation<ndim>* sim;
  Sph<ndim>* sph;
  SphNeighbourSearch<ndim>* sphneib;

  Part* parts() { return static_cast<Part*>(sph->GetSphParticleArray()); }

  //-----------------------------------------------------------------------------------------------
  void dump_particles(const string& stage)
  {
    const int N = sph->Nhydro;
    Part* p = parts();
    vector<double> r(N*ndim), v(N*ndim), a(N*ndim), atree(N*ndim);
    vector<double> m(N), h(N), hrangesqd(N), hfactor(N), sound(N), rho(N), pressure(N), u(N),
                   dudt(N), gpot(N), div_v(N), alpha(N), dalphadt(N), invomega(N), zeta(N);
    vector<int> iorig(N), ptype(N), flags(N), level(N), levelneib(N), nstep(N), nlast(N);
    vector<double> r0(N*ndim), v0(N*ndim), a0(N*ndim), u0(N), dudt0(N), pdt(N), dt_next(N), tlast(N);
    for (int i = 0; i < N; i++) {
      for (int k = 0; k < ndim; k++) { r0[i*ndim+k] = p[i].r0[k]; v0[i*ndim+k] = p[i].v0[k]; a0[i*ndim+k] = p[i].a0[k]; }
      u0[i] = p[i].u0; dudt0[i] = p[i].dudt0; pdt[i] = p[i].dt; dt_next[i] = p[i].dt_next; tlast[i] = p[i].tlast;
      nstep[i] = p[i].nstep; nlast[i] = p[i].nlast;
      for (int k = 0; k < ndim; k++) {
        r[i*ndim+k] = p[i].r[k]; v[i*ndim+k] = p[i].v[k];
        a[i*ndim+k] = p[i].a[k]; atree[i*ndim+k] = p[i].atree[k];
      }
      m[i] = p[i].m; h[i] = p[i].h; hrangesqd[i] = p[i].hrangesqd; hfactor[i] = p[i].hfactor;
      sound[i] = p[i].sound; rho[i] = p[i].rho; pressure[i] = p[i].pressure; u[i] = p[i].u;
      dudt[i] = p[i].dudt; gpot[i] = p[i].gpot; div_v[i] = p[i].div_v; alpha[i] = p[i].alpha;
      dalphadt[i] = p[i].dalphadt; invomega[i] = p[i].invomega; zeta[i] = p[i].zeta;
      iorig[i] = p[i].iorig; ptype[i] = p[i].ptype; flags[i] = (int) p[i].flags.get();
      level[i] = p[i].level; levelneib[i] = p[i].levelneib;
    }
    dump_f64(stage + ".r", r, N, ndim); dump_f64(stage + ".v", v, N, ndim);
    dump_f64(stage + ".a", a, N, ndim); dump_f64(stage + ".atree", atree, N, ndim);
    dump_f64(stage + ".m", m, N); dump_f64(stage + ".h", h, N);
    dump_f64(stage + ".hrangesqd", hrangesqd, N); dump_f64(stage + ".hfactor", hfactor, N);
    dump_f64(stage + ".sound", sound, N); dump_f64(stage + ".rho", rho, N);
    dump_f64(stage + ".pressure", pressure, N); dump_f64(stage + ".u", u, N);
    dump_f64(stage + ".dudt", dudt, N); dump_f64(stage + ".gpot", gpot, N);
    dump_f64(stage + ".div_v", div_v, N); dump_f64(stage + ".alpha", alpha, N);
    dump_f64(stage + ".dalphadt", dalphadt, N); dump_f64(stage + ".invomega", invomega, N);
    dump_f64(stage + ".zeta", zeta, N);
    dump_i32(stage + ".iorig", iorig, N); dump_i32(stage + ".ptype", ptype, N);
    dump_i32(stage + ".flags", flags, N); dump_i32(stage + ".level", level, N);
    dump_i32(stage + ".levelneib", levelneib, N);
    // integrator state (block time steps, Particle.h: nstep, nlast, dt, dt_next, tlast, r0, v0, a0, u0, dudt0)
    dump_i32(stage + ".nstep", nstep, N); dump_i32(stage + ".nlast", nlast, N);
    dump_f64(stage + ".dt", pdt, N); dump_f64(stage + ".dt_next", dt_next, N); dump_f64(stage + ".tlast", tlast, N);
    dump_f64(stage + ".r0", r0, N, ndim); dump_f64(stage + ".v0", v0, N, ndim); dump_f64(stage + ".a0", a0, N, ndim);
    dump_f64(stage + ".u0", u0, N); dump_f64(stage + ".dudt0", dudt0, N);
    // star particles (the N-body side of the main loop)
    const int Ns = sim->nbody->Nstar;
    if (Ns > 0) {
      vector<double> sr(Ns*ndim), sv(Ns*ndim), sa(Ns*ndim), sm(Ns), sh(Ns), sg(Ns), sad(Ns*ndim), sa2(Ns*ndim), sa3(Ns*ndim);
      for (int i = 0; i < Ns; i++) {
        for (int k = 0; k < ndim; k++) {
          sr[i*ndim+k] = sim->nbody->stardata[i].r[k]; sv[i*ndim+k] = sim->nbody->stardata[i].v[k];
          sa[i*ndim+k] = sim->nbody->stardata[i].a[k];
          sad[i*ndim+k] = sim->nbody->stardata[i].adot[k]; sa2[i*ndim+k] = sim->nbody->stardata[i].a2dot[k];
          sa3[i*ndim+k] = sim->nbody->stardata[i].a3dot[k];
        }
        sm[i] = sim->nbody->stardata[i].m; sh[i] = sim->nbody->stardata[i].h; sg[i] = sim->nbody->stardata[i].gpot;
      }
      dump_f64(stage + ".star_r", sr, Ns, ndim); dump_f64(stage + ".star_v", sv, Ns, ndim);
      dump_f64(stage + ".star_a", sa, Ns, ndim); dump_f64(stage + ".star_m", sm, Ns);
      dump_f64(stage + ".star_h", sh, Ns); dump_f64(stage + ".star_gpot", sg, Ns);
      // block time steps of the stars (NbodyParticle: level, nstep, nlast, tlast)
      {
        vector<int> sl(Ns), sns(Ns), snl(Ns); vector<double> stl(Ns);
        for (int i = 0; i < Ns; i++) {
          sl[i] = sim->nbody->stardata[i].level; sns[i] = sim->nbody->stardata[i].nstep; snl[i] = sim->nbody->stardata[i].nlast;
          stl[i] = sim->nbody->stardata[i].tlast;
        }
        dump_i32(stage + ".star_level", sl, Ns); dump_i32(stage + ".star_nstep", sns, Ns); dump_i32(stage + ".star_nlast", snl, Ns);
        dump_f64(stage + ".star_tlast", stl, Ns);
      }
      // Hermite-4 derivatives (NbodyParticle: adot, a2dot, a3dot)
      dump_f64(stage + ".star_adot", sad, Ns, ndim); dump_f64(stage + ".star_a2dot", sa2, Ns, ndim);
      dump_f64(stage + ".star_a3dot", sa3, Ns, ndim);
    }
  }

  //-----------------------------------------------------------------------------------------------
  void set_all_active()
  {
    for (int i = 0; i < sph->Nhydro; i++) sph->GetSphParticlePointer(i).flags.set(active);
  }

  // BuildTree + boundary ghosts + ghost tree, as SphSimulation.cpp:396-401
  void build_trees()
  {
    sphneib->BuildTree(true, 0, sim->ntreebuildstep, sim->ntreestockstep, sim->timestep, sph);
    sphneib->SearchBoundaryGhostParticles((FLOAT) 0.0, sim->simbox, sph);
    sphneib->BuildGhostTree(true, 0, sim->ntreebuildstep, sim->ntreestockstep, sim->timestep, sph);
  }

  // zero the force accumulators exactly as SphSimulation.cpp:409-422 does
  void zero_forces()
  {
    for (int i = 0; i < sph->Nhydro; i++) {
      SphParticle<ndim>& part = sph->GetSphParticlePointer(i);
      part.levelneib = 0;
      part.dalphadt  = (FLOAT) 0.0;
      part.div_v     = (FLOAT) 0.0;
      part.dudt      = (FLOAT) 0.0;
      part.gpot      = (FLOAT) 0.0;
      part.gpot_hydro= (FLOAT) 0.0;
      for (int k = 0; k < ndim; k++) part.a[k] = (FLOAT) 0.0;
      for (int k = 0; k < ndim; k++) part.atree[k] = (FLOAT) 0.0;
    }
  }

  void run_forces()
  {
    if (sph->self_gravity == 1)
      sphneib->UpdateAllSphForces(sph, sim->nbody, sim->simbox, sim->ewald);
    else
      sphneib->UpdateAllSphHydroForces(sph, sim->nbody, sim->simbox);
  }

  //-----------------------------------------------------------------------------------------------
  // Deterministic perturbation of h so that the h-rho iteration has real work to do.
  //-----------------------------------------------------------------------------------------------
  void perturb_h(double amp)
  {
    Part* p = parts();
    for (int i = 0; i < sph->Nhydro; i++) {
      uint32_t x = (uint32_t) p[i].iorig * 2654435761u;
      x ^= x >> 16; x *= 2246822519u; x ^= x >> 13;
      double f = (double)(x & 0xFFFFFF) / (double) 0x1000000;      // [0,1)
      p[i].h *= (1.0 + amp*(2.0*f - 1.0));
      p[i].hrangesqd = sph->kernp->kernrangesqd*p[i].h*p[i].h;
    }
  }

  //-----------------------------------------------------------------------------------------------
  // True reference neighbour lists for the hydro force pass, taken from the reference's own
  // tree walk + NeighbourManager (same calls as GradhSphTree.cpp:327-360), stored as CSR of
  // neighbour iorig (sorted per particle).  With self-gravity on, also the per-particle
  // direct-sum and smoothed-gravity lists and the per-leaf accepted cell ids
  // (GradhSphTree.cpp:498-531).
  //-----------------------------------------------------------------------------------------------
  void dump_neighbour_lists(const string& stage, bool gravlists)
  {
    const int N = sph->Nhydro;
    Part* sphdata = parts();
    TreeBase<ndim>* tree = sphneib->GetTree();
    vector<TreeCellBase<ndim> > celllist;
    int cactive = tree->ComputeActiveCellList(celllist);
    NeighbourManager<ndim,HydroParticle> neibmanager(sph, sim->simbox);
    vector<int> activelist(tree->MaxNumPartInLeafCell() + 8);

    vector<vector<int> > hyd(N), dir(N), smo(N);
    vector<int> leaf_of(N, -1);
    vector<long long> cell_off(1, 0);
    vector<int> cell_ids, cell_leafid;

    for (int cc = 0; cc < cactive; cc++) {
      TreeCellBase<ndim>& cell = celllist[cc];
      const int Nactive = tree->ComputeActiveParticleList(cell, sphdata, &activelist[0]);
      neibmanager.clear();
      if (gravlists) {
        tree->ComputeGravityInteractionAndGhostList(cell, neibmanager);
        neibmanager.EndSearchGravity(cell, sphdata);
        MultipoleMoment<ndim>* gravcell;
        const int Ngravcell = neibmanager.GetGravCell(&gravcell);
        for (int j = 0; j < Ngravcell; j++) cell_ids.push_back(gravcell[j].id);
        cell_off.push_back((long long) cell_ids.size());
        cell_leafid.push_back(cell.id);
      }
      else {
        tree->ComputeNeighbourAndGhostList(cell, neibmanager);
        neibmanager.EndSearch(cell, sphdata);
      }
      for (int j = 0; j < Nactive; j++) {
        const int i = activelist[j];
        Part& part = sphdata[i];
        leaf_of[i] = cell.id;
        Typemask hydromask = sph->types[part.ptype].hydromask;
        if (gravlists) {
          GravityNeighbourLists<HydroParticle> nl = neibmanager.GetParticleNeibGravity(part, hydromask);
          for (size_t n = 0; n < nl.neiblist.size(); n++) hyd[i].push_back(nl.neiblist[n].iorig);
          for (size_t n = 0; n < nl.smooth_gravlist.size(); n++) smo[i].push_back(nl.smooth_gravlist[n].iorig);
          for (size_t n = 0; n < nl.directlist.size(); n++) dir[i].push_back(nl.directlist[n].iorig);
        }
        else {
          NeighbourList<HydroParticle> nl = neibmanager.GetParticleNeib(part, hydromask, false);
          for (size_t n = 0; n < nl.size(); n++) hyd[i].push_back(nl[n].iorig);
        }
      }
    }

    auto dump_csr = [&](const string& nm, vector<vector<int> >& L) {
      vector<long long> off(N + 1, 0);
      vector<int> flat;
      // store rows in iorig order so that the file is independent of array order
      vector<int> row_of_iorig(N);
      for (int i = 0; i < N; i++) row_of_iorig[sphdata[i].iorig] = i;
      for (int io = 0; io < N; io++) {
        vector<int>& l = L[row_of_iorig[io]];
        sort(l.begin(), l.end());
        flat.insert(flat.end(), l.begin(), l.end());
        off[io + 1] = (long long) flat.size();
      }
      dump_i64(stage + "." + nm + "_off", off, N + 1);
      dump_i32(stage + "." + nm + "_idx", flat, flat.size());
    };
    dump_csr("neib", hyd);
    if (gravlists) {
      dump_csr("direct", dir);
      dump_csr("smooth", smo);
      vector<int> leaf_by_iorig(N);
      for (int i = 0; i < N; i++) leaf_by_iorig[sphdata[i].iorig] = leaf_of[i];
      dump_i32(stage + ".leaf_of_iorig", leaf_by_iorig, N);
      dump_i64(stage + ".gravcell_off", cell_off, cell_off.size());
      dump_i32(stage + ".gravcell_ids", cell_ids, cell_ids.size());
      dump_i32(stage + ".gravcell_leaf", cell_leafid, cell_leafid.size());
    }
  }

  //-----------------------------------------------------------------------------------------------
  // KD-tree dump (only valid for neib_search = kdtree): cell table + leaf membership.
  //-----------------------------------------------------------------------------------------------
  void dump_tree(const string& stage)
  {
    if (sim->simparams->stringparams["neib_search"] != "kdtree") return;
    typedef KDTree<ndim,GradhSphParticle,KDTreeCell> KD;
    KD* kd = static_cast<KD*>(sphneib->GetTree());
    Part* sphdata = parts();
    const int Nc = kd->Ncell;
    vector<int> meta(Nc*8);
    vector<double> box(Nc*ndim*4), com(Nc*ndim*2), sc(Nc*5), q(Nc*5);
    vector<int> leaf_of_iorig(sph->Nhydro, -1);
    for (int c = 0; c < Nc; c++) {
      KDTreeCell<ndim>& cell = kd->celldata[c];
      int* mm = &meta[c*8];
      mm[0] = cell.id; mm[1] = cell.level; mm[2] = cell.N; mm[3] = cell.copen; mm[4] = cell.cnext;
      mm[5] = cell.c1; mm[6] = cell.c2; mm[7] = cell.Nactive;
      for (int k = 0; k < ndim; k++) {
        box[(c*4+0)*ndim+k] = cell.bb.min[k];   box[(c*4+1)*ndim+k] = cell.bb.max[k];
        box[(c*4+2)*ndim+k] = cell.hbox.min[k]; box[(c*4+3)*ndim+k] = cell.hbox.max[k];
        com[(c*2+0)*ndim+k] = cell.rcell[k];    com[(c*2+1)*ndim+k] = cell.r[k];
      }
      sc[c*5+0] = cell.m; sc[c*5+1] = cell.rmax; sc[c*5+2] = cell.hmax;
      sc[c*5+3] = cell.cdistsqd; sc[c*5+4] = cell.mac;
      for (int k = 0; k < 5; k++) q[c*5+k] = cell.q[k];
      if (cell.copen == -1 && cell.N > 0) {
        int i = cell.ifirst;
        while (i != -1) {
          leaf_of_iorig[sphdata[i].iorig] = c;
          if (i == cell.ilast) break;
          i = kd->inext[i];
        }
      }
    }
    vector<int> hdr(4); hdr[0] = Nc; hdr[1] = kd->ltot; hdr[2] = kd->gtot; hdr[3] = kd->Nleafmax;
    dump_i32(stage + ".tree_hdr", hdr, 4);
    dump_i32(stage + ".tree_meta", meta, Nc, 8);
    dump_f64(stage + ".tree_box", box, Nc, ndim*4);
    dump_f64(stage + ".tree_com", com, Nc, ndim*2);
    dump_f64(stage + ".tree_scalars", sc, Nc, 5);
    dump_f64(stage + ".tree_q", q, Nc, 5);
    dump_i32(stage + ".tree_leaf_of_iorig", leaf_of_iorig, sph->Nhydro);
  }
};

//-------------------------------------------------------------------------------------------------
template <int ndim>
static int run(Parameters* params, const string& icfile, double perturb, const string& dumps,
               int bench_steps)
{
  auto want = [&](const char* w) { return (("," + dumps + ",").find(string(",") + w + ",") != string::npos); };

  // the simulation object copies the Parameters, so ic/Nhydro must be set before it is created
  FILE* f = NULL;
  int64_t N = 0;
  if (!icfile.empty()) {
    f = fopen(icfile.c_str(), "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", icfile.c_str()); return 2; }
    char magic[8]; int32_t nd, pad;
    if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "GSPHIC1", 7) != 0) { fprintf(stderr, "bad IC magic\n"); return 2; }
    if (fread(&nd, 4, 1, f) != 1 || fread(&pad, 4, 1, f) != 1 || fread(&N, 8, 1, f) != 1) return 2;
    if (nd != ndim) { fprintf(stderr, "IC ndim %d != params ndim %d\n", nd, ndim); return 2; }
    params->stringparams["ic"] = "python";
    params->intparams["Nhydro"] = (int) N;
  }
  // star particles (at rest) through the reference's own python-IC import (Simulation.cpp:1325-1420)
  int64_t Nstar = 0;
  vector<vector<double> > starcols;
  if (!g_starfile.empty()) {
    FILE* fs = fopen(g_starfile.c_str(), "rb");
    if (!fs || fread(&Nstar, 8, 1, fs) != 1) { fprintf(stderr, "cannot read %s\n", g_starfile.c_str()); return 2; }
    starcols.assign(ndim + 2, vector<double>(Nstar));
    for (int c = 0; c < ndim + 2; c++)
      if (fread(starcols[c].data(), 8, Nstar, fs) != (size_t) Nstar) { fprintf(stderr, "short star file\n"); return 2; }
    fclose(fs);
    params->intparams["Nstar"] = (int) Nstar;
  }

  SimulationBase* sb = SimulationBase::SimulationFactory(ndim, params->stringparams["sim"], params);
  SphSimulation<ndim>* sim = static_cast<SphSimulation<ndim>*>(sb);

  bool h_given = false;
  if (f) {
    sb->PreSetupForPython();
    static const char* names[3][3] = {{"x","y","z"},{"vx","vy","vz"},{"m","h","u"}};
    vector<double> buf(N);
    for (int grp = 0; grp < 3; grp++) {
      int cnt = (grp < 2) ? ndim : 3;
      for (int k = 0; k < cnt; k++) {
        if (fread(buf.data(), 8, N, f) != (size_t) N) { fprintf(stderr, "short IC file\n"); return 2; }
        if (string(names[grp][k]) == "h") {
          h_given = (buf[0] > 0.0);
          if (!h_given) continue;
        }
        sb->ImportArray(buf.data(), (int) N, names[grp][k], "sph");
      }
    }
    fclose(f);
    static const char* snames[3] = {"x", "y", "z"};
    for (int c = 0; c < ndim && Nstar > 0; c++) sb->ImportArray(starcols[c].data(), (int) Nstar, snames[c], "star");
    if (Nstar > 0) {
      sb->ImportArray(starcols[ndim].data(), (int) Nstar, "m", "star");
      sb->ImportArray(starcols[ndim + 1].data(), (int) Nstar, "h", "star");
    }
    sb->initial_h_provided = h_given;
  }

  double t0 = now();
  sb->SetupSimulation();
  double t_setup = now() - t0;

  Harness<ndim> H;
  H.sim = sim; H.sph = sim->sph; H.sphneib = sim->sphneib;
  if (!g_quiet) fprintf(stderr, "[ref_harness] setup done: Nhydro=%d Ntot=%d (%.2fs)\n",
                        H.sph->Nhydro, H.sph->Ntot, t_setup);

  if (want("setup")) {
    H.dump_particles("setup");
    fprintf(g_out, "{\"ref_setup\": true, \"n\": %d, \"nresync\": %d, \"level_max\": %d, \"level_step\": %d, \"dt_max\": %.17g, "
            "\"timestep\": %.17g, \"t\": %.17g, \"Nlevels\": %d}\n", (int) sim->n, (int) sim->nresync, (int) sim->level_max,
            (int) sim->level_step, (double) sim->dt_max, (double) sim->timestep, (double) sim->t, (int) sim->Nlevels);
    fflush(g_out);
  }
  if (!g_snapshot.empty()) sb->WriteSnapshotFile(g_snapshot, "su");
  if (!g_snapshot_col.empty()) sb->WriteSnapshotFile(g_snapshot_col, "column");
  if (!g_snapshot_sf.empty()) sb->WriteSnapshotFile(g_snapshot_sf, "sf");

  // ---------------------------------------------------------------- operator-level captures
  if (want("density") || want("forces") || want("neib") || want("tree") || want("gravlists") || want("stargas")) {
    H.set_all_active();
    if (perturb != 0.0) H.perturb_h(perturb);
    // gpot is an input of the density pass only through the sink potmin test; keep as is.
    H.build_trees();
    if (want("density")) H.dump_particles("dens_in");
    H.sphneib->UpdateAllSphProperties(H.sph, sim->nbody);
    if (want("density")) H.dump_particles("dens_out");

    // force pass, same order as SphSimulation.cpp:394-465
    sim->LocalGhosts->CopyHydroDataToGhosts(sim->simbox, H.sph);
    H.build_trees();
    H.zero_forces();
    if (want("tree")) H.dump_tree("force_in");
    if (want("neib")) H.dump_neighbour_lists("force_in", false);
    if (want("gravlists") && H.sph->self_gravity == 1) H.dump_neighbour_lists("grav_in", true);
    if (want("forces")) H.dump_particles("force_in");
    H.run_forces();
    if (want("forces")) H.dump_particles("force_out");
    // gravity of the gas on the stars (HydroTree::UpdateAllStarGasForces), as MainLoop calls it
    // after the gas forces (SphSimulation.cpp:786-792)
    if (want("stargas") && sim->nbody->Nstar > 0 && H.sph->self_gravity == 1) {
      const int Ns = sim->nbody->Nstar;
      vector<double> sa(Ns*ndim), sadot(Ns*ndim), sgpot(Ns), sr(Ns*ndim), sv(Ns*ndim), sh(Ns), sm(Ns);
      for (int i = 0; i < Ns; i++) {
        NbodyParticle<ndim>* st = sim->nbody->nbodydata[i];
        st->flags.set(active);
        for (int k = 0; k < ndim; k++) { st->a[k] = 0.0; st->adot[k] = 0.0; }
        st->gpot = 0.0;
      }
      H.sphneib->UpdateAllStarGasForces(H.sph, sim->nbody, sim->simbox, sim->ewald);
      for (int i = 0; i < Ns; i++) {
        NbodyParticle<ndim>* st = sim->nbody->nbodydata[i];
        for (int k = 0; k < ndim; k++) {
          sa[i*ndim+k] = st->a[k]; sadot[i*ndim+k] = st->adot[k]; sr[i*ndim+k] = st->r[k]; sv[i*ndim+k] = st->v[k];
        }
        sgpot[i] = st->gpot; sh[i] = st->h; sm[i] = st->m;
      }
      dump_f64("stargas.a", sa, Ns, ndim); dump_f64("stargas.adot", sadot, Ns, ndim); dump_f64("stargas.gpot", sgpot, Ns);
      dump_f64("stargas.r", sr, Ns, ndim); dump_f64("stargas.v", sv, Ns, ndim); dump_f64("stargas.h", sh, Ns);
      dump_f64("stargas.m", sm, Ns);
    }
  }

  // ---------------------------------------------------------------- drop-in check: the reference's
  // own MainLoop (src/Hydrodynamics/SphSimulation.cpp:574-880) drives whichever neighbour-search
  // object sits behind sim->sphneib -- the CPU GradhSphTree, or (--gpu) our GpuSphTree.
  if (g_mainloop > 0) {
    if (g_gpu) {
#ifdef WITH_GPU
      Parameters* sp = sim->simparams;
      sim->sphneib = new GpuSphTree<ndim>(
          sp->intparams["Nleafmax"], 1, sp->intparams["pruning_level_min"], sp->intparams["pruning_level_max"],
          sp->floatparams["thetamaxsqd"], sim->sph->kernp->kernrange, sp->floatparams["macerror"],
          sp->stringparams["gravity_mac"], sp->stringparams["multipole"], &sim->simbox, sim->sph->kernp,
          sim->timing, sim->sph->types, sim->sph, sp, sim->simunits);
      sim->sphneib->SetTimingObject(sim->timing);
      H.sphneib = sim->sphneib;
#else
      fprintf(stderr, "--gpu needs the WITH_GPU build (ref_harness_gpu)\n"); return 2;
#endif
    }
    double t1 = now();
    const double timestep0 = sim->timestep;          // the step PostInitialConditionsSetup chose
    int done = 0;
#ifdef WITH_GPU
    if (g_gpu && g_resident > 0) {
      // the device-resident main loop takes the first K steps, then hands the particles back to the
      // reference's own MainLoop (gandalf_b200/host/GpuSphTree.h: ResidentBegin / MainLoop / Fetch)
      GpuSphTree<ndim>* gt = static_cast<GpuSphTree<ndim>*>(sim->sphneib);
      const int K = std::min(g_resident, g_mainloop);
      const bool with_stars = sim->nbody->Nstar > 0;
      if (with_stars) gt->ResidentBegin(sim->sph, sim->nbody, sim->simparams, sim->t);
      else gt->ResidentBegin(sim->sph, sim->simparams, sim->t);
      gt->ResidentMainLoop(K);
      double tt = 0.0, ts = 0.0; long long ns = 0;
      gsph_block_state bs;
      if (with_stars) { gt->ResidentFetch(sim->sph, sim->nbody, &tt, &ts, &ns); bs.n = 0; }
      else gt->ResidentFetch(sim->sph, &tt, &ts, &ns, &bs);
      sim->t = tt; sim->timestep = ts; sim->n = 0;
      if (sim->Nlevels > 1) {                         // the hierarchy of the block time steps as the device left it
        sim->n = (int) bs.n; sim->nresync = (int) bs.nresync; sim->level_max = bs.level_max;
        sim->level_step = bs.level_step; sim->dt_max = bs.dt_max;
      }
      sim->Nsteps += K; sim->Nfullsteps += K; sim->Nblocksteps += K;
      sim->rebuild_tree = false;
      done = K;
    }
#endif
    for (int s = done; s < g_mainloop; s++) sim->MainLoop();
    double t_loop = now() - t1;
    H.dump_particles("loop");
    fprintf(g_out, "{\"ref_mainloop\": true, \"gpu\": %s, \"N\": %d, \"steps\": %d, \"t\": %.17g, \"timestep\": %.17g, \"timestep0\": %.17g, \"wall_s\": %.4f, "
            "\"n\": %d, \"nresync\": %d, \"level_max\": %d, \"level_step\": %d, \"dt_max\": %.17g, \"Nlevels\": %d}\n",
            g_gpu ? "true" : "false", H.sph->Nhydro, g_mainloop, (double) sim->t, (double) sim->timestep, timestep0, t_loop,
            (int) sim->n, (int) sim->nresync, (int) sim->level_max, (int) sim->level_step, (double) sim->dt_max, (int) sim->Nlevels);
    fflush(g_out);
  }

  // ---------------------------------------------------------------- CPU timing baseline
  if (bench_steps > 0) {
    double tb = 0, tg = 0, tp = 0, tf = 0;
    int nthreads = 1;
#ifdef _OPENMP
    nthreads = omp_get_max_threads();
#endif
    for (int s = 0; s < bench_steps; s++) {
      H.set_all_active();
      if (perturb != 0.0) H.perturb_h(perturb);   // every step restarts from the same perturbed guess
      double a0 = now();
      H.sphneib->BuildTree(true, 0, sim->ntreebuildstep, sim->ntreestockstep, sim->timestep, H.sph);
      double a1 = now();
      H.sphneib->SearchBoundaryGhostParticles((FLOAT) 0.0, sim->simbox, H.sph);
      H.sphneib->BuildGhostTree(true, 0, sim->ntreebuildstep, sim->ntreestockstep, sim->timestep, H.sph);
      double a2 = now();
      H.sphneib->UpdateAllSphProperties(H.sph, sim->nbody);
      double a3 = now();
      sim->LocalGhosts->CopyHydroDataToGhosts(sim->simbox, H.sph);
      H.zero_forces();
      H.run_forces();
      double a4 = now();
      tb += a1 - a0; tg += a2 - a1; tp += a3 - a2; tf += a4 - a3;
    }
    double tot = tb + tg + tp + tf;
    fprintf(g_out, "{\"ref_bench\": true, \"N\": %d, \"steps\": %d, \"threads\": %d, "
           "\"BUILD_TREE\": %.6f, \"BUILD_GHOST_TREE\": %.6f, \"SPH_PROPERTIES\": %.6f, "
           "\"SPH_FORCES\": %.6f, \"total\": %.6f, \"updates_per_s\": %.6e, \"setup_s\": %.3f}\n",
           H.sph->Nhydro, bench_steps, nthreads, tb, tg, tp, tf, tot,
           (double) H.sph->Nhydro*bench_steps/tot, t_setup);
    fflush(g_out);
  }
  return 0;
}

//-------------------------------------------------------------------------------------------------
int main(int argc, char** argv)
{
  if (argc < 3) {
    fprintf(stderr, "usage: %s PARAMS.dat OUTDIR [--in IC.bin] [--set k=v]... [--perturb-h A] "
                    "[--dump list] [--bench STEPS] [--quiet]\n", argv[0]);
    return 1;
  }
  string paramfile = argv[1];
  g_outdir = argv[2];
  string icfile, dumps = "setup,density,forces,neib";
  double perturb = 0.0;
  int bench_steps = 0;
  vector<pair<string,string> > sets;
  for (int i = 3; i < argc; i++) {
    string a = argv[i];
    if (a == "--in" && i + 1 < argc) icfile = argv[++i];
    else if (a == "--stars" && i + 1 < argc) g_starfile = argv[++i];
    else if (a == "--resident" && i + 1 < argc) g_resident = atoi(argv[++i]);
    else if (a == "--snapshot" && i + 1 < argc) g_snapshot = argv[++i];
    else if (a == "--snapshot-column" && i + 1 < argc) g_snapshot_col = argv[++i];
    else if (a == "--snapshot-sf" && i + 1 < argc) g_snapshot_sf = argv[++i];
    else if (a == "--set" && i + 1 < argc) {
      string kv = argv[++i];
      size_t eq = kv.find('=');
      if (eq == string::npos) { fprintf(stderr, "bad --set %s\n", kv.c_str()); return 1; }
      sets.push_back(make_pair(kv.substr(0, eq), kv.substr(eq + 1)));
    }
    else if (a == "--perturb-h" && i + 1 < argc) perturb = atof(argv[++i]);
    else if (a == "--dump" && i + 1 < argc) dumps = argv[++i];
    else if (a == "--bench" && i + 1 < argc) bench_steps = atoi(argv[++i]);
    else if (a == "--quiet") g_quiet = true;
    else if (a == "--mainloop" && i + 1 < argc) g_mainloop = atoi(argv[++i]);
    else if (a == "--gpu") g_gpu = true;
    else { fprintf(stderr, "unknown argument %s\n", a.c_str()); return 1; }
  }

  // absolute paths before chdir (the reference writes run_id.* files into the cwd)
  char cwd[4096];
  if (!getcwd(cwd, sizeof cwd)) return 2;
  auto absolutise = [&](string& p) { if (!p.empty() && p[0] != '/') p = string(cwd) + "/" + p; };
  absolutise(paramfile); absolutise(g_outdir); absolutise(icfile); absolutise(g_starfile); absolutise(g_snapshot); absolutise(g_snapshot_col); absolutise(g_snapshot_sf);
  mkdir(g_outdir.c_str(), 0755);
  if (chdir(g_outdir.c_str()) != 0) { fprintf(stderr, "cannot chdir to %s\n", g_outdir.c_str()); return 2; }

  // silence the reference's stdout chatter (splash, diagnostics) unless benchmarking output
  g_out = fdopen(dup(1), "w");
  if (g_quiet) { FILE* nul = freopen("/dev/null", "w", stdout); (void) nul; }

  ExceptionHandler::makeExceptionHandler(cplusplus);
  Parameters* params = new Parameters();
  params->ReadParamsFile(paramfile);
  for (size_t i = 0; i < sets.size(); i++) {
    const string& k = sets[i].first; const string& v = sets[i].second;
    if (params->intparams.count(k)) params->intparams[k] = atoi(v.c_str());
    else if (params->floatparams.count(k)) params->floatparams[k] = atof(v.c_str());
    else if (params->stringparams.count(k)) params->stringparams[k] = v;
    else { fprintf(stderr, "unknown parameter %s\n", k.c_str()); return 1; }
  }
  const int ndim = params->intparams["ndim"];

  int rc;
  if (ndim == 1) rc = run<1>(params, icfile, perturb, dumps, bench_steps);
  else if (ndim == 2) rc = run<2>(params, icfile, perturb, dumps, bench_steps);
  else rc = run<3>(params, icfile, perturb, dumps, bench_steps);
  return rc;
}
