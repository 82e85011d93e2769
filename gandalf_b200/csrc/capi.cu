/*=================================================================================================
 *  capi.cu -- the C ABI of libgsph (include/gsph.h): lifecycle, host<->HBM data movement, the
 *  operator surface and the parity/inspection calls.  No CPU fallback: every compute entry point
 *  launches CUDA kernels and fails with GSPH_ENODEVICE / GSPH_ECUDA when that is impossible.
 *=================================================================================================*/
#include "gsph_internal.cuh"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

static std::string g_create_error;

namespace {

struct FieldMap { int pf; size_t off; };      /* offset of a gsph_field inside gsph_particle_view */
#define VOFF(member) offsetof(gsph_particle_view, member)
const FieldMap kDoubleFields[] = {
  {F_RX, VOFF(r[0])}, {F_RY, VOFF(r[1])}, {F_RZ, VOFF(r[2])},
  {F_VX, VOFF(v[0])}, {F_VY, VOFF(v[1])}, {F_VZ, VOFF(v[2])},
  {F_AX, VOFF(a[0])}, {F_AY, VOFF(a[1])}, {F_AZ, VOFF(a[2])},
  {F_ATX, VOFF(atree[0])}, {F_ATY, VOFF(atree[1])}, {F_ATZ, VOFF(atree[2])},
  {F_M, VOFF(m)}, {F_H, VOFF(h)}, {F_U, VOFF(u)}, {F_RHO, VOFF(rho)}, {F_HR2, VOFF(hrangesqd)},
  {F_HFAC, VOFF(hfactor)}, {F_INVOMEGA, VOFF(invomega)}, {F_ZETA, VOFF(zeta)},
  {F_SOUND, VOFF(sound)}, {F_PRESS, VOFF(pressure)}, {F_DUDT, VOFF(dudt)}, {F_DIVV, VOFF(div_v)},
  {F_ALPHA, VOFF(alpha)}, {F_DALPHADT, VOFF(dalphadt)}, {F_GPOT, VOFF(gpot)},
  {F_GPOTHYDRO, VOFF(gpot_hydro)},
};
const FieldMap kIntFields[] = {
  {I_FLAGS, VOFF(flags)}, {I_PTYPE, VOFF(ptype)}, {I_LEVEL, VOFF(level)},
  {I_LEVELNEIB, VOFF(levelneib)}, {I_IORIG, VOFF(iorig)}, {I_SINKID, VOFF(sinkid)},
};
const int kNDoubleFields = sizeof(kDoubleFields)/sizeof(kDoubleFields[0]);
const int kNIntFields = sizeof(kIntFields)/sizeof(kIntFields[0]);

inline const gsph_field& vfield(const gsph_particle_view* v, size_t off)
{
  return *reinterpret_cast<const gsph_field*>(reinterpret_cast<const char*>(v) + off);
}

double default_double(int pf) { return (pf == F_H || pf == F_INVOMEGA) ? 1.0 : 0.0; }
int default_int(int f)
{
  if (f == I_FLAGS) return (int) GSPH_FLAG_ACTIVE;     /* absent flags: every particle is active */
  if (f == I_SINKID) return -1;
  if (f == I_IORIG) return -1;
  return 0;
}

template <class T>
__global__ void fill_kernel(int64_t n, T* a, T v)
{
  const int64_t i = (int64_t) blockIdx.x*blockDim.x + threadIdx.x;
  if (i < n) a[i] = v;
}

/* de-interleave fields of an AoS image that was copied to the device in one piece */
struct AosPlan { int nf; int ni; unsigned off_d[NF_DOUBLE]; int pf_d[NF_DOUBLE]; unsigned off_i[NF_INT]; int pf_i[NF_INT]; };
struct DevPtrs { double* f[NF_DOUBLE]; int32_t* i[NF_INT]; };

__global__ void aos_unpack_kernel(int64_t n0, int64_t n, const char* __restrict__ aos, size_t stride, AosPlan plan, DevPtrs out)
{
  const int64_t i = (int64_t) blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= n) return;
  const char* rec = aos + (size_t) i*stride;
  for (int k = 0; k < plan.nf; k++) out.f[plan.pf_d[k]][n0 + i] = *reinterpret_cast<const double*>(rec + plan.off_d[k]);
  for (int k = 0; k < plan.ni; k++) out.i[plan.pf_i[k]][n0 + i] = *reinterpret_cast<const int32_t*>(rec + plan.off_i[k]);
}

__global__ void scatter_fields_kernel(int N, const int32_t* perm, DevPtrs P, DevPtrs H, unsigned long long dmask, unsigned imask)
{
  const int p = blockIdx.x*blockDim.x + threadIdx.x;
  if (p >= N) return;
  const int i = perm[p];
  for (int f = 0; f < NF_DOUBLE; f++) if ((dmask >> f) & 1ull) H.f[f][i] = P.f[f][p];
  for (int f = 0; f < NF_INT; f++) if ((imask >> f) & 1u) H.i[f][i] = P.i[f][p];
}

template <int NDIM>
__global__ void zero_acc_kernel(int lo, int N, DevPtrs P)
{
  const int i = lo + blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= N) return;
  if (!(((unsigned) P.i[I_FLAGS][i]) & GSPH_FLAG_ACTIVE)) return;
  P.i[I_LEVELNEIB][i] = 0;
  P.f[F_DIVV][i] = 0.0; P.f[F_DUDT][i] = 0.0; P.f[F_GPOT][i] = 0.0; P.f[F_GPOTHYDRO][i] = 0.0;
#pragma unroll
  for (int k = 0; k < NDIM; k++) { P.f[F_AX + k][i] = 0.0; P.f[F_ATX + k][i] = 0.0; }
}

/* with an active slice and inactive particles the devices' levelneib arrays are combined by a MAX reduction
   (the scatter half of the force pass lands on every slice): the active particles of the OTHER slices start
   from 0 here as they do on their own device */
__global__ void zero_levelneib_outside_kernel(int N, int lo, int hi, DevPtrs P)
{
  const int i = blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= N || (i >= lo && i < hi)) return;
  if (((unsigned) P.i[I_FLAGS][i]) & GSPH_FLAG_ACTIVE) P.i[I_LEVELNEIB][i] = 0;
}

/* register-resident FMA chains for the FP roofline denominators */
template <class T>
__global__ void fma_peak_kernel(T* out, int iters)
{
  T a0 = (T) threadIdx.x*(T) 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const T b = (T) 0.999999, c = (T) 1e-7;
  for (int i = 0; i < iters; i++) {
    a0 = a0*b + c; a1 = a1*b + c; a2 = a2*b + c; a3 = a3*b + c;
    a4 = a4*b + c; a5 = a5*b + c; a6 = a6*b + c; a7 = a7*b + c;
  }
  out[blockIdx.x*blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

DevPtrs devptrs(double* const* f, int32_t* const* i)
{
  DevPtrs d;
  for (int k = 0; k < NF_DOUBLE; k++) d.f[k] = f[k];
  for (int k = 0; k < NF_INT; k++) d.i[k] = i[k];
  return d;
}

int ensure_capacity(gsph_handle* h, int64_t N)
{
  if (N <= h->Ncap) return GSPH_OK;
  const int64_t cap = N + N/16 + 1024;
  /* nothing of the old state survives a regrow: invalidate it BEFORE the first free, so that a failed
     cudaMalloc half-way leaves a handle that refuses work instead of one that launches on freed arrays */
  h->Ncap = 0; h->integ.cap = 0;
  h->uploaded = h->have_tree = h->sorted_valid = h->density_done = false;
  h->integ.on = false; h->results_newer = false; h->v_pending = false; h->vcom_valid = false;
  h->present_d = h->hdef_d = h->pvalid_d = 0ull; h->present_i = h->hdef_i = h->pvalid_i = 0u;
  for (int f = 0; f < NF_DOUBLE; f++) {
    if (h->H[f]) cudaFree(h->H[f]); if (h->P[f]) cudaFree(h->P[f]);
    h->H[f] = h->P[f] = nullptr;
    CUDA_TRY(h, cudaMalloc(&h->H[f], cap*sizeof(double)));
    CUDA_TRY(h, cudaMalloc(&h->P[f], cap*sizeof(double)));
    /* defined contents from the start: a device that only ever writes its slice of a result array (multi-GPU)
       leaves zeros, not heap garbage, in the rest */
    CUDA_TRY(h, cudaMemsetAsync(h->P[f], 0, cap*sizeof(double), h->stream));
  }
  for (int f = 0; f < NF_INT; f++) {
    if (h->HI[f]) cudaFree(h->HI[f]); if (h->PI[f]) cudaFree(h->PI[f]);
    h->HI[f] = h->PI[f] = nullptr;
    CUDA_TRY(h, cudaMalloc(&h->HI[f], cap*sizeof(int32_t)));
    CUDA_TRY(h, cudaMalloc(&h->PI[f], cap*sizeof(int32_t)));
    CUDA_TRY(h, cudaMemsetAsync(h->PI[f], 0, cap*sizeof(int32_t), h->stream));
  }
  if (h->perm) cudaFree(h->perm);
  h->perm = nullptr;
  CUDA_TRY(h, cudaMalloc(&h->perm, cap*sizeof(int32_t)));
  if (h->geo4) cudaFree(h->geo4); if (h->geoh) cudaFree(h->geoh); if (h->frec) cudaFree(h->frec);
  if (h->hnew) cudaFree(h->hnew);
  h->geo4 = h->geoh = h->frec = h->hnew = nullptr;
  CUDA_TRY(h, cudaMalloc(&h->hnew, cap*sizeof(double)));
  CUDA_TRY(h, cudaMalloc(&h->geo4, cap*4*sizeof(double)));
  CUDA_TRY(h, cudaMalloc(&h->geoh, cap*4*sizeof(double)));
  CUDA_TRY(h, cudaMalloc(&h->frec, cap*16*sizeof(double)));
  h->Ncap = cap;
  h->have_tree = false;
  return GSPH_OK;
}

int ensure_pinned(gsph_handle* h, size_t bytes)
{
  if (bytes <= h->pinned_bytes) return GSPH_OK;
  if (h->pinned) cudaFreeHost(h->pinned);
  h->pinned = nullptr; h->pinned_bytes = 0;
  CUDA_TRY(h, cudaMallocHost(&h->pinned, bytes));
  h->pinned_bytes = bytes;
  return GSPH_OK;
}

void fill_devparams(gsph_handle* h)
{
  const gsph_config& c = h->cfg;
  DevParams& d = h->dp;
  memset(&d, 0, sizeof d);
  d.ndim = c.ndim; d.kernel = c.kernel; d.tabulated = c.tabulated_kernel;
  d.kernrange = c.kernel == GSPH_KERNEL_M4 ? 2.0 : 3.0;
  d.kernrangesqd = d.kernrange*d.kernrange;
  d.kernnorm = gsph_kernnorm(c.ndim, c.kernel);
  d.h_fac = c.h_fac; d.h_converge = c.h_converge;
  d.avisc = c.avisc; d.acond = c.acond; d.tdavisc = c.tdavisc;
  d.alpha_visc = c.alpha_visc; d.beta_visc = c.beta_visc; d.alpha_visc_min = c.alpha_visc_min;
  d.eos = c.eos; d.gamma_eos = c.gamma_eos; d.gammam1 = c.gamma_eos - 1.0;
  d.temp0 = c.temp0; d.mu_bar = c.mu_bar; d.rho_bary = c.rho_bary;
  d.any_periodic = 0;
  for (int k = 0; k < 3; k++) {
    d.boxmin[k] = c.boxmin[k]; d.boxmax[k] = c.boxmax[k];
    d.boxsize[k] = c.boxmax[k] - c.boxmin[k]; d.boxhalf[k] = 0.5*d.boxsize[k];
    d.periodic[k] = (k < c.ndim && c.boundary_lhs[k] == GSPH_BOUNDARY_PERIODIC) ? 1 : 0;
    d.any_periodic |= d.periodic[k];
  }
  d.star_zeta = c.conservative_sph_star_gravity != 0; d.nbody_softening = c.nbody_softening;
  d.stars = nullptr; d.nstar = 0;
  d.slice_lo = 0; d.slice_hi = 0x7fffffff;
  d.any_mirror = 0; d.nimg = 1;
  for (int k = 0; k < 3; k++) {
    d.mirror[k][0] = (k < c.ndim && c.boundary_lhs[k] == GSPH_BOUNDARY_MIRROR) ? 1 : 0;
    d.mirror[k][1] = (k < c.ndim && c.boundary_rhs[k] == GSPH_BOUNDARY_MIRROR) ? 1 : 0;
    d.any_mirror |= d.mirror[k][0] | d.mirror[k][1];
    d.nimg *= 1 + d.mirror[k][0] + d.mirror[k][1];
  }
  d.hydro_forces = c.hydro_forces; d.self_gravity = c.self_gravity; d.multipole = c.multipole;
  d.gravity_mac = c.gravity_mac; d.thetamaxsqd = c.thetamaxsqd; d.invthetamaxsqd = 1.0/c.thetamaxsqd;
  d.macerror = c.macerror; d.Nleafmax = c.Nleafmax; d.create_sinks = c.create_sinks;
}

int validate_config(const gsph_config* c, std::string& err)
{
  if (c->abi_version != GSPH_ABI_VERSION) { err = "gsph_config.abi_version mismatch"; return GSPH_EINVAL; }
  if (c->ndim < 1 || c->ndim > 3) { err = "ndim must be 1, 2 or 3"; return GSPH_EINVAL; }
  if (c->kernel < GSPH_KERNEL_M4 || c->kernel > GSPH_KERNEL_GAUSSIAN) { err = "kernel must be m4, quintic or gaussian"; return GSPH_EINVAL; }
  if (c->kernel == GSPH_KERNEL_GAUSSIAN && c->self_gravity) { err = "the gaussian kernel has no gravity softening (wgrav = wpot = 0, SmoothingKernel.h:497-521): not with self_gravity"; return GSPH_EINVAL; }
  if (!(c->h_fac > 0) || !(c->h_converge > 0)) { err = "h_fac and h_converge must be positive"; return GSPH_EINVAL; }
  if (c->Nleafmax < 1) { err = "Nleafmax must be >= 1"; return GSPH_EINVAL; }
  if (c->eos < GSPH_EOS_ADIABATIC || c->eos > GSPH_EOS_BAROTROPIC2) { err = "gas_eos is not on the accelerated path"; return GSPH_EINVAL; }
  if (c->tdavisc < GSPH_TDAVISC_NONE || c->tdavisc > GSPH_TDAVISC_CD2010) { err = "unknown time_dependent_avisc"; return GSPH_EINVAL; }
  if (c->tdavisc == GSPH_TDAVISC_CD2010)
    for (int k = 0; k < c->ndim; k++)
      if (c->boundary_lhs[k] == GSPH_BOUNDARY_MIRROR || c->boundary_rhs[k] == GSPH_BOUNDARY_MIRROR) {
        err = "time_dependent_avisc = cd2010 with mirror boundaries is not on the accelerated path"; return GSPH_EINVAL;
      }
  for (int k = 0; k < c->ndim; k++) {
    if ((c->boundary_lhs[k] == GSPH_BOUNDARY_MIRROR || c->boundary_rhs[k] == GSPH_BOUNDARY_MIRROR)) {
      if (c->self_gravity) { err = "self-gravity with mirror boundaries is not supported (the reference asserts it, Tree.cpp:636)"; return GSPH_EINVAL; }
      if (!(c->boxmax[k] > c->boxmin[k])) { err = "mirror dimension needs boxmax > boxmin"; return GSPH_EINVAL; }
    }
    if ((c->boundary_lhs[k] == GSPH_BOUNDARY_PERIODIC) != (c->boundary_rhs[k] == GSPH_BOUNDARY_PERIODIC)) {
      err = "periodic boundaries must be set on both sides of a dimension"; return GSPH_EINVAL;
    }
    if (c->boundary_lhs[k] == GSPH_BOUNDARY_PERIODIC && !(c->boxmax[k] > c->boxmin[k])) {
      err = "periodic dimension needs boxmax > boxmin"; return GSPH_EINVAL;
    }
  }
  if (c->self_gravity) {
    if (c->gravity_mac < GSPH_MAC_GEOMETRIC || c->gravity_mac > GSPH_MAC_GADGET2) { err = "unknown gravity_mac"; return GSPH_EINVAL; }
    if (c->gravity_mac != GSPH_MAC_GEOMETRIC && !(c->macerror > 0.0)) { err = "gravity_mac = eigenmac / gadget2 needs macerror > 0"; return GSPH_EINVAL; }
    if (!(c->thetamaxsqd > 0)) { err = "thetamaxsqd must be positive"; return GSPH_EINVAL; }
    for (int k = 0; k < c->ndim; k++)
      if (c->boundary_lhs[k] == GSPH_BOUNDARY_PERIODIC) { err = "periodic (Ewald) gravity is not on the accelerated path"; return GSPH_EINVAL; }
  }
  return GSPH_OK;
}

struct Timer {
  gsph_handle* h; float* out;
  Timer(gsph_handle* h_, float* o) : h(h_), out(o) { cudaEventRecord(h->ev0, h->stream); }
  void stop() { cudaEventRecord(h->ev1, h->stream); cudaEventSynchronize(h->ev1); cudaEventElapsedTime(out, h->ev0, h->ev1); }
};

int reset_status(gsph_handle* h)
{
  CUDA_TRY(h, cudaMemsetAsync(h->status, 0, sizeof(DevStatus), h->stream));
  return GSPH_OK;
}

int fetch_status(gsph_handle* h)
{
  CUDA_TRY(h, cudaMemcpyAsync(h->status_host, h->status, sizeof(DevStatus), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return GSPH_OK;
}

}  // namespace

/*------------------------------------------------------------------------------------------------*/
int gsph_ensure_cand(gsph_handle* h, int warps, size_t cap_per_warp)
{
  if (cap_per_warp == 0) cap_per_warp = h->cand_cap_per_warp;
  if (cap_per_warp == 0) {
    const char* e = getenv("GSPH_CAND_CAP");
    cap_per_warp = e ? (size_t) atol(e) : 8192;
  }
  if (h->cand && warps <= h->cand_warps && cap_per_warp <= h->cand_cap_per_warp) return GSPH_OK;
  if (h->cand) cudaFree(h->cand);
  h->cand = nullptr;
  if (warps < h->cand_warps) warps = h->cand_warps;
  CUDA_TRY(h, cudaMalloc(&h->cand, (size_t) warps*cap_per_warp*sizeof(int32_t)));
  h->cand_warps = warps; h->cand_cap_per_warp = cap_per_warp;
  return GSPH_OK;
}

/* cohort range [c_lo, c_hi) of `level` that can hold targets of this device: all of them, or (multi-GPU) those
   meeting the tree-order slice -- from the data-independent topology of the balanced tree */
static void cohort_range(const gsph_handle* h, int level, int& c_lo, int& c_hi)
{
  const int ncohort = 1 << level;
  c_lo = 0; c_hi = ncohort;
  if (!h->slice_mode) return;
  auto range = [&](int k, int64_t& a, int64_t& b) {
    a = 0; b = h->N;
    for (int bit = level - 1; bit >= 0; bit--) { const int64_t left = (b - a)/2; if ((k >> bit) & 1) a += left; else b = a + left; }
  };
  int lo = 0, hi = ncohort;
  while (lo < hi) { const int mid = (lo + hi)/2; int64_t a, b; range(mid, a, b); if (b <= h->dp.slice_lo) lo = mid + 1; else hi = mid; }
  c_lo = lo; hi = ncohort;
  while (lo < hi) { const int mid = (lo + hi)/2; int64_t a, b; range(mid, a, b); if (a < h->dp.slice_hi) lo = mid + 1; else hi = mid; }
  c_hi = lo;
}

/* Buffers of the pre-walk kernels for the cohorts of this device.  GSPH_PREWALK=0 disables them (the main kernels
   walk themselves), GSPH_PREWALK_CAP = candidate ids per cohort, GSPH_PREWALK_MB = memory budget.            */
int gsph_prewalk_prepare(gsph_handle* h, int cohort_level)
{
  h->pre_n = 0;
  const char* e0 = getenv("GSPH_PREWALK");
  if (e0 && atoi(e0) == 0) return GSPH_OK;
  int c_lo, c_hi;
  cohort_range(h, cohort_level, c_lo, c_hi);
  const int n = c_hi - c_lo;
  if (n <= 0) return GSPH_OK;
  const char* e1 = getenv("GSPH_PREWALK_CAP"); const char* e2 = getenv("GSPH_PREWALK_MB");
  size_t cap = e1 ? (size_t) std::max(64, atoi(e1)) : (h->cfg.kernel == GSPH_KERNEL_M4 ? 2048 : 4096);
  const size_t budget = (e2 ? (size_t) atol(e2) : (size_t) 24576) << 20;
  while (cap > 512 && (size_t) n*cap*sizeof(int32_t) > budget) cap /= 2;
  if ((size_t) n*cap*sizeof(int32_t) > budget) return GSPH_OK;           /* too many cohorts: in-kernel walks */
  const size_t need = (size_t) n*cap*sizeof(int32_t);
  if (need > h->pre_bytes) {
    if (h->pre_cand) cudaFree(h->pre_cand);
    h->pre_cand = nullptr; h->pre_bytes = 0;
    if (cudaMalloc(&h->pre_cand, need) != cudaSuccess) { cudaGetLastError(); return GSPH_OK; }   /* no room: in-kernel walks */
    h->pre_bytes = need;
  }
  if ((size_t) n > h->pre_count_cap) {
    if (h->pre_count) cudaFree(h->pre_count);
    h->pre_count = nullptr; h->pre_count_cap = 0;
    CUDA_TRY(h, cudaMalloc(&h->pre_count, (size_t) n*sizeof(int)));
    h->pre_count_cap = (size_t) n;
  }
  h->pre_c0 = c_lo; h->pre_n = n; h->pre_cap = (int) cap;
  return GSPH_OK;
}

int gsph_ensure_nlist(gsph_handle* h, int warps, size_t cap)
{
  if (cap == 0) cap = h->nlist_cap;
  if (cap == 0) {
    const char* e = getenv("GSPH_NLIST_CAP");
    cap = e ? (size_t) atol(e) : (h->cfg.kernel == GSPH_KERNEL_M4 ? 256 : 768);
  }
  if (h->nlist && warps <= h->nlist_warps && cap <= h->nlist_cap) return GSPH_OK;
  if (h->nlist) cudaFree(h->nlist);
  h->nlist = nullptr;
  if (warps < h->nlist_warps) warps = h->nlist_warps;
  CUDA_TRY(h, cudaMalloc(&h->nlist, (size_t) warps*cap*32*sizeof(int32_t)));
  h->nlist_warps = warps; h->nlist_cap = cap;
  return GSPH_OK;
}

/*------------------------------------------------------------------------------------------------
 * device-resident main loop: streaming kernels over the host-order (staging) arrays.  The
 * arithmetic follows the reference's expressions operation by operation (no contraction), so a
 * step reproduces the CPU integrator's rounding.
 *-----------------------------------------------------------------------------------------------*/
struct IntegPtrs { double* r0[3]; double* v0[3]; double* a0[3]; double* u0; double* dudt0; };

IntegPtrs integ_ptrs(gsph_handle* h)
{
  IntegPtrs S;
  for (int k = 0; k < 3; k++) { S.r0[k] = h->integ.s0[k]; S.v0[k] = h->integ.s0[3 + k]; S.a0[k] = h->integ.s0[6 + k]; }
  S.u0 = h->integ.s0[9]; S.dudt0 = h->integ.s0[10];
  return S;
}

/* SphLeapfrogKDK::AdvanceParticles (SphLeapfrogKDK.cpp:76-125) + TimeIntegration::CheckBoundaries
 * (Integration.cpp:11-75).  dt = t - tlast, timestep = the global step (they differ by rounding). */
struct BlkPtrs;
template <int NDIM>
__global__ void integ_advance_kernel(int N, DevPtrs H, IntegPtrs S, DevParams dp, double dt, double timestep,
                                     int tdav, int energy, int block, double t, int n, const int32_t* __restrict__ nstep,
                                     const int32_t* __restrict__ nlast, const double* __restrict__ tlast)
{
  const int i = blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= N) return;
  unsigned fl = (unsigned) H.i[I_FLAGS][i];
  if (fl & GSPH_FLAG_DEAD) return;
  bool act = true;                                          /* one global step: dn == nstep for every particle */
  if (block) {                                              /* time since the start of THIS particle's step */
    dt = __dadd_rn(t, -tlast[i]);
    act = (n - nlast[i]) == nstep[i];
  }
#pragma unroll
  for (int k = 0; k < NDIM; k++) {
    const double r0 = S.r0[k][i], v0 = S.v0[k][i], a0 = S.a0[k][i];
    /* r0 + v0*dt + 0.5*a0*dt*dt */
    H.f[F_RX + k][i] = __dadd_rn(__dadd_rn(r0, __dmul_rn(v0, dt)), __dmul_rn(__dmul_rn(__dmul_rn(0.5, a0), dt), dt));
    H.f[F_VX + k][i] = __dadd_rn(v0, __dmul_rn(a0, dt));
  }
  if (tdav) H.f[F_ALPHA][i] = __dadd_rn(H.f[F_ALPHA][i], __dmul_rn(H.f[F_DALPHADT][i], timestep));
  if (energy) H.f[F_U][i] = __dadd_rn(S.u0[i], __dmul_rn(S.dudt0[i], dt));
  H.i[I_FLAGS][i] = (int32_t)(act ? (fl | GSPH_FLAG_ACTIVE) : (fl & ~(unsigned) GSPH_FLAG_ACTIVE));
  /* boundaries */
#pragma unroll
  for (int k = 0; k < NDIM; k++) {
    if (!dp.periodic[k] && !dp.mirror[k][0] && !dp.mirror[k][1]) continue;
    if (H.f[F_RX + k][i] < dp.boxmin[k]) {
      if (dp.periodic[k]) {
        H.f[F_RX + k][i] = __dadd_rn(H.f[F_RX + k][i], dp.boxsize[k]);
        S.r0[k][i] = __dadd_rn(S.r0[k][i], dp.boxsize[k]);
      }
      if (dp.mirror[k][0]) {
        H.f[F_RX + k][i] = __dadd_rn(__dmul_rn(2.0, dp.boxmin[k]), -H.f[F_RX + k][i]);
        S.r0[k][i] = __dadd_rn(__dmul_rn(2.0, dp.boxmin[k]), -S.r0[k][i]);
        H.f[F_VX + k][i] = -H.f[F_VX + k][i]; S.v0[k][i] = -S.v0[k][i];
        H.f[F_AX + k][i] = -H.f[F_AX + k][i]; S.a0[k][i] = -S.a0[k][i];
      }
    }
    if (H.f[F_RX + k][i] > dp.boxmax[k]) {
      if (dp.periodic[k]) {
        H.f[F_RX + k][i] = __dadd_rn(H.f[F_RX + k][i], -dp.boxsize[k]);
        S.r0[k][i] = __dadd_rn(S.r0[k][i], -dp.boxsize[k]);
      }
      if (dp.mirror[k][1]) {
        H.f[F_RX + k][i] = __dadd_rn(__dmul_rn(2.0, dp.boxmax[k]), -H.f[F_RX + k][i]);
        S.r0[k][i] = __dadd_rn(__dmul_rn(2.0, dp.boxmax[k]), -S.r0[k][i]);
        H.f[F_VX + k][i] = -H.f[F_VX + k][i]; S.v0[k][i] = -S.v0[k][i];
        H.f[F_AX + k][i] = -H.f[F_AX + k][i]; S.a0[k][i] = -S.a0[k][i];
      }
    }
  }
}

/* SphIntegration::Timestep of one particle (SphIntegration.cpp:76-124), host-order state */
template <int NDIM>
__device__ __forceinline__ double particle_timestep(int i, const DevPtrs& H, const DevParams& dp, const gsph_integrator_config& ic, bool& bad)
{
  double ts;
  const double hh = H.f[F_H][i], dv = fabs(H.f[F_DIVV][i]);
  if (dp.hydro_forces == 1) ts = __ddiv_rn(__dmul_rn(ic.courant_mult, hh), __dadd_rn(__dadd_rn(H.f[F_SOUND][i], __dmul_rn(hh, dv)), 1.0e-50));
  else ts = __ddiv_rn(__dmul_rn(ic.courant_mult, hh), __dadd_rn(__dmul_rn(hh, dv), 1.0e-50));
  double a2 = 0.0;
#pragma unroll
  for (int k = 0; k < NDIM; k++) { const double a = H.f[F_AX + k][i]; a2 = k == 0 ? __dmul_rn(a, a) : __dadd_rn(a2, __dmul_rn(a, a)); }
  const double amag = __dsqrt_rn(a2);
  ts = fmin(ts, __dmul_rn(ic.accel_mult, __dsqrt_rn(__ddiv_rn(hh, __dadd_rn(amag, 1.0e-50)))));
  if (ic.energy_integration && H.i[I_PTYPE][i] == 0)
    ts = fmin(ts, __dmul_rn(ic.energy_mult, __ddiv_rn(H.f[F_U][i], __dadd_rn(fabs(H.f[F_DUDT][i]), 1.0e-20))));
  bad = !(ts <= 1.0e20) || ts <= 0.0;
  return ts;
}

/* every particle's time step and the global minimum (Simulation::ComputeGlobalTimestep, Simulation.cpp:1690-1730).
 * Positive doubles order like their bit patterns, so the minimum is an integer atomicMin.                */
template <int NDIM>
__global__ void integ_timestep_kernel(int N, DevPtrs H, DevParams dp, gsph_integrator_config ic, unsigned long long* red)
{
  const int i = blockIdx.x*blockDim.x + threadIdx.x;
  double ts = 9.9e50;
  bool bad = false;
  if (i < N) ts = particle_timestep<NDIM>(i, H, dp, ic, bad);
  unsigned long long bits = bad ? 0x7ff0000000000000ull : (unsigned long long) __double_as_longlong(ts);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const unsigned long long other = __shfl_xor_sync(0xffffffffu, bits, o); bits = other < bits ? other : bits; }
  const unsigned nbad = __ballot_sync(0xffffffffu, bad);
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&red[0], bits);
    if (nbad) atomicAdd(&red[1], (unsigned long long) __popc(nbad));
  }
}

/*------------------------------------------------------------------------------------------------
 * Hierarchical block time steps (gsph_integrator_config::Nlevels > 1).  Per-particle integer step data live in
 * host order next to r0, v0, ...; the level logic follows Simulation::ComputeBlockTimesteps
 * (Simulation.cpp:1764-2199) statement by statement, incl. (int) pow(2, negative) = 0 for a particle that has just
 * moved below the deepest level (its dt_next is 0 for that step, as in the reference).
 *-----------------------------------------------------------------------------------------------*/
struct BlkPtrs { int32_t* nstep; int32_t* nlast; double* dt; double* dt_next; double* tlast; };
#define GSPH_INVLOGETWO 1.44269504088896            /* src/Headers/Constants.h:81 */

__host__ __device__ __forceinline__ int pow2int(int e) { return e >= 0 ? (e < 31 ? (1 << e) : 0x40000000) : 0; }

/* ComputeTimestepLevel, InlineFuncs.h:550-558 */
__device__ __forceinline__ int timestep_level(double dt, double dt_max)
{
  const int l = (int)(GSPH_INVLOGETWO*log(__ddiv_rn(dt_max, dt))) + 1;
  return l > 0 ? l : 0;
}

/* SphLeapfrogKDK::CheckTimesteps (SphLeapfrogKDK.cpp:284-339) after "unset active for every particle"
 * (SphSimulation.cpp:735-738): red[1] += particles woken up */
__global__ void blk_check_kernel(int N, DevPtrs H, BlkPtrs B, int n, int level_step, int ldm, unsigned long long* red)
{
  const int i = blockIdx.x*blockDim.x + threadIdx.x;
  bool woke = false;
  if (i < N) {
    unsigned fl = ((unsigned) H.i[I_FLAGS][i]) & ~(unsigned) GSPH_FLAG_ACTIVE;
    if (!(fl & GSPH_FLAG_DEAD)) {
      const int dn = n - B.nlast[i];
      const int level = H.i[I_LEVEL][i], levelneib = H.i[I_LEVELNEIB][i];
      if (dn != B.nstep[i] && levelneib - level > ldm) {
        const int level_new = levelneib - ldm;
        const int nnew = pow2int(level_step - level_new);
        if (nnew > 0 && dn % nnew == 0) {
          if (dn > 0) B.nstep[i] = dn;
          H.i[I_LEVEL][i] = level_new;
          fl |= GSPH_FLAG_ACTIVE;
          woke = true;
        }
      }
    }
    H.i[I_FLAGS][i] = (int32_t) fl;
  }
  const unsigned w = __ballot_sync(0xffffffffu, woke);
  if ((threadIdx.x & 31) == 0 && w) atomicAdd(&red[1], (unsigned long long) __popc(w));
}

/* mode 0  resynchronisation, first half: dt_next = Timestep(part), global minimum           (Simulation.cpp:1808-1818)
 * mode 1  resynchronisation, second half: levels from dt_next                                   (:1905-1917)
 * mode 2  regular step: particles at the end of their step choose their new level; red[2] = max level (:1947-2009)
 * mode 3  integer time rescaled by nfactor (op 1 multiply, 2 divide), then nstep of the particles that start a
 *         step now                                                                              (:2113-2155)  */
template <int NDIM>
__global__ void blk_timestep_kernel(int N, DevPtrs H, BlkPtrs B, DevParams dp, gsph_integrator_config ic, int mode, int n,
                                    int level_step, int level_max, double dt_max, double timestep, int op, int nfactor,
                                    unsigned long long* red, int single_level)
{
  /* single_level >= 0: sph_single_timestep = 1, the level every gas particle that starts a step now is put on
     (level_max_hydro; Simulation.cpp:1889-1902, 2092-2099) */
  const int i = blockIdx.x*blockDim.x + threadIdx.x;
  const int ldm = ic.level_diff_max > 0 ? ic.level_diff_max : 1;
  bool bad = false;
  unsigned long long bits = 0x7ff0000000000000ull;
  int lev_here = 0;
  const bool alive = i < N && !(((unsigned) H.i[I_FLAGS][i]) & GSPH_FLAG_DEAD);
  if (alive) {
    if (mode == 0) {
      const double ts = particle_timestep<NDIM>(i, H, dp, ic, bad);
      B.dt_next[i] = ts;
      if (!bad) bits = (unsigned long long) __double_as_longlong(ts);
    }
    else if (mode == 1) {
      int level = timestep_level(B.dt_next[i], dt_max);
      if (level > level_max) level = level_max;
      if (single_level >= 0) level = single_level;
      H.i[I_LEVEL][i] = level; H.i[I_LEVELNEIB][i] = level;
      const int ns = pow2int(level_step - level);
      B.nstep[i] = ns; B.nlast[i] = n;
      B.dt_next[i] = __dmul_rn((double) ns, timestep);
      H.i[I_FLAGS][i] = (int32_t)(((unsigned) H.i[I_FLAGS][i]) | GSPH_FLAG_END_TIMESTEP);
    }
    else if (mode == 2) {
      const int nstep = B.nstep[i];
      int plevel = H.i[I_LEVEL][i];
      if (n - B.nlast[i] == nstep) {
        const double ts = particle_timestep<NDIM>(i, H, dp, ic, bad);
        int level = timestep_level(ts, dt_max);
        const int lnb = H.i[I_LEVELNEIB][i] - ldm;
        if (lnb > level) level = lnb;
        if (nstep != pow2int(level_step - plevel)) {          /* step cut short by the neighbour-level limiter */
          if (level > plevel) plevel = level;
          H.i[I_LEVELNEIB][i] = plevel;
        }
        else {                                               /* natural end of the step */
          const int last_level = plevel;
          if (level < last_level && last_level > 1 && n % (2*nstep) == 0) plevel = last_level - 1;
          else if (level > last_level) plevel = level;
          H.i[I_LEVELNEIB][i] = level;
        }
        H.i[I_LEVEL][i] = plevel;
        B.nlast[i] = n;
        const int ns = pow2int(level_step - plevel);
        B.nstep[i] = ns;
        B.dt_next[i] = __dmul_rn((double) ns, timestep);
        H.i[I_FLAGS][i] = (int32_t)(((unsigned) H.i[I_FLAGS][i]) | GSPH_FLAG_END_TIMESTEP);
      }
      lev_here = plevel;
    }
    else {
      int ns = B.nstep[i], nl = B.nlast[i];
      if (op == 1) { ns *= nfactor; nl *= nfactor; }
      else if (op == 2) { ns /= nfactor; nl /= nfactor; }
      if (nl == n) {
        if (single_level >= 0) H.i[I_LEVEL][i] = single_level;
        ns = pow2int(level_step - H.i[I_LEVEL][i]);
      }
      B.nstep[i] = ns; B.nlast[i] = nl;
    }
  }
  if (mode == 0 || mode == 2) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, bits, o); bits = other < bits ? other : bits;
      lev_here = max(lev_here, __shfl_xor_sync(0xffffffffu, lev_here, o));
    }
    const unsigned nbad = __ballot_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0) {
      if (mode == 0) atomicMin(&red[0], bits);
      if (nbad) atomicAdd(&red[1], (unsigned long long) __popc(nbad));
      if (mode == 2) atomicMax(&red[2], (unsigned long long) lev_here);
    }
  }
}

/* first half of PostInitialConditionsSetup's tail: r0 = r, v0 = v, a0 = a (SphSimulation.cpp:480-486) */
template <int NDIM>
__global__ void integ_init_kernel(int N, DevPtrs H, IntegPtrs S)
{
  const int i = blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= N) return;
#pragma unroll
  for (int k = 0; k < NDIM; k++) { S.r0[k][i] = H.f[F_RX + k][i]; S.v0[k][i] = H.f[F_VX + k][i]; S.a0[k][i] = H.f[F_AX + k][i]; }
  S.u0[i] = H.f[F_U][i]; S.dudt0[i] = 0.0;
}

/* SphLeapfrogKDK::EndTimestep (SphLeapfrogKDK.cpp:226-270); dtp = part.dt, the step just taken */
template <int NDIM>
__global__ void integ_end_kernel(int N, DevPtrs H, IntegPtrs S, double dtp, int energy, int block, int n, double t,
                                 int32_t* __restrict__ nlast, double* __restrict__ pdt, double* __restrict__ dt_next,
                                 double* __restrict__ tlast)
{
  const int i = blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= N) return;
  const unsigned fl = (unsigned) H.i[I_FLAGS][i];
  if (fl & GSPH_FLAG_DEAD) return;
  if (block) {                                              /* only the particles whose step ends now */
    if (!(fl & GSPH_FLAG_END_TIMESTEP)) return;
    dtp = pdt[i];                                           /* part.dt: the step just taken */
    nlast[i] = n; tlast[i] = t; pdt[i] = dt_next[i]; dt_next[i] = 0.0;
  }
  const double hdt = __dmul_rn(0.5, dtp);
#pragma unroll
  for (int k = 0; k < NDIM; k++) {
    const double a = H.f[F_AX + k][i];
    const double v = __dadd_rn(H.f[F_VX + k][i], __dmul_rn(hdt, __dadd_rn(a, -S.a0[k][i])));
    H.f[F_VX + k][i] = v;
    S.r0[k][i] = H.f[F_RX + k][i]; S.v0[k][i] = v; S.a0[k][i] = a;
  }
  if (energy) {
    const double dudt = H.f[F_DUDT][i];
    double u = __dadd_rn(H.f[F_U][i], __dmul_rn(__dmul_rn(0.5, __dadd_rn(dudt, -S.dudt0[i])), dtp));
    if (u <= 0.0) u = __dadd_rn(S.u0[i], __dmul_rn(S.dudt0[i], dtp));
    H.f[F_U][i] = u; S.u0[i] = u; S.dudt0[i] = dudt;
  }
  H.i[I_FLAGS][i] = (int32_t)(fl & ~(unsigned)(GSPH_FLAG_ACTIVE | GSPH_FLAG_END_TIMESTEP));
}

int gsph_check_status(gsph_handle* h, const char* where)
{
  int rc = fetch_status(h);
  if (rc) return rc;
  const DevStatus& s = *h->status_host;
  if (s.err_stack_overflow) { h->err = std::string(where) + ": tree-walk stack overflow"; return GSPH_ENOMEM; }
  if (s.err_converge) { h->err = "Problem with convergence of h-rho iteration"; return GSPH_ECONVERGE; }
  if (s.err_hsearch) { h->err = std::string(where) + ": search radius grew more than 60 times"; return GSPH_ECONVERGE; }
  return GSPH_OK;
}

extern "C" {

int gsph_default_config(gsph_config* c)
{
  if (!c) return GSPH_EINVAL;
  memset(c, 0, sizeof *c);
  c->abi_version = GSPH_ABI_VERSION;
  c->ndim = 3; c->kernel = GSPH_KERNEL_M4; c->tabulated_kernel = 1;
  c->h_fac = 1.2; c->h_converge = 0.01;
  c->avisc = GSPH_AVISC_MON97; c->acond = GSPH_ACOND_NONE; c->tdavisc = GSPH_TDAVISC_NONE;
  c->conservative_sph_star_gravity = 1; c->nbody_softening = 1;
  c->alpha_visc = 1.0; c->beta_visc = 2.0; c->alpha_visc_min = 0.1;
  c->eos = GSPH_EOS_ADIABATIC; c->gamma_eos = 1.6666666666666;
  c->temp0 = 1.0; c->mu_bar = 1.0; c->rho_bary = 1.0e-14;
  for (int k = 0; k < 3; k++) { c->boxmin[k] = 0.0; c->boxmax[k] = 0.0; c->boundary_lhs[k] = c->boundary_rhs[k] = GSPH_BOUNDARY_OPEN; }
  c->hydro_forces = 1; c->self_gravity = 0;
  c->thetamaxsqd = 0.1; c->gravity_mac = GSPH_MAC_GEOMETRIC; c->macerror = 1.0e-4;
  c->multipole = GSPH_MULTIPOLE_QUADRUPOLE; c->Nleafmax = 6;
  for (int i = 0; i < GSPH_NTYPES; i++) { c->hmask[i] = 0; c->hydromask[i] = 0; }
  c->hmask[GSPH_TYPE_GAS] = 1u << GSPH_TYPE_GAS; c->hydromask[GSPH_TYPE_GAS] = 1u << GSPH_TYPE_GAS;
  c->gravmask = 1u << GSPH_TYPE_GAS; c->hydro_forces_type = 1u << GSPH_TYPE_GAS;
  c->self_gravity_type = 1u << GSPH_TYPE_GAS;
  c->create_sinks = 0; c->device = 0;
  return GSPH_OK;
}

int gsph_create(const gsph_config* cfg, gsph_handle** out)
{
  if (!cfg || !out) { g_create_error = "gsph_create: NULL argument"; return GSPH_EINVAL; }
  *out = nullptr;
  int rc = validate_config(cfg, g_create_error);
  if (rc) return rc;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_error = "gsph_create: no CUDA device (libgsph has no CPU fallback)";
    return GSPH_ENODEVICE;
  }
  if (cfg->device < 0 || cfg->device >= ndev) { g_create_error = "gsph_create: bad device ordinal"; return GSPH_EINVAL; }
  gsph_handle* h = new gsph_handle();
  h->cfg = *cfg; h->device = cfg->device;
  memset(&h->stats, 0, sizeof h->stats);
  auto fail = [&](const char* what, cudaError_t ce) {
    g_create_error = std::string(what) + ": " + cudaGetErrorString(ce);
    gsph_destroy(h);
    return GSPH_ECUDA;
  };
  if ((e = cudaSetDevice(h->device)) != cudaSuccess) return fail("cudaSetDevice", e);
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, h->device)) != cudaSuccess) return fail("cudaGetDeviceProperties", e);
  h->sm_count = prop.multiProcessorCount;
  if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", e);
  if ((e = cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", e);
  if ((e = cudaEventCreateWithFlags(&h->ev_geom, cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e);
  if ((e = cudaEventCreateWithFlags(&h->ev_all, cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e);
  if ((e = cudaEventCreateWithFlags(&h->ev_early, cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e);
  if ((e = cudaEventCreateWithFlags(&h->ev_compute, cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e);
  for (int k = 0; k < 3; k++)
    if ((e = cudaEventCreateWithFlags(&h->ev_axis[k], cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e);
  if ((e = cudaEventCreate(&h->ev0)) != cudaSuccess) return fail("cudaEventCreate", e);
  if ((e = cudaEventCreate(&h->ev1)) != cudaSuccess) return fail("cudaEventCreate", e);
  for (int k = 0; k < 2; k++) if ((e = cudaEventCreate(&h->evt[k])) != cudaSuccess) return fail("cudaEventCreate", e);
  for (int k = 0; k < 6; k++) for (int j = 0; j < 2; j++)
    if ((e = cudaEventCreate(&h->evk[k][j])) != cudaSuccess) return fail("cudaEventCreate", e);
  fill_devparams(h);
  std::vector<double> tab(6*GSPH_TABLE_RES);
  gsph_make_tables(cfg->ndim, cfg->kernel, tab.data());
  if ((e = cudaMalloc(&h->tables, tab.size()*sizeof(double))) != cudaSuccess) return fail("cudaMalloc", e);
  if ((e = cudaMemcpy(h->tables, tab.data(), tab.size()*sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess) return fail("cudaMemcpy", e);
  if ((e = cudaMalloc(&h->status, sizeof(DevStatus))) != cudaSuccess) return fail("cudaMalloc", e);
  if ((e = cudaMallocHost(&h->status_host, sizeof(DevStatus))) != cudaSuccess) return fail("cudaMallocHost", e);
  if ((e = cudaMalloc(&h->work_counter, 4*sizeof(int))) != cudaSuccess) return fail("cudaMalloc", e);
  *out = h;
  return GSPH_OK;
}

void gsph_destroy(gsph_handle* h)
{
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
  auto fre = [](void* p) { if (p) cudaFree(p); };
  for (int f = 0; f < NF_DOUBLE; f++) { fre(h->H[f]); fre(h->P[f]); }
  for (int f = 0; f < NF_INT; f++) { fre(h->HI[f]); fre(h->PI[f]); }
  for (int k = 0; k < 11; k++) fre(h->integ.s0[k]);
  fre(h->integ.nstep); fre(h->integ.nlast); fre(h->integ.pdt); fre(h->integ.dt_next); fre(h->integ.ptlast);
  fre(h->integ.red); fre(h->perm); fre(h->tables); fre(h->scratch); fre(h->cand); fre(h->gravlist); fre(h->status); fre(h->work_counter); fre(h->done); fre(h->needed); fre(h->pre_cand); fre(h->pre_count); fre(h->geo4); fre(h->geoh); fre(h->frec); fre(h->nlist); fre(h->hnew); fre(h->stars); fre(h->gstack);
  TreeDev& T = h->tree;
  fre(T.first); fre(T.count); fre(T.nactive); fre(T.m); fre(T.rmax); fre(T.hmax); fre(T.cdistsqd);
  for (int k = 0; k < 3; k++) { fre(T.bbmin[k]); fre(T.bbmax[k]); fre(T.hbmin[k]); fre(T.hbmax[k]); fre(T.rcell[k]); fre(T.com[k]); fre(T.vcom[k]); }
  for (int k = 0; k < 5; k++) fre(T.q[k]);
  if (h->status_host) cudaFreeHost(h->status_host);
  if (h->pinned) cudaFreeHost(h->pinned);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  for (int k = 0; k < 2; k++) if (h->evt[k]) cudaEventDestroy(h->evt[k]);
  for (int k = 0; k < 6; k++) for (int j = 0; j < 2; j++) if (h->evk[k][j]) cudaEventDestroy(h->evk[k][j]);
  if (h->ev_geom) cudaEventDestroy(h->ev_geom);
  if (h->ev_all) cudaEventDestroy(h->ev_all);
  if (h->ev_early) cudaEventDestroy(h->ev_early);
  if (h->ev_compute) cudaEventDestroy(h->ev_compute);
  for (int k = 0; k < 3; k++) if (h->ev_axis[k]) cudaEventDestroy(h->ev_axis[k]);
  fre(h->tree.nodeb); fre(h->tree.nodeh); fre(h->tree.gnode); fre(h->tree.mnode);
  for (int k = 0; k < 5; k++) if (h->ev_grav[k]) cudaEventDestroy(h->ev_grav[k]);
  if (h->aux_stream) cudaStreamDestroy(h->aux_stream);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

const char* gsph_last_error(const gsph_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

/*------------------------------------------------------------------------------------------------
 * upload: contiguous (stride == element size) fields go straight to the device; strided fields
 * that all live inside one AoS record (the reference's GradhSphParticle array) are moved as ONE
 * bulk copy per chunk and de-interleaved on the device.
 *-----------------------------------------------------------------------------------------------*/
static int upload_impl(gsph_handle* h, const gsph_particle_view* view, int64_t N, int64_t r0, int64_t r1)
{
  if (!h) return GSPH_EINVAL;
  if (!view || N <= 0 || N > 2000000000ll) { h->err = "gsph_upload: bad arguments"; return GSPH_EINVAL; }
  if (r0 < 0 || r1 < r0 || r1 > N) { h->err = "gsph_upload_range: bad range"; return GSPH_EINVAL; }
  if (h->dp.any_mirror && N > (int64_t) 1 << 26) { h->err = "mirror boundaries support at most 2^26 particles (image codes live in the upper id bits)"; return GSPH_EINVAL; }
  const bool whole = (r0 == 0 && r1 == N);
  const int64_t nr = r1 - r0;
  cudaSetDevice(h->device);
  Timer tm(h, &h->stats.ms_upload);
  int rc = GSPH_OK;
  if (h->v_pending && N == h->N) rc = gsph_ensure_velocities(h);   /* before the staging copy of v is overwritten */
  if (rc) return rc;
  h->v_pending = false;
  rc = ensure_capacity(h, N);
  if (rc) return rc;
  const int nd = h->cfg.ndim;
  if (!vfield(view, VOFF(r[0])).base || !vfield(view, VOFF(m)).base || !vfield(view, VOFF(h)).base) {
    h->err = "gsph_upload: r, m and h are required"; return GSPH_EINVAL;
  }
  /* classify */
  AosPlan plan; plan.nf = plan.ni = 0;
  const char* lo = nullptr; const char* hi = nullptr; size_t stride = 0; bool bulk_ok = true;
  auto consider = [&](const gsph_field& f, size_t esize) {
    if (!f.base || f.stride == esize) return;
    const char* b = (const char*) f.base;
    if (!lo || b < lo) lo = b;
    if (!hi || b + esize > hi) hi = b + esize;
    if (stride == 0) stride = f.stride; else if (stride != f.stride) bulk_ok = false;
  };
  for (int k = 0; k < kNDoubleFields; k++) consider(vfield(view, kDoubleFields[k].off), sizeof(double));
  for (int k = 0; k < kNIntFields; k++) consider(vfield(view, kIntFields[k].off), sizeof(int32_t));
  if (lo && (size_t)(hi - lo) > stride) bulk_ok = false;

  const int nb = (int)((N + 255)/256);
  /* Contiguous fields travel on the copy stream, geometry (r, h, integer fields) first: the tree
     build only needs those and starts while v, m, u, ... are still in flight.  The copy stream first
     waits for the compute stream, which may still be reading the staging arrays of the last step. */
  CUDA_TRY(h, cudaEventRecord(h->ev_compute, h->stream));
  CUDA_TRY(h, cudaStreamWaitEvent(h->copy_stream, h->ev_compute, 0));
  /* pass 0: geometry, 1: everything else but the velocities, 2: the velocities (first needed by the force pass) */
  for (int pass = 0; pass < 3; pass++) {
    if (pass == 2) CUDA_TRY(h, cudaEventRecord(h->ev_early, h->copy_stream));
    for (int k = 0; k < kNDoubleFields; k++) {
      const int pf = kDoubleFields[k].pf;
      const bool geom = (pf >= F_RX && pf <= F_RZ) || pf == F_H;
      const bool vel = pf >= F_VX && pf <= F_VZ;
      if ((geom ? 0 : (vel ? 2 : 1)) != pass) continue;
      const gsph_field& f = vfield(view, kDoubleFields[k].off);
      const bool unused_dim = (pf >= F_RX && pf <= F_ATZ) && ((pf - F_RX) % 3 >= nd);
      if (!f.base || unused_dim) {
        /* absent: the reference's constructor default; the staging array keeps it until something writes there */
        if (!((h->hdef_d >> pf) & 1ull)) fill_kernel<double><<<nb, 256, 0, h->copy_stream>>>(N, h->H[pf], default_double(pf));
        h->hdef_d |= 1ull << pf; h->present_d &= ~(1ull << pf);
        if (pf >= F_RX && pf <= F_RZ && !(bulk_ok && (plan.nf || plan.ni)))
          CUDA_TRY(h, cudaEventRecord(h->ev_axis[pf - F_RX], h->copy_stream));
        continue;
      }
      h->present_d |= 1ull << pf; h->hdef_d &= ~(1ull << pf);
      if (f.stride == sizeof(double)) {
        if (nr > 0) CUDA_TRY(h, cudaMemcpyAsync(h->H[pf] + r0, (const double*) f.base + r0, nr*sizeof(double), cudaMemcpyHostToDevice, h->copy_stream));
      }
      else if (!whole) { h->err = "gsph_upload_range needs contiguous (SoA) fields"; return GSPH_EINVAL; }
      else if (bulk_ok) { plan.off_d[plan.nf] = (unsigned)((const char*) f.base - lo); plan.pf_d[plan.nf++] = pf; }
      else {
        rc = ensure_pinned(h, N*sizeof(double)); if (rc) return rc;
        CUDA_TRY(h, cudaStreamSynchronize(h->copy_stream));
        double* st = (double*) h->pinned; const char* b = (const char*) f.base;
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < N; i++) st[i] = *(const double*)(b + (size_t) i*f.stride);
        CUDA_TRY(h, cudaMemcpyAsync(h->H[pf], st, N*sizeof(double), cudaMemcpyHostToDevice, h->copy_stream));
      }
      if (pf >= F_RX && pf <= F_RZ && !(bulk_ok && (plan.nf || plan.ni)))
        CUDA_TRY(h, cudaEventRecord(h->ev_axis[pf - F_RX], h->copy_stream));
    }
    if (pass >= 1) continue;
    for (int k = 0; k < kNIntFields; k++) {
      const int pf = kIntFields[k].pf;
      const gsph_field& f = vfield(view, kIntFields[k].off);
      if (!f.base) {
        if (!((h->hdef_i >> pf) & 1u)) fill_kernel<int32_t><<<nb, 256, 0, h->copy_stream>>>(N, h->HI[pf], default_int(pf));
        h->hdef_i |= 1u << pf; h->present_i &= ~(1u << pf);
        continue;
      }
      h->present_i |= 1u << pf; h->hdef_i &= ~(1u << pf);
      if (f.stride == sizeof(int32_t)) {
        if (nr > 0) CUDA_TRY(h, cudaMemcpyAsync(h->HI[pf] + r0, (const int32_t*) f.base + r0, nr*sizeof(int32_t), cudaMemcpyHostToDevice, h->copy_stream));
      }
      else if (!whole) { h->err = "gsph_upload_range needs contiguous (SoA) fields"; return GSPH_EINVAL; }
      else if (bulk_ok) { plan.off_i[plan.ni] = (unsigned)((const char*) f.base - lo); plan.pf_i[plan.ni++] = pf; }
      else {
        rc = ensure_pinned(h, N*sizeof(int32_t)); if (rc) return rc;
        CUDA_TRY(h, cudaStreamSynchronize(h->copy_stream));
        int32_t* st = (int32_t*) h->pinned; const char* b = (const char*) f.base;
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < N; i++) st[i] = *(const int32_t*)(b + (size_t) i*f.stride);
        CUDA_TRY(h, cudaMemcpyAsync(h->HI[pf], st, N*sizeof(int32_t), cudaMemcpyHostToDevice, h->copy_stream));
      }
    }
    if (!(bulk_ok && (plan.nf || plan.ni))) CUDA_TRY(h, cudaEventRecord(h->ev_geom, h->copy_stream));
  }
  if (bulk_ok && (plan.nf || plan.ni)) {
    /* AoS records [lo + i*stride, +span) in chunks through the tree-build scratch buffer */
    const size_t chunk_bytes = (size_t) 256 << 20;
    int64_t per = (int64_t)(chunk_bytes/stride); if (per < 1) per = 1;
    if (h->scratch_bytes < (size_t) std::min<int64_t>(per, N)*stride) {
      if (h->scratch) cudaFree(h->scratch);
      h->scratch = nullptr; h->scratch_bytes = 0;
      const size_t want = (size_t) std::min<int64_t>(per, N)*stride;
      CUDA_TRY(h, cudaMalloc(&h->scratch, want)); h->scratch_bytes = want;
    }
    DevPtrs out = devptrs(h->H, h->HI);
    for (int64_t n0 = 0; n0 < N; n0 += per) {
      const int64_t n = std::min<int64_t>(per, N - n0);
      size_t bytes = (size_t) n*stride;
      if (n0 + n == N) bytes = (size_t)(n - 1)*stride + (size_t)(hi - lo);   /* do not read past the last record */
      CUDA_TRY(h, cudaMemcpyAsync(h->scratch, lo + (size_t) n0*stride, bytes, cudaMemcpyHostToDevice, h->copy_stream));
      aos_unpack_kernel<<<(int)((n + 255)/256), 256, 0, h->copy_stream>>>(n0, n, (const char*) h->scratch, stride, plan, out);
      CUDA_TRY(h, cudaStreamSynchronize(h->copy_stream));
    }
    CUDA_TRY(h, cudaEventRecord(h->ev_geom, h->copy_stream));
    for (int k = 0; k < 3; k++) CUDA_TRY(h, cudaEventRecord(h->ev_axis[k], h->copy_stream));
    CUDA_TRY(h, cudaEventRecord(h->ev_early, h->copy_stream));
  }
  CUDA_TRY(h, cudaEventRecord(h->ev_all, h->copy_stream));
  h->upload_pending = true;
  CUDA_TRY(h, cudaGetLastError());
  if (N != h->N) h->have_tree = false;
  h->N = N; h->uploaded = true; h->sorted_valid = false; h->density_done = false;
  h->slice_mode = false; h->dp.slice_lo = 0; h->dp.slice_hi = 0x7fffffff;
  h->integ.on = false; h->results_newer = false; h->p_stale = false;
  h->stats.Nhydro = N;
  tm.stop();
  return GSPH_OK;
}

int gsph_upload(gsph_handle* h, const gsph_particle_view* view, int64_t N) { return upload_impl(h, view, N, 0, N); }

/* multi-process hosts: this rank ships only the host-order particles [lo, hi) of the N-particle set;
   the other ranks' shares reach the staging arrays through the host layer's NCCL all-gather
   (gsph_device_array(..., GSPH_ARRAY_STAGING, ...)) -- NVLink instead of R x the PCIe traffic */
int gsph_upload_range(gsph_handle* h, const gsph_particle_view* view, int64_t N, int64_t lo, int64_t hi)
{
  return upload_impl(h, view, N, lo, hi);
}

static int download_impl(gsph_handle* h, const gsph_particle_view* view, uint32_t field_mask, bool wait,
                         int64_t r0 = 0, int64_t r1 = -1)
{
  if (!h) return GSPH_EINVAL;
  if (!view) { h->err = "gsph_download: NULL view"; return GSPH_EINVAL; }
  if (!h->sorted_valid && !h->integ.on) { h->err = "gsph_download before gsph_build_tree"; return GSPH_ESTATE; }
  cudaSetDevice(h->device);
  Timer tm(h, &h->stats.ms_download);
  const int64_t N = h->N;
  if (r1 < 0) r1 = N;
  if (r0 < 0 || r1 < r0 || r1 > N) { h->err = "gsph_download_range: bad range"; return GSPH_EINVAL; }
  const bool whole = (r0 == 0 && r1 == N);
  const int64_t nr = r1 - r0;
  unsigned long long dmask = 0; unsigned imask = 0;
  auto want_d = [&](int pf) {
    if (field_mask == GSPH_F_ALL) return true;
    const bool dens = pf == F_H || pf == F_RHO || pf == F_HR2 || pf == F_HFAC || pf == F_INVOMEGA || pf == F_ZETA ||
                      pf == F_U || pf == F_SOUND || pf == F_PRESS || pf == F_DIVV ||
                      ((pf == F_ALPHA || pf == F_DALPHADT) && h->dp.tdavisc == GSPH_TDAVISC_CD2010);
    const bool forc = (pf >= F_AX && pf <= F_ATZ) || pf == F_DUDT || pf == F_DIVV || pf == F_DALPHADT ||
                      pf == F_GPOT || pf == F_GPOTHYDRO;
    return ((field_mask & GSPH_F_DENSITY) && dens) || ((field_mask & GSPH_F_FORCES) && forc);
  };
  auto want_i = [&](int pf) {
    if (field_mask == GSPH_F_ALL) return true;
    return ((field_mask & GSPH_F_DENSITY) && pf == I_FLAGS) || ((field_mask & GSPH_F_FORCES) && pf == I_LEVELNEIB);
  };
  for (int k = 0; k < kNDoubleFields; k++)
    if (vfield(view, kDoubleFields[k].off).base && want_d(kDoubleFields[k].pf)) dmask |= 1ull << kDoubleFields[k].pf;
  for (int k = 0; k < kNIntFields; k++)
    if (vfield(view, kIntFields[k].off).base && want_i(kIntFields[k].pf)) imask |= 1u << kIntFields[k].pf;
  /* tree order -> host order on the compute stream, then the copies on the copy stream: they
     overlap whatever the compute stream does next (the force pass, for a density download) */
  /* (the device-resident main loop keeps the state in host order already: there only results of passes
     that ran since the last scatter -- operator calls between the halves of a step -- have to come back) */
  if (!h->integ.on || h->results_newer) {
    /* only the fields whose tree-order copy carries information (permuted inputs, results of the passes);
       the others are still exactly what the host uploaded (or the defaults) in host order */
    const unsigned long long sd = dmask & h->pvalid_d; const unsigned si = imask & h->pvalid_i;
    if (sd || si) {
      scatter_fields_kernel<<<(int)((N + 255)/256), 256, 0, h->stream>>>((int) N, h->perm, devptrs(h->P, h->PI),
                                                                       devptrs(h->H, h->HI), sd, si);
      h->stats.kernel_launches++;
      CUDA_TRY(h, cudaGetLastError());
      h->hdef_d &= ~sd; h->hdef_i &= ~si;
    }
  }
  CUDA_TRY(h, cudaEventRecord(h->ev_compute, h->stream));
  CUDA_TRY(h, cudaStreamWaitEvent(h->copy_stream, h->ev_compute, 0));
  cudaStream_t cs = h->copy_stream;
  for (int k = 0; k < kNDoubleFields; k++) {
    const int pf = kDoubleFields[k].pf;
    if (!((dmask >> pf) & 1ull)) continue;
    const gsph_field& f = vfield(view, kDoubleFields[k].off);
    if (f.stride == sizeof(double)) {
      if (nr > 0) CUDA_TRY(h, cudaMemcpyAsync((double*) f.base + r0, h->H[pf] + r0, nr*sizeof(double), cudaMemcpyDeviceToHost, cs));
    }
    else if (!whole) { h->err = "gsph_download_range needs contiguous (SoA) fields"; return GSPH_EINVAL; }
    else {
      int rc = ensure_pinned(h, N*sizeof(double)); if (rc) return rc;
      CUDA_TRY(h, cudaMemcpyAsync(h->pinned, h->H[pf], N*sizeof(double), cudaMemcpyDeviceToHost, cs));
      CUDA_TRY(h, cudaStreamSynchronize(cs));
      const double* st = (const double*) h->pinned; char* b = (char*) f.base;
#pragma omp parallel for schedule(static)
      for (int64_t i = 0; i < N; i++) *(double*)(b + (size_t) i*f.stride) = st[i];
    }
  }
  for (int k = 0; k < kNIntFields; k++) {
    const int pf = kIntFields[k].pf;
    if (!((imask >> pf) & 1u)) continue;
    const gsph_field& f = vfield(view, kIntFields[k].off);
    if (f.stride == sizeof(int32_t)) {
      if (nr > 0) CUDA_TRY(h, cudaMemcpyAsync((int32_t*) f.base + r0, h->HI[pf] + r0, nr*sizeof(int32_t), cudaMemcpyDeviceToHost, cs));
    }
    else if (!whole) { h->err = "gsph_download_range needs contiguous (SoA) fields"; return GSPH_EINVAL; }
    else {
      int rc = ensure_pinned(h, N*sizeof(int32_t)); if (rc) return rc;
      CUDA_TRY(h, cudaMemcpyAsync(h->pinned, h->HI[pf], N*sizeof(int32_t), cudaMemcpyDeviceToHost, cs));
      CUDA_TRY(h, cudaStreamSynchronize(cs));
      const int32_t* st = (const int32_t*) h->pinned; char* b = (char*) f.base;
#pragma omp parallel for schedule(static)
      for (int64_t i = 0; i < N; i++) *(int32_t*)(b + (size_t) i*f.stride) = st[i];
    }
  }
  if (wait) CUDA_TRY(h, cudaStreamSynchronize(cs));
  tm.stop();
  return GSPH_OK;
}

int gsph_download(gsph_handle* h, const gsph_particle_view* view, uint32_t field_mask)
{
  return download_impl(h, view, field_mask, true);
}

int gsph_download_async(gsph_handle* h, const gsph_particle_view* view, uint32_t field_mask)
{
  return download_impl(h, view, field_mask, false);
}

/* host-order particles [lo, hi) only (the share of this rank in a multi-process host); wait = 0 queues
   the copies like gsph_download_async */
int gsph_download_range(gsph_handle* h, const gsph_particle_view* view, uint32_t field_mask, int64_t lo, int64_t hi, int wait)
{
  return download_impl(h, view, field_mask, wait != 0, lo, hi);
}

/* device pointer of one particle array by name: "rx" "ry" "rz" "vx" "vy" "vz" "ax" .. "atz" "m" "h" "u" "rho"
   "hrangesqd" "hfactor" "invomega" "zeta" "sound" "pressure" "dudt" "div_v" "alpha" "dalphadt" "gpot"
   "gpot_hydro" (double) and "flags" "ptype" "level" "levelneib" "iorig" "sinkid" (int32) */
int gsph_device_array(gsph_handle* h, int which, const char* name, void** ptr, int64_t* capacity, int32_t* elem_size)
{
  if (!h || !name || !ptr) return GSPH_EINVAL;
  static const char* const dn[NF_DOUBLE] = {"rx", "ry", "rz", "vx", "vy", "vz", "ax", "ay", "az", "atx", "aty", "atz",
    "m", "h", "u", "rho", "hrangesqd", "hfactor", "invomega", "zeta", "sound", "pressure", "dudt", "div_v", "alpha",
    "dalphadt", "gpot", "gpot_hydro"};
  static const char* const in[NF_INT] = {"flags", "ptype", "level", "levelneib", "iorig", "sinkid"};
  if (which != GSPH_ARRAY_TREE_ORDER && which != GSPH_ARRAY_STAGING) { h->err = "gsph_device_array: bad array class"; return GSPH_EINVAL; }
  if (!h->H[0]) { h->err = "gsph_device_array before gsph_upload"; return GSPH_ESTATE; }
  for (int k = 0; k < NF_DOUBLE; k++) if (!strcmp(name, dn[k])) {
    *ptr = which == GSPH_ARRAY_STAGING ? h->H[k] : h->P[k];
    if (capacity) *capacity = h->Ncap;
    if (elem_size) *elem_size = 8;
    return GSPH_OK;
  }
  for (int k = 0; k < NF_INT; k++) if (!strcmp(name, in[k])) {
    *ptr = which == GSPH_ARRAY_STAGING ? (void*) h->HI[k] : (void*) h->PI[k];
    if (capacity) *capacity = h->Ncap;
    if (elem_size) *elem_size = 4;
    return GSPH_OK;
  }
  h->err = std::string("gsph_device_array: unknown array ") + name;
  return GSPH_EINVAL;
}

/*------------------------------------------------------------------------------------------------*/
int gsph_build_tree(gsph_handle* h, int mode, double dt)
{
  if (!h) return GSPH_EINVAL;
  if (!h->uploaded) { h->err = "gsph_build_tree before gsph_upload"; return GSPH_ESTATE; }
  if (mode < 0 || mode > 3) { h->err = "gsph_build_tree: mode must be 0, 1, 2 or 3"; return GSPH_EINVAL; }
  cudaSetDevice(h->device);
  Timer tm(h, &h->stats.ms_build_tree);
  if (!h->have_tree) mode = 0;
  int rc = gsph_tree_build(h, mode, dt);
  if (rc) return rc;
  int nact = 0;
  CUDA_TRY(h, cudaMemcpyAsync(&nact, h->tree.nactive, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  tm.stop();
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  h->all_active = (nact == h->N);
  h->stats.n_active = nact;
  h->p_stale = false;
  return GSPH_OK;
}

/* Multi-GPU tree build: rank `rank` of `world` (a power of two) builds the top
 * log2(world) levels and then only ITS subtree; see build_tree_v2 in tree.cu. */
int gsph_tree_shard_range(gsph_handle* h, int rank, int world, int64_t* lo, int64_t* hi)
{
  if (!h || !lo || !hi) return GSPH_EINVAL;
  if (world < 1 || (world & (world - 1)) || rank < 0 || rank >= world) { h->err = "gsph_tree_shard_range: world must be a power of two, 0 <= rank < world"; return GSPH_EINVAL; }
  int ltop = 0; while ((1 << ltop) < world) ltop++;
  int64_t a = 0, b = h->N;
  for (int bit = ltop - 1; bit >= 0; bit--) {
    const int64_t left = (b - a)/2;
    if ((rank >> bit) & 1) a += left; else b = a + left;
  }
  *lo = a; *hi = b;
  return GSPH_OK;
}

int gsph_build_tree_shard_begin(gsph_handle* h, int rank, int world)
{
  if (!h) return GSPH_EINVAL;
  if (!h->uploaded) { h->err = "gsph_build_tree_shard_begin before gsph_upload"; return GSPH_ESTATE; }
  if (world < 1 || (world & (world - 1)) || rank < 0 || rank >= world) { h->err = "gsph_build_tree_shard_begin: world must be a power of two, 0 <= rank < world"; return GSPH_EINVAL; }
  int ltot = 0;
  while ((int64_t) h->cfg.Nleafmax*((int64_t)1 << ltot) < h->N) ltot++;
  if ((1 << ltot) < world) { h->err = "gsph_build_tree_shard_begin: more ranks than leaf cells"; return GSPH_EINVAL; }
  cudaSetDevice(h->device);
  Timer tm(h, &h->stats.ms_build_tree);
  int rc = gsph_tree_build_sharded(h, 1, rank, world);
  if (rc) return rc;
  tm.stop();
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return GSPH_OK;
}

int gsph_build_tree_shard_finish(gsph_handle* h)
{
  if (!h) return GSPH_EINVAL;
  if (!h->uploaded) { h->err = "gsph_build_tree_shard_finish before gsph_upload"; return GSPH_ESTATE; }
  cudaSetDevice(h->device);
  float ms_begin = h->stats.ms_build_tree;
  Timer tm(h, &h->stats.ms_build_tree);
  int rc = gsph_tree_build_sharded(h, 2, 0, 1);
  if (rc) return rc;
  int nact = 0;
  CUDA_TRY(h, cudaMemcpyAsync(&nact, h->tree.nactive, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  tm.stop();
  h->stats.ms_build_tree += ms_begin;
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  h->all_active = (nact == h->N);
  h->stats.n_active = nact;
  h->p_stale = false;
  return GSPH_OK;
}

int gsph_update_active_counters(gsph_handle* h)
{
  if (!h) return GSPH_EINVAL;
  if (!h->have_tree || !h->sorted_valid) { h->err = "gsph_update_active_counters before gsph_build_tree"; return GSPH_ESTATE; }
  cudaSetDevice(h->device);
  int rc = gsph_tree_update_active(h);
  if (rc) return rc;
  int nact = 0;
  CUDA_TRY(h, cudaMemcpyAsync(&nact, h->tree.nactive, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  h->all_active = (nact == h->N); h->stats.n_active = nact;
  return GSPH_OK;
}

static int pass_gravity(gsph_handle* h, int);
static int run_with_regrow(gsph_handle* h, const char* where, int (*pass)(gsph_handle*, int), int arg)
{
  /* a cohort whose candidate list overflows its scratch slot is skipped and flagged (nothing of it
     is written); the pass is repeated for the flagged cohorts with doubled scratch -- the
     reference's "return -1 and double the buffer" convention (Tree.cpp:263-265,
     GradhSphTree.cpp:172-185) */
  for (int attempt = 0; attempt < 8; attempt++) {
    int rc = pass(h, arg);
    if (rc) return rc;
    if (pass == pass_gravity) {
      /* a gravity walk that outgrew its shared-memory stack: redo the unfinished cohorts with stacks
         in HBM (16 Ki entries per warp, quadrupled on every further overflow) */
      rc = fetch_status(h);
      if (rc) return rc;
      if (h->status_host->err_stack_overflow && attempt < 5) {
        if (!h->gstack_on && h->gstack_cap) h->gstack_on = true;        /* buffers of an earlier step */
        else {
          h->gstack_cap = h->gstack_cap ? h->gstack_cap*4 : 16384;
          if (h->gstack) cudaFree(h->gstack);
          h->gstack = nullptr; h->gstack_warps = 0;
          h->gstack_on = true;
        }
        CUDA_TRY(h, cudaMemsetAsync(&h->status->err_stack_overflow, 0, sizeof(int), h->stream));
        continue;
      }
    }
    rc = gsph_check_status(h, where);
    if (rc) return rc;
    const DevStatus& s = *h->status_host;
    if (!s.err_cand_overflow && !s.err_nlist_overflow) return GSPH_OK;
    if (s.err_nlist_overflow) {
      const int w = h->nlist_warps; const size_t want = h->nlist_cap*2;
      cudaFree(h->nlist); h->nlist = nullptr; h->nlist_warps = 0;
      rc = gsph_ensure_nlist(h, w, want);
      if (rc) return rc;
      CUDA_TRY(h, cudaMemsetAsync(&h->status->err_nlist_overflow, 0, sizeof(int), h->stream));
    }
    if (s.err_cand_overflow) {
      size_t want = h->cand_cap_per_warp*2;
      while (want < s.cand_need) want *= 2;
      const int w = h->cand_warps;
      cudaFree(h->cand); h->cand = nullptr; h->cand_warps = 0;
      rc = gsph_ensure_cand(h, w, want);
      if (rc) return rc;
    }
    CUDA_TRY(h, cudaMemsetAsync(&h->status->err_cand_overflow, 0, sizeof(int), h->stream));
  }
  h->err = std::string(where) + ": candidate scratch overflow persists";
  return GSPH_ENOMEM;
}

static int pass_density(gsph_handle* h, int) { return gsph_run_density(h); }
static int pass_hydro(gsph_handle* h, int g) { return gsph_run_hydro(h, g); }
static int pass_gravity(gsph_handle* h, int) { return gsph_run_gravity(h, 0); }
int gsph_prepare_done_flags(gsph_handle* h)
{
  const size_t need = (size_t) 1 << h->tree.ltot;
  if (need > h->done_cap) {
    if (h->done) cudaFree(h->done);
    h->done = nullptr; h->done_cap = 0;
    CUDA_TRY(h, cudaMalloc(&h->done, need));
    h->done_cap = need;
  }
  CUDA_TRY(h, cudaMemsetAsync(h->done, 0, need, h->stream));
  return GSPH_OK;
}

__global__ void commit_h_kernel(int lo, int N, const int32_t* flags, const double* __restrict__ hnew, double* __restrict__ h)
{
  const int i = lo + blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= N) return;
  const unsigned f = (unsigned) flags[i];
  if ((f & GSPH_FLAG_ACTIVE) && !(f & GSPH_FLAG_DEAD)) h[i] = hnew[i];
}

static int update_properties_impl(gsph_handle* h, bool refresh)
{
  if (!h) return GSPH_EINVAL;
  if (!h->have_tree || !h->sorted_valid) { h->err = "gsph_update_properties before gsph_build_tree"; return GSPH_ESTATE; }
  if (h->integ.on && h->p_stale) { h->err = "gsph_update_properties: the device-resident main loop changed v, u after the last BuildTree; call gsph_build_tree (or gsph_upload) first"; return GSPH_ESTATE; }
  h->results_newer = true;
  cudaSetDevice(h->device);
  Timer tm(h, &h->stats.ms_properties);
  int rc = reset_status(h); if (rc) return rc;
  rc = gsph_prepare_done_flags(h); if (rc) return rc;
  rc = run_with_regrow(h, "gsph_update_properties", pass_density, 0);
  if (rc) return rc;
  {
    /* the kernel leaves h untouched (a cohort may have to be redone with larger scratch and must
       then start from the same guess): commit the new smoothing lengths now */
    const int lo = h->dp.slice_lo, N = (int) std::min<int64_t>(h->N, h->dp.slice_hi);
    commit_h_kernel<<<std::max(1, (N - lo + 255)/256), 256, 0, h->stream>>>(lo, N, h->PI[I_FLAGS], h->hnew, h->P[F_H]);
    h->stats.kernel_launches++;
    CUDA_TRY(h, cudaGetLastError());
  }
  if (refresh) {
    rc = gsph_tree_update_hmax(h);                      /* tree->UpdateAllHmaxValues, GradhSphTree.cpp:268 */
    if (rc) return rc;
  }
  tm.stop();
  const DevStatus& s = *h->status_host;
  h->stats.n_active = (int64_t) s.ntarget;
  h->stats.hiter_mean = s.ntarget ? (double) s.hiter_sum/(double) s.ntarget : 0.0;
  h->stats.hiter_max = s.hiter_max;
  h->stats.density_retries = s.density_retries;
  h->density_done = true;
  h->pvalid_d |= GSPH_OUT_DENSITY; h->pvalid_i |= 1u << I_FLAGS;
  return GSPH_OK;
}

int gsph_update_properties(gsph_handle* h) { return update_properties_impl(h, true); }
int gsph_density_pass(gsph_handle* h) { return update_properties_impl(h, false); }

int gsph_refresh_hmax(gsph_handle* h)
{
  if (!h) return GSPH_EINVAL;
  if (!h->have_tree || !h->sorted_valid) { h->err = "gsph_refresh_hmax before gsph_build_tree"; return GSPH_ESTATE; }
  cudaSetDevice(h->device);
  return gsph_tree_update_hmax(h);
}

/* multi-GPU: after the exchange of h, rho, invomega, zeta -- the derived density fields of the tree-order
   particles outside [lo, hi) (the other ranks' slices; all of them active, see gsph_set_active_slice) */
int gsph_derive_density_fields(gsph_handle* h, int64_t lo, int64_t hi)
{
  if (!h) return GSPH_EINVAL;
  if (!h->have_tree || !h->sorted_valid) { h->err = "gsph_derive_density_fields before gsph_build_tree"; return GSPH_ESTATE; }
  if (lo < 0 || hi < lo || hi > h->N) { h->err = "gsph_derive_density_fields: bad range"; return GSPH_EINVAL; }
  cudaSetDevice(h->device);
  return gsph_derive_density_fields_impl(h, (int) lo, (int) hi);
}

/* multi-GPU: this device updates only the tree-order particles [lo, hi); the others are sources.  The
   kernels test the range themselves (DevParams::slice_lo / slice_hi): no flag is rewritten and the
   active counters of the tree stay those of the particle flags.                                        */
int gsph_set_active_slice(gsph_handle* h, int64_t lo, int64_t hi)
{
  if (!h) return GSPH_EINVAL;
  if (!h->have_tree || !h->sorted_valid) { h->err = "gsph_set_active_slice before gsph_build_tree"; return GSPH_ESTATE; }
  if (lo < 0 || hi < lo || hi > h->N) { h->err = "gsph_set_active_slice: bad range"; return GSPH_EINVAL; }
  h->dp.slice_lo = (int) lo; h->dp.slice_hi = (int) hi;
  h->slice_mode = true;
  return GSPH_OK;
}

int gsph_device_field(gsph_handle* h, int field, void** ptr, int64_t* capacity)
{
  if (!h || !ptr) return GSPH_EINVAL;
  static const int map[GSPH_DF_COUNT] = {F_H, F_RHO, F_HR2, F_HFAC, F_INVOMEGA, F_ZETA, F_SOUND, F_PRESS, F_U, -1};
  if (field < 0 || field >= GSPH_DF_COUNT) { h->err = "gsph_device_field: unknown field"; return GSPH_EINVAL; }
  if (field == GSPH_DF_PERM) {
    if (!h->perm) { h->err = "gsph_device_field before gsph_upload"; return GSPH_ESTATE; }
    *ptr = h->perm;
    if (capacity) *capacity = h->Ncap;
    return GSPH_OK;
  }
  if (!h->P[map[field]]) { h->err = "gsph_device_field before gsph_upload"; return GSPH_ESTATE; }
  *ptr = h->P[map[field]];
  if (capacity) *capacity = h->Ncap;
  return GSPH_OK;
}

int gsph_zero_accelerations(gsph_handle* h)
{
  if (!h) return GSPH_EINVAL;
  if (!h->sorted_valid) { h->err = "gsph_zero_accelerations before gsph_build_tree"; return GSPH_ESTATE; }
  if (h->integ.on && h->p_stale) { h->err = "gsph_zero_accelerations: the device-resident main loop changed v, u after the last BuildTree; call gsph_build_tree (or gsph_upload) first"; return GSPH_ESTATE; }
  h->results_newer = true;
  cudaSetDevice(h->device);
  const int lo = h->dp.slice_lo, N = (int) std::min<int64_t>(h->N, h->dp.slice_hi);   /* this device's targets */
  DevPtrs P = devptrs(h->P, h->PI);
  const int nb = std::max(1, (N - lo + 255)/256);
  switch (h->cfg.ndim) {
  case 1: zero_acc_kernel<1><<<nb, 256, 0, h->stream>>>(lo, N, P); break;
  case 2: zero_acc_kernel<2><<<nb, 256, 0, h->stream>>>(lo, N, P); break;
  default: zero_acc_kernel<3><<<nb, 256, 0, h->stream>>>(lo, N, P); break;
  }
  h->stats.kernel_launches++;
  CUDA_TRY(h, cudaGetLastError());
  if (h->slice_mode && !h->all_active) {
    zero_levelneib_outside_kernel<<<(int)((h->N + 255)/256), 256, 0, h->stream>>>((int) h->N, lo, N, P);
    h->stats.kernel_launches++;
    CUDA_TRY(h, cudaGetLastError());
  }
  h->pvalid_d |= GSPH_OUT_FORCES & ~GSPH_DMASK(F_DALPHADT); h->pvalid_i |= 1u << I_LEVELNEIB;
  h->acc_zeroed = true;
  return GSPH_OK;
}

/* Accumulators the host did not supply start from their default 0 (what the reference's constructors give):
   if no gsph_zero_accelerations ran since BuildTree, clear the ones the tree-order arrays do not hold yet. */
__global__ void zero_fields_kernel(int N, DevPtrs P, unsigned long long dmask, unsigned imask)
{
  const int i = blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= N) return;
  for (int f = 0; f < NF_DOUBLE; f++) if ((dmask >> f) & 1ull) P.f[f][i] = 0.0;
  for (int f = 0; f < NF_INT; f++) if ((imask >> f) & 1u) P.i[f][i] = 0;
}

static int prepare_accumulators(gsph_handle* h)
{
  unsigned long long acc = GSPH_OUT_FORCES;
  if (h->dp.tdavisc == GSPH_TDAVISC_CD2010) acc &= ~GSPH_DMASK(F_DALPHADT);      /* set by the density pass */
  const unsigned long long zd = acc & ~h->pvalid_d; const unsigned zi = (1u << I_LEVELNEIB) & ~h->pvalid_i;
  if (zd || zi) {
    const int N = (int) h->N;
    zero_fields_kernel<<<(N + 255)/256, 256, 0, h->stream>>>(N, devptrs(h->P, h->PI), zd, zi);
    h->stats.kernel_launches++;
    CUDA_TRY(h, cudaGetLastError());
  }
  h->pvalid_d |= GSPH_OUT_FORCES; h->pvalid_i |= 1u << I_LEVELNEIB;
  return GSPH_OK;
}

static void force_stats(gsph_handle* h)
{
  const DevStatus& s = *h->status_host;
  h->stats.n_active = (int64_t) s.ntarget;
  h->stats.ngb_mean = s.ntarget ? (double) s.ngb_sum/(double) s.ntarget : 0.0;
  h->stats.ngb_max = s.ngb_max;
  h->stats.cand_mean = s.ntarget ? (double) s.cand_sum/(double) s.ntarget : 0.0;
  h->stats.ndirect_mean = s.nleaf_grav ? (double) s.ndirect_sum/(double) s.nleaf_grav : 0.0;
  h->stats.ncell_mean = s.nleaf_grav ? (double) s.ncell_sum/(double) s.nleaf_grav : 0.0;
}

/* stars = Nstar x {r[ndim], m, h} doubles (the fields ComputeStarGravForces reads from
 * nbody->nbodydata, GradhSph.cpp:722-731); kept on the device until the next call; Nstar = 0 clears */
int gsph_set_stars(gsph_handle* h, const double* stars, int Nstar)
{
  if (!h) return GSPH_EINVAL;
  if (Nstar < 0 || (Nstar > 0 && !stars)) { h->err = "gsph_set_stars: bad arguments"; return GSPH_EINVAL; }
  cudaSetDevice(h->device);
  if (Nstar > h->stars_cap) {
    if (h->stars) cudaFree(h->stars);
    h->stars = nullptr; h->stars_cap = 0;
    CUDA_TRY(h, cudaMalloc(&h->stars, (size_t) Nstar*5*sizeof(double)));
    h->stars_cap = Nstar;
  }
  if (Nstar > 0) {
    CUDA_TRY(h, cudaMemcpyAsync(h->stars, stars, (size_t) Nstar*(h->cfg.ndim + 2)*sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  }
  h->Nstar = Nstar;
  h->dp.stars = h->stars; h->dp.nstar = Nstar;
  return GSPH_OK;
}

int gsph_update_hydro_forces(gsph_handle* h)
{
  if (!h) return GSPH_EINVAL;
  if (!h->have_tree || !h->sorted_valid) { h->err = "gsph_update_hydro_forces before gsph_build_tree"; return GSPH_ESTATE; }
  if (h->integ.on && h->p_stale) { h->err = "gsph_update_hydro_forces: the device-resident main loop changed v, u after the last BuildTree; call gsph_build_tree (or gsph_upload) first"; return GSPH_ESTATE; }
  h->results_newer = true;
  cudaSetDevice(h->device);
  Timer tm(h, &h->stats.ms_forces);
  int rc = gsph_ensure_velocities(h); if (rc) return rc;
  rc = reset_status(h); if (rc) return rc;
  rc = prepare_accumulators(h); if (rc) return rc;
  rc = gsph_prepare_done_flags(h); if (rc) return rc;
  rc = run_with_regrow(h, "gsph_update_hydro_forces", pass_hydro, 0);
  if (rc) return rc;
  rc = gsph_run_star_forces(h);                       /* GradhSphTree.cpp:385-392 */
  if (rc) return rc;
  tm.stop();
  force_stats(h);
  return GSPH_OK;
}

/* UpdateAllSphForces in two halves, so that a host can fetch what the first half already finished
 * (du/dt, div_v, dalpha/dt, levelneib) while the second half runs:
 *   begin  : the neighbour half -- hydro + smoothed-gravity pair terms (GradhSphTree.cpp:515-555);
 *   finish : the tree-gravity half -- interaction lists, direct sums, cells (:557-598) -- and the stars. */
int gsph_update_all_forces_begin(gsph_handle* h)
{
  if (!h) return GSPH_EINVAL;
  if (!h->have_tree || !h->sorted_valid) { h->err = "gsph_update_all_forces before gsph_build_tree"; return GSPH_ESTATE; }
  if (!h->cfg.self_gravity) { h->err = "gsph_update_all_forces needs self_gravity = 1 in the config"; return GSPH_EINVAL; }
  if (h->integ.on && h->p_stale) { h->err = "gsph_update_all_forces: the device-resident main loop changed v, u after the last BuildTree; call gsph_build_tree (or gsph_upload) first"; return GSPH_ESTATE; }
  h->results_newer = true;
  cudaSetDevice(h->device);
  h->forces_half_done = false;
  Timer tm(h, &h->stats.ms_forces);
  int rc = gsph_ensure_velocities(h); if (rc) return rc;
  rc = reset_status(h); if (rc) return rc;
  rc = prepare_accumulators(h); if (rc) return rc;
  rc = gsph_prepare_done_flags(h); if (rc) return rc;
  rc = run_with_regrow(h, "gsph_update_all_forces(neighbours)", pass_hydro, 1);
  if (rc) return rc;
  tm.stop();
  force_stats(h);
  h->forces_half_done = true;
  return GSPH_OK;
}

int gsph_update_all_forces_finish(gsph_handle* h)
{
  if (!h) return GSPH_EINVAL;
  if (!h->forces_half_done) { h->err = "gsph_update_all_forces_finish before gsph_update_all_forces_begin"; return GSPH_ESTATE; }
  cudaSetDevice(h->device);
  h->forces_half_done = false;
  const gsph_stats first = h->stats;
  Timer tm(h, &h->stats.ms_forces);
  int rc = gsph_prepare_done_flags(h); if (rc) return rc;
  h->gstack_on = false;                               /* first launch: shared-memory stacks */
  /* first the decoupled walk / consume kernels; the fused kernel then finishes what they left (cohorts whose
     lists or walk stacks outgrew their slots), with its own retry ladder.  GSPH_GRAV_FUSED=1: fused only. */
  const char* fused_only = getenv("GSPH_GRAV_FUSED");
  bool need_fused = fused_only && atoi(fused_only) != 0;
  if (!need_fused) {
    rc = gsph_run_gravity(h, 1); if (rc) return rc;
    rc = fetch_status(h); if (rc) return rc;
    need_fused = h->status_host->grav_left != 0;
    if (getenv("GSPH_DEBUG"))
      fprintf(stderr, "gsph: decoupled gravity pass left %d cohorts to the fused kernel; longest lists %d cells, %d direct leaves, deepest stack %d\n",
              h->status_host->grav_left, h->status_host->grav_cell_max, h->status_host->grav_dir_max, h->status_host->grav_stack_max);
  }
  if (need_fused) rc = run_with_regrow(h, "gsph_update_all_forces(gravity)", pass_gravity, 0);
  else rc = gsph_check_status(h, "gsph_update_all_forces(gravity)");
  if (rc) return rc;
  rc = gsph_run_star_forces(h);                       /* GradhSphTree.cpp:600-607 */
  if (rc) return rc;
  tm.stop();
  h->stats.ms_forces += first.ms_forces;
  force_stats(h);
  return GSPH_OK;
}

int gsph_update_all_forces(gsph_handle* h, const double* stars, int Nstar)
{
  if (!h) return GSPH_EINVAL;
  cudaSetDevice(h->device);
  if (stars || Nstar > 0) { const int rs = gsph_set_stars(h, stars, Nstar); if (rs) return rs; }
  const int rc = gsph_update_all_forces_begin(h);
  if (rc) return rc;
  return gsph_update_all_forces_finish(h);
}

/* HydroTree::UpdateAllStarGasForces (HydroTree.cpp:552-654) */
int gsph_update_star_gas_forces(gsph_handle* h, const double* stars, int Nstar, double* a, double* adot, double* gpot,
                                int32_t* counts)
{
  if (!h) return GSPH_EINVAL;
  if (!h->have_tree || !h->sorted_valid) { h->err = "gsph_update_star_gas_forces before gsph_build_tree"; return GSPH_ESTATE; }
  if (Nstar < 0 || (Nstar > 0 && (!stars || !a || !gpot))) { h->err = "gsph_update_star_gas_forces: bad arguments"; return GSPH_EINVAL; }
  if (h->cfg.kernel == GSPH_KERNEL_GAUSSIAN) { h->err = "gsph_update_star_gas_forces: the gaussian kernel has no gravity softening"; return GSPH_EINVAL; }
  for (int k = 0; k < h->cfg.ndim; k++)
    if (h->dp.periodic[k]) { h->err = "gsph_update_star_gas_forces: periodic (Ewald) gravity is not on the accelerated path"; return GSPH_EINVAL; }
  if (Nstar == 0) return GSPH_OK;
  cudaSetDevice(h->device);
  { const int rv = gsph_ensure_velocities(h); if (rv) return rv; }       /* adot reads the gas velocities */
  const int nd = h->cfg.ndim, row = 2*nd + 1;
  double* d = nullptr; int* di = nullptr;
  CUDA_TRY(h, cudaMalloc(&d, (size_t) Nstar*row*2*sizeof(double)));
  if (cudaMalloc(&di, ((size_t) Nstar*3 + 1)*sizeof(int)) != cudaSuccess) { cudaFree(d); h->err = "cudaMalloc failed"; return GSPH_ENOMEM; }
  double* dout = d + (size_t) Nstar*row;
  int rc = GSPH_OK;
  std::vector<double> res((size_t) Nstar*row);
  std::vector<int> cnt((size_t) Nstar*3 + 1);
  auto fail = [&](cudaError_t e) { if (e != cudaSuccess && rc == GSPH_OK) { h->err = cudaGetErrorString(e); rc = GSPH_ECUDA; } };
  fail(cudaMemcpyAsync(d, stars, (size_t) Nstar*row*sizeof(double), cudaMemcpyHostToDevice, h->stream));
  fail(cudaMemsetAsync(di, 0, ((size_t) Nstar*3 + 1)*sizeof(int), h->stream));
  if (rc == GSPH_OK) rc = gsph_run_star_gas(h, d, Nstar, dout, di, di + (size_t) Nstar*3);
  fail(cudaMemcpyAsync(res.data(), dout, res.size()*sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  fail(cudaMemcpyAsync(cnt.data(), di, cnt.size()*sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  fail(cudaStreamSynchronize(h->stream));
  cudaFree(d); cudaFree(di);
  if (rc) return rc;
  if (cnt[(size_t) Nstar*3]) { h->err = "gsph_update_star_gas_forces: tree-walk stack overflow"; return GSPH_ENOMEM; }
  for (int s = 0; s < Nstar; s++) {
    for (int k = 0; k < nd; k++) { a[(size_t) s*nd + k] += res[(size_t) s*row + k]; if (adot) adot[(size_t) s*nd + k] += res[(size_t) s*row + nd + k]; }
    gpot[s] += res[(size_t) s*row + 2*nd];
    if (counts) for (int k = 0; k < 3; k++) counts[3*s + k] = cnt[(size_t) 3*s + k];
  }
  return GSPH_OK;
}

int gsph_get_hmax(gsph_handle* h, double* hmax)
{
  if (!h || !hmax) return GSPH_EINVAL;
  if (!h->have_tree) { h->err = "gsph_get_hmax before gsph_build_tree"; return GSPH_ESTATE; }
  cudaSetDevice(h->device);
  CUDA_TRY(h, cudaMemcpyAsync(hmax, h->tree.hmax, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return GSPH_OK;
}

int gsph_synchronize(gsph_handle* h)
{
  if (!h) return GSPH_EINVAL;
  cudaSetDevice(h->device);
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->copy_stream));
  return GSPH_OK;
}

/*------------------------------------------------------------------------------------------------*/
int gsph_neighbour_counts(gsph_handle* h, int32_t* counts)
{
  if (!h || !counts) return GSPH_EINVAL;
  if (!h->have_tree || !h->sorted_valid) { h->err = "gsph_neighbour_counts before gsph_build_tree"; return GSPH_ESTATE; }
  cudaSetDevice(h->device);
  const int64_t N = h->N;
  int32_t* d = nullptr;
  CUDA_TRY(h, cudaMalloc(&d, N*sizeof(int32_t)));
  CUDA_TRY(h, cudaMemsetAsync(d, 0, N*sizeof(int32_t), h->stream));
  int rc = reset_status(h);
  if (!rc) rc = gsph_run_neiblist(h, d, nullptr, nullptr);
  if (!rc) rc = gsph_check_status(h, "gsph_neighbour_counts");
  if (!rc && h->status_host->err_cand_overflow) { h->err = "gsph_neighbour_counts: candidate scratch overflow (set GSPH_CAND_CAP)"; rc = GSPH_ENOMEM; }
  if (!rc && cudaMemcpy(counts, d, N*sizeof(int32_t), cudaMemcpyDeviceToHost) != cudaSuccess) { h->err = "cudaMemcpy failed"; rc = GSPH_ECUDA; }
  cudaFree(d);
  return rc;
}

int gsph_neighbour_lists(gsph_handle* h, const int64_t* offsets, int32_t* idx)
{
  if (!h || !offsets || !idx) return GSPH_EINVAL;
  if (!h->have_tree || !h->sorted_valid) { h->err = "gsph_neighbour_lists before gsph_build_tree"; return GSPH_ESTATE; }
  cudaSetDevice(h->device);
  const int64_t N = h->N, total = offsets[N];
  int64_t* doff = nullptr; int32_t* didx = nullptr;
  CUDA_TRY(h, cudaMalloc(&doff, (N + 1)*sizeof(int64_t)));
  if (cudaMalloc(&didx, std::max<int64_t>(total, 1)*sizeof(int32_t)) != cudaSuccess) { cudaFree(doff); h->err = "cudaMalloc failed"; return GSPH_ENOMEM; }
  cudaMemcpy(doff, offsets, (N + 1)*sizeof(int64_t), cudaMemcpyHostToDevice);
  int rc = reset_status(h);
  if (!rc) rc = gsph_run_neiblist(h, nullptr, doff, didx);
  if (!rc) rc = gsph_check_status(h, "gsph_neighbour_lists");
  if (!rc && h->status_host->err_cand_overflow) { h->err = "gsph_neighbour_lists: candidate scratch overflow (set GSPH_CAND_CAP)"; rc = GSPH_ENOMEM; }
  if (!rc && cudaMemcpy(idx, didx, total*sizeof(int32_t), cudaMemcpyDeviceToHost) != cudaSuccess) { h->err = "cudaMemcpy failed"; rc = GSPH_ECUDA; }
  cudaFree(doff); cudaFree(didx);
  if (!rc) {
#pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t i = 0; i < N; i++) std::sort(idx + offsets[i], idx + offsets[i + 1]);
  }
  return rc;
}

/* depth-first id (KDTree::CreateTreeStructure, KDTree.cpp:362-433) of heap node (level, k) */
static inline int dfs_id(int ltot, int level, int k)
{
  int id = 0;
  for (int j = 0; j < level; j++) {
    const int bit = (k >> (level - 1 - j)) & 1;
    id += bit ? (1 << (ltot - j)) : 1;
  }
  return id;
}

int gsph_tree_info_get(gsph_handle* h, gsph_tree_info* info)
{
  if (!h || !info) return GSPH_EINVAL;
  if (!h->have_tree) { h->err = "gsph_tree_info_get before gsph_build_tree"; return GSPH_ESTATE; }
  info->ltot = h->tree.ltot; info->Ncell = h->tree.ncell; info->gtot = 1 << h->tree.ltot;
  info->Nleafmax = h->cfg.Nleafmax; info->Ntot = h->N;
  return GSPH_OK;
}

int gsph_tree_leaf_of(gsph_handle* h, int32_t* leaf_of)
{
  if (!h || !leaf_of) return GSPH_EINVAL;
  if (!h->have_tree) { h->err = "gsph_tree_leaf_of before gsph_build_tree"; return GSPH_ESTATE; }
  cudaSetDevice(h->device);
  const TreeDev& T = h->tree;
  const int nleaf = 1 << T.ltot;
  std::vector<int> first(nleaf), count(nleaf), perm(h->N);
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  CUDA_TRY(h, cudaMemcpy(first.data(), T.first + (nleaf - 1), nleaf*sizeof(int), cudaMemcpyDeviceToHost));
  CUDA_TRY(h, cudaMemcpy(count.data(), T.count + (nleaf - 1), nleaf*sizeof(int), cudaMemcpyDeviceToHost));
  CUDA_TRY(h, cudaMemcpy(perm.data(), h->perm, h->N*sizeof(int), cudaMemcpyDeviceToHost));
  for (int g = 0; g < nleaf; g++) {
    const int id = dfs_id(T.ltot, T.ltot, g);
    for (int p = first[g]; p < first[g] + count[g]; p++) leaf_of[perm[p]] = id;
  }
  return GSPH_OK;
}

int gsph_tree_cells(gsph_handle* h, double* box, double* com, double* scalars, double* q, int32_t* Nout)
{
  if (!h) return GSPH_EINVAL;
  if (!h->have_tree) { h->err = "gsph_tree_cells before gsph_build_tree"; return GSPH_ESTATE; }
  cudaSetDevice(h->device);
  { const int rv = gsph_ensure_vcom(h); if (rv) return rv; }
  const TreeDev& T = h->tree;
  const int nd = h->cfg.ndim, nc = T.ncell;
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  std::vector<double> tmp(nc);
  std::vector<int> dfs(nc), cnt(nc);
  for (int l = 0; l <= T.ltot; l++)
    for (int k = 0; k < (1 << l); k++) dfs[(1 << l) - 1 + k] = dfs_id(T.ltot, l, k);
  auto pull = [&](const double* dev) { return cudaMemcpy(tmp.data(), dev, nc*sizeof(double), cudaMemcpyDeviceToHost); };
  if (box) {
    const double* src[4][3] = {{T.bbmin[0], T.bbmin[1], T.bbmin[2]}, {T.bbmax[0], T.bbmax[1], T.bbmax[2]},
                               {T.hbmin[0], T.hbmin[1], T.hbmin[2]}, {T.hbmax[0], T.hbmax[1], T.hbmax[2]}};
    for (int b = 0; b < 4; b++) for (int k = 0; k < nd; k++) {
      CUDA_TRY(h, pull(src[b][k]));
      for (int n = 0; n < nc; n++) box[((size_t) dfs[n]*4 + b)*nd + k] = tmp[n];
    }
  }
  if (com) {
    const double* src[2][3] = {{T.rcell[0], T.rcell[1], T.rcell[2]}, {T.com[0], T.com[1], T.com[2]}};
    for (int b = 0; b < 2; b++) for (int k = 0; k < nd; k++) {
      CUDA_TRY(h, pull(src[b][k]));
      for (int n = 0; n < nc; n++) com[((size_t) dfs[n]*2 + b)*nd + k] = tmp[n];
    }
  }
  if (scalars) {
    const double* src[4] = {T.m, T.rmax, T.hmax, T.cdistsqd};
    for (int b = 0; b < 4; b++) {
      CUDA_TRY(h, pull(src[b]));
      for (int n = 0; n < nc; n++) scalars[(size_t) dfs[n]*5 + b] = tmp[n];
    }
    for (int n = 0; n < nc; n++) scalars[(size_t) n*5 + 4] = 0.0;      /* eigen-MAC value: not on this path */
  }
  if (q) for (int b = 0; b < 5; b++) {
    CUDA_TRY(h, pull(T.q[b]));
    for (int n = 0; n < nc; n++) q[(size_t) dfs[n]*5 + b] = tmp[n];
  }
  if (Nout) {
    CUDA_TRY(h, cudaMemcpy(cnt.data(), T.count, nc*sizeof(int), cudaMemcpyDeviceToHost));
    for (int n = 0; n < nc; n++) Nout[dfs[n]] = cnt[n];
  }
  return GSPH_OK;
}

/*================================================================================================
 * device-resident main loop
 *================================================================================================*/
#define INTEG_LAUNCH(KERN, ...)                                                                      \
  do {                                                                                               \
    const int nb_ = (int)((h->N + 255)/256);                                                         \
    switch (h->cfg.ndim) {                                                                           \
    case 1: KERN<1><<<nb_, 256, 0, h->stream>>>(__VA_ARGS__); break;                                \
    case 2: KERN<2><<<nb_, 256, 0, h->stream>>>(__VA_ARGS__); break;                                \
    default: KERN<3><<<nb_, 256, 0, h->stream>>>(__VA_ARGS__); break;                               \
    }                                                                                                \
    h->stats.kernel_launches++;                                                                      \
    CUDA_TRY(h, cudaGetLastError());                                                                 \
  } while (0)

/* the particle state of the reference lives in the host-order arrays: bring the results of the
   density and force passes back from tree order (r, v, m never change there) */
static int integ_scatter_results(gsph_handle* h)
{
  unsigned long long dmask = 0;
  const int fields[] = {F_H, F_U, F_RHO, F_HR2, F_HFAC, F_INVOMEGA, F_ZETA, F_SOUND, F_PRESS, F_DUDT, F_DIVV, F_ALPHA,
                        F_DALPHADT, F_GPOT, F_GPOTHYDRO, F_AX, F_AY, F_AZ, F_ATX, F_ATY, F_ATZ};
  for (int f : fields) dmask |= 1ull << f;
  dmask &= h->pvalid_d;
  const unsigned imask = ((1u << I_FLAGS) | (1u << I_LEVELNEIB)) & h->pvalid_i;
  scatter_fields_kernel<<<(int)((h->N + 255)/256), 256, 0, h->stream>>>((int) h->N, h->perm, devptrs(h->P, h->PI),
                                                                      devptrs(h->H, h->HI), dmask, imask);
  h->stats.kernel_launches++;
  CUDA_TRY(h, cudaGetLastError());
  h->results_newer = false;
  h->hdef_d = 0ull; h->hdef_i = 0u;                     /* the host-order arrays are the particle state now */
  return GSPH_OK;
}

static int integ_forces(gsph_handle* h)
{
  int rc = gsph_update_properties(h); if (rc) return rc;
  rc = gsph_zero_accelerations(h); if (rc) return rc;
  /* nradstep = 1: MainLoop refreshes the thermal properties of every particle here (SphSimulation.cpp:669-679);
     with one global step the density pass has just done exactly that for all of them */
  if (h->integ.cfg.Nlevels > 1 && (h->integ.on || h->integ.starting)) { rc = gsph_thermal_all_impl(h); if (rc) return rc; }
  if (h->cfg.self_gravity) rc = gsph_update_all_forces(h, NULL, 0);
  else if (h->cfg.hydro_forces) rc = gsph_update_hydro_forces(h);
  else { h->err = "Error: No forces included in simulation"; return GSPH_EINVAL; }   /* SphSimulation.cpp:467 */
  return rc;
}

/* Simulation::ComputeGlobalTimestep on the host-order state */
static int integ_global_timestep(gsph_handle* h, double* out)
{
  const unsigned long long init[2] = {0x7ff0000000000000ull, 0ull};
  CUDA_TRY(h, cudaMemcpyAsync(h->integ.red, init, sizeof init, cudaMemcpyHostToDevice, h->stream));
  INTEG_LAUNCH(integ_timestep_kernel, (int) h->N, devptrs(h->H, h->HI), h->dp, h->integ.cfg, h->integ.red);
  unsigned long long res[2];
  CUDA_TRY(h, cudaMemcpyAsync(res, h->integ.red, sizeof res, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  if (res[1]) { h->err = "Error : Hydro timestep too large"; return GSPH_ESTATE; }     /* SphIntegration.cpp:113-120 */
  double ts; memcpy(&ts, &res[0], sizeof ts);
  *out = ts;
  return GSPH_OK;
}

/* ---- star particles of the device-resident loop: the reference's N-body leapfrog, host side ---------- */
static int integ_push_stars(gsph_handle* h)
{
  IntegState& I = h->integ;
  const int nd = h->cfg.ndim;
  std::vector<double> rows((size_t) I.nstar*(nd + 2));
  for (int s = 0; s < I.nstar; s++) {
    for (int k = 0; k < nd; k++) rows[(size_t) s*(nd + 2) + k] = I.sr[(size_t) s*nd + k];
    rows[(size_t) s*(nd + 2) + nd] = I.sm[s]; rows[(size_t) s*(nd + 2) + nd + 1] = I.sh[s];
  }
  return gsph_set_stars(h, rows.data(), I.nstar);
}

/* MainLoop's N-body block (SphSimulation.cpp:771-812, Npec = 1): zero, gravity of the gas
   (UpdateAllStarGasForces), softened star-star gravity (NbodyLeapfrogKDK::CalculateDirectSmoothedGravForces) */
static int integ_star_forces(gsph_handle* h)
{
  IntegState& I = h->integ;
  const int nd = h->cfg.ndim, ns = I.nstar;
  /* with block time steps only the stars at the end of their own step get new forces (the `active` flag of
     Nbody*::AdvanceParticles); the others keep a, adot, gpot */
  const bool blk = I.cfg.Nlevels > 1 && (int) I.sact.size() == ns;
  std::vector<int> act;
  for (int s = 0; s < ns; s++) if (!blk || I.sact[s]) act.push_back(s);
  const int na = (int) act.size();
  if (na == 0) return GSPH_OK;
  for (int s : act) {
    for (int k = 0; k < nd; k++) { I.sa[(size_t) s*nd + k] = 0.0; if (I.hermite4) I.sadot[(size_t) s*nd + k] = 0.0; }
    I.sgpot[s] = 0.0;
  }
  if (h->cfg.self_gravity) {
    std::vector<double> rows((size_t) na*(2*nd + 1)), ga((size_t) na*nd), gad((size_t) na*nd), gg(na);
    for (int q = 0; q < na; q++) {
      const int s = act[q];
      for (int k = 0; k < nd; k++) { rows[(size_t) q*(2*nd + 1) + k] = I.sr[(size_t) s*nd + k]; rows[(size_t) q*(2*nd + 1) + nd + k] = I.sv[(size_t) s*nd + k]; }
      rows[(size_t) q*(2*nd + 1) + 2*nd] = I.sh[s];
    }
    const int rc = gsph_update_star_gas_forces(h, rows.data(), na, ga.data(), I.hermite4 ? gad.data() : NULL, gg.data(), NULL);
    if (rc) return rc;
    for (int q = 0; q < na; q++) {
      const int s = act[q];
      for (int k = 0; k < nd; k++) { I.sa[(size_t) s*nd + k] = ga[(size_t) q*nd + k]; if (I.hermite4) I.sadot[(size_t) s*nd + k] = gad[(size_t) q*nd + k]; }
      I.sgpot[s] = gg[q];
    }
  }
  const double twopi = 6.283185307179586;
  for (int i : act)
    for (int j = 0; j < ns; j++) {
      if (i == j) continue;
      double dr[3] = {0, 0, 0}, dv[3] = {0, 0, 0}, drsqd = 0.0, dvdr = 0.0;
      for (int k = 0; k < nd; k++) { dr[k] = I.sr[(size_t) j*nd + k] - I.sr[(size_t) i*nd + k]; dv[k] = I.sv[(size_t) j*nd + k] - I.sv[(size_t) i*nd + k]; }
      for (int k = 0; k < nd; k++) { drsqd = k == 0 ? dr[0]*dr[0] : drsqd + dr[k]*dr[k]; dvdr = k == 0 ? dv[0]*dr[0] : dvdr + dv[k]*dr[k]; }
      const double invhmean = 2.0/(I.sh[i] + I.sh[j]);
      if (!I.hermite4) {                                   /* NbodyLeapfrogKDK::CalculateDirectSmoothedGravForces */
        const double drmag = sqrt(drsqd) + 1.0e-20, invdrmag = 1.0/drmag;
        const double paux = I.sm[j]*invhmean*invhmean*
                            gsph_kernel_grav_host(nd, h->cfg.kernel, h->cfg.tabulated_kernel, 4, drmag*invhmean)*invdrmag;
        I.sgpot[i] += I.sm[j]*invhmean*gsph_kernel_grav_host(nd, h->cfg.kernel, h->cfg.tabulated_kernel, 5, drmag*invhmean);
        for (int k = 0; k < nd; k++) I.sa[(size_t) i*nd + k] += paux*dr[k];
      }
      else {                                               /* NbodyHermite4::CalculateDirectSmoothedGravForces (NbodyHermite4.cpp:77-140) */
        const double drmag = sqrt(drsqd), invdrmag = 1.0/sqrt(drsqd);
        const double drdt = dvdr*invdrmag;
        const double paux = I.sm[j]*invhmean*invhmean*
                            gsph_kernel_grav_host(nd, h->cfg.kernel, h->cfg.tabulated_kernel, 4, drmag*invhmean)*invdrmag;
        const double wmean = gsph_kernel_grav_host(nd, h->cfg.kernel, h->cfg.tabulated_kernel, 0, drmag*invhmean)*
                             (double) powf((float) invhmean, (float) nd);            /* (single precision in the reference too) */
        I.sgpot[i] += I.sm[j]*invhmean*gsph_kernel_grav_host(nd, h->cfg.kernel, h->cfg.tabulated_kernel, 5, drmag*invhmean);
        for (int k = 0; k < nd; k++) {
          I.sa[(size_t) i*nd + k] += paux*dr[k];
          I.sadot[(size_t) i*nd + k] += paux*dv[k] - 3.0*paux*drdt*invdrmag*dr[k] + 2.0*twopi*I.sm[j]*drdt*wmean*invdrmag*dr[k];
        }
      }
    }
  return GSPH_OK;
}

/* NbodyHermite4::CalculateAllStartupQuantities (NbodyHermite4.cpp:257-330): a2dot, a3dot of the stars from the
   (unsoftened) star-star terms, once, after the first force computation */
static void integ_star_startup_hermite(gsph_handle* h)
{
  IntegState& I = h->integ;
  const int nd = h->cfg.ndim, ns = I.nstar;
  std::fill(I.sa2dot.begin(), I.sa2dot.end(), 0.0); std::fill(I.sa3dot.begin(), I.sa3dot.end(), 0.0);
  auto dot = [&](const double* a, const double* b) { double s = 0.0; for (int k = 0; k < nd; k++) s = k == 0 ? a[0]*b[0] : s + a[k]*b[k]; return s; };
  for (int i = 0; i < ns; i++)
    for (int j = 0; j < ns; j++) {
      if (i == j) continue;
      double dr[3], dv[3], da[3], dadot[3], a[3], adot[3], a2dot[3];
      for (int k = 0; k < nd; k++) {
        dr[k] = I.sr[(size_t) j*nd + k] - I.sr[(size_t) i*nd + k]; dv[k] = I.sv[(size_t) j*nd + k] - I.sv[(size_t) i*nd + k];
        da[k] = I.sa[(size_t) j*nd + k] - I.sa[(size_t) i*nd + k]; dadot[k] = I.sadot[(size_t) j*nd + k] - I.sadot[(size_t) i*nd + k];
      }
      const double drsqd = dot(dr, dr) + 1.0e-50, dvsqd = dot(dv, dv);
      const double invdrsqd = 1.0/drsqd, invdrmag = sqrt(invdrsqd);
      const double drdt = dot(dv, dr)*invdrmag;
      const double i3 = pow(invdrmag, 3);
      for (int k = 0; k < nd; k++) { a[k] = I.sm[j]*dr[k]*i3; adot[k] = I.sm[j]*i3*(dv[k] - 3.0*drdt*invdrmag*dr[k]); }
      const double afac = dot(dv, dr)*invdrsqd;
      const double bfac = dvsqd*invdrsqd + afac*afac + dot(da, dr)*invdrsqd;
      const double cfac = 3.0*dot(dv, da)*invdrsqd + dot(dr, dadot)*invdrsqd + afac*(3.0*bfac - 4.0*afac*afac);
      for (int k = 0; k < nd; k++) {
        a2dot[k] = I.sm[j]*da[k]*invdrsqd*invdrmag - 6.0*afac*adot[k] - 3.0*bfac*a[k];
        I.sa2dot[(size_t) i*nd + k] += a2dot[k];
        I.sa3dot[(size_t) i*nd + k] += I.sm[j]*dadot[k]*invdrsqd*invdrmag - 9.0*afac*a2dot[k] - 9.0*bfac*adot[k] - 3.0*cfac*a[k];
      }
    }
}

/* NbodyHermite4::CorrectionTerms (NbodyHermite4.cpp:388-440): Hermite interpolation of a2dot, a3dot over the step,
   positions to 5th and velocities to 4th order; dt^3 in single precision like the reference's powf */
static void integ_star_correct_hermite(gsph_handle* h, double dt, int star = -1)
{
  IntegState& I = h->integ;
  const int nd = h->cfg.ndim;
  const double dt3 = (double) powf((float) dt, 3.0f), invdt = 1.0/dt;
  const size_t q0 = star < 0 ? 0 : (size_t) star*nd, q1 = star < 0 ? I.sa.size() : (size_t)(star + 1)*nd;
  for (size_t q = q0; q < q1; q++) {
    I.sa2dot[q] = (-6.0*(I.sa0[q] - I.sa[q]) - dt*(4.0*I.sadot0[q] + 2.0*I.sadot[q]))*invdt*invdt;
    I.sa3dot[q] = (12.0*(I.sa0[q] - I.sa[q]) + 6.0*dt*(I.sadot0[q] + I.sadot[q]))*invdt*invdt*invdt;
    I.sr[q] += I.sa2dot[q]*dt3*dt/24.0 + I.sa3dot[q]*dt3*dt*dt/120.0;
    I.sv[q] += I.sa2dot[q]*dt3/6.0 + I.sa3dot[q]*dt3*dt/24.0;
  }
}

/* NbodyLeapfrogKDK::Timestep over the stars */
static double integ_star_timestep(const gsph_handle* h, int only = -1)
{
  const IntegState& I = h->integ;
  const int nd = h->cfg.ndim;
  double ts = 9.9e50;
  for (int s = 0; s < I.nstar; s++) {
    if (only >= 0 && s != only) continue;
    auto sq = [&](const std::vector<double>& v) { double x = 0.0; for (int k = 0; k < nd; k++) { const double a = v[(size_t) s*nd + k]; x = k == 0 ? a*a : x + a*a; } return x; };
    const double asqd = sq(I.sa);
    if (!I.hermite4) { ts = std::min(ts, I.nbody_mult*sqrt(I.sh[s]/(sqrt(asqd) + 1.0e-50))); continue; }
    /* NbodyHermite4::Timestep (NbodyHermite4.cpp:542-578): the Aarseth criterion */
    const double small = 1.0e-50, a1sqd = sq(I.sadot), a2sqd = sq(I.sa2dot), a3sqd = sq(I.sa3dot);
    double t1;
    if (a1sqd > small && a2sqd > small) t1 = I.nbody_mult*sqrt((sqrt(asqd*a2sqd) + a1sqd)/(sqrt(a1sqd*a3sqd) + a2sqd));
    else if (asqd > small && a2sqd > small) t1 = I.nbody_mult*sqrt(asqd/(a2sqd + small));
    else if (asqd > small) t1 = sqrt(I.sh[s]/(sqrt(asqd) + small));
    else t1 = 9.9e50;
    ts = std::min(ts, t1);
  }
  return ts;
}

/* ---- block time steps: host side of Simulation::ComputeBlockTimesteps (Simulation.cpp:1764-2199) ------------ */
static BlkPtrs blk_ptrs(gsph_handle* h)
{
  BlkPtrs B; B.nstep = h->integ.nstep; B.nlast = h->integ.nlast; B.dt = h->integ.pdt; B.dt_next = h->integ.dt_next; B.tlast = h->integ.ptlast;
  return B;
}
static inline bool integ_block(const gsph_handle* h) { return h->integ.cfg.Nlevels > 1; }

/* ComputeTimestepLevel (InlineFuncs.h:550-558) on the host, for the stars */
static inline int timestep_level_host(double dt, double dt_max)
{
  const int lv = (int)(1.44269504088896*log(dt_max/dt)) + 1;
  return lv > 0 ? lv : 0;
}

#define BLK_LAUNCH(MODE, OP, NFACTOR)                                                                        \
  INTEG_LAUNCH(blk_timestep_kernel, (int) h->N, devptrs(h->H, h->HI), blk_ptrs(h), h->dp, I.cfg, MODE, (int) I.n, I.level_step, \
               I.level_max, I.dt_max, I.timestep, OP, NFACTOR, I.red, single_level)

static int blk_compute_timesteps(gsph_handle* h)
{
  IntegState& I = h->integ;
  const unsigned long long init[3] = {0x7ff0000000000000ull, 0ull, 0ull};
  unsigned long long res[3];
  int single_level = -1;                                  /* sph_single_timestep: set below where the level is known */
  CUDA_TRY(h, cudaMemcpyAsync(I.red, init, sizeof init, cudaMemcpyHostToDevice, h->stream));
  if (I.n == I.nresync) {                                 /* synchronise every step and rebuild the hierarchy */
    I.n = 0;
    BLK_LAUNCH(0, 0, 1);
    CUDA_TRY(h, cudaMemcpyAsync(res, I.red, sizeof res, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    if (res[1]) { h->err = "Error : Hydro timestep too large"; return GSPH_ESTATE; }
    double ts; memcpy(&ts, &res[0], sizeof ts);
    I.timestep = ts;
    std::vector<double> sdt(I.nstar);
    for (int s = 0; s < I.nstar; s++) { sdt[s] = integ_star_timestep(h, s); I.timestep = std::min(I.timestep, sdt[s]); }
    I.level_max = I.cfg.Nlevels - 1;
    I.level_step = I.level_max;                            /* + integration_step - 1, = 0 for the KDK leapfrog */
    I.dt_max = I.timestep*pow(2.0, I.level_max);
    const int lmh = std::min(timestep_level_host(ts, I.dt_max), I.level_max);     /* level_max_hydro */
    if (I.cfg.sph_single_timestep) single_level = lmh;
    if (I.nstar > 0) {
      /* the stars sit at least on the deepest level the gas occupies (Simulation.cpp:1858-1872) */
      for (int s = 0; s < I.nstar; s++) {
        I.slevel[s] = std::max(std::min(timestep_level_host(sdt[s], I.dt_max), I.level_max), lmh);
        I.snlast[s] = I.n; I.snstep[s] = pow2int(I.level_step - (int) I.slevel[s]); I.stlast[s] = I.t; I.send[s] = 1;
      }
    }
    BLK_LAUNCH(1, 0, 1);
    I.nresync = (long long) pow2int(I.level_step);
    I.timestep = I.dt_max/(double) I.nresync;
  }
  else {
    const int level_max_old = I.level_max;
    BLK_LAUNCH(2, 0, 1);
    CUDA_TRY(h, cudaMemcpyAsync(res, I.red, sizeof res, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    if (res[1]) { h->err = "Error : Hydro timestep too large"; return GSPH_ESTATE; }
    I.level_max = (int) res[2];
    if (I.cfg.sph_single_timestep) single_level = I.level_max;    /* level_max_hydro of this step */
    if (I.nstar > 0) {                                     /* the stars at the end of their step (Simulation.cpp:2024-2068) */
      const int lmh = I.level_max;                         /* level_max_hydro */
      for (int s = 0; s < I.nstar; s++) {
        if (I.n - I.snlast[s] == I.snstep[s]) {
          const long long nstep = I.snstep[s]; const int last = (int) I.slevel[s];
          const int lv = std::max(timestep_level_host(integ_star_timestep(h, s), I.dt_max), lmh);
          if (lv < last && lv > lmh && last > 1 && I.n % (2*nstep) == 0) I.slevel[s] = last - 1;
          else if (lv > last) I.slevel[s] = lv;
          I.snlast[s] = I.n; I.snstep[s] = pow2int(I.level_step - (int) I.slevel[s]); I.stlast[s] = I.t; I.send[s] = 1;
        }
        I.level_max = std::max(I.level_max, (int) I.slevel[s]);
      }
    }
    const long long istep = pow2int(I.level_step - level_max_old + 1);
    int op = 0, nfactor = 1;
    if (I.level_max > level_max_old) {                     /* levels added: finer integer time */
      nfactor = pow2int(I.level_max - level_max_old); op = 1;
      I.n *= nfactor;
      for (int s = 0; s < I.nstar; s++) { I.snstep[s] *= nfactor; I.snlast[s] *= nfactor; }
    }
    else if (I.level_max <= level_max_old - 1 && level_max_old > 1 && istep > 0 && I.n % istep == 0) {
      I.level_max = level_max_old - 1;                     /* one level removed */
      nfactor = pow2int(level_max_old - I.level_max); op = 2;
      I.n /= nfactor;
      for (int s = 0; s < I.nstar; s++) { I.snlast[s] /= nfactor; I.snstep[s] /= nfactor; }
    }
    else I.level_max = level_max_old;
    I.level_step = I.level_max;
    I.nresync = (long long) pow2int(I.level_step);
    I.timestep = I.dt_max/(double) I.nresync;
    for (int s = 0; s < I.nstar; s++) if (I.snlast[s] == I.n) I.snstep[s] = pow2int(I.level_step - (int) I.slevel[s]);
    BLK_LAUNCH(3, op, nfactor);
  }
  if (!(I.timestep > 0.0)) { h->err = "Error : timestep fallen to zero"; return GSPH_ESTATE; }
  return GSPH_OK;
}

/* after a pass of a block-time-step MainLoop: results back to host order, every active flag cleared,
   SphLeapfrogKDK::CheckTimesteps; *woken = particles whose step was cut short (they need the passes now) */
static int blk_check_timesteps(gsph_handle* h, int* woken)
{
  IntegState& I = h->integ;
  int rc = integ_scatter_results(h); if (rc) return rc;
  const unsigned long long init[3] = {0x7ff0000000000000ull, 0ull, 0ull};
  unsigned long long res[3];
  CUDA_TRY(h, cudaMemcpyAsync(I.red, init, sizeof init, cudaMemcpyHostToDevice, h->stream));
  const int ldm = I.cfg.level_diff_max > 0 ? I.cfg.level_diff_max : 1;
  blk_check_kernel<<<(int)((h->N + 255)/256), 256, 0, h->stream>>>((int) h->N, devptrs(h->H, h->HI), blk_ptrs(h), (int) I.n, I.level_step, ldm, I.red);
  h->stats.kernel_launches++;
  CUDA_TRY(h, cudaMemcpyAsync(res, I.red, sizeof res, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  *woken = (int) res[1];
  return GSPH_OK;
}

/* the woken particles' flags and levels reach the tree-order arrays; active counters of the cells
   (sphneib->UpdateActiveParticleCounters, SphSimulation.cpp:661) */
__global__ void blk_push_flags_kernel(int N, const int32_t* __restrict__ perm, DevPtrs H, DevPtrs P)
{
  const int p = blockIdx.x*blockDim.x + threadIdx.x;
  if (p >= N) return;
  const int i = perm[p];
  P.i[I_FLAGS][p] = H.i[I_FLAGS][i]; P.i[I_LEVEL][p] = H.i[I_LEVEL][i];
}
static int blk_push_flags(gsph_handle* h)
{
  blk_push_flags_kernel<<<(int)((h->N + 255)/256), 256, 0, h->stream>>>((int) h->N, h->perm, devptrs(h->H, h->HI), devptrs(h->P, h->PI));
  h->stats.kernel_launches++;
  CUDA_TRY(h, cudaGetLastError());
  return gsph_update_active_counters(h);
}

/* rows = Nstar x {r[ndim], v[ndim], a[ndim], m, h}: the stars the next gsph_integrator_begin starts with
   (a is only read by compute_forces = 0).  Nstar = 0 removes them.                                    */
int gsph_integrator_set_stars(gsph_handle* h, const double* rows, int Nstar, double nbody_mult)
{
  if (!h) return GSPH_EINVAL;
  if (Nstar < 0 || (Nstar > 0 && !rows) || (Nstar > 0 && !(nbody_mult > 0))) { h->err = "gsph_integrator_set_stars: bad arguments"; return GSPH_EINVAL; }
  IntegState& I = h->integ;
  const int nd = h->cfg.ndim, row = 3*nd + 2;
  I.on = false;
  I.hermite4 = false;
  I.nstar = Nstar; I.nbody_mult = nbody_mult;
  I.sr.assign((size_t) Nstar*nd, 0.0); I.sv = I.sr; I.sa = I.sr;
  I.sm.assign(Nstar, 0.0); I.sh = I.sm; I.sgpot = I.sm;
  for (int s = 0; s < Nstar; s++) {
    for (int k = 0; k < nd; k++) {
      I.sr[(size_t) s*nd + k] = rows[(size_t) s*row + k]; I.sv[(size_t) s*nd + k] = rows[(size_t) s*row + nd + k];
      I.sa[(size_t) s*nd + k] = rows[(size_t) s*row + 2*nd + k];
    }
    I.sm[s] = rows[(size_t) s*row + 3*nd]; I.sh[s] = rows[(size_t) s*row + 3*nd + 1];
  }
  I.sr0 = I.sr; I.sv0 = I.sv; I.sa0 = I.sa;
  /* the gas passes see them from now on (a host layer may run the first passes itself, compute_forces = 2) */
  return Nstar > 0 ? integ_push_stars(h) : gsph_set_stars(h, NULL, 0);
}

/* rows = Nstar x {r[ndim], v[ndim], a[ndim], m, h, gpot} as they stand after the last step */
int gsph_integrator_get_stars(gsph_handle* h, double* rows, int Nstar)
{
  if (!h) return GSPH_EINVAL;
  const IntegState& I = h->integ;
  if (Nstar != I.nstar || (Nstar > 0 && !rows)) { h->err = "gsph_integrator_get_stars: bad arguments"; return GSPH_EINVAL; }
  const int nd = h->cfg.ndim, row = 3*nd + 3;
  for (int s = 0; s < Nstar; s++) {
    for (int k = 0; k < nd; k++) {
      rows[(size_t) s*row + k] = I.sr[(size_t) s*nd + k]; rows[(size_t) s*row + nd + k] = I.sv[(size_t) s*nd + k];
      rows[(size_t) s*row + 2*nd + k] = I.sa[(size_t) s*nd + k];
    }
    rows[(size_t) s*row + 3*nd] = I.sm[s]; rows[(size_t) s*row + 3*nd + 1] = I.sh[s]; rows[(size_t) s*row + 3*nd + 2] = I.sgpot[s];
  }
  return GSPH_OK;
}

/* the same for nbody = hermite4 (the reference's default N-body integrator, NbodyHermite4.cpp): rows = Nstar x
   {r[ndim], v[ndim], a[ndim], adot[ndim], a2dot[ndim], a3dot[ndim], m, h}; the derivatives are only read by
   gsph_integrator_begin(compute_forces = 0) -- otherwise begin computes them like PostInitialConditionsSetup */
int gsph_integrator_set_stars_hermite4(gsph_handle* h, const double* rows, int Nstar, double nbody_mult)
{
  if (!h) return GSPH_EINVAL;
  if (Nstar <= 0 || !rows || !(nbody_mult > 0)) { h->err = "gsph_integrator_set_stars_hermite4: bad arguments"; return GSPH_EINVAL; }
  IntegState& I = h->integ;
  const int nd = h->cfg.ndim, row = 6*nd + 2;
  std::vector<double> lf((size_t) Nstar*(3*nd + 2));
  for (int s = 0; s < Nstar; s++) {
    for (int k = 0; k < 3*nd; k++) lf[(size_t) s*(3*nd + 2) + k] = rows[(size_t) s*row + k];
    lf[(size_t) s*(3*nd + 2) + 3*nd] = rows[(size_t) s*row + 6*nd]; lf[(size_t) s*(3*nd + 2) + 3*nd + 1] = rows[(size_t) s*row + 6*nd + 1];
  }
  const int rc = gsph_integrator_set_stars(h, lf.data(), Nstar, nbody_mult);
  if (rc) return rc;
  I.hermite4 = true;
  I.sadot.assign((size_t) Nstar*nd, 0.0); I.sa2dot = I.sadot; I.sa3dot = I.sadot;
  for (int s = 0; s < Nstar; s++)
    for (int k = 0; k < nd; k++) {
      I.sadot[(size_t) s*nd + k] = rows[(size_t) s*row + 3*nd + k];
      I.sa2dot[(size_t) s*nd + k] = rows[(size_t) s*row + 4*nd + k];
      I.sa3dot[(size_t) s*nd + k] = rows[(size_t) s*row + 5*nd + k];
    }
  I.sadot0 = I.sadot;
  return GSPH_OK;
}

/* adot, a2dot, a3dot of the stars after the last step (Nstar x ndim each, NULL = not wanted); nbody = hermite4 only */
int gsph_integrator_get_star_derivatives(gsph_handle* h, double* adot, double* a2dot, double* a3dot)
{
  if (!h) return GSPH_EINVAL;
  const IntegState& I = h->integ;
  if (!I.hermite4 || I.nstar == 0) { h->err = "gsph_integrator_get_star_derivatives: no Hermite-4 stars"; return GSPH_ESTATE; }
  if (adot) memcpy(adot, I.sadot.data(), I.sadot.size()*sizeof(double));
  if (a2dot) memcpy(a2dot, I.sa2dot.data(), I.sa2dot.size()*sizeof(double));
  if (a3dot) memcpy(a3dot, I.sa3dot.data(), I.sa3dot.size()*sizeof(double));
  return GSPH_OK;
}

/* block time steps of the stars: level, nstep, nlast (int64[Nstar]) and tlast (double[Nstar]); NULL = not wanted */
int gsph_integrator_get_star_levels(gsph_handle* h, int64_t* level, int64_t* nstep, int64_t* nlast, double* tlast)
{
  if (!h) return GSPH_EINVAL;
  const IntegState& I = h->integ;
  if (!I.on || I.cfg.Nlevels <= 1 || (int) I.slevel.size() != I.nstar) { h->err = "gsph_integrator_get_star_levels: no star particles on block time steps"; return GSPH_ESTATE; }
  for (int s = 0; s < I.nstar; s++) {
    if (level) level[s] = I.slevel[s];
    if (nstep) nstep[s] = I.snstep[s];
    if (nlast) nlast[s] = I.snlast[s];
    if (tlast) tlast[s] = I.stlast[s];
  }
  return GSPH_OK;
}

int gsph_integrator_begin(gsph_handle* h, const gsph_integrator_config* cfg, int compute_forces, double t0)
{
  if (!h) return GSPH_EINVAL;
  if (!cfg) { h->err = "gsph_integrator_begin: NULL config"; return GSPH_EINVAL; }
  if (!h->uploaded) { h->err = "gsph_integrator_begin before gsph_upload"; return GSPH_ESTATE; }
  if (!(cfg->accel_mult > 0) || !(cfg->courant_mult > 0) || !(cfg->energy_mult > 0)) { h->err = "gsph_integrator_begin: the time-step multipliers must be positive"; return GSPH_EINVAL; }
  if (cfg->ntreebuildstep < 1 || cfg->ntreestockstep < 1) { h->err = "gsph_integrator_begin: ntreebuildstep and ntreestockstep must be >= 1"; return GSPH_EINVAL; }
  if (h->Nstar > 0 && h->integ.nstar == 0) { h->err = "gsph_integrator_begin: star particles of gsph_set_stars are not integrated (hand them to gsph_integrator_set_stars)"; return GSPH_EINVAL; }
  if (h->integ.nstar > 0 && h->cfg.nbody_softening != 1) { h->err = "gsph_integrator_begin: star particles need nbody_softening = 1"; return GSPH_EINVAL; }
  if (h->integ.nstar > 0) for (int k = 0; k < h->cfg.ndim; k++)
    if (h->dp.periodic[k] || h->dp.mirror[k][0] || h->dp.mirror[k][1]) { h->err = "gsph_integrator_begin: star particles need open boundaries"; return GSPH_EINVAL; }
  if (cfg->sink_particles != 0 || h->cfg.create_sinks != 0) {
    /* the reference's MainLoop would search for new sinks, accrete and update the stellar properties
       (SphSimulation.cpp:815-833); stepping on without them would silently diverge */
    h->err = "gsph_integrator_begin: sink_particles / create_sinks are not on the accelerated path of the device-resident main loop";
    return GSPH_EINVAL;
  }
  if (compute_forces < 0 || compute_forces > 2) { h->err = "gsph_integrator_begin: compute_forces must be 0, 1 or 2"; return GSPH_EINVAL; }
  if (compute_forces == 1 && h->slice_mode) { h->err = "gsph_integrator_begin: compute_forces = 1 not with an active slice (the multi-GPU host layer runs the passes itself and passes 2)"; return GSPH_EINVAL; }
  if (compute_forces == 2 && !h->sorted_valid) { h->err = "gsph_integrator_begin(compute_forces = 2) before the passes"; return GSPH_ESTATE; }
  cudaSetDevice(h->device);
  IntegState& I = h->integ;
  if (I.cap < h->Ncap) {
    for (int k = 0; k < 11; k++) { if (I.s0[k]) cudaFree(I.s0[k]); I.s0[k] = nullptr; }
    I.cap = 0;
    for (int k = 0; k < 11; k++) CUDA_TRY(h, cudaMalloc(&I.s0[k], (size_t) h->Ncap*sizeof(double)));
    I.cap = h->Ncap;
  }
  if (!I.red) CUDA_TRY(h, cudaMalloc(&I.red, 4*sizeof(unsigned long long)));
  if (cfg->Nlevels > 1) {
    if (cfg->Nlevels > 30) { h->err = "gsph_integrator_begin: Nlevels must be <= 30"; return GSPH_EINVAL; }
    if (I.blk_cap < h->Ncap) {
      auto fre = [](void* p) { if (p) cudaFree(p); };
      fre(I.nstep); fre(I.nlast); fre(I.pdt); fre(I.dt_next); fre(I.ptlast);
      I.nstep = I.nlast = nullptr; I.pdt = I.dt_next = I.ptlast = nullptr; I.blk_cap = 0;
      CUDA_TRY(h, cudaMalloc(&I.nstep, (size_t) h->Ncap*sizeof(int32_t)));
      CUDA_TRY(h, cudaMalloc(&I.nlast, (size_t) h->Ncap*sizeof(int32_t)));
      CUDA_TRY(h, cudaMalloc(&I.pdt, (size_t) h->Ncap*sizeof(double)));
      CUDA_TRY(h, cudaMalloc(&I.dt_next, (size_t) h->Ncap*sizeof(double)));
      CUDA_TRY(h, cudaMalloc(&I.ptlast, (size_t) h->Ncap*sizeof(double)));
      I.blk_cap = h->Ncap;
    }
  }
  I.cfg = *cfg; I.t = I.tlast = t0; I.Nsteps = 0; I.rebuild = true; I.on = false; I.mid_step = false;
  I.sact.clear(); I.send.clear();                        /* (no star is "inactive" before the hierarchy exists) */
  I.n = I.nresync = 0; I.level_max = I.level_step = 0; I.dt_max = 0.0;
  h->hdef_d = 0ull; h->hdef_i = 0u;                     /* the integrator writes the host-order arrays */
  CUDA_TRY(h, cudaStreamWaitEvent(h->stream, h->ev_all, 0));          /* the upload has landed */
  if (I.nstar > 0) { const int rs = integ_push_stars(h); if (rs) return rs; }
  if (compute_forces == 1) {
    int rc = gsph_build_tree(h, 0, 0.0); if (rc) return rc;
    I.starting = true;
    rc = integ_forces(h);
    I.starting = false;
    if (rc) return rc;
  }
  if (compute_forces) { const int rc = integ_scatter_results(h); if (rc) return rc; }
  if (I.nstar > 0) {
    if (compute_forces) {
      const int rs = integ_star_forces(h); if (rs) return rs;
      if (I.hermite4) integ_star_startup_hermite(h);       /* PostInitialConditionsSetup, SphSimulation.cpp:513-516 */
    }
    I.sr0 = I.sr; I.sv0 = I.sv; I.sa0 = I.sa;
    if (I.hermite4) I.sadot0 = I.sadot;
  }
  INTEG_LAUNCH(integ_init_kernel, (int) h->N, devptrs(h->H, h->HI), integ_ptrs(h));
  int rc;
  if (integ_block(h)) {
    /* PostInitialConditionsSetup: ComputeBlockTimesteps with n = nresync = 0 (a resynchronisation), then
       EndTimestep with part.dt = 0 (SphSimulation.cpp:544-563) */
    CUDA_TRY(h, cudaMemsetAsync(I.nstep, 0, (size_t) h->N*sizeof(int32_t), h->stream));
    CUDA_TRY(h, cudaMemsetAsync(I.nlast, 0, (size_t) h->N*sizeof(int32_t), h->stream));
    CUDA_TRY(h, cudaMemsetAsync(I.pdt, 0, (size_t) h->N*sizeof(double), h->stream));
    CUDA_TRY(h, cudaMemsetAsync(I.dt_next, 0, (size_t) h->N*sizeof(double), h->stream));
    fill_kernel<double><<<(int)((h->N + 255)/256), 256, 0, h->stream>>>(h->N, I.ptlast, t0);
    I.slevel.assign(I.nstar, 0); I.snstep.assign(I.nstar, 0); I.snlast.assign(I.nstar, 0); I.stlast.assign(I.nstar, t0);
    I.send.assign(I.nstar, 0); I.sact.assign(I.nstar, 0);
    rc = blk_compute_timesteps(h); if (rc) return rc;
    std::fill(I.send.begin(), I.send.end(), 0);             /* nbody->EndTimestep: r0 = r ... was set above */
    INTEG_LAUNCH(integ_end_kernel, (int) h->N, devptrs(h->H, h->HI), integ_ptrs(h), 0.0, I.cfg.energy_integration, 1,
                 (int) I.n, I.t, I.nlast, I.pdt, I.dt_next, I.ptlast);
    I.on = true;
    return GSPH_OK;
  }
  rc = integ_global_timestep(h, &I.timestep); if (rc) return rc;
  if (I.nstar > 0) I.timestep = std::min(I.timestep, integ_star_timestep(h));
  /* EndTimestep with part.dt = 0 (SphSimulation.cpp:561): u0 = u, dudt0 = dudt */
  INTEG_LAUNCH(integ_end_kernel, (int) h->N, devptrs(h->H, h->HI), integ_ptrs(h), 0.0, I.cfg.energy_integration, 0, 0, 0.0,
               (int32_t*) nullptr, (double*) nullptr, (double*) nullptr, (double*) nullptr);
  I.on = true;
  return GSPH_OK;
}

/* first half of one MainLoop: time advance, prediction, boundaries (SphSimulation.cpp:586-608); the
 * host-order arrays then hold a fresh "upload".  *tree_mode = what BuildTree has to do (0 rebuild,
 * 1 restock, 2 extrapolate), *timestep = the step being taken (the extrapolation interval).       */
int gsph_integrator_advance(gsph_handle* h, int* tree_mode, double* timestep)
{
  if (!h) return GSPH_EINVAL;
  if (!h->integ.on) { h->err = "gsph_integrator_advance before gsph_integrator_begin"; return GSPH_ESTATE; }
  if (h->integ.mid_step) { h->err = "gsph_integrator_advance: the previous step was not finished"; return GSPH_ESTATE; }
  cudaSetDevice(h->device);
  IntegState& I = h->integ;
  I.Nsteps++;
  I.n++;
  const double ts = I.timestep;
  I.t = I.t + ts;
  const double dt = I.t - I.tlast;
  INTEG_LAUNCH(integ_advance_kernel, (int) h->N, devptrs(h->H, h->HI), integ_ptrs(h), h->dp, dt, ts,
               h->dp.tdavisc != GSPH_TDAVISC_NONE ? 1 : 0, I.cfg.energy_integration, integ_block(h) ? 1 : 0, I.t, (int) I.n,
               (const int32_t*) I.nstep, (const int32_t*) I.nlast, (const double*) I.ptlast);
  if (I.nstar > 0 && integ_block(h)) {                 /* Nbody*::AdvanceParticles, every star from the start of ITS step */
    const int nd = h->cfg.ndim;
    for (int s = 0; s < I.nstar; s++) {
      const double ds = I.t - I.stlast[s];
      for (int k = 0; k < nd; k++) {
        const size_t q = (size_t) s*nd + k;
        I.sr[q] = I.sr0[q] + I.sv0[q]*ds + 0.5*I.sa0[q]*ds*ds;
        I.sv[q] = I.sv0[q] + I.sa0[q]*ds;
        if (I.hermite4) {
          I.sr[q] = I.sr0[q] + I.sv0[q]*ds + 0.5*I.sa0[q]*ds*ds + 0.16666666666666666666666*I.sadot0[q]*ds*ds*ds;
          I.sv[q] = I.sv0[q] + I.sa0[q]*ds + 0.5*I.sadot0[q]*ds*ds;
        }
      }
      I.sact[s] = (I.n - I.snlast[s] == I.snstep[s]) ? 1 : 0;
    }
    const int rs = integ_push_stars(h); if (rs) return rs;
  }
  else if (I.nstar > 0) {                              /* NbodyLeapfrogKDK::AdvanceParticles */
    for (size_t q = 0; q < I.sr.size(); q++) {
      I.sr[q] = I.sr0[q] + I.sv0[q]*dt + 0.5*I.sa0[q]*dt*dt;
      I.sv[q] = I.sv0[q] + I.sa0[q]*dt;
      if (I.hermite4) {                                /* NbodyHermite4::AdvanceParticles (NbodyHermite4.cpp:340-378) */
        I.sr[q] = I.sr0[q] + I.sv0[q]*dt + 0.5*I.sa0[q]*dt*dt + 0.16666666666666666666666*I.sadot0[q]*dt*dt*dt;
        I.sv[q] = I.sv0[q] + I.sa0[q]*dt + 0.5*I.sadot0[q]*dt*dt;
      }
    }
    const int rs = integ_push_stars(h); if (rs) return rs;
  }
  CUDA_TRY(h, cudaEventRecord(h->ev_geom, h->stream));
  CUDA_TRY(h, cudaEventRecord(h->ev_all, h->stream));
  CUDA_TRY(h, cudaEventRecord(h->ev_early, h->stream));
  for (int k = 0; k < 3; k++) CUDA_TRY(h, cudaEventRecord(h->ev_axis[k], h->stream));
  h->sorted_valid = false; h->density_done = false; h->results_newer = false;
  h->slice_mode = false; h->dp.slice_lo = 0; h->dp.slice_hi = 0x7fffffff;
  /* sphneib->BuildTree(rebuild_tree, Nsteps, ntreebuildstep, ntreestockstep, timestep, sph) */
  const bool full = I.rebuild || I.Nsteps % I.cfg.ntreebuildstep == 0;
  if (tree_mode) *tree_mode = full ? 0 : (I.Nsteps % I.cfg.ntreestockstep == 0 ? 1 : 2);
  if (timestep) *timestep = ts;
  I.mid_step = true;
  return GSPH_OK;
}

/* second half, after BuildTree, UpdateAllSphProperties, ZeroAccelerations and the force pass (with
 * several ranks: after the result slices were all-gathered in tree order): results back to host
 * order, smoothing-length sanity check, next global time step, end-of-step terms
 * (SphSimulation.cpp:757-767, 836-870).                                                           */
int gsph_integrator_finish(gsph_handle* h)
{
  if (!h) return GSPH_EINVAL;
  if (!h->integ.on || !h->integ.mid_step) { h->err = "gsph_integrator_finish before gsph_integrator_advance"; return GSPH_ESTATE; }
  if (integ_block(h) && h->results_newer) { h->err = "gsph_integrator_finish: block time steps (Nlevels > 1) need gsph_integrator_check_timesteps after the passes (until it reports 0)"; return GSPH_ESTATE; }
  if (!h->sorted_valid) { h->err = "gsph_integrator_finish before the passes of this step"; return GSPH_ESTATE; }
  cudaSetDevice(h->device);
  IntegState& I = h->integ;
  int rc = GSPH_OK;
  if (!integ_block(h)) { rc = integ_scatter_results(h); if (rc) return rc; }      /* (block steps: blk_check_timesteps did) */
  if (h->dp.any_periodic) {
    double hmax = 0.0;
    rc = gsph_get_hmax(h, &hmax); if (rc) return rc;
    hmax *= h->dp.kernrange;
    for (int k = 0; k < h->cfg.ndim; k++)
      if (h->dp.periodic[k] && h->dp.boxhalf[k] < 2*hmax) { h->err = "Error: Smoothing length too large, self-interaction will occur"; return GSPH_ESTATE; }
  }
  if (I.nstar > 0) {
    rc = integ_star_forces(h); if (rc) return rc;
    const double dts = I.t - I.tlast;
    if (integ_block(h)) {                              /* Nbody*::CorrectionTerms for the stars at the end of their step */
      const int nd = h->cfg.ndim;
      for (int s = 0; s < I.nstar; s++) {
        if (!I.sact[s]) continue;
        const double ds = I.t - I.stlast[s];
        if (I.hermite4) integ_star_correct_hermite(h, ds, s);
        else for (int k = 0; k < nd; k++) { const size_t q = (size_t) s*nd + k; I.sv[q] += 0.5*(I.sa[q] - I.sa0[q])*ds; }
      }
    }
    else if (I.hermite4) integ_star_correct_hermite(h, dts);
    else for (size_t q = 0; q < I.sv.size(); q++) I.sv[q] += 0.5*(I.sa[q] - I.sa0[q])*dts;     /* NbodyLeapfrogKDK::CorrectionTerms */
  }
  if (integ_block(h)) {
    rc = blk_compute_timesteps(h); if (rc) return rc;   /* new levels; I.n, I.timestep of the next step */
    I.rebuild = false;
    INTEG_LAUNCH(integ_end_kernel, (int) h->N, devptrs(h->H, h->HI), integ_ptrs(h), 0.0, I.cfg.energy_integration, 1,
                 (int) I.n, I.t, I.nlast, I.pdt, I.dt_next, I.ptlast);
    for (int s = 0; s < I.nstar; s++) {                 /* Nbody*::EndTimestep for the stars flagged end_timestep */
      if (!I.send[s]) continue;
      const int nd = h->cfg.ndim;
      for (int k = 0; k < nd; k++) {
        const size_t q = (size_t) s*nd + k;
        I.sr0[q] = I.sr[q]; I.sv0[q] = I.sv[q]; I.sa0[q] = I.sa[q];
        if (I.hermite4) I.sadot0[q] = I.sadot[q];
      }
      I.snlast[s] = I.n; I.stlast[s] = I.t; I.send[s] = 0;
    }
    I.tlast = I.t;
    I.mid_step = false;
    h->p_stale = true;
    return GSPH_OK;
  }
  double next = 0.0;
  rc = integ_global_timestep(h, &next); if (rc) return rc;
  if (I.nstar > 0) {
    next = std::min(next, integ_star_timestep(h));
    I.sr0 = I.sr; I.sv0 = I.sv; I.sa0 = I.sa;          /* Nbody{LeapfrogKDK,Hermite4}::EndTimestep */
    if (I.hermite4) I.sadot0 = I.sadot;
  }
  I.rebuild = false;
  INTEG_LAUNCH(integ_end_kernel, (int) h->N, devptrs(h->H, h->HI), integ_ptrs(h), I.timestep, I.cfg.energy_integration, 0, 0, 0.0,
               (int32_t*) nullptr, (double*) nullptr, (double*) nullptr, (double*) nullptr);
  I.tlast = I.t;
  I.timestep = next;
  I.mid_step = false;
  h->p_stale = true;                                    /* v, u were corrected in host order only */
  return GSPH_OK;
}

/* block time steps, host-driven loop (several devices): SphLeapfrogKDK::CheckTimesteps after the passes of
 * this step (with several ranks: after the result slices were combined).  *woken = particles whose step is
 * cut short; when > 0 their active flags and new levels are in place and the passes have to run again
 * (the do-while loop of MainLoop, SphSimulation.cpp:658-756).                                        */
int gsph_integrator_check_timesteps(gsph_handle* h, int* woken)
{
  if (!h) return GSPH_EINVAL;
  if (!h->integ.on || !h->integ.mid_step) { h->err = "gsph_integrator_check_timesteps before gsph_integrator_advance"; return GSPH_ESTATE; }
  if (!woken) { h->err = "gsph_integrator_check_timesteps: NULL argument"; return GSPH_EINVAL; }
  if (!h->sorted_valid) { h->err = "gsph_integrator_check_timesteps before the passes of this step"; return GSPH_ESTATE; }
  cudaSetDevice(h->device);
  *woken = 0;
  if (!integ_block(h)) return GSPH_OK;                 /* one global step: nothing to check */
  int rc = blk_check_timesteps(h, woken); if (rc) return rc;
  if (*woken > 0) rc = blk_push_flags(h);
  return rc;
}

/* GradhSph::ComputeThermalProperties for every particle, active or not, in tree order: what MainLoop does
 * between the density and the force pass when nradstep = 1 (SphSimulation.cpp:669-679).  gsph_integrator_step
 * does it itself; a host layer that drives the passes of a block-time-step run calls this.          */
int gsph_update_thermal_properties(gsph_handle* h)
{
  if (!h) return GSPH_EINVAL;
  if (!h->sorted_valid) { h->err = "gsph_update_thermal_properties before gsph_build_tree"; return GSPH_ESTATE; }
  cudaSetDevice(h->device);
  return gsph_thermal_all_impl(h);
}

int gsph_integrator_step(gsph_handle* h, int nsteps)
{
  if (!h) return GSPH_EINVAL;
  if (!h->integ.on) { h->err = "gsph_integrator_step before gsph_integrator_begin"; return GSPH_ESTATE; }
  if (nsteps < 0) { h->err = "gsph_integrator_step: nsteps < 0"; return GSPH_EINVAL; }
  for (int s = 0; s < nsteps; s++) {
    int mode = 0; double ts = 0.0;
    int rc = gsph_integrator_advance(h, &mode, &ts);
    if (rc) return rc;
    rc = gsph_build_tree(h, mode, ts); if (rc) return rc;
    rc = integ_forces(h); if (rc) return rc;
    if (integ_block(h)) {
      /* the passes are repeated for the particles whose step CheckTimesteps cuts short (SphSimulation.cpp:658-756) */
      for (int it = 0; ; it++) {
        int woken = 0;
        rc = blk_check_timesteps(h, &woken); if (rc) return rc;
        if (woken == 0) break;
        if (it > 64) { h->err = "gsph_integrator_step: the neighbour time-step limiter does not settle"; return GSPH_ESTATE; }
        rc = blk_push_flags(h); if (rc) return rc;
        rc = integ_forces(h); if (rc) return rc;
      }
    }
    rc = gsph_integrator_finish(h);
    if (rc) return rc;
  }
  return GSPH_OK;
}

int gsph_integrator_state(gsph_handle* h, double* t, double* timestep, int64_t* nsteps)
{
  if (!h) return GSPH_EINVAL;
  if (!h->integ.on) { h->err = "gsph_integrator_state before gsph_integrator_begin"; return GSPH_ESTATE; }
  if (t) *t = h->integ.t;
  if (timestep) *timestep = h->integ.timestep;
  if (nsteps) *nsteps = h->integ.Nsteps;
  return GSPH_OK;
}

int gsph_integrator_block_state(gsph_handle* h, gsph_block_state* st, int32_t* nstep, int32_t* nlast, double* dt, double* tlast)
{
  if (!h) return GSPH_EINVAL;
  if (!h->integ.on) { h->err = "gsph_integrator_block_state before gsph_integrator_begin"; return GSPH_ESTATE; }
  const IntegState& I = h->integ;
  if (st) { st->n = I.n; st->nresync = I.nresync; st->level_max = I.level_max; st->level_step = I.level_step; st->dt_max = I.dt_max; }
  if (nstep || nlast || dt || tlast) {
    if (I.cfg.Nlevels <= 1) { h->err = "gsph_integrator_block_state: per-particle step data exist with Nlevels > 1 only"; return GSPH_EINVAL; }
    cudaSetDevice(h->device);
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    const size_t N = (size_t) h->N;
    if (nstep) CUDA_TRY(h, cudaMemcpy(nstep, I.nstep, N*sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (nlast) CUDA_TRY(h, cudaMemcpy(nlast, I.nlast, N*sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (dt) CUDA_TRY(h, cudaMemcpy(dt, I.pdt, N*sizeof(double), cudaMemcpyDeviceToHost));
    if (tlast) CUDA_TRY(h, cudaMemcpy(tlast, I.ptlast, N*sizeof(double), cudaMemcpyDeviceToHost));
  }
  return GSPH_OK;
}

/* the start-of-step state of every particle in host order (SphLeapfrogKDK keeps it in the particle: r0, v0, a0,
   u0, dudt0; with block time steps also dt_next): what a host needs to carry a run on in the reference's own
   MainLoop.  r0, v0, a0: ndim arrays of N doubles one after the other (component-major); NULL = not wanted. */
int gsph_integrator_start_state(gsph_handle* h, double* r0, double* v0, double* a0, double* u0, double* dudt0, double* dt_next)
{
  if (!h) return GSPH_EINVAL;
  if (!h->integ.on) { h->err = "gsph_integrator_start_state before gsph_integrator_begin"; return GSPH_ESTATE; }
  const IntegState& I = h->integ;
  if (dt_next && I.cfg.Nlevels <= 1) { h->err = "gsph_integrator_start_state: dt_next exists with Nlevels > 1 only"; return GSPH_EINVAL; }
  cudaSetDevice(h->device);
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  const size_t N = (size_t) h->N;
  const int nd = h->cfg.ndim;
  double* const out3[3] = {r0, v0, a0};
  for (int g = 0; g < 3; g++)
    if (out3[g]) for (int k = 0; k < nd; k++)
      CUDA_TRY(h, cudaMemcpy(out3[g] + (size_t) k*N, I.s0[3*g + k], N*sizeof(double), cudaMemcpyDeviceToHost));
  if (u0) CUDA_TRY(h, cudaMemcpy(u0, I.s0[9], N*sizeof(double), cudaMemcpyDeviceToHost));
  if (dudt0) CUDA_TRY(h, cudaMemcpy(dudt0, I.s0[10], N*sizeof(double), cudaMemcpyDeviceToHost));
  if (dt_next) CUDA_TRY(h, cudaMemcpy(dt_next, I.dt_next, N*sizeof(double), cudaMemcpyDeviceToHost));
  return GSPH_OK;
}

int gsph_timer_start(gsph_handle* h)
{
  if (!h) return GSPH_EINVAL;
  cudaSetDevice(h->device);
  CUDA_TRY(h, cudaEventRecord(h->evt[0], h->stream));
  return GSPH_OK;
}

int gsph_timer_stop(gsph_handle* h, float* ms)
{
  if (!h || !ms) return GSPH_EINVAL;
  cudaSetDevice(h->device);
  CUDA_TRY(h, cudaEventRecord(h->evt[1], h->stream));
  CUDA_TRY(h, cudaEventSynchronize(h->evt[1]));
  CUDA_TRY(h, cudaEventElapsedTime(ms, h->evt[0], h->evt[1]));
  return GSPH_OK;
}

int gsph_get_stats(gsph_handle* h, gsph_stats* st)
{
  if (!h || !st) return GSPH_EINVAL;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  float* dst[6] = {&h->stats.ms_kernel_density, &h->stats.ms_kernel_hydro, &h->stats.ms_kernel_gravity,
                   &h->stats.ms_tree_sort, &h->stats.ms_tree_levels, &h->stats.ms_tree_stock};
  for (int k = 0; k < 6; k++)
    if (h->evk_used[k]) { cudaEventElapsedTime(dst[k], h->evk[k][0], h->evk[k][1]); h->evk_used[k] = false; }
  *st = h->stats;
  return GSPH_OK;
}

int gsph_kernel_table(int ndim, int kernel, int which, double* out)
{
  if (ndim < 1 || ndim > 3 || (kernel < GSPH_KERNEL_M4 || kernel > GSPH_KERNEL_GAUSSIAN) || which < 0 || which > 5 || !out)
    return GSPH_EINVAL;
  std::vector<double> tab(6*GSPH_TABLE_RES);
  gsph_make_tables(ndim, kernel, tab.data());
  memcpy(out, tab.data() + (size_t) which*GSPH_TABLE_RES, GSPH_TABLE_RES*sizeof(double));
  return GSPH_OK;
}

int gsph_measure_fp_peak(int device, double* fp64_tflops, double* fp32_tflops)
{
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return GSPH_ENODEVICE;
  if (device < 0 || device >= ndev) return GSPH_EINVAL;
  cudaSetDevice(device);
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, device);
  const int blocks = prop.multiProcessorCount*8, threads = 256, iters = 1 << 15;
  void* out = nullptr;
  if (cudaMalloc(&out, (size_t) blocks*threads*sizeof(double)) != cudaSuccess) return GSPH_ENOMEM;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms = 0;
  const double flops = 2.0*8.0*(double) iters*(double) blocks*threads;
  fma_peak_kernel<double><<<blocks, threads>>>((double*) out, 1024);
  cudaEventRecord(e0); fma_peak_kernel<double><<<blocks, threads>>>((double*) out, iters); cudaEventRecord(e1);
  cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
  if (fp64_tflops) *fp64_tflops = flops/(ms*1e-3)/1e12;
  fma_peak_kernel<float><<<blocks, threads>>>((float*) out, 1024);
  cudaEventRecord(e0); fma_peak_kernel<float><<<blocks, threads>>>((float*) out, iters); cudaEventRecord(e1);
  cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
  if (fp32_tflops) *fp32_tflops = flops/(ms*1e-3)/1e12;
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
  return cudaGetLastError() == cudaSuccess ? GSPH_OK : GSPH_ECUDA;
}

}  /* extern "C" */
