"""Synthetic particle sets of the BASELINE.json configurations: generated on the
GPU with torch (plumbing only) so that 16M-64M particles take seconds, returned as numpy arrays.

 cloud : "turbulent self-gravitating molecular cloud": uniform-density sphere (jittered cubic
         lattice clipped to R = 1, which also removes coordinate ties), M = 1, divergence-free
         random velocity field with P(k) ~ k^-4 built from `nmodes` solenoidal Fourier modes,
         scaled to E_turb/|E_grav| = alpha_turb; isothermal-ish u.  Open boundaries, self-gravity.
 glass : 3-D periodic unit box, jittered lattice r = lattice + 0.2*dx*U(-1,1), m = 1/N, u = 1.5,
         hydro only.
"""
import numpy as np


def _torch():
    import torch
    return torch


def ref_perturb_factor(iorig, amp):
    """(1 + amp*(2f-1)) with the integer hash of oracle/ref_harness.cpp perturb_h, so that both
    bench arms start the h-rho iteration from the same perturbed guess."""
    x = (iorig.astype(np.uint64) * np.uint64(2654435761)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(2246822519)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(13)
    f = (x & np.uint64(0xFFFFFF)).astype(np.float64) / float(0x1000000)
    return 1.0 + amp * (2.0 * f - 1.0)


def cloud(N_target, seed=2024, alpha_turb=0.5, nmodes=48, device="cuda", jitter=0.25):
    torch = _torch()
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    # lattice spacing so that the unit sphere holds ~N_target points
    n = int(np.ceil((N_target * 6.0 / np.pi) ** (1.0 / 3.0))) + 2
    dx = 2.0 / n
    ax = (torch.arange(n, device=device, dtype=torch.float64) + 0.5) * dx - 1.0
    chunks = []
    for i0 in range(0, n, 64):          # slabs keep the temporary grids small
        xs = ax[i0:i0 + 64]
        X, Y, Z = torch.meshgrid(xs, ax, ax, indexing="ij")
        P = torch.stack([X.reshape(-1), Y.reshape(-1), Z.reshape(-1)], 0)
        P = P + jitter * dx * (2.0 * torch.rand(P.shape, device=device, dtype=torch.float64, generator=g) - 1.0)
        keep = (P * P).sum(0) < 1.0
        chunks.append(P[:, keep])
    r = torch.cat(chunks, 1)
    if r.shape[1] > N_target:           # drop the outermost points to hit N_target exactly
        rad = (r * r).sum(0)
        idx = torch.argsort(rad)[:N_target]
        idx, _ = torch.sort(idx)
        r = r[:, idx]
    N = r.shape[1]
    # solenoidal random field: v = sum_k A_k (e_k x khat) cos(k.x + phi), |A_k|^2 ~ k^-4
    rng = np.random.default_rng(seed)
    v = torch.zeros_like(r)
    for _ in range(nmodes):
        kmag = rng.uniform(1.0, 8.0)
        kd = rng.normal(size=3)
        kd /= np.linalg.norm(kd)
        e = np.cross(kd, rng.normal(size=3))
        e /= np.linalg.norm(e)
        amp = kmag ** -2.0 * rng.normal()
        k = torch.tensor(kd * kmag * np.pi, device=device, dtype=torch.float64)
        phase = (k[:, None] * r).sum(0) + rng.uniform(0, 2 * np.pi)
        v += torch.tensor(e * amp, device=device, dtype=torch.float64)[:, None] * torch.cos(phase)[None, :]
    m = torch.full((N,), 1.0 / N, device=device, dtype=torch.float64)
    egrav = 0.6                          # 3/5 G M^2 / R with G = M = R = 1
    ekin = 0.5 * (m[None, :] * v * v).sum().item()
    v *= np.sqrt(alpha_turb * egrav / max(ekin, 1e-300))
    u = torch.full((N,), 0.05, device=device, dtype=torch.float64)
    h = torch.full((N,), 1.2 * (4.0 * np.pi / 3.0 / N) ** (1.0 / 3.0), device=device, dtype=torch.float64)
    out = dict(r=r, v=v, m=m, u=u, h=h)
    params = {"ndim": 3, "self_gravity": 1, "hydro_forces": 1, "kernel": "m4", "tabulated_kernel": 1,
              "multipole": "fast_monopole", "thetamaxsqd": 0.15, "gravity_mac": "geometric",
              "gas_eos": "isothermal", "temp0": 0.05 * (1.6666666666666667 - 1.0), "mu_bar": 1.0,
              "gamma_eos": 1.6666666666666667, "Nleafmax": 6, "avisc": "mon97", "acond": "none"}
    return params, out


def glass(N_target, seed=12345, device="cuda", jitter=0.2):
    torch = _torch()
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    n = int(round(N_target ** (1.0 / 3.0)))
    dx = 1.0 / n
    ax = (torch.arange(n, device=device, dtype=torch.float64) + 0.5) * dx
    X, Y, Z = torch.meshgrid(ax, ax, ax, indexing="ij")
    r = torch.stack([X.reshape(-1), Y.reshape(-1), Z.reshape(-1)], 0)
    r = (r + jitter * dx * (2.0 * torch.rand(r.shape, device=device, dtype=torch.float64, generator=g) - 1.0)) % 1.0
    N = r.shape[1]
    v = 0.1 * torch.sin(2 * np.pi * r.flip(0))
    m = torch.full((N,), 1.0 / N, device=device, dtype=torch.float64)
    u = torch.full((N,), 1.5, device=device, dtype=torch.float64)
    h = torch.full((N,), 1.2 * dx, device=device, dtype=torch.float64)
    params = {"ndim": 3, "self_gravity": 0, "hydro_forces": 1, "kernel": "m4", "tabulated_kernel": 1,
              "gas_eos": "energy_eqn", "gamma_eos": 1.6666666666666667, "Nleafmax": 6, "avisc": "mon97"}
    for k in range(3):
        params.update({"boxmin[%d]" % k: 0.0, "boxmax[%d]" % k: 1.0, "boundary_lhs[%d]" % k: "periodic",
                       "boundary_rhs[%d]" % k: "periodic"})
    return params, dict(r=r, v=v, m=m, u=u, h=h)


def khi(N_target=983552, device="cuda", **_):
    """BASELINE config 2 (tests/hydro_tests/khi.dat scaled up): 2-D periodic [-0.5,0.5]^2, two cubic
    lattices rho = 1 (|y| > 0.25, v_x = -0.5... ) / rho = 2 (|y| < 0.25), equal pressure, v_x = -/+0.5, velocity
    perturbation amp 0.1 lambda 0.5; quintic kernel tabulated, Nleafmax = 12, wadsley2008, alpha 0.1 beta 0.2.
    Deterministic (no RNG)."""
    torch = _torch()
    # lattice 1 (rho 1) fills the box half |y| >= 0.25, lattice 2 (rho 2) the half |y| < 0.25: equal particle mass
    # -> lattice 2 has sqrt(2) finer spacing; sizes like the SURVEY's 800x400 + 1152x576 for ~983k
    scale = (N_target/983552.0)**0.5
    nx1, ny1 = max(8, int(round(800*scale))), max(4, int(round(400*scale)))
    nx2, ny2 = max(8, int(round(1152*scale))), max(4, int(round(576*scale)))
    def lattice(nx, ny, y0, y1):
        ax = (torch.arange(nx, device=device, dtype=torch.float64) + 0.5)/nx - 0.5
        ay = y0 + (torch.arange(ny, device=device, dtype=torch.float64) + 0.5)/ny*(y1 - y0)
        X, Y = torch.meshgrid(ax, ay, indexing="ij")
        return torch.stack([X.reshape(-1), Y.reshape(-1)], 0)
    top = lattice(nx1, ny1//2, 0.25, 0.5); bot = lattice(nx1, ny1 - ny1//2, -0.5, -0.25)
    mid = lattice(nx2, ny2, -0.25, 0.25)
    r = torch.cat([bot, mid, top], 1)
    n1, n2 = bot.shape[1] + top.shape[1], mid.shape[1]
    N = r.shape[1]
    inner = r[1].abs() < 0.25
    m = torch.where(inner, torch.full((N,), 2.0*0.5/n2, device=device, dtype=torch.float64),
                    torch.full((N,), 1.0*0.5/n1, device=device, dtype=torch.float64))
    rho = torch.where(inner, 2.0, 1.0).to(torch.float64)
    vx = torch.where(inner, 0.5, -0.5).to(torch.float64)
    vy = 0.1*torch.sin(2*np.pi*r[0]/0.5)*(torch.exp(-(r[1] - 0.25)**2/0.005) + torch.exp(-(r[1] + 0.25)**2/0.005))
    v = torch.stack([vx, vy], 0)
    press, gamma = 2.5, 1.6666666666666667
    u = press/(rho*(gamma - 1.0))
    h = 1.2*torch.sqrt(m/rho)
    params = {"ndim": 2, "self_gravity": 0, "hydro_forces": 1, "kernel": "quintic", "tabulated_kernel": 1,
              "gas_eos": "energy_eqn", "gamma_eos": gamma, "Nleafmax": 12, "avisc": "mon97", "acond": "wadsley2008",
              "alpha_visc": 0.1, "beta_visc": 0.2}
    for k in range(2):
        params.update({"boxmin[%d]" % k: -0.5, "boxmax[%d]" % k: 0.5, "boundary_lhs[%d]" % k: "periodic",
                       "boundary_rhs[%d]" % k: "periodic"})
    return params, dict(r=r, v=v, m=m, u=u, h=h)


def bossbodenheimer(N_target=4000000, device="cuda", seed=7, **_):
    """BASELINE config 3 (tests/astro_tests/bossbodenheimer.dat at 4 M): uniform sphere in solid-body
    rotation with an m = 2 azimuthal perturbation of amplitude 0.5 (here carried by the particle masses), barotropic
    EOS, M4 tabulated, self-gravity with fast_monopole, thetamaxsqd = 0.15, Nleafmax = 6; Nlevels = 1, no sinks.
    Dimensionless: M = R = G = 1, ratio of rotational to gravitational energy 0.16 like the original."""
    torch = _torch()
    params, ic = cloud(N_target, seed=seed, alpha_turb=0.0, nmodes=0, device=device, jitter=0.25)
    r = ic["r"]
    N = r.shape[1]
    phi = torch.atan2(r[1], r[0])
    m = (1.0 + 0.5*torch.cos(2.0*phi))/N
    m = m/m.sum()
    omega = np.sqrt(0.16*0.6/0.2)                         # E_rot = (1/5) M R^2 Omega^2 = 0.16 * (3/5) G M^2 / R
    v = torch.stack([-omega*r[1], omega*r[0], torch.zeros_like(r[0])], 0)
    params = dict(params, gas_eos="barotropic", rho_bary=1.0e3, temp0=0.01, mu_bar=2.35, gamma_eos=1.4, Nleafmax=6)
    return params, dict(r=r, v=v, m=m, u=ic["u"], h=ic["h"])


WORKLOADS = {"cloud": cloud, "glass": glass, "khi": khi, "bossbodenheimer": bossbodenheimer}
