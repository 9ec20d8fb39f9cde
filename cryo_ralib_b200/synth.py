"""Deterministic synthetic particles: noisy analytic projections of a seeded
random Gaussian-blob density.  A 3-D isotropic Gaussian projects to a 2-D Gaussian, and 2-D
Gaussians are separable, so every image is Gy^T diag(a) Gx: one small batched matmul, done
with numpy on the CPU (tests) or with torch on a device (bench: plumbing only)."""
import math

import numpy as np


def make_density(nx, nblobs=64, seed=1234):
    rng = np.random.default_rng(seed)
    r = 0.30 * nx * np.cbrt(rng.uniform(0, 1, nblobs))
    v = rng.standard_normal((nblobs, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    pos = v * r[:, None]
    sigma = rng.uniform(1.5, 4.0, nblobs)
    amp = rng.uniform(0.5, 1.5, nblobs)
    return pos, sigma, amp


def random_rotations(n, rng):
    q = rng.standard_normal((n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    a, b, c, d = q.T
    return np.stack([np.stack([a*a+b*b-c*c-d*d, 2*(b*c-a*d), 2*(b*d+a*c)], 1),
                     np.stack([2*(b*c+a*d), a*a-b*b+c*c-d*d, 2*(c*d-a*b)], 1),
                     np.stack([2*(b*d-a*c), 2*(c*d+a*b), a*a-b*b-c*c+d*d], 1)], 1)


def render(nx, cx, cy, sigma, amp, xp=np, chunk=4096):
    """Images [n][nx][nx] from blob centres cx,cy [n][B] (pixels relative to nx//2)."""
    n = cx.shape[0]
    if xp is np:
        grid = np.arange(nx, dtype=cx.dtype) - (nx // 2)
    else:
        grid = xp.arange(nx, dtype=cx.dtype, device=cx.device) - (nx // 2)
    out = []
    for s in range(0, n, chunk):
        gx = xp.exp(-(grid[None, None, :] - cx[s:s+chunk, :, None]) ** 2 / (2 * sigma[None, :, None] ** 2))
        gy = xp.exp(-(grid[None, None, :] - cy[s:s+chunk, :, None]) ** 2 / (2 * sigma[None, :, None] ** 2))
        gy = gy * (amp * sigma * math.sqrt(2 * math.pi))[None, :, None]
        if xp is np:
            out.append(np.einsum("nby,nbx->nyx", gy, gx))
        else:
            out.append(xp.einsum("nby,nbx->nyx", gy, gx))
    return xp.concatenate(out, 0) if len(out) > 1 else out[0]


def make_particles(P, nx, nviews, max_shift=3, snr=0.5, seed=2024, density_seed=1234, device=None):
    """Returns (images float32 [P][nx][nx], truth dict).  device: None -> numpy; a torch device
    string -> images are generated there and returned as a torch tensor."""
    pos, sigma, amp = make_density(nx, seed=density_seed)
    rng = np.random.default_rng(seed)
    rots = random_rotations(nviews, rng)
    view = rng.integers(0, nviews, P)
    psi = rng.uniform(0, 360, P)
    shift = rng.integers(-max_shift, max_shift + 1, (P, 2)).astype(np.float64)
    mirror = rng.integers(0, 2, P)
    proj = np.einsum("vij,bj->vbi", rots, pos)[:, :, :2]          # [V][B][2]
    p = proj[view]                                                 # [P][B][2]
    c, s = np.cos(np.radians(psi))[:, None], np.sin(np.radians(psi))[:, None]
    x = p[:, :, 0] * c - p[:, :, 1] * s
    y = p[:, :, 0] * s + p[:, :, 1] * c
    x = np.where(mirror[:, None] == 1, -x, x)
    x = x + shift[:, 0:1]
    y = y + shift[:, 1:2]
    truth = dict(view=view, psi=psi, shift=shift, mirror=mirror)
    if device is None:
        img = render(nx, x, y, sigma, amp)
        sig_var = img.var(axis=(1, 2)).mean()
        noise = rng.standard_normal(img.shape) * math.sqrt(sig_var / snr)
        return (img + noise).astype(np.float32), truth
    import torch
    dev = torch.device(device)
    tx = torch.as_tensor(x, dtype=torch.float32, device=dev)
    ty = torch.as_tensor(y, dtype=torch.float32, device=dev)
    ts = torch.as_tensor(sigma, dtype=torch.float32, device=dev)
    ta = torch.as_tensor(amp, dtype=torch.float32, device=dev)
    img = render(nx, tx, ty, ts, ta, xp=torch, chunk=2048)
    sig_var = img.var(dim=(1, 2)).mean()
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    img = img + torch.randn(img.shape, generator=g, device=dev, dtype=torch.float32) * torch.sqrt(sig_var / snr)
    return img, truth


def initial_references(images, R, per_ref=200, seed=99):
    """Mean of R random disjoint subsets of raw particles (generate_random_averages,
    notebook/00 cell 0)."""
    P = images.shape[0]
    per_ref = max(1, min(per_ref, P // R))
    rng = np.random.default_rng(seed)
    idx = rng.permutation(P)[:R * per_ref].reshape(R, per_ref)
    if isinstance(images, np.ndarray):
        return np.stack([images[np.sort(i)].mean(axis=0) for i in idx]).astype(np.float32)
    import torch
    return torch.stack([images[torch.as_tensor(np.sort(i), device=images.device)].mean(dim=0) for i in idx])
