"""Seeded synthetic configurations shared by the oracle-side and GPU-side tests (BASELINE.json configs,
SURVEY.md s8(d)).  Inputs only -- no reference data is read at run time."""
import numpy as np


def readme_points(m=1000, seed=1, noise=0.5):
    """README quick start (README.md:84-95): m uniform points on the unit square,
    z = sin(8x) + cos(8y) + noise * N(0,1)."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(0, 1, m)
    y = rng.uniform(0, 1, m)
    z = np.sin(8 * x) + np.cos(8 * y) + noise * rng.standard_normal(m)
    return np.column_stack([x, y]), z


def sphere_points(m=2000, seed=2):
    rng = np.random.default_rng(seed)
    lon = rng.uniform(-180, 180, m)
    lat = np.degrees(np.arcsin(rng.uniform(-1, 1, m)))
    z = 380 + 3 * np.sin(np.radians(lat) * 2) + np.cos(np.radians(lon)) + 0.5 * rng.standard_normal(m)
    return np.column_stack([lon, lat]), z


def pseudo_isea3h(seed=7):
    """Quasi-uniform centroid sets with the ISEA3H counts for resolutions 0..3 (12/32/92/272): Fibonacci
    lattices.  The real DGGRID centroids ship with the reference as data/isea3h.rda and cannot travel; the
    kernels only need *some* centroids on the sphere, identical for oracle and device."""
    out = {}
    for res, n in enumerate([12, 32, 92, 272]):
        k = np.arange(n) + 0.5
        lat = np.degrees(np.arcsin(1 - 2 * k / n))
        lon = (np.degrees(k * np.pi * (3 - np.sqrt(5))) + 180.0) % 360.0 - 180.0
        out[res] = np.column_stack([lon, lat])
    return out


def areal_tiles(tile=0.1, noise=0.1, seed=51):
    """Square supports tiling the unit square (edges kept off the BAU centroids), one tile row left out; each observes the average of a
    smooth field plus noise.  Returns (ptr, vx, vy, z, std)."""
    rng = np.random.default_rng(seed)
    ptr, vx, vy, z = [0], [], [], []
    nt = int(round(1.0 / tile))
    for iy in range(nt):
        for ix in range(nt):
            if iy == nt // 2 - 1:
                continue
            x0, y0 = tile * ix + 0.003, tile * iy + 0.003
            vx += [x0, x0 + tile, x0 + tile, x0, x0]
            vy += [y0, y0, y0 + tile, y0 + tile, y0]
            ptr.append(len(vx))
            xm, ym = x0 + tile / 2, y0 + tile / 2
            z.append(np.sin(6 * xm) + np.cos(5 * ym) + 0.5 * (xm + ym) + noise * rng.standard_normal())
    m = len(z)
    std = 0.1 + 0.05 * rng.uniform(size=m)
    return np.array(ptr), np.array(vx), np.array(vy), np.array(z), std
