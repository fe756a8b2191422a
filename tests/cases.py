"""Seeded input cases shared by the parity tests (sizes the CPU oracle finishes in seconds)."""
import os

import numpy as np

from fembrain_b200 import meshes

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def cube_case(nx, ny=None, nz=None):
    v, t = meshes.truth_cube(nx, ny, nz)
    fixed = meshes.cube_bottom_vertices(nx, ny, nz)
    load = meshes.cube_corner_vertex(nx, ny, nz)
    return v, t, fixed, load


def golden_mesh(name):
    """Meshes taken from the reference's data directory, committed as fixtures by tests/golden/make_golden.py."""
    z = np.load(os.path.join(GOLDEN, f"mesh_{name}.npz"))
    return z["verts"], z["tets"], z["fixed"]


def point_load(r, vertex, f=(1e4, 0.0, 0.0)):
    out = np.zeros(r)
    out[3 * vertex:3 * vertex + 3] = f
    return out


def perturbation(verts, scale=1.0, seed=0):
    rng = np.random.default_rng(seed)
    u = meshes.warm_displacement(verts) * 5.0 * scale + 0.02 * scale * rng.standard_normal(verts.shape)
    return u.reshape(-1)


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    den = np.abs(b).max() if b.size else 0.0
    return 0.0 if a.size == 0 else float(np.abs(a - b).max() / (den if den > 0 else 1.0))
