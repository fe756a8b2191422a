"""Synthetic and file tet-mesh inputs for the FEM step (host side, numpy only).

`truth_cube` restates VolMeshSamples::CreateTruthCube (reference src/deformable/VolMeshSamples.cpp:67-130):
node (i,j,k) sits at start + (i,j,k)*cellsize with start = (-nx/2, 0, -nz/2)*cellsize and index
i*ny*nz + j*nz + k; every cell is split into the same six tetrahedra in the same corner order.
`read_veg` reads the subset of the Vega .veg text format the reference's models use
(vegafem/volumetricMesh/volumetricMesh.cpp:45-..., *VERTICES / *ELEMENTS TET sections).
"""
from __future__ import annotations

import numpy as np

# CellCorners enum of VolMeshSamples.cpp:92: LBN, LBF, LTN, LTF, RBN, RBF, RTN, RTF
_LBN, _LBF, _LTN, _LTF, _RBN, _RBF, _RTN, _RTF = range(8)
# the six tets per cell, VolMeshSamples.cpp:110-115
_CELL_TETS = np.array(
    [
        [_LBN, _LTN, _RBN, _LBF],
        [_RTN, _LTN, _LBF, _RBN],
        [_RTN, _LTN, _LTF, _LBF],
        [_RTN, _RBN, _LBF, _RBF],
        [_RTN, _LBF, _LTF, _RBF],
        [_RTN, _LTF, _RTF, _RBF],
    ],
    dtype=np.int64,
)


def truth_cube(nx: int, ny: int | None = None, nz: int | None = None, cellsize: float = 0.2):
    """Return (verts float64 [nV,3], tets int32 [nT,4]) of CreateTruthCube(nx, ny, nz, cellsize)."""
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    if nx < 2 or ny < 2 or nz < 2:
        raise ValueError("Invalid input param to create a truth cube")
    start = np.array([-(float(nx)) / 2.0, 0.0, -(float(nz)) / 2.0]) * cellsize
    i, j, k = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    ijk = np.stack([i.ravel(), j.ravel(), k.ravel()], axis=1).astype(np.float64)
    # vec3d(i,j,k) * cellsize, then start + ...  (same operation order as the reference)
    verts = start[None, :] + ijk * cellsize

    ci, cj, ck = np.meshgrid(np.arange(nx - 1), np.arange(ny - 1), np.arange(nz - 1), indexing="ij")
    ci, cj, ck = ci.ravel().astype(np.int64), cj.ravel().astype(np.int64), ck.ravel().astype(np.int64)
    base = ci * ny * nz + cj * nz + ck
    corners = np.stack(
        [
            base,                      # LBN
            base + 1,                  # LBF
            base + nz,                 # LTN
            base + nz + 1,             # LTF
            base + ny * nz,            # RBN
            base + ny * nz + 1,        # RBF
            base + ny * nz + nz,       # RTN
            base + ny * nz + nz + 1,   # RTF
        ],
        axis=1,
    )
    tets = corners[:, _CELL_TETS].reshape(-1, 4)
    return np.ascontiguousarray(verts), np.ascontiguousarray(tets.astype(np.int32))


def cube_bottom_vertices(nx: int, ny: int | None = None, nz: int | None = None) -> np.ndarray:
    """Indices of all nodes with j == 0 (the y = 0 plane): the fixed set of SURVEY.md §8d."""
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    i, k = np.meshgrid(np.arange(nx), np.arange(nz), indexing="ij")
    return np.ascontiguousarray((i.ravel() * ny * nz + k.ravel()).astype(np.int32))


def cube_corner_vertex(nx: int, ny: int | None = None, nz: int | None = None) -> int:
    """Index of node (nx-1, ny-1, nz-1): where the pick-mode haptic load is applied."""
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    return (nx - 1) * ny * nz + (ny - 1) * nz + (nz - 1)


def warm_displacement(verts: np.ndarray, cellsize: float = 0.2) -> np.ndarray:
    """Deterministic warm state u_v = 0.05*cs*(sin x, cos y, sin z) of SURVEY.md §8d (R != I)."""
    u = np.empty_like(verts)
    u[:, 0] = np.sin(verts[:, 0])
    u[:, 1] = np.cos(verts[:, 1])
    u[:, 2] = np.sin(verts[:, 2])
    return np.ascontiguousarray(0.05 * cellsize * u)


def two_tetra():
    """VolMeshSamples::CreateTwoTetra (VolMeshSamples.cpp:41-65) — the mesh main.cpp:833 really builds."""
    verts = np.array([[-1, 0, 0], [1, 0, 0], [0, 0, -1], [0, 0, 1], [0, 2, 0]], dtype=np.float64)
    tets = np.array([[0, 2, 3, 4], [1, 2, 3, 4]], dtype=np.int32)
    return verts, tets


def one_tetra():
    """VolMeshSamples::CreateOneTetra (VolMeshSamples.cpp:15-39)."""
    verts = np.array([[-1, 0, 0], [0, 0, -2], [1, 0, 0], [0, 2, -1]], dtype=np.float64)
    tets = np.array([[0, 1, 2, 3]], dtype=np.int32)
    return verts, tets


def read_veg(path: str):
    """Parse *VERTICES and *ELEMENTS (TET) of a .veg file -> (verts [nV,3] f64, tets [nT,4] i32, 0-based)."""
    verts, tets = None, None
    with open(path, "r") as fh:
        lines = [ln.strip() for ln in fh]
    n = len(lines)
    p = 0

    def skip(p):
        while p < n and (not lines[p] or lines[p].startswith("#")):
            p += 1
        return p

    while p < n:
        ln = lines[p]
        if ln.startswith("*VERTICES"):
            p = skip(p + 1)
            hdr = lines[p].split()
            nv, dim = int(hdr[0]), int(hdr[1])
            assert dim == 3
            p += 1
            rows = []
            while len(rows) < nv:
                p = skip(p)
                rows.append(lines[p].split())
                p += 1
            arr = np.array(rows, dtype=np.float64)
            first = int(arr[0, 0])
            verts = np.ascontiguousarray(arr[:, 1:4])
            vbase = first
        elif ln.startswith("*ELEMENTS"):
            p = skip(p + 1)
            assert lines[p].upper().startswith("TET"), "only TET meshes are on the path"
            p = skip(p + 1)
            hdr = lines[p].split()
            ne, npe = int(hdr[0]), int(hdr[1])
            assert npe == 4
            p += 1
            rows = []
            while len(rows) < ne:
                p = skip(p)
                rows.append(lines[p].split())
                p += 1
            arr = np.array(rows, dtype=np.int64)
            tets = arr[:, 1:5]
        else:
            p += 1
    assert verts is not None and tets is not None
    tets = np.ascontiguousarray((tets - vbase).astype(np.int32))
    return verts, tets


def write_veg(path, verts, tets, materials=(), sets=None, regions=(), sep=" "):
    """Write a Vega .veg file (the format VolMeshIO::writeVega / PS_VegWriter produce for FemBrain's models).

    materials: [(name, density, E, nu)], sets: {name: iterable of 0-based element ids}, regions: [(set name, material name)];
    "allElements" is the implicit set of every element.  Vertices and elements are written 1-indexed like the reference's files."""
    with open(path, "w") as fh:
        fh.write("# Vega mesh file written by fembrain_b200.meshes.write_veg\n\n*VERTICES\n")
        fh.write(f"{len(verts)} 3 0 0\n")
        for i, p in enumerate(verts):
            fh.write(f"{i + 1}{sep}{float(p[0])!r}{sep}{float(p[1])!r}{sep}{float(p[2])!r}\n")
        fh.write("\n*ELEMENTS\nTET\n")
        fh.write(f"{len(tets)} 4 0\n")
        for i, t in enumerate(tets):
            fh.write(f"{i + 1}{sep}{int(t[0]) + 1}{sep}{int(t[1]) + 1}{sep}{int(t[2]) + 1}{sep}{int(t[3]) + 1}\n")
        for name, density, E, nu in materials:
            fh.write(f"\n*MATERIAL {name}\nENU, {float(density)!r}, {float(E)!r}, {float(nu)!r}\n")
        for name, els in (sets or {}).items():
            fh.write(f"\n*SET {name}\n")
            els = [int(e) + 1 for e in els]
            for k in range(0, len(els), 8):
                fh.write(", ".join(str(e) for e in els[k:k + 8]) + ",\n")
        for set_name, mat_name in regions:
            fh.write(f"\n*REGION\n{set_name}, {mat_name}\n")
