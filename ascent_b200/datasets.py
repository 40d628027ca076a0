"""Synthetic Blueprint-shaped inputs for tests and bench (no Conduit in this image).

``braid``  : Conduit ``blueprint::mesh::examples::braid("uniform"|"rectilinear", nx, ny, nz)``
             [Conduit, recalled]; corroborated in tree by
             src/examples/tutorial/ascent_intro/cpp/blueprint_example3.cpp:61-86.
``radial`` : ``create_3d_example_dataset`` (src/tests/t_utils.hpp:335-462), field radial_vert.
A "domain" here is a dict with the keys the host mirror's ``DataSet.add_domain`` takes:
kind, dims (POINT dims, x fastest), origin/spacing or axes, field (flat numpy), bounds.
"""
import math

import numpy as np

PI_VALUE = 3.14159265359  # Conduit's PI_VALUE


def braid_values(nx, ny, nz, i0=0, j0=0, k0=0, gx=None, gy=None, gz=None, dtype=np.float64):
    """Braid vertex field of a (nx,ny,nz)-point window starting at global point (i0,j0,k0) of a
    (gx,gy,gz)-point global grid (defaults: the window is the whole grid).  x fastest."""
    gx, gy, gz = gx or nx, gy or ny, gz or nz
    dx = float(np.float32(4.0 * PI_VALUE)) / float(gx - 1)
    dy = float(np.float32(2.0 * PI_VALUE)) / float(gy - 1)
    dz = float(np.float32(3.0 * PI_VALUE)) / float(gz - 1)
    cx = (np.arange(i0, i0 + nx, dtype=np.float64) * dx) + (2.0 * PI_VALUE)
    cy = (np.arange(j0, j0 + ny, dtype=np.float64) * dy) - PI_VALUE
    out = np.empty((nz, ny, nx), dtype)
    CX, CY = np.meshgrid(cx, cy)  # (ny, nx)
    base = np.sin(CX) + np.sin(CY) + 2 * np.cos(np.sqrt((CX * CX) / 2.0 + CY * CY) / .75) \
        + 4 * np.cos(CX * CY / 4.0)
    r2 = CX * CX + CY * CY
    for k in range(nz):
        cz = ((k0 + k) * dz) - (1.5 * PI_VALUE)
        out[k] = base + math.sin(cz) + 1.5 * np.cos(np.sqrt(r2 + cz * cz) / .75)
    return out.reshape(-1)


def braid_uniform(nx, ny=None, nz=None, dtype=np.float64):
    """One uniform domain on [-10,10]^3 (origin -10, spacing 20/(n-1))."""
    ny, nz = ny or nx, nz or nx
    dims = (nx, ny, nz)
    spacing = [20.0 / (n - 1) for n in dims]
    return dict(kind="uniform", dims=dims, origin=[-10.0] * 3, spacing=spacing,
                field=braid_values(nx, ny, nz, dtype=dtype), assoc="point")


def braid_uniform_blocks(n_block, blocks_per_axis, dtype=np.float32):
    """blocks_per_axis^3 uniform domains of n_block^3 points sharing one point layer
    (global (b*(n-1)+1)^3 points on [-10,10]^3): configs c3 (512, 2) and c5 (128, 8)."""
    b = blocks_per_axis
    g = b * (n_block - 1) + 1
    sp = 20.0 / (g - 1)
    doms = []
    for kz in range(b):
        for jy in range(b):
            for ix in range(b):
                i0, j0, k0 = ix * (n_block - 1), jy * (n_block - 1), kz * (n_block - 1)
                origin = [-10.0 + i0 * sp, -10.0 + j0 * sp, -10.0 + k0 * sp]
                doms.append(dict(kind="uniform", dims=(n_block,) * 3, origin=origin,
                                 spacing=[sp] * 3, assoc="point",
                                 field=braid_values(n_block, n_block, n_block, i0, j0, k0, g, g, g,
                                                    dtype=dtype)))
    return doms


def warp_axis(n, power=1.5):
    """Monotone non-uniform axis on [-10, 10] (config c4): -10 + 20*(i/(n-1))^power."""
    t = np.arange(n, dtype=np.float64) / float(n - 1)
    return -10.0 + 20.0 * t ** power


def braid_rectilinear(nx, ny=None, nz=None, power=1.0, dtype=np.float64):
    """Rectilinear braid; power=1 reproduces Conduit's explicit uniform axes, power!=1 is the
    c4 warp (field values stay those of the index-space braid)."""
    ny, nz = ny or nx, nz or nx
    axes = [warp_axis(n, power) for n in (nx, ny, nz)]
    return dict(kind="rectilinear", dims=(nx, ny, nz), axes=axes,
                field=braid_values(nx, ny, nz, dtype=dtype), assoc="point")


def radial_example(cell_dim, par_rank, par_size):
    """create_3d_example_dataset(data, cell_dim, rank, size): rectilinear, f64."""
    size = par_size * cell_dim
    nx, ny, nz = size // par_size, size, size
    start = 0.0 - float(size) / 2.0
    rank_offset = start + float(np.float32(par_rank * nx))
    x = rank_offset + np.arange(nx + 1, dtype=np.float64)
    y = start + np.arange(ny + 1, dtype=np.float64)
    z = start / 2.0 + np.arange(nz + 1, dtype=np.float64)
    Z, Y, X = np.meshgrid(z, y, x, indexing="ij")
    vert = 10.0 * np.sqrt(X * X + Y * Y + Z * Z)
    Zc, Yc, Xc = np.meshgrid(z[:-1], y[:-1], x[:-1], indexing="ij")
    ele = 10.0 * np.sqrt(Xc * Xc + Yc * Yc + Zc * Zc)
    return dict(kind="rectilinear", dims=(nx + 1, ny + 1, nz + 1), axes=[x, y, z],
                field=vert.reshape(-1), field_ele=ele.reshape(-1), assoc="point")


def domain_bounds(dom):
    """f64 coordinate bounds (xmin,xmax,ymin,ymax,zmin,zmax) as vtkm::cont::CoordinateSystem::
    GetBounds reports them: uniform origin/spacing are f32 in Ascent
    (ascent_vtkh_data_adapter.cpp:1264-1270), rectilinear axes f64 (:1336-1395)."""
    b = np.zeros(6, np.float64)
    for a in range(3):
        if dom["kind"] == "uniform":
            o = float(np.float32(dom["origin"][a]))
            s = float(np.float32(dom["spacing"][a]))
            b[2 * a], b[2 * a + 1] = o, o + s * float(dom["dims"][a] - 1)
        else:
            b[2 * a], b[2 * a + 1] = dom["axes"][a][0], dom["axes"][a][-1]
    return b


def union_bounds(bounds_list):
    bl = np.asarray(bounds_list, np.float64).reshape(-1, 6)
    out = np.zeros(6, np.float64)
    out[0::2] = bl[:, 0::2].min(axis=0)
    out[1::2] = bl[:, 1::2].max(axis=0)
    return out


def structured_to_hexes(dims, origin, spacing, drop_cells=()):
    """Explicit hexahedra (VTK vertex order) of a uniform grid: (points [n,3] f32, conn [n_cells,8] int32).
    drop_cells: flat cell ids (x fastest) to leave out -- what vtk-h's ghost stripper does to a ragged ghost field."""
    nx, ny, nz = [int(d) for d in dims]
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    pts = np.stack([np.float32(origin[0]) + np.float32(spacing[0]) * i.astype(np.float32),
                    np.float32(origin[1]) + np.float32(spacing[1]) * j.astype(np.float32),
                    np.float32(origin[2]) + np.float32(spacing[2]) * k.astype(np.float32)], -1).reshape(-1, 3)
    ck, cj, ci = np.meshgrid(np.arange(nz - 1), np.arange(ny - 1), np.arange(nx - 1), indexing="ij")
    i0 = ((ck * ny + cj) * nx + ci).reshape(-1)
    conn = np.stack([i0, i0 + 1, i0 + 1 + nx, i0 + nx, i0 + nx * ny, i0 + nx * ny + 1, i0 + nx * ny + 1 + nx,
                     i0 + nx * ny + nx], 1).astype(np.int32)
    if len(drop_cells):
        keep = np.ones(conn.shape[0], bool)
        keep[np.asarray(drop_cells, int)] = False
        conn = conn[keep]
    return pts.astype(np.float32), conn


def hexes_to_tets(conn):
    """Six tetrahedra per hexahedron around the 0-6 diagonal (conforming across faces of a structured mesh)."""
    c = np.asarray(conn)
    order = [(0, 1, 2, 6), (0, 2, 3, 6), (0, 3, 7, 6), (0, 7, 4, 6), (0, 4, 5, 6), (0, 5, 1, 6)]
    return np.concatenate([c[:, list(o)] for o in order], 0).astype(np.int32)
