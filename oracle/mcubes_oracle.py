"""CPU restatement (numpy loops) of the device marching cubes (csrc/mcubes.cu) for small lattices: same generated case
table (ho-nerf_b200/mcubes_tables.py), same vertex order (axis by axis, lattice order) and triangle order (cell by cell), so
the CUDA path can be compared index for index.  TEST INFRASTRUCTURE ONLY.  The algorithm it restates is the reference's
call mcubes.marching_cubes(u, threshold) (utils/renderer.py:279); PyMCubes itself is an un-vendored dependency
(requirements.txt: PyMCubes==0.1.4), not installed here: parity unpinned."""
import importlib.util
import os

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _tables():
    spec = importlib.util.spec_from_file_location("mcubes_tables", os.path.join(_ROOT, "ho-nerf_b200", "mcubes_tables.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def marching_cubes(u, iso=0.0):
    T = _tables()
    n_tris, tris, owner = T.build_tables()
    u = np.asarray(u, dtype=np.float32)
    nx, ny, nz = u.shape
    inside = u < np.float32(iso)
    flags = np.zeros((3, nx, ny, nz), dtype=np.int32)
    flags[0, :-1] = inside[:-1] != inside[1:]
    flags[1, :, :-1] = inside[:, :-1] != inside[:, 1:]
    flags[2, :, :, :-1] = inside[:, :, :-1] != inside[:, :, 1:]
    slot = np.cumsum(flags.reshape(-1)).reshape(flags.shape) - 1
    verts = np.zeros((int(flags.sum()), 3), dtype=np.float32)
    for axis in range(3):
        for (i, j, k) in np.argwhere(flags[axis]):
            p1 = [i, j, k]
            p1[axis] += 1
            v0, v1 = u[i, j, k], u[tuple(p1)]
            t = np.float32((np.float64(np.float32(iso)) - np.float64(v0)) / (np.float64(v1) - np.float64(v0)))
            pos = np.array([i, j, k], dtype=np.float32)
            pos[axis] += t
            verts[slot[axis, i, j, k]] = pos
    out = []
    for i in range(nx - 1):
        for j in range(ny - 1):
            for k in range(nz - 1):
                cs = 0
                for c, (dx, dy, dz) in enumerate(T.CORNERS):
                    cs |= int(inside[i + dx, j + dy, k + dz]) << c
                for q in range(n_tris[cs]):
                    tri = []
                    for e in tris[cs][3 * q: 3 * q + 3]:
                        dx, dy, dz, axis = owner[e]
                        tri.append(slot[axis, i + dx, j + dy, k + dz])
                    out.append(tri)
    return verts, np.asarray(out, dtype=np.int32).reshape(-1, 3)
