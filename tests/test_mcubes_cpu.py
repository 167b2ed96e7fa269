"""The generated marching-cubes case table (ho-nerf_b200/mcubes_tables.py), checked on the CPU: every sign configuration's
iso lines close, its triangles use exactly the crossing edges, complementary configurations use the same edges, and the
boundary of each cell's patch is exactly the set of face segments -- which depend on the face's corner signs only, so two
cells sharing a face always agree (no cracks)."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _tables():
    spec = importlib.util.spec_from_file_location("mcubes_tables", os.path.join(ROOT, "ho-nerf_b200", "mcubes_tables.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_case_table_is_consistent():
    T = _tables()
    assert T.check_tables() == 5                    # at most five triangles per cell, like the classic table
    n_tris, tris, owner = T.build_tables()
    assert n_tris[0] == 0 and n_tris[255] == 0 and n_tris[1] == 1 and n_tris[0b00001111] == 2
    assert sum(n_tris) == 820
    # every edge is owned by its lower corner along its axis
    for (a, b), (dx, dy, dz, axis) in zip(T.EDGES, owner):
        pa, pb = T.CORNERS[a], T.CORNERS[b]
        assert (dx, dy, dz) == tuple(min(x, y) for x, y in zip(pa, pb)) and pa[axis] != pb[axis]


def test_face_rule_depends_on_the_face_only():
    """Two cells that share a face see the same four corner signs on it; the segments drawn on that face must be the same
    (as undirected edge pairs, in the face's own edge numbering) whatever the other four corners of either cell are."""
    T = _tables()
    faces = T._ccw_faces()
    for fi, q in enumerate(faces):
        seen = {}
        edges = [T._EDGE_ID[frozenset((q[i], q[(i + 1) % 4]))] for i in range(4)]
        for case in range(256):
            key = tuple((case >> c) & 1 for c in q)
            segs = frozenset(frozenset((a, b)) for a, b in T._segments(case) if a in edges and b in edges and
                             _same_face(T, a, b, q))
            assert seen.setdefault(key, segs) == segs, (fi, case)


def _same_face(T, a, b, q):
    ca, cb = set(T.EDGES[a]), set(T.EDGES[b])
    return ca <= set(q) and cb <= set(q)


def test_no_triangle_lies_in_a_face():
    """A triangle with all three vertices on one face of the cell would be mirrored by the neighbouring cell (double sheet)."""
    T = _tables()
    n_tris, tris, _ = T.build_tables()
    faces = T._face_edge_sets()
    for case in range(256):
        for q in range(n_tris[case]):
            tri = set(tris[case][3 * q: 3 * q + 3])
            assert not any(tri <= f for f in faces), case


def _edge_stats(tri):
    d = {}
    for a, b, c in tri:
        for x, y in ((a, b), (b, c), (c, a)):
            d[(x, y)] = d.get((x, y), 0) + 1
    return d


def test_oracle_mesh_of_a_sphere_is_a_closed_oriented_manifold():
    """oracle/mcubes_oracle.py is the checker of the device mesh extraction; it is pinned here by the properties of a correct
    shared-vertex mesh (PyMCubes, the reference's dependency, is not installed: parity unpinned)."""
    import numpy as np
    import mcubes_oracle as MO
    res = 18
    ax = np.linspace(-1.0, 1.0, res)
    xx, yy, zz = np.meshgrid(ax, ax * 1.1, ax * 0.9, indexing="ij")
    u = (np.sqrt(xx * xx + yy * yy + zz * zz) - 0.6).astype(np.float32)
    v, t = MO.marching_cubes(u, 0.0)
    d = _edge_stats(t)
    assert all(n == 1 for n in d.values()) and all((b, a) in d for (a, b) in d)
    assert len(v) - len(d) // 2 + len(t) == 2
    # on the zero set to interpolation accuracy, normals towards the centre (lower values)
    w = v / (res - 1) * 2.0 - 1.0
    r = np.sqrt(w[:, 0] ** 2 + (1.1 * w[:, 1]) ** 2 + (0.9 * w[:, 2]) ** 2)
    assert np.abs(r - 0.6).max() < 0.02
    p0, p1, p2 = v[t[:, 0]], v[t[:, 1]], v[t[:, 2]]
    nrm = np.cross(p1 - p0, p2 - p0)
    assert ((nrm * ((p0 + p1 + p2) / 3.0 - (res - 1) / 2.0)).sum(1) < 0).all()


def test_oracle_mesh_of_a_field_with_ambiguous_faces_has_no_cracks():
    import numpy as np
    import mcubes_oracle as MO
    rng = np.random.default_rng(3)
    res = 14
    ax = np.linspace(0, 1, res)
    xx, yy, zz = np.meshgrid(ax, ax, ax, indexing="ij")
    u = np.zeros((res, res, res))
    for _ in range(12):
        k = rng.integers(1, 5, 3)
        ph = rng.random(3) * 6.28
        u += rng.standard_normal() * np.sin(6.28 * k[0] * xx + ph[0]) * np.sin(6.28 * k[1] * yy + ph[1]) * np.sin(6.28 * k[2] * zz + ph[2])
    v, t = MO.marching_cubes(u.astype(np.float32), 0.0)
    v = v.astype(np.float64)
    assert len(t) > 500
    d = _edge_stats(t)
    assert all(n == 1 for n in d.values())
    on_boundary = lambda p: bool(((p < 1e-9) | (p > res - 1 - 1e-9)).any())
    assert all(on_boundary(v[a]) and on_boundary(v[b]) for (a, b) in d if (b, a) not in d)
