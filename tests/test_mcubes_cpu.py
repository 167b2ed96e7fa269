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
