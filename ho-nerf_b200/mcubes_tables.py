"""Marching-cubes case table for the device mesh extraction (csrc/mcubes.cu), GENERATED, not transcribed.

The reference calls PyMCubes (``mcubes.marching_cubes(u, threshold)``, utils/renderer.py:279,561; utils/renderer_batch.py:309;
requirements.txt pins PyMCubes==0.1.4, which is not part of the reference tree and not installed here: parity unpinned).
Instead of copying a 256 x 16 table from memory, the triangulation of every sign configuration is derived here from first
principles:

* corner / edge numbering of the classic algorithm (corner i at (i&1 ^ (i>>1)&1, (i>>1)&1, (i>>2)&1), i.e. 0..3 counter-
  clockwise on z = 0 and 4..7 above them; edges 0-3 bottom ring, 4-7 top ring, 8-11 verticals); a corner is "inside" when its
  value is below the iso value;
* on each of the six faces the iso line is traced between the face's crossing edges; a face with four crossings (diagonal
  corners inside) is resolved by cutting each inside corner off separately -- a rule that depends on the face's corner signs
  only, so the two cells sharing a face always agree and the extracted surface has no cracks (the classic 1987 table does
  not have this property for every configuration);
* the face segments are oriented (inside on the left, seen from outside the cell), chained into closed loops and each loop is
  triangulated so that no diagonal (hence no triangle) lies inside a face of the cell (see _triangulate): at most five
  triangles per cell.

Triangles come out with their normal pointing to the INSIDE (towards lower values), which is the orientation the reference
then flips with ``triangles[..., ::-1]`` to get outward normals of a signed distance field.
"""
import itertools

CORNERS = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]
EDGES = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]
_FACES = [((0, 1, 2, 3), (0, 0, -1)), ((4, 5, 6, 7), (0, 0, 1)), ((0, 1, 5, 4), (0, -1, 0)), ((3, 2, 6, 7), (0, 1, 0)),
          ((0, 3, 7, 4), (-1, 0, 0)), ((1, 2, 6, 5), (1, 0, 0))]
_EDGE_ID = {frozenset(e): i for i, e in enumerate(EDGES)}
MAX_TRIS = 8          # asserted below


def _sub(a, b):
    return tuple(x - y for x, y in zip(a, b))


def _cross(a, b):
    return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])


def _dot(a, b):
    return sum(x * y for x, y in zip(a, b))


def _ccw_faces():
    """corner cycles counter-clockwise when seen from outside the cell"""
    out = []
    for cyc, n in _FACES:
        pa, pb, pc = (CORNERS[c] for c in cyc[:3])
        if _dot(_cross(_sub(pb, pa), _sub(pc, pa)), n) < 0:
            cyc = tuple(reversed(cyc))
        out.append(cyc)
    return out


def _segments(case):
    """directed iso-line segments (edge_from, edge_to) of every face, inside on the left seen from outside"""
    segs = []
    for q in _ccw_faces():
        s = [(case >> c) & 1 for c in q]
        e = [_EDGE_ID[frozenset((q[i], q[(i + 1) % 4]))] for i in range(4)]
        cross = [i for i in range(4) if s[i] != s[(i + 1) % 4]]
        if len(cross) == 2:
            i, j = cross
            # walking e_i -> e_j (i < j) the corners q_{i+1} .. q_j are on the right
            segs.append((e[j], e[i]) if s[(i + 1) % 4] else (e[i], e[j]))
        elif len(cross) == 4:
            for k in range(4):
                if s[k]:                      # cut the inside corner q_k off: edges e_{k-1} and e_k, q_k on the left
                    segs.append((e[k], e[(k - 1) % 4]))
    return segs


def _loops(case):
    nxt = {}
    for a, b in _segments(case):
        assert a not in nxt, "edge leaves twice"
        nxt[a] = b
    assert sorted(nxt) == sorted(nxt.values()), "iso lines of case %d do not close" % case
    loops, seen = [], set()
    for start in sorted(nxt):
        if start in seen:
            continue
        loop, cur = [], start
        while cur not in seen:
            seen.add(cur)
            loop.append(cur)
            cur = nxt[cur]
        assert cur == start
        loops.append(loop)
    return loops


def _edge_point(e):
    a, b = EDGES[e]
    return tuple((x + y) / 2.0 for x, y in zip(CORNERS[a], CORNERS[b]))


def _orientation_flip():
    """+1 when the fan triangles of _loops already have inward normals (checked on the single-corner case), else -1"""
    (loop,) = _loops(1)
    p = [_edge_point(e) for e in loop[:3]]
    n = _cross(_sub(p[1], p[0]), _sub(p[2], p[0]))
    centroid = tuple(sum(c) / 3.0 for c in zip(*p))
    to_inside = _sub(CORNERS[0], centroid)
    return 1 if _dot(n, to_inside) > 0 else -1


def _face_edge_sets():
    return [frozenset(_EDGE_ID[frozenset((q[i], q[(i + 1) % 4]))] for i in range(4)) for q in _ccw_faces()]


def _triangulations(poly):
    """every triangulation of a convex polygon given as a vertex list (Catalan many; loops have at most 7 vertices here... 12 in
    principle), as lists of triangles keeping the polygon's orientation"""
    if len(poly) < 3:
        return [[]]
    if len(poly) == 3:
        return [[tuple(poly)]]
    out = []
    a, b = poly[0], poly[-1]
    for k in range(1, len(poly) - 1):
        for left in _triangulations(poly[:k + 1]):
            for right in _triangulations(poly[k:]):
                out.append(left + [(a, poly[k], b)] + right)
    return out


def _triangulate(loop):
    """A triangulation of the loop none of whose DIAGONALS joins two cell edges of one face: such a diagonal lies in that
    face without being one of its iso-line segments, and the neighbouring cell (which sees the same four crossings on an
    ambiguous face) may draw it too -- four triangles on one edge, or, when a whole triangle lies in the face, a zero-volume
    double sheet.  Plain fans do this for some loops through an ambiguous face."""
    faces = _face_edge_sets()
    ring = {frozenset((loop[i], loop[(i + 1) % len(loop)])) for i in range(len(loop))}

    def in_face_diagonals(tri):
        n = 0
        for t in tri:
            for x, y in ((t[0], t[1]), (t[1], t[2]), (t[2], t[0])):
                if frozenset((x, y)) not in ring and any(x in f and y in f for f in faces):
                    n += 1
        return n

    best = None
    for tri in _triangulations(loop):
        n_bad = in_face_diagonals(tri)
        if best is None or n_bad < best[0]:
            best = (n_bad, tri)
        if n_bad == 0:
            break
    assert best[0] == 0, ("no triangulation without in-face diagonals", loop)
    return best[1]


def build_tables():
    """(n_tris[256], tris[256][MAX_TRIS * 3] edge ids (-1 padded), edge_owner[12] = (dx, dy, dz, axis))."""
    flip = _orientation_flip()
    n_tris, tris = [], []
    for case in range(256):
        row = []
        for loop in _loops(case):
            if flip < 0:
                loop = list(reversed(loop))
            for t in _triangulate(loop):
                row += list(t)
        assert len(row) <= MAX_TRIS * 3, (case, len(row))
        n_tris.append(len(row) // 3)
        tris.append(row + [-1] * (MAX_TRIS * 3 - len(row)))
    owner = []
    for a, b in EDGES:
        pa, pb = CORNERS[a], CORNERS[b]
        lo = tuple(min(x, y) for x, y in zip(pa, pb))
        axis = [i for i in range(3) if pa[i] != pb[i]][0]
        owner.append(lo + (axis,))
    return n_tris, tris, owner


def check_tables():
    """Structural self-checks used by tests/test_mcubes_cpu.py: crossing edges of a case == edges used by its triangles,
    complementary cases use the same edges, every directed triangle edge inside a cell that lies in the INTERIOR of the cell
    (fan diagonals) is matched by its reverse."""
    n_tris, tris, _ = build_tables()
    for case in range(256):
        crossing = {i for i, (a, b) in enumerate(EDGES) if ((case >> a) & 1) != ((case >> b) & 1)}
        used = {e for e in tris[case] if e >= 0}
        assert used == crossing, case
        comp = {e for e in tris[255 - case] if e >= 0}
        assert comp == used, case
        directed = set()
        for t in range(n_tris[case]):
            a, b, c = tris[case][3 * t: 3 * t + 3]
            for x, y in ((a, b), (b, c), (c, a)):
                assert (x, y) not in directed, case
                directed.add((x, y))
        # the boundary of the patch (directed edges without their reverse) is exactly the set of face segments
        boundary = {(x, y) for (x, y) in directed if (y, x) not in directed}
        segs = set(_segments(case))
        if _orientation_flip() < 0:
            segs = {(b, a) for a, b in segs}
        assert boundary == segs, case
    return max(n_tris)


if __name__ == "__main__":
    print("max triangles per cell:", check_tables())
    n, t, o = build_tables()
    print("triangles over all 256 cases:", sum(n), "owner:", o)
    _ = itertools
