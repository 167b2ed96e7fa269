"""Marching cubes on the device (ops.marching_cubes, csrc/mcubes.cu) -- replaces the host call mcubes.marching_cubes(u, thr) of
utils/renderer.py:279,561 and utils/renderer_batch.py:309.  PyMCubes is an un-vendored dependency of the reference and is
not installed: PARITY UNPINNED.  What is checked instead are the properties any correct shared-vertex marching-cubes mesh
has, on fields where they can be computed independently with numpy: one vertex per crossing lattice edge at the linearly
interpolated position, a closed consistently oriented 2-manifold (every directed edge matched by its reverse), the Euler
characteristic of the surface, normals pointing towards lower values, and -- through extract_geometry -- vertices on the
network's zero level set inside the bounding box with outward normals after the reference's triangle flip."""
import numpy as np
import pytest
import torch

from gpu_util import DEV, obj_modules

pytestmark = pytest.mark.gpu


def _crossing_edges(u, iso):
    inside = u < iso
    return [np.argwhere(inside[:-1] != inside[1:]), np.argwhere(inside[:, :-1] != inside[:, 1:]),
            np.argwhere(inside[:, :, :-1] != inside[:, :, 1:])]


def _edge_stats(tri):
    d = {}
    for a, b, c in tri:
        for x, y in ((a, b), (b, c), (c, a)):
            d[(x, y)] = d.get((x, y), 0) + 1
    return d


@pytest.mark.parametrize("res,iso", [(33, 0.0), (20, 0.07)])
def test_sphere_mesh_properties(res, iso):
    import honerf_b200 as H
    ax = torch.linspace(-1.0, 1.0, res)
    xx, yy, zz = torch.meshgrid(ax, ax * 1.1, ax * 0.9, indexing="ij")
    u = (torch.sqrt(xx * xx + yy * yy + zz * zz) - 0.6).float()
    v, t = H.ops.marching_cubes(u.to(DEV), iso)
    v, t = v.cpu().numpy().astype(np.float64), t.cpu().numpy().astype(np.int64)
    un = u.numpy().astype(np.float64)
    cross = _crossing_edges(un, iso)
    assert len(v) == sum(len(c) for c in cross) and t.min() == 0 and t.max() == len(v) - 1
    # every vertex sits on one crossing lattice edge, at the interpolated position
    frac = v - np.floor(v)
    assert ((frac > 0).sum(axis=1) <= 1).all()
    off = 0
    for axis, c in enumerate(cross):
        vv = v[off:off + len(c)]
        off += len(c)
        base = np.floor(vv).astype(np.int64)
        # vertices are emitted axis by axis in lattice order: the same order numpy's argwhere produces
        assert (base == c).all()
        nxt = c.copy(); nxt[:, axis] += 1
        v0, v1 = un[tuple(c.T)], un[tuple(nxt.T)]
        assert np.allclose(vv[:, axis] - c[:, axis], (iso - v0) / (v1 - v0), atol=1e-6)
    # closed, consistently oriented 2-manifold: every directed edge exactly once, its reverse exactly once
    d = _edge_stats(t)
    assert all(n == 1 for n in d.values()) and all((b, a) in d for (a, b) in d)
    n_edges = len(d) // 2
    assert len(v) - n_edges + len(t) == 2                      # Euler characteristic of a sphere
    # normals point towards lower values (inside the sphere): the reference reverses them afterwards
    centre = np.array([(res - 1) / 2.0] * 3)
    p0, p1, p2 = v[t[:, 0]], v[t[:, 1]], v[t[:, 2]]
    nrm = np.cross(p1 - p0, p2 - p0)
    assert ((nrm * ((p0 + p1 + p2) / 3.0 - centre)).sum(1) < 0).all()


def test_random_field_with_ambiguous_cells_has_no_cracks():
    """A band-limited random field (several components, saddle cells: ambiguous faces): no directed edge twice, and every
    edge that does not lie on the lattice boundary is matched by its reverse (no holes between cells)."""
    import honerf_b200 as H
    g = torch.Generator().manual_seed(3)
    res = 28
    ax = torch.linspace(0, 1, res)
    xx, yy, zz = torch.meshgrid(ax, ax, ax, indexing="ij")
    u = torch.zeros(res, res, res)
    for _ in range(12):
        k = torch.randint(1, 7, (3,), generator=g).float()
        ph = torch.rand(3, generator=g) * 6.28
        u += torch.randn(1, generator=g) * torch.sin(6.28 * k[0] * xx + ph[0]) * torch.sin(6.28 * k[1] * yy + ph[1]) * \
            torch.sin(6.28 * k[2] * zz + ph[2])
    v, t = H.ops.marching_cubes(u.to(DEV), 0.0)
    v, t = v.cpu().numpy().astype(np.float64), t.cpu().numpy().astype(np.int64)
    assert len(t) > 2000
    d = _edge_stats(t)
    assert all(n == 1 for n in d.values())
    on_boundary = lambda p: bool(((p < 1e-9) | (p > res - 1 - 1e-9)).any())
    open_edges = [(a, b) for (a, b) in d if (b, a) not in d]
    assert all(on_boundary(v[a]) and on_boundary(v[b]) for a, b in open_edges), len(open_edges)
    # degenerate-free: no triangle uses a vertex twice
    assert ((t[:, 0] != t[:, 1]) & (t[:, 1] != t[:, 2]) & (t[:, 0] != t[:, 2])).all()


@pytest.mark.parametrize("seed,res,iso", [(3, (28, 28, 28), 0.0), (5, (17, 23, 9), 0.1), (7, (2, 2, 2), 0.0), (8, (33, 2, 5), -0.2)])
def test_mesh_equals_the_oracle_index_for_index(seed, res, iso):
    """oracle/mcubes_oracle.py (numpy restatement, same generated case table): vertices bit for bit (same double
    interpolation, rounded once), triangle indices identical, on cubic, ragged and single-cell lattices."""
    import honerf_b200 as H
    import mcubes_oracle as MO
    g = torch.Generator().manual_seed(seed)
    u = torch.randn(*res, generator=g)
    if min(res) > 8:                                    # smooth it a little so that there are sheets, not only noise
        u = torch.nn.functional.avg_pool3d(u[None, None], 3, 1, 1)[0, 0].contiguous()
    v, t = H.ops.marching_cubes(u.to(DEV), iso)
    vo, to = MO.marching_cubes(u.numpy(), iso)
    assert v.shape == (len(vo), 3) and t.shape == (len(to), 3)
    assert np.array_equal(v.cpu().numpy(), vo.astype(np.float32))
    assert np.array_equal(t.cpu().numpy().astype(np.int64), to.astype(np.int64))


def test_empty_mesh():
    import honerf_b200 as H
    v, t = H.ops.marching_cubes(torch.ones(6, 7, 8, device=DEV), 0.0)
    assert v.shape == (0, 3) and t.shape == (0, 3)


def test_extract_geometry_runs_on_the_device_and_lies_on_the_zero_set():
    """NeuSRenderer.extract_geometry (utils/renderer.py:260-284): numpy (vertices float64, triangles) like the reference,
    vertices inside the bounding box and near the zero level set of SDFNetwork_OBJ.sdf, outward normals after the
    reference's triangle flip.  The synthetic (untrained) weights give a steep, rough field (|grad| ~ 5, turning within a
    48^3 cell), so the linear interpolation error is bounded statistically (median < cell / 2, 99th percentile < 1.5
    cells; the numpy oracle on the CPU field gives 0.22 / 0.89 cells), not by its maximum."""
    import honerf_b200 as H
    import ref_conf
    sdf, col, dev, _, _ = obj_modules(requires_grad=False)
    r = H.NeuSRenderer(sdf, dev, col, "obj", **ref_conf.RENDERER_CONF)
    lo, hi = torch.full((3,), -0.7), torch.full((3,), 0.7)
    res = 48
    verts, tris = r.extract_geometry(lo, hi, res, None, None, None, None, threshold=0.0)
    assert isinstance(verts, np.ndarray) and verts.dtype == np.float64 and isinstance(tris, np.ndarray) and tris.shape[1] == 3
    assert len(verts) > 500 and len(tris) > 1000
    assert verts.min() >= -0.7 - 1e-6 and verts.max() <= 0.7 + 1e-6
    x = torch.from_numpy(verts).float().to(DEV)
    s, _, n = sdf.fused(x)
    cell = 1.4 / (res - 1)
    print("mesh: %d vertices, %d triangles; max |sdf| at vertices %.2e (cell %.2e)" % (len(verts), len(tris), float(s.abs().max()), cell))
    assert float(s.abs().median()) < 0.5 * cell and float(s.abs().quantile(0.99)) < 1.5 * cell
    # orientation after the reference's flip: from the cell's inside corners (u < 0) towards its outside corners.  (The
    # network's own gradient at the vertices is no use here: on the synthetic field it turns faster than the lattice.)
    u = r.sdf_grid(lo, hi, res).cpu().numpy()
    vi = (verts + 0.7) / cell
    p0, p1, p2 = (vi[tris[:, i]] for i in range(3))
    nrm = np.cross(p1 - p0, p2 - p0)
    base = np.clip(np.floor((p0 + p1 + p2) / 3.0).astype(np.int64), 0, res - 2)
    offs = np.array([(a, b, c) for a in (0, 1) for b in (0, 1) for c in (0, 1)])
    corners = base[:, None, :] + offs[None]                                        # [T, 8, 3]
    vals = u[corners[..., 0], corners[..., 1], corners[..., 2]]
    inside = (vals < 0)[..., None]
    c_in = (corners * inside).sum(1) / np.maximum(inside.sum(1), 1)
    c_out = (corners * ~inside).sum(1) / np.maximum((~inside).sum(1), 1)
    assert ((nrm * (c_out - c_in)).sum(1) > 0).mean() > 0.995


def test_lattice_and_mesh_128():
    """sdf lattice generated inside the SDF kernel (hn_sdf_obj_grid) equals the point-list path bit for bit; timing of the
    128^3 lattice + mesh extraction printed."""
    import honerf_b200 as H
    import ref_conf
    sdf, col, dev, _, _ = obj_modules(requires_grad=False)
    r = H.NeuSRenderer(sdf, dev, col, "obj", **ref_conf.RENDERER_CONF)
    lo, hi = torch.full((3,), -0.6), torch.full((3,), 0.6)
    res = 40
    u = r.sdf_grid(lo, hi, res)
    ax = [torch.linspace(float(lo[i]), float(hi[i]), res).to(DEV) for i in range(3)]
    xx, yy, zz = torch.meshgrid(*ax, indexing="ij")
    ref = sdf.sdf(torch.stack([xx, yy, zz], -1).reshape(-1, 3)).reshape(res, res, res)
    assert torch.equal(u, ref)
    assert torch.equal(r.sdf_grid(lo, hi, res, x_range=(7, 19)), u[7:19])
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    u = r.sdf_grid(lo, hi, 128)
    e1.record()
    v, t = H.ops.marching_cubes(u, 0.0)
    e2.record()
    torch.cuda.synchronize()
    print("128^3: lattice %.2f ms, marching cubes %.2f ms (%d vertices, %d triangles)" % (e0.elapsed_time(e1), e1.elapsed_time(e2), len(v), len(t)))
