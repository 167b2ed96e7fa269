"""Host-side logic of the product that needs no GPU: buffer layouts, camera records, NDC axes, the mask / mat-vec
formulation of the contact-stability loss (with the nearest-neighbour kernel replaced by its CPU model), and the
no-CPU-fallback rule of the new entry points."""
import pytest
import torch

import analytic as A
import cases
import honerf_oracle as O
from golden_util import load_golden, rel_err


def test_stable_loss_host_formulation_matches_reference(monkeypatch):
    """honerf_b200.ops.stable_loss_from_sdf itself (frame filter, upstream's boolean-mask setdiff1d quirk, the two
    [F,P] x [P] products, the in_time <= 1 guard) on CPU tensors, hn_nn_select swapped for oracle/analytic.py's
    brute-force model: 1e-6 relative against the reference's own get_stable_loss_cross output."""
    from honerf_b200 import ops
    monkeypatch.setattr(ops, "nn_select", lambda p, i, o, return_nearest=False: A.nn_select_bruteforce(p, i, o)[0])
    g, c = load_golden("stable"), cases.stable_case()
    sel = c["sel"][g["keep"]]
    hs = g["hand_sdf"].clone().requires_grad_(True)
    loss = ops.stable_loss_from_sdf(hs, sel)
    assert rel_err(loss, g["loss"]) < 1e-6
    d, = torch.autograd.grad(loss, [hs])
    hs_ref = g["hand_sdf"].clone().requires_grad_(True)
    d_ref, = torch.autograd.grad(O.stable_loss_from_sdf(hs_ref, sel), [hs_ref])
    assert rel_err(d, d_ref) < 1e-6
    one = g["hand_sdf"].clone()
    one[1:] = one[1:].abs() + 1e-3
    assert float(ops.stable_loss_from_sdf(one, sel)) == 0.0
    # fixed=True uses the complement of the in-set: the selected vertices are then never in-set vertices
    fixed = ops.stable_loss_from_sdf(g["hand_sdf"], sel, fixed=True)
    assert torch.isfinite(fixed) and float(fixed) > 0 and abs(float(fixed) - float(g["loss"])) > 0


def test_packed_gradient_layout_round_trip():
    """PackedMLP's flat packed gradient [dW | db]: sizes, 4-float alignment of every piece, and split_flat_grad returning
    views of exactly the regions new_grad carves (what _ParamTokenFn relies on when autograd sums two flats)."""
    from honerf_b200 import ops
    dims = [(63, 256), (256, 193), (256, 257)]
    layers = [(torch.zeros(o, 1), torch.zeros(o, i), torch.zeros(o)) for i, o in dims]
    pk = ops.PackedMLP(layers, [1.0] * 3)
    n = pk.flat_grad_floats()
    assert n == sum(o * ((i + 3) // 4 * 4) for i, o in dims) + sum((o + 3) // 4 * 4 for _, o in dims)
    flat = torch.arange(n, dtype=torch.float32)
    dW, db = pk.split_flat_grad(flat)
    assert dW.data_ptr() == flat.data_ptr() and dW.numel() == pk.total
    off = pk.total
    for (i, o), b in zip(dims, db):
        assert b.shape == (o,) and float(b[0]) == off and off % 4 == 0
        off += (o + 3) // 4 * 4
    assert all(x % 4 == 0 for x in pk.offsets) and pk.shared_token is None


def test_camera_record_and_ndc_axes():
    from honerf_b200 import ops, rays
    c = cases.rays_case()
    cam = rays.PerspectiveCameras(c["R"], c["T"], c["focal"], c["pp"])
    assert cam.record.shape == (2, 16)
    assert torch.equal(cam.record[1, :9], c["R"][1].reshape(9)) and torch.equal(cam.record[1, 9:12], c["T"][1])
    assert torch.equal(cam.record[1, 12:14], c["focal"][1]) and torch.equal(cam.record[1, 14:], c["pp"][1])
    one = ops.pack_cameras(c["R"][0], c["T"][0], c["focal"][0], c["pp"][0])      # a missing batch axis is added
    assert one.shape == (1, 16) and torch.equal(one[0], cam.record[0])
    for H_, W_ in ((5, 7), (7, 5), (4, 4)):
        xs, ys = ops.ndc_grid_axes(H_, W_, "cpu")
        xy = O.ndc_grid_xy(H_, W_).reshape(H_, W_, 2)
        assert torch.equal(xy[0, :, 0], xs) and torch.equal(xy[:, 0, 1], ys)


def test_new_entry_points_have_no_cpu_fallback():
    import honerf_b200 as H
    from honerf_b200 import ops
    c, r = cases.loss_case(n=8), cases.rays_case()
    out = {"color_fine": c["color"], "weight_sum": c["wsum"], "gradient_error": c["grad_err"],
           "sdf_hand": c["sdf_h"], "sdf_obj": c["sdf_o"]}
    with pytest.raises(H.HonerfError):
        H.losses.training_loss(out, c["true_rgb"], c["true_mask"])
    with pytest.raises(H.HonerfError):
        H.losses.fitting_render_loss(out, c["true_rgb"], c["true_mask"])
    with pytest.raises(H.HonerfError):
        H.losses.interaction_loss(out)
    with pytest.raises(H.HonerfError):
        ops.rays_from_ndc(r["xy"], ops.pack_cameras(r["R"], r["T"], r["focal"], r["pp"]))
    with pytest.raises(H.HonerfError):
        ops.nn_select(torch.zeros(4, 3), torch.zeros(1, 4, dtype=torch.bool), torch.ones(1, 4, dtype=torch.bool))
    with pytest.raises(H.HonerfError):
        ops.rays_to_local(torch.zeros(4, 3), torch.zeros(4, 3), torch.eye(3), torch.zeros(3))
    with pytest.raises(H.HonerfError):
        ops.mid_points(torch.zeros(2, 3), torch.zeros(2, 3), torch.zeros(2, 8), 0.1, with_dirs=True)
