"""Builders shared by the GPU parity tests: product modules loaded with the seeded synthetic
parameters of oracle/synth.py."""
import torch

import ref_conf
import synth

DEV = "cuda"


def obj_modules(device=DEV, requires_grad=True):
    import honerf_b200 as H
    sp, cp = synth.obj_states()
    emb = H.Embedding()
    sdf = H.SDFNetwork_OBJ(emb, 4, "real", **ref_conf.OBJ_SDF_CONF)
    col = H.RenderingNetwork_OBJ(emb, "real", **ref_conf.OBJ_COLOR_CONF)
    dev = H.SingleVarianceNetwork(ref_conf.VARIANCE_INIT)
    sdf.load_state_dict(sp)
    col.load_state_dict(cp)
    for m in (sdf, col, dev):
        m.to(device)
        for p in m.parameters():
            p.requires_grad_(requires_grad)
    return sdf, col, dev, sp, cp


def to_dev(d, device=DEV):
    return {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in d.items()}


def oracle_core_fp64(case, z_vals):
    """fp64 oracle of render_core + training loss + all gradients on GIVEN z_vals (the north star's
    'when the same z_vals are fed').  fp32 autograd through the reference itself is 1-2 % off on a
    few ill-conditioned colour-net tensors, so gradients are referenced in double precision."""
    import honerf_oracle as O
    R = case["R"]
    sp, cp = synth.obj_states()
    spd = {k: v.double().requires_grad_(k != "se3_refine") for k, v in sp.items()}
    cpd = {k: v.double().requires_grad_(True) for k, v in cp.items()}
    var = torch.tensor(0.3, dtype=torch.float64, requires_grad=True)
    Ro = R["Ro"].double().requires_grad_(True)
    To = R["To"].double().requires_grad_(True)
    lo, ld = O.rays_to_local(R["rays_o"].double(), R["rays_d"].double(), Ro, To)
    core = O.render_core_obj(spd, cpd, var, lo, ld, z_vals.double(), 1.1 / 64)
    out = {"color_fine": core["color"], "weight_sum": core["weights"].sum(-1, keepdim=True),
           "gradient_error": core["gradient_error"]}
    loss = O.training_loss(out, case["true_rgb"].double(), case["true_mask"].double())
    names = ["sdf." + k for k in spd if k != "se3_refine"] + ["color." + k for k in cpd] + ["variance", "Ro", "To"]
    tens = [v for k, v in spd.items() if k != "se3_refine"] + list(cpd.values()) + [var, Ro, To]
    grads = dict(zip(names, torch.autograd.grad(loss, tens)))
    return core, loss, grads, names


def reference_fp32_own_error(case, z_vals, ref_grads, names):
    """rel-L2 error of the reference algorithm's OWN fp32 arithmetic (oracle port, torch CPU) against the fp64 gradients
    `ref_grads` on the same z_vals: the conditioning yardstick for gradient bounds.  The colour net's weight gradients are
    sums over 65 536 points of terms gated by ReLUs; a rounding-level perturbation of its inputs flips a fraction f of the
    gates and moves such a sum by ~sqrt(f), so fp32 itself is 5-7e-3 away from fp64 on color.lin0 at 512 rays."""
    import honerf_oracle as O
    from golden_util import rel_l2
    R = case["R"]
    sp, cp = synth.obj_states()
    spf = {k: v.clone().requires_grad_(k != "se3_refine") for k, v in sp.items()}
    cpf = {k: v.clone().requires_grad_(True) for k, v in cp.items()}
    var = torch.tensor(0.3, requires_grad=True)
    Ro, To = R["Ro"].clone().requires_grad_(True), R["To"].clone().requires_grad_(True)
    lo, ld = O.rays_to_local(R["rays_o"], R["rays_d"], Ro, To)
    core = O.render_core_obj(spf, cpf, var, lo, ld, z_vals.float(), 1.1 / 64)
    out = {"color_fine": core["color"], "weight_sum": core["weights"].sum(-1, keepdim=True),
           "gradient_error": core["gradient_error"]}
    loss = O.training_loss(out, case["true_rgb"], case["true_mask"])
    tens = [v for k, v in spf.items() if k != "se3_refine"] + list(cpf.values()) + [var, Ro, To]
    g32 = dict(zip(names, torch.autograd.grad(loss, tens)))
    return {k: rel_l2(g32[k], ref_grads[k]) for k in names}


def hand_modules(device=DEV, requires_grad=True, use_batch=False):
    import honerf_b200 as H
    sp, cp = synth.hand_states()
    emb = H.Embedding()
    sdf = H.SDFNetwork(emb, 4, "real", use_batch=use_batch, **ref_conf.HAND_SDF_CONF)
    col = H.RenderingNetwork(emb, "real", **ref_conf.HAND_COLOR_CONF)
    dev = H.SingleVarianceNetwork(ref_conf.VARIANCE_INIT)
    sdf.load_state_dict(sp)
    col.load_state_dict(cp)
    for m in (sdf, col, dev):
        m.to(device)
        for p in m.parameters():
            p.requires_grad_(requires_grad)
    return sdf, col, dev, sp, cp
