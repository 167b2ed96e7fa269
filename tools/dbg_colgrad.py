import sys
sys.path[:0] = ['/root/repo', '/root/repo/oracle', '/root/repo/tests']
import torch
import cases, honerf_oracle as O, synth, ref_conf
import honerf_b200 as H
from honerf_b200 import ops
from gpu_util import DEV, obj_modules
from golden_util import rel_err, max_abs
c = cases.obj_render_case(); R = c["R"]
def ref_run():
    sdf, col, dev, sp, cp = obj_modules()
    spr = {k: v.clone().requires_grad_(k != "se3_refine") for k, v in sp.items()}
    cpr = {k: v.clone().requires_grad_(True) for k, v in cp.items()}
    var = torch.tensor(0.3, requires_grad=True)
    ref = O.render_obj(spr, cpr, var, R["rays_o"], R["rays_d"], R["near"], R["far"], R["Ro"], R["To"], R["t_rand"])
    loss = O.training_loss(ref, c["true_rgb"], c["true_mask"])
    names = ["sdf." + k for k in spr if k != "se3_refine"] + ["color." + k for k in cpr] + ["variance"]
    tens = [v for k, v in spr.items() if k != "se3_refine"] + list(cpr.values()) + [var]
    return ref, dict(zip(names, torch.autograd.grad(loss, tens))), names
ref, ref_g, names = ref_run()
def run(sdf_prec, col_prec):
    sdf, col, dev, sp, cp = obj_modules()
    r = H.NeuSRenderer(sdf, dev, col, "obj", **ref_conf.RENDERER_CONF)
    sdf.fused = lambda x: ops.sdf_obj(sdf.packed(), x, 1.0, precision=sdf_prec)
    col.forward = lambda p, d, f, n, i=None: ops.color_obj(col.packed(), p, d, f, n, precision=col_prec)
    lo, ld = r.convert_obj_to_local(R["rays_o"].to(DEV), R["rays_d"].to(DEV), R["Ro"].to(DEV), R["To"].to(DEV))
    r.index = 0
    core = r.render_core(lo, ld, None, None, None, ref["z_vals"].to(DEV), 1.1 / 64, sdf, dev, col)
    out = {"color_fine": core["color"], "weight_sum": core["weights"].sum(-1, keepdim=True), "gradient_error": core["gradient_error"]}
    loss = O.training_loss(out, c["true_rgb"].to(DEV), c["true_mask"].to(DEV))
    loss.backward()
    got = {"sdf." + k: p.grad for k, p in sdf.named_parameters() if p.grad is not None}
    got.update({"color." + k: p.grad for k, p in col.named_parameters() if p.grad is not None})
    got["variance"] = dev.variance.grad
    worst = sorted(((rel_err(got[k], ref_g[k]), k) for k in names), reverse=True)[:3]
    print("sdf_prec %d col_prec %d: colour err %.2e  worst grads %s" % (sdf_prec, col_prec, max_abs(out["color_fine"], ref["color_fine"]), [(k, "%.2e" % e) for e, k in worst]))
for sp_, cp_ in [(0, 0), (2, 0), (0, 1), (2, 1), (1, 1)]:
    run(sp_, cp_)
