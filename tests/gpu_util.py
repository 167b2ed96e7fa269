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
