"""Launch-list workload: forward-only hand-field render of 4096 rays (bench.py's hand_render_fwd)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import honerf_b200 as H
import ref_conf, synth
hsp, hcp = synth.hand_states()
emb = H.Embedding()
hs = H.SDFNetwork(emb, 4, "real", use_batch=False, **ref_conf.HAND_SDF_CONF)
hc = H.RenderingNetwork(emb, "real", **ref_conf.HAND_COLOR_CONF)
hd = H.SingleVarianceNetwork(ref_conf.VARIANCE_INIT)
hs.load_state_dict(hsp); hc.load_state_dict(hcp)
for m in (hs, hc, hd):
    m.cuda()
rh = H.NeuSRenderer(hs, hd, hc, "hand", **ref_conf.RENDERER_CONF)
bt, T, J = synth.hand_pose()
n = int(os.environ.get("PROF_RAYS", 4096))
HR = synth.hand_rays(n, J, seed=7)
ro, rd, bt, T = HR["rays_o"].cuda(), HR["rays_d"].cuda(), bt.cuda(), T.cuda()
for it in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    with torch.no_grad():
        out = rh.render(ro, rd, HR["near"], HR["far"], bt, T, None, None, None, 0)
    e1.record()
    torch.cuda.synchronize()
    print("render %d rays: %.3f ms" % (n, e0.elapsed_time(e1)))
