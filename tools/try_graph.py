"""Debug helper: capture one object-field train step in a CUDA graph and print what breaks."""
import os
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import bench  # noqa: E402

dev = torch.device("cuda", 0)
H, renderer, params, opt = bench.build_gpu_model(dev, "tc_bf16x3")
host = bench.synthetic_batch(512, seed=7)
b = {k: v.to(dev) for k, v in host.items() if torch.is_tensor(v)}
Ro = b["Ro"].clone().requires_grad_(True)
To = b["To"].clone().requires_grad_(True)


def step():
    out = renderer.render(b["rays_o"], b["rays_d"], host["near"], host["far"], None, None, None, Ro, To, 0)
    loss = bench.training_loss(out, b["true_rgb"], b["true_mask"])
    opt.zero_grad(set_to_none=True)
    Ro.grad = None; To.grad = None
    loss.backward()
    opt.step()
    return loss


side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3):
        step()
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(g):
        loss = step()
    g.replay()
    torch.cuda.synchronize()
    print("captured OK, loss", float(loss))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print("graph replay %.3f ms/step" % (e0.elapsed_time(e1) / 10), "loss", float(loss))
except Exception:
    traceback.print_exc()
