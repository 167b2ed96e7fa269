#!/usr/bin/env python
"""bench.py -- rays/s of the object-field NeuS training step (render fwd + second-order bwd,
64 coarse + 64 importance samples) on N B200s, with roofline, CPU baseline and end-to-end numbers.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--rays R] [--precision simt_fp32|...]
    python bench.py --impl reference ...      # the reference algorithm on the host cores (oracle port)

One JSON line on stdout (rank 0).  A "step" = NeuSRenderer.render on one batch of synthetic rays +
the training loss of exp_runner.py:206-227 (masked L1 + BCE + eikonal, no VGG) + backward +
gradient all-reduce (N > 1) + Adam step.  Rays are sharded across ranks (weak scaling: --rays per
GPU); the only collective is the flat MLP-gradient all-reduce.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

N_SAMPLES, N_IMPORTANCE = 64, 64
# algorithmic FLOPs per ray of the train step: 112 F_o + 128 (6 F_o + 3 C_o)
F_O, C_O = 1_049_088, 585_728
FLOPS_PER_RAY_TRAIN = 112 * F_O + 128 * (6 * F_O + 3 * C_O)


WORKLOAD = ("obj-field train step (BASELINE configs[2]): %d rays/GPU x (64+64) samples, masked-L1+BCE+eikonal loss, "
            "2nd-order bwd, Adam")


def training_loss(out, true_rgb, true_mask, igr_weight=1.0, mask_weight=1.0):
    """exp_runner.py:206-227 without the VGG term (caller-side code of the reference)."""
    mask_sum = true_mask.sum() + 1e-5
    color_error = (out["color_fine"] - true_rgb) * true_mask
    color_loss = F.l1_loss(color_error, torch.zeros_like(color_error), reduction="sum") / mask_sum
    mask_loss = F.binary_cross_entropy(out["weight_sum"].clip(1e-3, 1.0 - 1e-3), true_mask)
    return color_loss + mask_loss * mask_weight + out["gradient_error"] * igr_weight


def synthetic_batch(n_rays, seed):
    import synth
    R = synth.object_rays(n_rays, seed=seed)
    g = torch.Generator().manual_seed(seed + 1000)
    R["true_rgb"] = torch.rand(n_rays, 3, generator=g)
    R["true_mask"] = (torch.rand(n_rays, 1, generator=g) > 0.5).float()
    return R


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md's clocks line).  The timed
    region of the default run lasts tens of milliseconds, shorter than one `nvidia-smi` query, so the samples are taken
    through NVML (pynvml: the same counters nvidia-smi reads) from a thread polling every ~2 ms; `nvidia-smi -lms` is
    the fallback when pynvml is missing.  CUDA_VISIBLE_DEVICES is honoured when mapping the torch index to NVML's."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        ids = [v for v in vis.split(",") if v.strip() != ""]
        try:
            self.index = int(ids[index]) if ids else index
        except Exception:
            self.index = index
        self.rows, self.proc, self.nvml, self.stop_flag, self.t = [], None, None, False, None
        self.sm, self.mx, self.reasons = [], None, set()

    def _poll(self):
        n = self.nvml
        names = {n.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 n.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 n.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 n.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
                mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if mask & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.001)

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.t.join(timeout=1)
            sm = sorted(self.sm)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.mx, "reasons": sorted(self.reasons),
                    "samples": len(sm), "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except Exception:
                continue
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvidia-smi"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm (oracle port, torch CPU, all host threads)
# ------------------------------------------------------------------------------------------------
def _reference_modules():
    """The UNMODIFIED reference's hot-path modules (oracle/ref_loader.py) when its tree is reachable: HONERF_REFERENCE_ROOT,
    baseline/_ref or /root/reference (the build container).  None on the GPU box, where only the oracle port travels."""
    if os.environ.get("HONERF_REFERENCE_ROOT") == "none":       # force the oracle port (what the GPU box runs)
        return None, None
    for root in (os.environ.get("HONERF_REFERENCE_ROOT"), os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if root and os.path.isfile(os.path.join(root, "utils", "renderer.py")):
            os.environ["HONERF_REFERENCE_ROOT"] = root
            try:
                import importlib
                import ref_loader
                importlib.reload(ref_loader)
                return ref_loader, ref_loader.load_reference()
            except Exception as e:      # noqa: BLE001
                sys.stderr.write("bench.py: reference tree at %s could not be imported (%s); using the oracle port\n" % (root, e))
    return None, None


def cpu_step_fns(n_rays, seed=7):
    """(train_step, forward_only, kind): the reference algorithm on the host.  kind = "reference" when the reference's own
    NeuSRenderer / networks run (utils/renderer.py, utils/fields.py through oracle/ref_loader.py), else "port"
    (oracle/honerf_oracle.py).  Same synthetic batch, weights and loss either way."""
    import honerf_oracle as O
    import synth
    B = synthetic_batch(n_rays, seed)
    rl, ref = _reference_modules()
    if ref is not None:
        sp, cp = synth.obj_states()
        emb = ref.fields.Embedding()
        sdf = ref.fields.SDFNetwork_OBJ(emb, 4, "real", **rl.OBJ_SDF_CONF)
        col = ref.fields.RenderingNetwork_OBJ(emb, "real", **rl.OBJ_COLOR_CONF)
        dev = ref.fields.SingleVarianceNetwork(rl.VARIANCE_INIT)
        sdf.load_state_dict(sp); col.load_state_dict(cp)
        r = ref.renderer.NeuSRenderer(sdf, dev, col, "obj", **rl.RENDERER_CONF)
        params = list(sdf.parameters()) + list(dev.parameters()) + list(col.parameters())
        opt = torch.optim.Adam(params, lr=1e-4)
        zb, zT = torch.zeros(21, 4, 4), torch.zeros(21, 3)

        def render():
            return r.render(B["rays_o"], B["rays_d"], B["near"], B["far"], zb, zT, None, B["Ro"], B["To"], 0)

        def step():
            out = render()
            loss = training_loss(out, B["true_rgb"], B["true_mask"])
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
            return float(loss)

        def fwd():
            with torch.no_grad():
                # the reference's gradient() needs autograd even for a forward render (utils/fields.py:336-347)
                with torch.enable_grad():
                    return float(render()["color_fine"].sum())
        return step, fwd, "reference"
    sp, cp = synth.obj_states()
    sp = {k: v.clone().requires_grad_(k != "se3_refine") for k, v in sp.items()}
    cp = {k: v.clone().requires_grad_(True) for k, v in cp.items()}
    var = torch.tensor(0.3, requires_grad=True)
    params = [v for k, v in sp.items() if k != "se3_refine"] + list(cp.values()) + [var]
    opt = torch.optim.Adam(params, lr=1e-4)

    def render():
        return O.render_obj(sp, cp, var, B["rays_o"], B["rays_d"], B["near"], B["far"], B["Ro"], B["To"], B["t_rand"])

    def step():
        loss = O.training_loss(render(), B["true_rgb"], B["true_mask"])
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return float(loss)

    def fwd():
        return float(render()["color_fine"].sum())
    return step, fwd, "port"


def time_cpu(n_rays, steps, warmup, forward_steps=0):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step, fwd, kind = cpu_step_fns(n_rays)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    fwd_rps = None
    if forward_steps > 0:
        fwd()
        t0 = time.perf_counter()
        for _ in range(forward_steps):
            fwd()
        fwd_rps = n_rays / ((time.perf_counter() - t0) / forward_steps)
    return n_rays / dt, dt, cores, kind, fwd_rps


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_rays = min(args.rays, 512)      # bounded sample: one 512-ray batch per step (~1.5-4 s on the host)
    # exactly the K steps / W warm-ups asked for, bounded so that the arm stays within a few minutes
    steps, warmup = max(1, min(args.steps, 40)), max(0, min(args.warmup, 10))
    rps, dt, cores, kind, fwd_rps = time_cpu(n_rays, steps, warmup, forward_steps=2)
    what = ("the reference's own NeuSRenderer / SDFNetwork_OBJ / RenderingNetwork_OBJ (oracle/ref_loader.py)" if kind == "reference"
            else "oracle/honerf_oracle.py (functional port, pinned to the reference by tests/golden)")
    line = {
        "impl": "reference", "metric": "rays/sec render fwd+bwd (64+64 samples), object-field train step",
        "value": rps, "unit": "rays/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD % n_rays, "rays_per_gpu": n_rays, "precision": "f32 (torch CPU)",
                   "parallelism": "CPU threads x%d" % cores},
        "cpu_baseline": {"value": rps, "unit": "rays/s", "cores": cores, "kind": kind,
                         "sample": "%d steps of one %d-ray batch: %s, torch CPU" % (steps, n_rays, what)},
        "forward_only": {"value": fwd_rps, "unit": "rays/s", "workload": "BASELINE configs[0]: render_core forward, %d rays x (64+64) "
                         "samples, on CPU" % n_rays},
        "e2e": {"value": rps, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def build_gpu_model(device, precision, optimizer="flat"):
    import honerf_b200 as H
    import ref_conf
    import synth
    H.set_default_precision(precision)
    sp, cp = synth.obj_states()
    sdf = H.SDFNetwork_OBJ(H.Embedding(), 4, "real", **ref_conf.OBJ_SDF_CONF)
    col = H.RenderingNetwork_OBJ(H.Embedding(), "real", **ref_conf.OBJ_COLOR_CONF)
    var = H.SingleVarianceNetwork(ref_conf.VARIANCE_INIT)
    sdf.load_state_dict(sp); col.load_state_dict(cp)
    for m in (sdf, col, var):
        m.to(device)
    r = H.NeuSRenderer(sdf, var, col, "obj", **ref_conf.RENDERER_CONF)
    params = list(sdf.parameters()) + list(var.parameters()) + list(col.parameters())
    if optimizer == "flat":
        from honerf_b200.optim import FlatAdam
        opt = FlatAdam(params, lr=1e-4)         # every parameter in one buffer, one Adam launch per step
    else:
        opt = torch.optim.Adam(params, lr=1e-4, fused=True, capturable=True)   # capturable in a CUDA graph
    return H, r, params, opt


def fitting_extra(H, device, n_rays, precision):
    """Informational: one pose-fitting iteration of fitting_single.py through NeuSRenderer_fitting (utils/renderer.py:
    286-572): hand + object fields frozen, gradients to the hand pose (bt_inv) and the object pose (Ro, To) only."""
    import ref_conf
    import synth
    hsp, hcp = synth.hand_states()
    osp, ocp = synth.obj_states()
    emb = H.Embedding()
    hs = H.SDFNetwork(emb, 4, "real", use_batch=False, **ref_conf.HAND_SDF_CONF)
    hc = H.RenderingNetwork(emb, "real", **ref_conf.HAND_COLOR_CONF)
    os_ = H.SDFNetwork_OBJ(emb, 4, "real", **ref_conf.OBJ_SDF_CONF)
    oc = H.RenderingNetwork_OBJ(emb, "real", **ref_conf.OBJ_COLOR_CONF)
    hd, od = H.SingleVarianceNetwork(ref_conf.VARIANCE_INIT), H.SingleVarianceNetwork(ref_conf.VARIANCE_INIT)
    hs.load_state_dict(hsp); hc.load_state_dict(hcp); os_.load_state_dict(osp); oc.load_state_dict(ocp)
    for m in (hs, hc, hd, os_, oc, od):
        m.to(device)
        for q in m.parameters():
            q.requires_grad_(False)
    r = H.renderer.NeuSRenderer_fitting(hs, hd, hc, os_, od, oc, **ref_conf.RENDERER_CONF)
    bt, T, J = synth.hand_pose()
    HR = synth.hand_rays(n_rays, J, seed=7)
    g = torch.Generator().manual_seed(106)
    Ro = synth.random_rotation(g).to(device).requires_grad_(True)
    To = (J.mean(0) + 0.02 * torch.randn(3, generator=g)).to(device).requires_grad_(True)
    true_rgb = torch.rand(n_rays, 3, generator=g).to(device)
    true_mask = (torch.rand(n_rays, 1, generator=g) > 0.3).float().to(device)
    bt = bt.to(device).requires_grad_(True)
    T = T.to(device)
    ro, rd = HR["rays_o"].to(device), HR["rays_d"].to(device)

    def step():
        out = r.render(ro, rd, HR["near"], HR["far"], bt, T, None, Ro, To)
        # fit_type '12' of fitting_single.py:253-283 (render loss + 30 contact + 20 penetration), fused loss kernels
        loss = H.losses.fitting_render_loss(out, true_rgb, true_mask) + H.losses.interaction_loss(out)
        bt.grad = None; Ro.grad = None; To.grad = None
        loss.backward()
        return out

    def timed_steps():
        for _ in range(2):
            o = step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            step()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 3, o
    ms, out = timed_steps()
    r.field_streams = True              # the two fields on two CUDA streams (NeuSRenderer_fitting.field_streams)
    ms_fs, _ = timed_steps()
    return {"rays": n_rays, "samples_per_ray_and_field": int(out["sdf_hand"].shape[0] // n_rays), "ms_per_step": ms,
            "value": n_rays / (ms * 1e-3), "unit": "rays/s", "precision": precision,
            "field_streams": {"ms_per_step": ms_fs, "value": n_rays / (ms_fs * 1e-3)},
            "loss": "fitting_single.py:253-283 (render + 30 contact + 20 penetration), fused loss kernels",
            "finite_pose_grads": bool(torch.isfinite(bt.grad).all() and torch.isfinite(Ro.grad).all()),
            "note": "eager launches; hand field on the per-layer pre-packed bf16x3 contractions (gemm_bx3), object field on the chain kernels"}


def _time_calls(fn, warm=2, reps=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def forward_extras(H, device, renderer, batch, host, Ro, To, hand_rays_n):
    """Informational forward-only renders (no autograd): BASELINE configs[0]'s shape on the GPU (object field, 512 rays)
    and configs[3]'s hand field on a full-image chunk of rays."""
    import ref_conf
    import synth
    res = {}
    with torch.no_grad():
        n = batch["rays_o"].shape[0]
        ms = _time_calls(lambda: renderer.render(batch["rays_o"], batch["rays_d"], host["near"], host["far"], None, None,
                                                 None, Ro, To, 0))
        res["obj_render_fwd"] = {"rays": n, "ms": ms, "value": n / (ms * 1e-3), "unit": "rays/s",
                                 "algorithmic_tflops": n * (112 * F_O + 128 * (2 * F_O + C_O)) / (ms * 1e-3) / 1e12}
        hsp, hcp = synth.hand_states()
        emb = H.Embedding()
        hs = H.SDFNetwork(emb, 4, "real", use_batch=False, **ref_conf.HAND_SDF_CONF)
        hc = H.RenderingNetwork(emb, "real", **ref_conf.HAND_COLOR_CONF)
        hd = H.SingleVarianceNetwork(ref_conf.VARIANCE_INIT)
        hs.load_state_dict(hsp); hc.load_state_dict(hcp)
        for m in (hs, hc, hd):
            m.to(device)
        rh = H.NeuSRenderer(hs, hd, hc, "hand", **ref_conf.RENDERER_CONF)
        bt, T, J = synth.hand_pose()
        HR = synth.hand_rays(hand_rays_n, J, seed=7)
        bt, T = bt.to(device), T.to(device)
        ro, rd = HR["rays_o"].to(device), HR["rays_d"].to(device)
        ms = _time_calls(lambda: rh.render(ro, rd, HR["near"], HR["far"], bt, T, None, None, None, 0))
        res["hand_render_fwd"] = {"rays": hand_rays_n, "ms": ms, "value": hand_rays_n / (ms * 1e-3), "unit": "rays/s",
                                  "note": "hand SDF net through its chain kernels (csrc/chain16_hand.cu), hand colour net per layer"}
    return res


def _dist_max_ms(ms, device, world):
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)
    return ms


def _timed_region(fn, device, world):
    """fn() between barrier + synchronize on both sides, CUDA events, max over ranks; returns (ms, fn's result)."""
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = fn()
    e1.record()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()
    return _dist_max_ms(e0.elapsed_time(e1), device, world), out


def lattice_extra(H, renderer, device, res, rank, world):
    """BASELINE configs[1]: SDFNetwork_OBJ.sdf on the res^3 lattice of extract_geometry (utils/renderer.py:260-284,
    bbox [-0.2, 0.2]^3 of exp_runner.py:511-517), sharded over the ranks in x slabs (honerf_b200.dist.shard_rays on the x
    axis) with a final all-gather of u; the lattice points are generated inside the kernel (hn_sdf_obj_grid)."""
    from honerf_b200 import dist as hdist
    bmin, bmax = torch.full((3,), -0.2), torch.full((3,), 0.2)
    sizes = [hdist.shard_rays(res, r, world) for r in range(world)]
    lo, hi = sizes[rank]
    # warm-up at 64^3 INCLUDING the all-gather: the first collective of a kind pays NCCL's lazy channel set-up (it put 150 ms
    # into one 8-GPU run of this line)
    wsz = [hdist.shard_rays(64, r, world) for r in range(world)]
    w = renderer.sdf_grid(bmin, bmax, 64, x_range=wsz[rank])
    if world > 1:
        hdist.gather_slabs(w, [b - a for a, b in wsz], dim=0)
    del w

    def run():
        u = renderer.sdf_grid(bmin, bmax, res, x_range=(lo, hi))
        return hdist.gather_slabs(u, [b - a for a, b in sizes], dim=0) if world > 1 else u
    ms, u = _timed_region(run, device, world)
    npts = res ** 3
    out = {"resolution": res, "ms": ms, "points_per_s": npts / (ms * 1e-3), "algorithmic_tflops": npts * F_O / (ms * 1e-3) / 1e12,
           "finite": bool(torch.isfinite(u).all()), "shape": list(u.shape), "sharding": "x slabs x%d + all-gather of u" % world,
           "n_gpus": world}
    del u
    return out


def hand_views_extra(H, device, rank, world, image=512, chunk=4096):
    """BASELINE configs[3]: full-image render of synthetic image x image views of the hand field (HALO pose-conditioned,
    wmask_realhand_hand1.conf), ONE VIEW PER GPU (weak scaling: N GPUs render N views), rays generated on the device
    from the NDC grid in chunks (exp_runner.py:338-369)."""
    import math
    import ref_conf
    import synth
    from honerf_b200 import rays as hrays
    hsp, hcp = synth.hand_states()
    emb = H.Embedding()
    hs = H.SDFNetwork(emb, 4, "real", use_batch=False, **ref_conf.HAND_SDF_CONF)
    hc = H.RenderingNetwork(emb, "real", **ref_conf.HAND_COLOR_CONF)
    hd = H.SingleVarianceNetwork(ref_conf.VARIANCE_INIT)
    hs.load_state_dict(hsp); hc.load_state_dict(hcp)
    for m in (hs, hc, hd):
        m.to(device)
    rh = H.NeuSRenderer(hs, hd, hc, "hand", **dict(ref_conf.RENDERER_CONF, perturb=0.0))
    bt, T, J = synth.hand_pose()
    bt, T = bt.to(device), T.to(device)
    # camera `rank` of 8 on a ring around the hand, looking at its centroid (pytorch3d row-vector convention)
    a = 2.0 * math.pi * rank / 8.0
    c = J.mean(0)
    eye = c + 0.9 * torch.tensor([math.sin(a), 0.0, -math.cos(a)])
    zax = torch.nn.functional.normalize(c - eye, dim=0)
    xax = torch.nn.functional.normalize(torch.linalg.cross(torch.tensor([0.0, 1.0, 0.0]), zax), dim=0)
    yax = torch.linalg.cross(zax, xax)
    R = torch.stack([xax, yax, zax], dim=1)[None]
    Tc = (-(eye @ R[0]))[None]
    cam = hrays.PerspectiveCameras(R, Tc, torch.tensor([[6.0, 6.0]]), torch.zeros(1, 2)).to(device)

    def run():
        img = torch.empty(image * image, 3, device=device)
        with torch.no_grad():
            first = 0
            for ro, rd in hrays.image_ray_chunks(cam, image, image, chunk):
                out = rh.render(ro, rd, 0.4, 1.5, bt, T, None, None, None, 0)
                img[first:first + len(ro)] = out["color_fine"]
                first += len(ro)
        return img
    with torch.no_grad():
        ro, rd = next(iter(hrays.image_ray_chunks(cam, image, image, chunk)))
        rh.render(ro, rd, 0.4, 1.5, bt, T, None, None, None, 0)
    ms, img = _timed_region(run, device, world)
    n = image * image
    return {"views": world, "image": [image, image], "rays": world * n, "ms": ms, "value": world * n / (ms * 1e-3), "unit": "rays/s",
            "finite": bool(torch.isfinite(img).all()), "coverage": float((img.sum(-1) > 0).float().mean()),
            "algorithmic_tflops": world * n * (112 * 2_468_352 + 128 * (2 * 2_468_352 + 1_249_280)) / (ms * 1e-3) / 1e12,
            "sharding": "one %dx%d view per GPU" % (image, image)}


def fitting_views_extra(H, device, rank, world, n_views=8, rays_per_view=196, iters=2):
    """BASELINE configs[4]: one pose-fitting iteration of fitting_single.py:200-291 over the 8 views of a frame (fit_12_8views:
    196 rays per view), views sharded over the ranks, pose gradients (bt_inv, Ro, To) SUMMED with one flat NCCL all-reduce
    (honerf_b200.dist.allreduce_gradients(average=False)).  Rank 0 also renders all views alone and checks the summed
    gradient against it."""
    import ref_conf
    import synth
    from honerf_b200 import dist as hdist
    hsp, hcp = synth.hand_states()
    osp, ocp = synth.obj_states()
    emb = H.Embedding()
    hs = H.SDFNetwork(emb, 4, "real", use_batch=False, **ref_conf.HAND_SDF_CONF)
    hc = H.RenderingNetwork(emb, "real", **ref_conf.HAND_COLOR_CONF)
    os_ = H.SDFNetwork_OBJ(emb, 4, "real", **ref_conf.OBJ_SDF_CONF)
    oc = H.RenderingNetwork_OBJ(emb, "real", **ref_conf.OBJ_COLOR_CONF)
    hd, od = H.SingleVarianceNetwork(ref_conf.VARIANCE_INIT), H.SingleVarianceNetwork(ref_conf.VARIANCE_INIT)
    hs.load_state_dict(hsp); hc.load_state_dict(hcp); os_.load_state_dict(osp); oc.load_state_dict(ocp)
    for m in (hs, hd, hc, os_, od, oc):
        m.to(device)
        for q in m.parameters():
            q.requires_grad_(False)
    r = H.renderer.NeuSRenderer_fitting(hs, hd, hc, os_, od, oc, **dict(ref_conf.RENDERER_CONF, perturb=0.0))
    bt0, T, J = synth.hand_pose()
    g = torch.Generator().manual_seed(106)
    Ro0 = synth.random_rotation(g)
    To0 = J.mean(0) + 0.02 * torch.randn(3, generator=g)
    views = []
    for v in range(n_views):
        HR = synth.hand_rays(rays_per_view, J, seed=40 + v)
        gv = torch.Generator().manual_seed(200 + v)
        views.append((HR["rays_o"].to(device), HR["rays_d"].to(device), torch.rand(rays_per_view, 3, generator=gv).to(device),
                      (torch.rand(rays_per_view, 1, generator=gv) > 0.3).float().to(device)))
    T = T.to(device)
    bt = bt0.to(device).requires_grad_(True)
    Ro, To = Ro0.to(device).requires_grad_(True), To0.to(device).requires_grad_(True)
    pose = [bt, Ro, To]

    def iteration(lo, hi, reduce):
        for q in pose:
            q.grad = None
        for v in range(lo, hi):
            ro, rd, rgb, mask = views[v]
            out = r.render(ro, rd, 0.4, 1.5, bt, T, None, Ro, To)
            loss = H.losses.fitting_render_loss(out, rgb, mask) + H.losses.interaction_loss(out)
            loss.backward()                              # gradients of the views of this rank accumulate
        if reduce:
            hdist.allreduce_gradients(pose, world, average=False)
        return [q.grad.clone() for q in pose]

    def iteration_batched(lo, hi, reduce):
        """The same iteration with the views of this rank rendered as ONE ray batch (the views of a frame share the hand and
        object pose; rays are independent), the reference's per-view losses evaluated on the slices: the sum of the per-view
        gradients in one forward / backward, 8 x fewer launches and whole waves of tiles for the persistent field kernels."""
        for q in pose:
            q.grad = None
        if hi <= lo:
            for q in pose:
                q.grad = torch.zeros_like(q)
        else:
            ro = torch.cat([views[v][0] for v in range(lo, hi)])
            rd = torch.cat([views[v][1] for v in range(lo, hi)])
            out = r.render(ro, rd, 0.4, 1.5, bt, T, None, Ro, To)
            nb = ro.shape[0]
            loss = 0.0
            for i, v in enumerate(range(lo, hi)):
                a, b = i * rays_per_view, (i + 1) * rays_per_view
                sl = {k: (t[a:b] if torch.is_tensor(t) and t.dim() > 0 and t.shape[0] == nb else t) for k, t in out.items()}
                loss = loss + H.losses.fitting_render_loss(sl, views[v][2], views[v][3]) + H.losses.interaction_loss(sl)
            loss.backward()
        if reduce:
            hdist.allreduce_gradients(pose, world, average=False)
        return [q.grad.clone() for q in pose]
    lo, hi = hdist.shard_views(n_views, rank, world)
    iteration(lo, hi, world > 1)
    ms_loop, grads_loop = _timed_region(lambda: [iteration(lo, hi, world > 1) for _ in range(iters)][-1], device, world)
    ms_loop /= iters
    iteration_batched(lo, hi, world > 1)
    ms, grads = _timed_region(lambda: [iteration_batched(lo, hi, world > 1) for _ in range(iters)][-1], device, world)
    ms /= iters
    rel = lambda xs, ys: max(float((a - b).norm() / (b.norm() + 1e-30)) for a, b in zip(xs, ys))
    batched_vs_loop = rel(grads, grads_loop)
    check = None
    if world > 1:
        ref = iteration(0, n_views, False) if rank == 0 else None
        if rank == 0:
            check = rel(grads_loop, ref)
    n = n_views * rays_per_view
    return {"views": n_views, "rays_per_view": rays_per_view, "samples_per_ray_and_field": 192, "ms_per_iteration": ms,
            "value": n / (ms * 1e-3), "unit": "rays/s", "finite_pose_grads": bool(all(torch.isfinite(q).all() for q in grads)),
            "views_per_render_call": "all views of the rank in one ray batch, per-view losses on the slices",
            "per_view_loop": {"ms_per_iteration": ms_loop, "value": n / (ms_loop * 1e-3), "unit": "rays/s",
                              "note": "one render call per view, as fitting_single.py:200-291 loops"},
            "pose_grad_batched_vs_per_view_loop_rel_err": batched_vs_loop,
            "pose_grad_allreduce_vs_single_rank_rel_err": check,
            "algorithmic_tflops": n * 3_799_990_000 / (ms * 1e-3) / 1e12,
            "sharding": "views x%d, one flat all-reduce (SUM) of the 336 + 9 + 3 pose-gradient floats" % world,
            "loss": "fitting_single.py:253-283 (render + 30 contact + 20 penetration), fused loss kernels"}


def _dbg(msg):
    if os.environ.get("BENCH_DEBUG"):
        sys.stderr.write("[bench rank %s %.1fs] %s\n" % (os.environ.get("RANK", "0"), time.perf_counter(), msg))
        sys.stderr.flush()


def run_gpu_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    # stdout carries exactly ONE JSON line: anything a library prints there (NCCL's version banner) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
    H, renderer, params, opt = build_gpu_model(device, args.precision, args.optimizer)
    renderer.ray_streams = args.ray_streams
    if args.ray_shards:
        renderer.ray_shard_sizes = [int(v) for v in args.ray_shards.split(",")]
    flat_opt = args.optimizer == "flat"
    from honerf_b200 import dist as hdist
    n_rays = args.rays
    host = synthetic_batch(n_rays, seed=7 + rank)
    pinned = {k: v.pin_memory() for k, v in host.items() if torch.is_tensor(v)}
    dev_batch = {k: v.to(device) for k, v in pinned.items()}
    Ro = dev_batch["Ro"].clone().requires_grad_(True)
    To = dev_batch["To"].clone().requires_grad_(True)

    flat_holder = {}

    def fwd_bwd(b):
        """render + loss + backward; for N > 1 also packs every gradient into one flat buffer"""
        if args.loss == "fused" and args.shard_loss:
            # per-shard loss on the shard's own stream (no join between forward and backward): the shards' totals add
            # up to the batch loss -- batch-wide mask_sum + 1e-5 as a device scalar, BCE / eikonal means weighted by
            # the shard's share of the rays
            div = b["true_mask"].sum() + 1e-5

            def shard_loss(out, lo, hi):
                w = (hi - lo) / float(n_rays)
                return H.ops.render_loss(out["color_fine"], out["weight_sum"], b["true_rgb"][lo:hi], b["true_mask"][lo:hi],
                                         out["gradient_error"], div, 1.0, w, w)[0]
            parts = renderer.render_sharded(b["rays_o"], b["rays_d"], host["near"], host["far"], None, None, None, Ro, To,
                                            0, shard_loss)
            loss = parts[0] if len(parts) == 1 else torch.stack(parts).sum()
        else:
            out = renderer.render(b["rays_o"], b["rays_d"], host["near"], host["far"], None, None, None, Ro, To, 0)
            if args.loss == "fused":
                # csrc/loss.cu: one forward + one backward launch instead of ~30 torch launches
                loss = H.ops.render_loss(out["color_fine"], out["weight_sum"], b["true_rgb"], b["true_mask"],
                                         out["gradient_error"], 0.0, 1.0, 1.0, 1.0)[0]
            else:
                loss = training_loss(out, b["true_rgb"], b["true_mask"])
        opt.zero_grad(set_to_none=True)
        Ro.grad = None; To.grad = None
        loss.backward()
        if flat_opt:
            flat_holder["runs"] = opt.gather_grads()        # gradients -> opt.flat_grad
        elif world > 1:
            flat_holder["flat"] = hdist.flatten_gradients(params)
        return loss

    # ---- the step's ONE exchange: the flat gradient buffer ------------------------------------------------------------
    # default: summed over the ranks INSIDE the Adam kernel, over NVLink peer memory (hn_peer_adam_flat, csrc/peer.cu);
    # --exchange nccl (or a failed set-up / check, decided collectively): ncclAllReduce followed by hn_adam_flat
    peer = {"on": False, "check": None, "why": None}
    if world > 1 and flat_opt and args.exchange == "peer":
        import torch.distributed as dist
        if opt.enable_peer_exchange():
            gen = torch.Generator(device=device)
            gen.manual_seed(1234 + rank)
            opt.flat_grad.copy_(torch.randn(opt.n, device=device, generator=gen))
            want = opt.flat_grad.clone()
            dist.all_reduce(want, op=dist.ReduceOp.SUM)
            got = opt.peer_allreduce()
            torch.cuda.synchronize()
            rel = float((got - want).abs().max() / want.abs().max())
            good = torch.tensor([1 if (opt.peer_error() == 0 and rel < 1e-5) else 0], device=device)
            dist.all_reduce(good, op=dist.ReduceOp.MIN)
            opt.flat_grad.zero_()
            peer["check"] = rel
            if int(good) == 1:
                peer["on"] = True
            else:
                peer["why"] = "check against ncclAllReduce failed (rel %.3g, err %d)" % (rel, opt.peer_error())
        else:
            peer["why"] = "peer-memory set-up failed (cudaIpc handles)"
        if not peer["on"]:
            sys.stderr.write("bench.py: peer-memory exchange off (%s); using ncclAllReduce\n" % peer["why"])

    def reduce_grads():
        """the NCCL form of the exchange: all-reduce of the flat gradient buffer"""
        if world > 1 and not peer["on"]:
            hdist.allreduce_flat(opt.flat_grad if flat_opt else flat_holder["flat"])

    def apply_grads(local=False):
        if flat_opt:
            # averaging folded into the Adam kernel; `local`: rank 0's solo kernel-timing pass (no collective)
            opt.step(flat_holder["runs"], grad_scale=1.0 / world, peer_exchange=peer["on"] and not local)
            return
        if world > 1:
            hdist.unflatten_gradients(params, flat_holder["flat"], world)
        opt.step()

    def train_step(b):
        loss = fwd_bwd(b)
        reduce_grads()
        apply_grads()
        return loss

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    # ---- device-resident number -------------------------------------------------------------
    # The whole step (render, loss, backward, gradient all-reduce, Adam) is a fixed launch sequence: capture it once
    # in a CUDA graph and replay it, as a training loop would.  Everything before the capture runs on the capture's
    # side stream (autograd's AccumulateGrad nodes remember the stream they were created on).
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    static = {k: v.clone() for k, v in dev_batch.items()}
    # the per-step inputs live in ONE flat device buffer (views below) mirrored by ONE pinned host buffer, so the end-to-
    # end step uploads a batch with a single copy, the way a data loader would hand it over
    in_keys = ("rays_o", "rays_d", "t_rand", "true_rgb", "true_mask")
    offs, off = {}, 0
    for k in in_keys:
        offs[k] = off
        off += (pinned[k].numel() + 3) // 4 * 4
    flat_host = torch.zeros(off, dtype=torch.float32).pin_memory()
    flat_dev = torch.zeros(off, dtype=torch.float32, device=device)
    for k in in_keys:
        assert pinned[k].dtype == torch.float32
        flat_host[offs[k]:offs[k] + pinned[k].numel()].copy_(pinned[k].reshape(-1))
        static[k] = flat_dev[offs[k]:offs[k] + pinned[k].numel()].view(pinned[k].shape)
    flat_dev.copy_(flat_host)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(args.warmup):
            train_step(static)
        torch.cuda.synchronize()
        l0 = H.launch_count()
        train_step(static)
        launches_per_step = H.launch_count() - l0
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    _dbg("warm-up done, %d launches per step" % launches_per_step)
    graph, graph_b, loss_static = None, None, None
    nccl_in_graph = False
    if not args.no_graph:
        try:
            # N = 1: one graph for the whole step.  N > 1: first try ONE graph as well, with the NCCL all-reduce captured as a
            # node between the backward pass and Adam (no launch gap around the collective); if this NCCL / driver
            # refuses, fall back to graph A = render + loss + backward + gradient packing, the all-reduce launched eagerly,
            # graph B = Adam.  Every rank takes the same path (the outcome of the capture is all-reduced).
            if world > 1 and not args.no_nccl_graph:
                ok = 1
                try:
                    g1 = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g1):
                        loss_static = train_step(static)
                    g1.replay()
                    torch.cuda.synchronize()
                except Exception as e:      # noqa: BLE001
                    sys.stderr.write("bench.py: NCCL in a CUDA graph refused (%s); using the two-graph step\n" % str(e)[:200])
                    ok = 0
                import torch.distributed as dist
                flag = torch.tensor([ok], device=device)
                dist.all_reduce(flag, op=dist.ReduceOp.MIN)
                if int(flag) == 1:
                    graph, nccl_in_graph = g1, True
            if graph is None:
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    loss_static = fwd_bwd(static) if world > 1 else train_step(static)
                if world > 1:
                    graph_b = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph_b, pool=graph.pool()):
                        apply_grads()
                graph.replay()
                if world > 1:
                    reduce_grads()
                    graph_b.replay()
            torch.cuda.synchronize()
            if not torch.isfinite(loss_static).all():
                raise RuntimeError("non-finite loss from the captured step")
        except Exception as e:      # noqa: BLE001
            # a failed capture leaves the CUDA RNG registered to a dead graph: start over without graphs
            sys.stderr.write("bench.py: CUDA graph capture failed (%s: %s); re-running with --no-graph\n"
                             % (type(e).__name__, str(e)[:300]))
            sys.stderr.flush()
            if world > 1:
                raise           # under torchrun every rank must take the same path: fail loudly instead
            os.execv(sys.executable, [sys.executable] + sys.argv + ["--no-graph"])
    _dbg("graph capture done: %s" % (graph is not None))

    def resident_step():
        if graph is not None:
            graph.replay()
            if graph_b is not None:
                reduce_grads()
                graph_b.replay()
        else:
            train_step(dev_batch)

    for _ in range(2):
        resident_step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms = timed(resident_step, args.steps)
    launches = launches_per_step * args.steps
    ms_per_step = ms / args.steps
    value = world * n_rays / (ms_per_step * 1e-3)

    # ---- end to end: pinned host inputs -> H2D -> step -> D2H loss, every step --------------------
    loss_host = torch.empty((), pin_memory=True)
    keys = in_keys
    h2d = flat_host.numel() * 4

    def e2e_step():
        if graph is not None:
            flat_dev.copy_(flat_host, non_blocking=True)       # the step's inputs: one pinned -> device copy
            resident_step()
            loss_host.copy_(loss_static.detach(), non_blocking=False)
        else:
            b = {k: pinned[k].to(device, non_blocking=True) for k in keys}
            loss = train_step(b)
            loss_host.copy_(loss.detach(), non_blocking=False)

    for _ in range(2):
        e2e_step()
    _dbg("device-resident timing done")
    ms_e2e = timed(e2e_step, args.steps) / args.steps
    # the sampler ran through both timed regions (device-resident and end-to-end steps, back to back under load)
    clocks = sampler.stop() if rank == 0 else None
    _dbg("e2e timing done")
    e2e_value = world * n_rays / (ms_e2e * 1e-3)

    # ---- roofline of the dominant kernel family (the MLP contractions), rank 0, separate pass ----
    roof, comp = None, None
    if rank == 0 and not args.no_roofline:
        # rank 0 alone runs this pass: the step without its collective
        # per-kernel durations are taken with the shards serialised (one stream): events around a launch that shares
        # the SMs with the other shard's kernels would time the sharing, not the kernel
        renderer.ray_streams = 1
        roof = mlp_roofline(H, lambda: (fwd_bwd(dev_batch), apply_grads(local=True)), n_rays)
        renderer.ray_streams = args.ray_streams
        if roof is not None:
            roof["measured_with"] = "ray_streams=1 (kernels serialised, one launch per family and step)"
        comp = compositor_roofline(H, device)
    large = None
    if rank == 0 and world == 1 and args.large_rays > 0:
        # informational: the same step on a batch that fills the 148 SMs for many waves (not the headline config)
        try:
            hb = synthetic_batch(args.large_rays, seed=99)
            lb = {k: v.to(device) for k, v in hb.items() if torch.is_tensor(v)}
            for _ in range(2):
                train_step(lb)
            ms_l = timed(lambda: train_step(lb), 3) / 3      # world == 1 here: no collective inside
            large = {"rays_per_gpu": args.large_rays, "ms_per_step": ms_l, "value": args.large_rays / (ms_l * 1e-3),
                     "unit": "rays/s", "note": "eager launches, device-resident inputs"}
            del lb
        except Exception as e:      # noqa: BLE001
            large = {"error": str(e)[:200]}
        torch.cuda.empty_cache()
    # ---- strong scaling: the SAME 512 rays split over the ranks, eager launches ------------------------------
    strong = None
    if peer["on"]:
        e = opt.peer_error()
        if e:
            raise RuntimeError("bench.py: hn_peer_adam_flat gave up waiting for rank %d inside the timed region" % (e - 1))
        barrier()           # rank 0 ran its solo pass: line the ranks up before the next peer-memory step
    if world > 1 and args.strong_rays > 0:
        from honerf_b200 import dist as hdist
        hb = synthetic_batch(args.strong_rays, seed=7)        # the same batch on every rank
        lo, hi = hdist.shard_rays(args.strong_rays, rank, world)
        sb = {k: (v[lo:hi] if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == args.strong_rays else v) for k, v in hb.items()}
        sb = {k: v.to(device) for k, v in sb.items() if torch.is_tensor(v)}
        n_keep = n_rays
        n_rays = hi - lo                                      # fwd_bwd's shard weights use the local ray count
        renderer.ray_streams = 1
        for _ in range(3):
            train_step(sb)
        ms_s = timed(lambda: train_step(sb), 5) / 5
        renderer.ray_streams = args.ray_streams
        n_rays = n_keep
        strong = {"rays_total": args.strong_rays, "rays_per_gpu": hi - lo, "ms_per_step": ms_s,
                  "value": args.strong_rays / (ms_s * 1e-3), "unit": "rays/s",
                  "note": "strong scaling: one 512-ray batch split over the ranks (per-rank mean losses, gradients averaged), eager launches"}
    # ---- the other BASELINE configs, sharded over the ranks (every rank takes part: they contain collectives) --------------
    grid = None
    if args.grid_res > 0:
        try:
            grid = lattice_extra(H, renderer, device, args.grid_res, rank, world)
        except Exception as e:      # noqa: BLE001
            if world > 1:
                raise
            grid = {"error": str(e)[:200]}
        torch.cuda.empty_cache()
    fit = None
    if args.fit_rays > 0:
        try:
            fit = fitting_views_extra(H, device, rank, world)
            if world == 1:
                fit["one_512_ray_batch"] = fitting_extra(H, device, args.fit_rays, args.precision)
        except Exception as e:      # noqa: BLE001
            if world > 1:
                raise
            fit = {"error": str(e)[:200]}
        torch.cuda.empty_cache()
    fwd_extra = None
    if args.fit_rays > 0:
        try:
            fwd_extra = {}
            if world == 1:
                renderer.ray_streams = 1          # eager launches: extra streams only add CPU launch work here
                fwd_extra = forward_extras(H, device, renderer, dev_batch, host, Ro, To, 4096)
                renderer.ray_streams = args.ray_streams
            fwd_extra["hand_views"] = hand_views_extra(H, device, rank, world, image=args.view_size)
        except Exception as e:      # noqa: BLE001
            if world > 1:
                raise
            fwd_extra = {"error": str(e)[:200]}
        torch.cuda.empty_cache()
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rps, dt, cores, kind, fwd_rps = time_cpu(min(n_rays, 512), 3, 1, forward_steps=2)
        cpu = {"value": rps, "unit": "rays/s", "cores": cores, "kind": kind,
               "sample": "3 steps of one %d-ray batch (%s, torch CPU)" % (
                   min(n_rays, 512), "the reference's own classes" if kind == "reference" else "oracle/honerf_oracle.py"),
               "forward_only_rays_per_s": fwd_rps}
    if rank == 0:
        # every extra that states its algorithmic TFLOP/s also states the fraction of the measured dense bf16 peak (same
        # denominator as `roofline`: sustained figure, these are long runs), per GPU
        peaks, peak_src = measured_peaks()
        peak_tf = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1400.0)))

        def add_frac(d):
            if isinstance(d, dict):
                if isinstance(d.get("algorithmic_tflops"), (int, float)):
                    per_gpu = d["algorithmic_tflops"] / max(int(d.get("n_gpus", world if "sharding" in d else 1)), 1)
                    d["roofline"] = {"bound": "tensor", "achieved": per_gpu, "peak": peak_tf, "unit": "TFLOP/s per GPU",
                                     "frac": per_gpu / peak_tf, "peak_source": peak_src}
                for v in list(d.values()):
                    add_frac(v)
        for extra in (grid, fit, fwd_extra):
            add_frac(extra)
        line = {
            "metric": "rays/sec render fwd+bwd (64+64 samples), object-field train step",
            "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"simt_fp32": "f32", "tc_tf32": "tf32", "tc_tf32x3": "tf32 (3xTF32 split operands)",
                      "tc_bf16x3": "bf16 (split hi+lo operands, 3 MMAs per product, fp32 accumulate)",
                      "tc_mixed16": "fp16/bf16 (value trunk: fp16 hi+lo, 3 MMAs; gradient sweeps: one 16-bit operand x hi+lo "
                                    "weights, 2 MMAs; weight gradients: bf16, 1 MMA; fp32 accumulate)"}[args.precision],
            "data": "synthetic",
            "config": {"workload": WORKLOAD % n_rays,
                       "rays_per_gpu": n_rays, "precision": args.precision, "parallelism": "rays sharded x%d" % world,
                       "cuda_graph": graph is not None,
                       "allreduce": (None if world == 1 else
                                     "two-shot sum over NVLink peer memory fused with Adam (hn_peer_adam_flat, one kernel, captured "
                                     "in the step's CUDA graph); checked against ncclAllReduce before timing: max rel diff %.2g"
                                     % peer["check"] if peer["on"] else
                                     "NCCL all-reduce of the flat gradient buffer captured in the step's CUDA graph"
                                     if nccl_in_graph else "NCCL all-reduce launched eagerly between two CUDA graphs"),
                       "exchange_fallback": peer["why"],
                       "optimizer": ("FlatAdam (hn_peer_adam_flat, one launch)" if peer["on"] else "FlatAdam (hn_adam_flat, one launch)")
                       if flat_opt else "torch.optim.Adam(fused, capturable)",
                       "loss": "hn_render_loss_fwd/_bwd (fused)" if args.loss == "fused" else "torch ops",
                       "ray_streams": args.ray_streams, "ray_shards": args.ray_shards or "equal",
                       "shard_loss": bool(args.shard_loss) and args.loss == "fused",
                       "l2": "per-step activation stash (~2 GB at 512 rays) exceeds the 126 MB L2; no explicit flush"},
            "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e},
            "gpu_launches": int(launches),
            "clocks": clocks, "roofline": roof, "compositor": comp, "cpu_baseline": cpu, "large_batch": large,
            "strong_scaling": strong, "sdf_grid": grid, "fitting_step": fit, "forward_only": fwd_extra,
            "mlp_flops_per_ray": FLOPS_PER_RAY_TRAIN,
        }
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        # Leave without tearing NCCL down: the step's CUDA graph holds captured ncclAllReduce nodes, and destroying the
        # communicator (or the graph) while the other exists has hung the 2-GPU run after the JSON line was written.  Every
        # rank has finished its work once the barrier returns; the result is on stdout; exit code 0 is all torchrun needs.
        import torch.distributed as dist
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


KERNEL_FAMILIES = ["per-layer contractions (gemm_* kernels)", "chain::sdf_only_kernel", "chain::sdf_fwd_kernel",
                   "chain::sdf_bwd_kernel", "chain::dw_kernel", "chain::color_fwd_kernel", "chain::color_bwd_kernel"]


def _family_flops_per_step(n_rays):
    """ALGORITHMIC FLOPs per train step of each kernel family (sums to n_rays * FLOPS_PER_RAY_TRAIN)."""
    n = n_rays * (N_SAMPLES + N_IMPORTANCE)
    return [0, n_rays * 112 * F_O,       # sampler queries: 64 + 3 x 16 points per ray, F_o each
            n * 2 * F_O,                 # value trunk + feature head, normal sweep
            n * 2 * F_O,                 # tangent sweep, reverse sweep
            n * (2 * F_O + C_O),         # weight gradients of both nets
            n * C_O, n * C_O]


def mlp_roofline(H, step_fn, n_rays):
    """Tensor roofline of the MLP kernels: algorithmic FLOPs (per ray 112 F_o + 128 (6 F_o + 3 C_o)) over
    device time taken with CUDA events recorded on the launching stream inside the library around every such launch
    (hn_timing_*), per kernel family.  `achieved`/`frac` at the top level are those of the DOMINANT kernel (largest
    share of the step); `step` aggregates all families."""
    import ctypes
    from honerf_b200 import _lib
    peaks, src = measured_peaks()
    _lib.lib.hn_timing_enable(1)
    step_fn()
    torch.cuda.synchronize()
    _lib.lib.hn_timing_reset()
    reps = 3
    for _ in range(reps):
        step_fn()
    torch.cuda.synchronize()
    nt = len(KERNEL_FAMILIES)
    ms = (ctypes.c_double * nt)()
    cnt = (ctypes.c_int64 * nt)()
    _lib.lib.hn_timing_collect_tags(ms, cnt, nt)
    _lib.lib.hn_timing_enable(0)
    if sum(cnt) == 0:
        return None
    peak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
    flops = _family_flops_per_step(n_rays)
    # DRAM traffic per launch comes from an `ncu --set full` capture condensed by tools/ncu_metrics.py; it is only quoted
    # when that capture was taken on THIS build (sha1 over csrc/), a stale file yields traffic = null
    traffic, traffic_src = {}, None
    tp = os.path.join(ROOT, "profiles", "r02_ncu_metrics.json")
    if os.path.isfile(tp):
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        from ncu_metrics import build_id
        d = json.load(open(tp))
        if d.get("build_id") == build_id():
            traffic = d.get("dram_bytes_per_launch", {})
            traffic_src = "dram__bytes_read.sum + dram__bytes_write.sum per launch, profiles/r02_ncu_metrics.json (build %s)" % d["build_id"]
        else:
            traffic_src = "profiles/r02_ncu_metrics.json was captured on build %s, this library is %s: not quoted" % (d.get("build_id"), build_id())
    fam = []
    for t in range(nt):
        if cnt[t] == 0:
            continue
        ms_step = ms[t] / reps
        fam.append({"kernel": KERNEL_FAMILIES[t], "launches_per_step": cnt[t] // reps, "ms_per_step": ms_step,
                    "ms_per_launch": ms[t] / cnt[t], "algorithmic_flops_per_step": flops[t],
                    "achieved": flops[t] / (ms_step * 1e-3) / 1e12 if flops[t] else None,
                    "frac": flops[t] / (ms_step * 1e-3) / 1e12 / peak if flops[t] else None,
                    "traffic": traffic.get(KERNEL_FAMILIES[t])})
    dom = max(fam, key=lambda f: f["ms_per_step"])
    tot_ms = sum(f["ms_per_step"] for f in fam)
    tot_fl = n_rays * FLOPS_PER_RAY_TRAIN
    return {"bound": "tensor", "kernel": dom["kernel"],
            "achieved": dom["achieved"], "peak": peak, "unit": "TFLOP/s", "frac": dom["frac"], "traffic": dom["traffic"],
            "traffic_source": traffic_src,
            "algorithmic_flops_per_launch": dom["algorithmic_flops_per_step"] // max(dom["launches_per_step"], 1),
            "ms_per_launch": dom["ms_per_launch"], "share_of_mlp_time": dom["ms_per_step"] / tot_ms,
            "peak_source": src + ", sustained bf16 (dense, no split: the 3-MMA split caps algorithmic FLOPs at 1/3 of it)",
            "step": {"achieved": tot_fl / (tot_ms * 1e-3) / 1e12, "frac": tot_fl / (tot_ms * 1e-3) / 1e12 / peak,
                     "mlp_ms_per_step": tot_ms, "algorithmic_flops_per_step": tot_fl},
            "families": fam}


def compositor_roofline(H, device, n_rays=1 << 18, n=128):
    """HBM roofline of the compositor alone: 2^18 rays x 128 samples, fwd 40 B + bwd 64 B per sample."""
    from honerf_b200 import ops
    peaks, src = measured_peaks()
    N = n_rays * n
    g = torch.Generator(device=device).manual_seed(0)
    sdf = (torch.rand(N, 1, device=device, generator=g) - 0.3).requires_grad_(True)
    nrm = torch.randn(N, 3, device=device, generator=g).requires_grad_(True)
    rgb = torch.rand(N, 3, device=device, generator=g).requires_grad_(True)
    dists = torch.full((n_rays, n), 1.1 / 128, device=device)
    d = F.normalize(torch.randn(n_rays, 3, device=device, generator=g), dim=-1)
    var = torch.tensor(0.3, device=device, requires_grad=True)
    res = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    from honerf_b200._lib import lib, check
    import ctypes
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    w = torch.empty(n_rays, n, device=device); c = torch.empty_like(w)
    col = torch.empty(n_rays, 3, device=device); ws = torch.empty(n_rays, device=device)
    wm = torch.empty_like(ws); ek = torch.empty_like(ws)
    ds, dn, dr = torch.empty(N, device=device), torch.empty(N, 3, device=device), torch.empty(N, 3, device=device)
    dv = torch.zeros(1, device=device)
    gc = torch.randn(n_rays, 3, device=device)

    def fwd():
        check(lib.hn_neus_composite_fwd(P(sdf), P(nrm), P(rgb), P(dists), P(d), P(var), n_rays, n, 1, P(w), P(c), None,
                                        P(col), P(ws), P(wm), P(ek), st), "fwd")

    def bwd():
        check(lib.hn_neus_composite_bwd(P(sdf), P(nrm), P(rgb), P(dists), P(d), P(var), P(w), n_rays, n, 1, P(gc), None,
                                        None, None, P(ds), P(dn), P(dr), None, P(dv), st), "bwd")

    for name, fn, bytes_per_sample in (("fwd", fwd, 40), ("bwd", bwd, 64)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        gbs = N * bytes_per_sample / (ms * 1e-3) / 1e9
        res[name] = {"ms": ms, "achieved": gbs, "unit": "GB/s", "peak": peaks["hbm_gbs"], "frac": gbs / peaks["hbm_gbs"],
                     "algorithmic_bytes": N * bytes_per_sample}
    res["bound"] = "hbm"
    res["workload"] = "%d rays x %d samples (3.4 GB fwd working set > L2)" % (n_rays, n)
    res["peak_source"] = src
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--rays", type=int, default=512, help="rays per GPU per step")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="tc_mixed16", choices=["simt_fp32", "tc_tf32", "tc_tf32x3", "tc_bf16x3", "tc_mixed16"])
    ap.add_argument("--optimizer", default="flat", choices=["flat", "torch"],
                    help="flat: honerf_b200.optim.FlatAdam (one launch); torch: torch.optim.Adam(fused, capturable)")
    ap.add_argument("--ray-streams", type=int, default=2,
                    help="render each GPU's rays as this many shards on concurrent CUDA streams (NeuSRenderer.ray_streams)")
    ap.add_argument("--ray-shards", default="", help="explicit shard sizes, e.g. 148,148,216 (overrides --ray-streams)")
    ap.add_argument("--shard-loss", type=int, default=1,
                    help="1: the fused loss is evaluated per ray shard on the shard's stream (render_sharded); 0: one "
                         "loss launch on the merged outputs")
    ap.add_argument("--loss", default="fused", choices=["fused", "torch"],
                    help="fused: hn_render_loss_fwd/_bwd (default); torch: the reference's loss lines as torch ops")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true",
                    help="skip the per-kernel event-timing passes (for an ncu launch list of the train step alone)")
    ap.add_argument("--no-graph", action="store_true", help="launch every step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--large-rays", type=int, default=4096, help="extra informational measurement (0 disables)")
    ap.add_argument("--fit-rays", type=int, default=512, help="extra informational measurements: two-field fitting step, forward-only renders (0 disables)")
    ap.add_argument("--grid-res", type=int, default=512, help="extra informational SDF-lattice measurement (0 disables)")
    ap.add_argument("--view-size", type=int, default=512, help="side of the synthetic hand views of the full-image extra")
    ap.add_argument("--strong-rays", type=int, default=512, help="N > 1: strong-scaling extra on this many rays in total (0 disables)")
    ap.add_argument("--no-nccl-graph", action="store_true", help="N > 1: keep the all-reduce outside the CUDA graphs")
    ap.add_argument("--exchange", choices=["peer", "nccl"], default="peer",
                    help="N > 1: how the flat gradient is summed over the ranks -- peer: inside the Adam kernel over NVLink peer "
                         "memory (hn_peer_adam_flat; falls back to nccl if the set-up or its check fails); nccl: ncclAllReduce")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
