"""Seeded synthetic inputs shared by the tests, the golden-vector generator and bench.py
(SURVEY.md section 8d).  TEST / BENCH INFRASTRUCTURE ONLY -- plain torch CPU tensors, no reference code.
"""
import math

import torch


def perturb_state_dict(sd, seed=1, v_noise=0.05, b_noise=0.05):
    """Dense-noise perturbation so that every input column of every layer matters (geometric
    init zeroes most of lin0 and of the skip columns)."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, v in sd.items():
        v = v.detach().clone()
        if k.endswith("weight_v"):
            v = v + v_noise * torch.randn(v.shape, generator=g)
        elif k.endswith("weight_g"):
            v = v * (0.8 + 0.4 * torch.rand(v.shape, generator=g))
        elif k.endswith("bias"):
            v = v + b_noise * torch.randn(v.shape, generator=g)
        out[k] = v
    return out


def random_rotation(g):
    q = torch.randn(4, generator=g)
    q = q / q.norm()
    w, x, y, z = q.tolist()
    return torch.tensor([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                         [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                         [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def object_rays(B, seed=2, small_pose=False):
    """Camera near -0.95 z looking at the origin so rays cross the ~0.5-radius init sphere."""
    g = torch.Generator().manual_seed(seed)
    o = torch.tensor([0.0, 0.0, -0.95]) + 0.05 * torch.randn(B, 3, generator=g)
    d = torch.nn.functional.normalize(torch.tensor([0.0, 0.0, 1.0]) + 0.1 * torch.randn(B, 3, generator=g), dim=-1)
    Ro = random_rotation(g)
    To = 0.02 * torch.randn(3, generator=g)
    # world rays such that the object-frame rays are (o, d):  o_w = Ro^T o + To, d_w = Ro^T d
    o_w = o @ Ro + To
    d_w = d @ Ro
    t_rand = torch.rand(B, 1, generator=g) - 0.5
    return dict(rays_o=o_w, rays_d=d_w, Ro=Ro, To=To, t_rand=t_rand, near=0.4, far=1.5)


# A rest-pose right hand skeleton: wrist + 5 fingers x 4 joints, 0.09 m root bones, 0.03 m others.
def hand_skeleton(seed=3, noise=0.003):
    g = torch.Generator().manual_seed(seed)
    joints = [torch.zeros(3)]
    for f in range(5):
        ang = math.radians(-40 + 20 * f)
        dirv = torch.tensor([math.sin(ang), math.cos(ang), 0.0])
        base = dirv * 0.09
        joints.append(base)
        for k in range(1, 4):
            joints.append(base + dirv * 0.03 * k)
    J = torch.stack(joints) + noise * torch.randn(21, 3, generator=g)
    return J


def hand_pose(seed=3, n_frames=None):
    """Synthetic per-bone un-pose transforms bt_inv [21,4,4] (rigid) and T_pose_21 [21,3].

    The real chain (halo_util/converter_fit_batch.py) is a caller of the hot path and is out of
    scope; any set of rigid 4x4s exercises the same kernel code."""
    def one(s):
        g = torch.Generator().manual_seed(s)
        J = hand_skeleton(s)
        bt = torch.zeros(21, 4, 4)
        for j in range(21):
            # small random rotation about a random axis
            ax = torch.nn.functional.normalize(torch.randn(3, generator=g), dim=0)
            th = 0.3 * torch.randn(1, generator=g).item()
            K = torch.tensor([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
            R = torch.eye(3) + math.sin(th) * K + (1 - math.cos(th)) * (K @ K)
            bt[j, :3, :3] = R
            bt[j, :3, 3] = 0.01 * torch.randn(3, generator=g)
            bt[j, 3, 3] = 1.0
        T = torch.einsum("jab,jb->ja", bt[:, :3, :3], J) + bt[:, :3, 3]
        return bt, T, J
    if n_frames is None:
        return one(seed)
    bts, Ts, Js = zip(*[one(seed + 17 * f) for f in range(n_frames)])
    return torch.stack(bts), torch.stack(Ts), torch.stack(Js)


def hand_rays(B, J, seed=4, near=0.4, far=1.5):
    """Rays from a camera ~0.9 m in front of the hand, aimed at points around the joints."""
    g = torch.Generator().manual_seed(seed)
    centre = J.mean(0)
    o = centre + torch.tensor([0.0, 0.0, -0.9]) + 0.02 * torch.randn(B, 3, generator=g)
    tgt = J[torch.randint(0, 21, (B,), generator=g)] + 0.01 * torch.randn(B, 3, generator=g)
    d = torch.nn.functional.normalize(tgt - o, dim=-1)
    t_rand = torch.rand(B, 1, generator=g) - 0.5
    return dict(rays_o=o, rays_d=d, t_rand=t_rand, near=near, far=far)


# ------------------------------------------------------------------------------------------
# Seeded network parameters with the reference's state_dict keys and shapes.
# Built here (not by the reference constructors) so that tests on the GPU box, which has no
# reference tree, can rebuild the exact tensors the golden vectors were generated with.
# ------------------------------------------------------------------------------------------
OBJ_SDF_DIMS = [(63, 256), (256, 256), (256, 256), (256, 193), (256, 256), (256, 256), (256, 256),
                (256, 256), (256, 257)]
OBJ_COLOR_DIMS = [(373, 256), (256, 256), (256, 256), (256, 256), (256, 3)]
HAND_SDF_DIMS = [(1386, 256), (256, 256), (256, 256), (256, 256), (1642, 256), (256, 256),
                 (256, 256), (256, 256), (256, 257)]
HAND_COLOR_DIMS = [(1669, 256), (256, 256), (256, 256), (256, 256), (256, 3)]


def make_state(dims, seed, kind="sdf", noise=None, n_se3=0, se3_width=9):
    """kind == 'sdf': geometric-init-like statistics (utils/fields.py:99-118, 286-305) plus dense
    noise on every entry; kind == 'color': default nn.Linear-like uniform init plus noise."""
    g = torch.Generator().manual_seed(seed)
    if noise is None:
        # SDF nets: the perturbation must stay well below the per-unit signal of the geometric
        # init (~0.05) or the zero level set disappears and every ray renders empty space.
        noise = 0.003 if kind == "sdf" else 0.05
    sd = {}
    n = len(dims)
    d0 = dims[0][0]
    for l, (din, dout) in enumerate(dims):
        if kind == "sdf":
            if l == n - 1:
                w = math.sqrt(math.pi) / math.sqrt(din) + 1e-4 * torch.randn(dout, din, generator=g)
                b = torch.full((dout,), -0.5)
            else:
                w = torch.randn(dout, din, generator=g) * (math.sqrt(2) / math.sqrt(dout))
                b = torch.zeros(dout)
                if l == 0:
                    w[:, 3:] = 0.0
                elif l == 4:              # skip layer: zero the re-injected encoding columns
                    w[:, -(d0 - 3):] = 0.0
        else:
            bound = 1.0 / math.sqrt(din)
            w = (torch.rand(dout, din, generator=g) * 2 - 1) * bound
            b = (torch.rand(dout, generator=g) * 2 - 1) * bound
        scale = noise
        if kind == "sdf" and din > 512:
            scale = scale * 0.3           # 1386/1642-wide layers: keep the summed perturbation small
        w = w + scale * torch.randn(dout, din, generator=g)
        b = b + noise * torch.randn(dout, generator=g)
        gnorm = w.norm(dim=1, keepdim=True) * (0.9 + 0.2 * torch.rand(dout, 1, generator=g))
        sd["lin%d.weight_v" % l] = w
        sd["lin%d.weight_g" % l] = gnorm
        sd["lin%d.bias" % l] = b
    if n_se3:
        se3 = torch.zeros(n_se3, se3_width)
        se3[:, 0] = 1
        se3[:, 3] = 1
        sd["se3_refine"] = se3
    return sd


def obj_states(seed=10):
    return (make_state(OBJ_SDF_DIMS, seed, "sdf", n_se3=4, se3_width=9),
            make_state(OBJ_COLOR_DIMS, seed + 1, "color"))


def hand_states(seed=20):
    """Hand nets.  With the geometric init the 1386-wide HALO feature barely reaches the output
    (only 3 live columns), so lin0 and the skip columns get dense weights, and the output bias is
    re-centred so that empty space (all-zero feature) sits at sdf = +0.04: rays then see a mix of
    free space and surface crossings near the bones."""
    sd = make_state(HAND_SDF_DIMS, seed, "sdf", n_se3=4, se3_width=36)
    g = torch.Generator().manual_seed(seed + 7)
    sd["lin0.weight_v"] = sd["lin0.weight_v"] + 0.006 * torch.randn(256, 1386, generator=g)
    sd["lin4.weight_v"][:, 256:] += 0.003 * torch.randn(256, 1386, generator=g)
    for l in (0, 4):
        sd["lin%d.weight_g" % l] = sd["lin%d.weight_v" % l].norm(dim=1, keepdim=True) * (
            0.9 + 0.2 * torch.rand(256, 1, generator=g))
    # far-field value with an all-zero feature
    x = torch.zeros(1, 1386)
    feat0 = x
    for l in range(9):
        if l == 4:
            x = torch.cat([x, feat0], 1) / math.sqrt(2.0)
        v = sd["lin%d.weight_v" % l]
        w = v * (sd["lin%d.weight_g" % l] / v.norm(dim=1, keepdim=True))
        x = x @ w.T + sd["lin%d.bias" % l]
        if l < 8:
            x = torch.nn.functional.softplus(x, beta=100)
    sd["lin8.bias"][0] -= x[0, 0] - 0.04
    return sd, make_state(HAND_COLOR_DIMS, seed + 1, "color")
