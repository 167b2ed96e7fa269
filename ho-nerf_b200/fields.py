"""Drop-in ``nn.Module``s for the reference's field networks (utils/fields.py).

Same constructor arguments, same parameter names and shapes (``lin{l}.weight_g``,
``lin{l}.weight_v``, ``lin{l}.bias``, ``se3_refine``, ``variance``), so reference checkpoints load
with ``load_state_dict`` unchanged.  The arithmetic runs in libhonerf_b200.so; these classes only
own the parameters.  Unsupported architectures raise instead of silently falling back.
"""
import math

import numpy as np
import torch
import torch.nn as nn

from . import ops


class Embedding(nn.Module):
    """Placeholder for the reference's ``Embedding`` (utils/fields.py:8-20).  The BARF-style
    encoding is fused into the first layer of every network, so this module has no work to do; it
    exists because callers construct it and pass it as ``barf_encoding``."""

    def forward(self, input, L):  # noqa: A002 - reference signature
        raise NotImplementedError(
            "honerf_b200 fuses the positional encoding into its kernels; there is no stand-alone "
            "Embedding.forward")


class _WNLinear(nn.Module):
    """Parameter holder with the state_dict layout of nn.utils.weight_norm(nn.Linear(...)):
    keys ``bias``, ``weight_g`` [out,1], ``weight_v`` [out,in]."""

    def __init__(self, weight, bias):
        super().__init__()
        self.bias = nn.Parameter(bias)
        self.weight_g = nn.Parameter(weight.norm(dim=1, keepdim=True))
        self.weight_v = nn.Parameter(weight)
        self.in_features, self.out_features = weight.shape[1], weight.shape[0]


def _default_linear_init(in_dim, out_dim):
    """nn.Linear's default init, through nn.Linear itself so the RNG stream matches a model built
    with the reference constructors under the same seed."""
    lin = nn.Linear(in_dim, out_dim)
    return lin.weight.detach().clone(), lin.bias.detach().clone()


def _geometric_sdf_layers(dims_in, dims_out, skip_layer, d0, bias, inside_outside):
    """Geometric (sphere) initialisation of IDR/NeuS as used at utils/fields.py:99-118, 286-305."""
    layers = []
    n = len(dims_in)
    for l in range(n):
        w, b = _default_linear_init(dims_in[l], dims_out[l])
        if l == n - 1:
            mean = math.sqrt(math.pi) / math.sqrt(dims_in[l])
            if not inside_outside:
                nn.init.normal_(w, mean=mean, std=0.0001)
                nn.init.constant_(b, -bias)
            else:
                nn.init.normal_(w, mean=-mean, std=0.0001)
                nn.init.constant_(b, bias)
        elif l == 0:
            nn.init.constant_(b, 0.0)
            nn.init.constant_(w[:, 3:], 0.0)
            nn.init.normal_(w[:, :3], 0.0, math.sqrt(2) / math.sqrt(dims_out[l]))
        elif l == skip_layer:
            nn.init.constant_(b, 0.0)
            nn.init.normal_(w, 0.0, math.sqrt(2) / math.sqrt(dims_out[l]))
            nn.init.constant_(w[:, -(d0 - 3):], 0.0)
        else:
            nn.init.constant_(b, 0.0)
            nn.init.normal_(w, 0.0, math.sqrt(2) / math.sqrt(dims_out[l]))
        layers.append(_WNLinear(w, b))
    return layers


class SingleVarianceNetwork(nn.Module):
    """utils/fields.py:243-249."""

    def __init__(self, init_val):
        super().__init__()
        self.register_parameter("variance", nn.Parameter(torch.tensor(init_val)))

    def forward(self, x):
        return torch.ones([len(x), 1], device=self.variance.device) * torch.exp(self.variance * 10.0)


class _PackedNet(nn.Module):
    """Shared plumbing: lazily built PackedMLP over lin0..lin{n-1}, dropped when parameters move."""

    _post_scales = None
    _gaps = None
    _chain_kind = "bx3"

    def packed(self):
        if getattr(self, "_packed", None) is None:
            n = self.num_layers - 1
            layers = [(getattr(self, "lin%d" % l).weight_g, getattr(self, "lin%d" % l).weight_v,
                       getattr(self, "lin%d" % l).bias) for l in range(n)]
            scales = self._post_scales or [1.0] * n
            self._packed = ops.PackedMLP(layers, scales, self._gaps, chain_kind=self._chain_kind)
        return self._packed

    def _apply(self, fn, *a, **k):      # .to()/.cuda()/.float() replace the parameter tensors
        self._packed = None
        return super()._apply(fn, *a, **k)


class SDFNetwork(_PackedNet):
    """Hand SDF network (utils/fields.py:56-177): HALO feature 21 x 66 = 1386 -> 256 x4 -> [cat 1386]/sqrt2
    -> 256 x4 -> 257, softplus(beta=100), weight norm.  ``forward`` returns (out [N,257], xyz_feature
    [N,1386], None, None): the reference's extra ``r, h`` are not consumed by any caller
    (RenderingNetwork ignores ``h``, SURVEY D-5)."""
    _chain_kind = "sdf_hand"

    def __init__(self, barf_encoding, traindata_num, data_type, d_in, d_out, d_hidden, n_layers, skip_in=(4,),
                 v_multires=10, r_multires=4, bias=0.5, scale=1, geometric_init=True, weight_norm=True,
                 inside_outside=False, use_batch=False):
        super().__init__()
        if not (d_in == 3 and d_out == 257 and d_hidden == 256 and n_layers == 8 and tuple(skip_in) == (4,) and
                v_multires == 10 and r_multires == 7 and weight_norm and geometric_init):
            raise NotImplementedError(
                "honerf_b200 implements the HO-NeRF hand SDF architecture only (d_in=3, d_out=257, d_hidden=256, "
                "n_layers=8, skip_in=[4], v_multires=10, r_multires=7, weight_norm, geometric_init)")
        self.barf_encoding = barf_encoding
        self.data_type = data_type
        self.v_multires, self.r_multires = v_multires, r_multires
        self.use_batch = use_batch
        self.skip_in = tuple(skip_in)
        self.scale = scale
        d0 = ((v_multires * 2 + 1) + (r_multires * 2 * d_in + d_in)) * 21
        dims = [d0] + [d_hidden] * n_layers + [d_out]
        self.num_layers = len(dims)
        dims_in = [dims[l] + d0 if l in self.skip_in else dims[l] for l in range(self.num_layers - 1)]
        dims_out = [dims[l + 1] for l in range(self.num_layers - 1)]
        for l, lin in enumerate(_geometric_sdf_layers(dims_in, dims_out, 4, d0, bias, inside_outside)):
            setattr(self, "lin" + str(l), lin)
        se3_refine = torch.zeros((traindata_num, 6 + 3 + 20 + 7))
        se3_refine[:, 0] = 1
        se3_refine[:, 3] = 1
        self.se3_refine = nn.Parameter(se3_refine, requires_grad=True)
        self._post_scales = [ops.SQRT1_2 if l in self.skip_in else 1.0 for l in range(self.num_layers - 1)]
        self._packed = None

    def fused(self, x, bt_inv, T_pose_21):
        """(sdf [N,1], feature [N,256], normal [N,3], xyz_feature [N,1386])."""
        return ops.sdf_hand(self.packed(), x, bt_inv, T_pose_21)

    def forward(self, x, bt_inv, T_pose_21):
        sdf, feat, _, xyz = self.fused(x, bt_inv, T_pose_21)
        return torch.cat([sdf, feat], dim=-1), xyz, None, None

    def sdf(self, x, bt_inv, T_pose_21):
        needs = torch.is_grad_enabled() and (x.requires_grad or bt_inv.requires_grad or
                                             any(p.requires_grad for p in self.parameters()))
        if needs:
            return self.fused(x, bt_inv, T_pose_21)[0]
        return ops.sdf_hand_sdf_only(self.packed(), x, bt_inv, T_pose_21)

    def gradient(self, x, bt_inv, T_pose_21):
        return self.fused(x, bt_inv, T_pose_21)[2].unsqueeze(1)


class RenderingNetwork(_PackedNet):
    """Hand colour network (utils/fields.py:179-240): [xyz_feature 1386, feature 256, normal+enc4 27] = 1669
    -> 256 x4 -> 3, ReLU, sigmoid."""
    _chain_kind = "color_hand"

    def __init__(self, barf_encoding, data_type, d_feature, d_in, d_out, d_hidden, n_layers, weight_norm=True,
                 v_multires=10, r_multires=4, grad_multires=4, squeeze_out=True, use_gradients=False):
        super().__init__()
        if not (d_feature == 256 and d_in == 3 and d_out == 3 and d_hidden == 256 and n_layers == 4 and weight_norm and
                v_multires == 10 and r_multires == 7 and grad_multires == 4 and squeeze_out and use_gradients):
            raise NotImplementedError(
                "honerf_b200 implements the HO-NeRF hand colour architecture only (d_feature=256, d_hidden=256, "
                "n_layers=4, multires 10/7/4, weight_norm, squeeze_out, use_gradients)")
        self.barf_encoding = barf_encoding
        self.data_type = data_type
        self.squeeze_out, self.use_gradients = squeeze_out, use_gradients
        self.v_multires, self.r_multires, self.grad_multires = v_multires, r_multires, grad_multires
        d0 = ((v_multires * 2 + 1) + (r_multires * 2 * d_in + d_in)) * 21 + d_feature + (grad_multires * 2 * d_in + d_in)
        dims = [d0] + [d_hidden] * n_layers + [d_out]
        self.num_layers = len(dims)
        for l in range(self.num_layers - 1):
            w, b = _default_linear_init(dims[l], dims[l + 1])
            setattr(self, "lin" + str(l), _WNLinear(w, b))
        self._gaps = [(1386, 2)] + [(dims[l], 0) for l in range(1, self.num_layers - 1)]
        self._packed = None

    def forward(self, d, xyz_feature, feature_vector, h, gradients, index=None):
        return ops.color_hand(self.packed(), xyz_feature, feature_vector, gradients)


class SDFNetwork_OBJ(nn.Module):
    """Object SDF network (utils/fields.py:251-347): 63 -> 256 x3 -> 193 -> [cat 63]/sqrt2 -> 256 x4
    -> 257, softplus(beta=100), weight norm."""

    def __init__(self, barf_encoding, traindata_num, data_type, d_in, d_out, d_hidden, n_layers,
                 skip_in=(4,), v_multires=10, r_multires=4, bias=0.5, scale=1, geometric_init=True,
                 weight_norm=True, inside_outside=False):
        super().__init__()
        if not (d_in == 3 and d_out == 257 and d_hidden == 256 and n_layers == 8 and
                tuple(skip_in) == (4,) and v_multires == 10 and weight_norm and geometric_init):
            raise NotImplementedError(
                "honerf_b200 implements the HO-NeRF object SDF architecture only (d_in=3, d_out=257, "
                "d_hidden=256, n_layers=8, skip_in=[4], v_multires=10, weight_norm, geometric_init)")
        self.barf_encoding = barf_encoding
        self.data_type = data_type
        self.v_multires = v_multires
        self.skip_in = tuple(skip_in)
        self.scale = scale
        d0 = v_multires * 2 * d_in + d_in
        dims = [d0] + [d_hidden] * n_layers + [d_out]
        self.num_layers = len(dims)
        dims_in = [dims[l] for l in range(self.num_layers - 1)]
        dims_out = [dims[l + 1] - d0 if (l + 1) in self.skip_in else dims[l + 1]
                    for l in range(self.num_layers - 1)]
        for l, lin in enumerate(_geometric_sdf_layers(dims_in, dims_out, 4, d0, bias, inside_outside)):
            setattr(self, "lin" + str(l), lin)
        se3_refine = torch.zeros((traindata_num, 6 + 3))
        se3_refine[:, 0] = 1
        se3_refine[:, 3] = 1
        self.se3_refine = nn.Parameter(se3_refine, requires_grad=True)
        self._packed = None

    def packed(self):
        if self._packed is None:
            layers = [(getattr(self, "lin%d" % l).weight_g, getattr(self, "lin%d" % l).weight_v,
                       getattr(self, "lin%d" % l).bias) for l in range(self.num_layers - 1)]
            scales = [ops.SQRT1_2 if l in self.skip_in else 1.0 for l in range(self.num_layers - 1)]
            self._packed = ops.PackedMLP(layers, scales, chain_kind="sdf_obj")
        return self._packed

    def _apply(self, fn, *a, **k):      # .to()/.cuda()/.float() replace the parameter tensors
        self._packed = None
        return super()._apply(fn, *a, **k)

    def fused(self, x):
        """(sdf [N,1], feature [N,256], normal [N,3]) in one differentiable operator."""
        return ops.sdf_obj(self.packed(), x, 1.0 / float(self.scale))

    def forward(self, inputs):
        sdf, feat, _ = self.fused(inputs)
        return torch.cat([sdf, feat], dim=-1)

    def sdf(self, x):
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            return self.fused(x)[0]
        return ops.sdf_obj_sdf_only(self.packed(), x, 1.0 / float(self.scale))

    def sdf_hidden_appearance(self, x):
        return self.forward(x)

    def sdf_lattice(self, xs, ys, zs):
        """sdf on the ij-meshgrid lattice of three axes (extract_geometry): [len(xs), len(ys), len(zs)], no gradients."""
        with torch.no_grad():
            return ops.sdf_obj_lattice(self.packed(), xs, ys, zs, 1.0 / float(self.scale))

    def gradient(self, x):
        return self.fused(x)[2].unsqueeze(1)


class RenderingNetwork_OBJ(nn.Module):
    """Object colour network (utils/fields.py:349-405): 373 -> 256 x4 -> 3, ReLU, sigmoid."""

    def __init__(self, barf_encoding, data_type, d_feature, d_in, d_out, d_hidden, n_layers,
                 weight_norm=True, v_multires=10, r_multires=4, grad_multires=4, squeeze_out=True,
                 use_gradients=False):
        super().__init__()
        if not (d_feature == 256 and d_in == 3 and d_out == 3 and d_hidden == 256 and n_layers == 4 and
                weight_norm and v_multires == 10 and r_multires == 4 and grad_multires == 4 and squeeze_out):
            raise NotImplementedError(
                "honerf_b200 implements the HO-NeRF object colour architecture only (d_feature=256, "
                "d_hidden=256, n_layers=4, multires 10/4/4, weight_norm, squeeze_out)")
        self.barf_encoding = barf_encoding
        self.data_type = data_type
        self.v_multires, self.r_multires, self.grad_multires = v_multires, r_multires, grad_multires
        self.squeeze_out = squeeze_out
        self.use_gradients = use_gradients
        d0 = (r_multires * 2 * d_in + d_in) + (v_multires * 2 * d_in + d_in) + d_feature + \
            (grad_multires * 2 * d_in + d_in)
        dims = [d0] + [d_hidden] * n_layers + [d_out]
        self.num_layers = len(dims)
        for l in range(self.num_layers - 1):
            w, b = _default_linear_init(dims[l], dims[l + 1])
            setattr(self, "lin" + str(l), _WNLinear(w, b))
        self._packed = None

    def packed(self):
        if self._packed is None:
            layers = [(getattr(self, "lin%d" % l).weight_g, getattr(self, "lin%d" % l).weight_v,
                       getattr(self, "lin%d" % l).bias) for l in range(self.num_layers - 1)]
            self._packed = ops.PackedMLP(layers, [1.0] * len(layers), chain_kind="color_obj")
        return self._packed

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    def forward(self, points, view_dirs, feature_vectors, gradients, index=None):
        return ops.color_obj(self.packed(), points, view_dirs, feature_vectors, gradients)
