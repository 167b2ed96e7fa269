"""Torch-facing operators over the C ABI: tensor plumbing (allocation, streams, autograd glue) only.
All arithmetic happens in libhonerf_b200.so; nothing here falls back to PyTorch math."""
import ctypes
import math

import torch

from . import _lib
from ._lib import HN_SIMT_FP32, HN_WS_BWD, HN_WS_SDF_ONLY, check, hn_mlp_grad_t, hn_mlp_t, lib

_PRECISIONS = {"simt_fp32": _lib.HN_SIMT_FP32, "tc_tf32": _lib.HN_TC_TF32, "tc_tf32x3": _lib.HN_TC_TF32X3,
               "tc_bf16x3": _lib.HN_TC_BF16X3, "tc_mixed16": _lib.HN_TC_MIXED16}
# Product default: 'tc_mixed16' = the fused tcgen05 chain kernels with the 16-bit activation stash for the object SDF field
# (csrc/chain16*.cu), 'tc_bf16x3' arithmetic everywhere else.  'simt_fp32' is the verification path (the north star's
# "fp32 SIMT path kept for verification"); HONERF_PRECISION overrides the default at import.
import os as _os

_default_precision = _PRECISIONS[_os.environ.get("HONERF_PRECISION", "tc_mixed16")]


def set_default_precision(name):
    """'simt_fp32' (verification path) | 'tc_tf32' | 'tc_tf32x3' (split TF32 operands, per-layer kernels) |
    'tc_bf16x3' (fused tile-chain kernels, split bf16 operands) | 'tc_mixed16' (object SDF field: fp16x3 value trunk,
    single-16-bit-operand sweeps in tensor memory, 16-bit activation stash; everything else as 'tc_bf16x3')."""
    global _default_precision
    _default_precision = _PRECISIONS[name]


def default_precision():
    return _default_precision


def _stream(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _require_cuda(t, what):
    if not t.is_cuda:
        raise _lib.HonerfError("%s: honerf_b200 has no CPU path; tensor is on %s" % (what, t.device))


def _f32c(t):
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _round4(x):
    return (x + 3) // 4 * 4


# ------------------------------------------------------------------------------------------------
# packed (weight-normalised) MLP parameters
# ------------------------------------------------------------------------------------------------
class PackedMLP:
    """Effective weights of a stack of weight-normalised Linears, packed once per parameter
    version by hn_wn_pack (the reference recomputes g*v/||v|| inside every Linear call)."""

    def __init__(self, layers, post_scales, gaps=None, chain_kind=None):
        # layers: list of (weight_g [out,1], weight_v [out,in], bias [out]) parameters
        # gaps: per layer (gap_at, gap): `gap` zero columns inserted after input column gap_at of
        #       the packed layout (hand colour net, see include/honerf_b200.h)
        self.layers = layers
        self.post_scales = list(post_scales)
        self.raw_in = [v.shape[1] for (_, v, _) in layers]
        self.gaps = list(gaps) if gaps is not None else [(i, 0) for i in self.raw_in]
        self.dims = [(v.shape[1] + gp[1], v.shape[0]) for (_, v, _), gp in zip(layers, self.gaps)]
        self.lds = [_round4(i) for (i, _) in self.dims]
        self.offsets = []
        off = 0
        for (i, o), ld in zip(self.dims, self.lds):
            self.offsets.append(off)
            off += o * ld
        self.total = off
        self.ldTs = [_round4(o) for (_, o) in self.dims]
        self.offsetsT = []
        off = 0
        for (i, o), ldT in zip(self.dims, self.ldTs):
            self.offsetsT.append(off)
            off += i * ldT
        self.totalT = off
        self.WT = None
        self.W = None
        self.bias = None
        self.chain_kind = chain_kind        # 'sdf_obj' | 'color_obj' | 'bx3': also pack the HN_TC_BF16X3 operands
        self.chain = None
        self.struct = None
        self._key = None
        self.shared_token = None            # see shared_param_tokens()

    def _version_key(self):
        return tuple((p.data_ptr(), p._version) for layer in self.layers for p in layer)

    def get(self):
        key = self._version_key()
        if key == self._key:
            return self
        dev = self.layers[0][1].device
        _require_cuda(self.layers[0][1], "PackedMLP")
        if self.W is None or self.W.device != dev:
            self.W = torch.empty(self.total, device=dev, dtype=torch.float32)
            self.WT = torch.empty(self.totalT, device=dev, dtype=torch.float32)
        st = hn_mlp_t()
        st.n_layers = len(self.layers)
        keep = []
        jobs = (_lib.hn_wn_job_t * len(self.layers))()
        for l, (g, v, b) in enumerate(self.layers):
            gd, vd, bd = _f32c(g.detach()), _f32c(v.detach()), _f32c(b.detach())
            keep.append((gd, vd, bd))
            i, o = self.dims[l]
            Wl = self.W[self.offsets[l]: self.offsets[l] + o * self.lds[l]]
            WTl = self.WT[self.offsetsT[l]: self.offsetsT[l] + i * self.ldTs[l]]
            j = jobs[l]
            j.v, j.g, j.W, j.WT = vd.data_ptr(), gd.data_ptr(), Wl.data_ptr(), WTl.data_ptr()
            j.out_dim, j.in_dim, j.ld, j.ldT = o, self.raw_in[l], self.lds[l], self.ldTs[l]
            j.gap_at, j.gap, j.post_scale = self.gaps[l][0], self.gaps[l][1], self.post_scales[l]
            st.in_dim[l], st.out_dim[l], st.ld[l] = i, o, self.lds[l]
            st.W[l] = Wl.data_ptr()
            st.WT[l] = WTl.data_ptr()
            st.ldT[l] = self.ldTs[l]
            st.b[l] = bd.data_ptr()
        # every layer's g * v / ||v|| (and its transposed copy) in one launch
        check(lib.hn_wn_pack_batch(jobs, len(self.layers), _stream(self.W)), "hn_wn_pack_batch")
        if self.chain_kind in ("bx3", "sdf_hand", "color_hand"):
            # pre-packed bf16 hi/lo operands of every layer for the per-layer HN_TC_BF16X3 contractions (hand colour net), plus --
            # for the hand SDF net -- the tile-chain operands of its 256 x 256 layers (HN_TC_MIXED16)
            size_fn, pack_fn = {"bx3": (lib.hn_mlp_bx3_bytes, lib.hn_mlp_bx3_pack),
                                "sdf_hand": (lib.hn_sdf_hand_chain_bytes, lib.hn_sdf_hand_chain_pack),
                                "color_hand": (lib.hn_color_hand_chain_bytes, lib.hn_color_hand_chain_pack)}[self.chain_kind]
            nbytes = int(size_fn(ctypes.byref(st)))
            if self.chain is None or self.chain.device != dev or self.chain.numel() < nbytes:
                self.chain = torch.empty(nbytes, device=dev, dtype=torch.uint8)
            check(pack_fn(ctypes.byref(st), _ptr(self.chain), nbytes, _stream(self.W)), "hn_%s_pack" % self.chain_kind)
            st.chain = self.chain.data_ptr()
            st.chain_bytes = nbytes
        elif self.chain_kind is not None:
            size_fn, pack_fn = {"sdf_obj": (lib.hn_sdf_obj_chain_bytes, lib.hn_sdf_obj_chain_pack),
                                "color_obj": (lib.hn_color_obj_chain_bytes, lib.hn_color_obj_chain_pack)}[self.chain_kind]
            nbytes = int(size_fn())
            if self.chain is None or self.chain.device != dev:
                self.chain = torch.empty(nbytes, device=dev, dtype=torch.uint8)
            check(pack_fn(ctypes.byref(st), _ptr(self.chain), nbytes, _stream(self.W)), "hn_%s_chain_pack" % self.chain_kind)
            st.chain = self.chain.data_ptr()
            st.chain_bytes = nbytes
        self._keep = keep
        self.struct = st
        self._key = key
        return self

    def new_grad(self):
        """Zeroed packed gradient buffers + the struct pointing at them."""
        return self.new_grad_flat()[1:]

    def new_grad_flat(self):
        """(flat, dW, db, struct): `flat` is the ONE allocation dW and the per-layer db are carved from."""
        dev = self.W.device
        # ONE zero-filled allocation (one fill kernel) carved into the packed dW and the per-layer db
        n_b = sum(_round4(o) for (_, o) in self.dims)
        flat = torch.zeros(self.total + n_b, device=dev, dtype=torch.float32)
        dW = flat[:self.total]
        db, off = [], self.total
        for (_, o) in self.dims:
            db.append(flat[off:off + o])
            off += _round4(o)
        gs = hn_mlp_grad_t()
        for l in range(len(self.layers)):
            gs.dW[l] = dW[self.offsets[l]:].data_ptr()
            gs.db[l] = db[l].data_ptr()
        return flat, dW, db, gs

    def flat_grad_floats(self):
        return self.total + sum(_round4(o) for (_, o) in self.dims)

    def split_flat_grad(self, flat):
        """(dW, [db per layer]) views of a flat packed gradient laid out like new_grad()'s allocation."""
        dW = flat[:self.total]
        db, off = [], self.total
        for (_, o) in self.dims:
            db.append(flat[off:off + o])
            off += _round4(o)
        return dW, db

    def unpack_grads(self, dW, db):
        """(dg, dv, db) per layer from packed dW via hn_wn_bwd.  Returns a flat list in the order
        of ``flat_params``."""
        out = []
        jobs = (_lib.hn_wn_job_t * len(self._keep))()
        for l, (gd, vd, bd) in enumerate(self._keep):
            i, o = self.dims[l]
            dv = torch.empty_like(vd)
            dg = torch.empty(o, 1, device=vd.device, dtype=torch.float32)
            dWl = dW[self.offsets[l]: self.offsets[l] + o * self.lds[l]]
            j = jobs[l]
            j.v, j.g, j.dW, j.dv, j.dg = vd.data_ptr(), gd.data_ptr(), dWl.data_ptr(), dv.data_ptr(), dg.data_ptr()
            j.out_dim, j.in_dim, j.ld = o, self.raw_in[l], self.lds[l]
            j.gap_at, j.gap, j.post_scale = self.gaps[l][0], self.gaps[l][1], self.post_scales[l]
            out += [dg, dv, db[l]]
        check(lib.hn_wn_bwd_batch(jobs, len(self._keep), _stream(dW)), "hn_wn_bwd_batch")
        return out

    def flat_params(self):
        return [p for layer in self.layers for p in layer]


def _check_same_weights(ctx, what):
    """The backward reads the PackedMLP's shared W / WT / chain buffers and the raw parameter storage: if the parameters
    were updated in place (or the net re-packed for a new parameter version) between this call's forward and its
    backward, torch would raise a version-counter error for an ordinary op -- so do we."""
    if ctx.packed._key != ctx.pkey or ctx.packed._version_key() != ctx.pkey:
        raise _lib.HonerfError("%s: the network's parameters changed between forward and backward (in-place update or "
                               "re-pack); the saved activations no longer match the packed weights" % what)


class _ParamTokenFn(torch.autograd.Function):
    """One autograd edge for ALL parameters of a net: forward packs the weights (once per parameter version) and
    returns a token shaped like the flat packed gradient; every field call made with that token returns its flat
    packed gradient for it, autograd adds the flats of all consumers (one add per extra consumer instead of one per
    parameter), and this backward unpacks the sum to (dg, dv, db) per layer once (hn_wn_bwd_batch, one launch)."""

    @staticmethod
    def forward(ctx, packed, *params):
        pk = packed.get()
        ctx.packed = pk
        return torch.empty(pk.flat_grad_floats(), device=pk.W.device, dtype=torch.float32)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, flat):
        pk = ctx.packed
        dW, db = pk.split_flat_grad(_f32c(flat))
        return (None,) + tuple(pk.unpack_grads(dW, db))


class shared_param_tokens:
    """with shared_param_tokens(packed_a, packed_b): every sdf_obj / color_obj call inside shares ONE parameter edge
    per net.  Used when a batch is rendered as several ray shards on concurrent streams: the weights are packed once
    on the current stream BEFORE the shards fork, and the parameter gradients are unpacked once after they join."""

    def __init__(self, *packed):
        self.packed = packed

    def __enter__(self):
        for pk in self.packed:
            params = pk.flat_params()
            if torch.is_grad_enabled() and any(q.requires_grad for q in params):
                pk.shared_token = _ParamTokenFn.apply(pk, *params)
            else:
                pk.get()
        return self

    def __exit__(self, *exc):
        for pk in self.packed:
            pk.shared_token = None
        return False


# ------------------------------------------------------------------------------------------------
# object SDF field
# ------------------------------------------------------------------------------------------------
def sdf_obj_sdf_only(packed, pts, inv_scale=1.0, precision=None):
    """SDFNetwork_OBJ.sdf under no_grad (utils/fields.py:330-331): [N,3] -> [N,1]."""
    pts = _f32c(pts.detach())
    _require_cuda(pts, "sdf_obj_sdf_only")
    pk = packed.get()
    n = pts.shape[0]
    sdf = torch.empty(n, 1, device=pts.device, dtype=torch.float32)
    if n == 0:
        return sdf
    wsf = lib.hn_sdf_obj_ws_floats(n, HN_WS_SDF_ONLY)
    ws = torch.empty(wsf, device=pts.device, dtype=torch.float32)
    check(lib.hn_sdf_obj_sdf(ctypes.byref(pk.struct), _ptr(pts), n, inv_scale, _ptr(sdf), _ptr(ws), wsf,
                             _default_precision if precision is None else precision, _stream(pts)),
          "hn_sdf_obj_sdf")
    return sdf


def sdf_obj_lattice(packed, xs, ys, zs, inv_scale=1.0):
    """u[ix, iy, iz] = SDFNetwork_OBJ.sdf((xs[ix], ys[iy], zs[iz])) of extract_geometry (utils/renderer.py:262-278), one
    launch, lattice points generated in the kernel (hn_sdf_obj_grid)."""
    xs, ys, zs = _f32c(xs.detach()), _f32c(ys.detach()), _f32c(zs.detach())
    _require_cuda(xs, "sdf_obj_lattice")
    pk = packed.get()
    u = torch.empty(xs.numel(), ys.numel(), zs.numel(), device=xs.device, dtype=torch.float32)
    if u.numel():
        check(lib.hn_sdf_obj_grid(ctypes.byref(pk.struct), _ptr(xs), xs.numel(), _ptr(ys), ys.numel(), _ptr(zs), zs.numel(),
                                  inv_scale, _ptr(u), _stream(xs)), "hn_sdf_obj_grid")
    return u


class _SdfObjFn(torch.autograd.Function):
    """(sdf, feature, normal) = f(pts, params) with a second-order-aware backward."""

    @staticmethod
    def forward(ctx, pts, packed, inv_scale, precision, *params):
        pts_c = _f32c(pts.detach())
        _require_cuda(pts_c, "sdf_obj")
        pk = packed.get()
        n = pts_c.shape[0]
        dev = pts_c.device
        sdf = torch.empty(n, 1, device=dev, dtype=torch.float32)
        feat = torch.empty(n, 256, device=dev, dtype=torch.float32)
        normal = torch.empty(n, 3, device=dev, dtype=torch.float32)
        stf = lib.hn_sdf_obj_stash_floats(n)
        stash = torch.empty(max(stf, 4), device=dev, dtype=torch.float32)
        if n > 0:
            check(lib.hn_sdf_obj_fwd(ctypes.byref(pk.struct), _ptr(pts_c), n, inv_scale, _ptr(sdf), _ptr(feat),
                                     256, _ptr(normal), _ptr(stash), stf, None, 0, precision, _stream(pts_c)),
                  "hn_sdf_obj_fwd")
        ctx.packed, ctx.stash, ctx.n = pk, stash, n
        ctx.pkey = pk._key
        ctx.inv_scale, ctx.precision = inv_scale, precision
        ctx.struct = pk.struct
        ctx.pts_needs_grad = pts.requires_grad
        ctx.params_need_grad = any(p.requires_grad for p in params)
        ctx.token_mode = len(params) == 1 and len(pk.layers) * 3 != 1
        return sdf, feat, normal

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, d_sdf, d_feat, d_normal):
        n, pk = ctx.n, ctx.packed
        if ctx.stash is None:
            raise _lib.HonerfError("sdf_obj backward called twice (the stash is consumed)")
        _check_same_weights(ctx, "sdf_obj backward")
        dev = ctx.stash.device
        d_sdf = _f32c(d_sdf) if d_sdf is not None else None
        d_feat = _f32c(d_feat) if d_feat is not None else None
        d_normal = _f32c(d_normal) if d_normal is not None else torch.zeros(n, 3, device=dev)
        d_pts = torch.empty(n, 3, device=dev, dtype=torch.float32) if ctx.pts_needs_grad else None
        grads = [None] * (1 if ctx.token_mode else 3 * len(pk.layers))
        if n > 0:
            gs = None
            if ctx.params_need_grad:
                flat, dW, db, gs = pk.new_grad_flat()
            wsf = lib.hn_sdf_obj_ws_floats(n, HN_WS_BWD)
            ws = torch.empty(wsf, device=dev, dtype=torch.float32)
            check(lib.hn_sdf_obj_bwd(ctypes.byref(ctx.struct), n, ctx.inv_scale, _ptr(ctx.stash), _ptr(d_sdf),
                                     _ptr(d_feat), 256, _ptr(d_normal), _ptr(d_pts),
                                     ctypes.byref(gs) if gs is not None else None, _ptr(ws), wsf,
                                     ctx.precision, _stream(ctx.stash)), "hn_sdf_obj_bwd")
            if gs is not None:
                grads = [flat] if ctx.token_mode else pk.unpack_grads(dW, db)
        elif d_pts is not None:
            d_pts.zero_()
        ctx.stash = None
        return (d_pts, None, None, None) + tuple(grads)


def sdf_obj(packed, pts, inv_scale=1.0, precision=None):
    """Fused SDFNetwork_OBJ.forward + .gradient (utils/fields.py:316-347).
    Returns sdf [N,1], feature [N,256], normal [N,3]; differentiable w.r.t. pts and parameters."""
    precision = _default_precision if precision is None else precision
    if packed.shared_token is not None:
        return _SdfObjFn.apply(pts, packed, float(inv_scale), precision, packed.shared_token)
    return _SdfObjFn.apply(pts, packed, float(inv_scale), precision, *packed.flat_params())


# ------------------------------------------------------------------------------------------------
# object colour field
# ------------------------------------------------------------------------------------------------
class _ColorObjFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pts, dirs, feat, normal, packed, precision, *params):
        pts_c, dirs_c = _f32c(pts.detach()), _f32c(dirs.detach())
        feat_c, nrm_c = _f32c(feat.detach()), _f32c(normal.detach())
        _require_cuda(pts_c, "color_obj")
        pk = packed.get()
        n, dev = pts_c.shape[0], pts_c.device
        rgb = torch.empty(n, 3, device=dev, dtype=torch.float32)
        stf = lib.hn_color_obj_stash_floats(n)
        stash = torch.empty(max(stf, 4), device=dev, dtype=torch.float32)
        if n > 0:
            check(lib.hn_color_obj_fwd(ctypes.byref(pk.struct), _ptr(pts_c), _ptr(dirs_c), _ptr(feat_c),
                                       feat_c.shape[1], _ptr(nrm_c), n, _ptr(rgb), _ptr(stash), stf, precision,
                                       _stream(pts_c)), "hn_color_obj_fwd")
        ctx.packed, ctx.stash, ctx.n, ctx.precision, ctx.struct = pk, stash, n, precision, pk.struct
        ctx.pkey = pk._key
        ctx.rgb = rgb
        ctx.need = (pts.requires_grad, dirs.requires_grad, feat.requires_grad, normal.requires_grad)
        ctx.params_need_grad = any(p.requires_grad for p in params)
        ctx.token_mode = len(params) == 1 and len(pk.layers) * 3 != 1
        return rgb

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, d_rgb):
        if ctx.stash is None:
            raise _lib.HonerfError("color_obj backward called twice (the stash is consumed)")
        _check_same_weights(ctx, "color_obj backward")
        n, pk, dev = ctx.n, ctx.packed, ctx.stash.device
        d_rgb = _f32c(d_rgb)
        d_pts = torch.empty(n, 3, device=dev) if ctx.need[0] else None
        d_dirs = torch.empty(n, 3, device=dev) if ctx.need[1] else None
        d_feat = torch.empty(n, 256, device=dev) if ctx.need[2] else None
        d_nrm = torch.empty(n, 3, device=dev) if ctx.need[3] else None
        grads = [None] * (1 if ctx.token_mode else 3 * len(pk.layers))
        if n > 0:
            gs = None
            if ctx.params_need_grad:
                flat, dW, db, gs = pk.new_grad_flat()
            wsf = lib.hn_color_obj_ws_floats(n, HN_WS_BWD)
            ws = torch.empty(wsf, device=dev, dtype=torch.float32)
            check(lib.hn_color_obj_bwd(ctypes.byref(ctx.struct), n, _ptr(ctx.stash), _ptr(ctx.rgb), _ptr(d_rgb),
                                       _ptr(d_pts), _ptr(d_dirs), _ptr(d_feat), 256, _ptr(d_nrm),
                                       ctypes.byref(gs) if gs is not None else None, _ptr(ws), wsf,
                                       ctx.precision, _stream(ctx.stash)), "hn_color_obj_bwd")
            if gs is not None:
                grads = [flat] if ctx.token_mode else pk.unpack_grads(dW, db)
        ctx.stash = None
        return (d_pts, d_dirs, d_feat, d_nrm, None, None) + tuple(grads)


def color_obj(packed, pts, dirs, feat, normal, precision=None):
    """RenderingNetwork_OBJ.forward (utils/fields.py:387-405): -> rgb [N,3]."""
    precision = _default_precision if precision is None else precision
    if packed.shared_token is not None:
        return _ColorObjFn.apply(pts, dirs, feat, normal, packed, precision, packed.shared_token)
    return _ColorObjFn.apply(pts, dirs, feat, normal, packed, precision, *packed.flat_params())


# ------------------------------------------------------------------------------------------------
# hand SDF / colour fields
# ------------------------------------------------------------------------------------------------
def _hand_pose_args(pts, bt_inv, T_pose_21):
    """Flatten (possibly frame-batched) hand inputs: pts [N,3] | [F,P,3], bt_inv [21,4,4] | [F,21,4,4],
    T_pose_21 [21,3] | [F,21,3]  ->  pts [n,3], bt [F,21,4,4], T [F,21,3], points per frame."""
    if pts.dim() == 3:
        ppf = pts.shape[1]
        pts2 = pts.reshape(-1, 3)
    else:
        pts2, ppf = pts, max(int(pts.shape[0]), 1)
    bt = bt_inv if bt_inv.dim() == 4 else bt_inv[None]
    T = T_pose_21 if T_pose_21.dim() == 3 else T_pose_21[None]
    if pts.dim() == 2 and bt.shape[0] != 1:
        raise ValueError("un-batched points need a single [21,4,4] bone transform set")
    return pts2, bt, T, ppf


def sdf_hand_sdf_only(packed, pts, bt_inv, T_pose_21, precision=None):
    """SDFNetwork.sdf under no_grad (utils/fields.py:158-160)."""
    pts2, bt, T, ppf = _hand_pose_args(pts, bt_inv, T_pose_21)
    pts2, bt, T = _f32c(pts2.detach()), _f32c(bt.detach()), _f32c(T.detach())
    _require_cuda(pts2, "sdf_hand_sdf_only")
    pk = packed.get()
    n = pts2.shape[0]
    sdf = torch.empty(n, 1, device=pts2.device, dtype=torch.float32)
    if n == 0:
        return sdf
    wsf = lib.hn_sdf_hand_ws_floats(n, HN_WS_SDF_ONLY)
    ws = torch.empty(wsf, device=pts2.device, dtype=torch.float32)
    check(lib.hn_sdf_hand_sdf(ctypes.byref(pk.struct), _ptr(pts2), _ptr(bt), _ptr(T), n, ppf, _ptr(sdf), _ptr(ws), wsf,
                              _default_precision if precision is None else precision, _stream(pts2)),
          "hn_sdf_hand_sdf")
    return sdf


_HAND_ROW_LD = 1644      # csrc/fields_hand.cu HROW_LD


def _rows_f32(t):
    """(tensor, leading dimension) of a 2-D fp32 tensor whose rows are contiguous and 16-byte aligned (a column slice of a
    wider buffer is passed as it is); anything else is made contiguous first."""
    t = t.detach()
    if (t.dtype == torch.float32 and t.dim() == 2 and t.stride(1) == 1 and t.stride(0) % 4 == 0 and t.stride(0) >= t.shape[1]
            and t.data_ptr() % 16 == 0):
        return t, t.stride(0)
    t = _f32c(t)
    return t, t.shape[1]


class _SdfHandFn(torch.autograd.Function):
    """(sdf, feature, normal, xyz_feature) = f(pts, bt_inv, T_pose_21, params)."""

    @staticmethod
    def forward(ctx, pts, bt_inv, T_pose, ppf, packed, precision, differentiable, *params):
        pts_c, bt_c, T_c = _f32c(pts.detach()), _f32c(bt_inv.detach()), _f32c(T_pose.detach())
        _require_cuda(pts_c, "sdf_hand")
        pk = packed.get()
        ctx.set_materialize_grads(False)      # an output nobody differentiates arrives as None, not as a zero tensor
        ctx.params_need_grad = any(p.requires_grad for p in params)
        n, dev = pts_c.shape[0], pts_c.device
        sdf = torch.empty(n, 1, device=dev)
        feat = torch.empty(n, 256, device=dev)
        normal = torch.empty(n, 3, device=dev)
        stf = lib.hn_sdf_hand_stash_floats(n)
        stash = torch.empty(max(stf, 4), device=dev, dtype=torch.float32)
        # xyz_feature is a view of the stash's skip-input rows [h3 256 | feature 1386 | pad 2] (ld 1644): no copy; the
        # backward only reads that part of the stash
        xyz = stash[:n * _HAND_ROW_LD].view(n, _HAND_ROW_LD)[:, 256:256 + 1386]
        if n > 0:
            fwd = lib.hn_sdf_hand_fwd if differentiable else lib.hn_sdf_hand_fwd_render
            check(fwd(ctypes.byref(pk.struct), _ptr(pts_c), _ptr(bt_c), _ptr(T_c), n, ppf, _ptr(sdf), _ptr(feat), 256, _ptr(normal),
                      None, 0, _ptr(stash), stf, precision, _stream(pts_c)), "hn_sdf_hand_fwd")
        ctx.packed, ctx.stash, ctx.n, ctx.ppf, ctx.precision, ctx.struct = pk, stash, n, ppf, precision, pk.struct
        ctx.pkey = pk._key
        ctx.pts_c, ctx.bt_c, ctx.T_c = pts_c, bt_c, T_c
        ctx.need = (pts.requires_grad, bt_inv.requires_grad, T_pose.requires_grad)
        return sdf, feat, normal, xyz

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, d_sdf, d_feat, d_normal, d_xyz):
        if ctx.stash is None:
            raise _lib.HonerfError("sdf_hand backward called twice (the stash is consumed)")
        _check_same_weights(ctx, "sdf_hand backward")
        n, pk, dev = ctx.n, ctx.packed, ctx.stash.device
        d_sdf = _f32c(d_sdf) if d_sdf is not None else None
        d_feat = _f32c(d_feat) if d_feat is not None else None
        d_xyz, ld_dxyz = _rows_f32(d_xyz) if d_xyz is not None else (None, 1386)
        d_normal = _f32c(d_normal) if d_normal is not None else torch.zeros(n, 3, device=dev)
        d_pts = torch.empty(n, 3, device=dev) if ctx.need[0] else None
        d_bt = torch.zeros_like(ctx.bt_c) if ctx.need[1] else None
        d_T = torch.zeros_like(ctx.T_c) if ctx.need[2] else None
        grads = [None] * (3 * len(pk.layers))
        if n > 0:
            gs = None
            if ctx.params_need_grad:
                dW, db, gs = pk.new_grad()
            wsf = lib.hn_sdf_hand_ws_floats(n, HN_WS_BWD)
            ws = torch.empty(wsf, device=dev, dtype=torch.float32)
            check(lib.hn_sdf_hand_bwd(ctypes.byref(ctx.struct), _ptr(ctx.pts_c), _ptr(ctx.bt_c), _ptr(ctx.T_c), n, ctx.ppf,
                                      _ptr(ctx.stash), _ptr(d_sdf), _ptr(d_feat), 256, _ptr(d_normal), _ptr(d_xyz), ld_dxyz,
                                      _ptr(d_pts), _ptr(d_bt), _ptr(d_T), ctypes.byref(gs) if gs is not None else None,
                                      _ptr(ws), wsf, ctx.precision, _stream(ctx.stash)), "hn_sdf_hand_bwd")
            if gs is not None:
                grads = pk.unpack_grads(dW, db)
        elif d_pts is not None:
            d_pts.zero_()
        ctx.stash = None
        return (d_pts, d_bt, d_T, None, None, None, None) + tuple(grads)


def sdf_hand(packed, pts, bt_inv, T_pose_21, precision=None):
    """Fused anerf_emb_point(_batch) + SDFNetwork.forward + .gradient (utils/fields.py:22-52, 132-177).
    Returns sdf [N,1], feature [N,256], normal [N,3], xyz_feature [N,1386]."""
    precision = _default_precision if precision is None else precision
    params = packed.flat_params()
    if precision == _lib.HN_TC_MIXED16 and torch.is_grad_enabled() and any(p.requires_grad for p in params):
        # the tile-chain kernels of the hand net keep a 16-bit stash and compute no weight gradients (pose fitting, rendering);
        # a call whose backward trains the net stays on the per-layer contractions with their fp32 stash
        precision = _lib.HN_TC_BF16X3
    pts2, bt, T, ppf = _hand_pose_args(pts, bt_inv, T_pose_21)
    # nothing to differentiate (rendering): the operator skips what only its backward would read
    differentiable = torch.is_grad_enabled() and (pts2.requires_grad or bt.requires_grad or T.requires_grad or
                                                  any(p.requires_grad for p in params))
    return _SdfHandFn.apply(pts2, bt, T, ppf, packed, precision, differentiable, *params)


class _ColorHandFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz, feat, normal, packed, precision, differentiable, *params):
        (xyz_c, ld_xyz), feat_c, nrm_c = _rows_f32(xyz), _f32c(feat.detach()), _f32c(normal.detach())
        _require_cuda(xyz_c, "color_hand")
        pk = packed.get()
        n, dev = xyz_c.shape[0], xyz_c.device
        rgb = torch.empty(n, 3, device=dev)
        stf = lib.hn_color_hand_stash_floats(n)
        stash = torch.empty(max(stf, 4), device=dev, dtype=torch.float32)
        if n > 0:
            fwd = lib.hn_color_hand_fwd if differentiable else lib.hn_color_hand_fwd_render
            check(fwd(ctypes.byref(pk.struct), _ptr(xyz_c), ld_xyz, _ptr(feat_c), feat_c.shape[1], _ptr(nrm_c), n, _ptr(rgb),
                      _ptr(stash), stf, precision, _stream(xyz_c)), "hn_color_hand_fwd")
        ctx.packed, ctx.stash, ctx.n, ctx.precision, ctx.struct, ctx.rgb = pk, stash, n, precision, pk.struct, rgb
        ctx.pkey = pk._key
        ctx.need = (xyz.requires_grad, feat.requires_grad, normal.requires_grad)
        ctx.params_need_grad = any(p.requires_grad for p in params)
        return rgb

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, d_rgb):
        if ctx.stash is None:
            raise _lib.HonerfError("color_hand backward called twice (the stash is consumed)")
        _check_same_weights(ctx, "color_hand backward")
        n, pk, dev = ctx.n, ctx.packed, ctx.stash.device
        d_rgb = _f32c(d_rgb)
        # rows padded to 1388 floats: the hand SDF backward takes this buffer as a 16-byte aligned aux operand
        d_xyz = torch.empty(n, 1388, device=dev)[:, :1386] if ctx.need[0] else None
        d_feat = torch.empty(n, 256, device=dev) if ctx.need[1] else None
        d_nrm = torch.empty(n, 3, device=dev) if ctx.need[2] else None
        grads = [None] * (3 * len(pk.layers))
        if n > 0:
            gs = None
            if ctx.params_need_grad:
                dW, db, gs = pk.new_grad()
            wsf = lib.hn_color_hand_ws_floats(n, HN_WS_BWD)
            ws = torch.empty(wsf, device=dev, dtype=torch.float32)
            check(lib.hn_color_hand_bwd(ctypes.byref(ctx.struct), n, _ptr(ctx.stash), _ptr(ctx.rgb), _ptr(d_rgb),
                                        _ptr(d_xyz), 1388, _ptr(d_feat), 256, _ptr(d_nrm),
                                        ctypes.byref(gs) if gs is not None else None, _ptr(ws), wsf, ctx.precision,
                                        _stream(ctx.stash)), "hn_color_hand_bwd")
            if gs is not None:
                grads = pk.unpack_grads(dW, db)
        ctx.stash = None
        return (d_xyz, d_feat, d_nrm, None, None, None) + tuple(grads)


def color_hand(packed, xyz_feature, feat, normal, precision=None):
    """RenderingNetwork.forward (utils/fields.py:222-240): -> rgb [N,3]."""
    precision = _default_precision if precision is None else precision
    params = packed.flat_params()
    differentiable = torch.is_grad_enabled() and (xyz_feature.requires_grad or feat.requires_grad or normal.requires_grad or
                                                  any(p.requires_grad for p in params))
    return _ColorHandFn.apply(xyz_feature, feat, normal, packed, precision, differentiable, *params)


# ------------------------------------------------------------------------------------------------
# rays and hierarchical sampling (all under no_grad in the reference)
# ------------------------------------------------------------------------------------------------
def ray_points(rays_o, rays_d, z):
    """[B,3],[B,3],[B,n] -> [B*n,3], bit-exact with rays_o[:,None]+rays_d[:,None]*z[...,None]."""
    o, d, z = _f32c(rays_o.detach()), _f32c(rays_d.detach()), _f32c(z.detach())
    _require_cuda(z, "ray_points")
    B, n = z.shape
    pts = torch.empty(B * n, 3, device=z.device, dtype=torch.float32)
    check(lib.hn_ray_points(_ptr(o), _ptr(d), _ptr(z), B, n, _ptr(pts), _stream(z)), "hn_ray_points")
    return pts


class _RaysToLocalFn(torch.autograd.Function):
    """o' = Ro (o - To), d' = Ro d (utils/renderer.py:180-188) with gradients to Ro, To (and the rays if needed)."""

    @staticmethod
    def forward(ctx, rays_o, rays_d, Ro, To):
        o, d = _f32c(rays_o.detach()), _f32c(rays_d.detach())
        R, T = _f32c(Ro.detach()), _f32c(To.detach())
        _require_cuda(o, "rays_to_local")
        _require_cuda(R, "rays_to_local")
        B = o.shape[0]
        lo, ld = torch.empty(B, 3, device=o.device), torch.empty(B, 3, device=o.device)
        check(lib.hn_rays_to_local(_ptr(o), _ptr(d), _ptr(R), _ptr(T), B, _ptr(lo), _ptr(ld), _stream(o)),
              "hn_rays_to_local")
        ctx.save_for_backward(o, d, R, T)
        ctx.need_rays = (rays_o.requires_grad, rays_d.requires_grad)
        ctx.shapes = (Ro.shape, To.shape)
        return lo, ld

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_lo, g_ld):
        o, d, R, T = ctx.saved_tensors
        B = o.shape[0]
        g_lo = _f32c(g_lo) if g_lo is not None else None
        g_ld = _f32c(g_ld) if g_ld is not None else None
        d_R, d_T = torch.empty(9, device=o.device), torch.empty(3, device=o.device)
        d_o = torch.empty(B, 3, device=o.device) if ctx.need_rays[0] else None
        d_d = torch.empty(B, 3, device=o.device) if ctx.need_rays[1] else None
        check(lib.hn_rays_to_local_bwd(_ptr(g_lo), _ptr(g_ld), _ptr(o), _ptr(d), _ptr(R), _ptr(T), B, _ptr(d_R), _ptr(d_T),
                                       _ptr(d_o), _ptr(d_d), _stream(o)), "hn_rays_to_local_bwd")
        return d_o, d_d, d_R.reshape(ctx.shapes[0]), d_T.reshape(ctx.shapes[1])


def rays_to_local(rays_o, rays_d, Ro, To):
    """[B,3], [B,3], Ro [3,3], To [3] -> rays in the object frame; one launch forward, one backward."""
    return _RaysToLocalFn.apply(rays_o, rays_d, Ro, To)


_u_cache = {}


def _u_samples(n, device):
    key = (n, str(device))
    if key not in _u_cache:
        # generated by torch on the host so the values are the reference's (utils/renderer.py:19)
        _u_cache[key] = torch.linspace(0.0 + 0.5 / n, 1.0 - 0.5 / n, steps=n).to(device)
    return _u_cache[key]


def up_sample(z_vals, sdf, n_importance, inv_s):
    """NeuSRenderer.up_sample (utils/renderer.py:60-86): [B,m],[B,m] -> [B,n_importance]."""
    z, s = _f32c(z_vals.detach()), _f32c(sdf.detach()).reshape(z_vals.shape)
    _require_cuda(z, "up_sample")
    B, m = z.shape
    out = torch.empty(B, n_importance, device=z.device, dtype=torch.float32)
    u = _u_samples(n_importance, z.device)
    check(lib.hn_up_sample(_ptr(z), _ptr(s), _ptr(u), B, m, n_importance, float(inv_s), _ptr(out), _stream(z)),
          "hn_up_sample")
    return out


def inverse_cdf(bins, cdf, n_samples, return_indices=False):
    bins, cdf = _f32c(bins), _f32c(cdf)
    _require_cuda(cdf, "inverse_cdf")
    B, m = cdf.shape
    out = torch.empty(B, n_samples, device=cdf.device, dtype=torch.float32)
    below = above = None
    if return_indices:
        below = torch.empty(B, n_samples, device=cdf.device, dtype=torch.int64)
        above = torch.empty_like(below)
    u = _u_samples(n_samples, cdf.device)
    check(lib.hn_inverse_cdf(_ptr(bins), _ptr(cdf), _ptr(u), B, m, n_samples, _ptr(out), _ptr(below), _ptr(above),
                             _stream(cdf)), "hn_inverse_cdf")
    return (out, below, above) if return_indices else out


def merge_sorted(z_a, z_b, sdf_a=None, sdf_b=None, sdf_row_mod=0, return_index=False):
    """cat_z_vals' sort + gather (utils/renderer.py:88-105) for two sorted rows."""
    za, zb = _f32c(z_a.detach()), _f32c(z_b.detach())
    _require_cuda(za, "merge_sorted")
    B, m = za.shape
    k = zb.shape[1]
    zo = torch.empty(B, m + k, device=za.device, dtype=torch.float32)
    idx = torch.empty(B, m + k, device=za.device, dtype=torch.int64) if return_index else None
    so = sa = sb = None
    if sdf_a is not None:
        sa, sb = _f32c(sdf_a.detach()).reshape(B, m), _f32c(sdf_b.detach()).reshape(B, k)
        so = torch.empty(B, m + k, device=za.device, dtype=torch.float32)
    check(lib.hn_merge_sorted(_ptr(za), m, _ptr(zb), k, B, _ptr(zo), _ptr(idx), _ptr(sa), _ptr(sb),
                              int(sdf_row_mod), _ptr(so), _stream(za)), "hn_merge_sorted")
    return zo, so, idx


def sort_rows(x, return_index=False):
    x = _f32c(x.detach())
    _require_cuda(x, "sort_rows")
    B, n = x.shape
    out = torch.empty_like(x)
    idx = torch.empty(B, n, device=x.device, dtype=torch.int64) if return_index else None
    check(lib.hn_sort_rows(_ptr(x), B, n, _ptr(out), _ptr(idx), _stream(x)), "hn_sort_rows")
    return (out, idx) if return_index else out


class _MidPointsFn(torch.autograd.Function):
    """pts = o + d * (z + dists/2) and dirs = expand(d) (utils/renderer.py:119-127); z is a constant (no_grad)."""

    @staticmethod
    def forward(ctx, rays_o, rays_d, z, sample_dist, with_dirs):
        o, d, zc = _f32c(rays_o.detach()), _f32c(rays_d.detach()), _f32c(z.detach())
        _require_cuda(zc, "mid_points")
        B, n = zc.shape
        pts = torch.empty(B * n, 3, device=zc.device, dtype=torch.float32)
        dists = torch.empty(B, n, device=zc.device, dtype=torch.float32)
        dirs = torch.empty(B * n, 3, device=zc.device, dtype=torch.float32) if with_dirs else None
        check(lib.hn_mid_points(_ptr(o), _ptr(d), _ptr(zc), B, n, float(sample_dist), _ptr(pts), _ptr(dists),
                                _ptr(dirs), _stream(zc)), "hn_mid_points")
        ctx.save_for_backward(zc, dists)
        ctx.mark_non_differentiable(dists)
        if with_dirs:
            return pts, dists, dirs
        return pts, dists

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, d_pts, _d_dists, d_dirs=None):
        zc, dists = ctx.saved_tensors
        B, n = zc.shape
        if d_pts is None and d_dirs is None:
            return None, None, None, None, None
        g = _f32c(d_pts) if d_pts is not None else torch.zeros(B * n, 3, device=zc.device)
        gd = _f32c(d_dirs) if d_dirs is not None else None
        d_o = torch.empty(B, 3, device=zc.device, dtype=torch.float32)
        d_d = torch.empty(B, 3, device=zc.device, dtype=torch.float32)
        check(lib.hn_mid_points_bwd(_ptr(g), _ptr(gd), _ptr(zc), _ptr(dists), B, n, _ptr(d_o), _ptr(d_d), _stream(zc)),
              "hn_mid_points_bwd")
        return d_o, d_d, None, None, None


def mid_points(rays_o, rays_d, z, sample_dist, with_dirs=False):
    """-> pts [B*n,3], dists [B,n] (and, with_dirs, the expanded view directions dirs [B*n,3], whose cotangent is
    folded into d_rays_d by the same backward kernel)."""
    return _MidPointsFn.apply(rays_o, rays_d, z, sample_dist, bool(with_dirs))


# ------------------------------------------------------------------------------------------------
# compositing
# ------------------------------------------------------------------------------------------------
class _CompositeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sdf, normal, rgb, dists, rays_d, variance, seed_with_c0):
        sdf_c, nrm_c, rgb_c = _f32c(sdf.detach()), _f32c(normal.detach()), _f32c(rgb.detach())
        dist_c, d_c, var_c = _f32c(dists.detach()), _f32c(rays_d.detach()), _f32c(variance.detach()).reshape(1)
        _require_cuda(sdf_c, "neus_composite")
        B, n = dist_c.shape
        dev = sdf_c.device
        weights = torch.empty(B, n, device=dev)
        cdf = torch.empty(B, n, device=dev)
        color = torch.empty(B, 3, device=dev)
        wsum = torch.empty(B, 1, device=dev)
        wmax = torch.empty(B, 1, device=dev)
        eik = torch.empty(B, device=dev)
        check(lib.hn_neus_composite_fwd(_ptr(sdf_c), _ptr(nrm_c), _ptr(rgb_c), _ptr(dist_c), _ptr(d_c), _ptr(var_c),
                                        B, n, int(seed_with_c0), _ptr(weights), _ptr(cdf), None, _ptr(color),
                                        _ptr(wsum), _ptr(wmax), _ptr(eik), _stream(sdf_c)), "hn_neus_composite_fwd")
        ctx.save_for_backward(sdf_c, nrm_c, rgb_c, dist_c, d_c, var_c, weights)
        ctx.seed = int(seed_with_c0)
        ctx.need_d = rays_d.requires_grad
        ctx.var_shape = variance.shape
        ctx.rgb_shape = rgb.shape
        ctx.mark_non_differentiable(cdf, wmax)
        return color, weights, cdf, wsum, wmax, eik

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, d_color, d_weights, _d_cdf, d_wsum, _d_wmax, d_eik):
        sdf_c, nrm_c, rgb_c, dist_c, d_c, var_c, weights = ctx.saved_tensors
        B, n = dist_c.shape
        dev = sdf_c.device
        d_color = _f32c(d_color) if d_color is not None else torch.zeros(B, 3, device=dev)
        d_weights = _f32c(d_weights) if d_weights is not None else None
        d_wsum = _f32c(d_wsum) if d_wsum is not None else None
        d_eik = _f32c(d_eik) if d_eik is not None else None
        d_sdf = torch.empty(B * n, 1, device=dev)
        d_nrm = torch.empty(B * n, 3, device=dev)
        d_rgb = torch.empty(B * n, 3, device=dev)
        d_rd = torch.empty(B, 3, device=dev) if ctx.need_d else None
        d_var = torch.zeros(1, device=dev)
        check(lib.hn_neus_composite_bwd(_ptr(sdf_c), _ptr(nrm_c), _ptr(rgb_c), _ptr(dist_c), _ptr(d_c), _ptr(var_c),
                                        _ptr(weights), B, n, ctx.seed, _ptr(d_color), _ptr(d_wsum), _ptr(d_weights),
                                        _ptr(d_eik), _ptr(d_sdf), _ptr(d_nrm), _ptr(d_rgb), _ptr(d_rd), _ptr(d_var),
                                        _stream(sdf_c)), "hn_neus_composite_bwd")
        return d_sdf, d_nrm, d_rgb.reshape(ctx.rgb_shape), None, d_rd, d_var.reshape(ctx.var_shape), None


def neus_composite(sdf, normal, rgb, dists, rays_d, variance, seed_with_c0=True):
    """alpha / transmittance / compositing of render_core (utils/renderer.py:144-169).
    sdf [B*n,1], normal [B*n,3], rgb [B*n,3] or [B,n,3], dists [B,n], rays_d [B,3], variance scalar.
    Returns color [B,3], weights [B,n], cdf [B,n], weight_sum [B,1], weight_max [B,1],
    eik [B] (per-ray sums of (||n||-1)^2)."""
    return _CompositeFn.apply(sdf, normal, rgb, dists, rays_d, variance, seed_with_c0)


class _AlphaFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sdf, normal, dists, rays_d, variance):
        sdf_c, nrm_c = _f32c(sdf.detach()), _f32c(normal.detach())
        dist_c, d_c, var_c = _f32c(dists.detach()), _f32c(rays_d.detach()), _f32c(variance.detach()).reshape(1)
        _require_cuda(sdf_c, "neus_alpha")
        B, n = dist_c.shape
        alpha = torch.empty(B, n, device=sdf_c.device)
        eik = torch.empty(B, device=sdf_c.device)
        check(lib.hn_neus_alpha_fwd(_ptr(sdf_c), _ptr(nrm_c), _ptr(dist_c), _ptr(d_c), _ptr(var_c), B, n, _ptr(alpha),
                                    _ptr(eik), _stream(sdf_c)), "hn_neus_alpha_fwd")
        ctx.save_for_backward(sdf_c, nrm_c, dist_c, d_c, var_c)
        ctx.need_d = rays_d.requires_grad
        ctx.shapes = (sdf.shape, normal.shape, rays_d.shape, variance.shape)
        return alpha, eik

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, d_alpha, d_eik):
        sdf_c, nrm_c, dist_c, d_c, var_c = ctx.saved_tensors
        B, n = dist_c.shape
        dev = sdf_c.device
        d_alpha = _f32c(d_alpha) if d_alpha is not None else None
        d_eik = _f32c(d_eik) if d_eik is not None else None
        d_sdf = torch.empty(B * n, 1, device=dev)
        d_nrm = torch.empty(B * n, 3, device=dev)
        d_rd = torch.empty(B, 3, device=dev) if ctx.need_d else None
        d_var = torch.zeros(1, device=dev)
        check(lib.hn_neus_alpha_bwd(_ptr(sdf_c), _ptr(nrm_c), _ptr(dist_c), _ptr(d_c), _ptr(var_c), B, n, _ptr(d_alpha),
                                    _ptr(d_eik), _ptr(d_sdf), _ptr(d_nrm), _ptr(d_rd), _ptr(d_var), _stream(sdf_c)),
              "hn_neus_alpha_bwd")
        s_sdf, s_nrm, s_rd, s_var = ctx.shapes
        return (d_sdf.reshape(s_sdf), d_nrm.reshape(s_nrm), None, d_rd.reshape(s_rd) if d_rd is not None else None,
                d_var.reshape(s_var))


def neus_alpha(sdf, normal, dists, rays_d, variance):
    """alpha of get_alpha_sample_color (utils/renderer.py:396-420): sdf [B*n,1], normal [B*n,3],
    dists [B,n], rays_d [B,3] -> alpha [B,n], eik [B] (per-ray sums of (||n||-1)^2)."""
    return _AlphaFn.apply(sdf, normal, dists, rays_d, variance)


class _FitCompositeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, alpha_h, rgb_h, alpha_o, rgb_o):
        ah, ao = _f32c(alpha_h.detach()), _f32c(alpha_o.detach())
        ch, co = _f32c(rgb_h.detach()), _f32c(rgb_o.detach())
        _require_cuda(ah, "fit_composite")
        B, n = ah.shape
        dev = ah.device
        trans = torch.empty(B, n, device=dev)
        color = torch.empty(B, 3, device=dev)
        wsum = torch.empty(B, 1, device=dev)
        check(lib.hn_fit_composite_fwd(_ptr(ah), _ptr(ch), _ptr(ao), _ptr(co), B, n, _ptr(trans), _ptr(color),
                                       _ptr(wsum), _stream(ah)), "hn_fit_composite_fwd")
        ctx.save_for_backward(ah, ch, ao, co, trans)
        ctx.rgb_shapes = (rgb_h.shape, rgb_o.shape)
        return color, wsum

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, d_color, d_wsum):
        ah, ch, ao, co, trans = ctx.saved_tensors
        B, n = ah.shape
        dev = ah.device
        d_color = _f32c(d_color) if d_color is not None else None
        d_wsum = _f32c(d_wsum) if d_wsum is not None else None
        d_ah, d_ao = torch.empty(B, n, device=dev), torch.empty(B, n, device=dev)
        d_ch, d_co = torch.empty(B * n, 3, device=dev), torch.empty(B * n, 3, device=dev)
        check(lib.hn_fit_composite_bwd(_ptr(ah), _ptr(ch), _ptr(ao), _ptr(co), _ptr(trans), B, n, _ptr(d_color),
                                       _ptr(d_wsum), _ptr(d_ah), _ptr(d_ch), _ptr(d_ao), _ptr(d_co), _stream(ah)),
              "hn_fit_composite_bwd")
        return d_ah, d_ch.reshape(ctx.rgb_shapes[0]), d_ao, d_co.reshape(ctx.rgb_shapes[1])


def fit_composite(alpha_h, rgb_h, alpha_o, rgb_o):
    """Joint two-field compositing (utils/renderer.py:512-524): alpha [B,n], rgb [B,n,3] (or [B*n,3])
    -> color [B,3], weight_sum [B,1]."""
    return _FitCompositeFn.apply(alpha_h, rgb_h, alpha_o, rgb_o)


SQRT1_2 = 1.0 / math.sqrt(2.0)


# ------------------------------------------------------------------------------------------------
# ray generation
# ------------------------------------------------------------------------------------------------
def pack_cameras(R, T, focal_length, principal_point):
    """pytorch3d PerspectiveCameras arguments ([n,3,3], [n,3], [n,2], [n,2]; a missing batch axis is added) ->
    [n,16] device floats  R | T | fx fy | px py  (the camera record of include/honerf_b200.h)."""
    R = R.reshape(-1, 3, 3)
    n = R.shape[0]
    parts = [_f32c(R).reshape(n, 9), _f32c(T).reshape(n, 3), _f32c(focal_length).reshape(-1, 2).expand(n, 2),
             _f32c(principal_point).reshape(-1, 2).expand(n, 2)]
    return torch.cat(parts, dim=1).contiguous()


def rays_from_ndc(xy, cams):
    """utils/utils.py:31-115 (_xy_to_ray_bundle, unit directions): xy [n_cams, ..., 2] NDC, cams [n_cams,16]
    -> origins, directions [n_cams, ..., 3]."""
    xy_c, cams_c = _f32c(xy.detach()), _f32c(cams.detach())
    _require_cuda(xy_c, "rays_from_ndc")
    _require_cuda(cams_c, "rays_from_ndc")
    n_cams = cams_c.shape[0]
    if xy_c.shape[0] != n_cams or xy_c.shape[-1] != 2 or cams_c.shape[-1] != 16:
        raise _lib.HonerfError("rays_from_ndc: xy %s does not match cams %s" % (tuple(xy.shape), tuple(cams.shape)))
    per = xy_c[0].numel() // 2 if n_cams else 0
    o = torch.empty(*xy_c.shape[:-1], 3, device=xy_c.device)
    d = torch.empty_like(o)
    check(lib.hn_rays_from_ndc(_ptr(xy_c), _ptr(cams_c), n_cams, per, _ptr(o), _ptr(d), _stream(xy_c)),
          "hn_rays_from_ndc")
    return o, d


def ndc_grid_axes(H, W, device):
    """The reference's full-image NDC axes (exp_runner.py:338-348), made by torch so they are its values."""
    if W >= H:
        range_x, range_y = W / H, 1.0
    else:
        range_x, range_y = 1.0, H / W
    return (torch.linspace(range_x, -range_x, W).to(device), torch.linspace(range_y, -range_y, H).to(device))


def rays_ndc_grid(xs, ys, cam, first, count):
    """Rays of pixels [first, first+count) of the H x W image whose NDC axes are xs [W], ys [H] (pixel = row*W + col):
    the chunk rays_o.split(batch_size)[k] of exp_runner.py:349-355 without the full-image ray list."""
    xs_c, ys_c, cam_c = _f32c(xs), _f32c(ys), _f32c(cam).reshape(-1)
    _require_cuda(xs_c, "rays_ndc_grid")
    if cam_c.numel() != 16:
        raise _lib.HonerfError("rays_ndc_grid: one camera (16 floats) expected")
    o = torch.empty(count, 3, device=xs_c.device)
    d = torch.empty_like(o)
    check(lib.hn_rays_ndc_grid(_ptr(xs_c), _ptr(ys_c), xs_c.numel(), ys_c.numel(), _ptr(cam_c), first, count, _ptr(o),
                               _ptr(d), _stream(xs_c)), "hn_rays_ndc_grid")
    return o, d


# ------------------------------------------------------------------------------------------------
# loss epilogues
# ------------------------------------------------------------------------------------------------
_loss_ws = {}


def _loss_workspace(device):
    """Per device AND stream: the kernels leave the workspace clean, so it is zeroed once."""
    key = (str(device), torch.cuda.current_stream(device).cuda_stream)
    if key not in _loss_ws:
        _loss_ws[key] = torch.zeros(int(lib.hn_loss_ws_floats()), device=device)
    return _loss_ws[key]


class _RenderLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, color, wsum, grad_err, true_rgb, true_mask, color_div, color_w, mask_w, igr_w):
        c, w = _f32c(color.detach()).reshape(-1, 3), _f32c(wsum.detach()).reshape(-1)
        t, m = _f32c(true_rgb.detach()).reshape(-1, 3), _f32c(true_mask.detach()).reshape(-1)
        _require_cuda(c, "render_loss")
        n = c.shape[0]
        if w.numel() != n or t.shape[0] != n or m.numel() != n:
            raise _lib.HonerfError("render_loss: color %s, weight_sum %s, true_rgb %s, true_mask %s disagree"
                                   % (tuple(color.shape), tuple(wsum.shape), tuple(true_rgb.shape), tuple(true_mask.shape)))
        ge = _f32c(grad_err.detach()).reshape(1) if grad_err is not None else None
        div_dev = None
        if torch.is_tensor(color_div):          # device scalar: the divisor of the whole batch this shard belongs to
            div_dev = _f32c(color_div.detach()).reshape(1)
            _require_cuda(div_dev, "render_loss")
            color_div = 0.0
        out = torch.empty(8, device=c.device)
        check(lib.hn_render_loss_fwd(_ptr(c), _ptr(w), _ptr(t), _ptr(m), _ptr(ge), n, float(color_div), _ptr(div_dev),
                                     float(color_w), float(mask_w), float(igr_w), _ptr(_loss_workspace(c.device)),
                                     _ptr(out), _stream(c)), "hn_render_loss_fwd")
        ctx.save_for_backward(c, w, t, m, out)
        ctx.coef = (float(color_w), float(mask_w), float(igr_w))
        ctx.shapes = (color.shape, wsum.shape, grad_err.shape if grad_err is not None else None)
        stats = out[1:4]
        ctx.mark_non_differentiable(stats)
        return out[0], stats

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g, _g_stats):
        c, w, t, m, out = ctx.saved_tensors
        n = c.shape[0]
        g = _f32c(g).reshape(1)
        d_c = torch.empty(n, 3, device=c.device)
        d_w = torch.empty(n, device=c.device)
        d_ge = torch.empty(1, device=c.device) if ctx.shapes[2] is not None else None
        check(lib.hn_render_loss_bwd(_ptr(g), _ptr(c), _ptr(w), _ptr(t), _ptr(m), _ptr(out), n, ctx.coef[0], ctx.coef[1],
                                     ctx.coef[2], _ptr(d_c), _ptr(d_w), _ptr(d_ge), _stream(c)), "hn_render_loss_bwd")
        return (d_c.reshape(ctx.shapes[0]), d_w.reshape(ctx.shapes[1]),
                d_ge.reshape(ctx.shapes[2]) if d_ge is not None else None, None, None, None, None, None, None)


def render_loss(color, weight_sum, true_rgb, true_mask, gradient_error=None, color_div=0.0, color_weight=1.0,
                mask_weight=1.0, igr_weight=1.0):
    """Fused render loss: returns (total, stats) with stats = [color_loss, mask_loss, psnr] (not differentiable).
    color_div <= 0: the training normaliser mask_sum + 1e-5 (exp_runner.py:207,221); > 0: an explicit divisor
    (fitting_single.py:254: n rays; fitting_video.py:288: F * P); a device scalar tensor: that value (a ray shard
    passes its batch's mask_sum + 1e-5).  true_mask must already be 0/1."""
    return _RenderLossFn.apply(color, weight_sum, gradient_error, true_rgb, true_mask, color_div, color_weight,
                               mask_weight, igr_weight)


class _InteractionLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sdf_h, sdf_o, thr, w_c, w_p):
        h, o = _f32c(sdf_h.detach()), _f32c(sdf_o.detach())
        _require_cuda(h, "interaction_loss")
        h2, o2 = h.reshape(h.shape[0], -1), o.reshape(o.shape[0], -1)
        n = h2.shape[0]
        if o2.shape[0] != n:
            raise _lib.HonerfError("interaction_loss: %d hand samples vs %d object samples" % (n, o2.shape[0]))
        out = torch.empty(8, device=h.device)
        check(lib.hn_interaction_loss_fwd(_ptr(h2), h2.shape[1], _ptr(o2), o2.shape[1], n, float(thr), float(w_c),
                                          float(w_p), _ptr(_loss_workspace(h.device)), _ptr(out), _stream(h)),
              "hn_interaction_loss_fwd")
        ctx.save_for_backward(h2, o2, out)
        ctx.args = (float(thr), float(w_c), float(w_p))
        ctx.shapes = (sdf_h.shape, sdf_o.shape)
        stats = out[1:5]
        ctx.mark_non_differentiable(stats)
        return out[0], stats

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g, _g_stats):
        h2, o2, out = ctx.saved_tensors
        n = h2.shape[0]
        g = _f32c(g).reshape(1)
        d_h = torch.empty(n, device=h2.device)
        d_o = torch.empty(n, device=h2.device)
        check(lib.hn_interaction_loss_bwd(_ptr(g), _ptr(h2), h2.shape[1], _ptr(o2), o2.shape[1], _ptr(out), n,
                                          ctx.args[0], ctx.args[1], ctx.args[2], _ptr(d_h), _ptr(d_o), _stream(h2)),
              "hn_interaction_loss_bwd")

        def widen(d, like):
            if like.shape[1] == 1:
                return d.reshape(-1, 1)
            full = torch.zeros_like(like)
            full[:, 0] = d
            return full
        return widen(d_h, h2).reshape(ctx.shapes[0]), widen(d_o, o2).reshape(ctx.shapes[1]), None, None, None


def interaction_loss(sdf_hand, sdf_obj, contact_thr=1e-2, w_contact=30.0, w_penet=20.0):
    """Contact + penetration terms of fitting_single.py:268-283 on column 0 of render_out['sdf_hand'] / ['sdf_obj']:
    returns (w_contact * contact + w_penet * penet, stats = [contact, penet, contact_num, penet_num])."""
    return _InteractionLossFn.apply(sdf_hand, sdf_obj, contact_thr, w_contact, w_penet)


# ------------------------------------------------------------------------------------------------
# temporal contact-stability loss
# ------------------------------------------------------------------------------------------------
def nn_select(pts, in_mask, out_mask, return_nearest=False):
    """pts [P,3]; in_mask, out_mask [T,P] bool -> flag [T,P] bool: flag[t,q] = q is the nearest out-candidate of some
    in-point of frame t (np.unique of cKDTree.query(k=1) as a flag array, utils/renderer_batch.py:354-357)."""
    p = _f32c(pts.detach())
    _require_cuda(p, "nn_select")
    im, om = in_mask.to(torch.uint8).contiguous(), out_mask.to(torch.uint8).contiguous()
    T, P = im.shape
    if p.shape != (P, 3) or om.shape != (T, P):
        raise _lib.HonerfError("nn_select: pts %s, in_mask %s, out_mask %s disagree"
                               % (tuple(pts.shape), tuple(in_mask.shape), tuple(out_mask.shape)))
    flag = torch.zeros(T, P, dtype=torch.uint8, device=p.device)
    nearest = torch.empty(T, P, dtype=torch.int64, device=p.device) if return_nearest else None
    check(lib.hn_nn_select(_ptr(p), _ptr(im), _ptr(om), T, P, _ptr(flag), _ptr(nearest), _stream(p)), "hn_nn_select")
    return (flag.bool(), nearest) if return_nearest else flag.bool()


def stable_loss_from_sdf(hand_sdf, pts0, fixed=False):
    """utils/renderer_batch.py:328-369 given the hand SDF of every frame at the (strided) object vertices:
    hand_sdf [F,P] (differentiable), pts0 [P,3] = frame 0's vertices in the object frame.  Device-resident: the frame
    filter, the in/out sets, the nearest-neighbour selection (hn_nn_select) and the sums never visit the host.

    fixed=False reproduces upstream bit of behaviour: `np.setdiff1d(vert_id_all, cur_in_id)` receives the BOOLEAN
    mask, so the "out" set is every vertex except the ids {0, 1} that occur as mask VALUES (0 if some vertex is
    outside, 1 if some is inside); fixed=True uses the complement of the in-set, which is what the code reads like."""
    F_, P = hand_sdf.shape
    neg = hand_sdf.detach() < 0                                  # [F,P] in-sets
    valid = neg.any(dim=1)                                       # frames that penetrate at all
    in_time = valid.sum()
    if fixed:
        out_mask = ~neg
    else:
        out_mask = torch.ones_like(neg)
        if P > 0:
            out_mask[:, 0] = ~(~neg).any(dim=1)                  # id 0 leaves the set when a False is present
        if P > 1:
            out_mask[:, 1] = ~neg.any(dim=1)                     # id 1 leaves the set when a True is present
    flag = nn_select(pts0, neg & valid[:, None], out_mask)
    vf = valid.to(hand_sdf.dtype)
    s_pos = (hand_sdf.clip(0, 1e7) * vf[:, None]).sum(0)         # [P] sums over the penetrating frames
    s_neg = (hand_sdf.clip(-1e7, 0).abs() * vf[:, None]).sum(0)
    n_in = neg.sum(dim=1).to(hand_sdf.dtype)                     # [F]
    denom = ((in_time - 1).to(hand_sdf.dtype) * n_in).clamp_min(1.0)
    in_err = (neg.to(hand_sdf.dtype) @ s_pos) / denom
    out_err = (flag.to(hand_sdf.dtype) @ s_neg) / denom
    total = ((in_err + 0.05 * out_err) * vf).sum() / in_time.clamp_min(1).to(hand_sdf.dtype)
    return torch.where(in_time > 1, total, torch.zeros_like(total))


# ------------------------------------------------------------------------------------------------
# marching cubes on the device
# ------------------------------------------------------------------------------------------------
_mc_devices = set()


def _mc_tables(device):
    key = str(device)
    if key in _mc_devices:
        return
    import numpy as np
    from . import mcubes_tables as T
    n_tris, tris, owner = T.build_tables()
    a = np.asarray(n_tris, dtype=np.int8)
    b = np.ascontiguousarray(np.asarray(tris, dtype=np.int8)[:, :15])
    c = np.ascontiguousarray(np.asarray(owner, dtype=np.int8))
    with torch.cuda.device(device):
        check(lib.hn_mc_set_tables(a.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p),
                                   c.ctypes.data_as(ctypes.c_void_p)), "hn_mc_set_tables")
    _mc_devices.add(key)


def outside_points(rays_o, rays_d, z_vals, sample_dist):
    """Inverted-sphere query points of render_core_outside: -> (pts4 [B*n,4] = (p/r, 1/r), dirs [B*n,3], dists [B,n]).
    Rays are data here (no gradient), like everywhere in the sampling stage."""
    o, d, z = _f32c(rays_o.detach()), _f32c(rays_d.detach()), _f32c(z_vals.detach())
    _require_cuda(z, "outside_points")
    B, n = z.shape
    pts4 = torch.empty(B * n, 4, device=z.device)
    dirs = torch.empty(B * n, 3, device=z.device)
    dists = torch.empty(B, n, device=z.device)
    check(lib.hn_outside_points(_ptr(o), _ptr(d), _ptr(z), float(sample_dist), B, n, _ptr(pts4), _ptr(dirs), _ptr(dists),
                                _stream(z)), "hn_outside_points")
    return pts4, dirs, dists


class _OutsideCompositeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, density, raw_rgb, dists, background):
        dn, rw, ds = _f32c(density.detach()), _f32c(raw_rgb.detach()), _f32c(dists.detach())
        _require_cuda(dn, "outside_composite")
        B, n = ds.shape
        bg = _f32c(background.detach().reshape(3)) if background is not None else None
        sampled = torch.empty(B, n, 3, device=dn.device)
        alpha = torch.empty(B, n, device=dn.device)
        weights = torch.empty(B, n, device=dn.device)
        color = torch.empty(B, 3, device=dn.device)
        check(lib.hn_outside_composite_fwd(_ptr(dn), _ptr(rw), _ptr(ds), _ptr(bg), B, n, _ptr(sampled), _ptr(alpha), _ptr(weights),
                                           _ptr(color), _stream(dn)), "hn_outside_composite_fwd")
        ctx.set_materialize_grads(False)
        ctx.saved = (dn, ds, bg, sampled, alpha, weights)
        ctx.shapes = (density.shape, raw_rgb.shape)
        return color, sampled, alpha, weights

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_color, g_sampled, g_alpha, g_weights):
        dn, ds, bg, sampled, alpha, weights = ctx.saved
        B, n = ds.shape
        d_density = torch.empty(B, n, device=dn.device)
        d_raw = torch.empty(B, n, 3, device=dn.device)
        f = lambda t: _f32c(t) if t is not None else None
        g_color, g_sampled, g_alpha, g_weights = f(g_color), f(g_sampled), f(g_alpha), f(g_weights)
        check(lib.hn_outside_composite_bwd(_ptr(dn), _ptr(ds), _ptr(bg), _ptr(sampled), _ptr(alpha), _ptr(weights), B, n,
                                           _ptr(g_color), _ptr(g_sampled), _ptr(g_alpha), _ptr(g_weights), _ptr(d_density),
                                           _ptr(d_raw), _stream(dn)), "hn_outside_composite_bwd")
        return d_density.reshape(ctx.shapes[0]), d_raw.reshape(ctx.shapes[1]), None, None


def outside_composite(density, raw_rgb, dists, background=None):
    """alpha / transmittance / colour compositing of a density field along rays (render_core_outside): density [B*n,1] or
    [B,n], raw_rgb [B*n,3] or [B,n,3] (the NeRF's raw outputs), dists [B,n] -> (color [B,3], sampled_color [B,n,3],
    alpha [B,n], weights [B,n]); differentiable in density and raw_rgb through all four outputs."""
    return _OutsideCompositeFn.apply(density, raw_rgb, dists, background)


def marching_cubes(u, threshold=0.0):
    """mcubes.marching_cubes(u, threshold) (utils/renderer.py:279) on the device: u [nx,ny,nz] fp32 CUDA tensor ->
    (vertices [V,3] fp32 in index coordinates, triangles [T,3] int32); vertices are shared between cells, triangle normals
    point towards lower values (the reference reverses the triangles afterwards).  The lattice never leaves the device; the
    only host syncs are the two totals that size the outputs."""
    u = _f32c(u.detach())
    _require_cuda(u, "marching_cubes")
    if u.dim() != 3 or min(u.shape) < 2:
        raise ValueError("marching_cubes: u must be [nx, ny, nz] with every side >= 2")
    nx, ny, nz = u.shape
    _mc_tables(u.device)
    n = nx * ny * nz
    flags = torch.empty(3 * n, device=u.device, dtype=torch.int32)
    cell = torch.empty((nx - 1) * (ny - 1) * (nz - 1), device=u.device, dtype=torch.int32)
    check(lib.hn_mc_classify(_ptr(u), nx, ny, nz, float(threshold), _ptr(flags), _ptr(cell), _stream(u)), "hn_mc_classify")
    vscan = torch.cumsum(flags, 0, dtype=torch.int32)
    tscan = torch.cumsum(cell, 0, dtype=torch.int32)
    nv, nt = int(vscan[-1]), int(tscan[-1])
    vertices = torch.empty(nv, 3, device=u.device, dtype=torch.float32)
    triangles = torch.empty(nt, 3, device=u.device, dtype=torch.int32)
    if nv > 0:
        check(lib.hn_mc_emit(_ptr(u), nx, ny, nz, float(threshold), _ptr(flags), _ptr(vscan), _ptr(cell), _ptr(tscan),
                             _ptr(vertices), _ptr(triangles), _stream(u)), "hn_mc_emit")
    return vertices, triangles
