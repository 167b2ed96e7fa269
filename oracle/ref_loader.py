"""Import the UNMODIFIED reference (iscas3dv/HO-NeRF) hot-path modules on CPU.

TEST INFRASTRUCTURE ONLY.  Used by ``oracle/make_golden.py`` (fixture generation, in the
build container only: ``/root/reference`` does not exist on the GPU box) and by
``tests/test_oracle_vs_reference.py`` (skipped when the reference tree is absent).
Nothing in the product package may import this file.

The reference imports ``mcubes``, ``matplotlib.pyplot`` and ``mpl_toolkits.mplot3d`` at module
top level (utils/renderer.py:6,8; utils/renderer_batch.py:6-10); they are absent here and are
not used by the code we exercise, so empty stub modules are placed in ``sys.modules`` first
(SURVEY.md section 8c).
"""
import importlib
import os
import sys
import types
import warnings

REFERENCE_ROOT = os.environ.get("HONERF_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "utils", "renderer.py"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def load_reference():
    """Returns a namespace with .fields, .renderer, .renderer_batch (reference modules)."""
    if not reference_available():
        raise FileNotFoundError("reference tree not found at %s" % REFERENCE_ROOT)
    _stub("mcubes")
    mpl = _stub("matplotlib")
    plt = _stub("matplotlib.pyplot", Axes=object)
    mpl.pyplot = plt
    tk = _stub("mpl_toolkits")
    m3d = _stub("mpl_toolkits.mplot3d", Axes3D=object)
    tk.mplot3d = m3d
    warnings.filterwarnings("ignore", category=FutureWarning)
    warnings.filterwarnings("ignore", category=UserWarning)
    # The reference uses the top-level package name ``utils``; import it under an alias so it
    # cannot collide with anything in this repository.
    if "honerf_ref_utils" not in sys.modules:
        spec_dir = os.path.join(REFERENCE_ROOT, "utils")
        pkg = types.ModuleType("honerf_ref_utils")
        pkg.__path__ = [spec_dir]
        sys.modules["honerf_ref_utils"] = pkg
    ns = types.SimpleNamespace()
    ns.fields = importlib.import_module("honerf_ref_utils.fields")
    ns.renderer = importlib.import_module("honerf_ref_utils.renderer")
    ns.renderer_batch = importlib.import_module("honerf_ref_utils.renderer_batch")
    return ns


# Network hyper-parameters of the reference configs (confs/wmask_realobj_bean.conf:40-77,
# confs/wmask_realhand_hand1.conf:40-77, fit_confs/fit_12_8views.conf:26-91).
OBJ_SDF_CONF = dict(d_out=257, d_in=3, d_hidden=256, n_layers=8, skip_in=[4], v_multires=10,
                    r_multires=4, bias=0.5, scale=1.0, geometric_init=True, weight_norm=True)
OBJ_COLOR_CONF = dict(d_feature=256, d_in=3, d_out=3, d_hidden=256, n_layers=4, weight_norm=True,
                      v_multires=10, r_multires=4, grad_multires=4, squeeze_out=True,
                      use_gradients=True)
HAND_SDF_CONF = dict(d_out=257, d_in=3, d_hidden=256, n_layers=8, skip_in=[4], v_multires=10,
                     r_multires=7, bias=0.5, scale=1.0, geometric_init=True, weight_norm=True)
HAND_COLOR_CONF = dict(d_feature=256, d_in=3, d_out=3, d_hidden=256, n_layers=4,
                       weight_norm=True, v_multires=10, r_multires=7, grad_multires=4,
                       squeeze_out=True, use_gradients=True)
RENDERER_CONF = dict(n_samples=64, n_importance=64, n_outside=0, up_sample_steps=4, perturb=1.0)
VARIANCE_INIT = 0.3
