"""CPU-only checks of the boundary: the C-ABI library loads, exports every symbol the header
declares, the ctypes prototypes cover the header, and the product refuses CPU tensors."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "honerf_b200.h")).read()
    return sorted(set(re.findall(r"HN_API\s+[\w\s\*]+?\b(hn_\w+)\s*\(", txt)))


def test_library_exports_every_header_symbol():
    from honerf_b200 import _lib
    syms = _header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(_lib.lib, s), "libhonerf_b200.so does not export %s" % s
        assert s in _lib.PROTOTYPES, "ctypes prototype missing for %s" % s
    assert set(_lib.PROTOTYPES) == set(syms)
    assert _lib.lib.hn_version() >= 100
    # the bring-up self-test GEMMs are a separate library, not product symbols
    txt = open(os.path.join(ROOT, "include", "honerf_b200_selftest.h")).read()
    tsyms = set(re.findall(r"HN_API\s+[\w\s\*]+?\b(hn_\w+)\s*\(", txt))
    assert tsyms == set(_lib.SELFTEST_PROTOTYPES) and not (tsyms & set(syms))
    st = _lib.load_selftest()
    for t in tsyms:
        assert hasattr(st, t) and not hasattr(_lib.lib._load(), t)


def test_size_queries_run_without_a_gpu():
    from honerf_b200 import _lib
    # stashes are sized for whole 128-point tiles and for the largest layout of any precision: HN_TC_BF16X3 keeps
    # E | H[8] | D[8] | EB in fp32 (64 + 16 * 256 + 64 floats per point), HN_TC_MIXED16 keeps E | EB (fp32), E16 and the
    # 16-bit tiles EM[8] | EML[8] | A16[8] | D16[8] (128 + 32 + 32 * 128 floats per point)
    assert _lib.lib.hn_sdf_obj_stash_floats(1000) == 1024 * (128 + 32 + 32 * 128)
    assert _lib.lib.hn_sdf_obj_stash_floats(1024) == 1024 * (128 + 32 + 32 * 128)
    assert _lib.lib.hn_sdf_obj_ws_floats(10, _lib.HN_WS_SDF_ONLY) > 0
    assert _lib.lib.hn_sdf_obj_ws_floats(1000, _lib.HN_WS_BWD) >= 1024 * (64 + 24 * 256 + 64)
    # colour stash: fp32 layout ENC | FEAT | R[4] (128 + 5 * 256) or, HN_TC_MIXED16, ENC + hi / lo T16 tiles of six arrays (128 + 12 * 128)
    assert _lib.lib.hn_color_obj_stash_floats(10) == 128 * (128 + 12 * 128)
    assert _lib.lib.hn_sdf_obj_chain_bytes() > 4 * 1024 * 1024 and _lib.lib.hn_color_obj_chain_bytes() > 2 * 1024 * 1024


def test_no_cpu_fallback():
    import honerf_b200 as H
    import ref_conf
    sdf = H.SDFNetwork_OBJ(H.Embedding(), 2, "real", **ref_conf.OBJ_SDF_CONF)
    with pytest.raises(H.HonerfError):
        sdf.sdf(torch.zeros(4, 3))
    with pytest.raises(H.HonerfError):
        H.ops.up_sample(torch.zeros(2, 8), torch.zeros(2, 8), 4, 64.0)


def test_bad_arguments_are_reported_not_crashed():
    from honerf_b200 import _lib
    r = _lib.lib.hn_wn_pack(None, None, 4, 4, 4, 1.0, None, None, 0, None)
    assert r == -1
    assert b"hn_wn_pack" in _lib.lib.hn_last_error()


def test_state_dict_keys_match_reference_layout():
    import honerf_b200 as H
    import ref_conf
    import synth
    sp, cp = synth.obj_states()
    sdf = H.SDFNetwork_OBJ(H.Embedding(), 4, "real", **ref_conf.OBJ_SDF_CONF)
    col = H.RenderingNetwork_OBJ(H.Embedding(), "real", **ref_conf.OBJ_COLOR_CONF)
    assert set(sdf.state_dict()) == set(sp)
    assert set(col.state_dict()) == set(cp)
    sdf.load_state_dict(sp)          # strict
    col.load_state_dict(cp)
    for k, v in sp.items():
        assert sdf.state_dict()[k].shape == v.shape
    with pytest.raises(NotImplementedError):
        H.SDFNetwork_OBJ(H.Embedding(), 4, "real", **dict(ref_conf.OBJ_SDF_CONF, d_hidden=128))


def test_same_seed_same_init_as_reference():
    import ref_loader
    if not ref_loader.reference_available():
        pytest.skip("reference tree not present (GPU box)")
    import honerf_b200 as H
    import ref_conf
    ref = ref_loader.load_reference()
    torch.manual_seed(0)
    mine = H.SDFNetwork_OBJ(H.Embedding(), 3, "real", **ref_conf.OBJ_SDF_CONF).state_dict()
    torch.manual_seed(0)
    theirs = ref.fields.SDFNetwork_OBJ(ref.fields.Embedding(), 3, "real", **ref_conf.OBJ_SDF_CONF).state_dict()
    assert set(mine) == set(theirs)
    for k in mine:
        assert torch.equal(mine[k], theirs[k]), k
