"""The C-ABI library loads and exports every symbol include/casmvs.h declares
(no compute calls: there is no GPU in the CPU test tier)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "casmvs.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(casmvs_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_expected_surface():
    syms = _declared_symbols()
    for must in ("casmvs_warp_cost_fwd", "casmvs_conv3d_fwd", "casmvs_costreg_fwd",
                 "casmvs_regress_fwd", "casmvs_depth_hypotheses_fwd", "casmvs_homo_warp_fwd",
                 "casmvs_device_check", "casmvs_last_error"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from casmvsnet_pl_b200 import _lib
    assert os.path.isfile(_lib.LIB_PATH), "build the extension first (__graft_entry__.build())"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in _declared_symbols():
        assert hasattr(lib, s), f"libcasmvs.so lacks {s}"


def test_binding_covers_header():
    from casmvsnet_pl_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared_symbols()
    lib = _lib.load()
    assert lib.casmvs_version() == 100


def test_host_side_queries_without_gpu():
    from casmvsnet_pl_b200 import _lib, ops
    lib = _lib.load()
    assert lib.casmvs_packed_conv3d_weight_floats(32, 8) == 27 * 32 * 8
    assert lib.casmvs_warp_cost_workspace_bytes(_lib.NHWC, 1, 3, 32, 128, 160) == 0
    assert lib.casmvs_warp_cost_workspace_bytes(_lib.NCHW, 1, 3, 32, 128, 160) == 3 * 32 * 128 * 160 * 4
    total = lib.casmvs_costreg_param_floats(32)
    last = ops.costreg_layer_info(32, 10)
    assert last["cin"] == 8 and last["cout"] == 1 and last["shift_off"] + 1 == total
    first = ops.costreg_layer_info(32, 0)
    assert first["w_off"] == 0 and first["cin"] == 32 and first["cout"] == 8
    up = ops.costreg_layer_info(32, 7)
    assert up["kind"] == _lib.CONV_TRANSPOSE and (up["cin"], up["cout"]) == (64, 32)


def test_errors_are_reported_not_thrown():
    """bad arguments -> negative status + message (no GPU needed: validation is first)."""
    from casmvsnet_pl_b200 import _lib
    lib = _lib.load()
    rc = lib.casmvs_warp_cost_fwd(None, 1, None, None, None, 1, 1, 3, 8, 8, 16, 16, 1, None, 0, None)
    assert rc < 0 and b"null" in lib.casmvs_last_error()
    one = ctypes.c_void_p(16)
    rc = lib.casmvs_warp_cost_fwd(one, 1, one, one, one, 1, 1, 3, 12, 8, 16, 16, 1, None, 0, None)
    assert rc < 0 and b"multiple of 8" in lib.casmvs_last_error()
    rc = lib.casmvs_costreg_fwd(one, one, one, 1, 8, 12, 16, 16, 0, one, 1 << 40, None)
    assert rc < 0 and b"divisible by 8" in lib.casmvs_last_error()
    with pytest.raises(_lib.CasMVSError):
        _lib.check(rc, "costreg")
