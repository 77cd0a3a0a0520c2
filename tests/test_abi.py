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


def _header_constants():
    src = open(os.path.join(ROOT, "include", "casmvs.h")).read()
    consts = {k: int(v) for k, v in re.findall(r"#define\s+(CASMVS_[A-Z0-9_]+)\s+(\d+)", src)}
    for body in re.findall(r"enum\s+\w+\s*\{(.*?)\}", src, flags=re.S):
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        consts.update({k: int(v) for k, v in re.findall(r"(CASMVS_[A-Z0-9_]+)\s*=\s*(\d+)", body)})
    return consts


def test_binding_constants_match_header():
    """The ctypes binding re-states the header's enums and flags: keep them in lock step."""
    from casmvsnet_pl_b200 import _lib
    c = _header_constants()
    assert (c["CASMVS_NCHW"], c["CASMVS_NHWC"]) == (_lib.NCHW, _lib.NHWC)
    assert (c["CASMVS_FP32"], c["CASMVS_TF32"]) == (_lib.FP32, _lib.TF32)
    assert "CASMVS_TF32X3" not in c and "tf32x3" not in _lib.PRECISIONS   # removed, not lying
    assert (c["CASMVS_CONV"], c["CASMVS_CONV_TRANSPOSE"], c["CASMVS_CONV_PLANAR"]) == \
        (_lib.CONV, _lib.CONV_TRANSPOSE, _lib.CONV_PLANAR)
    assert c["CASMVS_ROUND_TF32"] == _lib.ROUND_TF32
    assert c["CASMVS_KEEP_FP32_OUT"] == _lib.KEEP_FP32_OUT
    # flags must not collide with the enum values they are OR-ed into
    assert c["CASMVS_ROUND_TF32"] > max(c["CASMVS_NCHW"], c["CASMVS_NHWC"])
    assert c["CASMVS_KEEP_FP32_OUT"] > c["CASMVS_TF32"]


def test_new_entry_points_validate_arguments():
    """FeatureNet entry points: argument validation comes before any device work."""
    from casmvsnet_pl_b200 import _lib
    lib = _lib.load()
    one = ctypes.c_void_p(16)
    rc = lib.casmvs_conv3d_fwd(one, one, None, None, 1.0, None, one, 1, 8, 8, 4, 16, 16,
                               _lib.CONV_PLANAR, 2, _lib.TF32, None)
    assert rc < 0 and b"stride" in lib.casmvs_last_error()
    rc = lib.casmvs_conv3d_fwd(one, one, None, None, 1.0, None, one, 1, 8, 8, 4, 16, 16,
                               _lib.CONV, 1, _lib.TF32 | 512, None)
    assert rc < 0 and b"precision" in lib.casmvs_last_error()
    rc = lib.casmvs_conv3d_fwd(one, one, None, None, 1.0, None, one, 1, 8, 8, 4, 16, 16,
                               _lib.CONV, 1, 2, None)            # the removed TF32X3 value
    assert rc < 0 and b"precision" in lib.casmvs_last_error()
    assert lib.casmvs_release_weight_images(one, 64) == 0         # nothing cached: no-op
    assert lib.casmvs_fallback_count() == 0 and lib.casmvs_weight_cache_generation() == 0
    rc = lib.casmvs_fpn_merge_fwd(one, one, one, one, one, 1, 15, 16, 8, 0, None)
    assert rc < 0 and b"even" in lib.casmvs_last_error()
    rc = lib.casmvs_fpn_merge_fwd(None, one, one, one, one, 1, 16, 16, 6, 0, None)
    assert rc < 0 and b"multiple of 4" in lib.casmvs_last_error()
    rc = lib.casmvs_conv2d_5x5s2_fwd(one, one, one, 0.01, one, 1, 8, 8, 32, 32, 0, None)
    assert rc < 0 and b"8->16" in lib.casmvs_last_error()
    rc = lib.casmvs_conv2d_rgb8_fwd(None, one, one, 0.01, one, 1, 32, 32, 0, None)
    assert rc < 0 and b"null" in lib.casmvs_last_error()
    assert lib.casmvs_conv2d_rgb8_fwd(one, one, one, 0.01, one, 0, 32, 32, 0, None) == 0   # empty batch


def test_work_item_magic_division_is_exact_in_its_range():
    """The persistent conv kernels decompose a work-item index with q = umulhi(n, m),
    m = floor(2^32 / d) + 1 (common.cuh: make_fastdiv / fastdivmod) and the host only launches when
    items * max_divisor < 2^32 (fastdiv_ok).  Restated here in integer arithmetic: the quotient is
    exact for every n with n * d < 2^32, at the divisors and item counts of the BASELINE shapes
    and at the edge of the range."""
    import random
    rnd = random.Random(0)

    def fastdiv(n, d):
        if d <= 1:
            return n
        m = ((1 << 32) // d + 1) & 0xFFFFFFFF
        return (n * m) >> 32

    divisors = [1, 2, 3, 5, 7, 11, 16, 22, 30, 39, 64, 128, 216, 264, 1000, 4097, 65535]
    for d in divisors:
        top = ((1 << 32) - 1) // d                       # largest n with n * d < 2^32
        samples = {0, 1, d - 1, d, d + 1, top, top - 1, top // 2}
        samples |= {rnd.randrange(0, top + 1) for _ in range(2000)}
        samples |= {k * d + r for k in (top // d, top // d - 1, 12345 % (top // d + 1)) for r in (0, d - 1)
                    if 0 <= k * d + r <= top}
        for n in samples:
            if n < 0:
                continue
            assert fastdiv(n, d) == n // d, (n, d)
    # cfg5 (1920x1056, D = 64): tile columns x rows x chunks of the Cout <= 8 kernel stay in range
    items = 64 * 264 * 64
    assert items * 264 < (1 << 32)
