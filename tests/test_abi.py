"""CPU: the C-ABI library loads and exports exactly what include/ahv_b200.h declares."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ahv_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"AHV_API\s+[\w\s\*]+?\b(ahv_\w+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for s in ("ahv_score", "ahv_topk", "ahv_topk_merge", "ahv_so3_from_normals", "ahv_rotate_volume",
              "ahv_forward_3d2d", "ahv_workspace_bytes", "ahv_predict_host", "ahv_version", "ahv_status_string"):
        assert s in syms


def test_library_exports_every_declared_symbol(ahv):
    lib = ahv._lib.lib()
    for s in declared_symbols():
        assert hasattr(lib, s), s
    assert set(ahv._lib.SIGNATURES) == set(declared_symbols())
    assert lib.ahv_version() == 200


def test_status_strings(ahv):
    lib = ahv._lib.lib()
    assert lib.ahv_status_string(0) == b"ok"
    assert b"no fallback" in lib.ahv_status_string(ahv._lib.AHV_ENOTSUP)
    with pytest.raises(RuntimeError):
        ahv._lib.check(ahv._lib.AHV_EINVAL, "x")


def test_argument_validation_without_gpu(ahv):
    lib = ahv._lib.lib()
    # invalid arguments are rejected before any device work
    assert lib.ahv_topk(None, 1, 10, 0, 0, None, None, None, 0, None) == ahv._lib.AHV_EINVAL
    assert lib.ahv_topk(None, 1, 10, 33, 0, None, None, None, 0, None) == ahv._lib.AHV_EINVAL
    assert lib.ahv_score(None, 7, None, None, 0, None, None, None, None, None, None, None, 1, 0, 1, 1, 0, None, 0, None) == ahv._lib.AHV_EINVAL
    assert lib.ahv_workspace_bytes(32, 50000, 1) >= 32 * 50000 * 4


def test_library_has_sm100a_code_only():
    so = os.path.join(ROOT, "3dahv_b200", "lib3dahv_b200.so")
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", so], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_ops_refuse_cpu_tensors(ahv):
    import torch

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ahv.ops.forward_3d2d(torch.zeros(1, 16, 8, 8, 8), torch.zeros(32, 384), torch.zeros(32, 32), torch.zeros(32))
    with pytest.raises(RuntimeError):
        ahv.so3.random_rotations(4, device="cpu")


def test_argument_validation_of_round2_entries_without_gpu(ahv):
    """Every entry added in round 2 rejects malformed arguments before it touches the device."""
    import ctypes

    lib, E = ahv._lib.lib(), ahv._lib.AHV_EINVAL
    assert lib.ahv_peer_bytes(0, 1) == 0 and lib.ahv_peer_bytes(4, 33) == 0
    assert lib.ahv_peer_bytes(4, 8) == 512 + 2 * 8 * 4 * 8 * 64                     # header + flags + entries[2][8][pairs][k]
    p = ctypes.c_void_p()
    assert lib.ahv_peer_alloc(0, 1, ctypes.byref(p)) == E and lib.ahv_peer_alloc(4, 0, ctypes.byref(p)) == E
    assert lib.ahv_so3_perturb(None, 4, 0, 5.0, 0, None, None) == E                  # m >= 1
    assert lib.ahv_so3_perturb(None, 4, 8, -1.0, 0, None, None) == E                 # angle in [0, 180]
    assert lib.ahv_resblock3d(None, None, None, None, None, 3, None) == E
    assert lib.ahv_infonce(None, None, 0, None, 15.0, 0.0, None, None, 2, 10, None) == E    # temperature > 0
    assert lib.ahv_refine_workspace_bytes(2, 1000, 0, 8) == 0 and lib.ahv_refine_workspace_bytes(2, 1000, 33, 8) == 0
    assert lib.ahv_refine_workspace_bytes(2, 1000, 8, 8) >= lib.ahv_workspace_bytes(2, 1000, 8) + lib.ahv_workspace_bytes(2, 64, 1)
    args = [None] * 9
    assert lib.ahv_refine(None, 0, None, None, 0, None, None, None, None, 8, 4, 5.0, 0, None, None, None, None, None, None, None,
                          2, 4, 0, None, 0, None) == E                               # k > N
    assert lib.ahv_topk_exchange(None, None, None, 0, 0, 10, 2, 1, None, None, None, 0, 2, None, 4, 1, None) == E
    assert lib.ahv_verify_sharded(None, 0, None, None, 0, None, None, None, None, None, None, None, 0, 0, 2, 10, 0, None, 0, 0, 2,
                                  None, 4, 1, None) == E                             # k >= 1
    s = ctypes.c_void_p()
    assert lib.ahv_predict_host_ex(None, None, 0, None, None, 0, None, None, None, None, None, None, None, None, 1, 0, 1, 1, 0,
                                   0, 1, None, 0, 0, None) == E                      # no session
    assert lib.ahv_host_session_destroy(None) == 0
    # saved-activation training entries: null buffers, unknown arithmetic mode, misaligned activation buffer
    assert lib.ahv_score_train(None, None, None, 0, None, None, None, None, None, None, None, 2, 10, None, 0, None) == E
    one = ctypes.c_void_p(16)
    bwd = lambda h1, mode: lib.ahv_score_backward_saved(one, one, one, 0, one, one, one, one, one, h1, one, one, one, one, one, one,
                                                        2, 10, mode, None)
    assert bwd(one, 7) == E and bwd(ctypes.c_void_p(24), ahv.MATH_TC) == E and bwd(None, ahv.MATH_TC) == E
    assert lib.ahv_version() == 200


def _compile_c_example(tmp_path):
    exe = str(tmp_path / "predict_host")
    lib_dir = os.path.join(ROOT, "3dahv_b200")
    cmd = ["gcc", "-O2", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "predict_host.c"),
           "-o", exe, "-L" + lib_dir, "-l:lib3dahv_b200.so", "-Wl,-rpath," + lib_dir]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    return exe


def test_plain_c_host_compiles_against_the_header(ahv, tmp_path):
    """include/ahv_b200.h is a C header (no C++, no torch types): a plain-C host builds and links against the library."""
    exe = _compile_c_example(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 2 and "usage" in out.stderr
