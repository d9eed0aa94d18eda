"""The C-ABI library loads and exports every symbol include/b200mpc.h declares (no compute calls: no GPU here)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "b200mpc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200mpc_\w+)\s*\(", src)))


def test_header_declares_the_expected_entry_points():
    names = _declared_functions()
    for must in ("b200mpc_create", "b200mpc_destroy", "b200mpc_solve_batch", "b200mpc_solve_batch_device", "b200mpc_solve_batch_multi",
                 "b200mpc_eval_batch", "b200mpc_last_error", "b200mpc_default_options",
                 "b200mpc_obstacles_batch", "b200mpc_obstacles_batch_device", "b200mpc_goals_batch",
                 "b200mpc_goals_batch_device", "b200mpc_reftraj_batch", "b200mpc_reftraj_batch_device",
                 "b200mpc_control_step_device", "b200mpc_dilate_batch", "b200mpc_dilate_batch_device",
                 "b200mpc_inflate_batch", "b200mpc_inflate_batch_device", "b200mpc_local_costmap_batch",
                 "b200mpc_local_costmap_batch_device", "b200mpc_raycast_batch", "b200mpc_raycast_batch_device",
                 "b200mpc_headings_batch", "b200mpc_headings_batch_device"):
        assert must in names


def test_library_exports_every_declared_symbol(built):
    from ros2_mpc_b200 import _shim
    lib = C.CDLL(_shim.LIB_PATH)
    for name in _declared_functions():
        assert hasattr(lib, name), f"{name} declared in include/b200mpc.h but not exported"
    assert lib.b200mpc_abi_version() == 1


def test_params_struct_layout_matches(built):
    from ros2_mpc_b200 import _shim
    lib = _shim.lib()
    assert lib.b200mpc_sizeof_params() == C.sizeof(_shim.Params)
    p = _shim.default_params()
    assert (p.tol, p.max_iter, p.acceptable_tol, p.acceptable_iter, p.mu_init, p.max_soc) == (1e-8, 3000, 1e-6, 15, 0.1, 4)


def test_sass_is_sm100a_fp64(built):
    """The shipped cubin targets sm_100a and the solve kernel is FP64 (DFMA) code."""
    import shutil
    import subprocess
    from ros2_mpc_b200 import _shim
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    out = subprocess.run(["cuobjdump", "-lelf", _shim.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_create_fails_loudly_without_a_gpu(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from ros2_mpc_b200 import MpcPointStabilizationLocal
    with pytest.raises(RuntimeError, match="no CUDA device|CPU fallback"):
        MpcPointStabilizationLocal()


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under ros2_mpc_b200/ may import, link or mention loading it."""
    pkg = os.path.join(ROOT, "ros2_mpc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if not f.endswith((".py", ".cu", ".h", ".cuh", ".cpp")):
                continue
            text = open(os.path.join(dirpath, f)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
            assert "libmpc_oracle" not in text and "mpc_oracle.h" not in text, f
