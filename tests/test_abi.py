"""The C-ABI shared library: loads without a GPU, exports every symbol include/lmpc_b200.h declares,
fails loudly (no CPU fallback) when there is no device."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "lmpc_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(lmpc_[a-z_0-9]+)\s*\(", hdr)))


def test_header_symbols_are_exported(pkg):
    from racing_lmpc_ros2_b200 import binding
    lib = binding.load_library()
    syms = _declared_symbols()
    assert len(syms) >= 16
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/lmpc_b200.h but not exported"
    assert set(binding.EXPORTS) == set(syms)
    assert lib.lmpc_version() >= 100


def test_pod_layout_matches_header(pkg):
    """sizeof of the ctypes mirrors == the header's structs (compiled with gcc here)."""
    import subprocess, tempfile
    from racing_lmpc_ros2_b200 import binding
    src = '#include <stdio.h>\n#include "lmpc_b200.h"\nint main(){printf("%zu %zu %zu %zu\\n", sizeof(lmpc_vehicle_params), sizeof(lmpc_mpc_config), sizeof(lmpc_batch_in), sizeof(lmpc_batch_out));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")])
        sizes = [int(v) for v in subprocess.check_output([os.path.join(d, "t")]).split()]
    assert sizes == [C.sizeof(binding.VehicleParams), C.sizeof(binding.MpcConfig), C.sizeof(binding.BatchIn), C.sizeof(binding.BatchOut)]


def test_no_cpu_fallback(pkg):
    """Without a CUDA device the product refuses to construct; it never routes to the oracle."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from racing_lmpc_ros2_b200.solver import BatchedRacingMPC, LmpcError
    with pytest.raises(LmpcError, match="no CUDA device"):
        BatchedRacingMPC(pkg.configs.BARC_VEHICLE, pkg.configs.barc_lmpc_config(20), max_batch=4)
    # and the product package does not import the oracle
    import sys
    pkgdir = os.path.join(ROOT, "racing-lmpc-ros2_b200")
    for root, _, files in os.walk(pkgdir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                txt = open(os.path.join(root, f)).read()
                assert "liblmpc_oracle" not in txt and "import oracle" not in txt and "from oracle" not in txt, f
