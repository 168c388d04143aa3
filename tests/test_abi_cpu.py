"""CPU tests of the boundary: the C-ABI library builds for sm_100a, loads, and exports every
symbol include/pmc.h declares.  No compute calls (no GPU here)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "pmc.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pmc_[a-z0-9_]+)\s*\(", txt)))


def test_header_declares_the_four_call_sites():
    syms = declared_symbols()
    for s in ("pmc_init_r", "pmc_assign", "pmc_subsweep", "pmc_shift_cells", "pmc_sweep",
              "pmc_create", "pmc_destroy", "pmc_run_host"):
        assert s in syms


def test_library_builds_loads_and_exports_everything(built):
    import pmc_b200
    assert os.path.exists(pmc_b200.LIB_PATH)
    L = ctypes.CDLL(pmc_b200.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(L, s), f"{s} declared in include/pmc.h but not exported"
    assert sorted(pmc_b200.EXPORTS) == declared_symbols()


def test_library_is_sm100a_sass_with_packed_fp32(built):
    import pmc_b200
    out = subprocess.run(["cuobjdump", "-sass", pmc_b200.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    # Blackwell packed-FP32 pair tests in the sub-sweep hot loop
    assert "FFMA2" in out and "FADD2" in out and "FMUL2" in out


def test_error_strings_and_no_gpu_failure(built):
    import pmc_b200
    L = pmc_b200.lib()
    assert L.pmc_error_string(0) == b"success"
    assert b"nmax" in L.pmc_error_string(-2)
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            pmc_b200.ParallelMC(4096)        # product path fails loudly without a GPU


def test_host_schedule_matches_oracle_without_gpu(built):
    """pmc_schedule / pmc_colour_to_off are pure host functions; they need a handle, which
    needs a device, so here only the stateless one is compared."""
    import pmc_b200
    from oracle import oracle as O
    for c in range(4):
        assert pmc_b200.ParallelMC.colour_to_off(c) == O.Oracle.colour_to_off(c)
