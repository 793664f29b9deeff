"""The C-ABI shared library: builds, loads, exports exactly what include/rrmpg_b200.h declares, and fails
loudly (no CPU fallback) when no GPU is present.  No compute calls are made here."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from rrmpg_b200 import _lib, engine

HEADER = os.path.join(ROOT, "include", "rrmpg_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"RRB_API\s+[\w\s\*]+?\b(rrb_\w+)\s*\(", src)))


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for need in ["rrb_abc_simulate", "rrb_hbvedu_simulate", "rrb_gr4j_simulate", "rrb_cemaneige_simulate",
                 "rrb_cemaneigegr4j_simulate", "rrb_init", "rrb_shutdown", "rrb_last_error", "rrb_host_alloc",
                 "rrb_host_free"]:
        assert need in syms


def test_library_loads_and_exports_every_declared_symbol():
    handle = _lib.lib()
    for name in declared_symbols():
        assert hasattr(handle, name), f"{name} declared in the header but not exported"
    assert sorted(_lib.exported_symbols()) == declared_symbols(), "ctypes table and header disagree"
    nm = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(re.findall(r" T (rrb_\w+)", nm))
    assert exported == declared_symbols(), "the .so exports symbols the header does not declare (or misses some)"
    assert handle.rrb_version() == 110


def test_header_is_plain_c():
    """The boundary must be consumable from C (cgo / JNI / ctypes style bindings)."""
    src = '#include "rrmpg_b200.h"\nint main(void){ rrb_opts o; o.struct_size = (int)sizeof(o); return o.struct_size == 0; }\n'
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"),
                        "-x", "c", "-"], input=src, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_opts_struct_layout_matches_header():
    fields = ["stream", "block", "variant", "x4_max", "qobs", "mse", "slab_steps", "objective", "n_devices", "devices",
              "obs_stats", "state_in", "state_out", "out_row_pitch"]
    src = ('#include <stdio.h>\n#include <stddef.h>\n#include "rrmpg_b200.h"\nint main(void){printf("%zu", sizeof(rrb_opts));'
           + "".join(f'printf(" %zu", offsetof(rrb_opts, {f}));' for f in fields) + 'printf("\\n");return 0;}\n')
    exe = os.path.join(ROOT, "oracle", "_layout_probe")
    r = subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-x", "c", "-", "-o", exe], input=src,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    try:
        got = [int(v) for v in subprocess.run([exe], capture_output=True, text=True).stdout.split()]
    finally:
        os.remove(exe)
    O = _lib.Opts
    assert got == [C.sizeof(O)] + [getattr(O, f).offset for f in fields]


@pytest.mark.skipif(_lib.device_count() > 0, reason="a GPU is present")
def test_no_gpu_means_loud_failure_not_a_cpu_fallback():
    with pytest.raises(RuntimeError, match="no CUDA device|no CPU fallback"):
        engine.abc(np.ones(10), 0.0, np.array([[0.1, 0.2, 0.3]]))
    with pytest.raises(RuntimeError):
        _lib.require_gpu()


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "rrmpg_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", text, re.M), fn
                assert "liboracle" not in text and "rr_oracle" not in text.replace("oracle/rr_oracle.c header", ""), fn


def test_one_cta_per_sm_launch_shape():
    """rr_kernels.h one_cta_block: the CTA size of a mid-sized single-catchment launch -- ceil(warps / SMs) warps, so that
    warp w lands on sub-partition w % 4 of its own SM -- and 0 where the shape does not apply (host-side logic, compiled
    here with the host compiler)."""
    src = ('#include <stdio.h>\n#include "rr_kernels.h"\nint main(){using rrb::one_cta_block;'
           'printf("%d %d %d %d %d %d %d %d\\n", one_cta_block(65536, 148, 512, 5), one_cta_block(32768, 148, 512, 5),'
           'one_cta_block(37888, 148, 512, 9), one_cta_block(104192, 148, 512, 5), one_cta_block(52096, 148, 512, 5),'
           'one_cta_block(65536, 148, 384, 9), one_cta_block(1786, 4, 512, 5), one_cta_block(65536, 0, 512, 5));return 0;}\n')
    exe = os.path.join(ROOT, "oracle", "_shape_probe")
    r = subprocess.run(["g++", "-std=c++17", "-I", os.path.join(ROOT, "rrmpg_b200", "csrc"), "-I", "/usr/local/cuda/include",
                        "-x", "c++", "-", "-o", exe], input=src, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    try:
        got = [int(x) for x in subprocess.run([exe], capture_output=True, text=True).stdout.split()]
    finally:
        os.remove(exe)
    # 2048 warps / 148 SMs -> 14 warps; 1024 -> 7; 8 warps < min 9; 22 warps > 16; 11; 14 warps need 448 > 384 threads;
    # 56 warps on 4 "SMs" -> 14; sm_count 0 = 148
    assert got == [448, 224, 0, 0, 352, 0, 448, 448]
