"""CPU-side checks of the drop-in boundary: libmcfgpu.so loads, exports every symbol include/mcfgpu.h declares, the
ctypes mirrors have the C struct sizes, and - with no GPU - the engine refuses to run instead of falling back."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import mincostflow_b200 as mcf
from mincostflow_b200 import solver

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mcfgpu.h")


def _declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mcf_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = mcf.load_library()
    names = _declared_functions()
    assert len(names) >= 23
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/mcfgpu.h but not exported by libmcfgpu.so"
    assert lib.mcf_api_version() == 1


def test_struct_sizes_match_the_header(tmp_path):
    """Compile a 10-line C program against include/mcfgpu.h and compare sizeof with the ctypes mirrors."""
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "mcfgpu.h"\nint main(void){printf("%zu %zu %zu\\n", sizeof(mcf_optimization_config),'
                   ' sizeof(mcf_options), sizeof(mcf_metrics)); return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    a, b, c = (int(x) for x in subprocess.check_output([str(exe)]).split())
    assert (a, b, c) == (C.sizeof(solver._CConfig), C.sizeof(solver._COptions), C.sizeof(solver.SolverMetrics))


def test_default_options_are_the_reference_defaults():
    lib = mcf.load_library()
    o = solver._COptions()
    lib.mcf_default_options(C.byref(o))
    assert (o.supply_type, o.pivot_rule, o.auto_configuration, o.optimized_pivot) == (0, 2, 1, 0)   # NetworkSimplex.cs:38, :77, :90
    c = o.config                                                                                    # OptimizationTypes.cs:25-38
    assert (c.flags, c.max_block_size, c.min_block_size, c.consecutive_hits_before_adapt) == (0, 100, 25, 3)
    assert (c.block_size_growth_factor, c.block_size_shrink_factor, c.low_hit_rate_threshold,
            c.high_hit_rate_threshold, c.min_block_size_ratio) == (1.2, 0.8, 0.05, 0.3, 0.125)


def test_invalid_arguments_are_rejected_without_a_device():
    lib = mcf.load_library()
    h = C.c_void_p()
    src = np.array([0, 5], np.int32); tgt = np.array([1, 0], np.int32)
    assert lib.mcf_create(C.c_int32(2), C.c_int32(2), src.ctypes.data_as(C.c_void_p), tgt.ctypes.data_as(C.c_void_p), C.byref(h)) == -1
    assert lib.mcf_create(C.c_int32(-1), C.c_int32(0), None, None, C.byref(h)) == -1
    assert lib.mcf_solve(None, None) == -1 and lib.mcf_get_flows(None, None) == -1


def test_no_cpu_fallback_when_no_gpu():
    if mcf.device_count() > 0:
        pytest.skip("a B200 is visible")
    g = mcf.GraphBuilder().AddNodes(2).AddArc(0, 1).Build()
    with pytest.raises(mcf.EngineError) as ei:
        mcf.NetworkSimplex(g)
    assert ei.value.code == -2                                       # MCF_ERR_NO_DEVICE


def test_product_package_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "mincostflow_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "ns_oracle" not in text and "from oracle" not in text and "import oracle" not in text, f


def test_graph_builder_mirror():
    """GraphBuilder.cs:10-104 error behaviour."""
    b = mcf.GraphBuilder().AddNodes(3)
    b.AddArc(0, 1).AddArc(1, 2)
    with pytest.raises(mcf.ArgumentException):
        b.AddArc(0, 7)
    with pytest.raises(mcf.ArgumentException):
        b.AddNode(1)
    g = b.Build()
    assert g.NodeCount == 3 and g.ArcCount == 2 and int(g.Source(mcf.Arc(1))) == 1 and int(g.Target(mcf.Arc(1))) == 2
