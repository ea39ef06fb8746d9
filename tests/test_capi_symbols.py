"""CPU-side checks of the C-ABI boundary: the library builds, loads and exports every symbol
include/md2_loss.h declares; argument validation works without touching a GPU."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from monodepth2_b200 import build, _capi
    build.build()
    return _capi.load_library()


def header_symbols():
    src = open(os.path.join(ROOT, "include", "md2_loss.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(md2_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported(lib):
    from monodepth2_b200 import _capi
    syms = header_symbols()
    assert len(syms) >= 18
    assert sorted(_capi.SYMBOLS) == syms
    for s in syms:
        assert hasattr(lib, s), "libmd2loss.so does not export %s" % s


def test_struct_layout_matches_header(tmp_path):
    """sizeof / offsetof of every field of the header's structs, as gcc lays them out, against the ctypes mirror."""
    import subprocess
    from monodepth2_b200._capi import Md2Problem, Md2Tensors
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "md2_loss.h"', 'int main(void) {',
             'printf("md2_problem %zu\\n", sizeof(md2_problem));', 'printf("md2_tensors %zu\\n", sizeof(md2_tensors));']
    for cname, cls in (("md2_problem", Md2Problem), ("md2_tensors", Md2Tensors)):
        for fname, _ in cls._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines += ['return 0; }']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    assert int(got["md2_problem"]) == C.sizeof(Md2Problem)
    assert int(got["md2_tensors"]) == C.sizeof(Md2Tensors)
    for cname, cls in (("md2_problem", Md2Problem), ("md2_tensors", Md2Tensors)):
        for fname, _ in cls._fields_:
            assert int(got["%s.%s" % (cname, fname)]) == getattr(cls, fname).offset, (cname, fname)


def test_validation_without_gpu(lib):
    from monodepth2_b200._capi import Md2Problem
    n = C.c_size_t(0)
    good = Md2Problem(batch=12, height=192, width=640, num_scales=4, num_src=2, automask=1,
                      min_depth=0.1, max_depth=100.0, disparity_smoothness=1e-3, want_grad=1)
    assert lib.md2_loss_workspace_bytes(C.byref(good), C.byref(n)) == 0
    assert 30e6 < n.value < 200e6
    bad = Md2Problem(batch=12, height=190, width=640, num_scales=4, num_src=2, min_depth=0.1, max_depth=100.0)
    assert lib.md2_loss_workspace_bytes(C.byref(bad), C.byref(n)) == -1
    four = Md2Problem(batch=12, height=192, width=640, num_scales=4, num_src=4, min_depth=0.1, max_depth=100.0)
    assert lib.md2_loss_workspace_bytes(C.byref(four), C.byref(n)) == 0          # MD2_MAX_SRC sources are supported
    bad2 = Md2Problem(batch=12, height=192, width=640, num_scales=4, num_src=5, min_depth=0.1, max_depth=100.0)
    assert lib.md2_loss_workspace_bytes(C.byref(bad2), C.byref(n)) == -1
    assert lib.md2_status_string(-3) == b"workspace too small"
    assert lib.md2_version() >= 100


def test_missing_library_fails_loudly(tmp_path):
    from monodepth2_b200 import _capi
    with pytest.raises(_capi.Md2Error):
        _capi.load_library(str(tmp_path / "nope.so"))


def test_register_budget():
    """ptxas -v log of the shipped build: no marching-kernel instantiation spills (VERDICT r1 bar), and none of the
    5-CTA instantiations (96 threads, <= 2 sources, no --avg_reprojection with gradients) really uses 129-136 registers -
    such a kernel is given 4 CTAs per SM although 5 x 96 x 136 < 65536 (profiles/r02_optimization_log.md)."""
    import os
    import re
    log = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "monodepth2_b200", "lib",
                       "libmd2loss.md2_kernels.cu.ptxas.log")
    if not os.path.exists(log):
        pytest.skip("no ptxas log (library not built here)")
    txt = open(log).read()
    entries = re.findall(r"Compiling entry function '(\S+)'.*?\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n"
                         r"ptxas info\s+: Used (\d+) registers", txt)
    march = [e for e in entries if "md2_march_roles" in e[0]]
    assert len(march) >= 60, len(march)
    for name, stack, st, ld, regs in march:
        assert int(st) == 0 and int(ld) == 0, (name, st, ld)
        m = re.search(r"CfgILi(\d)ELb(\d)ELb(\d)ELb(\d)ELb(\d)", name)
        nsrc, avg, auto_, grad = int(m.group(1)), int(m.group(2)), int(m.group(3)), int(m.group(4))
        five_ctas = nsrc <= 2 and not (avg and grad) and grad        # RoleCfg::MIN_CTAS == 5 and 3 roles = 96 threads
        if five_ctas:
            assert not (128 < int(regs) <= 136), (name, regs)
