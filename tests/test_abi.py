"""The C-ABI library loads, exports every symbol include/focalsv_cuda.h declares, and its host-only helpers
work.  No compute calls here (no GPU in this container)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from focalsv_b200 import _abi, api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "focalsv_cuda.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(fsv_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = api.load_library()
    syms = _header_symbols()
    assert len(syms) >= 17
    for s in syms:
        assert hasattr(lib, s), "libfocalsv_cuda.so does not export %s" % s
    assert sorted(api.EXPORTS) == syms
    assert lib.fsv_abi_version() == _abi.ABI_VERSION


def test_struct_layouts_match_header():
    assert _abi.TASK_DTYPE.itemsize == 40 and _abi.RESULT_DTYPE.itemsize == 64
    assert C.sizeof(_abi.Scoring) == 40
    assert _abi.RESULT_DTYPE.fields["cigar_off"][1] == 48 and _abi.RESULT_DTYPE.fields["cells"][1] == 56


def test_header_compiles_as_c_and_layouts_equal_the_python_mirrors(tmp_path):
    """include/focalsv_cuda.h is plain C: gcc compiles a probe against it and the sizes / offsets it prints are the ones
    the ctypes structures and numpy dtypes of focalsv_b200/_abi.py use."""
    import shutil, subprocess
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    structs = {"fsv_task": _abi.TASK_DTYPE, "fsv_result": _abi.RESULT_DTYPE, "fsv_signature": _abi.SIGNATURE_DTYPE,
               "fsv_pair": _abi.PAIR_DTYPE, "fsv_record": _abi.RECORD_DTYPE}
    cstructs = {"fsv_scoring": _abi.Scoring, "fsv_stats": _abi.Stats, "fsv_preset": _abi.PresetC}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "focalsv_cuda.h"', 'int main(void) {']
    for name, dt in structs.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (name, name))
        for f in dt.names:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (name, f, name, f))
    for name, st in cstructs.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (name, name))
        for f, _ in st._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (name, f, name, f))
    lines += ['printf("abi %d\\n", FSV_ABI_VERSION);', 'return 0; }']
    src = tmp_path / "probe.c"
    src.write_text("\n".join(lines))
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", inc, str(src), "-o", str(tmp_path / "probe")])
    got = dict(l.split() for l in subprocess.check_output([str(tmp_path / "probe")]).decode().splitlines())
    for name, dt in structs.items():
        assert int(got[name]) == dt.itemsize, name
        for f in dt.names:
            assert int(got["%s.%s" % (name, f)]) == dt.fields[f][1], (name, f)
    for name, st in cstructs.items():
        assert int(got[name]) == C.sizeof(st), name
        for f, _ in st._fields_:
            assert int(got["%s.%s" % (name, f)]) == getattr(st, f).offset, (name, f)
    assert int(got["abi"]) == _abi.ABI_VERSION


def test_no_cpu_fallback_without_device():
    lib = api.load_library()
    if lib.fsv_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    h = C.c_void_p()
    assert lib.fsv_init(0, C.byref(h)) == _abi.ERR_NO_DEVICE
    with pytest.raises(api.FsvError):
        api.Aligner(0)
    assert b"no CPU path" in lib.fsv_strerror(_abi.ERR_NO_DEVICE)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "focalsv_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "ksw2_oracle" not in src, f


def test_task_cells_helper_matches_oracle(oracle):
    for ql, tl, w in ((100, 100, 10), (1, 50, -1), (777, 333, 50), (2000, 2100, 501), (0, 5, 3)):
        assert api.task_cells(ql, tl, w) == oracle.task_cells(ql, tl, w)


def test_lpt_bins_balance_and_determinism():
    rng = np.random.default_rng(1)
    lens = rng.integers(1000, 200000, 500)
    tasks = api.make_tasks(lens, lens, 3001, 200)
    b1 = api.lpt_bins(tasks, 8)
    b2 = api.lpt_bins(tasks, 8)
    assert np.array_equal(b1, b2) and b1.min() == 0 and b1.max() == 7
    est = (2 * lens - 1) * np.minimum(lens, 3002)
    load = np.array([est[b1 == k].sum() for k in range(8)], dtype=np.float64)
    assert load.max() / load.mean() < 1.05
