"""CPU: the C-ABI library loads, exports exactly what include/vrb200.h declares, and fails loudly
(no CPU fallback) when there is no CUDA device. No compute calls here."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "vrb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vrb_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from volren_b200 import _capi
    lib = _capi.load_library()
    declared = header_symbols()
    assert declared, "no declarations parsed from include/vrb200.h"
    for name in declared:
        assert hasattr(lib, name), f"libvrb200.so does not export {name}"
    assert sorted(_capi.SYMBOLS) == declared
    assert lib.vrb_abi_version() == 1


def test_struct_layout_matches_header(tmp_path):
    """sizeof/offsetof of the PODs as gcc sees include/vrb200.h == the ctypes mirrors."""
    import subprocess
    from volren_b200 import _capi
    src = tmp_path / "layout.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "vrb200.h"\n'
        'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(vrb_params), sizeof(vrb_counters), sizeof(vrb_brick_view),'
        ' offsetof(vrb_params, vol_density_transform), offsetof(vrb_params, use_transferfunc), offsetof(vrb_params, env_strength),'
        ' offsetof(vrb_brick_view, range_mips));return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    P = _capi.Params
    want = [C.sizeof(P), C.sizeof(_capi.Counters), C.sizeof(_capi.BrickView), P.vol_density_transform.offset,
            P.use_transferfunc.offset, P.env_strength.offset, _capi.BrickView.range_mips.offset]
    assert got == want
    # the oracle shares the same PODs by construction
    from oracle import binding
    assert binding.Params is _capi.Params


def test_status_strings():
    from volren_b200 import _capi
    lib = _capi.load_library()
    assert lib.vrb_status_string(0) == b"ok"
    assert b"1024" in lib.vrb_status_string(_capi.VRB_ERR_TOO_MANY_BRICKS)


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from volren_b200 import Context, VrbError, _capi
    with pytest.raises(VrbError) as e:
        Context(0)
    assert e.value.status == _capi.VRB_ERR_NO_DEVICE


def test_missing_library_raises(monkeypatch, tmp_path):
    from volren_b200 import _capi
    monkeypatch.setattr(_capi, "_lib", None)
    with pytest.raises(ImportError):
        _capi.load_library(str(tmp_path / "nope.so"))


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under volren_b200/ may reference it."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "volren_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")) or f == "Makefile":
                s = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"vr_oracle|libvr_oracle|from oracle|import oracle|oracle/", s):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad
