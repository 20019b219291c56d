"""CPU test: the C-ABI library loads without a GPU and exports every symbol include/sptag_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "sptag_b200.h")
LIB = os.path.join(ROOT, "sptag_b200", "lib", "libsptag_b200.so")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sptag_b200_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for must in ("sptag_b200_create", "sptag_b200_load", "sptag_b200_search", "sptag_b200_search_device",
                 "sptag_b200_set_param", "sptag_b200_destroy", "sptag_b200_merge_topk"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    if not os.path.exists(LIB):
        import __graft_entry__
        __graft_entry__.build()
    L = ctypes.CDLL(LIB)
    for s in declared_symbols():
        assert hasattr(L, s), "libsptag_b200.so does not export %s" % s


def test_python_binding_lists_the_same_symbols():
    from sptag_b200 import capi
    assert sorted(capi.EXPORTS) == declared_symbols()


def test_null_handle_is_rejected_without_a_gpu():
    from sptag_b200 import capi
    L = capi.lib()
    assert L.sptag_b200_search(None, None, 1, 1, None, None, None) == 0x15  # EmptyIndex
    assert b"null handle" in L.sptag_b200_last_error()
    assert L.sptag_b200_num_vectors(None) == 0


def test_no_cpu_fallback_without_a_gpu():
    """The product path must fail loudly where there is no B200: creating an index on a box without a GPU returns an
    error code through the C ABI (and the binding raises) -- it never computes anything on the host."""
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip("this check is for GPU-less boxes")
    from sptag_b200 import B200Index, capi
    x = np.zeros((10, 4), np.float32)
    with pytest.raises(capi.SptagB200Error) as e:
        B200Index.create(algo=capi.ALGO_BKT, value_type=capi.VT_FLOAT, metric=capi.METRIC_L2, vectors=x,
                         graph=np.full((10, 4), -1, np.int32), tree_starts=np.array([0], np.int32),
                         tree_nodes=np.array([[10, 1, 2], [0, -1, -1], [-1, -1, -1]], np.int32))
    assert e.value.code != 0


def test_load_of_a_missing_folder_is_an_error():
    from sptag_b200 import B200Index, capi
    with pytest.raises(capi.SptagB200Error) as e:
        B200Index.load("/nonexistent/index/folder")
    assert e.value.code == 0x02  # FailedOpenFile (DefinitionList.h:54-68)
    assert b"indexloader.ini" in capi.lib().sptag_b200_last_error()


def test_missing_library_is_reported_not_papered_over(monkeypatch):
    from sptag_b200 import capi
    monkeypatch.setattr(capi, "_lib", None)
    monkeypatch.setattr(capi, "LIB_PATH", "/nonexistent/libsptag_b200.so")
    with pytest.raises(capi.SptagB200Error) as e:
        capi.lib()
    assert "no CPU fallback" in str(e.value)


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: include/sptag_b200.h must compile as C99 (what cgo / JNI / ctypes generators consume)
    and as C++14 (what the reference would include)."""
    import shutil
    import subprocess
    src = tmp_path / "hdr.c"
    src.write_text('#include "sptag_b200.h"\nint main(void) { sptag_b200_index_desc d; (void)d; return 0; }\n')
    inc = os.path.join(ROOT, "include")
    if shutil.which("gcc"):
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", inc,
                               str(src)])
    if shutil.which("g++"):
        subprocess.check_call(["g++", "-std=c++14", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-x", "c++", "-I", inc,
                               str(src)])


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under sptag_b200/ or include/ may import, link or execute it, and the
    shipped library must not depend on the oracle's shared objects."""
    import subprocess
    pkg = os.path.join(ROOT, "sptag_b200")
    offenders = []
    for base, _, names in os.walk(pkg):
        for n in names:
            if n.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                text = open(os.path.join(base, n), errors="replace").read()
                for needle in ("libsptag_oracle", "libsptag_ref", "sptag_oracle.h", "import reflib", "oracle/_ref",
                               "oracle/_build"):
                    if needle in text:
                        offenders.append((os.path.relpath(os.path.join(base, n), ROOT), needle))
    assert not offenders, offenders
    if os.path.exists(LIB):
        out = subprocess.run(["ldd", LIB], capture_output=True, text=True).stdout
        assert "sptag_oracle" not in out and "sptag_ref" not in out, out
