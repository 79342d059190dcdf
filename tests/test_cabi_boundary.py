"""The drop-in boundary without a GPU: the C-ABI library loads, exports every symbol that
include/mgpoisson.h declares, and refuses to run (loudly) when there is no CUDA device."""
import ctypes as C
import os
import re
import subprocess

import pytest

from conftest import ROOT, _have_gpu


def test_library_exports_every_declared_symbol(mgp):
    names = mgp.header_functions()
    assert len(names) >= 40
    out = subprocess.run(["nm", "-D", "--defined-only", mgp.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (mg_\w+)", out))
    missing = [n for n in names if n not in exported]
    assert not missing, f"declared in include/mgpoisson.h but not exported: {missing}"
    # and nothing undeclared leaks out of the ABI
    assert exported <= set(names), exported - set(names)
    L = mgp.lib()
    for n in names:
        assert hasattr(L, n)


def test_cdef_block_is_plain_c(mgp):
    cdef = mgp.header_cdef()
    assert "#" not in re.sub(r"/\*.*?\*/", "", cdef, flags=re.S)  # no preprocessor inside the cdef block
    # the LuaJIT wrapper embeds the same declarations
    lua = open(os.path.join(ROOT, "lua-multigrid-poisson_b200", "lua", "multigrid-poisson", "cuda.lua")).read()
    for n in mgp.header_functions():
        assert n in lua, f"{n} missing from cuda.lua's ffi.cdef"


def test_header_compiles_as_c99(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "mgpoisson.h"\nint main(void){return MG_OK;}\n')
    subprocess.run(["/usr/bin/gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I",
                    os.path.join(ROOT, "include"), "-c", str(src), "-o", str(tmp_path / "t.o")], check=True)


def test_version_and_no_compute_without_gpu(mgp):
    L = mgp.lib()
    assert b"sm_100a" in L.mg_version()
    if _have_gpu():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = L.mg_create(2, 64, 0, 7, -1, C.byref(h))
    assert rc != 0 and not h.value
    assert b"no CPU fallback" in L.mg_last_error(None)
    with pytest.raises(mgp.MGError):
        mgp.MultigridCUDA(64)


def test_create_rejects_bad_arguments(mgp):
    L = mgp.lib()
    h = C.c_void_p()
    assert L.mg_create(2, 48, 0, 7, -1, C.byref(h)) == -1   # not a power of two
    assert L.mg_create(4, 64, 0, 7, -1, C.byref(h)) == -1   # dim
    assert L.mg_create(2, 64, 9, 7, -1, C.byref(h)) == -1   # real_kind
    assert L.mg_destroy(None) == 0


def test_product_never_touches_the_oracle():
    # the product path must not import, link or call anything under oracle/
    pkg = os.path.join(ROOT, "lua-multigrid-poisson_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".lua", "Makefile")):
                txt = open(os.path.join(dp, fn), errors="ignore").read()
                for pat in ("import oracle", "from oracle", "libmgoracle", "orc_", "oracle/"):
                    assert pat not in txt, (pat, os.path.join(dp, fn))
