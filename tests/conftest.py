"""pytest configuration: `gpu` marker, path-based import of the product package.

-m "not gpu": oracle vs the reference-source run (tests/golden/ref_*), known answers, golden fixtures, host logic,
              C-ABI load + symbols.
-m gpu      : parity tests proper -- CUDA path through the C ABI vs the CPU oracle.
"""
import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG_DIR = os.path.join(ROOT, "lua-multigrid-poisson_b200")


def load_package():
    """The package directory carries the reference's name (with '-'), so import it by path."""
    name = "lua_multigrid_poisson_b200"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(
        name, os.path.join(PKG_DIR, "__init__.py"), submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def mgp():
    pkg = load_package()
    if not os.path.exists(pkg.LIB_PATH):
        pkg.build()
    return pkg


@pytest.fixture(scope="session")
def orc():
    import oracle
    oracle.build()
    return oracle
