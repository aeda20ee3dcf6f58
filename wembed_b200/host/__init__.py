"""Host-side mirror of the reference's public interface: include/wembed.h implemented over the C ABI, plus the
pybind11 module `wembed` (same names as the reference's python/bindings.cpp).

    from wembed_b200.host import load
    wembed = load()          # builds on demand, returns the `wembed` extension module
"""
from __future__ import annotations

import importlib.util
import os
import subprocess
import sys
import sysconfig

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_LIBDIR = os.path.join(os.path.dirname(_HERE), "lib")
HOST_LIB = os.path.join(_LIBDIR, "libwembed_host.so")
PYMOD = os.path.join(_LIBDIR, "wembed" + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))
_SRCS = [os.path.join(_HERE, f) for f in ("graph.cpp", "embedder.cpp", "hierarchy.cpp", "wembed.cpp")]
_DEPS = _SRCS + [os.path.join(_HERE, f) for f in ("graph.hpp", "embedder.hpp", "hierarchy.hpp", "bindings.cpp")] + [
    os.path.join(_ROOT, "include", "wembed.h"), os.path.join(_ROOT, "include", "wembed_b200.h")]
CXX = "/usr/bin/g++"


def _stale(target):
    return not os.path.exists(target) or any(os.path.getmtime(d) > os.path.getmtime(target) for d in _DEPS)


def build(force: bool = False) -> str:
    """libwembed_host.so (C++ facade over libwembed_b200.so) and the pybind11 module wembed.*.so, both in wembed_b200/lib."""
    from .. import build as cuda_build
    cuda_build.build()
    inc = ["-I", os.path.join(_ROOT, "include"), "-I", _HERE]
    common = ["-std=c++17", "-O2", "-fPIC", "-shared", "-Wall", "-pthread", "-Wl,-rpath,$ORIGIN", "-L", _LIBDIR]
    if force or _stale(HOST_LIB):
        subprocess.run([CXX, *common, *inc, "-o", HOST_LIB, *_SRCS, "-lwembed_b200"], check=True)
    if force or _stale(PYMOD):
        import pybind11
        subprocess.run([CXX, *common, *inc, "-I", pybind11.get_include(), "-I", sysconfig.get_paths()["include"], "-fvisibility=hidden",
                        "-o", PYMOD, os.path.join(_HERE, "bindings.cpp"), "-lwembed_host", "-lwembed_b200"], check=True)
    return PYMOD


def load():
    build()
    if "wembed" in sys.modules:
        return sys.modules["wembed"]
    spec = importlib.util.spec_from_file_location("wembed", PYMOD)
    mod = importlib.util.module_from_spec(spec)
    sys.modules["wembed"] = mod
    spec.loader.exec_module(mod)
    return mod
