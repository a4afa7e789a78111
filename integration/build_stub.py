#!/usr/bin/env python
"""Compile integration/pybind_stub.cpp into integration/_build/ob_pybind_stub.so (g++, torch headers, linked against
omniserve_b200/lib/libomniserve_b200.so with an $ORIGIN-relative rpath).  Called by __graft_entry__.build()."""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(HERE, "_build")
SO = os.path.join(OUT, "ob_pybind_stub.so")


def build(force=False):
    import torch
    from torch.utils.cpp_extension import include_paths
    src = os.path.join(HERE, "pybind_stub.cpp")
    lib = os.path.join(ROOT, "omniserve_b200", "lib", "libomniserve_b200.so")
    if not force and os.path.exists(SO) and os.path.getmtime(SO) >= max(os.path.getmtime(src), os.path.getmtime(lib)):
        return SO
    os.makedirs(OUT, exist_ok=True)
    abi = 1 if torch._C._GLIBCXX_USE_CXX11_ABI else 0
    tl = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", src, "-o", SO, f"-D_GLIBCXX_USE_CXX11_ABI={abi}",
           "-DTORCH_EXTENSION_NAME=ob_pybind_stub", "-DTORCH_API_INCLUDE_EXTENSION_H", f"-I{os.path.join(ROOT, 'include')}",
           f"-I{sysconfig.get_paths()['include']}"] + [f"-I{p}" for p in include_paths("cuda")] + [
           f"-L{tl}", f"-Wl,-rpath,{tl}", f"-L{os.path.dirname(lib)}", "-Wl,-rpath,$ORIGIN/../../omniserve_b200/lib",
           "-lomniserve_b200", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-L/usr/local/cuda/lib64",
           "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("pybind stub failed to build:\n" + r.stderr[-3000:])
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
