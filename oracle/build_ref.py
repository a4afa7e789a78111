#!/usr/bin/env python
"""Rebuild the reference's own CUDA extensions for sm_100 into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing under ``omniserve_b200/`` may import this.

The reference (mit-han-lab/omniserve, /root/reference, read-only) ships 13
pybind11 torch extensions (``kernels/setup.py:156-333``).  This recipe compiles
the ones on the W4A8KV4 hot path *from the sources where they lie* -- no source
is copied into this repository -- with the reference's own compiler flags
(``kernels/setup.py:19-35``: ``-O2 --use_fast_math ...``) and the arch flag its
``setup.py`` would auto-detect on a B200 (``compute_100/sm_100``,
``kernels/setup.py:96-105,141-145``).  We do not run the reference's build
system (setuptools would try to write into the read-only tree); each module is
``nvcc -c`` / ``g++ -c`` per source + one link, driven from here.

Output: ``oracle/_ref/omniserve_backend/<module>.so`` (git-ignored; travels to
the GPU box with gpurun).  The modules are the *Ampere kernels rebuilt on the
same box* that BASELINE.md section 2 asks to be timed next to ours, and the
second parity oracle of SURVEY.md section 8(c).

Usage: python oracle/build_ref.py [--jobs N] [--only mod1,mod2]
Skips silently (exit 0) when /root/reference is absent (GPU box).
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys
import sysconfig
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("OMNISERVE_REFERENCE", "/root/reference")
CSRC = os.path.join(REF, "kernels", "csrc")
OUT = os.path.join(HERE, "_ref")
PKG = os.path.join(OUT, "omniserve_backend")
OBJ = os.path.join(OUT, "obj")

FA = "fused_attention"
# module -> sources, exactly as kernels/setup.py lists them
MODULES = {
    "qgemm_w4a8_per_chn": ["qgemm/w4a8_per_chn/pybind.cpp", "qgemm/w4a8_per_chn/gemm_cuda.cu"],
    "qgemm_w4a8_per_group": ["qgemm/w4a8_per_group/pybind.cpp", "qgemm/w4a8_per_group/gemm_cuda.cu"],
    "qgemm_w8a8": ["qgemm/w8a8/pybind.cpp", "qgemm/w8a8/w8a8_gemm_cuda.cu"],
    "fused_kernels": ["fused.cpp", "fused_kernels.cu"],
    "layernorm_ops": ["layernorm.cpp", "layernorm_kernels.cu"],
    "activation_ops": ["activation.cpp", "activation_kernels.cu"],
    "fused_attention_pure_dense": [
        f"{FA}/fused_attention_pure_dense/fused_attention.cpp",
        f"{FA}/fused_attention_pure_dense/decoderMaskedMultiheadAttention.cu",
        f"{FA}/fused_attention_pure_dense/update_kv_cache.cu",
        f"{FA}/fused_attention_pure_dense/input_metadata_helper.cu",
    ],
    "fused_attention_fine_grained_dense": [
        f"{FA}/fused_attention_fine_grained/dense_attention/fused_attention.cpp",
        f"{FA}/fused_attention_fine_grained/dense_attention/decoderMaskedMultiheadAttention.cu",
        f"{FA}/fused_attention_fine_grained/fine_grained_common/update_kv_cache.cu",
        f"{FA}/common/input_metadata_helper.cu",
    ],
    "fused_attention_fine_grained_sparse": [
        f"{FA}/fused_attention_fine_grained/sparse_attention/fused_attention.cpp",
        f"{FA}/fused_attention_fine_grained/sparse_attention/decoderMaskedMultiheadAttention.cu",
        f"{FA}/fused_attention_fine_grained/fine_grained_common/update_kv_cache.cu",
        f"{FA}/common/input_metadata_helper.cu",
    ],
    "fused_attention_per_tensor_dense": [
        f"{FA}/fused_attention_per_tensor/dense_attention/fused_attention.cpp",
        f"{FA}/fused_attention_per_tensor/dense_attention/decoderMaskedMultiheadAttention.cu",
        f"{FA}/fused_attention_per_tensor/per_tensor_common/update_kv_cache.cu",
        f"{FA}/common/input_metadata_helper.cu",
    ],
    "fused_attention_selector": [
        f"{FA}/sparse_utils/KVPageSelector/fused_kv_page_selector.cpp",
        f"{FA}/sparse_utils/KVPageSelector/KVPageSelector.cu",
    ],
    "fused_attention_ctx_pool": [
        f"{FA}/sparse_utils/ContextPool/pybind.cpp",
        f"{FA}/sparse_utils/ContextPool/context_pool_kernel.cu",
    ],
}


def _flags():
    import torch
    from torch.utils.cpp_extension import include_paths

    abi = 1 if torch._C._GLIBCXX_USE_CXX11_ABI else 0
    inc = [f"-I{p}" for p in include_paths("cuda")] + [f"-I{sysconfig.get_paths()['include']}"]
    common = ["-DENABLE_BF16", f"-D_GLIBCXX_USE_CXX11_ABI={abi}", "-DTORCH_API_INCLUDE_EXTENSION_H"]
    cxx = ["-g0", "-O3", "-fopenmp", "-std=c++17", "-fPIC"] + common
    nvcc = [
        "-O2", "-std=c++17", "-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__",
        "-U__CUDA_NO_BFLOAT16_OPERATORS__", "-U__CUDA_NO_BFLOAT16_CONVERSIONS__",
        "-U__CUDA_NO_BFLOAT162_OPERATORS__", "-U__CUDA_NO_BFLOAT162_CONVERSIONS__",
        "--expt-relaxed-constexpr", "--expt-extended-lambda", "--use_fast_math",
        "-gencode", "arch=compute_100,code=sm_100", "--threads", "4",
        "--compiler-options", "-fPIC", "-w",
    ] + common
    torch_lib = os.path.join(os.path.dirname(torch.__file__), "lib")
    link = [
        "-shared", f"-L{torch_lib}", f"-Wl,-rpath,{torch_lib}", "-L/usr/local/cuda/lib64",
        "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python",
        "-lcudart", "-lgomp",
    ]
    return inc, cxx, nvcc, link


def _compile(args):
    mod, src, inc, cxx, nvcc = args
    path = os.path.join(CSRC, src)
    tag = hashlib.sha1((mod + src).encode()).hexdigest()[:10]
    obj = os.path.join(OBJ, f"{mod}.{tag}.o")
    if os.path.exists(obj) and os.path.getmtime(obj) > os.path.getmtime(path):
        return obj, 0.0, ""
    t0 = time.time()
    ext = [f"-DTORCH_EXTENSION_NAME={mod}"]
    if src.endswith(".cu"):
        cmd = ["nvcc", "-c", path, "-o", obj] + inc + nvcc + ext
    else:
        cmd = ["g++", "-c", path, "-o", obj] + inc + cxx + ext
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        return None, time.time() - t0, f"{' '.join(cmd)}\n{r.stderr[-4000:]}"
    return obj, time.time() - t0, ""


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--jobs", type=int, default=max(1, (os.cpu_count() or 2) // 2))
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    if not os.path.isdir(CSRC):
        print(f"[build_ref] {CSRC} not present; nothing to do (prebuilt oracle/_ref is used if shipped)")
        return 0
    mods = {k: v for k, v in MODULES.items() if not a.only or k in a.only.split(",")}
    os.makedirs(PKG, exist_ok=True)
    os.makedirs(OBJ, exist_ok=True)
    open(os.path.join(PKG, "__init__.py"), "a").close()
    inc, cxx, nvcc, link = _flags()
    jobs = [(m, s, inc, cxx, nvcc) for m, srcs in mods.items() for s in srcs]
    objs: dict[str, list[str]] = {m: [] for m in mods}
    failed = set()
    with cf.ThreadPoolExecutor(a.jobs) as ex:
        for (m, s, *_), (obj, dt, err) in zip(jobs, ex.map(_compile, jobs)):
            if obj is None:
                failed.add(m)
                print(f"[build_ref] FAILED {m}:{s} ({dt:.0f}s)\n{err}", file=sys.stderr)
            else:
                objs[m].append(obj)
                print(f"[build_ref] {m}:{s} ok ({dt:.0f}s)")
    for m in mods:
        if m in failed:
            continue
        so = os.path.join(PKG, f"{m}.so")
        r = subprocess.run(["g++", "-o", so] + objs[m] + link, capture_output=True, text=True)
        if r.returncode != 0:
            failed.add(m)
            print(f"[build_ref] LINK FAILED {m}\n{r.stderr[-3000:]}", file=sys.stderr)
        else:
            print(f"[build_ref] linked {so}")
    # launch interposer for the reference selector's missing dynamic shared memory (see oracle/ref_launch_shim.c)
    shim_src = os.path.join(HERE, "ref_launch_shim.c")
    shim = os.path.join(OUT, "libref_launch_shim.so")
    if os.path.exists(shim_src) and (not os.path.exists(shim) or os.path.getmtime(shim) < os.path.getmtime(shim_src)):
        r = subprocess.run(["gcc", "-shared", "-fPIC", "-O2", "-o", shim, shim_src, "-ldl"], capture_output=True, text=True)
        print(f"[build_ref] {'built' if r.returncode == 0 else 'FAILED'} {shim} {r.stderr[-500:]}")
    print(f"[build_ref] done: {len(mods) - len(failed)}/{len(mods)} modules; failed={sorted(failed)}")
    return 1 if failed else 0


if __name__ == "__main__":
    sys.exit(main())
