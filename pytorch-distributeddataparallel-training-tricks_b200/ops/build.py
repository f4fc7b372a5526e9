"""In-tree build of the sm_100a extension (``_ext/_ddpt_C*.so``).

``nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo`` for every ``.cu`` (no torch headers in
device code, so they compile in seconds and cross-compile without a GPU), ``g++`` for the binding
translation units, one shared object linked against libtorch.  Objects are cached by source hash
so incremental rebuilds only touch what changed.  Run as ``python -m ddpt_b200.ops.build``.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys
import sysconfig
from typing import List

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.abspath(os.path.join(_HERE, os.pardir, "csrc"))
EXT_DIR = os.path.abspath(os.path.join(_HERE, os.pardir, "_ext"))
OBJ_DIR = os.path.join(EXT_DIR, "obj")
MODULE = "_ddpt_C"

NVCC_ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "--use_fast_math", "-Xcompiler", "-fPIC",
              "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"]


def _sources():
    cu, cpp = [], []
    for root, _dirs, files in os.walk(CSRC):
        for f in sorted(files):
            path = os.path.join(root, f)
            if f.endswith(".cu"):
                cu.append(path)
            elif f.endswith(".cpp"):
                cpp.append(path)
    return sorted(cu), sorted(cpp)


def _headers_digest() -> str:
    h = hashlib.sha1()
    for root, _dirs, files in os.walk(CSRC):
        for f in sorted(files):
            if f.endswith((".h", ".cuh", ".hpp")):
                with open(os.path.join(root, f), "rb") as fh:
                    h.update(fh.read())
    return h.hexdigest()


def _obj_path(src: str, digest: str) -> str:
    rel = os.path.relpath(src, CSRC).replace(os.sep, "__")
    return os.path.join(OBJ_DIR, f"{rel}.{digest[:12]}.o")


def _digest(src: str, flags: List[str], hdr: str) -> str:
    h = hashlib.sha1()
    with open(src, "rb") as f:
        h.update(f.read())
    h.update(" ".join(flags).encode())
    h.update(hdr.encode())
    return h.hexdigest()


def _run(cmd: List[str], verbose: bool) -> None:
    if verbose:
        print(" ".join(cmd), flush=True)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"build failed: {' '.join(cmd)}\n{res.stdout}\n{res.stderr}")
    if verbose and res.stderr.strip():
        print(res.stderr, file=sys.stderr)


def feature_defines(cu: List[str], cpp: List[str]) -> List[str]:
    names = {os.path.basename(p) for p in cu + cpp}
    defs = []
    if "bindings_nn.cpp" in names:
        defs.append("-DDDPT_HAVE_NN")
    if "bindings_comm.cpp" in names:
        defs.append("-DDDPT_HAVE_COMM")
    if "bindings_loader.cpp" in names:
        defs.append("-DDDPT_HAVE_LOADER")
    if "reducer.cpp" in names:
        defs.append("-DDDPT_HAVE_REDUCER")
    return defs


def build_all(verbose: bool = True, jobs: int = 0) -> str:
    import torch
    from torch.utils import cpp_extension as ce
    os.makedirs(OBJ_DIR, exist_ok=True)
    cu, cpp = _sources()
    hdr = _headers_digest()
    defs = feature_defines(cu, cpp)
    cuda_home = ce.CUDA_HOME or "/usr/local/cuda"
    nvcc = os.path.join(cuda_home, "bin", "nvcc")
    includes = [f"-I{p}" for p in ce.include_paths("cuda")] + [f"-I{sysconfig.get_paths()['include']}",
                                                              f"-I{CSRC}", f"-I{os.path.join(CSRC, 'include')}"]
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    cxx_flags = ["-O2", "-std=c++17", "-fPIC", "-fvisibility=hidden", f"-D_GLIBCXX_USE_CXX11_ABI={abi}",
                 f"-DTORCH_EXTENSION_NAME={MODULE}", "-DTORCH_API_INCLUDE_EXTENSION_H", "-Wno-attributes"] + defs
    nv_flags = NVCC_ARCH + NVCC_FLAGS + [f"-I{CSRC}", f"-I{os.path.join(CSRC, 'include')}",
                                         f"-I{os.path.join(cuda_home, 'include')}"]
    cxx = os.environ.get("CXX", "g++")

    tasks, objs = [], []
    for src in cu:
        obj = _obj_path(src, _digest(src, nv_flags, hdr))
        objs.append(obj)
        if not os.path.exists(obj):
            tasks.append([nvcc, *nv_flags, "-c", src, "-o", obj])
    for src in cpp:
        obj = _obj_path(src, _digest(src, cxx_flags, hdr + torch.__version__))
        objs.append(obj)
        if not os.path.exists(obj):
            tasks.append([cxx, *cxx_flags, *includes, "-c", src, "-o", obj])
    jobs = jobs or min(8, os.cpu_count() or 1)
    with cf.ThreadPoolExecutor(max_workers=jobs) as pool:
        for fut in [pool.submit(_run, t, verbose) for t in tasks]:
            fut.result()

    suffix = sysconfig.get_config_var("EXT_SUFFIX") or ".so"
    target = os.path.join(EXT_DIR, MODULE + suffix)
    stamp = os.path.join(EXT_DIR, "link.stamp")
    want = hashlib.sha1("\n".join(objs).encode()).hexdigest()
    have = open(stamp).read().strip() if os.path.exists(stamp) else ""
    if tasks or have != want or not os.path.exists(target):
        lib_dirs = ce.library_paths("cuda")
        link = [cxx, "-shared", *objs, "-o", target] + [f"-L{d}" for d in lib_dirs] + \
               [f"-Wl,-rpath,{d}" for d in lib_dirs] + \
               ["-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-lcudart",
                "-lpthread"]
        _run(link, verbose)
        with open(stamp, "w") as f:
            f.write(want)
    # drop stale objects
    keep = set(objs)
    for f in os.listdir(OBJ_DIR):
        p = os.path.join(OBJ_DIR, f)
        if p not in keep:
            os.remove(p)
    return target


if __name__ == "__main__":
    print(build_all(verbose="-q" not in sys.argv))
