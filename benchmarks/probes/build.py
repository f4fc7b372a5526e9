"""Build the hardware-probe extension (benchmarks/probes/_ddpt_probes*.so), separate from the shipped one.

The probes answer "how does the hardware address this?" questions (tcgen05 descriptor views that are not 1024-byte
aligned, where kind::mxf8f6f4.block_scale reads its scale factors in TMEM); their kernels reuse the product's
umma.cuh / TMA helpers but are not product code (keep them out of the shipped .so)."""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "pytorch-distributeddataparallel-training-tricks_b200", "csrc")
MODULE = "_ddpt_probes"


def build(verbose: bool = False) -> str:
    import torch
    from torch.utils import cpp_extension as ce
    cuda_home = ce.CUDA_HOME or "/usr/local/cuda"
    nvcc = os.path.join(cuda_home, "bin", "nvcc")
    obj_dir = os.path.join(HERE, "obj")
    os.makedirs(obj_dir, exist_ok=True)
    inc = [f"-I{CSRC}", f"-I{os.path.join(CSRC, 'include')}", f"-I{os.path.join(cuda_home, 'include')}"]
    objs = []
    # the probes need the tensor-map helper and the runtime helpers of the product sources
    cu = [os.path.join(HERE, "umma_probe.cu"), os.path.join(HERE, "umma_mx_probe.cu"),
          os.path.join(CSRC, "nn", "tensor_map.cu"), os.path.join(CSRC, "common", "runtime.cu")]
    for src in cu:
        obj = os.path.join(obj_dir, os.path.basename(src) + ".o")
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "--use_fast_math",
               "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", *inc, "-c", src, "-o", obj]
        subprocess.run(cmd, check=True, capture_output=not verbose)
        objs.append(obj)
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    cpp = os.path.join(HERE, "bindings_probes.cpp")
    obj = os.path.join(obj_dir, "bindings_probes.o")
    includes = [f"-I{p}" for p in ce.include_paths("cuda")] + [f"-I{sysconfig.get_paths()['include']}"]
    subprocess.run([os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-fPIC", f"-D_GLIBCXX_USE_CXX11_ABI={abi}",
                    f"-DTORCH_EXTENSION_NAME={MODULE}", "-DTORCH_API_INCLUDE_EXTENSION_H", "-Wno-attributes", *includes,
                    "-c", cpp, "-o", obj], check=True, capture_output=not verbose)
    objs.append(obj)
    target = os.path.join(HERE, MODULE + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))
    lib_dirs = ce.library_paths("cuda")
    subprocess.run([os.environ.get("CXX", "g++"), "-shared", *objs, "-o", target] + [f"-L{d}" for d in lib_dirs] +
                   [f"-Wl,-rpath,{d}" for d in lib_dirs] +
                   ["-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-lcudart", "-lcuda"],
                   check=True, capture_output=not verbose)
    return target


def load():
    if HERE not in sys.path:
        sys.path.insert(0, HERE)
    import importlib
    try:
        return importlib.import_module(MODULE)
    except ImportError:
        build()
        importlib.invalidate_caches()
        return importlib.import_module(MODULE)


if __name__ == "__main__":
    print(build(verbose=True))
