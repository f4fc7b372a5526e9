"""Static checks on the compiled sm_100a kernels (no GPU needed: nvcc cross-compiles here).

Round 2 lesson: an epilogue helper with five 32-float arrays pushed the persistent / halo GEMM kernels over their
register budget; ptxas spilled ~120 bytes per thread and every forward convolution lost 15-50 % (conv.3 forward
77.8 -> 115.7 us) without any test failing.  The tensor-core kernels must compile without local-memory spills."""
import os
import re
import shutil
import subprocess

import pytest

from conftest import ROOT

CSRC = os.path.join(ROOT, "pytorch-distributeddataparallel-training-tricks_b200", "csrc")
GEMM_SOURCES = ["nn/conv_gemm_persistent.cu", "nn/conv_gemm_halo.cu", "nn/wgrad_gemm_multitap.cu", "nn/wgrad_gemm.cu",
                "nn/conv_dgrad_kwmerge.cu", "nn/conv0_wgrad_tc.cu"]


@pytest.mark.parametrize("src", GEMM_SOURCES)
def test_tensor_core_kernels_do_not_spill(src, tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "--use_fast_math",
           "--expt-relaxed-constexpr", f"-I{CSRC}", f"-I{CSRC}/include", "-Xptxas", "-v", "-c",
           os.path.join(CSRC, src), "-o", str(tmp_path / "k.o")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-2000:]
    spills = [int(m) for m in re.findall(r"(\d+) bytes spill stores", res.stderr)]
    regs = [int(m) for m in re.findall(r"Used (\d+) registers", res.stderr)]
    assert spills and max(spills) == 0, f"{src}: ptxas spilled (max {max(spills)} bytes); registers {sorted(set(regs))}"
    assert max(regs) <= 168, f"{src}: {max(regs)} registers per thread leaves no room next to a communication CTA"
