"""Compiles nerf_qa_b200/csrc_torch/nqa_torch_ops.cpp into the in-tree extension (called by the Makefile).  Host code only:
g++ against the torch headers of the running interpreter, linked to libnerfqa_b200.so next to it ($ORIGIN rpath)."""
import os
import subprocess
import sys

import torch
from torch.utils import cpp_extension as ce

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(root, sys.argv[1]) if len(sys.argv) > 1 else os.path.join(root, "nerf_qa_b200", "_nqa_torch_ops.so")
cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}",
       "-DTORCH_API_INCLUDE_EXTENSION_H"]
cmd += [f"-I{p}" for p in ce.include_paths()] + [f"-I{cuda_home}/include", f"-I{root}/include"]
cmd += [os.path.join(root, "nerf_qa_b200", "csrc_torch", "nqa_torch_ops.cpp"), "-o", out]
cmd += [f"-L{p}" for p in ce.library_paths()] + ["-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch"]
cmd += [f"-L{root}/nerf_qa_b200", "-lnerfqa_b200", "-Wl,-rpath,$ORIGIN"]
print(" ".join(cmd))
subprocess.check_call(cmd)
