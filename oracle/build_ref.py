"""Compile the reference's own CPU attention (flash_attention_c) UNMODIFIED into oracle/_ref/.

Sources are compiled where they lie under /root/reference (never copied into this repo):
    /root/reference/flash_attention_c/csrc/attn.cpp   naive_attn / flash_attn (OpenMP)
    /root/reference/flash_attention_c/csrc/ops.cu     pybind glue (PYBIND11_MODULE(_kernels))
The reference's own CMake build is not run; this is the short recipe SURVEY.md section 8c verified:
torch.utils.cpp_extension.load with -O3 -fopenmp.  Output: oracle/_ref/_kernels.so (git-ignored,
NOT gpurun-ignored, so it travels to the GPU box where /root/reference does not exist).
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("TFA_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")


def build_ref(force: bool = False, verbose: bool = False):
    src_dir = os.path.join(REF, "flash_attention_c", "csrc")
    srcs = [os.path.join(src_dir, "attn.cpp"), os.path.join(src_dir, "ops.cu")]
    target = os.path.join(OUT, "_kernels.so")
    if not all(os.path.exists(s) for s in srcs):
        return target if os.path.exists(target) else None      # GPU box: use the prebuilt file
    if os.path.exists(target) and not force:
        return target
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    os.environ.setdefault("MAX_JOBS", "4")
    from torch.utils.cpp_extension import load

    build_dir = os.path.join(OUT, "_build")
    os.makedirs(build_dir, exist_ok=True)
    load(name="_kernels", sources=srcs, extra_cflags=["-O3", "-fopenmp"],
         extra_ldflags=["-L/usr/lib/gcc/x86_64-linux-gnu/13", "-lgomp"], with_cuda=True,
         build_directory=build_dir, verbose=verbose, is_python_module=False)
    built = os.path.join(build_dir, "_kernels.so")
    shutil.copy2(built, target)
    return target


if __name__ == "__main__":
    print(build_ref(force="--force" in sys.argv, verbose=True))
