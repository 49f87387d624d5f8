"""Stage the UNMODIFIED reference under baseline/_ref/ (git-ignored, NOT gpurun-ignored: it travels to the GPU box,
where /root/reference does not exist).  Nothing here is product code and nothing under baseline/_ref/ is committed.

What is staged (all from /root/reference, byte for byte):
  drivers/cutlass/test.py                 flash_attention_cutlass/test.py          the kernel's own driver (test.py:43-87)
  drivers/py/{main.py, tiny_flash_attn.py, tiny_flash_attn_triton.py, main_torch_only.py}
                                          flash_attention_py/*                     main.py:62-102 and what it imports
  attention_cutlass_ref.<abi>.so          the reference's CuTe/sm80 kernel (flash_attention_cutlass/csrc/
                                          {attention_api.cpp, flash_attention.cu, flash_api.cpp}) compiled from the
                                          sources where they lie, unmodified, for compute_100 / sm_100 with its own
                                          flags (build.py:55-73) and its vendored CUTLASS 3.4 headers -- the same-box
                                          GPU comparator of SURVEY.md section 6.  Module name attention_cutlass_ref so
                                          that it can be imported next to this repo's `attention_cutlass`.
The reference's own setup.py is not run (it targets sm_80/sm_90 only and installs into site-packages).

Usage:  python baseline/build_ref.py [--force] [--no-kernel]
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("TFA_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

DRIVERS = {
    "drivers/cutlass/test.py": "flash_attention_cutlass/test.py",
    "drivers/py/main.py": "flash_attention_py/main.py",
    "drivers/py/tiny_flash_attn.py": "flash_attention_py/tiny_flash_attn.py",
    "drivers/py/tiny_flash_attn_triton.py": "flash_attention_py/tiny_flash_attn_triton.py",
    "drivers/py/main_torch_only.py": "flash_attention_py/main_torch_only.py",
}
EXT = "attention_cutlass_ref" + (sysconfig.get_config_var("EXT_SUFFIX") or ".so")


def stage_drivers():
    staged = {}
    for dst, src in DRIVERS.items():
        s, d = os.path.join(REF, src), os.path.join(OUT, dst)
        if not os.path.exists(s):
            continue
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        staged[dst] = hashlib.sha256(open(s, "rb").read()).hexdigest()
    if staged:
        with open(os.path.join(OUT, "drivers", "SHA256SUMS"), "w") as f:     # lets the GPU-side test prove "unchanged"
            for k_, v_ in sorted(staged.items()):
                f.write(f"{v_}  {k_}\n")
    return staged


def build_ref_kernel(force=False):
    """nvcc (device code, sm_100) + g++ (pybind glue) on the reference's three source files, unmodified."""
    cdir = os.path.join(REF, "flash_attention_cutlass")
    srcs = [os.path.join(cdir, "csrc", f) for f in ("attention_api.cpp", "flash_attention.cu", "flash_api.cpp")]
    target = os.path.join(OUT, EXT)
    if not all(os.path.exists(s) for s in srcs):
        return target if os.path.exists(target) else None          # GPU box: use the prebuilt file
    if os.path.exists(target) and not force:
        return target
    import torch
    from torch.utils import cpp_extension as ce

    os.makedirs(OUT, exist_ok=True)
    bdir = os.path.join(OUT, "_build")
    os.makedirs(bdir, exist_ok=True)
    inc = ["-I" + os.path.join(cdir, d) for d in ("csrc", "include", "deps/cutlass/include",
                                                  "deps/cutlass/tools/utils/include", "deps/cutlass/examples/common")]
    for p in ce.include_paths("cuda"):
        inc += ["-isystem", p]
    inc += ["-isystem", sysconfig.get_paths()["include"]]
    cxx11 = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    defs = ["-DTORCH_EXTENSION_NAME=attention_cutlass_ref", "-DTORCH_API_INCLUDE_EXTENSION_H",
            f"-D_GLIBCXX_USE_CXX11_ABI={cxx11}"]
    # the reference's own nvcc flags (flash_attention_cutlass/build.py:55-73), arch swapped for the B200
    nvflags = ["-O3", "-std=c++17", "-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__",
               "-U__CUDA_NO_HALF2_OPERATORS__", "-U__CUDA_NO_BFLOAT16_CONVERSIONS__", "--expt-relaxed-constexpr",
               "--expt-extended-lambda", "--use_fast_math", "-lineinfo", "--ptxas-options=-O2",
               "-gencode", "arch=compute_100,code=sm_100", "-Xcompiler", "-fPIC", "-w"]
    objs = []
    for s in srcs:
        o = os.path.join(bdir, os.path.basename(s) + ".o")
        objs.append(o)
        if s.endswith(".cu"):
            cmd = [NVCC] + nvflags + defs + inc + ["-c", s, "-o", o]
        else:
            cmd = [os.environ.get("CXX", "g++"), "-O3", "-std=c++17", "-fPIC", "-w"] + defs + inc + ["-c", s, "-o", o]
        subprocess.run(cmd, check=True)
    ld = []
    for d in ce.library_paths("cuda"):
        ld += ["-L", d, "-Wl,-rpath," + d]
    subprocess.run([os.environ.get("CXX", "g++"), "-shared", "-o", target] + objs + ld +
                   ["-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-lcudart"], check=True)
    shutil.rmtree(bdir, ignore_errors=True)
    return target


if __name__ == "__main__":
    if not os.path.isdir(REF):
        print(f"[baseline] {REF} absent: nothing to stage (prebuilt baseline/_ref is used as is)")
        sys.exit(0)
    os.makedirs(OUT, exist_ok=True)
    print("[baseline] drivers:", ", ".join(sorted(stage_drivers())))
    if "--no-kernel" not in sys.argv:
        print("[baseline] reference kernel:", build_ref_kernel(force="--force" in sys.argv))
