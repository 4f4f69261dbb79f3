"""ORACLE (test infrastructure): build the REFERENCE's own CUDA ops into oracle/_ref/.

    python oracle/build_ref.py [--force]

What it builds (sources are read where they lie under /root/reference, patched IN MEMORY for the six torch-2.x
API renames SURVEY.md fact 0.5 lists -- kernel bodies untouched -- and written only into the git-ignored
oracle/_ref/src/; nothing of the reference enters the repository's history):

  oracle/_ref/build_cost_volume_cuda.so   snvc/extension/build_cost_volume/src/{BuildCostVolume.cpp,BuildCostVolume_cuda.cu}
  oracle/_ref/iou3d_nms_cuda.so           snvc/extension/iou3d_nms/src/{iou3d_cpu.cpp,iou3d_nms_api.cpp,iou3d_nms.cpp,iou3d_nms_kernel.cu}

Both are ordinary torch extension modules (pybind11), compiled with nvcc for sm_100 by explicit commands -- not the
reference's setup.py -- against the torch of this image.  nvcc cross-compiles without a GPU, so this runs in the
build container (where /root/reference exists); the .so files travel to the GPU box with the repo snapshot and
`tests/test_gpu_ref_pin.py` / `scripts/gpu_baselines.py` import them there through `load()`.  They are the checker
and the same-box GPU baseline, never the product: nothing under snvc_b200/ imports this file.
"""
import importlib.util
import os
import re
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = "/root/reference/snvc/extension"

MODULES = {
    "build_cost_volume_cuda": ("build_cost_volume/src", ["BuildCostVolume.cpp", "BuildCostVolume_cuda.cu"], ["-DWITH_CUDA"]),
    "iou3d_nms_cuda": ("iou3d_nms/src", ["iou3d_cpu.cpp", "iou3d_nms_api.cpp", "iou3d_nms.cpp", "iou3d_nms_kernel.cu"], []),
}
HEADERS = {"iou3d_nms_cuda": ["iou3d_cpu.h", "iou3d_nms.h"]}

# torch >= 2.0 API renames; applied to host code only, the __global__/__device__ bodies contain
# none of these tokens
SUBS = [
    (r"#include <THC/THC\.h>\s*\n", ""),
    (r"#include <THC/THCAtomics\.cuh>", "#include <ATen/cuda/Atomic.cuh>"),
    (r"#include <THC/THCDeviceUtils\.cuh>", "#include <ATen/cuda/DeviceUtils.cuh>\n#include <ATen/ceil_div.h>\n#include <c10/cuda/CUDAException.h>"),
    (r"THCCeilDiv", "at::ceil_div"),
    (r"THCudaCheck", "C10_CUDA_CHECK"),
    (r"\.type\(\)\.is_cuda\(\)", ".is_cuda()"),
    (r"AT_DISPATCH_FLOATING_TYPES\((\w+)\.type\(\)", r"AT_DISPATCH_FLOATING_TYPES(\1.scalar_type()"),
    (r"\.data<([^>]+)>\(\)", r".data_ptr<\1>()"),
    (r"data_ptr<long>", "data_ptr<int64_t>"),
    (r"long \* keep_data", "int64_t * keep_data"),
]


def patched(text):
    for pat, rep in SUBS:
        text = re.sub(pat, rep, text)
    return text


def available():
    return all(os.path.exists(os.path.join(OUT, m + ".so")) for m in MODULES)


def load(name):
    """Import oracle/_ref/<name>.so as a torch extension module (torch must be imported first: libc10/libtorch)."""
    import torch  # noqa: F401
    path = os.path.join(OUT, name + ".so")
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path}: run `python oracle/build_ref.py` in the build container (needs /root/reference)")
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def build(force=False, arch="100"):
    if not os.path.isdir(REF):
        raise SystemExit("oracle/build_ref.py: /root/reference is absent (GPU box) -- using the prebuilt oracle/_ref/*.so")
    import torch
    from torch.utils import cpp_extension as ce
    os.makedirs(os.path.join(OUT, "src"), exist_ok=True)
    inc = [f"-I{p}" for p in ce.include_paths("cuda")] + [f"-I{sysconfig.get_paths()['include']}"]
    abi = f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"
    libdirs = ce.library_paths("cuda")
    nvcc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")
    built = []
    for name, (sub, files, defs) in MODULES.items():
        so = os.path.join(OUT, name + ".so")
        srcs = [os.path.join(REF, sub, f) for f in files]
        if not force and os.path.exists(so) and all(os.path.getmtime(so) > os.path.getmtime(s) for s in srcs + [__file__]):
            built.append(so)
            continue
        wdir = os.path.join(OUT, "src", name)
        os.makedirs(wdir, exist_ok=True)
        for f in files + HEADERS.get(name, []):
            with open(os.path.join(REF, sub, f)) as fh:
                text = patched(fh.read())
            with open(os.path.join(wdir, f), "w") as fh:
                fh.write(text)
        common = inc + [abi, f"-DTORCH_EXTENSION_NAME={name}", "-DTORCH_API_INCLUDE_EXTENSION_H", "-std=c++17", "-O2"] + defs
        objs = []
        for f in files:
            o = os.path.join(wdir, f + ".o")
            if f.endswith(".cu"):
                # the reference's setup.py passes no arch and no math flags: nvcc defaults (-fmad=true), one real arch
                cmd = [nvcc, "-c", os.path.join(wdir, f), "-o", o, "-gencode", f"arch=compute_{arch},code=sm_{arch}",
                       "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC"] + common
            else:
                cmd = ["g++", "-c", os.path.join(wdir, f), "-o", o, "-fPIC", "-w"] + common
            subprocess.check_call(cmd)
            objs.append(o)
        link = ["g++", "-shared", "-o", so] + objs + [f"-L{d}" for d in libdirs] + \
               [f"-Wl,-rpath,{d}" for d in libdirs] + ["-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch",
                                                       "-ltorch_python", "-lcudart"]
        subprocess.check_call(link)
        built.append(so)
    return built


if __name__ == "__main__":
    for p in build(force="--force" in sys.argv):
        print(p)
