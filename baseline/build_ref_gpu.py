"""Build the REFERENCE's own soft_rasterize CUDA extension for sm_100a -- the "kernel to beat" and the
GPU oracle of SURVEY.md §8(c) -- from the sources where they lie under /root/reference.

Nothing of the reference is copied into the repository: the two source files are read from
/root/reference, the two documented API-drift shims are applied IN A TEMP DIR
(`-DAT_CHECK=TORCH_CHECK`; `faces.type()` -> `faces.scalar_type()` at kernel.cu:691,706,771, torch >= 1.5),
and only the built module lands in baseline/_ref/ (git-ignored; it travels to the GPU box with the gpurun
snapshot).  Cross-compiles without a GPU.

    python baseline/build_ref_gpu.py [--no-fma]     # --no-fma: -fmad=false build for the 3-way parity check
"""
import os
import subprocess
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/external/SoftRas/soft_renderer/cuda"
OUT = os.path.join(HERE, "_ref")


def build(no_fma=False):
    if not os.path.isdir(REF):
        return None
    import torch
    from torch.utils import cpp_extension
    name = "soft_rasterize_ref_nofma" if no_fma else "soft_rasterize_ref"
    os.makedirs(OUT, exist_ok=True)
    target = os.path.join(OUT, name + ".so")
    if os.path.exists(target):
        return target
    inc = cpp_extension.include_paths(device_type="cuda") if "device_type" in cpp_extension.include_paths.__code__.co_varnames \
        else cpp_extension.include_paths(cuda=True)
    inc.append(sysconfig.get_paths()["include"])
    incs = [x for p in inc for x in ("-I", p)]
    defs = ["-DAT_CHECK=TORCH_CHECK", "-DTORCH_EXTENSION_NAME=" + name, "-DTORCH_API_INCLUDE_EXTENSION_H",
            "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)]
    with tempfile.TemporaryDirectory() as tmp:
        cu = open(os.path.join(REF, "soft_rasterize_cuda_kernel.cu")).read()
        cu = cu.replace("AT_DISPATCH_FLOATING_TYPES(faces.type()", "AT_DISPATCH_FLOATING_TYPES(faces.scalar_type()")
        cpp = open(os.path.join(REF, "soft_rasterize_cuda.cpp")).read()
        cpp = cpp.replace("PYBIND11_MODULE(soft_rasterize, m)", "PYBIND11_MODULE(%s, m)" % name)
        open(os.path.join(tmp, "k.cu"), "w").write(cu)
        open(os.path.join(tmp, "b.cpp"), "w").write(cpp)
        common = ["-std=c++17", "-O3", "-Xcompiler", "-fPIC", "-w"] + defs + incs
        arch = ["-gencode", "arch=compute_100a,code=sm_100a"]
        subprocess.check_call(["nvcc", *arch, *common, *(["-fmad=false"] if no_fma else []), "-c",
                               os.path.join(tmp, "k.cu"), "-o", os.path.join(tmp, "k.o")])
        subprocess.check_call(["nvcc", *arch, *common, "-x", "cu", "-c", os.path.join(tmp, "b.cpp"), "-o",
                               os.path.join(tmp, "b.o")])
        libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
        subprocess.check_call(["nvcc", *arch, "-shared", "-o", target, os.path.join(tmp, "k.o"), os.path.join(tmp, "b.o"),
                               "-L", libdir, "-lc10", "-ltorch", "-ltorch_cpu", "-ltorch_python", "-lc10_cuda", "-ltorch_cuda",
                               "-Xlinker", "-rpath", "-Xlinker", libdir])
    return target


if __name__ == "__main__":
    print(build("--no-fma" in sys.argv))
