"""Build recipe for the CPU oracles.  TEST INFRASTRUCTURE ONLY -- the product path
(umr_b200/) never imports anything from oracle/.

* oracle B  (always):  oracle/softras_oracle.cpp  ->  oracle/_build/libsoftras_oracle.so
  our independent CPU restatement of the reference rasteriser (committed source).
* oracle A  (only where /root/reference exists, i.e. the build container):
  the reference's own device code, extracted verbatim at build time from
  external/SoftRas/soft_renderer/cuda/soft_rasterize_cuda_kernel.cu (its anonymous
  namespace, lines 22-659) into a temporary ref_device_code.inc and compiled for the host
  behind oracle/ref_host_shim.cpp  ->  oracle/_ref/libsoftras_ref_host.so
  (oracle/_ref/ is git-ignored: no reference source is ever committed; the built .so
  travels to the GPU box with the gpurun snapshot).

Usage:  python oracle/build_oracle.py [--force]
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_CU = "/root/reference/external/SoftRas/soft_renderer/cuda/soft_rasterize_cuda_kernel.cu"
REF_MD5 = "6fc801c6c2a494e523c53762737be737"  # reference @ 15ca8c87
CXXFLAGS = ["-O2", "-ffp-contract=off", "-fopenmp", "-std=c++17", "-shared", "-fPIC",
            "-fno-fast-math", "-Wall", "-Wno-unused-variable", "-Wno-maybe-uninitialized",
            "-Wno-unused-but-set-variable", "-Wno-uninitialized"]

LIB_B = os.path.join(HERE, "_build", "libsoftras_oracle.so")
LIB_A = os.path.join(HERE, "_ref", "libsoftras_ref_host.so")


def _newer(target, *sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources if os.path.exists(s))


def build_b(force=False):
    src = os.path.join(HERE, "softras_oracle.cpp")
    os.makedirs(os.path.dirname(LIB_B), exist_ok=True)
    if not force and _newer(LIB_B, src):
        return LIB_B
    subprocess.check_call(["g++", *CXXFLAGS, src, "-o", LIB_B])
    return LIB_B


def extract_reference_namespace(dst):
    """Copy the anonymous namespace (helpers + 3 kernels) of the reference .cu, verbatim."""
    with open(REF_CU, "rb") as f:
        raw = f.read()
    if hashlib.md5(raw).hexdigest() != REF_MD5:
        print("[oracle] warning: reference kernel file differs from the surveyed revision",
              file=sys.stderr)
    lines = raw.decode().splitlines(keepends=True)
    start = next(i for i, l in enumerate(lines) if l.startswith("namespace{"))
    end = next(i for i, l in enumerate(lines)
               if l.startswith("std::vector<at::Tensor> forward_soft_rasterize_cuda("))
    # last closing brace of the namespace before the host launchers
    while not lines[end - 1].startswith("}"):
        end -= 1
    with open(dst, "w") as f:
        f.write("// GENERATED at build time from the reference (lines %d-%d). DO NOT COMMIT.\n"
                % (start + 1, end))
        f.writelines(lines[start:end])


def build_a(force=False):
    """Returns the path of oracle A's library, or None when it cannot be (re)built here."""
    if not os.path.exists(REF_CU):
        return LIB_A if os.path.exists(LIB_A) else None  # GPU box: use the prebuilt file
    os.makedirs(os.path.dirname(LIB_A), exist_ok=True)
    shim = os.path.join(HERE, "ref_host_shim.cpp")
    if not force and _newer(LIB_A, shim, REF_CU):
        return LIB_A
    import tempfile
    # the extracted reference text only ever lives in a temp dir (never in the repo tree)
    with tempfile.TemporaryDirectory() as tmp:
        extract_reference_namespace(os.path.join(tmp, "ref_device_code.inc"))
        subprocess.check_call(["g++", *CXXFLAGS, "-w", "-I", tmp, shim, "-o", LIB_A])
    return LIB_A


def build_all(force=False):
    return build_b(force), build_a(force)


if __name__ == "__main__":
    b, a = build_all("--force" in sys.argv)
    print("oracle B:", b)
    print("oracle A:", a or "unavailable (no /root/reference and no prebuilt oracle/_ref)")
