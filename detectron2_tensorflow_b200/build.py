"""In-tree build of libd2b200.so (hand-written sm_100a kernels + the C-ABI).

    python -m detectron2_tensorflow_b200.build [--force]

nvcc cross-compiles without a GPU.  The library has no torch/TF dependency
(static cudart); -fmad=false keeps one IEEE rounding per written fp32 operation,
which the bit-parity contract with the reference's op order relies on.
"""
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OUT_DIR = os.path.join(PKG, "lib")
LIB = os.path.join(OUT_DIR, "libd2b200.so")

NVCC_FLAGS = [
    "-std=c++17", "-O3",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-O3",
    "-cudart", "static",
    "-I", os.path.join(ROOT, "include"),
    "-I", CSRC,
]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(ROOT, "include", "d2b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OUT_DIR, exist_ok=True)
    objs = []
    procs = []
    obj_dir = os.path.join(PKG, "build")
    os.makedirs(obj_dir, exist_ok=True)
    env = dict(os.environ)
    # the image may point CC/CXX at a wrapper; nvcc wants the system host compiler
    host = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    for src in sources():
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        extra = os.environ.get("D2B_EXTRA_NVCC", "").split()  # experiments only (e.g. -DD2B_RA_THREADS=128)
        cmd = [nvcc, "-ccbin", host] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(out.decode(errors="replace"))
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"nvcc failed on {src}\n")
    if failed:
        raise RuntimeError("libd2b200 build failed")
    cmd = [nvcc, "-ccbin", host, "-shared", "-cudart", "static", "-Wno-deprecated-gpu-targets", "-o", LIB] + objs
    subprocess.check_call(cmd, env=env)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
