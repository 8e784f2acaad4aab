"""In-tree build of lib/libfft_b200.so: sm_100a CUDA kernels + C-ABI (csrc/) and the C99 host library
that carries the reference's public API (host/). Usage: python build.py [--force] [--verbose]."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
# FFTB200_VARIANT=<name> + FFTB200_NVCC_EXTRA="-D..." build an A/B library lib/ab_<name>.so in its own object dir
_VAR = os.environ.get("FFTB200_VARIANT", "")
OBJ = os.path.join(HERE, "build" + ("_" + _VAR if _VAR else ""))
LIB = os.path.join(HERE, "lib", "ab_%s.so" % _VAR if _VAR else "libfft_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
GCC = "/usr/bin/gcc"
NVCC_FLAGS = ["-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
              "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++"] + os.environ.get("FFTB200_NVCC_EXTRA", "").split()
# The host tables replicate the reference's rounding sequence explicitly (host/ref_twiddle.c): no
# fast-math, no implicit FMA contraction. x86-64-v3 = AVX2+FMA so fma() is one instruction.
HOST_FLAGS = ["-O2", "-std=c99", "-march=x86-64-v3", "-ffp-contract=off", "-fPIC", "-Wall", "-Wextra",
              "-I" + os.path.join(ROOT, "include")]


def _newer(src_list, target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in src_list)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    csrc = os.path.join(HERE, "csrc")
    host = os.path.join(HERE, "host")
    headers = [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cuh", ".h"))]
    headers += [os.path.join(ROOT, "include", f) for f in os.listdir(os.path.join(ROOT, "include"))]
    headers += [os.path.join(host, f) for f in os.listdir(host) if f.endswith(".h")]
    headers.append(os.path.abspath(__file__))
    jobs = []
    objs = []
    for f in sorted(os.listdir(csrc)):
        if f.endswith(".cu"):
            o = os.path.join(OBJ, f[:-3] + ".o")
            objs.append(o)
            if force or _newer([os.path.join(csrc, f)] + headers, o):
                jobs.append([NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) +
                            ["-c", os.path.join(csrc, f), "-o", o])
    for f in sorted(os.listdir(host)):
        if f.endswith(".c"):
            o = os.path.join(OBJ, f[:-2] + ".o")
            objs.append(o)
            if force or _newer([os.path.join(host, f)] + headers, o):
                jobs.append([GCC] + HOST_FLAGS + ["-c", os.path.join(host, f), "-o", o])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0 or verbose:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("build failed: " + " ".join(cmd))

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        list(ex.map(run, jobs))
    if jobs or not os.path.exists(LIB):
        run([NVCC, "-shared", "-o", LIB] + objs + ["-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++",
                                                    "-lpthread", "-lm", "-cudart", "static"])
    build_programs(force or bool(jobs))
    return LIB


def build_programs(force=False):
    """C99 callers of the public API (programs/*.c -> bin/): the reference's benchmark protocol and a tour of every entry
    point. Plain gcc against include/ and the shared library, exactly how a user of the reference would build."""
    if _VAR:
        return
    src_dir, bin_dir = os.path.join(HERE, "programs"), os.path.join(HERE, "bin")
    os.makedirs(bin_dir, exist_ok=True)
    for f in sorted(os.listdir(src_dir)):
        if not f.endswith(".c"):
            continue
        src, exe = os.path.join(src_dir, f), os.path.join(bin_dir, f[:-2])
        if force or _newer([src, LIB], exe):
            r = subprocess.run([GCC, "-std=c99", "-O2", "-Wall", "-Wextra", "-I" + os.path.join(ROOT, "include"), src, "-o", exe,
                                "-L" + os.path.dirname(LIB), "-lfft_b200", "-Wl,-rpath,$ORIGIN/../lib", "-lm", "-lpthread"],
                               capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("build failed: " + f + "\n" + r.stderr)


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
