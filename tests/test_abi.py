"""CPU suite, part 2: the drop-in boundary. The shared library loads, exports every function that
include/*.h declares, and - on a machine without a GPU - refuses to compute instead of falling back."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "include")
# declared by the reference only under __CUDACC__ and defined nowhere in it (reference fft_gpu.h:191-200)
NOT_EXPORTED = {"fft_gpu_set_cuda_options"}


def declared_functions(header):
    src = open(os.path.join(INC, header)).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    names = []
    for m in re.finditer(r"^[A-Za-z_][\w\s\*]*?[\s\*]((?:fft|fftb200)_\w+)\s*\([^;{]*\)\s*;", src, flags=re.M):
        if "typedef" in m.group(0) or "static" in m.group(0):
            continue
        names.append(m.group(1))
    return names


@pytest.mark.parametrize("header", ["fft_auto.h", "fft_gpu.h", "fftb200.h", "fftb200_ext.h", "fftb200_dist.h"])
def test_every_declared_symbol_is_exported(F, header):
    if not os.path.exists(os.path.join(INC, header)):
        pytest.skip(header + " not present")
    names = [n for n in declared_functions(header) if n not in NOT_EXPORTED]
    assert len(names) >= 4, (header, names)
    missing = [n for n in names if not hasattr(F.lib, n)]
    assert not missing, missing


# platform-specific hooks the reference declares only for other toolchains (Objective-C / nvcc) and defines nowhere for Linux gcc
NOT_DECLARED_HERE = {"fft_gpu_set_mps_options", "fft_gpu_set_cuda_options"}


@pytest.mark.parametrize("header", ["fft_auto.h", "fft_gpu.h"])
def test_prototypes_match_the_reference_headers(header):
    """Declaration by declaration against tests/golden/reference_prototypes.json (written from the reference's own headers by
    tests/golden/make_prototypes.py): same return type and parameter types for every public function, same enum / flag values."""
    import json
    import proto_parse
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_prototypes.json")))
    text = open(os.path.join(INC, header)).read()
    mine = proto_parse.prototypes(text)
    for name, sig in ref["prototypes"][header].items():
        if name in NOT_DECLARED_HERE:
            continue
        assert mine.get(name) == sig, (name, sig, mine.get(name))
    consts = proto_parse.constants(text)
    if header == "fft_auto.h":   # the direction enum lives in fft_common.h in both trees
        consts.update(proto_parse.constants(open(os.path.join(INC, "fft_common.h")).read()))
        want = dict(ref["constants"]["fft_auto.h"], **ref["constants"]["fft_common.h"])
    else:
        want = ref["constants"][header]
    for name, value in want.items():
        assert consts.get(name) == value, (name, value, consts.get(name))


def test_expected_public_names_present():
    auto = set(declared_functions("fft_auto.h"))
    gpu = set(declared_functions("fft_gpu.h"))
    for n in ("fft_auto", "fft_plan_dft_1d", "fft_plan_r2c_1d", "fft_execute", "fft_execute_dft", "fft_destroy_plan",
              "fft_plan_c2r_1d", "fft_plan_dft_2d", "fft_alloc_complex", "fft_free", "fft_version"):
        assert n in auto, n
    for n in ("fft_gpu_available", "fft_gpu_init", "fft_gpu_alloc", "fft_gpu_free", "fft_gpu_copy_h2d",
              "fft_gpu_copy_d2h", "fft_gpu_plan_1d", "fft_gpu_execute", "fft_gpu_destroy_plan", "fft_gpu_dft_1d",
              "fft_gpu_dft_1d_batch", "fft_gpu_set_device"):
        assert n in gpu, n


def test_library_does_not_link_the_oracle(F):
    """The product must not depend on the CPU checker: no oracle symbols, no libfftref / liboracle."""
    out = subprocess.run(["ldd", F.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "fftref" not in out
    syms = subprocess.run(["nm", "-D", "--defined-only", F.LIB_PATH], capture_output=True, text=True).stdout
    for cpu_algo in ("radix2_dit_fft", "split_radix_fft", "bluestein_fft", "oracle_fft_pow2", "naive_dft"):
        assert cpu_algo not in syms, cpu_algo


def test_headers_compile_as_c99_against_a_reference_style_caller(tmp_path):
    """A caller written against the reference's API (examples/demo_v2_features.c:65-90,130-143 pattern)
    compiles and links against this library with a plain C99 compiler."""
    src = tmp_path / "caller.c"
    src.write_text(r'''
#include "fft_auto.h"
#include "fft_gpu.h"
int main(void) {
    int n = 1024;
    complex_t* in = fft_alloc_complex(n); complex_t* out = fft_alloc_complex(n);
    for (int i = 0; i < n; i++) in[i] = cos(TWO_PI * 5 * i / n) + I * 0.0;
    printf("%s gpu=%d\n", fft_version(), fft_gpu_available());
    fft_plan_t p = fft_plan_dft_1d(n, in, out, -1, FFT_ESTIMATE);
    if (p) { fft_execute(p); fft_destroy_plan(p); printf("peak %.1f\n", cabs(out[5])); }
    else printf("no plan (no GPU)\n");
    int rc = fft_auto(in, out, n, -1);
    fft_gpu_plan_t gp = fft_gpu_plan_1d(n, 4, FFT_FORWARD);
    fft_gpu_execute(gp, NULL, NULL); fft_gpu_destroy_plan(gp);
    fft_free(in); fft_free(out);
    return (p != NULL) == (rc == 0) ? 0 : 1;
}
''')
    exe = tmp_path / "caller"
    libdir = os.path.join(ROOT, "fft-implementation-in-c_b200", "lib")
    subprocess.run(["gcc", "-std=c99", "-O2", "-Wall", "-Werror", "-I" + INC, str(src), "-o", str(exe),
                    "-L" + libdir, "-lfft_b200", "-Wl,-rpath," + libdir, "-lm"], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_programs_build_and_refuse_to_run_on_the_cpu(F):
    """programs/*.c (the reference's benchmark protocol, the API tour) are plain C99 callers of the public headers;
    without a device they say so and exit 2 - they never compute on the CPU."""
    bindir = os.path.join(ROOT, "fft-implementation-in-c_b200", "bin")
    for name in ("demo_drop_in", "benchmark_all_gpu"):
        exe = os.path.join(bindir, name)
        assert os.path.exists(exe), exe
        if F.lib.fft_gpu_available() == 1:
            continue
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 2 and "no CUDA device" in r.stdout


def test_wisdom_header_and_import_contract(F):
    import ctypes as C
    p = F.lib.fft_export_wisdom_to_string()
    assert C.string_at(p).decode().startswith("# FFT Wisdom v2.0.0\n")
    assert F.lib.fft_import_wisdom_from_string(b"# FFT Wisdom v2.0.0\nplan 1024 1 -1 0 # x\n") == 1
    assert F.lib.fft_import_wisdom_from_string(None) == 0


def test_no_cpu_fallback_without_gpu(F):
    if F.lib.fft_gpu_available() == 1:
        pytest.skip("a GPU is visible: covered by the gpu suite")
    x = np.ones(64, dtype=np.complex128)
    y = np.zeros_like(x)
    assert F.lib.fft_auto(F.ptr(x), F.ptr(y), 64, -1) == -1
    assert not y.any(), "output must be untouched when there is no device"
    assert not F.lib.fft_plan_dft_1d(64, F.ptr(x), F.ptr(y), -1, 0)
    assert not F.lib.fft_gpu_plan_1d(64, 2, -1)
    assert not F.lib.fft_gpu_alloc(64)
    assert F.lib.fft_gpu_dft_1d_batch(F.ptr(x), F.ptr(y), 32, 2, -1) == -1
    assert F.lib.fft_gpu_init(F.FFT_GPU_AUTO) == -1
    # the rows built on top of the hot path (c2r, 2-D, convolution / correlation) have no CPU fallback either
    r = np.zeros(64)
    assert not F.lib.fft_plan_c2r_1d(64, F.ptr(x), F.ptr(r), 0)
    assert not F.lib.fft_plan_dft_2d(8, 8, F.ptr(x), F.ptr(y), -1, 0)
    assert not F.lib.fft_gpu_plan_2d(8, 8, -1)
    assert F.lib.fft_gpu_dft_2d(F.ptr(x), F.ptr(y), 8, 8, -1) == -1
    assert F.lib.fft_gpu_convolution(F.ptr(x), 8, F.ptr(x), 8, F.ptr(y)) == -1
    assert F.lib.fft_gpu_circular_convolution(F.ptr(x), F.ptr(x), 16, F.ptr(y)) == -1
    assert F.lib.fft_gpu_cross_correlation(F.ptr(x), F.ptr(x), 16, F.ptr(y)) == -1
    assert F.lib.fft_gpu_autocorrelation(F.ptr(x), 16, F.ptr(y)) == -1
    assert not y.any()
    with pytest.raises(RuntimeError):
        F.require_gpu()


def test_argument_errors_do_not_need_a_gpu(F):
    x = np.ones(8, dtype=np.complex128)
    assert not F.lib.fft_plan_dft_1d(0, F.ptr(x), F.ptr(x), -1, 0)
    assert not F.lib.fft_plan_dft_1d(8, None, F.ptr(x), -1, 0)
    assert not F.lib.fft_plan_r2c_1d(12, F.ptr(x), F.ptr(x), 0)  # r2c is power-of-two only
    assert not F.lib.fft_gpu_plan_1d(-4, 1, -1)
    assert F.lib.fft_auto(None, None, 8, -1) == -1
    F.lib.fft_execute(None)           # void functions ignore NULL handles (reference fft_auto.c:242)
    F.lib.fft_destroy_plan(None)
    F.lib.fft_gpu_execute(None, None, None)
    F.lib.fft_gpu_destroy_plan(None)
    F.lib.fft_gpu_free(None)
    assert b"2.0.0" in F.lib.fft_version()
    p = F.lib.fft_alloc_complex(100)
    assert p and p % 64 == 0
    F.lib.fft_free(p)
