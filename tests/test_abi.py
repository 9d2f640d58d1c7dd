"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, and fails loudly (no CPU fallback) when there is no CUDA device.  No compute calls here."""
import ctypes as C
import os
import re
import subprocess

import pytest

from carskit_b200 import capi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "carskit_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cars_[a-z_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert header_functions() == sorted(capi.EXPORTS)


def test_library_exports_every_declared_symbol(cars_lib):
    for name in header_functions():
        assert hasattr(cars_lib, name), name
    assert b"sm_100a" in cars_lib.cars_version()


def test_struct_layout_matches_header(cars_lib):
    # sizes the C compiler produces for the header's structs (checked by compiling a probe)
    import subprocess
    import tempfile
    probe = r'''
#include "carskit_b200.h"
#include <stdio.h>
#include <stddef.h>
int main(void){printf("%zu %zu %zu %zu %zu %zu\n", sizeof(cars_desc), offsetof(cars_desc,nnz), offsetof(cars_desc,global_mean),
 offsetof(cars_desc,stream), sizeof(cars_model_arrays), sizeof(cars_stats));return 0;}
'''
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "p.c")
        open(src, "w").write(probe)
        exe = os.path.join(td, "p")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", exe, src])
        out = subprocess.check_output([exe]).decode().split()
    got = [C.sizeof(capi.CarsDesc), capi.CarsDesc.nnz.offset, capi.CarsDesc.global_mean.offset,
           capi.CarsDesc.stream.offset, C.sizeof(capi.CarsModelArrays), C.sizeof(capi.CarsStats)]
    assert [int(x) for x in out] == got


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_create_rejects_bad_descriptors(cars_lib):
    ts, _ = synth.make_training_set(10, 5, [2, 2], 40, seed=1)
    h = C.c_void_p()
    d = capi.make_desc(ts, capi.CAMF_CI, 8)
    d.abi_version = 99
    assert cars_lib.cars_create(C.byref(d), C.byref(h)) == -1
    assert b"abi_version" in cars_lib.cars_last_error(None)
    d = capi.make_desc(ts, capi.CAMF_CI, 0)
    assert cars_lib.cars_create(C.byref(d), C.byref(h)) == -5
    d = capi.make_desc(ts, capi.FM, 8, num_context_dims=2)
    assert cars_lib.cars_create(C.byref(d), C.byref(h)) == -1  # FM has its own entry points
    assert b"cars_fm_create" in cars_lib.cars_last_error(None)
    d = capi.make_desc(ts, capi.CAMF_CI, 8)
    assert cars_lib.cars_fm_create(C.byref(d), C.byref(h)) == -1
    d = capi.make_desc(ts, capi.FM, 8, num_context_dims=0)
    assert cars_lib.cars_fm_create(C.byref(d), C.byref(h)) == -1
    assert b"num_context_dims" in cars_lib.cars_fm_last_error(None)
    d = capi.make_desc(ts, capi.CAMF_CI, 8)
    d.ctx_ptr = None
    assert cars_lib.cars_create(C.byref(d), C.byref(h)) == -1
    assert not h.value
    assert cars_lib.cars_create(None, C.byref(h)) == -1


@pytest.mark.skipif(_has_cuda(), reason="this check is for boxes without a GPU")
def test_no_cpu_fallback_without_a_device(cars_lib):
    ts, _ = synth.make_training_set(10, 5, [2, 2], 40, seed=1)
    d = capi.make_desc(ts, capi.CAMF_CI, 8)
    h = C.c_void_p()
    rc = cars_lib.cars_create(C.byref(d), C.byref(h))
    assert rc == -2, cars_lib.cars_last_error(None)
    assert b"no CPU path" in cars_lib.cars_last_error(None)
    with pytest.raises(capi.CarsError):
        capi.Engine(d)
    d = capi.make_desc(ts, capi.FM, 8, num_context_dims=2)
    assert cars_lib.cars_fm_create(C.byref(d), C.byref(h)) == -2
    with pytest.raises(capi.CarsError):
        capi.FmEngine(d)


def test_product_never_imports_the_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "carskit_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle\b|#include\s+\"[^\"]*oracle", txt, flags=re.M):
                    bad.append(f)
    assert not bad, bad


@pytest.mark.skipif(_has_cuda(), reason="on a GPU box the client trains; this check is for boxes without a GPU")
def test_plain_c_client_links_and_reports_no_device(cars_lib, tmp_path):
    # the header is consumable from plain C and the library links without Python: examples/c_client.c
    libdir = os.path.join(ROOT, "carskit_b200")
    exe = str(tmp_path / "c_client")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "c_client.c"), "-L", libdir, "-lcarskit_b200",
                           f"-Wl,-rpath,{libdir}", "-o", exe])
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode == 3 and "cars_create: -2" in p.stdout and "no CPU path" in p.stdout


def test_product_sources_read_no_environment_variable():
    # every developer knob travels in cars_desc.tuning (csrc/tuning.h); only the -DCARS_TRACE developer build may getenv
    bad = []
    csrc = os.path.join(ROOT, "carskit_b200", "csrc")
    for f in sorted(os.listdir(csrc)):
        txt = open(os.path.join(csrc, f)).read()
        txt = re.sub(r"#ifdef CARS_TRACE.*?#(else|endif)", "", txt, flags=re.S)
        if re.search(r"\bgetenv\s*\(", txt):
            bad.append(f)
    assert not bad, bad
