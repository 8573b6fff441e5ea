"""A compiled C consumer of include/ev2b.h against libev2b.so (catches drift between the header and the ctypes mirrors,
which the CPU suite can only compare by symbol name)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_c_program_links_and_steps_an_episode(tmp_path):
    from ev2gym_b200 import _lib
    _lib.load()                                   # builds the library if the sources are newer
    exe = str(tmp_path / "c_abi_smoke")
    libdir = os.path.dirname(_lib.LIB_PATH)
    libname = os.path.basename(_lib.LIB_PATH)
    subprocess.check_call(["gcc", "-O1", "-ffp-contract=off", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c_abi_smoke.c"), "-L", libdir, "-l:" + libname,
                           "-Wl,-rpath," + libdir, "-lm", "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "c_abi_smoke ok" in r.stdout


def test_c_consumer_compiles_against_the_header(tmp_path):
    """CPU side: the C program at least compiles against include/ev2b.h (syntax, types, every call's arity)."""
    obj = str(tmp_path / "c_abi_smoke.o")
    subprocess.check_call(["gcc", "-c", "-Wall", "-Werror=implicit-function-declaration", "-Werror=incompatible-pointer-types",
                           "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c_abi_smoke.c"), "-o", obj])
    assert os.path.getsize(obj) > 0


def test_c_consumer_runs_against_the_emulated_library(tmp_path):
    """CPU side: the same C program linked against the SIMT-emulated build of the same sources (tests/simt_emu) --
    the header's structs, the call sequence and the program's own known answers are checked without a GPU."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "simt_emu"))
    import emu_engine
    lib = emu_engine.build()
    exe = str(tmp_path / "c_abi_smoke_emu")
    subprocess.check_call(["gcc", "-O1", "-ffp-contract=off", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c_abi_smoke.c"), "-L", os.path.dirname(lib),
                           "-l:" + os.path.basename(lib), "-Wl,-rpath," + os.path.dirname(lib), "-lm", "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "c_abi_smoke ok" in r.stdout
