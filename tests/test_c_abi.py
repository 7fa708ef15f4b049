"""The C-ABI from plain C (what a Fortran BIND(C) shim links against): tests/c/abi_check.c is compiled with gcc
against include/svfsi_b200.h and lib/libsvfsi_b200.so and run without a GPU -- struct layout of the flattened
FSILS_lsType, FSILS_LS_CREATE defaults, the host-only part of FSILS_LHS_CREATE, refusal before gpu_init_.
(No Fortran compiler exists in this image, so the ISO_C_BINDING module of INTEGRATION.md itself stays
uncompiled; a BIND(C) derived type has by definition the layout of the companion C struct checked here.)"""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("gcc") is None, reason="needs gcc")
def test_header_is_plain_c_and_the_library_links_from_c(tmp_path):
    from svfsi_b200 import api
    assert os.path.exists(api.LIB_PATH), "build the library first (__graft_entry__.build())"
    exe = str(tmp_path / "abi_check")
    libdir = os.path.dirname(api.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c", "abi_check.c"), "-o", exe, "-L", libdir,
                           "-lsvfsi_b200", f"-Wl,-rpath,{libdir}"])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and "ABI OK" in out.stdout, out.stdout + out.stderr
