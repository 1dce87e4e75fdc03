"""The drop-in boundary is a C ABI: include/whisper_b200.h must compile as C99 and a plain C program must link against
libwhisper_b200.so and get status codes + messages back (no GPU needed for these entry points)."""
import os
import shutil
import subprocess

import pytest

from conftest import ROOT


@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not available")
def test_header_is_plain_c_and_library_links_from_c(tmp_path, lib):
    exe = str(tmp_path / "c_abi_demo")
    libdir = os.path.join(ROOT, "whisper_trtllm_b200")
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "c_abi_demo.c"), "-o", exe, "-L", libdir, "-lwhisper_b200", f"-Wl,-rpath,{libdir}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "wb_model_create(NULL) -> -1" in r.stdout and "null" in r.stdout
