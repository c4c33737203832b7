"""Compiles and runs tests/cpp/test_handles.cpp against the C++ host layer (no device needed): null-handle semantics of
deviceMemory / pinnedMemory, kernel_t / properties_t, settings_t, the process grid and the box labels."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_layer_handle_semantics(tmp_path):
    lib = os.path.join(ROOT, "streamparanumal_b200", "lib")
    assert os.path.exists(os.path.join(lib, "libspn_host.so")), "run `make` first"
    exe = str(tmp_path / "test_handles")
    cmd = ["g++", "-O1", "-std=c++17", "-fopenmp", "-Wall", f"-I{ROOT}/include", f"-I{ROOT}/streamparanumal_b200/host",
           f"-I{ROOT}/streamparanumal_b200/bs", "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_handles.cpp"),
           f"-L{lib}", "-lspn_host", "-lspn_b200", "-lspn_b200_comm", f"-Wl,-rpath,{lib}"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "host handle semantics ok" in r.stdout
