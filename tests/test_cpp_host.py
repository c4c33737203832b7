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


def test_mpi_stand_in_semantics(tmp_path):
    """tests/cpp/test_mpi_shim.cpp under spnrun at 1, 2, 3, 5 ranks: ordering of same-tag messages of different sizes,
    zero-length and 24 MB messages, separate communicators, ragged collectives, rank-ordered reductions — what the
    reference's comm_t / ogs setup assume of MPI and the multi-rank pins (tests/test_ref_multirank_cpu.py) rest on."""
    import pytest
    exe = str(tmp_path / "test_mpi_shim")
    cmd = ["g++", "-O1", "-std=c++17", "-Wall", f"-I{ROOT}/include/shim", "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_mpi_shim.cpp"),
           os.path.join(ROOT, "streamparanumal_b200", "shim", "mpi_shim.cpp")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    spnrun = os.path.join(ROOT, "streamparanumal_b200", "bin", "spnrun")
    if not os.access(spnrun, os.X_OK):
        pytest.fail("streamparanumal_b200/bin/spnrun missing: run `make` first")
    for ranks in (1, 2, 3, 5):
        cmdline = [exe] if ranks == 1 else [spnrun, "-np", str(ranks), exe]
        r = subprocess.run(cmdline, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        assert f"mpi shim ok ({ranks} ranks)" in r.stdout


def test_occa_stand_in_semantics(tmp_path):
    """tests/cpp/test_occa_shim.cpp: include/shim/occa.hpp against a recording back end — json paths, malloc / slicing /
    copies with their kinds, offsets and async flags, null handles, stream switching, kernel argument packing
    (the OCCA surface of SURVEY.md section 8b)."""
    exe = str(tmp_path / "test_occa_shim")
    cmd = ["g++", "-O1", "-std=c++17", "-Wall", f"-I{ROOT}/include/shim", "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_occa_shim.cpp")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "occa shim ok" in r.stdout, r.stdout + r.stderr
