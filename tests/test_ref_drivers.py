"""The reference's own sources behind the drop-in boundary.

build/ref_drivers/BS0..BS8 are the reference's BS drivers, libs/core, libs/primitives, libs/ogs and libs/mesh compiled
UNCHANGED from /root/reference against include/shim/{occa.hpp,mpi.h} and linked with the product backend of the OCCA shim
(streamparanumal_b200/shim/occa_spn.cpp -> C ABI -> sm_100a kernels): platform_t::buildKernel, kernel_t::operator(),
deviceMemory<T>, device.finish / setStream — every call site of SURVEY.md section 8(b) — run as written
(`make ref-drivers`; the binaries travel to the GPU box, /root/reference does not).
oracle/_ref/BS0..BS8 are the same sources with the HOST backend (test infrastructure): the reference's `-m Serial` path.
CPU: the binaries exist, keep the reference CLI and refuse anything but CUDA.  GPU: every driver runs on the B200 and
prints the reference's line; BS5 at the BASELINE size; DOF counts equal those of the reference's Serial executables."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRV = os.path.join(ROOT, "build", "ref_drivers")
SERIAL = os.path.join(ROOT, "oracle", "_ref")
NUM = r"[-+0-9.eE]+"
HAVE_REF = os.path.isdir("/root/reference/BS")


def run(args, timeout=600):
    return subprocess.run(args, capture_output=True, text=True, timeout=timeout)


def need(path):
    if not os.access(path, os.X_OK):
        if HAVE_REF:
            pytest.fail(f"{path} missing although /root/reference is here: run `make ref-drivers` / `make -C oracle`")
        pytest.skip(f"{path} not built (it is compiled where /root/reference exists)")
    return path


def test_reference_sources_compile_against_the_shims():
    for n in range(9):
        need(os.path.join(DRV, f"BS{n}"))


def test_reference_cli_unchanged_and_no_cpu_mode():
    exe = need(os.path.join(DRV, "BS5"))
    r = run([exe, "--help"])  # (the reference exits with LIBP_ERROR after printing its usage: libs/core/settings.cpp)
    for flag in ("-m", "--mode", "-b", "--bytes", "-bmin", "-bmax", "-s", "--bstep", "-v", "-d", "-pl"):
        assert flag in r.stdout
    r = run([exe, "-m", "Serial", "-b", "4096"])       # the product backend implements CUDA only
    assert r.returncode != 0 and "no CPU fallback" in r.stderr
    r = run([need(os.path.join(DRV, "BS8")), "--help"])
    for flag in ("-e", "--elements", "-nx", "-ny", "-nz", "-p", "--degree"):
        assert flag in r.stdout


def test_reference_serial_executables_run_on_the_cpu():
    """oracle/_ref/BSn -m Serial: the reference's CPU path (host code + its own OKL kernels), the CPU arm of bench.py"""
    r = run([need(os.path.join(SERIAL, "BS5")), "-m", "Serial", "-b", str(24 * 100000)])
    assert r.returncode == 0, r.stderr
    m = re.search(rf"BS5 = \[(\d+) ({NUM}) ({NUM}) +({NUM})\]; %CG Update", r.stdout)
    assert m and int(m.group(1)) == 100000
    r = run([need(os.path.join(SERIAL, "BS8")), "-m", "Serial", "-nx", "3", "-ny", "2", "-nz", "2", "-p", "5"])
    assert r.returncode == 0, r.stderr
    m = re.search(rf"BS8 = \[(\d+), (\d+), ({NUM}), ({NUM}), +({NUM}), +({NUM})\]; % GatherScatter", r.stdout)
    assert m and int(m.group(1)) == 5 and int(m.group(2)) == 16 * 11 * 11


@pytest.mark.gpu
def test_reference_bs5_runs_on_the_b200_at_the_baseline_size():
    r = run([need(os.path.join(DRV, "BS5")), "-b", "1610612736"])   # N = 2^26 (BASELINE.json configs[0])
    assert r.returncode == 0, r.stderr
    m = re.fullmatch(rf"BS5 = \[(\d+) ({NUM}) ({NUM}) +({NUM})\]; %CG Update \[DOFs, elapsed, DOFs/s, BW \(GB/s\)\]\n", r.stdout)
    assert m, r.stdout
    assert int(m.group(1)) == 1 << 26
    assert float(m.group(4)) > 3000.0, f"BS5 through the reference driver: {m.group(4)} GB/s"


@pytest.mark.gpu
def test_every_reference_driver_runs_on_the_b200():
    titles = {1: "Copy", 2: "AXPY", 3: "Norm", 4: "Dot", 5: "CG Update"}
    per_entry = {1: 16, 2: 24, 3: 8, 4: 16, 5: 24}
    for n, title in titles.items():
        r = run([need(os.path.join(DRV, f"BS{n}")), "-b", str(48 << 20)])
        assert r.returncode == 0, r.stderr
        m = re.fullmatch(rf"BS{n} = \[(\d+) ({NUM}) ({NUM}) +({NUM})\]; %{title} \[DOFs, elapsed, DOFs/s, BW \(GB/s\)\]\n", r.stdout)
        assert m, r.stdout
        assert int(m.group(1)) == (48 << 20) // per_entry[n]
    r = run([need(os.path.join(DRV, "BS0")), "-n", "100"])
    assert r.returncode == 0 and re.search(r"BS0 = \[100 ", r.stdout), r.stdout + r.stderr
    # the ogs drivers: the reference's mesh_t + ogs_t::Setup + ogsOperator_t on the device kernels; DOFs must equal what the
    # reference's Serial executable reports for the same mesh
    for n, title in ((6, "Gather"), (7, "Scatter"), (8, "GatherScatter")):
        args = ["-nx", "6", "-ny", "5", "-nz", "4", "-p", "7"]
        r = run([need(os.path.join(DRV, f"BS{n}"))] + args)
        assert r.returncode == 0, r.stderr
        m = re.search(rf"BS{n} = \[(\d+), (\d+), ({NUM}), ({NUM}), +({NUM}), +({NUM})\]; % {title} ", r.stdout)
        assert m, r.stdout
        assert int(m.group(1)) == 7 and int(m.group(2)) == 43 * 36 * 29
        s = run([need(os.path.join(SERIAL, f"BS{n}")), "-m", "Serial"] + args)
        ms = re.search(rf"BS{n} = \[(\d+), (\d+),", s.stdout)
        assert ms and ms.group(2) == m.group(2)


SPNRUN = os.path.join(ROOT, "streamparanumal_b200", "bin", "spnrun")
CHECK_LINE = r"ogs_check: P=(\d+) N=(\d+) Nlocal=(\d+) Ngather=(\d+) gatherScatter=(\d+) gather=(\d+) scatter=(\d+) cases=(\d+)"


def _ogs_check(exe, ranks, mesh_args, mode=()):
    cmd = ([SPNRUN, "-np", str(ranks)] if ranks > 1 else []) + [exe] + list(mode) + list(mesh_args)
    r = run(cmd)
    assert r.returncode == 0, r.stdout + r.stderr
    m = re.search(CHECK_LINE, r.stdout)
    assert m, r.stdout + r.stderr
    assert int(m.group(1)) == ranks
    assert (int(m.group(5)), int(m.group(6)), int(m.group(7))) == (0, 0, 0), r.stdout
    # GatherScatter(Add): 4 types x 3 modes x K in {1,3} = 24; Gather x3 + Scatter x2 per K, for 2 types = 20
    assert int(m.group(8)) == 44
    return int(m.group(4))


@pytest.mark.parametrize("ranks", [1, 2, 4])
def test_ogs_check_is_sound_on_the_reference_serial_backend(ranks):
    """tests/cpp/ref_ogs_device_check.cpp on the HOST backend: the reference's deviceMemory path (its own OKL kernels compiled
    for the CPU) against its host-memory path — validates the check that the GPU test below relies on.
    (Rank counts are powers of two: the reference's Auto setup builds its crystal router too, and that one breaks at any
    other count on any MPI — the upper half leaves the level loop at libs/ogs/ogsCrystalRouter.cpp:455-466 before the
    forwarding exchange its partner in the lower half still posts at :505-514.)"""
    ng = _ogs_check(need(os.path.join(SERIAL, "ogs_check")), ranks, ["-nx", "3", "-ny", "3", "-nz", "2", "-p", "4"], mode=["-m", "Serial"])
    # global nodes of the mesh: the process grid multiplies the per-rank box (libs/mesh/meshSetupBoxHex3D.cpp)
    grid = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1)}[ranks]
    assert ng == (3 * grid[0] * 4 + 1) * (3 * grid[1] * 4 + 1) * (2 * grid[2] * 4 + 1)


@pytest.mark.gpu
@pytest.mark.parametrize("ranks", [1, 2, 4])
def test_reference_ogs_device_path_equals_its_host_path_on_the_b200(ranks):
    """the reference's mesh_t + ogs_t, unchanged, on the sm_100a backend: GatherScatter / Gather / Scatter on device memory
    (ogsOperator_t -> kernel_t -> C ABI; at P > 1 the reference's own pairwise exchange over the MPI stand-in, one GPU per
    rank where the box has them) equal the reference's host-memory overloads bit for bit"""
    # fewer GPUs than ranks: the reference's own platform_t shares them (device_id % deviceCount with its oversubscription
    # warning, libs/core/platformDeviceConfig.cpp:88-95); its exchange is host-staged MPI, so nothing spins on the device
    exe = need(os.path.join(DRV, "ogs_check"))
    _ogs_check(exe, ranks, ["-nx", "6", "-ny", "5", "-nz", "4", "-p", "7"])
    _ogs_check(exe, ranks, ["-nx", "3", "-ny", "2", "-nz", "2", "-p", "1"])


def test_reference_device_gather_scatter_ignores_the_op():
    """A quirk of the reference the check exposes: ogsOperator_t::GatherScatter(deviceMemory) launches
    gatherScatterKernel[type][Add] for every op (libs/ogs/ogsOperator.cpp:443-480), so its device and host paths differ for
    Mul / Max / Min on the interior rows — on its own Serial backend, no kernel of ours involved.  The native layer
    (streamparanumal_b200/host) honours the op, as the reference's host path and its header (include/ogs.hpp:75-90) do."""
    exe = need(os.path.join(SERIAL, "ogs_check"))
    env = dict(os.environ, OGS_CHECK_ALL_OPS="1")
    r = subprocess.run([exe, "-m", "Serial", "-nx", "2", "-ny", "2", "-nz", "2", "-p", "2"], capture_output=True, text=True, env=env, timeout=600)
    m = re.search(CHECK_LINE, r.stdout)
    assert m and int(m.group(5)) > 0 and (int(m.group(6)), int(m.group(7))) == (0, 0), r.stdout
