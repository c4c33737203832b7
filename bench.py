#!/usr/bin/env python
"""bench.py — the driver-facing benchmark of the BS hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BS8, the reference's fused ogs gather+scatter benchmark
(BS/BS8/src/bs8Run.cpp:29-78) on the configuration BASELINE.json quotes it on — a box
of 64x64x64 degree-7 hex elements PER RANK (2^27 local nodes, FP64, K=1, Add, Sym).
One "step" is one mesh.ogs.GatherScatter(o_q, 1, Add, Sym) call.  `value` is the
reference's DOFs/s figure (NgatherGlobal / time per call: bs8Run.cpp:67-75) summed over
all ranks, in GDOF/s, with q resident in HBM; `e2e` is the same metric through the
host-memory overload (pinned host q -> H2D -> gatherScatter -> D2H every step).
With --suite (default on at 1 GPU) the same JSON line carries a `suite` object with the
effective GB/s of every BS1-BS8 kernel at its BASELINE.json size and BS0 us/launch,
bytes counted the way the reference counts them.

At N > 1 the line also carries `strong` (the same call with 64^3 elements IN TOTAL, BASELINE.json configs[3]) and, at every
N, `tail` (BS0 launch latency, BS5 at 1 KB - 1 MB, BS8 at 2^3 - 8^3 elements per rank: configs[4]).  Before anything is timed
every rank runs checked calls (`parity_checked`): all-ones -> the node multiplicity computed from the lattice position alone,
a label-valued vector -> multiplicity x value (both exact, both need the halo exchange to be right), and at one rank the
full-size result against the CPU oracle on the plan's CSR.

--impl reference times the reference's OWN CPU path on this box's host cores: oracle/_ref/libref_ogs.so is the reference's
libs/core + primitives + ogs compiled unchanged (oracle/Makefile), so the call timed is its mesh.ogs.GatherScatter(q, 1, Add, Sym)
on the SAME 64^3 degree-7 mesh — the OpenMP-threaded host-memory overload and the `-m Serial` device path, the faster one
reported.  Without that library (no /root/reference at build time) the OpenMP restatement in oracle/ stands in (kind "port").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SMI_QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")


# --------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------
def bs8_bytes(Nt, Ng):
    """BS/BS8/src/bs8Run.cpp:58-61"""
    return 4 * (Ng + 1) + 4 * Nt + 8 * Nt + 8 * Nt


def measured_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch from the round's ncu --set full capture
    (profiles/r02_traffic.json, written by tools/ncu_traffic.py from the .ncu-rep); None when no capture is committed"""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            return json.load(f).get(key)
    except Exception:
        return None


def bs6_bytes(Nt, Ng):
    """BS/BS6/src/bs6Run.cpp:60-63 (BS7 counts the same total: bs7Run.cpp:60-63)"""
    return 4 * (Ng + 1) + 4 * Nt + 8 * Nt + 8 * Ng


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled during the timed region."""

    def __init__(self, device_index):
        self.rows = []
        self.proc = None
        self.idx = device_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.idx}", f"--query-gpu={SMI_QUERY}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smmax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.rows:
            f = [x.strip() for x in row.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smmax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(smmax) if smmax else None,
                "power_w_max": max(power) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------
# reference arm (CPU)
# --------------------------------------------------------------------------------------
def reference_cpu(nx, p, steps, warmup, budget_s=None):
    """Times the reference's CPU path for BS8 on an nx^3 degree-p box.  Returns (results, Nt, Ng): results maps a variant
    name to {"t": seconds per call, "kind": "reference" | "port", "cores": threads}.  budget_s: keep calling for about
    that many seconds instead of a fixed step count."""
    import spn_paths  # noqa: F401
    from streamparanumal_b200 import host, synth
    from oracle import pyoracle, refogs
    cores = os.cpu_count() or 1
    ids, nglobal = host.box_hex_ids(0, 1, nx, nx, nx, p)
    Nt = ids.size
    results = {}
    if refogs.available():
        r = refogs.RefOgs(ids, kind=refogs.SIGNED, unique=True)   # libp::ogs::ogs_t::Setup, as mesh_t::GatherScatterSetup calls it
        Ng = r.NgatherGlobal
        assert Ng == nglobal
        for name, path in (("reference host-memory overload (OpenMP host operator)", 0), ("reference -m Serial (OKL kernel)", 1)):
            if budget_s is not None:
                t1, th = r.time_gather_scatter(path, 1, 1)
                n = max(1, min(200, int(budget_s / max(t1, 1e-9))))
                t, th = r.time_gather_scatter(path, n, 0)
            else:
                t, th = r.time_gather_scatter(path, steps, warmup)
            results[name] = {"t": t, "kind": "reference", "cores": th if path == 0 else 1}
        r.close()
    else:
        o = host.Ogs(ids.copy(), host.Comm(0, 1, None), kind=host.SIGNED, unique=True, platform=None)
        rs, ci = o.array("local_rowStartsT"), o.array("local_colIdsT")
        Ng = o.NgatherGlobal
        q = synth.splitmix64(Nt, 3)
        fn = lambda: pyoracle.ogs_gather_scatter(rs, ci, rs, ci, q, threads=cores)
        for _ in range(max(1, warmup if budget_s is None else 1)):
            fn()
        t0 = time.perf_counter()
        n = 0
        while (n < steps) if budget_s is None else (time.perf_counter() - t0 < budget_s):
            fn()
            n += 1
        results["OpenMP restatement of the host operator (oracle/)"] = {"t": (time.perf_counter() - t0) / n, "kind": "port", "cores": cores}
    return results, Nt, Ng


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    nx = args.ref_nx if args.ref_nx else args.nx
    results, Nt, Ng = reference_cpu(nx, args.p, args.steps, args.warmup)
    best = min(results, key=lambda k: results[k]["t"])
    t, kind, cores = results[best]["t"], results[best]["kind"], results[best]["cores"]
    value = Ng / t * 1e-9
    sample = (f"{best}: mesh.ogs.GatherScatter(q, 1, Add, Sym) on the {nx}^3 degree-{args.p} hex box (Nt={Nt}, Ng={Ng}), "
              f"{args.steps} calls after {args.warmup} warm-ups, q resident in the reference's own containers")
    line = {
        "impl": "reference", "metric": "BS8 gather+scatter throughput (DOFs/s, NgatherGlobal per call time)",
        "value": value, "unit": "GDOF/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"BS8 ogs_t::GatherScatter(Add,Sym,K=1), hex p={args.p}, {nx}^3 elements "
                               f"(mpirun -np 1 ./BS8 -m Serial -nx {nx} -ny {nx} -nz {nx} -p {args.p})",
                   "nx": nx, "p": args.p, "Nt": Nt, "Ng": Ng},
        "effective_gbs": bs8_bytes(Nt, Ng) / t * 1e-9,
        "cpu_baseline": {"value": value, "unit": "GDOF/s", "cores": cores, "kind": kind, "sample": sample,
                         "all": {k: {"GDOF/s": Ng / v["t"] * 1e-9, "GB/s": bs8_bytes(Nt, Ng) / v["t"] * 1e-9, "cores": v["cores"],
                                     "kind": v["kind"]} for k, v in results.items()}},
        "e2e": {"value": value, "unit": "GDOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------
def cuda_time(fn, iters, warm, torch):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / iters


def run_suite(torch, K, peak, ogs_ctx):
    """BS0-BS7 at the BASELINE.json sizes (BS8 is the headline and is added by the caller)."""
    from streamparanumal_b200 import host
    out = {}
    dev = "cuda"

    def add(name, bytes_per_call, t, dofs):
        out[name] = {"ms": t * 1e3, "GB/s": bytes_per_call / t * 1e-9, "frac_of_measured_peak": bytes_per_call / t * 1e-9 / peak,
                     "frac_of_8TBs": bytes_per_call / t * 1e-9 / 8000.0, "GDOF/s": dofs / t * 1e-9, "bytes": bytes_per_call}

    # BS1-BS4 at the top of the sweep (4 GiB moved), BS5 at N=2^26 (config 1)
    B = 1 << 32
    n1 = B // 16
    a = torch.rand(n1, dtype=torch.float64, device=dev)
    b = torch.empty_like(a)
    add("BS1 copy", 16 * n1, cuda_time(lambda: K.copy(n1, a, b), 20, 5, torch), n1)
    n2 = B // 24
    add("BS2 axpby", 24 * n2, cuda_time(lambda: K.axpby(n2, 1.0, a, 1.0, b), 20, 5, torch), n2)
    n4 = B // 16
    o = torch.zeros(1, dtype=torch.float64, device=dev)
    add("BS4 dot", 16 * n4, cuda_time(lambda: K.dot(n4, a, b, o), 20, 5, torch), n4)
    del b
    n3 = B // 8
    a3 = torch.cat([a, a])
    del a
    add("BS3 norm", 8 * n3, cuda_time(lambda: K.norm2(n3, a3, o), 20, 5, torch), n3)
    del a3
    n5 = 1 << 26
    Ap = torch.rand(n5, dtype=torch.float64, device=dev)
    r = torch.zeros(n5, dtype=torch.float64, device=dev)
    add("BS5 cg_update", 24 * n5, cuda_time(lambda: K.cg_update(n5, Ap, 1.0, r, o), 20, 5, torch), n5)
    # BS0: launch latency per launch, host wall clock around 1000 launches (bs0Run.cpp:70-77)
    for _ in range(100):
        K.bs0(1, Ap, r)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(1000):
        K.bs0(1, Ap, r)
    torch.cuda.synchronize()
    out["BS0 launch"] = {"us_per_launch": (time.perf_counter() - t0) / 1000 * 1e6}
    del Ap, r
    # BS6 / BS7 on the headline mesh, through ogs_t as bs6Run.cpp:47 / bs7Run.cpp:47 call it
    ogs, Nt, Ng, q = ogs_ctx
    gq = torch.empty(Ng, dtype=torch.float64, device=dev)
    add("BS6 gather", bs6_bytes(Nt, Ng), cuda_time(lambda: ogs.Gather(gq, q, 1, host.ADD, host.TRANS), 20, 5, torch), Ng)
    add("BS7 scatter", bs6_bytes(Nt, Ng), cuda_time(lambda: ogs.Scatter(q, gq, 1, host.NOTRANS), 20, 5, torch), Ng)
    return out


def parity_checks(torch, host, ogs, labels, rank, world, nx, p, full_oracle):
    """Independent expectations for ogs_t::GatherScatter(Add, Sym) on this rank's brick, exact in FP64:
      ones        -> the multiplicity of every node in the GLOBAL mesh, computed from the lattice position alone;
      g(label)    -> multiplicity x g(label)   (g integer-valued: a slot paired with the wrong label shows up);
      at one rank -> the CPU oracle walking the plan's CSR on seeded uniform(-1,1) data, bit for bit (full size)."""
    from streamparanumal_b200 import synth
    grid, rxyz = host.rank_grid(rank, world)
    mult = synth.box_multiplicity(rxyz, grid, (nx, nx, nx), p).astype(np.float64)
    n = labels.size
    out = {"checks": []}
    q = torch.ones(n, dtype=torch.float64, device="cuda")
    for rep in range(3):   # three calls: both receive halves of the peer exchange and the replayed graph are exercised
        q.fill_(1.0)
        ogs.GatherScatter(q, 1, host.ADD, host.SYM)
    ok1 = bool(np.array_equal(q.cpu().numpy(), mult))
    out["checks"].append({"all-ones -> lattice multiplicity": ok1})
    g = ((labels % 1021) - 510).astype(np.float64)
    q.copy_(torch.from_numpy(g))
    ogs.GatherScatter(q, 1, host.ADD, host.SYM)
    ok2 = bool(np.array_equal(q.cpu().numpy(), mult * g))
    out["checks"].append({"g(label) -> multiplicity * g(label)": ok2})
    ok3 = True
    if full_oracle:
        from oracle import pyoracle   # the checker, not the thing measured
        rs, ci = ogs.array("local_rowStartsT"), ogs.array("local_colIdsT")
        x = synth.splitmix64(n, 77)
        q.copy_(torch.from_numpy(x))
        ogs.GatherScatter(q, 1, host.ADD, host.SYM)
        exp = pyoracle.ogs_gather_scatter(rs, ci, rs, ci, x, threads=os.cpu_count() or 1)
        ok3 = bool(np.array_equal(q.cpu().numpy().view(np.int64), exp.view(np.int64)))
        out["checks"].append({"full-size result == CPU oracle on the plan's CSR (bit for bit)": ok3})
    out["ok"] = ok1 and ok2 and ok3
    del q
    return out


def strong_scaling(torch, dist, host, comm, platform, rank, world, p, timed):
    """BASELINE.json configs[3], strong: 64^3 elements IN TOTAL split over the process grid (-nx 32 at 8 ranks)"""
    from streamparanumal_b200 import synth
    grid, _ = host.rank_grid(rank, world)
    n = tuple(64 // g for g in grid)
    ids, _ = host.box_hex_ids(rank, world, n[0], n[1], n[2], p)
    ogs = host.Ogs(ids, comm, kind=host.SIGNED, unique=True, platform=platform)
    q = torch.from_numpy(synth.splitmix64(ids.size, 5 + rank)).cuda()
    t = timed(lambda: ogs.GatherScatter(q, 1, host.ADD, host.SYM), 50, 5)
    Ng, Nt = ogs.NgatherGlobal, ids.size * world
    res = {"elements_per_rank": list(n), "Ng": Ng, "ms_per_step": t * 1e3, "GDOF/s": Ng / t * 1e-9,
           "effective_gbs": bs8_bytes(Nt, Ng) / t * 1e-9, "note": "64^3 degree-7 elements in total; speed-up over 1 GPU = this GDOF/s / the 1-GPU value"}
    ogs.close()
    return res


def latency_tail(torch, dist, host, K, comm, platform, rank, world, p, timed):
    """BASELINE.json configs[4]: BS0 launch latency, BS5 at 1 KB - 1 MB (with the scalar allreduce at N > 1), BS8 at 2^3 - 8^3
    elements per rank — the launch-bound end of the sweeps (BS/BS8/runBS8.sh:76-90, BS/BS5/runBS5.sh)"""
    from streamparanumal_b200 import synth
    out = {"BS8_us": {}, "BS5_us": {}}
    for nx in (2, 4, 8):
        ids, _ = host.box_hex_ids(rank, world, nx, nx, nx, p)
        ogs = host.Ogs(ids, comm, kind=host.SIGNED, unique=True, platform=platform)
        q = torch.from_numpy(synth.splitmix64(ids.size, 9 + rank)).cuda()
        t = timed(lambda: ogs.GatherScatter(q, 1, host.ADD, host.SYM), 200, 20)
        out["BS8_us"][f"nx={nx}"] = round(t * 1e6, 2)
        ogs.close()
    o = torch.zeros(1, dtype=torch.float64, device="cuda")
    for nbytes in (1024, 16384, 262144, 1048576):
        n = nbytes // 24
        Ap = torch.rand(n, dtype=torch.float64, device="cuda")
        r = torch.zeros(n, dtype=torch.float64, device="cuda")

        def call():
            K.cg_update(n, Ap, 1.0, r, o)
            if world > 1:
                comm.allreduce_device(o)
        out["BS5_us"][f"{nbytes} B"] = round(timed(call, 200, 20) * 1e6, 2)
    a = torch.zeros(8, dtype=torch.float64, device="cuda")
    for _ in range(100):
        K.bs0(1, a, a)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(1000):
        K.bs0(1, a, a)
    torch.cuda.synchronize()
    out["BS0_us_per_launch"] = round((time.perf_counter() - t0) / 1000 * 1e6, 2)
    return out


def run_b200(args):
    import ctypes
    import torch
    import torch.distributed as dist
    from streamparanumal_b200 import kernels as K
    from streamparanumal_b200 import _lib, host
    from streamparanumal_b200 import synth as util

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: streamparanumal_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    lib = _lib.kernel_lib()  # fail loudly if the sm_100a library is missing
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    peak, peak_src = measured_peak()

    # ---- setup: box mesh labels -> ogs_t (host, once; not timed — the reference does not time it either:
    #      mesh_t::Setup + ogs.Setup run before bs8_t::Run, BS/BS8/bs8Main.cpp:45-51) ----
    nx, p = args.nx, args.p
    t_setup = time.perf_counter()
    comm = host.Comm()  # RANK / WORLD_SIZE / MASTER_PORT: host collectives over local sockets, device data over NCCL
    platform = host.Platform(comm, local_rank)
    ids, nglobal = host.box_hex_ids(rank, world, nx, nx, nx, p)
    ogs = host.Ogs(ids, comm, kind=host.SIGNED, unique=True, platform=platform)
    t_setup = time.perf_counter() - t_setup
    Nt_local = ids.size
    Nt, Ng = Nt_local * world, ogs.NgatherGlobal  # bs8Run.cpp:52-56: Nelements*Np summed over ranks, ogs.NgatherGlobal
    assert Ng == nglobal, "ogs setup lost labels"
    halo = ogs.exchange_info(host.SYM)
    labels = np.abs(ogs.ids)
    del ids

    # ---- checked calls before anything is timed (every rank; the exchange the timed loop uses) ----
    parity = parity_checks(torch, host, ogs, labels, rank, world, nx, p, full_oracle=(world == 1))
    del labels
    q = torch.from_numpy(util.splitmix64(Nt_local, 3 + rank)).cuda()

    step = lambda: ogs.GatherScatter(q, 1, host.ADD, host.SYM)  # mesh.ogs.GatherScatter(o_q, 1, ogs::Add, ogs::Sym)

    def timed(fn, steps, warm):
        """`steps` calls of fn between barrier + device sync; CUDA events; max over ranks (seconds per call)"""
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1) * 1e-3 / steps], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- timed region: device-resident ----
    sampler = ClockSampler(local_rank)
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    t_step = timed(step, args.steps, 0)
    # K steps last only ~25 ms, shorter than one nvidia-smi sampling period: keep the same step loop running (untimed) for
    # about 1.5 s so that the clock / throttle samples are taken under the load of the measured kernel
    n_load = max(args.steps, int(1.5 / max(t_step, 1e-6)))
    for _ in range(n_load):
        step()
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["sampled_over"] = f"the timed region ({args.steps} steps) + {n_load} more untimed steps of the same loop"

    # ---- e2e: host buffers through the host-buffer call of the host layer (ogs_t::GatherScatterStart/Finish on pinned
    #      host vectors: every step copies its q host->device, runs ogs_t::GatherScatter and copies the result back).
    #      Two host vectors alternate, so the upload of step s+1 overlaps the download of step s (full-duplex PCIe);
    #      every step's H2D and D2H are inside the timed region. ----
    nbytes = Nt_local * 8
    hbuf, hnp = [], []
    for i in range(2):
        hp = ctypes.c_void_p()
        _lib.check(lib.spn_host_malloc(ctypes.byref(hp), nbytes), "host_malloc")
        hbuf.append(hp)
        hnp.append(np.ctypeslib.as_array(ctypes.cast(hp, ctypes.POINTER(ctypes.c_double)), shape=(Nt_local,)))
        hnp[i][:] = util.splitmix64(Nt_local, 11 + 2 * rank + i)
    # one checked call: the host-buffer path must return what the device path computes on the same input
    q_chk = torch.from_numpy(hnp[0].copy()).cuda()
    ogs.GatherScatter(q_chk, 1, host.ADD, host.SYM)
    ogs.HostGatherScatterStart(hbuf[0], K.ogs_type(q_chk))
    ogs.HostGatherScatterFinish(hbuf[0])
    assert np.array_equal(hnp[0], q_chk.cpu().numpy()), "host-buffer GatherScatter differs from the device path"
    del q_chk
    ftype = K.ogs_type(q)

    def e2e_run(steps):
        for s_ in range(steps):
            ogs.HostGatherScatterStart(hbuf[s_ % 2], ftype)
            if s_ > 0:
                ogs.HostGatherScatterFinish(hbuf[(s_ - 1) % 2])
        ogs.HostGatherScatterFinish(hbuf[(steps - 1) % 2])

    e2e_steps = max(4, min(args.steps, 10))
    e2e_run(2)  # warm-up (allocates the staging buffers)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_run(e2e_steps)          # returns when the last result is back in host memory
    e1.record()
    torch.cuda.synchronize()
    t_ = torch.tensor([e0.elapsed_time(e1) * 1e-3 / e2e_steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_, op=dist.ReduceOp.MAX)
    t_e2e = float(t_.item())

    # copy-only ceiling of the same step: the two pinned vectors moved H2D and D2H concurrently on two streams, no kernel.
    # (On this pool the boxes expose ONE NUMA node — nvidia-smi topo: every GPU "CPU Affinity 0-31, NUMA 0" — so there is
    # nothing to bind ranks to; what the e2e number can reach at N ranks is what the host side sustains for N such copies.)
    d_a = torch.empty(Nt_local, dtype=torch.float64, device="cuda")
    d_b = torch.empty(Nt_local, dtype=torch.float64, device="cuda")
    s_up, s_dn = ctypes.c_void_p(), ctypes.c_void_p()
    _lib.check(lib.spn_stream_create(ctypes.byref(s_up)), "stream")
    _lib.check(lib.spn_stream_create(ctypes.byref(s_dn)), "stream")

    def copy_pair():
        _lib.check(lib.spn_memcpy_h2d(d_a.data_ptr(), hbuf[0], nbytes, s_up, 1), "h2d")
        _lib.check(lib.spn_memcpy_d2h(hbuf[1], d_b.data_ptr(), nbytes, s_dn, 1), "d2h")

    for _ in range(2):
        copy_pair()
    lib.spn_stream_sync(s_up); lib.spn_stream_sync(s_dn)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(4):
        copy_pair()
    lib.spn_stream_sync(s_up); lib.spn_stream_sync(s_dn)
    t_copy = torch.tensor([(time.perf_counter() - t0) / 4], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_copy, op=dist.ReduceOp.MAX)
    t_copy = float(t_copy.item())
    lib.spn_stream_destroy(s_up); lib.spn_stream_destroy(s_dn)
    del d_a, d_b
    for hp in hbuf:
        _lib.check(lib.spn_host_free(hp), "host_free")

    bytes_step = bs8_bytes(Nt, Ng)          # whole job, counted as bs8Run.cpp:58-61 does
    bytes_rank = bytes_step / world
    # launches per call: interior gatherScatter; + halo gather/pack and combine/scatter (ours) when the mesh is partitioned
    launches_per_step = 1 if world == 1 else 3  # halo gather+pack, interior gatherScatter, halo combine+scatter
    line = {
        "metric": "BS8 gather+scatter throughput (DOFs/s, NgatherGlobal per call time)",
        "value": Ng / t_step * 1e-9, "unit": "GDOF/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"BS8 ogs_t::GatherScatter(Add,Sym,K=1), hex p={p}, {nx}^3 elements per rank "
                               f"(mpirun -np {world} ./BS8 -nx {nx} -ny {nx} -nz {nx} -p {p})",
                   "nx": nx, "p": p, "Nt_per_rank": Nt_local, "Nt": Nt, "Ng": Ng, "Ngather_rank0": ogs.Ngather,
                   "tiles_rank0": ogs.local_tiles, "halo_rows_rank0": ogs.NhaloT, "halo_send_rank0": halo["Nsend"],
                   "neighbours_rank0": halo["NranksSend"],
                   "exchange": "none (one rank)" if world == 1 else (
                       "peer memory: the halo pack kernel stores partials straight into the neighbours' receive buffers over NVLink (P2P, CUDA IPC) and raises flags; the combine kernel waits on them after the interior kernel"
                       if ogs.peer_exchange else "NCCL send/recv group on a comm stream, overlapped with the interior kernel"),
                   "l2": "q (1.07 GB) + ids (0.54 GB) per call exceed the 126 MB L2; no flush needed",
                   "host_setup_s": round(t_setup, 2)},
        "effective_gbs": bytes_step / t_step * 1e-9,
        "roofline": {"bound": "hbm", "kernel": "ogs_tile_pf_kernel<double,Add,GATHER_SCATTER,U=4,128 threads,8 CTAs/SM>",
                     "achieved": bytes_rank / t_step * 1e-9, "peak": peak, "peak_source": peak_src,
                     "unit": "GB/s", "frac": bytes_rank / t_step * 1e-9 / peak,
                     "frac_of_8TBs_nominal": bytes_rank / t_step * 1e-9 / 8000.0,
                     "algorithmic_bytes_per_launch": bytes_rank,
                     "basis": "per GPU: reference-counted bytes of one call / ranks / step time"
                              + ("" if world == 1 else " (step = halo pack [+ exchange] + interior kernel + halo combine)"),
                     "traffic": measured_traffic("bs8_gather_scatter_64_p7") if (world == 1 and nx == 64 and p == 7) else None},
        "e2e": {"value": Ng / t_e2e * 1e-9, "unit": "GDOF/s", "ms_per_step": t_e2e * 1e3, "steps": e2e_steps,
                "h2d_bytes_per_step": nbytes * world, "d2h_bytes_per_step": nbytes * world,
                "api": "ogs_t::GatherScatterStart/Finish(pinned host q): H2D -> GatherScatter -> D2H per step, two vectors in flight",
                "copy_ceiling": {"ms_per_step": t_copy * 1e3, "GDOF/s": Ng / t_copy * 1e-9,
                                 "gbs_per_gpu_per_direction": nbytes / t_copy * 1e-9,
                                 "what": "the same H2D + D2H of one step on two streams with no kernel, max over ranks (host wall clock)",
                                 "e2e_fraction_of_ceiling": t_copy / t_e2e}},
        "gpu_launches": args.steps * launches_per_step,
        "clocks": clocks,
        "parity_checked": parity["ok"], "parity": parity,
    }
    if world > 1:
        ok = torch.tensor([1 if parity["ok"] else 0], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        line["parity_checked"] = bool(ok.item())
    if not line["parity_checked"]:
        raise SystemExit(f"bench.py: parity check failed on rank {rank}: {parity}")

    if args.strong and world > 1:
        line["strong"] = strong_scaling(torch, dist, host, comm, platform, rank, world, p, timed)
    if args.tail:
        line["tail"] = latency_tail(torch, dist, host, K, comm, platform, rank, world, p, timed)

    if args.suite and world == 1:
        line["suite"] = run_suite(torch, K, peak, (ogs, Nt, Ng, q))
        line["suite"]["BS8 gatherScatter"] = {"ms": t_step * 1e3, "GB/s": line["effective_gbs"],
                                              "frac_of_measured_peak": line["roofline"]["frac"],
                                              "frac_of_8TBs": line["roofline"]["frac_of_8TBs_nominal"],
                                              "GDOF/s": line["value"], "bytes": bytes_step}

    if args.cpu_baseline and rank == 0 and world == 1:
        line["cpu_baseline"] = cpu_baseline(args)
    if rank == 0:
        print(json.dumps(line), flush=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ogs.close()
    platform.close()
    comm.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def cpu_baseline(args):
    """The reference's CPU path on this box's host cores, bounded sample (rank 0, N=1): about 10 s per variant on a
    32^3 box of the same degree (the full 64^3 run is what `--impl reference` times)."""
    nx = args.cpu_nx
    results, Nt, Ng = reference_cpu(nx, args.p, 0, 0, budget_s=8.0)
    best = min(results, key=lambda k: results[k]["t"])
    v = results[best]
    return {"value": Ng / v["t"] * 1e-9, "unit": "GDOF/s", "cores": v["cores"], "kind": v["kind"],
            "sample": f"{best}: BS8 gatherScatter, {nx}^3 degree-{args.p} hex box (Nt={Nt}, Ng={Ng}), ~8 s of calls per variant",
            "all": {k: {"GDOF/s": Ng / r["t"] * 1e-9, "GB/s": bs8_bytes(Nt, Ng) / r["t"] * 1e-9, "cores": r["cores"], "kind": r["kind"]}
                    for k, r in results.items()}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nx", type=int, default=64, help="elements per direction per rank (reference -nx/-ny/-nz)")
    ap.add_argument("--p", type=int, default=7, help="polynomial degree (reference -p)")
    ap.add_argument("--ref-nx", type=int, default=0, help="--impl reference: elements per direction (0 = --nx, the same workload)")
    ap.add_argument("--cpu-nx", type=int, default=32, help="cpu_baseline: elements per direction of the bounded CPU sample")
    ap.add_argument("--no-tail", dest="tail", action="store_false")
    ap.add_argument("--no-strong", dest="strong", action="store_false")
    ap.add_argument("--no-suite", dest="suite", action="store_false")
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
