"""Seeded inputs of the golden cases (shared by make_golden.py and the tests)."""
import numpy as np

import spn_testutil as util

DTYPES = {"f64": np.float64, "f32": np.float32, "i32": np.int32, "i64": np.int64}


def stream_cases():
    return [dict(n=n, alpha=alpha, beta=beta, bs=bs)
            for (n, alpha, beta, bs) in [(1, 1.0, 1.0, 256), (257, 1.0, 1.0, 256), (4099, 0.37, -1.25, 256), (9001, 0.37, 1.0, 1024)]]


def stream_key(c):
    return f"stream/n{c['n']}_a{c['alpha']}_b{c['beta']}_bs{c['bs']}"


def stream_inputs(c):
    return util.splitmix64(c["n"], 101), util.splitmix64(c["n"], 102)


def ogs_cases():
    out = []
    for mesh in [(2, 2, 2, 3), (3, 2, 2, 1)]:
        for dt, K, op in [("f64", 1, 0), ("f64", 3, 0), ("f64", 1, 1), ("f64", 1, 2), ("f64", 1, 3), ("f32", 1, 0), ("i32", 1, 2),
                          ("i32", 2, 0), ("i64", 1, 3), ("i64", 1, 0)]:
            out.append(dict(mesh=mesh, dtype=dt, K=K, op=op))
    # degree 7 (the BASELINE degree: multiplicities 1, 2, 4, 8) and degree 15 (the largest the CLI accepts)
    out.append(dict(mesh=(2, 2, 2, 7), dtype="f64", K=1, op=0))
    out.append(dict(mesh=(1, 2, 1, 15), dtype="f64", K=1, op=0))
    return out


def ogs_key(c):
    return "ogs/m" + "x".join(map(str, c["mesh"])) + f"_{c['dtype']}_K{c['K']}_op{c['op']}"


def ogs_values(n, dtype, seed):
    dtype = np.dtype(dtype)
    if np.issubdtype(dtype, np.floating):
        return util.splitmix64(n, seed).astype(dtype)
    return ((np.arange(n, dtype=np.int64) * 7919 + seed) % 23 - 11).astype(dtype)


def ogs_inputs(c):
    ids = util.hex_ids(*c["mesh"])
    rs, ci = util.csr_from_ids(ids)
    dtype = DTYPES[c["dtype"]]
    v = ogs_values(ids.size * c["K"], dtype, 200 + sum(c["mesh"]))
    if c["op"] == 1 and not np.issubdtype(np.dtype(dtype), np.floating):
        v = np.where(v == 0, 1, np.sign(v)).astype(dtype)
    return rs, ci, v


def extract_ids(rs, ci, v, c):
    n = v.size // c["K"]
    return ((np.arange(1 + n // 3) * 31) % n).astype(np.int32)
