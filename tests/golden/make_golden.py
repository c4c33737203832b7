"""Generates tests/golden/ref_okl_golden.npz: outputs of the REFERENCE's own OKL kernels (compiled for the host by
oracle/Makefile into oracle/_ref/libref_okl.so, from the sources under /root/reference) on seeded inputs.

Run in the build container (the only place /root/reference exists):
    make -C oracle && python tests/golden/make_golden.py
Only OUTPUTS are stored; tests regenerate the inputs from the same seeds (tests/golden/cases.py), so the fixture stays
small.  The reference ships no golden vectors for this path: these are the pin that travels."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, HERE)

import cases  # noqa: E402
from oracle import pyoracle  # noqa: E402


def main():
    assert pyoracle.have_ref(), "oracle/_ref/libref_okl.so is missing: run `make -C oracle` where /root/reference exists"
    out = {}
    for c in cases.stream_cases():
        a, b = cases.stream_inputs(c)
        key = cases.stream_key(c)
        out[key + "/bs1"] = pyoracle.ref_bs1(a)
        out[key + "/bs2"] = pyoracle.ref_bs2(c["alpha"], a, c["beta"], b)
        out[key + "/bs3"] = np.float64(pyoracle.ref_norm2(a))
        out[key + "/bs4"] = np.float64(pyoracle.ref_dot(a, b))
        r, d = pyoracle.ref_cg_update(a, c["alpha"], b, bs=c["bs"])
        out[key + "/bs5_r"] = r
        out[key + "/bs5_rdotr"] = np.float64(d)
    for c in cases.ogs_cases():
        rs, ci, v = cases.ogs_inputs(c)
        bs = pyoracle.block_rows(rs, 512)
        key = cases.ogs_key(c)
        g = pyoracle.ref_ogs_gather(bs, rs, ci, v, K=c["K"], op=c["op"])
        out[key + "/gather"] = g
        out[key + "/gatherScatter"] = pyoracle.ref_ogs_gather_scatter(bs, rs, ci, rs, ci, v, K=c["K"], op=c["op"])
        if c["op"] == 0:
            out[key + "/scatter"] = pyoracle.ref_ogs_scatter(bs, rs, ci, g, np.zeros_like(v), K=c["K"])
            out[key + "/extract"] = pyoracle.ref_ogs_extract(cases.extract_ids(rs, ci, v, c), v, K=c["K"])
    path = os.path.join(HERE, "ref_okl_golden.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
