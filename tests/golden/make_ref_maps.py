"""Writes tests/golden/ref_ogs_maps.npz: the index maps and operator results of the reference's OWN ogs_t::Setup
(libs/ogs/ogsSetup.cpp:87-290, libs/ogs/ogsOperator.cpp:522-665) and box-hex mesh ids (libs/mesh), produced by
oracle/_ref/libref_ogs.so — the reference sources compiled unchanged (oracle/Makefile `ref-host`).
Run in the container that holds /root/reference:   python tests/golden/make_ref_maps.py
The fixture lets the pin hold where neither /root/reference nor oracle/_ref exists."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from golden.ref_map_cases import CASES, REF_MESHES, case_ids, case_values  # noqa: E402
from oracle import refogs  # noqa: E402


def main():
    out = {}
    for name in CASES:
        if name in REF_MESHES:  # labels of the reference's own Tri / Quad / Tet box mesh, signs dropped
            ids, kind, unique = np.abs(refogs.mesh_box_ids(*REF_MESHES[name])), refogs.SIGNED, True
            out[f"{name}/ids_in"] = ids.copy()
        else:
            ids, kind, unique = case_ids(name)
        r = refogs.RefOgs(ids, kind=kind, unique=unique)
        for k, v in r.maps().items():
            out[f"{name}/{k}"] = v
        out[f"{name}/ids_out"] = r.ids
        out[f"{name}/dims"] = np.array([r.N, r.Ngather, r.NlocalT, r.NlocalP, r.NhaloT, r.NhaloP, r.nnzT, r.nnzN], dtype=np.int64)
        v = case_values(name, ids.size)
        small = ids.size <= 8192  # keeps the committed fixture near 2 MB: the large cases store Sym only
        for tname, trans in (("sym", refogs.SYM), ("notrans", refogs.NOTRANS), ("trans", refogs.TRANS)):
            host_result = r.gather_scatter(v, 1, refogs.ADD, trans)
            # the reference's device path (kernel_t -> compiled OKL kernels) must agree with its host twin
            dev = r.gather_scatter(v, 1, refogs.ADD, trans, device=True)
            assert np.array_equal(dev.view(np.int64), host_result.view(np.int64)), (name, tname)
            if small or tname == "sym":
                out[f"{name}/gs_{tname}"] = host_result
        if kind != refogs.UNSIGNED and small:  # Gather / Scatter need a unique positive copy (libs/ogs/ogsSetup.cpp:827-830)
            out[f"{name}/gather_trans"] = r.gather(v, 1, refogs.ADD, refogs.TRANS)[: r.Ngather]
            gq = case_values(name + "g", r.NlocalT + r.NhaloT)
            out[f"{name}/scatter_notrans"] = r.scatter(gq, np.zeros(ids.size), 1, refogs.NOTRANS)
        r.close()
    for (nx, ny, nz, p) in ((2, 2, 2, 1), (3, 2, 2, 3), (2, 3, 4, 7)):
        out[f"mesh_{nx}_{ny}_{nz}_{p}"] = refogs.mesh_box_hex_ids(nx, ny, nz, p)
    path = os.path.join(ROOT, "tests", "golden", "ref_ogs_maps.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
