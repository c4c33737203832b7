"""Label sets behind tests/golden/ref_ogs_maps.npz (shared by the generator and the tests)."""
import numpy as np

import spn_testutil as util

SIGNED, UNSIGNED = 1, 0

# name -> (builder, kind, unique)
CASES = ("hex_p1_nx2", "hex_p1_nx4", "hex_p3_nx2", "hex_p3_nx4", "hex_p7_nx2", "hex_p7_nx4", "hex_p15_nx2",
         "hex_p7_masked", "hex_p3_unsigned", "hex_p3_presigned", "random_labels", "long_rows",
         "refmesh_tri_p4", "refmesh_quad_p3", "refmesh_tet_p3")

# labels of the reference's OWN box meshes of the other element types (-e Tri / Quad / Tet of BS6-8:
# libs/mesh/meshSetupBox{Tri2D,Quad2D,Tet3D}.cpp + meshConnectNodes.cpp); element, nx, ny, nz, degree.
# The generator takes them from oracle/_ref/libref_ogs.so and stores them in the fixture as <name>/ids_in.
REF_MESHES = {"refmesh_tri_p4": ("Tri", 4, 3, 1, 4), "refmesh_quad_p3": ("Quad", 4, 3, 1, 3), "refmesh_tet_p3": ("Tet", 3, 2, 2, 3)}
_fixture = None


def case_ids(name):
    if name in REF_MESHES:  # up to 24 copies of a node (a tet-mesh vertex), rows in the reference's element order
        global _fixture
        if _fixture is None:
            import os
            _fixture = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_ogs_maps.npz"))
        return _fixture[f"{name}/ids_in"].copy(), SIGNED, True
    if name.startswith("hex_p") and "_nx" in name:
        p, nx = int(name.split("_")[1][1:]), int(name.split("_")[2][2:])
        return util.hex_ids(nx, nx, nx, p), SIGNED, True
    if name == "hex_p7_masked":  # Dirichlet-style mask: every 5th node does not participate
        ids = util.hex_ids(3, 2, 2, 7).copy()
        ids[::5] = 0
        return ids, SIGNED, True
    if name == "hex_p3_unsigned":
        return util.hex_ids(3, 3, 2, 3), UNSIGNED, False
    if name == "hex_p3_presigned":  # caller-chosen positive copy (the first), unique = false
        ids = util.hex_ids(3, 2, 3, 3).copy()
        seen = set()
        for i, v in enumerate(ids):
            if v in seen:
                ids[i] = -v
            seen.add(v)
        return ids, SIGNED, False
    if name == "random_labels":  # no mesh structure: 5000 nodes over 1200 labels, some zeros
        rng = np.random.RandomState(7)
        ids = rng.randint(0, 1200, size=5000).astype(np.int64)
        return ids, SIGNED, True
    if name == "long_rows":  # rows longer than a warp (41 and 301 copies; the reference refuses rows above 512,
        # libs/ogs/ogsOperator.cpp:535-536) next to short ones
        ids = np.concatenate([np.full(40, 5), np.arange(6, 200), np.full(300, 3), np.arange(6, 120), [5, 3]]).astype(np.int64)
        return ids, SIGNED, True
    raise KeyError(name)


def case_values(name, n):
    seed = sum(ord(c) for c in name)
    return util.splitmix64(n, seed)
