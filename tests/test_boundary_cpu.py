"""CPU-only checks of the drop-in boundary: library loads and exports the C ABI, the module has the
reference's state_dict surface and initialiser, the native graph builder reproduces the reference's
hierarchy, the product path refuses to run without a GPU, and nothing in the product imports oracle/."""
import ast
import os
import re

import numpy as np
import pytest
import torch

from helpers import CASES, graph_from_fixture, load_npz, tensor_digest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_header_symbols():
    from pose2mesh_release_b200 import _lib

    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "p2m_b200.h")).read()
    declared = set(re.findall(r"\b(p2m_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/p2m_b200.h but not exported"
    assert set(_lib.EXPORTS) <= declared
    assert b"sm_100a" in lib.p2m_version()


def test_conv_kernel_register_split_is_balanced_in_the_build():
    """k_cheb_conv_umma redistributes registers between its warpgroups with setmaxnreg (csrc/cheb_umma.cu: REGS_*): the
    epilogue's setmaxnreg.inc draws on exactly what the utility warpgroup's setmaxnreg.dec released, which only works
    out if every instantiation is LAUNCHED with 80 registers per thread.  The host refuses to launch otherwise
    (check_launch_regs); this catches such a build here, without a GPU."""
    import shutil
    import subprocess

    from pose2mesh_release_b200 import build

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    lib = build.build()
    out = subprocess.run([cuobjdump, "-res-usage", lib], capture_output=True, text=True).stdout
    regs = re.findall(r"Function [^\n]*k_cheb_conv_umma[^\n]*\n[^\n]*REG:(\d+)", out)
    assert len(regs) >= 10, f"conv kernel instantiations not found in {lib}"
    assert set(regs) == {"80"}, f"launch register counts of k_cheb_conv_umma: {sorted(set(regs))}"
    src = open(os.path.join(ROOT, "pose2mesh_release_b200", "csrc", "cheb_umma.cu")).read()
    m = re.search(r"REGS_LAUNCH = (\d+), REGS_UTIL = (\d+), REGS_EPI = (\d+)", src)
    launch, util, epi = (int(v) for v in m.groups())
    assert launch == 80 and epi - launch <= launch - util and util % 8 == 0 and epi % 8 == 0


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "pose2mesh_release_b200")
    for fn in os.listdir(pkg):
        if not fn.endswith(".py"):
            continue
        tree = ast.parse(open(os.path.join(pkg, fn)).read())
        for node in ast.walk(tree):
            names = []
            if isinstance(node, ast.Import):
                names = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                names = [node.module or ""]
            assert not any(n.split(".")[0] == "oracle" for n in names), f"{fn} imports oracle"


@pytest.mark.parametrize("name", ["smpl_small", "mano_like"])
def test_module_state_dict_surface_and_init(name):
    from pose2mesh_release_b200.meshnet import Pose2Mesh

    n, seed, levels, mano = CASES[name]
    z = load_npz(f"meshnet_{name}.npz")
    mats, _ = graph_from_fixture(name)
    n_before = len(mats)
    torch.manual_seed(123)
    model = Pose2Mesh(5, 3, mats, joint_set="mano" if mano else "human36")
    assert len(mats) == n_before, "caller's list must not be mutated"
    sd = model.state_dict()
    assert sorted(sd.keys()) == [str(k) for k in z["keys"]]
    assert sum(p.numel() for p in model.parameters()) == int(z["n_param"])
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(z["shape/" + k]), k
        np.testing.assert_allclose(tensor_digest(v), z["init/" + k], rtol=1e-12, atol=0, err_msg=k)
    last = len(model.cl) - 1
    assert model.bn[last] is None and f"bn.{last}.weight" not in sd
    # reference-style checkpoints load (keys identical), incl. the DataParallel 'module.'-stripped form
    model.load_state_dict({k: v.clone() for k, v in sd.items()})


def test_forward_refuses_cpu_tensors():
    from pose2mesh_release_b200.meshnet import Pose2Mesh

    mats, _ = graph_from_fixture("mano_like")
    model = Pose2Mesh(5, 3, mats, joint_set="mano")
    with pytest.raises(RuntimeError, match="CUDA"):
        model(torch.zeros(1, 21, 5))


@pytest.mark.parametrize("name", ["smpl_small", "mano_like", "smpl_like"])
def test_native_graph_builder_matches_reference_fixture(name):
    import hashlib

    from pose2mesh_release_b200 import graph as pg

    def sha(a):
        return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()

    n, seed, levels, mano = CASES[name]
    z = load_npz(f"graph_{name}.npz")
    face = pg.synthetic_sphere_faces(n, seed)
    j, sk, fp = (21, pg.MANO_SKELETON, pg.MANO_HORI_CONN) if mano else (17, pg.H36M_SKELETON, pg.H36M_FLIP_PAIRS)
    adj, lap, perm, perm_rev = pg.build_coarse_graphs(face, j, sk, fp, levels=levels)
    assert np.array_equal(np.asarray(perm_rev), z["perm_reverse"])
    for i, m in enumerate(lap):
        c = m.tocsr()
        c.sort_indices()
        assert c.nnz == int(z[f"L{i}_nnz"])
        assert sha(c.indptr.astype(np.int64)) == str(z[f"L{i}_indptr_sha"])
        assert sha(c.indices.astype(np.int64)) == str(z[f"L{i}_indices_sha"])
        ref = z[f"L{i}_data32_sum"]
        assert abs(np.abs(c.data.astype(np.float32)).astype(np.float64).sum() - ref[1]) <= 1e-6 * ref[1]


def test_compute_perm_known_answer_native():
    from pose2mesh_release_b200 import graph as pg

    got = pg.compute_perm([np.array([4, 1, 1, 2, 2, 3, 0, 0, 3]), np.array([2, 1, 0, 1, 0])])
    assert got == [[3, 4, 0, 9, 1, 2, 5, 8, 6, 7, 10, 11], [2, 4, 1, 3, 0, 5], [0, 1, 2]]   # lib/coarsening.py:261-262


def test_model_create_without_gpu_fails_loudly():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pose2mesh_release_b200.meshnet import Pose2Mesh

    mats, _ = graph_from_fixture("mano_like")
    model = Pose2Mesh(5, 3, mats, joint_set="mano")
    with pytest.raises(RuntimeError, match="no usable CUDA device|no CPU path"):
        model._hier.handle(0)
