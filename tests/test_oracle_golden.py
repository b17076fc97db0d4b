"""Pin the CPU oracle (oracle/) against the reference: the reference's own compute_perm KAT
(lib/coarsening.py:261-262) and the fixtures produced by the unmodified reference
(tests/golden/make_golden.py)."""
import hashlib

import numpy as np
import pytest
import torch

from oracle import graph_oracle as go
from oracle import meshnet_oracle as mo

from helpers import CASES, graph_from_fixture, load_npz, rel_err, tensor_digest


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_compute_perm_known_answer():
    got = go.binary_tree_order([np.array([4, 1, 1, 2, 2, 3, 0, 0, 3]), np.array([2, 1, 0, 1, 0])])
    assert got == [[3, 4, 0, 9, 1, 2, 5, 8, 6, 7, 10, 11], [2, 4, 1, 3, 0, 5], [0, 1, 2]]


def _joint(mano):
    return (21, go.MANO_SKELETON, go.MANO_HORI_CONN) if mano else (17, go.H36M_SKELETON, go.H36M_FLIP_PAIRS)


@pytest.mark.parametrize("name", ["smpl_small", "mano_like", "smpl_like"])
def test_graph_oracle_matches_reference(name):
    n, seed, levels, mano = CASES[name]
    z = load_npz(f"graph_{name}.npz")
    face = go.synthetic_sphere_faces(n, seed)
    if "face" in z:
        assert np.array_equal(face, z["face"])
    j, sk, fp = _joint(mano)
    adj, lap, perm, perm_rev = go.build_coarse_graphs(face, j, sk, fp, levels=levels)
    assert np.array_equal(np.asarray(perm_rev), z["perm_reverse"])
    assert len(lap) == int(z["n_levels"])
    for i, m in enumerate(lap):
        c = m.tocsr()
        c.sort_indices()
        assert tuple(c.shape) == tuple(z[f"L{i}_shape"])
        assert c.nnz == int(z[f"L{i}_nnz"])
        assert sha(c.indptr.astype(np.int64)) == str(z[f"L{i}_indptr_sha"])
        assert sha(c.indices.astype(np.int64)) == str(z[f"L{i}_indices_sha"])
        d32 = c.data.astype(np.float32)
        ref = z[f"L{i}_data32_sum"]
        assert abs(d32.astype(np.float64).sum() - ref[0]) <= 1e-6 * max(1.0, abs(ref[0]))
        assert abs(np.abs(d32).astype(np.float64).sum() - ref[1]) <= 1e-6 * ref[1]
        if f"L{i}_data" in z:
            # ARPACK's start vector is not seeded, so lmax may move in the last f64 bits
            np.testing.assert_allclose(c.data, z[f"L{i}_data"], rtol=1e-12, atol=1e-14)


def test_smpl_like_level_sizes_match_real_smpl():
    z = load_npz("graph_smpl_like.npz")
    sizes = [int(z[f"L{i}_shape"][0]) for i in range(int(z["n_levels"]))]
    assert sizes == [12288, 6144, 3072, 1536, 768, 384, 192, 96, 48, 17]   # meshnet.py:27,35-37,117 comments
    assert int(z["L0_nnz"]) == 6890 + 2 * 20664 + (12288 - 6890)


def test_cheb_conv_oracle_matches_reference():
    z = load_npz("cheb_conv.npz")
    mats, _ = graph_from_fixture("smpl_small")
    lap = mo.laplacians_to_torch(mats, drop_second_coarsest=False)[int(z["level"])]
    x = torch.from_numpy(z["x"])
    w, b = torch.from_numpy(z["weight"]), torch.from_numpy(z["bias"])
    y = mo.cheb_conv(x, lap, w, b)
    assert rel_err(y, torch.from_numpy(z["y_plain"])) < 2e-6
    fout = w.shape[0]
    bn = dict(weight=torch.from_numpy(z["bn_weight"]), bias=torch.from_numpy(z["bn_bias"]),
              running_mean=torch.zeros(fout), running_var=torch.ones(fout))
    y = mo.cheb_conv(x, lap, w, b, bn, training=True)
    assert rel_err(y, torch.from_numpy(z["y_bn_train"])) < 2e-6
    np.testing.assert_allclose(bn["running_mean"].numpy(), z["bn_running_mean"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(bn["running_var"].numpy(), z["bn_running_var"], rtol=1e-5, atol=1e-7)
    y = mo.cheb_conv(x, lap, w, b, bn, training=False)
    assert rel_err(y, torch.from_numpy(z["y_bn_eval"])) < 2e-6


@pytest.mark.parametrize("name", ["smpl_small", "mano_like"])
def test_meshnet_oracle_matches_reference(name):
    n, seed, levels, mano = CASES[name]
    z = load_npz(f"meshnet_{name}.npz")
    mats, _ = graph_from_fixture(name)
    laps = mo.laplacians_to_torch(mats)
    sizes = [m.shape[0] for m in laps]
    torch.manual_seed(123)
    sd0 = mo.init_state_dict(5, 3, sizes, mano)
    assert sorted(sd0.keys()) == [str(k) for k in z["keys"]]
    assert sum(v.numel() for k, v in sd0.items() if "running" not in k and "num_batches" not in k) == int(z["n_param"])
    for k, v in sd0.items():
        assert tuple(v.shape) == tuple(z["shape/" + k]), k
        np.testing.assert_allclose(tensor_digest(v), z["init/" + k], rtol=1e-12, atol=0, err_msg=k)

    x = torch.from_numpy(z["x"])
    sd_eval = mo.randomize_bn_({k: v.clone() for k, v in sd0.items()}, seed=7)
    with torch.no_grad():
        y = mo.forward(sd_eval, laps, x, mano=mano, training=False)
    assert rel_err(y, torch.from_numpy(z["y_eval"])) < 1e-5

    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone())
          for k, v in sd0.items()}
    xg = x.clone().requires_grad_(True)
    y = mo.forward(sd, laps, xg, mano=mano, training=True)
    assert rel_err(y, torch.from_numpy(z["y_train"])) < 1e-5
    loss = (y - torch.from_numpy(z["target"])).abs().mean()
    assert abs(loss.item() - float(z["loss"])) < 1e-5
    loss.backward()
    assert rel_err(xg.grad, torch.from_numpy(z["dx"])) < 1e-4
    for k, v in sd.items():
        if v.requires_grad:
            got, ref = tensor_digest(v.grad), z["grad/" + k]
            assert abs(got[1] - ref[1]) <= 1e-3 * ref[1] + 1e-6, k      # sum |g| (bias grads under BN are pure rounding noise)
            assert abs(got[2] - ref[2]) <= 2e-3 * ref[2] + 1e-12, k     # sum g^2
        if "running" in k:
            np.testing.assert_allclose(v.numpy(), z["after/" + k], rtol=1e-4, atol=1e-6, err_msg=k)

    # the open-ReLU train step of the reference (strict/*: every BatchNorm bias = +6): the fixture of the tight GPU digests
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone())
          for k, v in sd0.items()}
    with torch.no_grad():
        for k in sd:
            if k.startswith("bn.") and k.endswith(".bias"):
                sd[k].fill_(6.0)
    xs = x.clone().requires_grad_(True)
    ys = mo.forward(sd, laps, xs, mano=mano, training=True)
    assert rel_err(ys, torch.from_numpy(z["strict/y_train"])) < 1e-5
    ls = (ys - torch.from_numpy(z["target"])).abs().mean()
    assert abs(ls.item() - float(z["strict/loss"])) < 1e-5
    ls.backward()
    assert rel_err(xs.grad, torch.from_numpy(z["strict/dx"])) < 1e-4
    for k, v in sd.items():
        if v.requires_grad:
            got, ref = tensor_digest(v.grad), z["strict/grad/" + k]
            assert abs(got[1] - ref[1]) <= 1e-3 * ref[1] + 1e-5, k
            assert abs(got[2] - ref[2]) <= 2e-3 * ref[2] + 1e-10, k


@pytest.mark.parametrize("name", ["smpl_small", "smpl_like"])
def test_padding_vertices_are_isolated_and_reduce_to_a_dense_map(name):
    """Structure the next kernel round builds on: the fake vertices that the binary-tree
    reorder pads each level with (lib/coarsening.py:214-258) are isolated in L~ and all carry the same diagonal
    value c = 2/lmax - 1 (laplacian(): row = 1 on the diagonal, rescale_L: L/(lmax/2) - I), so on those rows the
    Chebyshev conv is the dense map y = x (W0 + c W1 + (2c^2 - 1) W2)^T + b; and no real row ever reads a fake one."""
    n, seed, levels, mano = CASES[name]
    if name == "smpl_like":  # the fixture of the full-size case only stores digests: rebuild it with the oracle
        j, sk, fp = _joint(mano)
        _, mats, _, _ = go.build_coarse_graphs(go.synthetic_sphere_faces(n, seed), j, sk, fp, levels=levels)
    else:
        mats, _ = graph_from_fixture(name)
    frac = []
    for L in mats[: min(3, len(mats) - 1)]:
        c = L.tocsr().astype(np.float64)
        off = c.copy()
        off.setdiag(0)
        off.eliminate_zeros()
        iso = np.diff(off.indptr) == 0
        frac.append(iso.mean())
        diag = c.diagonal()
        assert iso.any() and np.ptp(diag[iso]) < 1e-12
        # no edge from a connected row into an isolated one (the matrix is symmetric, so also none out of it)
        assert not iso[off.indices].any()
        cval = float(diag[iso][0])
        g = torch.Generator().manual_seed(3)
        fin, fout = 8, 5
        x = torch.randn(2, c.shape[0], fin, generator=g)
        w = torch.randn(fout, fin * 3, generator=g) * 0.2
        b = torch.randn(fout, generator=g)
        lap = mo.laplacians_to_torch([L], drop_second_coarsest=False)[0]
        y = mo.cheb_conv(x, lap, w, b)
        w3 = w.view(fout, fin, 3)
        w_eff = w3[:, :, 0] + cval * w3[:, :, 1] + (2 * cval * cval - 1) * w3[:, :, 2]
        y_iso = x[:, iso] @ w_eff.t() + b
        assert rel_err(y[:, iso], y_iso) < 1e-5
    if name == "smpl_like":
        assert 0.40 < frac[0] < 0.48  # 5398 of 12288 rows at the finest SMPL-size level


def test_demo_oracle_matches_reference_golden():
    """demo/run.py:150-158 normalisation, PoseNet (lib/models/posenet.py) and the FlatPose2Mesh concat
    (pose2mesh_net.py:18-19), restated in oracle/demo_oracle.py, against the outputs of the unmodified reference
    functions (tests/golden/demo_pipeline.npz, made by tests/golden/make_golden_demo.py)."""
    from oracle import demo_oracle as do

    z = load_npz("demo_pipeline.npz")
    assert np.array_equal(z["joint_input"], load_npz("demo_input.npz")["joint_input"])
    bbox2 = do.process_bbox(do.get_bbox(z["joint_input"]).copy())
    np.testing.assert_allclose(bbox2, z["bbox2"], rtol=1e-7)
    np.testing.assert_allclose(do.normalize_pose2d(z["joint_input"]), z["joint_img"][0], atol=1e-6)
    torch.manual_seed(123)
    sd = do.posenet_init_state_dict(17)
    mo.randomize_bn_({("bn." + k): v for k, v in sd.items() if "batch_norm" in k}, seed=11)
    with torch.no_grad():
        pose3d = do.posenet_forward(sd, torch.from_numpy(z["pose2d"]).reshape(8, -1))
    np.testing.assert_allclose(pose3d.numpy(), z["pose3d"], rtol=1e-5, atol=1e-6)
    comb = do.flat_pose2mesh_input(torch.from_numpy(z["pose2d"]), pose3d)
    np.testing.assert_allclose(comb.numpy(), z["pose_combine"], rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("name", ["smpl_small", "mano_like"])
def test_children_of_padding_vertices_are_padding_vertices(name):
    """What the eval-mode duplicate elimination among the isolated rows rests on (DevLevel::rep_tiles): in the
    reference's binary-tree reorder (lib/coarsening.py:214-258) fake vertices are added bottom-up, so both children
    (rows 2p, 2p+1 of the next finer level) of an isolated row p are isolated too, with the level's common diagonal."""
    mats, _ = graph_from_fixture(name)
    mesh_levels = mats[:-1]                       # the joint graph is last

    def isolated(m):
        c = m.tocsr()
        deg = np.diff(c.indptr)
        return (deg == 1) & (c.indices[np.minimum(c.indptr[:-1], c.nnz - 1)] == np.arange(c.shape[0]))

    iso = [isolated(m) for m in mesh_levels]
    checked = 0
    for k in range(len(iso) - 1):
        if mesh_levels[k].shape[0] != 2 * mesh_levels[k + 1].shape[0]:
            continue
        parents = np.nonzero(iso[k + 1])[0]
        assert iso[k][2 * parents].all() and iso[k][2 * parents + 1].all(), k
        checked += len(parents)
    assert checked > 0


def test_loss_oracle_matches_reference_golden():
    """lib/core/loss.py:10-23,62-114 restated in oracle/loss_oracle.py vs the unmodified reference classes
    (tests/golden/mesh_losses.npz): the three losses and the gradient of their weighted sum."""
    from oracle import loss_oracle as lo

    z = load_npz("mesh_losses.npz")
    out = torch.from_numpy(z["out"]).requires_grad_(True)
    gt, valid, face = torch.from_numpy(z["gt"]), torch.from_numpy(z["valid"]), z["face"].astype(np.int64)
    ln, le, lc = lo.normal_vector_loss(out, gt, face), lo.edge_length_loss(out, gt, face), lo.coord_loss(out, gt, valid)
    assert abs(ln.item() - float(z["normal"])) < 1e-6
    assert abs(le.item() - float(z["edge"])) < 1e-6
    assert abs(lc.item() - float(z["coord"])) < 1e-6
    w = z["weights"]
    (w[0] * ln + w[1] * le + w[2] * lc).backward()
    np.testing.assert_allclose(out.grad.numpy(), z["grad"], rtol=1e-4, atol=1e-8)
