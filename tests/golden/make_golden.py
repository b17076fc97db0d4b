#!/usr/bin/env python
"""Generate the golden fixtures in this directory by running the UNMODIFIED reference
(/root/reference, hongsukchoi/Pose2Mesh_RELEASE) in-process through oracle/ref_shim.py.

Run once in the build container (the reference tree does not exist on the GPU box):

    python tests/golden/make_golden.py

Outputs (all small, committed):
    graph_<case>.npz     build_coarse_graphs() outputs of the reference for a seeded synthetic mesh
                         (full CSR for the small cases, digests for the 6890-vertex case)
    meshnet_<case>.npz   Pose2Mesh forward (eval + train) and backward of the reference for seeded
                         weights/inputs: outputs, BN running stats, input gradient, per-parameter
                         gradient digests, per-parameter init digests (torch.manual_seed(123)); and the
                         train step once more with the ReLUs held open (strict/*: BatchNorm bias +6)
    cheb_conv.npz        one graph_conv_cheby() call of the reference (with and without BatchNorm)
    demo_input.npz       the reference's only hot-path input fixture demo/h36m_joint_input.npy
"""
import contextlib
import hashlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import graph_oracle as go  # noqa: E402
from oracle import meshnet_oracle as mo  # noqa: E402
from oracle import ref_shim  # noqa: E402

CASES = {
    # name: (n_vertex, mesh seed, levels, joint set)
    "smpl_small": (1200, 0, 9, "human36"),
    "mano_like": (778, 1, 6, "mano"),
    "smpl_like": (6890, 2, 9, "human36"),
}


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def tensor_digest(t: torch.Tensor) -> np.ndarray:
    d = t.detach().double().flatten()
    return np.array([d.sum().item(), d.abs().sum().item(), (d * d).sum().item()] + d[:5].tolist())


def joint_setting(joint_set):
    if joint_set == "mano":
        return 21, go.MANO_SKELETON, go.MANO_HORI_CONN
    return 17, go.H36M_SKELETON, go.H36M_FLIP_PAIRS


def ref_graphs(gu, name):
    n, seed, levels, joint_set = CASES[name]
    face = go.synthetic_sphere_faces(n, seed)
    j, sk, fp = joint_setting(joint_set)
    with contextlib.redirect_stdout(io.StringIO()):
        adj, lap, perm, perm_rev = gu.build_coarse_graphs(face, j, sk, fp, levels=levels)
    return face, adj, lap, perm, perm_rev


def save_graph(gu, name):
    face, adj, lap, perm, perm_rev = ref_graphs(gu, name)
    out = {"n_levels": len(lap), "perm_reverse": np.asarray(perm_rev, dtype=np.int64)}
    full = name != "smpl_like"
    for i, m in enumerate(lap):
        c = m.tocsr()
        c.sort_indices()
        out[f"L{i}_shape"] = np.array(c.shape)
        out[f"L{i}_nnz"] = np.array(c.nnz)
        out[f"L{i}_indptr_sha"] = np.array(digest(c.indptr.astype(np.int64)))
        out[f"L{i}_indices_sha"] = np.array(digest(c.indices.astype(np.int64)))
        out[f"L{i}_data32_sum"] = np.array([c.data.astype(np.float32).astype(np.float64).sum(),
                                            np.abs(c.data.astype(np.float32)).astype(np.float64).sum()])
        if full:
            out[f"L{i}_indptr"] = c.indptr.astype(np.int32)
            out[f"L{i}_indices"] = c.indices.astype(np.int32)
            out[f"L{i}_data"] = c.data.astype(np.float64)
    if full:
        out["face"] = face.astype(np.int32)
    np.savez_compressed(os.path.join(HERE, f"graph_{name}.npz"), **out)
    print("wrote graph", name, [m.shape[0] for m in lap])


def save_meshnet(gu, mn, name, batch=2):
    n, seed, levels, joint_set = CASES[name]
    mano = joint_set == "mano"
    ref_shim.load(joint_set)  # sets cfg.DATASET.target_joint_set
    face, adj, lap, perm, perm_rev = ref_graphs(gu, name)
    j = joint_setting(joint_set)[0]
    lap_for_oracle = [m.copy() for m in lap]
    torch.manual_seed(123)  # main/train.py:12,22
    with contextlib.redirect_stdout(io.StringIO()):
        model = mn.get_model(5, 3, lap)
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    out = {"n_param": np.array(sum(p.numel() for p in model.parameters())),
           "keys": np.array(sorted(sd0.keys()))}
    for k, v in sd0.items():
        out["init/" + k] = tensor_digest(v)
        out["shape/" + k] = np.array(v.shape, dtype=np.int64)

    # the oracle's initialiser must reproduce the reference's RNG consumption exactly
    torch.manual_seed(123)
    sizes = [m.shape[0] for m in lap_for_oracle]
    del sizes[-2]
    sd_o = mo.init_state_dict(5, 3, sizes, mano)
    for k in sd0:
        assert torch.equal(sd0[k], sd_o[k]), k

    g = torch.Generator().manual_seed(0)
    x = torch.randn(batch, j, 5, generator=g)
    tgt = torch.randn(batch, lap_for_oracle[0].shape[0], 3, generator=g)
    out["x"] = x.numpy()
    out["target"] = tgt.numpy()

    # ---- eval forward with randomised BN -------------------------------------
    sd_eval = mo.randomize_bn_({k: v.clone() for k, v in sd0.items()}, seed=7)
    model.load_state_dict(sd_eval)
    model.eval()
    with ref_shim.cpu_cuda_noop(), torch.no_grad():
        y_eval = model(x)
    out["y_eval"] = y_eval.numpy()

    # ---- train forward + backward (L1 to random target) from the seeded init ---------------
    model.load_state_dict(sd0)
    model.train()
    xg = x.clone().requires_grad_(True)
    with ref_shim.cpu_cuda_noop():
        y_train = model(xg)
        loss = (y_train - tgt).abs().mean()
        loss.backward()
    out["y_train"] = y_train.detach().numpy()
    out["loss"] = np.array(loss.item())
    out["dx"] = xg.grad.numpy()
    for k, p in model.named_parameters():
        out["grad/" + k] = tensor_digest(p.grad)
    for k, v in model.state_dict().items():
        if "running" in k:
            out["after/" + k] = v.numpy().copy()
    # ---- the same with the ReLUs held open (every BatchNorm bias = +6): no activation sits near its kink, so two
    #      correct fp32 implementations agree on every gradient to rounding — the fixture for TIGHT gradient digests
    sd_open = {k: v.clone() for k, v in sd0.items()}
    for k in sd_open:
        if k.startswith("bn.") and k.endswith(".bias"):
            sd_open[k].fill_(6.0)
    model.load_state_dict(sd_open)
    model.train()
    model.zero_grad()
    xs = x.clone().requires_grad_(True)
    with ref_shim.cpu_cuda_noop():
        ys = model(xs)
        ls = (ys - tgt).abs().mean()
        ls.backward()
    out["strict/y_train"] = ys.detach().numpy()
    out["strict/loss"] = np.array(ls.item())
    out["strict/dx"] = xs.grad.numpy()
    for k, p in model.named_parameters():
        out["strict/grad/" + k] = tensor_digest(p.grad)
    np.savez_compressed(os.path.join(HERE, f"meshnet_{name}.npz"), **out)
    print("wrote meshnet", name, tuple(y_eval.shape), "loss", loss.item(), "strict loss", ls.item())


def save_cheb_conv(gu, cgc):
    face, adj, lap, perm, perm_rev = ref_graphs(gu, "smpl_small")
    level = 3  # V = 256
    L = gu.sparse_python_to_torch(lap[level])
    v = lap[level].shape[0]
    g = torch.Generator().manual_seed(11)
    b, fin, fout = 3, 20, 12
    x = torch.randn(b, v, fin, generator=g)
    cl = torch.nn.Linear(3 * fin, fout)
    cl.weight.data = torch.randn(fout, 3 * fin, generator=g) * 0.1
    cl.bias.data = torch.randn(fout, generator=g) * 0.1
    bn = torch.nn.BatchNorm1d(fout)
    bn.weight.data = torch.rand(fout, generator=g) + 0.5
    bn.bias.data = torch.randn(fout, generator=g) * 0.1
    bn.train()
    y_plain = cgc.graph_conv_cheby(x, cl, None, L, fout, 3)
    y_bn_train = cgc.graph_conv_cheby(x, cl, bn, L, fout, 3)
    bn.eval()
    y_bn_eval = cgc.graph_conv_cheby(x, cl, bn, L, fout, 3)
    np.savez_compressed(os.path.join(HERE, "cheb_conv.npz"), level=np.array(level), x=x.numpy(),
                        weight=cl.weight.detach().numpy(), bias=cl.bias.detach().numpy(),
                        bn_weight=bn.weight.detach().numpy(), bn_bias=bn.bias.detach().numpy(),
                        bn_running_mean=bn.running_mean.numpy(), bn_running_var=bn.running_var.numpy(),
                        y_plain=y_plain.detach().numpy(), y_bn_train=y_bn_train.detach().numpy(),
                        y_bn_eval=y_bn_eval.detach().numpy())
    print("wrote cheb_conv")


def save_demo_input():
    p = os.path.join(ref_shim.REF_ROOT, "demo", "h36m_joint_input.npy")
    a = np.load(p)
    np.savez_compressed(os.path.join(HERE, "demo_input.npz"), joint_input=a)
    print("wrote demo_input", a.shape, a.dtype)


def main():
    gu, co, mn, cgc = ref_shim.load()
    import warnings

    warnings.simplefilter("ignore")
    for name in CASES:
        save_graph(gu, name)
    save_cheb_conv(gu, cgc)
    save_demo_input()
    save_meshnet(gu, mn, "smpl_small")
    save_meshnet(gu, mn, "mano_like")


if __name__ == "__main__":
    main()
