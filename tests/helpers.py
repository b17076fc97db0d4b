"""Shared helpers for the test-suite: golden-fixture loading and oracle plumbing."""
import os

import numpy as np
import scipy.sparse as sp
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_npz(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def graph_from_fixture(name):
    """Rebuild the reference's graph_L list (scipy CSR float64, fine -> coarse, joint graph last)."""
    z = load_npz(f"graph_{name}.npz")
    mats = []
    for i in range(int(z["n_levels"])):
        shape = tuple(int(s) for s in z[f"L{i}_shape"])
        mats.append(sp.csr_matrix((z[f"L{i}_data"], z[f"L{i}_indices"], z[f"L{i}_indptr"]), shape=shape))
    return mats, z


def tensor_digest(t: torch.Tensor) -> np.ndarray:
    d = t.detach().double().cpu().flatten()
    return np.array([d.sum().item(), d.abs().sum().item(), (d * d).sum().item()] + d[:5].tolist())


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max |b|  (SURVEY.md §8d parity measure)."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


CASES = {"smpl_small": (1200, 0, 9, False), "mano_like": (778, 1, 6, True), "smpl_like": (6890, 2, 9, False)}
