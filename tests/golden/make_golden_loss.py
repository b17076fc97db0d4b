#!/usr/bin/env python
"""Golden vectors of the reference's mesh losses (lib/core/loss.py), produced by the UNMODIFIED reference classes in
the build container (CPU; Tensor.cuda -> identity):   python tests/golden/make_golden_loss.py -> mesh_losses.npz"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import graph_oracle as go  # noqa: E402
from oracle import ref_shim  # noqa: E402


def load_ref_loss():
    import importlib.util
    import types

    ref_shim.load("human36")
    if "funcs_utils" not in sys.modules:
        try:
            import funcs_utils  # noqa: F401
        except Exception:
            m = types.ModuleType("funcs_utils")
            m.stop = lambda *a, **k: None
            sys.modules["funcs_utils"] = m
    spec = importlib.util.spec_from_file_location("ref_core_loss", os.path.join(ref_shim.REF_ROOT, "lib", "core", "loss.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    ref_loss = load_ref_loss()
    face = go.synthetic_sphere_faces(300, 4).astype(np.int64)
    g = torch.Generator().manual_seed(12)
    B, nv = 3, 300
    gt = torch.randn(B, nv, 3, generator=g)
    out = (gt + 0.3 * torch.randn(B, nv, 3, generator=g)).requires_grad_(True)
    valid = (torch.rand(B, nv, 1, generator=g) > 0.2).float()
    with ref_shim.cpu_cuda_noop():
        ln = ref_loss.NormalVectorLoss(face)(out, gt)
        le = ref_loss.EdgeLengthLoss(face)(out, gt)
        lc = ref_loss.CoordLoss(has_valid=True)(out, gt, valid)
        (0.7 * ln + 1.3 * le + 0.5 * lc).backward()
    path = os.path.join(HERE, "mesh_losses.npz")
    np.savez_compressed(path, face=face.astype(np.int32), out=out.detach().numpy(), gt=gt.numpy(), valid=valid.numpy(),
                        normal=ln.item(), edge=le.item(), coord=lc.item(), weights=np.array([0.7, 1.3, 0.5]),
                        grad=out.grad.numpy())
    print("wrote", path, os.path.getsize(path))
