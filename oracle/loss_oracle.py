"""TEST INFRASTRUCTURE ONLY — CPU (torch) restatement of the reference's mesh losses, lib/core/loss.py:10-23,62-114.

Only tests/ may import this.  Parity status: PINNED against the unmodified reference classes
(tests/golden/mesh_losses.npz, made by tests/golden/make_golden_loss.py; tests/test_oracle_golden.py).
"""
import torch
import torch.nn.functional as F


def coord_loss(pred, target, valid=None):
    """loss.py:17-23."""
    if valid is not None:
        pred, target = pred * valid, target * valid
    return (pred - target).abs().mean()


def normal_vector_loss(coord_out, coord_gt, face):
    """loss.py:67-87: |cos| between the (unit) edges of the predicted face and the ground-truth face normal."""
    f = torch.as_tensor(face, dtype=torch.long)
    a, b, c = coord_out[:, f[:, 0]], coord_out[:, f[:, 1]], coord_out[:, f[:, 2]]
    v1, v2, v3 = F.normalize(b - a, dim=2), F.normalize(c - a, dim=2), F.normalize(c - b, dim=2)
    ga, gb, gc = coord_gt[:, f[:, 0]], coord_gt[:, f[:, 1]], coord_gt[:, f[:, 2]]
    n = F.normalize(torch.cross(F.normalize(gb - ga, dim=2), F.normalize(gc - ga, dim=2), dim=2), dim=2)
    cos = [(v * n).sum(2, keepdim=True).abs() for v in (v1, v2, v3)]
    return torch.cat(cos, 1).mean()


def edge_length_loss(coord_out, coord_gt, face):
    """loss.py:96-114: | |edge(out)| - |edge(gt)| | over the three edges (0,1), (0,2), (1,2) of every face."""
    f = torch.as_tensor(face, dtype=torch.long)

    def lens(x):
        a, b, c = x[:, f[:, 0]], x[:, f[:, 1]], x[:, f[:, 2]]
        return [((p - q) ** 2).sum(2, keepdim=True).sqrt() for p, q in ((a, b), (a, c), (b, c))]

    return torch.cat([(o - g).abs() for o, g in zip(lens(coord_out), lens(coord_gt))], 1).mean()
