"""Debug aid (GPU box): per-parameter gradient error of the training step vs autograd over the CPU oracle."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import graph_from_fixture, CASES
from oracle import meshnet_oracle as mo
from pose2mesh_release_b200.meshnet import Pose2Mesh

name = sys.argv[1] if len(sys.argv) > 1 else "mano_like"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n, seed, levels, mano = CASES[name]
mats, _ = graph_from_fixture(name)
torch.manual_seed(123)
model = Pose2Mesh(5, 3, mats, joint_set="mano" if mano else "human36").cuda()
laps = mo.laplacians_to_torch(mats)
sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
g = torch.Generator().manual_seed(5)
J = laps[-1].shape[0]
x = torch.randn(B, J, 5, generator=g); tgt = torch.randn(B, laps[0].shape[0], 3, generator=g)

def oracle(dt):
    l2 = [torch.sparse_csr_tensor(l.crow_indices(), l.col_indices(), l.values().to(dt), size=l.shape) for l in laps]
    s = {k: (v.clone().to(dt).requires_grad_(True) if v.is_floating_point() and 'running' not in k else
             (v.clone().to(dt) if v.is_floating_point() else v.clone())) for k, v in sd.items()}
    xo = x.clone().to(dt).requires_grad_(True)
    acts = []
    y = mo.forward(s, l2, xo, mano=mano, training=True, collect=acts)
    (y - tgt.to(dt)).abs().mean().backward()
    return xo.grad, {k: v.grad for k, v in s.items() if v.requires_grad}, y.detach()

gx, gp, yo = oracle(torch.float64)
model.train()
xg = x.cuda().requires_grad_(True)
y = model(xg)
(y - tgt.cuda()).abs().mean().backward()
print("y rel err", ((y.detach().cpu().double() - yo).abs().max() / yo.abs().max()).item())
print("dx rel err", ((xg.grad.cpu().double() - gx).abs().max() / gx.abs().max()).item())
for k, p in model.named_parameters():
    ref = gp[k]
    e = (p.grad.cpu().double() - ref).abs().max().item()
    print(f"{k:14s} max|ref| {ref.abs().max().item():.3e}  rel err {e / max(ref.abs().max().item(), 1e-30):.3e}")
