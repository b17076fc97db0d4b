"""Profiling aid (GPU box): N eval forwards of the bench workload (B=256, SMPL-size hierarchy, fp16x3),
run eagerly so that ncu sees every launch.  Usage: python tools/ncu_forward.py [n_forward=2] [batch=256] [mode=fwd|train]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from pose2mesh_release_b200.meshnet import Pose2Mesh  # noqa: E402

n_fwd = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
mode = sys.argv[3] if len(sys.argv) > 3 else "fwd"
graph_L, perm_rev = bench.build_problem("smpl")
torch.manual_seed(123)
model = Pose2Mesh(5, 3, graph_L, joint_set="human36")
model.load_state_dict(bench.randomize_bn_({k: v.clone() for k, v in model.state_dict().items()}))
model = model.cuda().set_precision("fp16x3")
x = torch.randn(B, 17, 5, generator=torch.Generator().manual_seed(1000)).cuda()
if mode == "fwd":
    model.eval()
    with torch.no_grad():
        for _ in range(n_fwd):
            y = model(x)
else:
    model.train()
    tgt = torch.randn(B, model.num_vertices, 3, generator=torch.Generator().manual_seed(7)).cuda()
    for _ in range(n_fwd):
        model.zero_grad()
        (model(x) - tgt).abs().mean().backward()
torch.cuda.synchronize()
print("ok", tuple(y.shape) if mode == "fwd" else "train")
