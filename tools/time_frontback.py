"""Timing aid (GPU box): PoseNet / FlatPose2Mesh front half, joint regression and mesh losses at B=256."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from pose2mesh_release_b200 import posenet, postprocess, loss as L, graph as pg

def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

B = 256
net = posenet.get_model(17, 4096, 2, 0.5).cuda().eval()
x = torch.randn(B, 34, device="cuda")
with torch.no_grad():
    print("posenet native  ms", timeit(lambda: net.forward_native(x, with_combine=True)))
    net.train(); net.eval()
    ref = torch.nn.Sequential()  # torch/cuBLAS reference of the same math for comparison
    def torch_path():
        y = net.w1(x)
        for st in net.linear_stages:
            y = st(y)
        return net.w2(y)
    print("posenet torch   ms", timeit(torch_path))
verts = torch.randn(B, 6890, 3, device="cuda"); jr = torch.rand(17, 6890, device="cuda")
print("regress_joints  ms", timeit(lambda: postprocess.regress_joints(verts, jr)))
face = pg.synthetic_sphere_faces(6890, 2)
ml = L.MeshLosses(face)
gt = torch.randn(B, 6890, 3, device="cuda")
def loss_fb():
    o = verts.clone().requires_grad_(True)
    a, b = ml(o, gt); (a + b).backward()
print("mesh losses f+b ms", timeit(loss_fb))
