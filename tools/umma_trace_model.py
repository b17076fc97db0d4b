"""Debug aid (GPU box, P2M_TRACE=1 build): event timeline of CTA 0 of the conv kernel of ONE layer of the eval forward
at the bench workload.  Usage: P2M_TRACE_V=12288 P2M_TRACE_UNPOOL=0 python tools/umma_trace_model.py [n_events=400]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from pose2mesh_release_b200 import _lib  # noqa: E402
from pose2mesh_release_b200.meshnet import Pose2Mesh  # noqa: E402

n_ev = int(sys.argv[1]) if len(sys.argv) > 1 else 400
graph_L, perm_rev = bench.build_problem("smpl")
torch.manual_seed(123)
model = Pose2Mesh(5, 3, graph_L, joint_set="human36")
model.load_state_dict(bench.randomize_bn_({k: v.clone() for k, v in model.state_dict().items()}))
model = model.cuda().set_precision("fp16x3").eval()
x = torch.randn(256, 17, 5, generator=torch.Generator().manual_seed(1000)).cuda()
lib = _lib.load()
buf = torch.zeros(8 * 512, dtype=torch.int64, device="cuda")
with torch.no_grad():
    model(x)
    h = model._hier.handle(0)
    _lib.check(lib.p2m_debug_set_trace(h, buf.data_ptr()), "set_trace (needs a P2M_TRACE=1 build)")
    model(x)
    torch.cuda.synchronize()
    lib.p2m_debug_set_trace(h, None)
t = buf.cpu().numpy().reshape(8, 512)
names = {0: "producer", 1: "bload", 2: "mma", 3: "epilogue", 4: "loader"}
ev_all = []
for role in range(5):
    for v in t[role]:
        if v:
            ev_all.append((int(v) & 0xFFFFFFFFFFFF, role, int(v) >> 48))
ev_all.sort()
t0 = ev_all[0][0]
pn = {1: "wait_x", 2: "x_ready", 4: "T1 gathered", 5: "T1 barrier passed", 6: "T2 gathered", 7: "blocks emitted", 8: "end barrier"}
for c, role, ev in ev_all[:n_ev]:
    if role == 0:
        label = pn.get(ev, str(ev))
    elif role == 1:
        label = f"slot free -> load B block {ev - 10}"
    elif role == 2:
        label = "acc buffer free" if ev == 1 else f"A/B block {ev - 10} full -> 6 MMAs"
    elif role == 3:
        label = "accumulator ready" if ev == 1 else "tile stored"
    else:
        label = "stage free" if ev == 1 else "copies issued"
    print(f"{c - t0:9d}  {names[role]:9s} {label}")
