"""Debug aid (GPU box): event timeline of CTA 0 of the tcgen05 conv kernel for one full-size layer."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from pose2mesh_release_b200 import graph as pg, cheby_graph_conv as cgc, _lib

fin, fout, B = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
level = int(sys.argv[4]) if len(sys.argv) > 4 else 0
face = pg.synthetic_sphere_faces(6890, 2)
_, graph_L, _, _ = pg.build_coarse_graphs(face, 17, pg.H36M_SKELETON, pg.H36M_FLIP_PAIRS, levels=9)
L = graph_L[level]
V = L.shape[0]
x = torch.randn(B, V, fin, device="cuda")
cl = torch.nn.Linear(3 * fin, fout).cuda()
cgc.set_default_precision("fp16x3")
lib = _lib.load()
buf = torch.zeros(8 * 512, dtype=torch.int64, device="cuda")
with torch.no_grad():
    cgc.graph_conv_cheby(x, cl, None, L, fout, 3)   # warm-up
    h = cgc.graph_handle(L).handle(0)
    _lib.check(lib.p2m_debug_set_trace(h, buf.data_ptr()), "set_trace (needs a P2M_TRACE=1 build)")
    cgc.graph_conv_cheby(x, cl, None, L, fout, 3)
    torch.cuda.synchronize()
    lib.p2m_debug_set_trace(h, None)
t = buf.cpu().numpy().reshape(8, 512)
names = {0: "producer", 1: "bload", 2: "mma", 3: "epilogue"}
ev_all = []
for role in range(4):
    for v in t[role]:
        if v == 0:
            continue
        ev_all.append((int(v) & 0xFFFFFFFFFFFF, role, int(v) >> 48))
ev_all.sort()
t0 = ev_all[0][0]
print(f"V={V} {fin}->{fout} B={B}: first 260 events of CTA 0 (cycles since start)")
pn = {1: "wait_x", 2: "x_ready", 4: "T1 gathered", 5: "T1 barrier passed", 6: "T2 gathered", 7: "blocks emitted", 8: "end barrier"}
for c, role, ev in ev_all[:260]:
    if role == 0:
        label = pn.get(ev, str(ev))
    elif role == 1:
        label = f"slot free -> load B block {ev - 10}"
    elif role == 2:
        label = "acc buffer free" if ev == 1 else f"A/B block {ev - 10} full -> 6 MMAs"
    else:
        label = "accumulator ready" if ev == 1 else "tile stored"
    print(f"{c - t0:9d}  {names[role]:9s} {label}")
