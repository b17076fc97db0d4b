"""Debug aid (GPU box): run one tcgen05 conv layer and print the error structure vs the CPU oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from helpers import graph_from_fixture
from oracle import meshnet_oracle as mo
from pose2mesh_release_b200 import cheby_graph_conv as cgc

def run(name, level, b, fin, fout, mode):
    mats, _ = graph_from_fixture(name)
    L = mats[level]; V = L.shape[0]
    lap = mo.laplacians_to_torch([L], drop_second_coarsest=False)[0]
    g = torch.Generator().manual_seed(17)
    x = torch.randn(b, V, fin, generator=g)
    w = torch.zeros(fout, fin, 3)
    if mode == "rand":
        w = (torch.rand(fout, fin, 3, generator=g) * 2 - 1) * 0.1
    else:  # only Chebyshev order k active, "identity-like"
        k = int(mode[1])
        for n in range(fout):
            w[n, n % fin, k] = 1.0
    w = w.reshape(fout, 3 * fin)
    bias = torch.zeros(fout)
    cl = torch.nn.Linear(3 * fin, fout).cuda()
    cl.weight.data.copy_(w); cl.bias.data.copy_(bias)
    cgc.set_default_precision("fp16x3")
    gh = cgc.graph_handle(L)
    with torch.no_grad():
        y = cgc.graph_conv_cheby(x.cuda(), cl, None, L, fout, 3).cpu()
    st = gh.kernel_status(0)
    cgc.set_default_precision("fp32")
    with torch.no_grad():
        y32 = cgc.graph_conv_cheby(x.cuda(), cl, None, L, fout, 3).cpu()
    yo = mo.cheb_conv(x, lap, w, bias)
    e = (y - yo).abs(); scale = yo.abs().max().item()
    print(f"[{name} L{level} V={V} B={b} {fin}->{fout} mode={mode}] status={st} rel err umma={e.max().item()/scale:.3e} "
          f"simt={(y32-yo).abs().max().item()/scale:.3e}  |y|max={y.abs().max().item():.3e} |yo|max={scale:.3e}")
    if e.max().item() / scale > 1e-5:
        er = e.max(dim=2).values.max(dim=0).values  # per vertex
        ec = e.max(dim=1).values.max(dim=0).values  # per column
        bad_r = (er > 1e-5 * scale).nonzero().flatten()
        bad_c = (ec > 1e-5 * scale).nonzero().flatten()
        print("   bad rows:", len(bad_r), "of", V, "first", bad_r[:16].tolist(), " bad cols:", len(bad_c), "of", fout, bad_c[:16].tolist())
        print("   y[0,0,:8]  =", y[0, 0, :8].tolist())
        print("   yo[0,0,:8] =", yo[0, 0, :8].tolist())
        print("   y[0,1,:8]  =", y[0, 1, :8].tolist())
        print("   yo[0,1,:8] =", yo[0, 1, :8].tolist())

if __name__ == "__main__":
    for mode in ("k0", "k1", "k2", "rand"):
        run("smpl_small", 4, 1, 32, 64, mode)      # V=128: one tile, one chunk
    run("smpl_small", 4, 1, 64, 128, "rand")
    run("smpl_small", 0, 2, 128, 128, "rand")
    run("smpl_small", 1, 2, 256, 256, "rand")
    run("mano_like", 0, 2, 128, 128, "rand")
