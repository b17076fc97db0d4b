"""Print the README results table from profiles/r2_bench_*.json."""
import json, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")
def load(n):
    p = os.path.join(P, n)
    return json.load(open(p)) if os.path.exists(p) and os.path.getsize(p) else None
rows = []
n1, n2, tr, mf, mt, ref, n8 = (load(f"r2_bench_{k}.json") for k in ("n1", "n2", "train_n1", "mano_fwd", "mano_train", "reference_arm", "n8"))
def k(v): return f"{v / 1e3:.2f} k"
if n1:
    rows.append(("B=256 SMPL-size eval forward (the bench line, configs[1])", f"{n1['ms_per_step']:.2f} ms/step, **{k(n1['value'])} meshes/s**", f"{k(n1['e2e']['value'])} meshes/s ({n1['e2e']['d2h_bytes_per_step'] / 1e6:.1f} MB back per step)"))
    t = n1.get("train")
    if t: rows.append(("B=256 forward+backward, L1 loss, train-mode BatchNorm (configs[2])", f"{t['ms_per_step']:.1f} ms/step, {k(t['value'])} meshes/s", f"{k(t['e2e']['value'])}"))
if n2:
    rows.append(("2 GPUs (torchrun, weak scaling, 256 per GPU): forward", f"{k(n2['value'])} meshes/s ({n2['value'] / n1['value']:.2f}x)" if n1 else k(n2['value']), f"{k(n2['e2e']['value'])}"))
    t = n2.get("train")
    if t: rows.append(("2 GPUs: training step incl. the 35 MB gradient all-reduce", f"{t['ms_per_step']:.1f} ms/step, {k(t['value'])} meshes/s; all-reduce {t['allreduce_ms']:.2f} ms exposed", f"{k(t['e2e']['value'])}"))
if n8:
    rows.append(("8 GPUs (configs[4]: 2048 poses, 256 per GPU): forward", f"{k(n8['value'])} meshes/s ({n8['value'] / n1['value']:.2f}x)" if n1 else k(n8['value']), f"{k(n8['e2e']['value'])}"))
    t = n8.get("train")
    if t: rows.append(("8 GPUs: training step incl. the gradient all-reduce", f"{t['ms_per_step']:.1f} ms/step, {k(t['value'])} meshes/s ({t['value'] / n1['train']['value']:.2f}x); all-reduce {t['allreduce_ms']:.2f} ms exposed", f"{k(t['e2e']['value'])}"))
if mf: rows.append(("MANO-size hierarchy (1088..68), B=1024 forward (configs[3])", f"{mf['ms_per_step']:.2f} ms/step, {k(mf['value'])} meshes/s", f"{k(mf['e2e']['value'])}"))
if mt: rows.append(("MANO-size, B=1024 forward+backward", f"{mt['ms_per_step']:.1f} ms/step, {k(mt['value'])} meshes/s", f"{k(mt['e2e']['value'])}"))
if ref: rows.append((f"CPU port of the reference on the same box ({ref['cpu_baseline']['cores']} of {ref['cpu_baseline']['host_cores']} threads, its best)", f"{ref['value']:.1f} meshes/s", "—"))
print("| config | device-timed | end to end (host buffers) |\n|---|---|---|")
for r in rows: print("| " + " | ".join(r) + " |")
if n1:
    r = n1["roofline"]
    print(f"\nDominant layer (V=12288, 128→128, B=256): {r['ms_per_launch']:.2f} ms → {r['achieved']:.0f} GB/s of algorithmic bytes = "
          f"**{r['frac']:.3f} of the measured HBM peak** ({r['peak']:.0f} GB/s); DRAM traffic of the layer {json.load(open(os.path.join(P, 'r2_ncu_traffic.json')))['layer_dram_bytes'] / 1e9:.2f} GB "
          f"vs {r['algorithmic_bytes'] / 1e9:.2f} GB algorithmic.")
