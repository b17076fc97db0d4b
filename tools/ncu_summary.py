"""Summarise an .ncu-rep: key raw metrics per kernel + stall mix, instruction mix and hottest SASS of one kernel."""
import collections, csv, io, re, subprocess, sys

rep = sys.argv[1]
kid = sys.argv[2] if len(sys.argv) > 2 else "0"
tiles = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__cycles_active.avg", "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active"]
for n, r in enumerate(rows[2:]):
    d = dict(zip(hdr, r))
    print(f"--- kernel {n}")
    for k in keys:
        if k in d:
            print(f"  {k} = {d[k]} {units[hdr.index(k)]}")

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + (["--kernel-id", f":::{kid}"] if kid != "-" else []), capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0].startswith("0x")]
seen, uniq = set(), []
for r in data:
    if r[0] not in seen:
        seen.add(r[0]); uniq.append(r)
data = uniq
cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter()
for r in data:
    for c in cols:
        if r[idx[c]]:
            tot[c] += int(r[idx[c]])
s = sum(tot.values())
print("stalls:", {c: round(100 * v / s, 1) for c, v in tot.most_common(9)})
inst = sum(int(r[idx["Instructions Executed"]]) for r in data)
print(f"warp instructions: {inst}  ({inst / tiles:.0f} per tile)")
byop = collections.Counter()
for r in data:
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[idx["Source"]].strip())
    byop[m.group(2) if m else r[idx["Source"]]] += int(r[idx["Instructions Executed"]])
print("instruction mix (per tile):", [(op, round(c / tiles)) for op, c in byop.most_common(18)])
ts = sum(int(r[idx["# Samples"]]) for r in data)
order = sorted(range(len(data)), key=lambda i: -int(data[i][idx["# Samples"]]))
print("hottest SASS (share of samples, executions, instruction):")
for i in order[:14]:
    r = data[i]
    ctx = data[i - 1][idx["Source"]].strip()[:60] if i else ""
    print(f"  {100 * int(r[idx['# Samples']]) / ts:5.1f}% {r[idx['Instructions Executed']]:>10s}  {r[idx['Source']].strip()[:70]}   <- after: {ctx}")
