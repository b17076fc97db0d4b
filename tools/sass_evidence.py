"""Write profiles/r2_sass_cheb_umma.txt: per tcgen05 kernel of libp2m_b200.so, the counts of the SASS mnemonics that
prove the Blackwell-native path (B200_PROFILING.md: UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG /
UBLKCP = TMA, UTCBAR = tcgen05.commit, SYNCS = mbarrier) plus an excerpt around the first MMA of each kernel.
Run here (no GPU needed): python tools/sass_evidence.py"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pose2mesh_release_b200", "libp2m_b200.so")
OUT = os.path.join(ROOT, "profiles", "r2_sass_cheb_umma.txt")
KEYS = ["UTCHMMA", "UTCBAR", "UTMALDG", "UBLKCP", "LDTM", "STTM", "UTCATOMSWS", "LDGSTS", "SYNCS", "HMMA", "HGMMA"]

sass = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", sass)
lines = [f"cuobjdump -sass {os.path.relpath(LIB, ROOT)}  (built {os.popen('date -u -r ' + LIB).read().strip()})",
         "per kernel: SASS mnemonic counts; kernels without any tensor / TMA instruction are listed by name only", ""]
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip() or n
plain = []
for f in funcs[1:]:
    name, _, body = f.partition("\n")
    cnt = collections.Counter()
    for ln in body.splitlines():
        m = re.search(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m:
            op = m.group(1).split(".")[0]
            if op in KEYS:
                cnt[op] += 1
    dn = demangle(name.strip())
    if not any(cnt[k] for k in ("UTCHMMA", "UTMALDG", "UBLKCP", "LDTM")):
        plain.append(dn)
        continue
    lines.append(f"== {dn}")
    lines.append("   " + "  ".join(f"{k}={cnt[k]}" for k in KEYS if cnt[k]))
    body_l = body.splitlines()
    first = next((i for i, ln in enumerate(body_l) if "UTCHMMA" in ln), None)
    if first is not None:
        for ln in body_l[max(0, first - 3):first + 8]:
            m = re.search(r"/\*[0-9a-f]+\*/\s+(.*?)\s*;", ln)
            if m:
                lines.append("      " + m.group(1))
    lines.append("")
lines.append("kernels without tensor-core / TMA instructions (SIMT: sparse basis, BatchNorm, thin head, packing, losses ...):")
lines += ["   " + p[:150] for p in plain]
open(OUT, "w").write("\n".join(lines) + "\n")
print("wrote", OUT, len(lines), "lines")
