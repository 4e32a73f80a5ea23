"""Turn the scratch outputs of tools/round_check.sh ncu (gpurun_out/) into the tracked summaries under profiles/:
    python tools/summarize_profiles.py r01_v3 "one line describing the build"
writes profiles/<tag>_launches.md, profiles/<tag>_ncu_full.md (tables only; append the reading by hand) and copies
the bench line to profiles/<tag>_bench.json."""
import collections
import csv
import io
import re
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "gpurun_out"
tag, note = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")

# ---- launch list --------------------------------------------------------------------------------------------------
lines = [ln for ln in open(OUT / "launches.csv") if not ln.startswith("==")]
r = csv.reader(lines)
hdr = next(r)
ki, vi, ui, mi = (hdr.index(k) for k in ("Kernel Name", "Metric Value", "Metric Unit", "Metric Name"))
agg, tot, n = collections.OrderedDict(), 0.0, 0
for row in r:
    if len(row) <= vi or row[mi] != "gpu__time_duration.sum":
        continue
    m = re.search(r"(\w+_kernel)(<[^>]*>)?", row[ki])
    key = (m.group(1) + (m.group(2) or "")) if m else row[ki][:40]
    v = float(row[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(row[ui], 1.0)
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += v
    tot += v
    n += 1
out = [f"# {tag} — ncu launch list of ONE fused CFG++ step (SDXL 1024x1024, batch 2 => UNet batch 4), B200", "",
       "command: ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none python tools/ncu_target.py 1",
       f"(per-launch times are cold-cache and serialised: compare SHARES)  total {tot / 1e3:.2f} ms over {n} launches",
       f"({note})", "", "| kernel | launches | total us | share | avg us |", "|---|---:|---:|---:|---:|"]
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| {k} | {c} | {t:.1f} | {100 * t / tot:.1f}% | {t / c:.1f} |")
(ROOT / "profiles" / f"{tag}_launches.md").write_text("\n".join(out) + "\n")

# ---- ncu --set full captures ----------------------------------------------------------------------------------------
WANT = ["Kernel Name", "Grid Size", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "launch__registers_per_thread", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max"]


def table(rep, title):
    raw = subprocess.run(["ncu", "-i", str(OUT / rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, units = rows[0], rows[1]
    idx = [h.index(w) if w in h else None for w in WANT]
    ls = [f"## {title}", "", "| " + " | ".join(f"{w} [{units[i]}]" if i is not None and units[i] else w for w, i in zip(WANT, idx)) + " |",
          "|" + "---|" * len(WANT)]
    for row in rows[2:]:
        cells = [(row[i] if i is not None else "-") for i in idx]
        m = re.search(r"(\w+_kernel<[^>]*>)", cells[0])
        cells[0] = m.group(1) if m else cells[0][:40]
        ls.append("| " + " | ".join(cells) + " |")
    return "\n".join(ls)


doc = [f"# {tag} — `ncu --set full --clock-control none --import-source on` captures (B200), selected raw metrics", "",
       f"({note})", "",
       table("prof_gemm.ncu-rep", "gemm_kernel<BN,GEGLU,CL>: 12 consecutive GEMM launches inside the fused step (-k regex:gemm_kernel -s 200 -c 12)"), "",
       table("prof_attn.ncu-rep", "attn_kernel<HD,NQT,KS>: 4 consecutive attention launches (-k regex:attn_kernel -s 30 -c 4)"), ""]
(ROOT / "profiles" / f"{tag}_ncu_full.md").write_text("\n".join(doc))
shutil.copy(OUT / "bench.json", ROOT / "profiles" / f"{tag.replace('_v', '_bench_v')}.json")
print((ROOT / "profiles" / f"{tag}_launches.md").read_text())
print((ROOT / "profiles" / f"{tag}_ncu_full.md").read_text())
