"""Turn the scratch outputs of tools/profile_round.sh (gpurun_out/) into the tracked summaries under profiles/:

    python tools/summarize_profiles.py r02 "one line describing the build"

writes profiles/<tag>_launches.md (shares of the serialised step), profiles/<tag>_kernels.md (per kernel class: time,
achieved DRAM GB/s against the measured HBM peak, tensor / XU pipe activity), profiles/<tag>_ncu_full.md (selected
raw metrics of the `--set full` captures) and profiles/roofline_traffic.json (what bench.py reports as
roofline.traffic: mean dram read + write bytes per launch over the full-set GEMM capture). Nothing is typed by hand."""
import collections
import csv
import io
import json
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "gpurun_out"
PROF = ROOT / "profiles"
tag, note = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
HBM = peaks.get("hbm_gbs", 6650.0)
UNIT = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "%": 1.0}


def kname(full):
    m = re.search(r"(\w+_kernel)(<[^>]*>)?", full)
    return (m.group(1) + (m.group(2) or "")) if m else full[:48]


def read_long_csv(path):
    """ncu --csv (one row per launch x metric) -> [ {ID, name, metric: value...} ] in launch order."""
    lines = [ln for ln in open(path) if not ln.startswith("==")]
    r = csv.reader(lines)
    hdr = next(r)
    ii, ki, vi, ui, mi = (hdr.index(k) for k in ("ID", "Kernel Name", "Metric Value", "Metric Unit", "Metric Name"))
    rows = collections.OrderedDict()
    for row in r:
        if len(row) <= vi:
            continue
        d = rows.setdefault(row[ii], {"name": kname(row[ki])})
        try:
            d[row[mi]] = float(row[vi].replace(",", "")) * UNIT.get(row[ui], 1.0)
        except ValueError:
            pass
    return list(rows.values())


# ---- launch list ---------------------------------------------------------------------------------------------------
rows = read_long_csv(OUT / "launches.csv")
agg, tot = collections.OrderedDict(), 0.0
for d in rows:
    a = agg.setdefault(d["name"], [0, 0.0])
    a[0] += 1
    a[1] += d.get("gpu__time_duration.sum", 0.0)
    tot += d.get("gpu__time_duration.sum", 0.0)
out = [f"# {tag} — ncu launch list of ONE fused CFG++ step (SDXL 1024x1024, batch 2 => UNet batch 4), B200", "",
       "command: ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none python tools/ncu_target.py 1",
       f"(per-launch times are cold-cache and serialised: compare SHARES)  total {tot / 1e3:.2f} ms over {len(rows)} launches",
       f"({note})", "", "| kernel | launches | total us | share | avg us |", "|---|---:|---:|---:|---:|"]
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| {k} | {c} | {t:.1f} | {100 * t / tot:.1f}% | {t / c:.1f} |")
(PROF / f"{tag}_launches.md").write_text("\n".join(out) + "\n")

# ---- per-class metrics -----------------------------------------------------------------------------------------------
if (OUT / "metrics.csv").exists():
    rows = read_long_csv(OUT / "metrics.csv")
    T, R, W = "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum"
    TP = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
    XU = "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"
    IS = "smsp__issue_active.avg.pct_of_peak_sustained_active"
    L2 = "lts__t_bytes.sum"
    agg = collections.OrderedDict()
    for d in rows:
        a = agg.setdefault(d["name"], collections.defaultdict(float))
        t = d.get(T, 0.0)
        a["n"] += 1
        a["t"] += t
        a["dram"] += d.get(R, 0.0) + d.get(W, 0.0)
        a["l2"] += d.get(L2, 0.0)
        for key, m in (("tp", TP), ("xu", XU), ("is", IS)):
            a[key] += d.get(m, 0.0) * t  # time-weighted
    tot = sum(a["t"] for a in agg.values())
    out = [f"# {tag} — per kernel class, all launches of ONE fused CFG++ step (SDXL, UNet batch 4), B200", "",
           "command: ncu --profile-from-start off --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,"
           "dram__bytes_write.sum,sm__pipe_tensor_cycles_active...,sm__inst_executed_pipe_xu...,smsp__issue_active...,lts__t_bytes.sum",
           f"(ncu serialises launches and flushes caches between replays: DRAM bytes are cold-cache upper bounds; pipe "
           f"percentages are time-weighted means; HBM peak = {HBM:.1f} GB/s measured)  ({note})", "",
           "| kernel | launches | total us | share | avg us | DRAM MB/launch | achieved DRAM GB/s | % of HBM peak | L2 MB/launch | tensor pipe % | XU pipe % | issue active % |",
           "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["t"]):
        gbs = a["dram"] / (a["t"] * 1e-6) / 1e9 if a["t"] else 0.0
        out.append(f"| {k} | {int(a['n'])} | {a['t']:.1f} | {100 * a['t'] / tot:.1f}% | {a['t'] / a['n']:.1f} | "
                   f"{a['dram'] / a['n'] / 1e6:.2f} | {gbs:.0f} | {100 * gbs / HBM:.1f}% | {a['l2'] / a['n'] / 1e6:.1f} | "
                   f"{a['tp'] / a['t']:.1f} | {a['xu'] / a['t']:.1f} | {a['is'] / a['t']:.1f} |")
    (PROF / f"{tag}_kernels.md").write_text("\n".join(out) + "\n")

# ---- ncu --set full captures ---------------------------------------------------------------------------------------
WANT = ["Kernel Name", "Grid Size", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "launch__registers_per_thread", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.max"]


def raw_rows(rep):
    raw = subprocess.run(["ncu", "-i", str(OUT / rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    return rows[0], rows[1], rows[2:]


def table(rep, title):
    if not (OUT / rep).exists():
        return f"## {title}\n\n(capture {rep} missing)"
    h, units, body = raw_rows(rep)
    idx = [h.index(w) if w in h else None for w in WANT]
    ls = [f"## {title}", "", "| " + " | ".join(f"{w} [{units[i]}]" if i is not None and units[i] else w for w, i in zip(WANT, idx)) + " |",
          "|" + "---|" * len(WANT)]
    for row in body:
        cells = [(row[i] if i is not None else "-") for i in idx]
        cells[0] = kname(cells[0])
        ls.append("| " + " | ".join(cells) + " |")
    return "\n".join(ls)


CAPS = [("prof_conv.ncu-rep", "implicit-GEMM conv3x3 (gemm_kernel, conv mode): down_blocks.0.resnets.0 conv1 / conv2, 4 x 128x128 x 320 -> 320"),
        ("prof_gemm.ncu-rep", "gemm_kernel<BN,GEGLU,CL>: 12 consecutive GEMM launches of the 1280-channel transformer blocks (-s 200 -c 12)"),
        ("prof_attn.ncu-rep", "attn_kernel<64,2,3>: self-attention N = 1024, 20 heads, UNet batch 4"),
        ("prof_xattn.ncu-rep", "xattn_kernel<64,2,80>: cross-attention, 77 keys"),
        ("prof_gn.ncu-rep", "gn_stats_kernel / gn_apply_kernel at 128x128x320 (HBM-bound)"),
        ("prof_convio.ncu-rep", "conv_in_kernel and conv_out_step_kernel (conv_out + CFG++ mix + DDIM update fused)"),
        ("prof_small.ncu-rep", "latency-class helpers: select_step, sincos, small_linear, im2col_s2, upsample2x")]
doc = [f"# {tag} — `ncu --set full --clock-control none --import-source on` captures (B200), selected raw metrics", "", f"({note})", ""]
for rep, title in CAPS:
    doc += [table(rep, title), ""]
(PROF / f"{tag}_ncu_full.md").write_text("\n".join(doc))

# ---- roofline.traffic for bench.py ---------------------------------------------------------------------------------
if (OUT / "prof_gemm.ncu-rep").exists():
    h, units, body = raw_rows("prof_gemm.ncu-rep")
    ri, wi = h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
    vals = [float(r[ri].replace(",", "")) * UNIT.get(units[ri], 1.0) + float(r[wi].replace(",", "")) * UNIT.get(units[wi], 1.0)
            for r in body]
    tj = {}
    if (PROF / "roofline_traffic.json").exists():
        tj = json.loads((PROF / "roofline_traffic.json").read_text())
    tj["sdxl_b2"] = {"dram_bytes_per_launch": sum(vals) / len(vals), "launches": len(vals),
                     "source": f"ncu --set full, prof_gemm.ncu-rep ({tag}): mean dram__bytes_read.sum + dram__bytes_write.sum over "
                               f"{len(vals)} consecutive gemm_kernel launches of the fused step (cold L2), summarised in "
                               f"profiles/{tag}_ncu_full.md by tools/summarize_profiles.py"}
    (PROF / "roofline_traffic.json").write_text(json.dumps(tj, indent=1) + "\n")
for f in (f"{tag}_launches.md", f"{tag}_kernels.md"):
    if (PROF / f).exists():
        print((PROF / f).read_text())
