"""Turn the ncu captures brought back in gpurun_out/ into the tracked text summaries under profiles/.
    python scripts/summarize_ncu.py <tag>
"""
import collections
import csv
import re
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
out = []

# ---- launch list (gpu__time_duration per launch; cold-cache, serialised: compare SHARES)
rows = list(csv.DictReader(l for l in open("gpurun_out/launches_b8.csv") if l.startswith('"')))
agg, tot = collections.OrderedDict(), 0.0
for r in rows:
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("opp::", "").replace("void ", "")
    t = float(r["Metric Value"]) / 1e3
    tot += t
    n, s = agg.get(name, (0, 0.0))
    agg[name] = (n + 1, s + t)
out.append(f"# ncu launch list — one forward, batch 8, 512x512, 5000 points, fp16x3 ({len(rows)} launches, {tot:.0f} us)\n")
out.append("command: ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off python scripts/profile_step.py 8\n")
out.append("| kernel | launches | total us | share |\n|---|---|---|---|")
for k, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| `{k}` | {n} | {s:.1f} | {100 * s / tot:.1f} % |")

WANT = [("gpu__time_duration.sum", "time"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
        ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem->TC wavefronts %"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
        ("launch__registers_per_thread", "regs"), ("launch__cluster_size", "cluster")]
for rep in ("prof_conv", "prof_xfmr", "prof_simt"):
    try:
        txt = subprocess.run(["ncu", "-i", f"gpurun_out/{rep}.ncu-rep", "--page", "raw", "--csv"],
                             capture_output=True, text=True).stdout
    except OSError:
        continue
    rr = list(csv.reader(txt.splitlines()))
    if len(rr) < 3:
        continue
    hdr, units = rr[0], rr[1]
    out.append(f"\n# ncu --set full: {rep} (batch 8)\n")
    out.append("| kernel | grid | " + " | ".join(n for _, n in WANT) + " |\n|---|---|" + "---|" * len(WANT))
    for r in rr[2:]:
        name = re.sub(r"\(.*", "", r[hdr.index("Kernel Name")]).replace("opp::", "").replace("void ", "")
        cells = []
        for key, _ in WANT:
            if key in hdr:
                i = hdr.index(key)
                cells.append(f"{r[i]} {units[i]}".strip())
            else:
                cells.append("-")
        out.append(f"| `{name}` | {r[hdr.index('Grid Size')]} | " + " | ".join(cells) + " |")
open(f"profiles/{tag}_ncu_summary.md", "w").write("\n".join(out) + "\n")
print("\n".join(out))
