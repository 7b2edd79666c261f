"""Summarise an `ncu --set full --csv --page raw` log of one forward into a per-launch markdown table
(time, tensor-pipe active %, DRAM bytes and achieved GB/s vs the measured HBM peak, L2 %, L1 LSU
wavefronts %, SM %) plus per-kernel-class totals.  Usage:
    python scripts/summarize_ncu_raw.py gpurun_out/s2/ncu_b64_raw.csv profiles/r2_ncu_b64.md [traffic.json]"""
import csv
import io
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src, dst = sys.argv[1], sys.argv[2]
traffic_out = sys.argv[3] if len(sys.argv) > 3 else None
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
txt = open(src).read().split("\n")
start = [i for i, l in enumerate(txt) if l.startswith('"ID"')][0]
rows = list(csv.reader(io.StringIO("\n".join(txt[start:]))))
hdr, data = rows[0], [r for r in rows[2:] if len(r) == len(rows[0])]


def col(name):
    return hdr.index(name)


C = {"t": col("gpu__time_duration.sum"), "tensor": col("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
     "dr": col("dram__bytes_read.sum"), "dw": col("dram__bytes_write.sum"),
     "l2": col("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
     "sm": col("sm__throughput.avg.pct_of_peak_sustained_elapsed"),
     "lsu": col("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
     "regs": col("launch__registers_per_thread"), "name": col("Kernel Name")}


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


def short(n):
    n = re.sub(r"opp::", "", n)
    m = re.match(r"gemm_kernel<\(?(?:int\))?(\d), *(\w+(?:<[^>]*>)?)(?:, *\(?(?:bool\))?(\w+))?>", n)
    if m:
        return f"gemm<{'CONV' if m.group(1) == '1' else 'ROWS'},{m.group(2)}{',dyn' if m.group(3) in ('1', 'true') else ''}>"
    return re.sub(r"\(.*", "", n)


out = ["| # | kernel | time us | tensor pipe % | DRAM read MB | DRAM write MB | DRAM GB/s | of HBM peak | L2 % | L1 LSU wavefronts % | SM % | regs |",
       "|---|---|---|---|---|---|---|---|---|---|---|---|"]
tot, traffic = {}, {}
for i, r in enumerate(data):
    n = short(r[C["name"]])
    t = num(r[C["t"]]) / 1e3
    dr, dw = num(r[C["dr"]]), num(r[C["dw"]])
    gbs = (dr + dw) / (t * 1e-6) / 1e9 if t else 0
    out.append(f"| {i} | `{n}` | {t:.1f} | {num(r[C['tensor']]):.1f} | {dr / 1e6:.1f} | {dw / 1e6:.1f} | {gbs:.0f} | "
               f"{gbs / peaks['hbm_gbs']:.2f} | {num(r[C['l2']]):.1f} | {num(r[C['lsu']]):.1f} | {num(r[C['sm']]):.1f} | {r[C['regs']]} |")
    a = tot.setdefault(n, [0, 0.0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += t
    a[2] += t * num(r[C["tensor"]])
    a[3] += dr + dw
    key = [k for k in ("EpiConvT<1>", "EpiConv", "EpiStoreF16", "EpiQ", "EpiLN", "EpiLseCol", "EpiConfCol") if k in r[C["name"]]]
    traffic.setdefault(key[0] if key else n, []).append({"us": t, "dram_read_bytes": dr, "dram_write_bytes": dw})
total_t = sum(a[1] for a in tot.values())
summ = ["| kernel | launches | total us | share | time-weighted tensor pipe % | DRAM GB/s (avg) | of HBM peak |", "|---|---|---|---|---|---|---|"]
for n, a in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    gbs = a[3] / (a[1] * 1e-6) / 1e9
    summ.append(f"| `{n}` | {a[0]} | {a[1]:.1f} | {100 * a[1] / total_t:.1f} % | {a[2] / a[1]:.1f} | {gbs:.0f} | {gbs / peaks['hbm_gbs']:.2f} |")
open(dst, "w").write("\n".join(summ) + f"\n\ntotal {total_t:.0f} us over {len(data)} launches (ncu serialises and replays: compare shares)\n\n"
                     + "\n".join(out) + "\n")
if traffic_out:
    json.dump(traffic, open(traffic_out, "w"))
print("\n".join(summ))
