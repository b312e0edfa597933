"""Turn ncu exports into the committed summaries under profiles/.

  python profiles/summarize.py <launch-list.csv> <raw-page.csv> <out.md> [bench.json]

launch-list.csv : ncu --metrics gpu__time_duration.sum --csv --log-file ...   (every launch)
raw-page.csv    : ncu -i prof.ncu-rep --page raw --csv                        (--set full capture)
"""
import collections
import csv
import json
import sys


def launches(path):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, start = r, i + 1
            break
    kn, mv, mn = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    d = collections.OrderedDict()
    for r in rows[start:]:
        if len(r) > mv and r[mn] == "gpu__time_duration.sum":
            name = r[kn].split("(")[0].replace("void ", "")
            d.setdefault(name, []).append(float(r[mv].replace(",", "")) / 1e3)  # ns -> us
    return d


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "launch__grid_size",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
    out = []
    units = rows[1]
    scale = {"Gbyte": 1e3, "Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6,   # bytes -> MB
             "s": 1e6, "ms": 1e3, "us": 1.0, "ns": 1e-3}                 # time -> us
    for r in rows[2:]:
        rec = {"kernel": r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")}
        for w in want:
            if w in hdr:
                v, u = r[hdr.index(w)], units[hdr.index(w)]
                if u in scale and (w.startswith("dram__bytes") or w == "gpu__time_duration.sum"):
                    v = f"{float(v.replace(',', '')) * scale[u]:.3f}"   # ncu picks a unit per column
                rec[w] = v
        out.append(rec)
    return out, rows[1], hdr


def main():
    ll, rp, outp = sys.argv[1:4]
    lines = ["# ncu summary\n"]
    if len(sys.argv) > 4:
        b = json.loads(open(sys.argv[4]).read().strip().splitlines()[-1])
        lines += ["bench.py line of the same build (CUDA events, not under ncu):\n", "```json", json.dumps(b), "```\n"]
    d = launches(ll)
    tot = sum(sum(v) for v in d.values())
    lines += ["## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: compare shares)\n",
              "| kernel | launches | total us | avg us | share |", "|---|---:|---:|---:|---:|"]
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        lines.append(f"| `{k}` | {len(v)} | {sum(v):.1f} | {sum(v) / len(v):.2f} | {sum(v) / tot:.3f} |")
    recs, units, hdr = raw(rp)
    lines += ["\n## `ncu --set full` captures (per launch)\n",
              "| kernel | time us | DRAM read MB | DRAM write MB | DRAM % of ncu peak | warps active % | regs | FP64 pipe % | tensor pipe % | grid |",
              "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|"]
    for r in recs:
        g = lambda k: r.get(k, "")
        lines.append(f"| `{r['kernel']}` | {g('gpu__time_duration.sum')} | {g('dram__bytes_read.sum')} | {g('dram__bytes_write.sum')} | "
                     f"{g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')} | {g('sm__warps_active.avg.pct_of_peak_sustained_active')} | "
                     f"{g('launch__registers_per_thread')} | {g('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active')} | "
                     f"{g('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')} | {g('launch__grid_size')} |")
    open(outp, "w").write("\n".join(lines) + "\n")
    print("wrote", outp)
    # per-launch DRAM traffic (read + write, bytes) keyed by bench.py's kernel names
    alias = {"rhs_blocks": "rhs_blocks_kernel", "cg_recompute_pass": "cg_pass_kernel", "cg_fused_pass": "cg_pass_kernel",
             "cg_solve_kernel": "cg_pass_kernel",
             "chisq": "chisq_kernel", "mh_suffstat": "mh_suffstat_kernel", "mh_perpixel": "mh_perpixel_kernel"}
    acc = collections.defaultdict(list)
    for r in recs:
        try:
            t = float(r["gpu__time_duration.sum"])
            by = (float(r["dram__bytes_read.sum"]) + float(r["dram__bytes_write.sum"])) * 1e6
        except (KeyError, ValueError):
            continue
        if t < 10.0:
            continue  # passes launched after convergence return at once
        for pre, name in alias.items():
            if r["kernel"].startswith(pre):
                acc[name].append(by)
    import os
    tp = os.path.join(os.path.dirname(os.path.abspath(outp)), "ncu_traffic.json")
    old = json.load(open(tp)) if os.path.exists(tp) else {}
    old.update({k: round(sum(v) / len(v)) for k, v in acc.items()})
    json.dump(old, open(tp, "w"), indent=1, sort_keys=True)
    print("wrote", tp)


if __name__ == "__main__":
    main()
