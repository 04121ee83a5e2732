"""Turns an `ncu --set full` report (read here on the CPU box with `ncu -i ... --page raw|source --csv`) and/or a
`--metrics gpu__time_duration.sum` launch list into the markdown summaries committed under profiles/.

    python tools/summarize_ncu.py rep   gpurun_out/X.ncu-rep  profiles/r01_X.md  "title"
    python tools/summarize_ncu.py list  gpurun_out/launches.csv profiles/r01_launches_X.md "title"
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "sm__cycles_elapsed.max",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def summarize_rep(rep, dst, title):
    raw = ncu_csv(rep, "raw")
    hdr, units, rows = raw[0], raw[1], raw[2:]
    name_i = hdr.index("Kernel Name")
    lines = [f"# {title}", "", f"Source: `{rep}` (`ncu --set full --clock-control none --import-source on`, one GPU, under gpurun; ",
             "numbers under the profiler are never bench values).", ""]
    for r in rows:
        lines += [f"## `{r[name_i][:110]}`", "", "| metric | value | unit |", "|---|---|---|"]
        for k in KEYS:
            if k in hdr:
                lines.append(f"| {k} | {r[hdr.index(k)]} | {units[hdr.index(k)]} |")
        st = [(h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", ""), float(r[i] or 0)) for i, h in enumerate(hdr)
              if "average_warps_issue_stalled" in h and "per_issue_active" in h]
        st = sorted(st, key=lambda x: -x[1])[:6]
        lines += ["", "Warp stall reasons (warps stalled per issue-active cycle): " + ", ".join(f"{n} {v:.2f}" for n, v in st), ""]
    src = ncu_csv(rep, "source")
    h2 = src[1]
    iS, iSrc, iE = h2.index("# Samples"), h2.index("Source"), h2.index("Instructions Executed")
    body, prev = [], -1
    for r in src[2:]:
        if len(r) <= iS:
            continue
        try:
            a = int(r[0], 16)
        except ValueError:
            continue
        if a < prev:
            break  # first kernel instance only
        body.append(r)
        prev = a
    tot = sum(int(r[iS] or 0) for r in body) or 1
    lines += ["## hottest SASS instructions of the first captured launch (stall samples)", "", "| samples | share | executed | SASS |", "|---|---|---|---|"]
    for r in sorted(body, key=lambda r: -int(r[iS] or 0))[:14]:
        lines.append(f"| {r[iS]} | {100 * int(r[iS]) / tot:.1f}% | {r[iE]} | `{r[iSrc][:100]}` |")
    open(dst, "w").write("\n".join(lines) + "\n")
    print("wrote", dst)


def summarize_list(path, dst, title):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    agg = {}
    for r in rows:
        name = r[4].split("(")[0][:70]
        t = float(r[-1])
        a = agg.setdefault(name, [0, 0.0, r[8], r[7]])
        a[0] += 1
        a[1] += t
    total = sum(a[1] for a in agg.values()) or 1
    lines = [f"# {title}", "", f"Source: `{path}` (`ncu --metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: read the SHARES).", "",
             "| kernel | launches | avg ns | share of listed time | grid | block |", "|---|---|---|---|---|---|"]
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{name}` | {a[0]} | {a[1] / a[0]:.0f} | {100 * a[1] / total:.1f}% | {a[2]} | {a[3]} |")
    open(dst, "w").write("\n".join(lines) + "\n")
    print("wrote", dst)


if __name__ == "__main__":
    mode, src, dst, title = sys.argv[1:5]
    (summarize_rep if mode == "rep" else summarize_list)(src, dst, title)
