"""Turns ncu CSV exports into the markdown tables committed next to them.

    python profiles/summarize.py launches <launches.csv[.gz]> <out.md> "<command that was profiled>"
    python profiles/summarize.py kernels  <out.md> "<workload note>" <title>=<raw.csv[.gz]> [<title>=<raw.csv[.gz]> ...]
    python profiles/summarize.py traffic  <out.json> "<workload> x <scenes>" "<ncu command>" <raw.csv[.gz]> ...

`launches` reads the list written by
    ncu --metrics gpu__time_duration.sum --clock-control none -c N --csv --log-file launches.csv <cmd>
`kernels` reads `ncu -i <rep> --page raw --csv` exports of `ncu --set full --clock-control none` captures and keeps
the first captured launch of every kernel.
"""
import csv
import gzip
import io
import re
import sys
from collections import OrderedDict, defaultdict


def _open(path):
    raw = gzip.open(path, "rt", errors="replace").read() if path.endswith(".gz") else open(path, errors="replace").read()
    # ncu log files carry ==PROF== lines around the CSV
    lines = [ln for ln in raw.splitlines() if ln.startswith('"')]
    return list(csv.reader(io.StringIO("\n".join(lines))))


def short_name(name):
    name = re.sub(r"<unnamed>::|\(anonymous namespace\)::|void ", "", name)
    m = re.match(r"([A-Za-z_0-9:]+(?:<[^(]*?>)?)\s*\(", name)
    return (m.group(1) if m else name)[:60]


def launches(path, out, command):
    rows = _open(path)
    hdr = rows[0]
    iname, ival, imet, iunit = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name"), hdr.index("Metric Unit")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[1:]:
        if len(r) <= ival or r[imet] != "gpu__time_duration.sum":
            continue
        v = float(r[ival].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iunit].strip(), 1e-6)
        k = short_name(r[iname])
        tot[k] += v
        cnt[k] += 1
    total = sum(tot.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list of `{command}`\n\n")
        f.write("Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c N --csv --log-file launches.csv "
                f"{command}`\n(cold-cache, serialised per-launch times: compare SHARES, not absolutes).  Raw list: `{path.split('/')[-1]}`.\n\n")
        f.write(f"Total captured: {sum(cnt.values())} launches, {total:.1f} ms of kernel time.\n\n")
        f.write("| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|\n")
        for k in sorted(tot, key=tot.get, reverse=True):
            if tot[k] / total < 1e-4:
                continue
            f.write(f"| `{k}` | {cnt[k]} | {tot[k]:.2f} | {100 * tot[k] / total:.2f}% | {1e3 * tot[k] / cnt[k]:.1f} |\n")
        fam = lambda pat: 100 * sum(v for k, v in tot.items() if re.search(pat, k)) / total
        f.write(f"\n`k_gemm<*>` family: {fam('k_gemm'):.1f}% of kernel time; `k_rl_*` (Cholesky sweep): {fam('k_rl_'):.1f}%; "
                f"`k_kgrad`: {fam('k_kgrad'):.1f}%; `k_build`: {fam('k_build$'):.1f}%.\n")
        f.write("Phase ids of `k_gemm<N>`: 2=A, 3=B, 5=G_A, 6=dT(+Adam), 8=G_C, 9=sym Phi(G_A A^T), 11=Y, 12=G_K.\n")


COLS = OrderedDict([
    ("time", "gpu__time_duration.sum"),
    ("grid", "launch__grid_size"),
    ("block", "launch__block_size"),
    ("regs", "launch__registers_per_thread"),
    ("dyn smem", "launch__shared_mem_per_block_dynamic"),
    ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("DMMA pipe %", "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active"),
    ("FP64 pipe %", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
    ("tensor pipe %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("L2 throughput %", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("SM throughput %", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("DRAM read", "dram__bytes_read.sum"),
    ("DRAM write", "dram__bytes_write.sum"),
    ("DRAM %", "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("L2 hit %", "lts__t_sector_hit_rate.pct"),
    ("smem bank conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
])


def kernels(out, note, sources):
    with open(out, "w") as f:
        f.write("# `ncu --set full --clock-control none` captures (one row per kernel, first captured launch)\n\n")
        f.write(note + "\n")
        for title, path in sources:
            rows = _open(path)
            hdr, units = rows[0], rows[1]
            iname = hdr.index("Kernel Name")
            stall = [(i, h.replace("smsp__pcsamp_warps_issue_stalled_", "")) for i, h in enumerate(hdr)
                     if h.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in h]
            f.write(f"\n## {title}\n\nRaw export: `{path.split('/')[-1]}`\n\n")
            f.write("| kernel | " + " | ".join(COLS) + " | top stalls (% of samples) |\n")
            f.write("|---|" + "---:|" * len(COLS) + "---|\n")
            seen = set()
            for r in rows[2:]:
                k = short_name(r[iname])
                if k in seen:
                    continue
                seen.add(k)
                cells = []
                for label, metric in COLS.items():
                    match = [i for i, h in enumerate(hdr) if h == metric or h.endswith("." + metric)]
                    if not match:
                        cells.append("n/a")
                        continue
                    i = match[0]
                    v, u = r[i], units[i]
                    try:
                        v = f"{float(v.replace(',', '')):.4g}"
                    except ValueError:
                        pass
                    cells.append(f"{v} {u}".strip() if u not in ("%", "") and label not in ("grid", "block", "regs") else v)
                st = []
                for i, n in stall:
                    try:
                        st.append((float(r[i].replace(",", "")), n))
                    except ValueError:
                        pass
                tot = sum(v for v, _ in st) or 1.0
                top = ", ".join(f"{n} {100 * v / tot:.0f}" for v, n in sorted(st, reverse=True)[:4])
                f.write(f"| `{k}` | " + " | ".join(cells) + f" | {top} |\n")


def traffic(out, workload, command, paths):
    """per-launch DRAM bytes of the first captured launch of every kernel -> json read by bench.py"""
    import json
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    res = {}
    for path in paths:
        rows = _open(path)
        hdr, units = rows[0], rows[1]
        iname, ir, iw = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        it = hdr.index("gpu__time_duration.sum")
        for r in rows[2:]:
            k = short_name(r[iname])
            if k in res:
                continue
            res[k] = {"dram_read_bytes": float(r[ir].replace(",", "")) * scale[units[ir]],
                      "dram_write_bytes": float(r[iw].replace(",", "")) * scale[units[iw]],
                      "duration": r[it] + " " + units[it]}
    with open(out, "w") as f:
        json.dump({"workload": workload, "command": command, "kernels": res}, f, indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5:])
    elif sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4])
    elif sys.argv[1] == "kernels":
        kernels(sys.argv[2], sys.argv[3], [tuple(a.split("=", 1)) for a in sys.argv[4:]])
    else:
        raise SystemExit(__doc__)
