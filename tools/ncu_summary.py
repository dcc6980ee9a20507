#!/usr/bin/env python3
"""Compact, committable summaries of the ncu captures made by tools/ncu_capture.sh.

  python tools/ncu_summary.py gpurun_out profiles/r02_ncu

* launches.csv            -> launch_list.csv (kernel, grid, block, microseconds) + launch_shares.txt
* <name>.source.csv       -> <name>.sass_mix.txt : executed warp instructions by opcode, stall samples by reason,
                             and the hottest SASS lines (the full source page is several MB; this keeps what is read)
* <name>.raw.csv          -> <name>.key_metrics.txt
Details / raw files are copied as they are (they are small).
"""
import collections
import csv
import os
import re
import shutil
import sys


def short_kernel(name):
    m = re.search(r"(\w+_kernel|\w+)(<[^(]*>)?\(", name)
    k = m.group(1) if m else name[:40]
    t = ""
    if "Fq2" in name or "Fq2Params" in name:
        t = "<G2>"
    elif "Curve<" in name:
        t = "<G2>" if "Field2" in name or "Fq2" in name else "<G1>"
    m2 = re.search(r"_kernel<([0-9, ]+)>", name)
    if m2:
        t = "<%s>" % m2.group(1).replace(" ", "")
    return k + t


def launches(src, dst):
    path = os.path.join(src, "launches.csv")
    if not os.path.exists(path):
        return
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.DictReader(lines)
    for r in rd:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        rows.append((int(r["ID"]), short_kernel(r["Kernel Name"]), r["Grid Size"], r["Block Size"], r["Stream"], us))
    with open(os.path.join(dst, "launch_list.csv"), "w") as f:
        f.write("id,kernel,grid,block,stream,us\n")
        for r in rows:
            f.write('%d,%s,"%s","%s",%s,%.2f\n' % r)
    tot = sum(r[5] for r in rows)
    agg = collections.OrderedDict()
    for r in rows:
        a = agg.setdefault(r[1], [0, 0.0])
        a[0] += 1
        a[1] += r[5]
    with open(os.path.join(dst, "launch_shares.txt"), "w") as f:
        f.write("ncu --metrics gpu__time_duration.sum --clock-control none: %d launches, %.1f us serialised "
                "(cold-cache, one kernel at a time: shares, not absolutes)\n" % (len(rows), tot))
        for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-40s %3d launches %10.1f us  %5.1f %%\n" % (k, n, us, 100.0 * us / tot))


def sass_mix(src, dst, name):
    path = os.path.join(src, name + ".source.csv")
    if not os.path.exists(path):
        return
    with open(path) as f:
        first = f.readline()
        rd = csv.DictReader(f)
        rows = list(rd)
    ops = collections.Counter()
    stall_cols = [c for c in rows[0].keys() if c.startswith("stall_") and "Not Issued" not in c]
    stalls = collections.Counter()
    tot_inst = 0
    for r in rows:
        try:
            n = int(r["Instructions Executed"])
        except (ValueError, KeyError, TypeError):
            continue
        src_line = r["Source"].strip()
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src_line)
        op = m.group(2) if m else "?"
        base = ".".join(op.split(".")[:3]) if op.startswith("IMAD") else op.split(".")[0]
        ops[base] += n
        tot_inst += n
        for c in stall_cols:
            try:
                stalls[c] += int(r[c])
            except (ValueError, TypeError):
                pass
    with open(os.path.join(dst, name + ".sass_mix.txt"), "w") as f:
        f.write(first.strip()[:300] + "\n")
        f.write("SASS lines %d, executed warp instructions %d\n\n" % (len(rows), tot_inst))
        f.write("executed warp instructions by opcode\n")
        for op, n in ops.most_common(24):
            f.write("  %-22s %14d  %5.1f %%\n" % (op, n, 100.0 * n / max(tot_inst, 1)))
        st = sum(stalls.values())
        f.write("\nwarp stall samples by reason (all samples)\n")
        for c, n in stalls.most_common(10):
            f.write("  %-24s %10d  %5.1f %%\n" % (c, n, 100.0 * n / max(st, 1)))
        f.write("\nhottest SASS lines by samples\n")
        hot = sorted([r for r in rows if (r.get("# Samples") or "").isdigit()], key=lambda r: -int(r["# Samples"]))[:40]
        for r in hot:
            f.write("  %8s samples  %12s exec  %s\n" % (r.get("# Samples"), r.get("Instructions Executed"), r["Source"].strip()[:110]))


KEYS = ("gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fmaheavy", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__average_warps_issue_stalled",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed")


def key_metrics(src, dst, name):
    path = os.path.join(src, name + ".raw.csv")
    if not os.path.exists(path):
        return
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    with open(os.path.join(dst, name + ".key_metrics.txt"), "w") as f:
        for vals in rows[2:]:
            kn = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""
            f.write("== %s\n" % kn[:160])
            for h, u, v in zip(hdr, units, vals):
                if any(h.startswith(k) for k in KEYS) and "not_issued" not in h and "per_warp_active" not in h:
                    f.write("  %-96s %s %s\n" % (h, v, u))


def main():
    src, dst = sys.argv[1], sys.argv[2]
    os.makedirs(dst, exist_ok=True)
    launches(src, dst)
    names = sorted({f[:-len(".raw.csv")] for f in os.listdir(src) if f.endswith(".raw.csv")})
    for n in names:
        sass_mix(src, dst, n)
        key_metrics(src, dst, n)
        for ext in (".details.txt",):
            p = os.path.join(src, n + ext)
            if os.path.exists(p):
                shutil.copy(p, os.path.join(dst, n + ext))
    for extra in ("ncu_plain.json",):
        p = os.path.join(src, extra)
        if os.path.exists(p):
            shutil.copy(p, os.path.join(dst, "plain_run_before_ncu.json"))


if __name__ == "__main__":
    main()
