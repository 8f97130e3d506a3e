"""Summaries for profiles/: (1) per-kernel shares from an ncu launch list (gpu__time_duration.sum
CSV), (2) per-launch DRAM traffic / throughput / occupancy from an `ncu --set full` report.

    python tools/ncu_summary.py launches gpurun_out/launches.csv [first_id last_id]
    python tools/ncu_summary.py full gpurun_out/prof.ncu-rep
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = name.replace("b3m::", "")
    m = re.match(r"(?:void )?(k_\w+)(<.*)?", name)
    if not m:
        return name[:60]
    base, targs = m.group(1), m.group(2) or ""
    tag = ""
    if base in ("k_scan_tile", "k_scan_reduce"):
        ops = re.findall(r"Op\w+", targs)
        lam = re.findall(r"lambda[^,>]*#(\d+)", targs)
        fn = re.findall(r"(k2_suffix_sort|k4_build_dict|node_merge|leaf_build|scan_exclusive_inplace|k8_\w+)", targs)
        tag = "<%s%s%s>" % (ops[0] if ops else "", ":" + fn[0] if fn else "", "#" + lam[-1] if lam else "")
    elif base == "k_radix_onesweep":
        m2 = re.search(r"<\(int\)(\d)>", targs)
        tag = "<%s>" % (m2.group(1) if m2 else "")
    return base + tag


def launches(path, lo=None, hi=None):
    rows = [r for r in csv.DictReader(l for l in open(path) if l.startswith('"'))]
    rows = [r for r in rows if r.get("Metric Name") == "gpu__time_duration.sum"]
    if lo is not None:
        rows = [r for r in rows if lo <= int(r["ID"]) <= hi]
    agg = OrderedDict()
    tot = 0.0
    for r in rows:
        k = short(r["Kernel Name"])
        ns = float(r["Metric Value"].replace(",", ""))
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ns
        tot += ns
    print("launches %d, total %.3f ms (cold-cache, serialised: compare shares, not absolutes)" % (len(rows), tot / 1e6))
    print("%-44s %8s %12s %8s" % ("kernel", "launches", "total_us", "share"))
    for k, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-44s %8d %12.1f %7.1f%%" % (k, c, ns / 1e3, 100 * ns / tot))


METRICS = [
    ("gpu__time_duration.sum", "dur_us", 1e-3),
    ("dram__bytes_read.sum", "dram_rd_MB", 1e-6),
    ("dram__bytes_write.sum", "dram_wr_MB", 1e-6),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct", 1),
    ("lts__t_sectors.sum", "l2_Msect", 1e-6),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_pct", 1),
    ("launch__registers_per_thread", "regs", 1),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct", 1),
]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = csv.reader(io.StringIO(out))
    hdr = next(rd)
    units = next(rd)
    idx = {h: i for i, h in enumerate(hdr)}
    print("%-40s %14s " % ("kernel", "grid") + " ".join("%10s" % m[1] for m in METRICS))
    for r in rd:
        vals = []
        for name, _, sc in METRICS:
            i = idx.get(name)
            if i is None:
                vals.append("-")
                continue
            try:
                v = float(r[i].replace(",", ""))
                u = units[i]
                if name.startswith("dram__bytes"):
                    v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
                if name == "gpu__time_duration.sum":
                    v *= {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(u, 1)
                vals.append("%.2f" % (v * sc))
            except ValueError:
                vals.append(r[i])
        print("%-40s %14s " % (short(r[idx["Kernel Name"]]), r[idx["Grid Size"]].replace(" ", "")) + " ".join("%10s" % v for v in vals))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], *(int(x) for x in sys.argv[3:5]))
    else:
        full(sys.argv[2])
