"""Per-source-line instruction counts and stall samples of one kernel: joins the SASS page of an
.ncu-rep with the line table of the cubin (nvdisasm -g) by instruction order.

    python tools/ncu_src_lines.py gpurun_out/prof.ncu-rep k_msd_finish bwtb3m_b200/libb3m.so sufsort [min_pct [mangled-substring]]
"""
import csv, os, re, subprocess, sys, tempfile

rep, kern, so, unit = sys.argv[1:5]
minpct = float(sys.argv[5]) if len(sys.argv) > 5 else 0.5
mangled = sys.argv[6] if len(sys.argv) > 6 else kern  # substring of the mangled name when the kernel is a template
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.startswith(unit)][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
# instruction -> (file, line), per function in order
lines, cur, infn = [], None, False
fn_re = re.compile(r"^\s*\.text\.(\S+):")
target = None
for l in dis.splitlines():
    m = fn_re.match(l)
    if m:
        infn = mangled in m.group(1) and (target is None or m.group(1) == target)
        if infn and target is None:
            target = m.group(1)
        continue
    if not infn:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if "Source" in r][0]
hdr = rows[hi]
ie, ss = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
body = []
for r in rows[hi + 1:]:
    if len(r) != len(hdr) or r[0] == "Address":
        break
    body.append(r)
print("function", target, "sass instructions", len(body), "line entries", len(lines))
agg = {}
for i, r in enumerate(body):
    k = lines[i] if i < len(lines) else None
    a = agg.setdefault(k, [0, 0])
    a[0] += int(r[ie] or 0)
    a[1] += int(r[ss] or 0)
ti, ts = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
print("warp instructions %d, samples %d" % (ti, ts))
src = {}
for k, a in sorted(agg.items(), key=lambda kv: (kv[0] or ("", 0))):
    if a[0] >= ti * minpct / 100 or a[1] >= ts * minpct / 100:
        text = ""
        if k:
            for d in ("bwtb3m_b200/csrc", "include"):
                p = os.path.join(d, k[0])
                if os.path.exists(p):
                    src.setdefault(p, open(p).read().splitlines())
                    text = src[p][k[1] - 1].strip()[:100] if k[1] <= len(src[p]) else ""
        print("%-14s %5s  inst %5.1f%%  stall %5.1f%%  %s" % (k[0] if k else "?", k[1] if k else "", 100 * a[0] / ti, 100 * a[1] / max(ts, 1), text))
