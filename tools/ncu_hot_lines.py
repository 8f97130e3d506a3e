"""Print the hottest SASS lines (by warp-stall samples) of a kernel from an .ncu-rep."""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if "Source" in r][0]
hdr = rows[hi]
si, ai = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_")]
body = []
for r in rows[hi + 1:]:
    if len(r) != len(hdr) or r[0] == "Address":
        break
    body.append(r)
tot = sum(int(r[ai]) for r in body)
print("kernel", kern, "instructions", len(body), "samples", tot)
idx = sorted(range(len(body)), key=lambda i: -int(body[i][ai]))[:top]
for i in sorted(idx):
    r = body[i]
    st = sorted(((int(r[c]), hdr[c]) for c in stall_cols if r[c].isdigit() and int(r[c])), reverse=True)[:2]
    print("%5d %5.1f%%  %-70s %s" % (i, 100 * int(r[ai]) / max(tot, 1), r[si].strip()[:70], st))
