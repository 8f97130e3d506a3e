"""Static SASS instruction mix of one kernel of an object file (cuobjdump -sass), per opcode and per
record: the CPU-side check of instruction counts before GPU time is spent on a kernel change.

    python tools/sass_mix.py bwtb3m_b200/csrc/sufsort.o 'k_radix_onesweepILi2ELb1ELb0ELb1E' 16
"""
import re
import signal
import subprocess
import sys
from collections import Counter


def main():
    signal.signal(signal.SIGPIPE, signal.SIG_DFL)  # `| head` is fine
    obj, pat, per = sys.argv[1], sys.argv[2], float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    cur, mix, name = None, Counter(), None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            if pat in cur and name is None:
                name = cur
            continue
        if cur is None or cur != name:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m:
            mix[m.group(1)] += 1
    if name is None:
        sys.exit("no function matches " + pat)
    total = sum(mix.values())
    print("%s\n%d SASS instructions, %.1f per record (%g records per thread)" % (name, total, total / per, per))
    for op, c in mix.most_common():
        print("%-14s %5d  %6.2f per record" % (op, c, c / per))


if __name__ == "__main__":
    main()
