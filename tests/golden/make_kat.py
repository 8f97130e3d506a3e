#!/usr/bin/env python
"""Generates tests/golden/kat.json: known-answer vectors for the circular BWT / SA / ISA.

The reference ships no golden files and cannot run here (libmaus2 absent), so the vectors are
derived from the reference's own executable definition in lcpbit's self-test,
    BWT[i] = s[(SA[i]+n-1)%n], ISA[SA[i]] = i      (/root/reference/src/lcpbit.cpp:3658-3669)
by a naive rotation sort written here (pure Python, independent of oracle/ and of the CUDA path).
Texts: "abbab#" is the reference's own test string (/root/reference/src/lcpbit.cpp:4053); the
others and the pacterm row are the vectors of SURVEY.md section 4.  The .pac bytes follow BWA's
fa2pac layout (2 bit/base MSB first, pad byte when l%4==0, count byte l%4).

    python tests/golden/make_kat.py   # rewrites kat.json next to this file
"""
import json
import os


def rotation_sa(t):
    n = len(t)
    return sorted(range(n), key=lambda i: t[i:] + t[:i])


def vectors(t):
    n = len(t)
    sa = rotation_sa(t)
    bwt = [t[(s + n - 1) % n] for s in sa]
    isa = [0] * n
    for r, s in enumerate(sa):
        isa[s] = r
    return sa, bwt, isa


def pac_bytes(bases):
    l = len(bases)
    out = bytearray((l + 3) // 4)
    for i, b in enumerate(bases):
        out[i >> 2] |= b << ((~i & 3) << 1)
    if l % 4 == 0:
        out.append(0)
    out.append(l % 4)
    return bytes(out)


def main():
    kats = []
    for s in ("abbab#", "banana", "mississippi", "ACGTACGTTGCA"):
        t = list(s.encode())
        sa, bwt, isa = vectors(t)
        kats.append({"name": s, "text": t, "sa": sa, "bwt": bwt, "isa": isa})
    bases = [0, 1, 2, 3, 3, 2, 1, 0]  # ACGTTGCA
    t = [b + 1 for b in bases] + [0]
    sa, bwt, isa = vectors(t)
    kats.append({"name": "pacterm:ACGTTGCA", "pac_hex": pac_bytes(bases).hex(), "text": t, "sa": sa, "bwt": bwt, "isa": isa,
                 "bwa_primary": isa[0]})
    # exhaustive small alphabets in the pattern of lcpbit.cpp:3777-3794: every string of length 4
    # over {0,1} followed by a unique larger last symbol
    ex = []
    for v in range(16):
        t = [(v >> k) & 1 for k in range(4)] + [2]
        sa, bwt, isa = vectors(t)
        ex.append({"text": t, "sa": sa, "bwt": bwt})
    out = {"source": "tests/golden/make_kat.py (naive rotation sort; definition: /root/reference/src/lcpbit.cpp:3658-3669)",
           "kats": kats, "exhaustive_2x4": ex}
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "kat.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
