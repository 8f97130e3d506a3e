"""Synthetic inputs of BASELINE.json's configs: fixed seeds, no network.

Every generator returns the bytes of the input FILE (what `bwtb3m <file>` would read)."""
import numpy as np

# name -> (inputtype, number of symbols, seed, description)
CONFIGS = {
    "cfg1": ("bytestream", 8_000_000, 1, "8 Mbp random ACGT bytestream, bwtonly=1"),
    "cfg2": ("pacterm", 48_000_000, 2, "48 Mbp chr21-sized DNA as BWA pacterm, sasamplingrate=32, full sampled SA/ISA"),
    "cfg3": ("pacterm", 3_100_000_000, 3, "3.1 Gbp GRCh38-sized pacterm genome"),
    "cfg4": ("pacterm", 3_200_000_000, 4, "64 mutated copies of a 50 Mbp genome (large-LCP path)"),
    "cfg5": ("bytestream", 1_000_000_000, 5, "1 GB random byte-alphabet bytestream"),
}


def pac_file_from_packed(packed, l):
    """packed: uint8 array with ceil(l/4) bytes (2 bit/base, MSB first).  Appends BWA's trailer:
    a zero byte when l % 4 == 0, then the count byte l % 4."""
    nb = (l + 3) // 4
    out = np.empty(nb + (2 if l % 4 == 0 else 1), dtype=np.uint8)
    out[:nb] = packed[:nb]
    if l % 4:
        # clear the unused low bits of the last byte
        keep = (0xFF << (2 * (4 - l % 4))) & 0xFF
        out[nb - 1] &= keep
        out[nb] = l % 4
    else:
        out[nb] = 0
        out[nb + 1] = 0
    return out


def random_pac(l, seed):
    """iid uniform 2-bit bases: a random byte holds four of them."""
    rng = np.random.default_rng(seed)
    nb = (l + 3) // 4
    packed = rng.integers(0, 256, size=nb, dtype=np.uint8)
    return pac_file_from_packed(packed, l)


def pack_bases(bases):
    """bases: uint8 array of values 0..3 -> packed pac payload."""
    l = bases.size
    pad = (-l) % 4
    b = np.concatenate([bases, np.zeros(pad, dtype=np.uint8)]) if pad else bases
    b = b.reshape(-1, 4)
    return ((b[:, 0] << 6) | (b[:, 1] << 4) | (b[:, 2] << 2) | b[:, 3]).astype(np.uint8)


def repetitive_pac(copies, base_len, seed, rate=1e-3):
    """`copies` mutated copies of one random genome: copy c gets independent substitutions at
    the given rate (seed*64+c), SURVEY 8d cfg4."""
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 4, size=base_len, dtype=np.uint8)
    parts = []
    for c in range(copies):
        r = np.random.default_rng(seed * 64 + c)
        x = base.copy()
        k = r.binomial(base_len, rate)
        pos = r.integers(0, base_len, size=k)
        x[pos] = (x[pos] + r.integers(1, 4, size=k).astype(np.uint8)) & 3
        parts.append(x)
    allb = np.concatenate(parts)
    return pac_file_from_packed(pack_bases(allb), allb.size)


def make(name, scale=1.0):
    """Returns (inputtype, file bytes as uint8 array, number of input symbols excluding the terminator)."""
    itype, n, seed, _ = CONFIGS[name]
    n = max(1, int(n * scale))
    if name == "cfg1":
        rng = np.random.default_rng(seed)
        return itype, np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=n)], n
    if name in ("cfg2", "cfg3"):
        return itype, random_pac(n, seed), n
    if name == "cfg4":
        base_len = max(1, n // 64)
        data = repetitive_pac(64, base_len, seed)
        return itype, data, base_len * 64
    if name == "cfg5":
        rng = np.random.default_rng(seed)
        return itype, rng.integers(0, 256, size=n, dtype=np.uint8), n
    raise KeyError(name)
