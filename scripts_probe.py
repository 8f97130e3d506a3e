import sys, time
import numpy as np
from bwtb3m_b200 import Engine
l = int(float(sys.argv[1])) if len(sys.argv) > 1 else 48_000_000
itype = sys.argv[2] if len(sys.argv) > 2 else "pacterm"
rng = np.random.default_rng(2)
if itype in ("pac", "pacterm"):
    nb = (l + 3) // 4
    data = rng.integers(0, 256, size=nb + 2, dtype=np.uint8)
    if l % 4 == 0:
        data[nb] = 0; data = data[: nb + 2]; data[nb + 1] = 0
    else:
        data = data[: nb + 1]; data[nb] = l % 4
else:
    sig = int(sys.argv[3]) if len(sys.argv) > 3 else 256
    data = rng.integers(0, sig, size=l, dtype=np.uint8)
e = Engine(0)
for it in range(3):
    t0 = time.time(); e.load_host(data, itype); t1 = time.time()
    e.build(); t2 = time.time()
    i = e.info()
    print("load %.1f ms build %.1f ms | dec %.2f sort %.2f ext %.2f dict %.2f walk %.2f total %.2f | rounds %d passes %d active_sum %d launches %d prerate %d" % (
        (t1 - t0) * 1e3, (t2 - t1) * 1e3, i["ms_decode"], i["ms_sort"], i["ms_extract"], i["ms_dict"], i["ms_walk"], i["ms_total"],
        i["sort_rounds"], i["radix_passes"], i["sort_active_sum"], i["launches"], i["preisarate"]), flush=True)
    print("  Mbp/s device: %.1f ; radix GB/s: %.1f ; walk Gsteps/s %.2f" % (i["n"] / i["ms_total"] / 1e3, i["radix_bytes"] / i["ms_sort"] / 1e6, i["walk_lf_steps"] / i["ms_walk"] / 1e6))
for nch in (1 << 14, 1 << 17, 1 << 19, 1 << 21):
    ms, cs = e.lf_bench(nch, 256)
    print("lfbench chains %d: %.3f ms -> %.2f Gsteps/s" % (nch, ms, nch * 256 / ms / 1e6))
