import sys
import numpy as np
from bwtb3m_b200 import Engine
l = int(float(sys.argv[1])); itype = sys.argv[2]; iters = int(sys.argv[3]) if len(sys.argv) > 3 else 1
rng = np.random.default_rng(2)
nb = (l + 3) // 4
data = rng.integers(0, 256, size=nb + 2, dtype=np.uint8)
data[nb] = 0; data[nb + 1] = 0
assert l % 4 == 0
e = Engine(0)
for it in range(iters):
    e.load_host(data, itype); e.build()
    i = e.info(); print(i["ms_total"], i["ms_sort"])
