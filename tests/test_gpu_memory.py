"""GPU: the device memory an engine holds follows the block plan and the sorter, not a fixed bytes-per-suffix
figure fixed at load time (b3m_info.arena_capacity / arena_peak)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _capacity(pac, **kw):
    from bwtb3m_b200 import Engine
    e = Engine(0)
    try:
        e.load_host(pac, "pacterm")
        loaded = e.info()["arena_capacity"]
        e.build(sasamplingrate=32, isasamplingrate=1024, **kw)
        i = e.info()
        return loaded, i["arena_capacity"], i["arena_peak"], e.fetch()
    finally:
        e.close()


def test_working_set_follows_block_plan():
    from bwtb3m_b200 import workloads
    n = 16_000_000
    pac = workloads.random_pac(n, 5)
    l1, c1, p1, r1 = _capacity(pac, numblocks=1, sortpath="lsd")
    l8, c8, p8, r8 = _capacity(pac, numblocks=8)
    lm, cm, pm, rm = _capacity(pac, numblocks=1)
    # after load only the text is resident: file + byte codes + packed text
    assert l1 == l8 == lm and l1 < 3 * n
    # eight blocks / the MSD sorter need far less than one block under the LSD sorter ...
    assert c8 < 0.7 * c1 and cm < 0.7 * c1
    # ... and what was reserved was enough: nothing grew during the build
    assert p1 <= c1 and p8 <= c8 and pm <= cm
    for k in ("bwt", "sa", "isa"):
        assert np.array_equal(r1[k], r8[k]) and np.array_equal(r1[k], rm[k])
