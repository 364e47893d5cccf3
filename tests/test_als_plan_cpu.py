"""CPU: the ALS work list (mf_als_plan, host-only — csrc/als.cu als_plan): long segments are cut into equal parts that other
CTAs accumulate separately.  Integer tier: every entry of every segment is covered exactly once, part lengths follow the
rule the kernel re-derives from (degree, nparts), slots are disjoint, the list is sorted longest-first."""
import numpy as np
import pytest


def _part_len(deg, nparts):
    return -(-(-(-deg // nparts)) // 32) * 32 if nparts > 1 else deg  # roundup32(ceil(deg / nparts))


@pytest.mark.parametrize("split", [32, 64, 1000, 8192, 1 << 30])
def test_plan_covers_every_entry_once(pkg, split):
    rng = np.random.default_rng(split % 97)
    deg = np.concatenate([np.zeros(5, np.int64), rng.integers(1, 40, 300), rng.integers(40, 5000, 200),
                          np.array([31, 32, 33, 63, 64, 65, 8191, 8192, 8193, 16384, 16385, 100000, 180000])])
    rng.shuffle(deg)
    ptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.uint32)
    items, slots = pkg.als_plan(ptr, split)
    seg, part, nparts, slot = (items[:, i].astype(np.int64) for i in range(4))
    # one item per unsplit segment, nparts items per split one, parts 0..nparts-1
    assert np.array_equal(np.bincount(seg, minlength=len(deg)), np.maximum(1, np.array([nparts[seg == s][0] for s in range(len(deg))])))
    covered = np.zeros(len(deg), np.int64)
    lengths = np.zeros(len(items), np.int64)
    for i, (s, p, n, sl) in enumerate(zip(seg, part, nparts, slot)):
        d = int(deg[s])
        plen = _part_len(d, n)
        lo, hi = min(d, p * plen), min(d, (p + 1) * plen)
        assert n == 1 or (hi > lo and d > split), (s, p, n, d)   # split segments have no empty parts
        assert n == 1 or plen % 32 == 0
        covered[s] += hi - lo
        lengths[i] = hi - lo
    assert np.array_equal(covered, deg)
    # longest-first
    assert np.all(np.diff(lengths) <= 0)
    # slots of split segments: contiguous ranges [slot, slot + nparts), disjoint, dense
    used = sorted({(int(sl), int(n)) for sl, n in zip(slot, nparts) if n > 1})
    pos = 0
    for sl, n in used:
        assert sl == pos
        pos += n
    assert pos == slots
    if split >= deg.max():
        assert slots == 0 and len(items) == len(deg)


def test_plan_degenerate_inputs(pkg):
    items, slots = pkg.als_plan(np.zeros(1, np.uint32))
    assert items.shape == (0, 4) and slots == 0
    items, slots = pkg.als_plan(np.array([0, 0, 0], np.uint32))
    assert items.shape == (2, 4) and slots == 0 and set(items[:, 0]) == {0, 1} and np.all(items[:, 2] == 1)
