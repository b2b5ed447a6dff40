"""Host-side check of the stream-K attention schedule (gtsfm_b200/csrc/attn_ps.cuh `attn_ps_decode`, linear.cuh
`run_flash2`): the arithmetic that cuts the (item, key tile) space into per-CTA ranges is restated here and its invariants
are checked over many shapes - every unit is covered exactly once, an item's partial slots are numbered 0..nsplits-1
without gaps, and nsplits never exceeds the `max_splits` the host sizes the partial buffers with."""
import itertools
import random


def cdiv(a, b):
    return (a + b - 1) // b


def plan(nq0, nk0, nq1, nk1, sm_count):
    """mirror of run_flash2's persistent branch"""
    p = []
    for nq, nk in ((nq0, nk0), (nq1, nk1)):
        qt = cdiv(nq, 256) if nq > 0 and nk > 0 else 0
        p.append({"qt": qt, "tiles": cdiv(nk, 64) if nk > 0 else 1})
    W0 = p[0]["qt"] * 4 * p[0]["tiles"]
    W = W0 + p[1]["qt"] * 4 * p[1]["tiles"]
    if W <= 0:
        return None
    ncta = min(W, sm_count)
    quota = cdiv(W, ncta)
    ncta = cdiv(W, quota)
    max_splits = cdiv(max(p[0]["tiles"], p[1]["tiles"]), quota) + 1
    return p, W0, W, quota, ncta, max_splits


def decode(p, W0, quota, w, w_end):
    """mirror of attn_ps_decode"""
    z = 1 if w >= W0 else 0
    base = W0 if z else 0
    wl = w - base
    tiles, qt = p[z]["tiles"], p[z]["qt"]
    item = wl // tiles
    tile0 = wl - item * tiles
    T = min(tiles - tile0, w_end - w)
    wi0 = base + item * tiles
    wi1 = wi0 + tiles
    c_first = wi0 // quota
    return {"z": z, "item": item, "h": item // qt, "q0": (item - (item // qt) * qt) * 256, "tile0": tile0, "T": T,
            "split": w // quota - c_first, "nsplits": (wi1 - 1) // quota - c_first + 1,
            "itemg": (p[0]["qt"] * 4 if z else 0) + item}


def check(nq0, nk0, nq1, nk1, sm):
    pl = plan(nq0, nk0, nq1, nk1, sm)
    if pl is None:
        return
    p, W0, W, quota, ncta, max_splits = pl
    assert ncta <= sm and ncta * quota >= W and (ncta - 1) * quota < W
    covered = {}
    splits = {}
    for c in range(ncta):
        w, w_end = c * quota, min(W, (c + 1) * quota)
        assert w < w_end, "no empty CTA"
        while w < w_end:
            s = decode(p, W0, quota, w, w_end)
            assert s["T"] >= 1 and 0 <= s["h"] < 4 and 0 <= s["split"] < s["nsplits"] <= max_splits
            for t in range(s["tile0"], s["tile0"] + s["T"]):
                key = (s["itemg"], t)
                assert key not in covered
                covered[key] = c
            splits.setdefault(s["itemg"], []).append((s["split"], s["nsplits"]))
            w += s["T"]
    n_items = (p[0]["qt"] + p[1]["qt"]) * 4
    assert len(covered) == W and len(splits) == n_items
    for itemg, lst in splits.items():
        ns = lst[0][1]
        assert sorted(x[0] for x in lst) == list(range(ns)) and all(x[1] == ns for x in lst)


def test_stream_k_schedule_invariants():
    rnd = random.Random(0)
    shapes = [(5000, 5000, 5000, 5000), (5000, 4321, 4321, 5000), (1, 1, 1, 1), (257, 63, 64, 65), (2048, 100, 100, 2048), (300, 0, 0, 300)]
    shapes += [tuple(rnd.randint(1, 6000) for _ in range(4)) for _ in range(60)]
    for (a, b, c, d), sm in itertools.product(shapes, (148, 140, 132, 1)):
        check(a, b, c, d, sm)
