"""numpy restatement of the PANEL layout (cuda-recommender_b200/csrc/layout.cuh, prep.cu) — the
integer-tier checker for the layout the GPU builds: pieces = (panel, segment) cuts of every segment,
panel-major storage, pieces padded to `pad` entries with idx16 = 4*panel_rows / val = 0 (indices stored as byte offsets), work items of <= chunk
entries, slots of a segment ordered (panel, chunk).  Test infrastructure only."""
import numpy as np

SMEM_MAX = 227 * 1024 - 256


def panel_cap(nvec):
    return (SMEM_MAX // (4 * nvec) - 8) // 8 * 8


def choose_panel_rows(gdim, cap):
    if gdim <= cap:
        return max(8, -(-gdim // 8) * 8)
    npan = -(-gdim // cap)
    return -(-(-(-gdim // npan)) // 8) * 8


def session_panel_rows(gdim, side_is_csr, user_panel_rows=0):
    cap = min(panel_cap(3 if side_is_csr else 2), 16376)
    want = user_panel_rows // 8 * 8 if user_panel_rows > 0 else 16376
    return choose_panel_rows(gdim, max(min(cap, want), 8))


def session_pad(nseg, nnz, gdim, panel_rows):
    """The padding granularity a session picks (session.cu): 8 for short-piece copies (fewer than 24 entries per piece on
    average: the sweeps then run one item per lane), else 32."""
    npan = -(-gdim // panel_rows)
    pieces = max(1, min(nseg * npan, max(nnz, 1)))
    return 8 if nnz // pieces < 24 else 32


def panel_layout(ptr, idx, val, gdim, panel_rows, chunk, pad=32):
    if pad < 8 or pad % 8 or chunk % pad:
        pad = 8
    ptr = ptr.astype(np.int64)
    nseg = len(ptr) - 1
    P = max(1, -(-gdim // panel_rows))
    idx16, pval, items, item_panel = [], [], [], []
    seg_items = np.zeros(nseg, np.int64)
    pieces = []
    for p in range(P):
        lo_key, hi_key = p * panel_rows, (p + 1) * panel_rows
        for s in range(nseg):
            seg = idx[ptr[s]:ptr[s + 1]]
            a = ptr[s] + np.searchsorted(seg, lo_key, "left")
            b = ptr[s] + np.searchsorted(seg, hi_key, "left")
            cnt = b - a
            pd = -(-cnt // pad) * pad
            nit = -(-pd // chunk)
            pieces.append((p, s, a, cnt, pd, nit))
            seg_items[s] += nit
    slot_ptr = np.concatenate([[0], np.cumsum(seg_items)])
    before = np.zeros(nseg, np.int64)
    pos = 0
    for p, s, a, cnt, pd, nit in pieces:
        if pd == 0:
            continue
        loc = (idx[a:a + cnt].astype(np.int64) - p * panel_rows)
        idx16.append((np.concatenate([loc, np.full(pd - cnt, panel_rows)]) * 4).astype(np.uint16))  # stored as byte offsets
        pval.append(np.concatenate([val[a:a + cnt], np.zeros(pd - cnt, np.float32)]).astype(np.float32))
        for j in range(nit):
            items.append((pos + j * chunk, min(chunk, pd - j * chunk), s, slot_ptr[s] + before[s] + j))
            item_panel.append(p)
        before[s] += nit
        pos += pd
    return dict(n_panels=P, n_padded=pos, n_items=len(items),
                idx16=np.concatenate(idx16) if idx16 else np.zeros(0, np.uint16),
                val=np.concatenate(pval) if pval else np.zeros(0, np.float32),
                items=np.array(items, np.uint32).reshape(-1, 4), item_panel=np.array(item_panel, np.int64))


def check_layout(got, want):
    """The GPU lists the same work items, re-ordered inside every panel (length-ranked, then dealt into lanes; the order
    among equal lengths is arbitrary), and stores the padded entries in that WORK-LIST order (STREAM order, layout.cuh):
    item i starts where items 0..i-1 end.  Bit-exact check: same set of items per panel (keyed by their slot), every item
    holds exactly the reference's padded entries, starts are the exclusive prefix sums of the lengths."""
    gi, wi = got["items"], want["items"]
    assert gi.shape == wi.shape
    assert got["n_padded"] == want["n_padded"] and len(got["idx16"]) >= want["n_padded"]
    if len(wi) == 0:
        return
    lens = gi[:, 1].astype(np.int64)
    assert np.array_equal(gi[:, 0].astype(np.int64), np.concatenate([[0], np.cumsum(lens)[:-1]]))
    assert int(lens.sum()) == want["n_padded"]
    by_slot = {int(r[3]): r for r in wi}
    assert len(by_slot) == len(wi)
    for g in gi:
        w = by_slot[int(g[3])]
        assert int(g[1]) == int(w[1]) and int(g[2]) == int(w[2])
        a, b, n = int(g[0]), int(w[0]), int(g[1])
        assert np.array_equal(got["idx16"][a:a + n], want["idx16"][b:b + n])
        assert np.array_equal(got["val"][a:a + n], want["val"][b:b + n])
    counts = np.bincount(want["item_panel"], minlength=want["n_panels"])
    lo = 0
    for c in counts:
        assert set(gi[lo:lo + c, 3].tolist()) == set(wi[lo:lo + c, 3].tolist())  # same panel membership
        lo += c
