"""TMEM column layout of the split-softmax attention kernel for 256 < seq <= 448 (attention.cu:
window_attention_tc2_kernel), replayed on the CPU: for every window length the two warps of a lane quarter must never
overwrite a score column that either of them still has to read, P must not touch O, and the P V k-steps must address
the columns where the keys' probabilities were packed.  Mirrors p_col_split / softmax_half / issue_pv_split.  (The same
layout rule was measured on the 256-column ping-pong slots too - slower there, removed; the replay still covers it.)"""
import pytest


def p_col_split(c, c0, n):
    return 16 * c if c < c0 else 16 * (n + c)


@pytest.mark.parametrize("seq", list(range(1, 449)))
def test_split_layout_is_hazard_free(seq):
    """seq <= 256: ping-pong kernel (256-column slots); 256 < seq <= 448: single-slot kernel (512 columns, O last)."""
    spad = (seq + 15) // 16 * 16
    n = (seq + 31) // 32
    c0 = (n + 1) // 2
    slot = 256 if seq <= 256 else 512
    deferred = spad <= 192 or slot == 512                      # O disjoint from the scores
    o_cols = set(range(slot - 64, slot)) if deferred else set(range(64, 128))
    s_cols = set(range(spad))                                  # written by S = Q K^T
    chunks = {0: list(range(0, c0)), 1: list(range(n - 1, c0 - 1, -1))}     # exp-pass order of each half
    region = {h: set(col for c in chunks[h] for col in range(32 * c, 32 * c + 32)) for h in (0, 1)}
    assert region[0].isdisjoint(region[1]) and s_cols <= (region[0] | region[1])
    p_cols = {}
    for h in (0, 1):
        unread = set(region[h])                                # both halves finished their max pass before any P write
        for c in chunks[h]:
            unread -= set(range(32 * c, 32 * c + 32))          # chunk c is in registers now
            w = set(range(p_col_split(c, c0, n), p_col_split(c, c0, n) + 16))
            assert w.isdisjoint(unread), "P chunk %d overwrites unread scores of its own half" % c
            assert w.isdisjoint(region[1 - h]), "P chunk %d lands in the other half's score range" % c
            assert max(w) < slot and w.isdisjoint(o_cols), "P chunk %d outside the slot or on O" % c
            for col in w:
                assert col not in p_cols, "two P chunks share column %d" % col
                p_cols[col] = (c, col - p_col_split(c, c0, n))
    if deferred:
        assert o_cols.isdisjoint(s_cols)                       # S(u+2) may be issued while O(u) is still unread
    # P V: k-step kk covers keys [16 kk, 16 kk + 16) = 8 packed columns
    for kk in range(spad // 16):
        base = p_col_split(kk // 2, c0, n) + 8 * (kk % 2)
        for j in range(8):
            c, off = p_cols[base + j]
            key0 = 32 * c + 2 * off                            # column `off` of chunk c packs keys (2 off, 2 off + 1)
            assert key0 == 16 * kk + 2 * j


# ---------------------------------------------------------------------------------------------------------------------
# Ping-pong kernel (seq <= 256, attention.cu: window_attention_pp_kernel / row_chunks / row_exp): one warp per lane quarter
# and slot walks the 32-column chunks upwards with the load of the NEXT chunk in flight, packs P chunk c at columns
# [16 c, 16 c + 16); a last chunk of <= 16 keys is read with a 16-column load AFTER the others and its P stored as 8
# columns; O sits at columns [192, 256) of the 256-column slot and S(u) waits for the epilogue of unit u - 2 iff its key
# columns reach them.
@pytest.mark.parametrize("kv", list(range(1, 257)))
def test_pp_layout_is_hazard_free(kv):
    spad = (kv + 15) // 16 * 16
    n = (kv + 31) // 32
    tl = kv - 32 * (n - 1)
    narrow = tl <= 16
    wide = n - 1 if narrow else n
    order = list(range(wide)) + ([n - 1] if narrow else [])       # exp-pass order
    width = lambda c: 16 if (narrow and c == n - 1) else 32
    unread = set(col for c in order for col in range(32 * c, 32 * c + width(c)))
    assert set(range(spad)) <= unread                              # every score column S = Q K^T wrote is read
    p_cols = {}
    for j, c in enumerate(order):
        unread -= set(range(32 * c, 32 * c + width(c)))            # chunk c is in registers
        pw = 8 if width(c) == 16 else 16
        w = set(range(16 * c, 16 * c + pw))
        # (the load of the next chunk is in flight when P(c) is stored: it must not be hit either - it is still `unread`)
        assert w.isdisjoint(unread), "P chunk %d overwrites unread scores" % c
        assert max(w) < 192 or spad > 192                          # P never reaches the O columns unless S does too
        for col in w:
            assert col not in p_cols
            p_cols[col] = (c, col - 16 * c)
    assert max(p_cols) < 128                                       # P (<= 256 keys) stays below the O columns in any case
    for kk in range(spad // 16):                                   # P V: k-step kk = keys [16 kk, 16 kk + 16) = columns [8 kk, 8 kk + 8)
        for j in range(8):
            c, off = p_cols[8 * kk + j]
            assert 32 * c + 2 * off == 16 * kk + 2 * j


# TMA boxes of one of Q / K / V of an item (attention.cu: pp_row_bytes / pp_load_rows): all seq rows -> one box of
# ceil16(seq) rows; fewer -> 64-row boxes, then 16-row boxes (or one 64-row box when the tail needs more than 48 rows).
def pp_boxes(rows, seq):
    if rows == seq:
        return [(0, (seq + 15) // 16 * 16)]
    nfull, rem = rows // 64, rows % 64
    boxes = [(64 * b, 64) for b in range(nfull)]
    r = 64 * nfull
    if rem > 48:
        boxes.append((r, 64))
    else:
        while r < rows:
            boxes.append((r, 16))
            r += 16
    return boxes


def pp_row_bytes(rows, seq):
    if rows == seq:
        return (seq + 15) // 16 * 16 * 128
    nfull, rem = rows // 64, rows % 64
    if rem > 48:
        return (nfull + 1) * 8192
    return nfull * 8192 + (rem + 15) // 16 * 2048


@pytest.mark.parametrize("seq", [1, 16, 33, 103, 129, 180, 192, 201, 256])
def test_pp_tma_boxes_cover_the_rows_and_stay_inside_the_item(seq):
    region = (seq + 15) // 16 * 16                                # rows of K / V in an item buffer (Q: whole 128-row tiles)
    for rows in range(1, seq + 1):
        boxes = pp_boxes(rows, seq)
        assert sum(n for _, n in boxes) * 128 == pp_row_bytes(rows, seq)      # expect_tx == bytes that arrive
        covered = set(r for r0, n in boxes for r in range(r0, r0 + n))
        assert set(range(rows)) <= covered
        assert max(covered) < region
        assert all(r0 % 16 == 0 for r0, _ in boxes)                # 2 KB-aligned destinations keep the 128-byte swizzle phase
        assert len(boxes) <= 6
