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
