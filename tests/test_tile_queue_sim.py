"""Lock-step simulation of the warp work queue of k_tile_walk (csrc/gsi_kernels.cu): the control
flow that flattens a tile's runs into a dense position list, hands positions to idle lanes and
keeps lanes with unfinished walks busy across list refills.  A transliteration of the kernel's
loop with the walk replaced by a coin (accept / move / defer); checks what the kernel must
guarantee whatever the geometry: it terminates, and every binned position of the tile is handed
out exactly once.  (A GPU hang costs a whole box lease; this runs in milliseconds.)"""
import numpy as np
import pytest

SPOS = 256


RING = 64


def simulate(runs, S, ranges, rng_id, blocks_of_warp, rng, p_accept=0.75, p_defer=0.01, p_out=0.06, max_steps=8, guard=10**6):
    """runs: (n_chunks, 2) array of (start, end) per chunk for one tile; blocks_of_warp: the 32-chunk
    blocks the shared counter hands to this warp, in order.  Returns the positions this warp
    finished (accepted, deferred or walked to the end by the straggler path), in order."""
    n_chunks = len(runs)
    n_blocks = (n_chunks + 31) >> 5
    lanes = np.arange(32)
    grabs = iter(list(blocks_of_warp) + [10**9] * 4)
    cb_nxt = next(grabs)
    r_pos0 = np.zeros(32, dtype=np.int64); r_len = np.zeros(32, dtype=np.int64); r_excl = np.zeros(32, dtype=np.int64)
    b_total = win_lo = win_n = nxt = 0
    cur_v = np.zeros(32, dtype=bool); cur_pos = np.zeros(32, dtype=np.int64); cur_steps = np.zeros(32, dtype=np.int64)
    nxt_v = np.zeros(32, dtype=bool); nxt_pos = np.zeros(32, dtype=np.int64)      # coordinates on their way (one refill ahead)
    out_of_tile = np.zeros(32, dtype=bool)
    spos = np.full(SPOS, -1, dtype=np.int64)
    ring = np.full(RING, -1, dtype=np.int64)
    sg_head = sg_count = 0
    done = []

    def run_stragglers(n):
        nonlocal sg_head, sg_count
        for l in range(n):
            e = (sg_head + l) & (RING - 1)
            assert ring[e] >= 0
            done.append(int(ring[e]))
            ring[e] = -1
        sg_head = (sg_head + n) & (RING - 1)
        sg_count -= n

    it = 0
    while True:
        it += 1
        assert it < guard, "work queue does not terminate"
        if nxt >= win_n:
            have = True
            win_lo += win_n
            if win_lo < b_total:
                pass
            elif cb_nxt < n_blocks:
                c = (cb_nxt << 5) + lanes
                ok = c < n_chunks
                start = np.where(ok, runs[np.minimum(c, n_chunks - 1), 0], 0)
                end = np.where(ok, runs[np.minimum(c, n_chunks - 1), 1], 0)
                r_len = end - start
                r_pos0 = c * S + start
                incl = np.cumsum(r_len)
                r_excl = incl - r_len
                b_total = int(incl[31])
                win_lo = 0
                cb_nxt = next(grabs)
            else:
                have = False
            nxt = 0
            win_n = 0
            if have:
                win_n = min(SPOS, b_total - win_lo)
                spos[:] = -1
                for l in range(32):
                    k0 = max(0, win_lo - r_excl[l])
                    k1 = min(r_len[l], max(0, win_lo + win_n - r_excl[l]))
                    for k in range(k0, k1):
                        spos[r_excl[l] + k - win_lo] = r_pos0[l] + k
                if win_n == 0:
                    continue
            elif not (cur_v | nxt_v).any():
                break
        need = ~cur_v
        rank = np.cumsum(need) - need
        for l in range(32):
            if need[l]:
                cur_v[l], cur_pos[l], cur_steps[l] = nxt_v[l], nxt_pos[l], 0
                off = nxt + rank[l]
                nxt_v[l] = off < win_n
                if nxt_v[l]:
                    assert spos[off] >= 0
                    nxt_pos[l] = spos[off]
                out_of_tile[l] = cur_v[l] and rng.random() < p_out      # seed outside the staged range
        nxt = min(nxt + int(need.sum()), win_n)
        for l in range(32):
            if cur_v[l] and not out_of_tile[l]:
                cur_steps[l] += 1
                u = rng.random()
                if u < p_accept or u < p_accept + p_defer or cur_steps[l] >= max_steps:
                    done.append(int(cur_pos[l]))
                    cur_v[l] = False
                elif rng.random() < p_out:
                    out_of_tile[l] = True                                # the walk left the tile
        push = cur_v & out_of_tile
        prank = np.cumsum(push) - push
        for l in range(32):
            if push[l]:
                e = (sg_head + sg_count + prank[l]) & (RING - 1)
                assert ring[e] < 0, "straggler ring overflow"
                ring[e] = cur_pos[l]
                cur_v[l] = False
                out_of_tile[l] = False
        sg_count += int(push.sum())
        if sg_count >= 32:
            run_stragglers(32)
    while sg_count > 0:
        run_stragglers(min(sg_count, 32))
    return done


@pytest.mark.parametrize("mean_run,n_chunks,ranges,wpc", [(4, 300, 1, 3), (8, 1000, 2, 4), (0.2, 500, 1, 2), (40, 70, 1, 1),
                                                           (300, 40, 2, 2), (4, 31, 4, 24), (2000, 33, 1, 2)])
def test_every_position_is_handed_out_exactly_once(mean_run, n_chunks, ranges, wpc):
    rng = np.random.default_rng(int(mean_run * 10) + n_chunks)
    S = 8192
    lens = np.minimum(rng.poisson(mean_run, n_chunks), S)
    if mean_run >= 2000:
        lens[5] = S                       # a whole chunk in one tile: many windows of one run
    starts = rng.integers(0, S - lens + 1)
    runs = np.stack([starts, starts + lens], axis=1)
    want = sorted(int(c * S + s + k) for c, (s, e) in enumerate(runs) for k in range(e - s))
    got = []
    n_blocks = (n_chunks + 31) >> 5
    for r in range(ranges):
        mine = list(range(r, n_blocks, ranges))          # blocks of work item (tile, r), dealt to the warps by a counter
        deal = [[] for _ in range(wpc)]
        for i, blk in enumerate(mine):
            deal[int(rng.integers(0, wpc))].append(blk)
        for w in range(wpc):
            got += simulate(runs, S, ranges, r, deal[w], rng)
    assert sorted(got) == want


def test_slow_walks_and_empty_tiles_terminate():
    rng = np.random.default_rng(1)
    runs = np.zeros((100, 2), dtype=np.int64)
    assert simulate(runs, 8192, 1, 0, [0, 1, 2, 3], rng) == []
    lens = rng.poisson(6, 200)
    runs = np.stack([np.zeros(200, dtype=np.int64), lens], axis=1)
    got = simulate(runs, 8192, 1, 0, range(7), rng, p_accept=0.05, p_defer=0.0, p_out=0.3, max_steps=50)
    assert len(got) == lens.sum() and len(set(got)) == len(got)
