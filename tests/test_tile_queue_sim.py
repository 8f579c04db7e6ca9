"""Lock-step simulation of the warp work queue of k_tile_walk (csrc/gsi_kernels.cu): the control
flow that flattens a tile's runs into a dense position list, hands positions to idle lanes and
keeps lanes with unfinished walks busy across list refills.  A transliteration of the kernel's
loop with the walk replaced by a coin (accept / move / defer); checks what the kernel must
guarantee whatever the geometry: it terminates, and every binned position of the tile is handed
out exactly once.  (A GPU hang costs a whole box lease; this runs in milliseconds.)"""
import numpy as np
import pytest

SPOS = 256


def simulate(runs, S, ranges, rng_id, wpc, warp, rng, p_accept=0.75, p_defer=0.01, max_steps=8, guard=10**6):
    """runs: (n_chunks,) array of (start, end) per chunk for one tile.  Returns the positions this
    warp processed, in order."""
    n_chunks = len(runs)
    n_blocks = (n_chunks + 31) >> 5
    lanes = np.arange(32)
    cb = rng_id * wpc + warp
    cb_stride = ranges * wpc
    r_pos0 = np.zeros(32, dtype=np.int64); r_len = np.zeros(32, dtype=np.int64); r_excl = np.zeros(32, dtype=np.int64)
    b_total = win_lo = win_n = nxt = 0
    cur_v = np.zeros(32, dtype=bool); cur_pos = np.zeros(32, dtype=np.int64); cur_steps = np.zeros(32, dtype=np.int64)
    nxt_v = np.zeros(32, dtype=bool); nxt_pos = np.zeros(32, dtype=np.int64)      # coordinates on their way (one refill ahead)
    nn_v = np.zeros(32, dtype=bool); nn_pos = np.zeros(32, dtype=np.int64)        # ... two refills ahead
    spos = np.full(SPOS, -1, dtype=np.int64)
    done = []
    it = 0
    while True:
        it += 1
        assert it < guard, "work queue does not terminate"
        if nxt >= win_n:
            have = True
            win_lo += win_n
            if win_lo < b_total:
                pass
            elif cb < n_blocks:
                c = (cb << 5) + lanes
                ok = c < n_chunks
                start = np.where(ok, runs[np.minimum(c, n_chunks - 1), 0], 0)
                end = np.where(ok, runs[np.minimum(c, n_chunks - 1), 1], 0)
                r_len = end - start
                r_pos0 = c * S + start
                incl = np.cumsum(r_len)
                r_excl = incl - r_len
                b_total = int(incl[31])
                win_lo = 0
                cb += cb_stride
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
            elif not (cur_v | nxt_v | nn_v).any():
                break
        need = ~cur_v
        rank = np.cumsum(need) - need
        for l in range(32):
            if need[l]:
                cur_v[l], cur_pos[l], cur_steps[l] = nxt_v[l], nxt_pos[l], 0
                nxt_v[l], nxt_pos[l] = nn_v[l], nn_pos[l]
                off = nxt + rank[l]
                nn_v[l] = off < win_n
                if nn_v[l]:
                    assert spos[off] >= 0
                    nn_pos[l] = spos[off]
        nxt = min(nxt + int(need.sum()), win_n)
        for l in range(32):
            if cur_v[l]:
                cur_steps[l] += 1
                u = rng.random()
                if u < p_accept or u < p_accept + p_defer or cur_steps[l] >= max_steps:
                    done.append(int(cur_pos[l]))
                    cur_v[l] = False
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
    for r in range(ranges):
        for w in range(wpc):
            got += simulate(runs, S, ranges, r, wpc, w, rng)
    assert sorted(got) == want


def test_slow_walks_and_empty_tiles_terminate():
    rng = np.random.default_rng(1)
    runs = np.zeros((100, 2), dtype=np.int64)
    assert simulate(runs, 8192, 1, 0, 2, 0, rng) == []
    lens = rng.poisson(6, 200)
    runs = np.stack([np.zeros(200, dtype=np.int64), lens], axis=1)
    got = simulate(runs, 8192, 1, 0, 1, 0, rng, p_accept=0.05, p_defer=0.0, max_steps=50)
    assert len(got) == lens.sum() and len(set(got)) == len(got)
