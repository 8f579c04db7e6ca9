"""Lock-step simulation of k_walkp's per-warp work queue (gsi_kernels.cu): two ready queries per
lane in rotation (P, Q), two feeder stages (N, M), chunks drawn from a shared counter in the middle
of a refill.  The CUDA kernel cannot run here; this restates its queue logic statement for
statement (same variable names) for 32 lanes x several warps and checks the invariants the kernel
relies on: every query of [0, nq) is handed out exactly once, a query is only processed after it
went through M -> N -> P, its leaf line is requested at least one iteration before it is read,
and the loop terminates."""
import random

import pytest


class Warp:
    def __init__(self, nq, chunk, work, n_chunks, dynamic, n_warps, warp_id, rng):
        self.nq, self.chunk, self.work, self.n_chunks = nq, chunk, work, n_chunks
        self.dynamic, self.n_warps, self.rng = dynamic, n_warps, rng
        self.base = self.next = self.total = 0
        self.dry = False
        self.ahead = self.draw() if dynamic else warp_id          # lane 0's register
        z = lambda v: [v] * 32
        self.p_v, self.q_v, self.n_v, self.m_v = z(False), z(False), z(False), z(False)
        self.p_q, self.q_q, self.n_q, self.m_q = z(0), z(0), z(0), z(0)
        self.p_steps, self.q_steps = z(0), z(0)
        self.p_requested_at, self.q_requested_at = z(-1), z(-1)   # iteration of the last prefetch
        self.it = 0
        self.done = False

    def draw(self):
        v = self.work[0]
        self.work[0] += 1
        return v

    def iteration(self, processed, handed):
        """one loop iteration of the kernel; returns False when the warp leaves the loop"""
        for l in range(32):
            if self.p_v[l]:
                # the line of this leaf must have been requested in an EARLIER iteration
                assert 0 <= self.p_requested_at[l] < self.it, (self.it, l)
                self.p_steps[l] += 1
                if self.rng.random() < 0.75 or self.p_steps[l] >= 6:   # accept / defer
                    processed.append(self.p_q[l])
                    self.p_v[l] = False
        need = [not v for v in self.p_v]
        for l in range(32):
            if need[l]:
                self.p_v[l], self.p_q[l], self.p_steps[l] = self.n_v[l], self.n_q[l], 0
                self.n_v[l], self.n_q[l] = self.m_v[l], self.m_q[l]
                self.m_v[l] = False
        take = [l for l in range(32) if need[l]]
        if take and not self.dry:
            cnt = len(take)
            rank = {l: i for i, l in enumerate(take)}
            rem = self.total - self.next
            if cnt > rem:
                ch = self.ahead
                for l in take:
                    if rank[l] < rem:
                        self.m_v[l], self.m_q[l] = True, self.base + self.next + rank[l]
                if ch < self.n_chunks:
                    self.ahead = self.draw() if self.dynamic else ch + self.n_warps
                    self.base = ch * self.chunk
                    self.total = min(self.chunk, self.nq - self.base)
                    want = cnt - rem
                    for l in take:
                        if rank[l] >= rem and rank[l] - rem < self.total:
                            self.m_v[l], self.m_q[l] = True, self.base + (rank[l] - rem)
                    self.next = min(want, self.total)
                else:
                    self.dry = True
                    self.next = self.total
            else:
                for l in take:
                    self.m_v[l], self.m_q[l] = True, self.base + self.next + rank[l]
                self.next += cnt
            for l in take:
                if self.m_v[l]:
                    handed.append(self.m_q[l])
        for l in range(32):
            if self.p_v[l]:
                self.p_requested_at[l] = self.it                  # prefetch.global.L2
        if not any(self.p_v + self.q_v + self.n_v + self.m_v):
            self.done = True
            return False
        self.p_v, self.q_v = self.q_v, self.p_v
        self.p_q, self.q_q = self.q_q, self.p_q
        self.p_steps, self.q_steps = self.q_steps, self.p_steps
        self.p_requested_at, self.q_requested_at = self.q_requested_at, self.p_requested_at
        self.it += 1
        return True


@pytest.mark.parametrize("nq,chunk,n_warps,dynamic", [
    (1, 256, 3, 1), (31, 256, 2, 1), (256, 256, 1, 1), (257, 256, 4, 1), (10_000, 256, 7, 1),
    (10_000, 256, 7, 0), (8192 + 5, 2048, 20, 1), (50_000, 512, 3, 0), (3 * 256 + 17, 256, 64, 1),
    (100, 32, 5, 1), (100, 32, 5, 0), (33, 32, 1, 1),
])
def test_every_query_handed_out_once_and_processed(nq, chunk, n_warps, dynamic):
    rng = random.Random(nq * 31 + chunk + n_warps)
    n_chunks = (nq + chunk - 1) // chunk
    work = [0]
    warps = [Warp(nq, chunk, work, n_chunks, dynamic, n_warps, w, rng) for w in range(n_warps)]
    processed, handed = [], []
    guard = 0
    while not all(w.done for w in warps):
        # warps advance in an arbitrary interleaving
        w = rng.choice([w for w in warps if not w.done])
        w.iteration(processed, handed)
        guard += 1
        assert guard < 50 * (nq + 100) * n_warps, "queue does not terminate"
    assert sorted(handed) == list(range(nq))
    assert sorted(processed) == list(range(nq))
