"""Model check (CPU) of the one-shot allreduce protocol of harl_b200/csrc/p2p_comm.cu: per exchange `seq` every rank
(1) writes its bucket into its own slot `seq & 1`, (2) publishes `seq` into its flag on every peer, (3) waits until all
its flags show >= seq, (4) reads every rank's slot `seq & 1` and adds them in rank order.  No barrier follows.

Claim checked under random interleavings of the ranks' steps (and random stalls between them): no rank ever reads a slot
while its owner may be overwriting it, every rank obtains the sum of the right exchange, and all ranks obtain the
bit-identical value -- i.e. two slots and a monotone flag per rank are enough."""
import random

import numpy as np
import pytest


class Rank:
    def __init__(self, r, world, n_exchanges, data):
        self.r, self.world, self.n, self.data = r, world, n_exchanges, data
        self.seq, self.pc, self.read_i, self.acc, self.results = 1, 0, 0, None, []

    def step(self, slots, flags, writing):
        """One atomic action of this rank; returns False when it has nothing left to do."""
        if self.seq > self.n:
            return False
        p = self.seq & 1
        if self.pc == 0:                                    # (1) write own slot -- in two halves, so a torn read would show
            writing[self.r][p] = True
            slots[self.r][p][: 2] = self.data[self.seq - 1][self.r][: 2]
            self.pc = 1
        elif self.pc == 1:
            slots[self.r][p][2:] = self.data[self.seq - 1][self.r][2:]
            writing[self.r][p] = False
            self.pc = 2
        elif self.pc == 2:                                  # (2) publish the flag on every peer (after the data: release order)
            for q in range(self.world):
                flags[q][self.r] = self.seq
            self.pc = 3
        elif self.pc == 3:                                  # (3) wait
            if all(flags[self.r][q] >= self.seq for q in range(self.world)):
                self.pc, self.read_i, self.acc = 4, 0, None
        elif self.pc == 4:                                  # (4) read peer read_i, rank order
            q = self.read_i
            assert not writing[q][p], f"rank {self.r} reads slot {p} of rank {q} while it is being written (seq {self.seq})"
            v = slots[q][p].copy()
            self.acc = v if self.acc is None else self.acc + v
            self.read_i += 1
            if self.read_i == self.world:
                self.results.append(self.acc)
                self.seq, self.pc = self.seq + 1, 0
        return True


@pytest.mark.parametrize("world,seed", [(2, 0), (2, 1), (4, 2), (8, 3), (8, 4), (3, 5)])
def test_two_slots_and_a_monotone_flag_suffice(world, seed):
    rng = random.Random(seed)
    n_ex = 40
    data = np.random.default_rng(seed).standard_normal((n_ex, world, 4)).astype(np.float32)
    slots = [[np.zeros(4, np.float32), np.zeros(4, np.float32)] for _ in range(world)]
    flags = [[0] * world for _ in range(world)]
    writing = [[False, False] for _ in range(world)]
    ranks = [Rank(r, world, n_ex, data) for r in range(world)]
    live = list(range(world))
    weights = [rng.random() ** 3 + 0.02 for _ in range(world)]        # some ranks are much slower than others
    steps = 0
    while live:
        r = rng.choices(live, [weights[i] for i in live])[0]
        if not ranks[r].step(slots, flags, writing):
            live.remove(r)
        steps += 1
        if steps % 97 == 0:                                           # the speeds change over time
            weights = [rng.random() ** 3 + 0.02 for _ in range(world)]
        assert steps < 10**6
    for s in range(n_ex):
        want = data[s][0].copy()
        for q in range(1, world):
            want = want + data[s][q]                                  # rank order, fp32
        for rk in ranks:
            assert np.array_equal(rk.results[s], want), (s, rk.r)


def test_a_single_slot_would_not_be_enough():
    """The same model with one slot (no parity) lets a fast rank overwrite data a slow rank has not read yet."""
    world, n_ex = 2, 6
    data = np.arange(n_ex * world * 4, dtype=np.float32).reshape(n_ex, world, 4)
    slots = [[np.zeros(4, np.float32)] * 2 for _ in range(world)]
    for r in range(world):
        slots[r][1] = slots[r][0]                                     # both parities alias ONE buffer
    flags = [[0] * world for _ in range(world)]
    writing = [[False, False] for _ in range(world)]
    ranks = [Rank(r, world, n_ex, data) for r in range(world)]
    # both publish exchange 1; rank 0 waits, reads both slots and starts WRITING exchange 2 before rank 1 has read exchange 1
    order = [0] * 3 + [1] * 3 + [0] * 3 + [0] + [1] * 3 + [0] * 60 + [1] * 60
    bad = False
    try:
        for r in order:
            ranks[r].step(slots, flags, writing)
        for s in range(min(len(ranks[1].results), n_ex)):
            want = data[s][0] + data[s][1]
            bad = bad or not np.array_equal(ranks[1].results[s], want)
    except AssertionError:
        bad = True
    assert bad
