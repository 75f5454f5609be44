"""Host model of the rank-level ring of dsgd_svd_kernel (csrc/sgd.cu, P > 1): the item super-blocks travel rank -> rank
through peer-mapped slabs with two buffers per rank, a "delivered" counter written by the sender (rflags) and a credit
returned by the receiver.  A discrete-event simulation with random speeds checks, per block slot:

  * no deadlock for P = 2 .. 8 ranks, K = 1 .. 4 outer steps per sub-epoch, any relative speed;
  * every read sees exactly the data it expects (super-block (rank + E) % P after E * K + T updates): nothing is
    overwritten before its last read, nothing is read before it was delivered.

The protocol, for rank g, sub-epoch E (counted from the start of the fit), slot ib, outer steps T = 0 .. K-1:
    T == 0 and E > 0 : wait rflags[g][ib] >= E                     (the right neighbour delivered the block of sub-epoch E)
    read  buffer[g][E & 1][ib]
    T == K - 1       : after the read, credit[g + 1][ib] = E + 1    (the sender may reuse this slot for sub-epoch E + 2)
    update the block
    T <  K - 1       : write buffer[g][E & 1][ib]                   (hand-off to the next cluster through L2)
    T == K - 1       : wait credit[g][ib] >= E, write buffer[g - 1][(E + 1) & 1][ib], then rflags[g - 1][ib] = E + 1
"""
import heapq
import random

import pytest


class Counter:
    def __init__(self):
        self.v = 0
        self.waiters = []

    def set(self, sim, v):
        assert v >= self.v
        self.v = v
        ready = [w for w in self.waiters if w[0] <= v]
        self.waiters = [w for w in self.waiters if w[0] > v]
        for _, cb in ready:
            sim.at(0.0, cb)

    def wait_ge(self, sim, v, cb):
        if self.v >= v:
            sim.at(0.0, cb)
        else:
            self.waiters.append((v, cb))


class Sim:
    def __init__(self, seed):
        self.t, self.q, self.n, self.rng = 0.0, [], 0, random.Random(seed)

    def at(self, dt, cb):
        self.n += 1
        heapq.heappush(self.q, (self.t + dt, self.n, cb))

    def run(self):
        while self.q:
            self.t, _, cb = heapq.heappop(self.q)
            cb()


class Slot:
    """One block slot of one rank: the chain of reads / updates / writes over the sub-epochs."""

    def __init__(self, sim, g, P, K, n_sub, ranks, speed):
        self.sim, self.g, self.P, self.K, self.n_sub, self.ranks, self.speed = sim, g, P, K, n_sub, ranks, speed
        self.buf = [(g, 0), None]            # buffer parity -> (super-block, updates so far)
        self.rflags, self.credit = Counter(), Counter()
        self.reading = [False, False]
        self.E = self.T = 0
        self.done = False

    def left(self):
        return self.ranks[(self.g - 1) % self.P]

    def right(self):
        return self.ranks[(self.g + 1) % self.P]

    def start(self):
        self.step()

    def step(self):
        if self.E == self.n_sub:
            self.done = True
            return
        if self.T == 0 and self.E > 0:
            self.rflags.wait_ge(self.sim, self.E, self.read)
        else:
            self.read()

    def read(self):
        par = self.E & 1
        self.reading[par] = True
        self.sim.at(self.sim.rng.expovariate(1.0) * 0.2 * self.speed, self.after_read)

    def after_read(self):
        par = self.E & 1
        want = ((self.g + self.E) % self.P, self.E * self.K + self.T)
        assert self.buf[par] == want, "rank %d E %d T %d read %s, expected %s" % (self.g, self.E, self.T, self.buf[par], want)
        self.reading[par] = False
        self.block = (want[0], want[1] + 1)
        if self.T == self.K - 1:
            self.right().credit.set(self.sim, self.E + 1)
        self.sim.at(self.sim.rng.expovariate(1.0) * self.speed, self.after_update)

    def after_update(self):
        if self.T < self.K - 1:
            self.buf[self.E & 1] = self.block
            self.T += 1
            self.step()
        else:
            self.credit.wait_ge(self.sim, self.E, self.send)

    def send(self):
        dst, par, E = self.left(), (self.E + 1) & 1, self.E
        assert not dst.reading[par], "store into a slot the neighbour is reading"
        # the neighbour must be done with what this slot held: its sub-epoch E - 1 (same parity) is over
        assert dst.E > E - 1 or (dst.E == E - 1 and dst.T == dst.K - 1 and not dst.reading[par]) or E == 0, \
            "slot overwritten before its last read (sender E %d, receiver E %d T %d)" % (E, dst.E, dst.T)
        dst.buf[par] = self.block

        def delivered():
            dst.rflags.set(self.sim, E + 1)
        # the flag store travels for a while; stores of one sender to one counter arrive in order (they are a whole
        # sub-epoch apart in the kernel)
        self.flag_eta = max(getattr(self, "flag_eta", 0.0), self.sim.t + self.sim.rng.expovariate(1.0) * 0.1)
        self.sim.at(self.flag_eta - self.sim.t, delivered)
        self.E += 1
        self.T = 0
        self.step()


@pytest.mark.parametrize("P,K", [(2, 1), (2, 3), (3, 2), (4, 1), (4, 4), (8, 1), (8, 2)])
def test_rank_ring_mailboxes(P, K):
    for seed in range(30):
        sim = Sim(seed * 17 + P + K)
        ranks = []
        for g in range(P):
            ranks.append(Slot(sim, g, P, K, 3 * P, ranks, sim.rng.choice([0.05, 0.5, 1.0, 1.0, 4.0, 20.0])))
        for s in ranks:
            s.start()
        sim.run()
        assert all(s.done for s in ranks), "deadlock: %s" % [(s.g, s.E, s.T) for s in ranks if not s.done]


def test_rank_ring_model_notices_a_missing_credit_wait():
    class NoCredit(Slot):
        def after_update(self):
            if self.T < self.K - 1:
                return Slot.after_update(self)
            self.send()

    bad = 0
    for seed in range(30):
        sim = Sim(seed)
        ranks = []
        for g in range(3):
            ranks.append(NoCredit(sim, g, 3, 2, 9, ranks, [0.05, 1.0, 20.0][g]))
        for s in ranks:
            s.start()
        try:
            sim.run()
            assert all(s.done for s in ranks)
        except AssertionError:
            bad += 1
    assert bad > 0
