"""Host model of the DSMEM hop of dsgd_svd_kernel (csrc/sgd.cu, bulk-copy hop): a discrete-event simulation of the
mailbox protocol with random work and transfer times.  It checks what the kernel relies on and cannot assert itself:

  * no deadlock for any cluster size (2 .. 16) and any relative speed of the CTAs;
  * a bulk copy never lands in a buffer its receiver is still working in, nor in one the copy engine is still reading
    for the receiver's own outgoing push (the "free" mailbox, signalled by the receiver of the PREVIOUS hop);
  * mailbox phases never run two ahead of a waiter (a one-bit parity wait would become ambiguous);
  * every CTA works on the item blocks in ring order: CTA c sees block (c + t) % C at step t.

The protocol, per CTA c and push k (= step k of a cluster's inner ring; two buffers, two "data" and two "free" mailboxes):
    wait  F[(k-1) & 1] phase (k-1) >> 1            (k > 0: the left neighbour's spare buffer is free)
    post  arrive.expect_tx on own D[k & 1]
    copy  working buffer -> left neighbour's spare buffer, completing on the left neighbour's D[k & 1]
    wait  own D[k & 1] phase k >> 1                 (the right neighbour's block has landed in the spare buffer)
    arrive on F[k & 1] of CTA c + 2                 (the right neighbour's source buffer may be overwritten)
    swap  working / spare buffer
"""
import heapq
import random

import pytest


class Mailbox:
    """mbarrier with one expected arrival per phase and a transaction count (data mailboxes only)."""

    def __init__(self):
        self.phase = 0          # completed phases
        self.arrivals = 0
        self.tx = 0
        self.waiters = []       # (phase_index, callback)

    def _maybe_complete(self, sim):
        if self.arrivals >= 1 and self.tx == 0:
            self.arrivals -= 1
            self.phase += 1
            ready = [w for w in self.waiters if w[0] < self.phase]
            self.waiters = [w for w in self.waiters if w[0] >= self.phase]
            for _, cb in ready:
                sim.call_soon(cb)

    def arrive(self, sim, expect=0):
        self.tx += expect
        self.arrivals += 1
        assert self.arrivals <= 1, "two arrivals pending in one phase: the mailbox count would underflow"
        self._maybe_complete(sim)

    def complete_tx(self, sim, n):
        self.tx -= n
        self._maybe_complete(sim)

    def wait(self, sim, use, cb):
        """wait for the completion of phase number `use` (0-based); the kernel waits on parity use & 1"""
        assert self.phase - use <= 1, "mailbox ran two phases ahead of its waiter: parity wait is ambiguous"
        if self.phase > use:
            sim.call_soon(cb)
        else:
            self.waiters.append((use, cb))


class Sim:
    def __init__(self, seed):
        self.t = 0.0
        self.q = []
        self.n = 0
        self.rng = random.Random(seed)

    def at(self, dt, cb):
        self.n += 1
        heapq.heappush(self.q, (self.t + dt, self.n, cb))

    def call_soon(self, cb):
        self.at(0.0, cb)

    def run(self):
        while self.q:
            self.t, _, cb = heapq.heappop(self.q)
            cb()


class Cta:
    def __init__(self, sim, c, C, steps, ring, speed):
        self.sim, self.c, self.C, self.steps, self.ring, self.speed = sim, c, C, steps, ring, speed
        self.D = [Mailbox(), Mailbox()]
        self.F = [Mailbox(), Mailbox()]
        self.buf = [("block", c), None]   # contents of the two item buffers
        self.slot = 0                     # working buffer
        self.busy = True                  # the threads are reading / writing buf[slot]
        self.reading = [False, False]     # the copy engine is reading buf[x] for an outgoing push
        self.k = 0
        self.seen = []
        self.done = False

    def left(self):
        return self.ring[(self.c - 1) % self.C]

    def start(self):
        self.work()

    def work(self):
        self.busy = True
        self.seen.append(self.buf[self.slot][1])
        self.sim.at(self.sim.rng.expovariate(1.0) * self.speed, self.after_work)

    def after_work(self):
        self.busy = False
        if self.k == self.steps - 1:      # the last step of the outer ring step: no push (L2 hand-off in the kernel)
            self.done = True
            return
        k = self.k
        if k > 0:
            self.F[(k - 1) & 1].wait(self.sim, (k - 1) >> 1, self.push)
        else:
            self.push()

    def push(self):
        k = self.k
        self.D[k & 1].arrive(self.sim, expect=1)          # own arrive.expect_tx (one "byte" = the whole block)
        src = self.slot
        self.reading[src] = True
        payload = self.buf[src]
        dst = self.left()
        self.sim.at(self.sim.rng.expovariate(1.0) * 0.5, lambda: self.landed(dst, k, src, payload))
        self.D[k & 1].wait(self.sim, k >> 1, self.received)

    def landed(self, dst, k, src, payload):
        # The copy writes into the receiver's spare buffer of ITS push k: the working buffer of step k is k & 1 (the
        # buffers swap at every push).  The receiver may still be a step behind (copies of two consecutive pushes go
        # to different buffers and different mailboxes, so push k may even overtake push k - 1) but never further.
        spare = (k & 1) ^ 1
        assert dst.k in (k - 1, k), "copy of push %d landed while the receiver is at step %d" % (k, dst.k)
        assert not (dst.busy and dst.slot == spare), "copy landed in the receiver's working buffer"
        assert not dst.reading[spare], "copy landed in a buffer the copy engine is still reading"
        dst.buf[spare] = payload
        self.reading[src] = False
        dst.D[k & 1].complete_tx(self.sim, 1)

    def received(self):
        k = self.k
        self.ring[(self.c + 2) % self.C].F[k & 1].arrive(self.sim)
        self.slot ^= 1
        self.k += 1
        self.work()


@pytest.mark.parametrize("C", [2, 3, 4, 8, 16])
def test_bulk_hop_protocol_never_deadlocks_or_clobbers(C):
    for seed in range(40):
        sim = Sim(seed * 131 + C)
        ring = []
        # very different speeds, including one straggler and one sprinter
        speeds = [sim.rng.choice([0.05, 0.3, 1.0, 1.0, 1.0, 3.0, 10.0]) for _ in range(C)]
        for c in range(C):
            ring.append(Cta(sim, c, C, C, ring, speeds[c]))
        for cta in ring:
            cta.start()
        sim.run()
        assert all(cta.done for cta in ring), "deadlock: %s" % [(cta.c, cta.k) for cta in ring if not cta.done]
        for cta in ring:
            assert cta.seen == [(cta.c + t) % C for t in range(C)]


def test_bulk_hop_protocol_many_rounds():
    """Several outer steps back to back (the push counter and the mailbox phases keep running across them)."""
    C, rounds = 4, 6
    for seed in range(20):
        sim = Sim(seed)
        ring = []
        for c in range(C):
            ring.append(Cta(sim, c, C, C * rounds, ring, sim.rng.choice([0.2, 1.0, 5.0])))
        for cta in ring:
            cta.start()
        sim.run()
        assert all(cta.done for cta in ring)
        for cta in ring:
            assert cta.seen == [(cta.c + t) % C for t in range(C * rounds)]


def _failures(cls, C=4, seeds=20):
    bad = 0
    for seed in range(seeds):
        sim = Sim(seed)
        ring = []
        for c in range(C):
            ring.append(cls(sim, c, C, C * 3, ring, sim.rng.choice([0.05, 1.0, 10.0])))
        for cta in ring:
            cta.start()
        try:
            sim.run()
            assert all(cta.done for cta in ring)
        except AssertionError:
            bad += 1
    return bad


def test_model_notices_broken_protocols():
    """The checks above are not vacuous: dropping the "free" wait, or signalling the wrong CTA, is caught."""

    class NoFreeWait(Cta):
        def after_work(self):
            self.busy = False
            if self.k == self.steps - 1:
                self.done = True
                return
            self.push()

    class FreeToWrongCta(Cta):
        def received(self):
            k = self.k
            self.ring[(self.c + 1) % self.C].F[k & 1].arrive(self.sim)
            self.slot ^= 1
            self.k += 1
            self.work()

    assert _failures(Cta) == 0
    assert _failures(NoFreeWait) > 0
    assert _failures(FreeToWrongCta) > 0
