"""The ring kernel's staging ring is a byte FIFO whose two sides — the issue cursor (where the next item goes) and
the consumer (where the next item is read) — never exchange positions: both apply the same placement rule to the
same sequence of item sizes (blinky_b200/csrc/warp_device.cu: room_for / issue / the consumer's wrap).  This is a
model of exactly those rules, fuzzed: an item must never be placed over one that is still in flight, the consumer
must find every item where it was put, and the pipeline must never stall for good.  (It found the one bug the
rules had: an empty ring has to restart at offset 0, or "everything in flight lies behind cpos" stops being true.)
The kernel itself is exercised with the smallest and the deepest ring on the GPU
(tests/test_gpu_parity.py::test_ring_kernel_schedules_and_ring_geometries)."""
import random

import pytest


class Ring:
    def __init__(self, R, max_inflight, restart_when_empty=True):
        self.R, self.max_inflight, self.restart = R, max_inflight, restart_when_empty
        self.ipos = self.cpos = self.inflight = 0
        self.queue = []  # ground truth: (start, size) of the items in flight, oldest first

    # -- issue side (warp_device.cu: room_for, can_issue, issue)
    def room_for(self, n):
        if self.inflight == 0:
            return True
        if self.ipos > self.cpos:
            return self.ipos + n <= self.R or n <= self.cpos
        return self.ipos + n <= self.cpos

    def can_issue(self, n):
        return self.inflight < self.max_inflight and self.room_for(n)

    def issue(self, n):
        if self.inflight == 0 and self.restart:
            self.ipos = self.cpos = 0
        if self.ipos + n > self.R:
            self.ipos = 0
        for s, m in self.queue:
            assert self.ipos + n <= s or s + m <= self.ipos, ("placed over an item in flight", self.ipos, n, self.queue)
        self.queue.append((self.ipos, n))
        self.ipos += n
        self.inflight += 1

    # -- consumer side (gather_frame / the entry unpack)
    def consume(self, n):
        if self.cpos + n > self.R:
            self.cpos = 0
        s, m = self.queue.pop(0)
        assert (s, m) == (self.cpos, n), ("the consumer looks in the wrong place", s, m, self.cpos, n)
        self.cpos += n
        self.inflight -= 1


def run(R, sizes, max_inflight, lookahead, seed, restart=True):
    rnd = random.Random(seed)
    ring = Ring(R, max_inflight, restart)
    issued = consumed = 0
    while consumed < len(sizes):
        may_issue = issued < len(sizes) and issued - consumed < lookahead and ring.can_issue(sizes[issued])
        if may_issue and (ring.inflight == 0 or rnd.random() < 0.6):
            ring.issue(sizes[issued])
            issued += 1
        elif ring.inflight > 0:
            ring.consume(sizes[consumed])
            consumed += 1
        else:
            raise AssertionError(("stalled for good", issued, consumed, ring.ipos, ring.cpos, sizes[issued]))


ENTRY = 2176  # kBoxBlockBytes


@pytest.mark.parametrize("R", [2176, 8192, 9728, 12800, 16128, 32768])
def test_byte_ring_rules_hold_for_random_item_sequences(R):
    for seed in range(1500):
        rnd = random.Random(seed * 7919 + R)
        # a unit = its entry block followed by 1..16 boxes of one size (multiples of 128 bytes, at most the ring)
        sizes = []
        while len(sizes) < 60:
            box = 128 * rnd.randint(1, min(64, R // 128))
            sizes += [min(ENTRY, R)] + [box] * rnd.choice([1, 1, 2, 3, 8, 16])
        run(R, sizes, max_inflight=rnd.randint(1, 6), lookahead=rnd.choice([2, 3, 6, 24]), seed=seed)


def test_without_the_restart_rule_the_model_fails():
    """the check has teeth: drop 'an empty ring restarts at the front' and an item lands on one in flight"""
    with pytest.raises(AssertionError):
        for seed in range(4000):
            rnd = random.Random(seed)
            R = rnd.choice([8320, 9728, 16128])
            sizes = [128 * rnd.randint(1, min(63, R // 128)) for _ in range(40)]
            run(R, sizes, max_inflight=6, lookahead=3, seed=seed, restart=False)
