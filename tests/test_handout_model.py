"""Model check of the tile hand-out protocol of the persistent ring kernels (csrc/fft_pipe.cuh "Tile order" / "End of work", the same
code in csrc/fft_lastpipe.cuh): a CPU restatement of the protocol as a state machine, run under random interleavings of the two thread
groups of every CTA. The GPU tests exercise the real kernels; the rare orders (a group that still holds a tile for the other group when
it finds its first closed slot) are a matter of timing there, here they are enumerated by the scheduler.

Protocol: ring slot s of a CTA belongs to group s % 2 and is loaded by the group that consumed slot s - 3 (slots 0..2: at start-up, by
position). A group reads the global counter at the end of a turn (`ahead`), parks it at the top of the next turn and loads that tile
into the slot three ahead in the middle of that turn. A number past the end closes the slot. A group that finds a closed slot passes
on what it holds, and leaves once it has (or start-up has) closed a slot of the other group.

Checked: every tile is consumed exactly once, nobody waits for a slot that is never posted (no deadlock), every counter read has been
parked when the CTA ends (the last CTA may reset the counter)."""
import random

import pytest

CLOSED = -1


class Group:
    def __init__(self, cta, g):
        self.cta, self.g = cta, g
        self.slot = g          # ring slot this group waits for next
        self.ahead = None      # counter value read, not parked yet
        self.parked = None     # next_of[g]
        self.pc = "prologue"   # next micro-step
        self.done = False


class Cta:
    def __init__(self, first, grid, ntiles, sim):
        self.first, self.grid, self.ntiles, self.sim = first, grid, ntiles, sim
        self.slots = {}        # slot -> tile or CLOSED
        self.told = [False, False]
        for k in range(3):     # start-up: the first three slots by position
            tile = first + k * grid
            self.post(k, tile if tile < ntiles else None, by=None)
        self.groups = [Group(self, 0), Group(self, 1)]

    def post(self, slot, tile, by):
        assert slot not in self.slots, "slot posted twice"
        if tile is None or tile >= self.ntiles:
            self.slots[slot] = CLOSED
            if by is None:
                self.told[(slot + 1) & 1] = True   # closed at start-up counts as closed by the other group
            else:
                self.told[by] = True
        else:
            self.slots[slot] = tile

    def take(self):
        v = 3 * self.grid + self.sim.counter
        self.sim.counter += 1
        return v


class Sim:
    def __init__(self, grid, ntiles, rng):
        self.counter = 0
        self.consumed = []
        self.ctas = [Cta(c, grid, ntiles, self) for c in range(grid)]
        self.rng = rng
        self.ntiles = ntiles

    def runnable(self, gr):
        if gr.done:
            return False
        if gr.pc == "wait":
            return gr.slot in gr.cta.slots
        return True

    def step(self, gr):
        c = gr.cta
        if gr.pc == "prologue":          # ahead = take() before the loop
            gr.ahead = c.take()
            gr.pc = "park"
        elif gr.pc == "park":            # top of the turn: next_of[g] = ahead
            gr.parked, gr.ahead = gr.ahead, None
            gr.pc = "wait"
        elif gr.pc == "wait":            # the slot has been posted
            tile = c.slots[gr.slot]
            if tile == CLOSED:
                c.post(gr.slot + 3, gr.parked, by=gr.g)
                gr.parked = None
                if c.told[gr.g]:
                    gr.done = True
                    return
                gr.pc = "draw"
            else:
                self.consumed.append(tile)
                gr.pc = "issue"
        elif gr.pc == "issue":           # after the last gather: load the slot three ahead
            c.post(gr.slot + 3, gr.parked, by=gr.g)
            gr.parked = None
            gr.pc = "draw"
        elif gr.pc == "draw":            # end of the turn: read the counter for the turn after the next
            gr.ahead = c.take()
            gr.slot += 2
            gr.pc = "park"

    def run(self, bias):
        groups = [g for c in self.ctas for g in c.groups]
        for _ in range(200000):
            ready = [g for g in groups if self.runnable(g)]
            if not ready:
                break
            # a biased scheduler: with probability `bias` keep running the group that ran last (long solo runs produce the skewed orders)
            if bias and getattr(self, "last", None) in ready and self.rng.random() < bias:
                g = self.last
            else:
                g = self.rng.choice(ready)
            self.last = g
            self.step(g)
        assert all(g.done for g in groups), "deadlock: a group waits for a slot that is never posted"
        assert sorted(self.consumed) == list(range(self.ntiles)), "a tile was lost or consumed twice"
        assert all(g.ahead is None for g in groups), "a counter read was still in flight at the end"


@pytest.mark.parametrize("bias", [0.0, 0.7, 0.95])
def test_tile_hand_out_protocol_under_random_interleavings(bias):
    rng = random.Random(1234 + int(bias * 100))
    for trial in range(1500):
        grid = rng.randint(1, 5)
        ntiles = rng.randint(0, 14 * grid)
        Sim(grid, ntiles, rng).run(bias)


def test_the_model_catches_a_group_that_leaves_on_its_first_closed_slot():
    """The first version of the kernel left a group at its first closed slot after closing the slot three ahead: the tile it still held for
    the other group was lost when the two groups were far enough out of step. The model must reject that protocol."""
    class EarlyLeave(Sim):
        def step(self, gr):
            c = gr.cta
            if gr.pc == "wait" and c.slots[gr.slot] == CLOSED:
                c.post(gr.slot + 3, None, by=gr.g)   # closes the slot ahead and leaves, whatever it holds
                gr.parked = None
                gr.done = True
                return
            super().step(gr)

    rng = random.Random(7)
    failures = 0
    for trial in range(3000):
        grid = rng.randint(1, 3)
        ntiles = rng.randint(3 * grid, 12 * grid)
        try:
            EarlyLeave(grid, ntiles, rng).run(0.9)
        except AssertionError:
            failures += 1
    assert failures > 0
