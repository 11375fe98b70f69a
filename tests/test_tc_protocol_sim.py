"""Protocol model of the tcgen05 IIC adjoint's hand-offs (csrc/iic_bwd_tc.cu), run under randomised schedules on the CPU.

The kernel's roles meet at mbarriers whose waits test a phase PARITY: a waiter that looks at a barrier one phase early reads
"complete" (the preceding-phase rule), one that looks two phases late blocks for ever.  Whether that can happen depends on the
ring depths and on how rows are dealt to the converter sets / issuers / epilogue sets — which is what this model checks, for
the committed role split and for the A/B variants of profiles/probes/iic_variants.sh: no schedule may deadlock and no issuer
may start an output row before the three input rows it reads are really in their slots.  (A pair-granular variant of the ring
— two rows per hand-off — failed exactly this check before it ever ran correctly on a GPU)"""
import random

import pytest


class _Bar:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count

    def done(self, parity):          # mbarrier.try_wait.parity: true iff the CURRENT phase's parity differs
        return (self.phase & 1) != parity


def _run(segs, seed, NA=8, ND=8, NCS=4, NISS=4, NES=2):
    rnd = random.Random(seed)
    a_full = [_Bar(1) for _ in range(NA)]        # the kernel counts 4 quarter warps; one converter per set here
    a_empty = [_Bar(3) for _ in range(NA)]       # three reader rows commit
    d_full = [_Bar(1) for _ in range(ND)]
    d_empty = [_Bar(1) for _ in range(ND)]
    in_slot = [-1] * NA                          # global input row that currently sits in the slot
    conv = [[] for _ in range(NCS)]
    iss = [[] for _ in range(NISS)]
    epi = [[] for _ in range(NES)]
    ar0 = or0 = 0
    for n_out in segs:                           # a segment: n_out output rows from n_out + 2 input rows
        for j in range(n_out + 2):
            conv[(ar0 + j) % NCS].append(ar0 + j)
        for i in range(n_out):
            iss[(or0 + i) % NISS].append((or0 + i, ar0 + i, i, n_out))
            epi[(or0 + i) % NES].append(or0 + i)
        ar0 += n_out + 2
        or0 += n_out
    ci, ii, ei, pipe = [0] * NCS, [0] * NISS, [0] * NES, []
    while True:
        progressed = False
        actors = [("c", s) for s in range(NCS)] + [("i", s) for s in range(NISS)] + [("e", s) for s in range(NES)] + [("t", 0)]
        rnd.shuffle(actors)
        for kind, s in actors:
            if kind == "c" and ci[s] < len(conv[s]):
                ar = conv[s][ci[s]]
                if a_empty[ar % NA].done(((ar // NA) & 1) ^ 1):
                    in_slot[ar % NA] = ar
                    a_full[ar % NA].arrive()
                    ci[s] += 1
                    progressed = True
            elif kind == "i" and ii[s] < len(iss[s]):
                orow, a, i, n_out = iss[s][ii[s]]
                if all(a_full[r % NA].done((r // NA) & 1) for r in (a, a + 1, a + 2)) and d_empty[orow % ND].done(((orow // ND) & 1) ^ 1):
                    for r in (a, a + 1, a + 2):
                        if in_slot[r % NA] != r:
                            return f"issuer {s} started output row {orow} while slot {r % NA} held row {in_slot[r % NA]}, not {r}"
                    arrivals = [d_full[orow % ND]]
                    for dyy in range(3):         # the first / last row of a segment stands in for the missing readers
                        arrivals += [a_empty[(a + dyy) % NA]] * (1 + (2 - dyy if i == 0 else 0) + (dyy if i == n_out - 1 else 0))
                    pipe.append(arrivals)        # tcgen05.commit: arrives when the (in-order) tensor pipe gets there
                    ii[s] += 1
                    progressed = True
            elif kind == "e" and ei[s] < len(epi[s]):
                orow = epi[s][ei[s]]
                if d_full[orow % ND].done((orow // ND) & 1):
                    d_empty[orow % ND].arrive()
                    ei[s] += 1
                    progressed = True
            elif kind == "t" and pipe:
                for b in pipe.pop(0):
                    b.arrive()
                progressed = True
        if all(ci[s] == len(conv[s]) for s in range(NCS)) and all(ii[s] == len(iss[s]) for s in range(NISS)) \
                and all(ei[s] == len(epi[s]) for s in range(NES)) and not pipe:
            return None
        if not progressed:
            return "deadlock"


SEGMENTS = ([194], [100, 94], [30, 164], [12], [5, 7], [1], [2], [3], [1, 1, 1], [7, 1, 9], [13], [1, 2, 3, 4, 5, 6, 7], [2, 2, 2, 2])


@pytest.mark.parametrize("cfg", [dict(), dict(NISS=2), dict(NCS=3), dict(NCS=3, NISS=2), dict(NES=1), dict(NCS=2, NISS=2)],
                         ids=lambda c: "committed" if not c else ",".join(f"{k}={v}" for k, v in c.items()))
def test_row_ring_has_no_deadlock_and_no_early_read(cfg):
    for segs in SEGMENTS:
        for seed in range(60):
            assert _run(segs, seed, **cfg) is None, (segs, seed, cfg)


def _converter_rows(segs, cset, NCS=4):
    """the converter cursor of csrc/iic_bwd_tc.cu (enter / step) restated: global input-row indices set ``cset`` writes"""
    out, ar0, k = [], 0, 0
    while k < len(segs):
        n_in = segs[k] + 2
        j = (cset + NCS - ar0 % NCS) % NCS
        while j < n_in:
            out.append(ar0 + j)
            j += NCS
        ar0 += n_in
        k += 1
    return out


def test_converter_sets_cover_every_input_row_exactly_once():
    """a one-row segment has 3 input rows, fewer than the 4 converter sets: the set without a row there must move on to the next
    segment instead of writing one (it would arrive twice on the A slot of the next segment's first row — the hang that
    [3,5,17,228] exposed).  Every global input row belongs to exactly one set, and a set's rows are 4 apart."""
    import random
    rng = random.Random(7)
    for _ in range(300):
        segs = [rng.choice([1, 1, 2, 3, 5, 17]) for _ in range(rng.randint(1, 6))]
        total = sum(n + 2 for n in segs)
        seen = []
        for cset in range(4):
            rows = _converter_rows(segs, cset)
            assert all(r % 4 == cset for r in rows), (segs, cset, rows)
            seen += rows
        assert sorted(seen) == list(range(total)), segs
