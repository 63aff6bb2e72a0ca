"""Model check of the mbarrier protocol of the persistent tcgen05 GEMM (rstnet_b200/csrc/gemm_tc.cu, gemm_tc_ts_kernel).

The kernel's warp roles (TMA producer, two transform warp groups, MMA issuer, drain) talk through parity-waited
mbarriers.  A parity wait is satisfied by "the phase with that parity has completed", so a waiter that can get a
whole phase ahead of the barrier is satisfied by the phase BEFORE the one it means -- which happened in round 1: with 5
shared-memory stages the previous fill of a slot belongs to the other transform group, the group waited on full[s]
first, and read an A tile whose TMA was still in flight.  This test replays the kernel's index / parity formulas in a
randomly scheduled discrete-event model (adversarial TMA and MMA completion times) and checks that every read sees
the data it means; it also shows that the original wait order is caught by the model."""
import random

import pytest

NA, CH = 4, 4


class Bar:
    def __init__(self):
        self.phase = 0          # index of the current (incomplete) phase

    def arrive(self):
        self.phase += 1

    def passed(self, parity: int) -> bool:   # mbarrier.try_wait.parity
        return (self.phase & 1) != parity


def simulate(S: int, n_stages: int, seed: int, full_first: bool, tma_in_order: bool, tile_k: int = 16):
    """returns None, or a string describing the first protocol violation.  The CTA's stream is n_stages / tile_k tiles of
    tile_k stages each; a tile is cut in chunks of up to 4 stages (the last one may be short), one accumulator hand-over per
    chunk -- the unpadded schedule of round 2."""
    rnd = random.Random(seed)
    full, empty = [Bar() for _ in range(S)], [Bar() for _ in range(S)]
    a_ready, a_free = [Bar() for _ in range(NA)], [Bar() for _ in range(NA)]
    acc_full, acc_empty = [Bar(), Bar()], [Bar(), Bar()]
    slot = [None] * S            # stage whose operands are completely in the smem slot (None while a TMA is writing it)
    tmem_a = [None] * NA
    acc = [None, None]
    in_flight = []               # TMA loads issued, not landed: (slot, stage)
    mma_q = []                   # MMAs issued, not retired (retire in order): stage
    prod = {"g": 0}
    xf = [{"g": 0, "step": 0}, {"g": 1, "step": 0}]
    mma = {"g": 0, "step": 0}
    drain = {"c": 0}
    assert n_stages % tile_k == 0
    chunk_of, first_of, last_of = [], [], []      # per stage: chunk index in the stream, first / last stage of its chunk?
    c = 0
    for t in range(n_stages // tile_k):
        for k in range(tile_k):
            chunk_of.append(c)
            first_of.append(k % CH == 0)
            last = k % CH == CH - 1 or k == tile_k - 1
            last_of.append(last)
            if last:
                c += 1
    n_chunks = c

    def step_producer():
        g = prod["g"]
        if g >= n_stages:
            return False
        s = g % S
        par = (g // S) & 1
        if not empty[s].passed(par ^ 1):
            return False
        slot[s] = None
        in_flight.append((s, g))
        prod["g"] += 1
        return True

    def step_land():
        if not in_flight:
            return False
        i = 0 if tma_in_order else rnd.randrange(len(in_flight))
        s, g = in_flight.pop(i)
        slot[s] = g
        full[s].arrive()
        return True

    def step_transform(r):
        st = xf[r]
        g = st["g"]
        if g >= n_stages:
            return False
        s, sa = g % S, g % NA
        w_full = lambda: full[s].passed((g // S) & 1)
        w_free = lambda: a_free[sa].passed(((g // NA) & 1) ^ 1)
        waits = [w_full, w_free] if full_first else [w_free, w_full]
        if st["step"] < 2:
            if not waits[st["step"]]():
                return False
            st["step"] += 1
            return True
        if slot[s] != g:
            raise AssertionError(f"transform group {r} read slot {s} for stage {g} but it holds {slot[s]}")
        tmem_a[sa] = g
        a_ready[sa].arrive()
        st["g"], st["step"] = g + 2, 0
        return True

    def step_mma():
        g = mma["g"]
        if g >= n_stages:
            return False
        cc, sa = chunk_of[g], g % NA
        buf, s = cc & 1, g % S
        if mma["step"] == 0:
            if first_of[g] and not acc_empty[buf].passed(((cc >> 1) & 1) ^ 1):
                return False
            mma["step"] = 1
            return True
        if not a_ready[sa].passed((g // NA) & 1):
            return False
        if tmem_a[sa] != g or slot[s] != g:
            raise AssertionError(f"MMA of stage {g} saw A stage {tmem_a[sa]} / smem slot {slot[s]}")
        mma_q.append(g)
        mma["g"], mma["step"] = g + 1, 0
        return True

    def step_retire():
        if not mma_q:
            return False
        g = mma_q.pop(0)
        empty[g % S].arrive()
        a_free[g % NA].arrive()
        if last_of[g]:
            acc[chunk_of[g] & 1] = chunk_of[g]
            acc_full[chunk_of[g] & 1].arrive()
        return True

    def step_drain():
        c = drain["c"]
        if c >= n_chunks:
            return False
        buf = c & 1
        if not acc_full[buf].passed((c >> 1) & 1):
            return False
        if acc[buf] != c:
            raise AssertionError(f"drain of chunk {c} read accumulator of chunk {acc[buf]}")
        acc_empty[buf].arrive()
        drain["c"] += 1
        return True

    actors = [step_producer, step_land, lambda: step_transform(0), lambda: step_transform(1), step_mma, step_retire, step_drain]
    # random per-actor speeds make one transform group or the TMA engine arbitrarily slow
    weights = [rnd.choice([1, 1, 3, 10]) for _ in actors]
    try:
        idle = 0
        while drain["c"] < n_chunks:
            a = rnd.choices(range(len(actors)), weights)[0]
            if actors[a]():
                idle = 0
            else:
                idle += 1
                if idle > 20000:
                    return "deadlock"
    except AssertionError as e:
        return str(e)
    return None


@pytest.mark.parametrize("S", [5, 6])
@pytest.mark.parametrize("tile_k", [1, 2, 5, 6, 16])
def test_kernel_wait_order_is_safe(S, tile_k):
    """the order the kernel uses (a_free, then full): no violation under any schedule tried, for short-K tiles (one
    stage per tile, short last chunks) as well as long ones."""
    n = 80 // tile_k * tile_k
    for seed in range(120):
        for in_order in (True, False):
            assert simulate(S, n, seed, full_first=False, tma_in_order=in_order, tile_k=tile_k) is None, (S, tile_k, seed, in_order)


def test_model_catches_the_original_wait_order_with_five_stages():
    """full first, then a_free, with an odd stage count: some schedule lets a transform group read a slot whose TMA is
    still in flight (the round-1 bug); with an even stage count both fills of a slot belong to the same group and the
    old order is safe too."""
    bad = [seed for seed in range(300) if simulate(5, 80, seed, full_first=True, tma_in_order=False)]
    assert bad, "the model no longer reproduces the aliasing hazard it documents"
    assert all(simulate(6, 80, seed, full_first=True, tma_in_order=False) is None for seed in range(100))
