"""Model check of the mbarrier protocol of the EXPERIMENTAL two-tile attention forward (ops/csrc/attention_sm100.cu,
attn_fwd2_kernel).  The kernel itself has never been executed (no GPU time was left when it was written); this is NOT a run
of the kernel.  It is a hand transcription of its synchronisation skeleton -- which role waits on / arrives at which
barrier, with which parity, how many arrivals complete a phase, that tcgen05.commit fires only after every earlier MMA
finished -- driven by random schedules.  It can find deadlocks and phase-parity aliasing (a wait that is satisfied by the
wrong completion) in the DESIGN; it cannot find transcription mistakes, address / swizzle / masking bugs or anything the
hardware does differently from this model.

mbarrier semantics modelled: `phase` = completed phases; arrive() decrements the pending count and completes the phase at 0;
wait(parity) passes iff the phase with that parity has completed, i.e. (phase & 1) != parity.  Every wait also carries the
completion index the code INTENDS to observe; passing with a different completion count is reported as aliasing.
"""
import random

import pytest

KV_STAGES = 3


class Bar:
    def __init__(self, name, count):
        self.name, self.count, self.pending, self.phase = name, count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, f"{self.name}: more arrivals than the barrier was initialised for"
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count

    def ready(self, parity):
        return (self.phase & 1) != parity


class Model:
    def __init__(self, nkv_t, rng):
        self.nkv_t, self.nkv, self.rng = nkv_t, max(nkv_t), rng
        B = lambda n, c: Bar(n, c)
        self.q_full = B("q_full", 1)
        self.kv_full = [B(f"kv_full{i}", 1) for i in range(KV_STAGES)]
        self.kv_empty = [B(f"kv_empty{i}", 1) for i in range(KV_STAGES)]
        self.s_full = [B(f"s_full{t}", 1) for t in range(2)]
        self.s_empty = [B(f"s_empty{t}", 4) for t in range(2)]
        self.p_full = [B(f"p_full{t}", 4) for t in range(2)]
        self.pv_full = [B(f"pv_full{t}", 1) for t in range(2)]
        self.inflight = []        # async completions: (remaining ticks, callback); MMAs complete in issue order
        self.mma_queue = []       # issued MMAs / commits in order: ("mma", ticks) | ("commit", bar)
        self.errors = []
        # what the hardware state would be, to catch use-before-ready / overwrite-before-read
        self.s_owner = [None, None]     # S_t currently holds block j (written by MMA completion)
        self.s_read_done = [-1, -1]     # softmax finished reading S_t(j) (min over its 4 warps)
        self.p_written = [-1, -1]
        self.s_reads = [[-1] * 4, [-1] * 4]
        self.p_busy = [None, None]              # PV_t(j) in flight: reads the P_t smem tile and accumulates into O_t
        self.pv_done = [-1, -1]                 # last PV_t that has COMPLETED (O_t holds blocks 0..pv_done, P_t is free again)
        self.kv_block = [None] * KV_STAGES      # which KV block a smem stage holds (set when its TMA lands)

    # ---- events -------------------------------------------------------------------------------------------------
    def wait(self, bar, parity, intended):
        """generator: block until the barrier shows `parity` complete; check it is completion number `intended`"""
        while not bar.ready(parity):
            yield
        if bar.phase != intended + 1:
            self.errors.append(f"{bar.name}: wait intended completion #{intended} but barrier has {bar.phase} completed phases")

    def tma(self, bar, stage=None, block=None):   # expect_tx + bulk copy: lands later
        def land():
            if stage is not None:
                self.kv_block[stage] = block
            bar.arrive()
        self.inflight.append([self.rng.randint(1, 6), land])

    def issue_mma(self, ticks, on_done=None):
        self.mma_queue.append(["mma", ticks, on_done])

    def commit(self, bar):
        self.mma_queue.append(["commit", bar, None])

    def tick(self):
        for it in self.inflight:
            it[0] -= 1
        done = [it for it in self.inflight if it[0] <= 0]
        self.inflight = [it for it in self.inflight if it[0] > 0]
        for it in done:
            it[1]()
        # tensor pipe: in-order; a commit arrives when everything before it has completed
        while self.mma_queue:
            head = self.mma_queue[0]
            if head[0] == "mma":
                head[1] -= 1
                if head[1] > 0:
                    break
                if head[2]:
                    head[2]()
                self.mma_queue.pop(0)
            else:
                head[1].arrive()
                self.mma_queue.pop(0)

    # ---- roles (transcribed from attn_fwd2_kernel) ----------------------------------------------------------------
    def producer(self):
        self.tma(self.q_full)
        for j in range(self.nkv):
            st = j % KV_STAGES
            yield from self.wait(self.kv_empty[st], ((j // KV_STAGES) & 1) ^ 1, j // KV_STAGES - 1)
            self.tma(self.kv_full[st], st, j)
            yield

    def mma(self):
        def issue_s(t, j):
            st = j % KV_STAGES
            yield from self.wait(self.kv_full[st], (j // KV_STAGES) & 1, j // KV_STAGES)
            if j > 0:
                yield from self.wait(self.s_empty[t], (j - 1) & 1, j - 1)
                if self.s_read_done[t] < j - 1:
                    self.errors.append(f"S_{t}({j}) issued before softmax finished reading S_{t}({j - 1})")

            if self.kv_block[st] != j:
                self.errors.append(f"S_{t}({j}) reads K stage {st} which holds block {self.kv_block[st]}")

            def done(t=t, j=j, st=st):
                if self.kv_block[st] != j:
                    self.errors.append(f"K stage {st} overwritten (block {self.kv_block[st]}) while S_{t}({j}) was in flight")
                self.s_owner[t] = j
            self.issue_mma(self.rng.randint(1, 4), done)
            self.commit(self.s_full[t])

        def issue_pv(t, j):
            yield from self.wait(self.p_full[t], j & 1, j)
            st = j % KV_STAGES
            if self.p_written[t] != j:
                self.errors.append(f"PV_{t}({j}) issued but P buffer holds block {self.p_written[t]}")
            if self.kv_block[st] != j:
                self.errors.append(f"PV_{t}({j}) reads V stage {st} which holds block {self.kv_block[st]}")
            self.p_busy[t] = j

            def done(t=t, j=j, st=st):
                if self.kv_block[st] != j:
                    self.errors.append(f"V stage {st} overwritten (block {self.kv_block[st]}) while PV_{t}({j}) was in flight")
                self.p_busy[t] = None
                self.pv_done[t] = j
            self.issue_mma(self.rng.randint(1, 4), done)
            self.commit(self.pv_full[t])

        yield from self.wait(self.q_full, 0, 0)
        yield from issue_s(0, 0)
        yield from issue_s(1, 0)
        for j in range(self.nkv):
            for t in range(2):
                if j < self.nkv_t[t]:
                    yield from issue_pv(t, j)
                    if j + 1 < self.nkv_t[t]:
                        yield from issue_s(t, j + 1)
            self.commit(self.kv_empty[j % KV_STAGES])
            yield

    def softmax_warp(self, t, w):
        n = self.nkv_t[t]
        for j in range(n):
            yield from self.wait(self.s_full[t], j & 1, j)
            if self.s_owner[t] != j:
                self.errors.append(f"softmax {t}.{w} block {j}: S buffer holds block {self.s_owner[t]}")
            for _ in range(self.rng.randint(1, 3)):   # pass 1
                yield
            if j > 0:
                yield from self.wait(self.pv_full[t], (j - 1) & 1, j - 1)
            if j > 0 and self.pv_done[t] < j - 1:
                self.errors.append(f"softmax {t}.{w} block {j}: rescales O_{t} / rewrites P_{t} before PV_{t}({j - 1}) completed "
                                   f"(last completed: {self.pv_done[t]})")
            for _ in range(self.rng.randint(1, 5)):   # (rescale) + pass 2
                yield
            if self.s_owner[t] != j:
                self.errors.append(f"softmax {t}.{w} block {j}: S buffer overwritten (now block {self.s_owner[t]}) during pass 2")
            self.s_reads[t][w] = j
            self.s_read_done[t] = min(self.s_reads[t])
            self.s_empty[t].arrive()
            yield
            self.p_full[t].arrive()
            if self.p_full[t].pending == self.p_full[t].count:   # this arrival completed the phase: all 4 warps wrote P(j)
                self.p_written[t] = j
        yield from self.wait(self.pv_full[t], (n - 1) & 1, n - 1)
        if self.pv_done[t] != n - 1:       # epilogue: O_t is read out of TMEM
            self.errors.append(f"softmax {t}.{w}: epilogue reads O_{t} before PV_{t}({n - 1}) completed (last completed: {self.pv_done[t]})")

    def run(self, max_ticks=100000):
        agents = [self.producer(), self.mma()] + [self.softmax_warp(t, w) for t in range(2) for w in range(4)]
        live = list(agents)
        for _ in range(max_ticks):
            if not live:
                break
            self.rng.shuffle(live)
            for a in list(live):
                if self.rng.random() < 0.7:     # random interleaving / stalls
                    try:
                        next(a)
                    except StopIteration:
                        live.remove(a)
            self.tick()
        else:
            self.errors.append(f"deadlock / livelock: {len(live)} roles still blocked after {max_ticks} ticks")
        return self.errors


def _configs():
    for qb2 in range(4):                       # causal, S up to 1024: tile A sees 2qb2+1 blocks, tile B 2qb2+2
        yield (2 * qb2 + 1, 2 * qb2 + 2)
    for n in (2, 3, 4, 8, 9):                  # non-causal: both tiles see every block
        yield (n, n)


@pytest.mark.parametrize("nkv_t", list(_configs()))
def test_fwd2_barrier_protocol_has_no_deadlock_or_phase_aliasing(nkv_t):
    for seed in range(60):
        errs = Model(list(nkv_t), random.Random(seed)).run()
        assert not errs, (nkv_t, seed, errs[:4])


def _drops(prefix):
    class Broken(Model):
        def wait(self, bar, parity, intended):
            if bar.name.startswith(prefix):
                return
                yield
            yield from super().wait(bar, parity, intended)
    return Broken


def test_checker_catches_epilogue_not_waiting_for_last_pv():
    """Seeded bug 1: without the pv_full waits the epilogue can read O_t while the last PV_t is still accumulating."""
    Broken = _drops("pv_full")
    assert any(any("epilogue reads" in e for e in Broken([4, 4], random.Random(s)).run()) for s in range(60))


def test_mid_loop_pv_full_wait_is_implied_by_s_full():
    """Finding (re-derived with the "PV_t(j-1) has COMPLETED" predicate; an earlier version of this model checked "nothing in
    flight", which was vacuously true and proved nothing): inside the loop s_full[t](j) is committed after PV_t(j-1) was
    issued, and a tcgen05.commit fires only once every earlier MMA has completed, so waiting for S_t(j) already orders the
    softmax after PV_t(j-1).  This rests on the MODEL's assumption that MMAs retire in issue order and a commit covers all
    of them (the assumption the hardware-validated kernels make for their kv_empty commits); it would stop holding if
    S_t(j+1) were ever issued ahead of PV_t(j).  Only the epilogue's pv_full wait is load-bearing."""
    class MidOnly(Model):
        def wait(self, bar, parity, intended):
            if bar.name.startswith("pv_full") and intended < self.nkv_t[int(bar.name[-1])] - 1:
                return
                yield
            yield from super().wait(bar, parity, intended)
    for cfg in ([4, 4], [3, 4], [7, 8]):
        assert not any(MidOnly(list(cfg), random.Random(s)).run() for s in range(60)), cfg


def test_checker_catches_producer_not_waiting_for_stage_release():
    """Seeded bug 2: the TMA producer refills a K/V stage without waiting for kv_empty -> must be reported."""
    Broken = _drops("kv_empty")
    assert any(any("stage" in e for e in Broken([8, 8], random.Random(s)).run()) for s in range(40))


def test_s_empty_wait_is_implied_by_p_full_in_the_current_issue_order():
    """Finding, not a guarantee: the MMA role issues S_t(j+1) only after PV_t(j), i.e. after p_full[t](j), and every softmax
    warp arrives at s_empty BEFORE p_full -- so dropping the s_empty wait changes nothing in this model.  The barrier only
    matters if the issue order is ever changed to launch S_t(j+1) ahead of PV_t(j)."""
    Broken = _drops("s_empty")
    assert not any(Broken([4, 4], random.Random(s)).run() for s in range(40))
