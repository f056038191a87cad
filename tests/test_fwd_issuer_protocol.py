"""Protocol model of the opt-in forward hand-off variants (fa_fwd_f16_sm100.cu, FwdCfg VAR bits 1 and 2; DESIGN.md
section 6b). The variants are compiled but have not run on a GPU yet, so their barrier protocol is checked here with a
randomised interleaving model: the TMA producer, the two softmax warpgroups, the MMA issuer (fixed order for VAR = 2,
polling for VAR = 6 / 7) and the in-order tensor pipe are state machines that share mbarriers with the hardware's
phase / parity semantics. Checked for every schedule: no deadlock, no parity test that succeeds on a stale phase, every
MMA reads the ring stage / TMEM contents it was issued for, the producer never overwrites a stage with reads outstanding,
and the softmax never reads S before both halves of Q K^T have executed.

This mirrors the issuer code line by line (same ring-slot arithmetic, same use counters); it is a model of the protocol,
not of the CUDA code generation."""
import random

import pytest

K_STAGES = 4


class Bar:
    """mbarrier: `done` completed phases; try/test_wait(parity) succeeds iff the current phase parity differs."""

    def __init__(self, count):
        self.count, self.pending, self.done = count, count, 0

    def arrive(self, n=1):
        self.pending -= n
        assert self.pending >= 0
        if self.pending == 0:
            self.done += 1
            self.pending = self.count

    def test(self, parity):
        return (self.done & 1) != parity


class Sim:
    def __init__(self, n, var, rng):
        self.n, self.var, self.rng = n, var, rng
        self.kv_full = [Bar(1) for _ in range(K_STAGES)]
        self.kv_empty = [Bar(1) for _ in range(K_STAGES)]
        self.s_full = [Bar(1), Bar(1)]
        self.p_half = [Bar(1), Bar(1)]      # 128 thread arrivals modelled as one
        self.p_ready = [Bar(1), Bar(1)]
        self.s_cons = [Bar(1), Bar(1)]
        self.o_final = [Bar(1), Bar(1)]
        self.stage_content = [None] * K_STAGES     # ring slot t currently held
        self.stage_reads = [0] * K_STAGES          # issued-but-not-executed MMAs reading the stage
        self.pipe = []                             # in-order tensor pipe: ("mma", fn) / ("commit", bar)
        self.qk_done = [[0, 0], [0, 0]]            # per warpgroup: tiles whose [lower, upper] S half has executed
        self.s_read = [0, 0]                       # tiles whose S row the softmax holds in registers
        self.p_written = [[0, 0], [0, 0]]          # per warpgroup: tiles whose [first, second] P half is stored
        self.pv_done = [[0, 0], [0, 0]]
        self.finished = [False, False]

    # ---- tensor pipe -------------------------------------------------------------------------------------------------
    def issue(self, stage, t, fn):
        assert self.stage_content[stage] == t, f"MMA issued on stage {stage} holding {self.stage_content[stage]}, wants {t}"
        self.stage_reads[stage] += 1

        def run():
            assert self.stage_content[stage] == t, "ring stage overwritten before the MMA executed"
            self.stage_reads[stage] -= 1
            fn()
        self.pipe.append(("mma", run))

    def commit(self, bar):
        self.pipe.append(("commit", bar))

    def pipe_step(self):
        if not self.pipe:
            return False
        kind, x = self.pipe.pop(0)
        if kind == "mma":
            x()
        else:
            x.arrive()
        return True

    # ---- MMAs --------------------------------------------------------------------------------------------------------
    def qk(self, i, j, halves):
        t = 2 * j

        def fn():
            for h in halves:
                if h == 1:   # upper half: the softmax must hold S of the previous tile in registers
                    assert j == 0 or self.s_read[i] >= j, "upper half of S overwritten before the softmax read it"
                else:        # lower half aliases P of the previous tile: its P V must have executed (pipe order)
                    assert j == 0 or self.pv_done[i][1] >= j, "lower half of S overwrites P before P V consumed it"
                assert self.qk_done[i][h] == j
                self.qk_done[i][h] = j + 1
        self.issue(t % K_STAGES, t, fn)

    def pv(self, i, j, half):
        t = 2 * j + 1

        def fn():
            assert self.p_written[i][half] >= j + 1, "P V executed before its half of P was stored"
            assert self.pv_done[i][half] == j
            self.pv_done[i][half] = j + 1
        self.issue(t % K_STAGES, t, fn)

    # ---- actors (generators: each `yield` is a point where another actor may run) -------------------------------------
    def producer(self):
        for t in range(2 * self.n):
            s, u = t % K_STAGES, t // K_STAGES
            while not self.kv_empty[s].test((u & 1) ^ 1):
                yield
            assert self.stage_reads[s] == 0, "producer overwrites a stage with MMAs outstanding"
            self.stage_content[s] = t
            for _ in range(self.rng.randrange(3)):
                yield
            self.kv_full[s].arrive()
            yield

    def softmax(self, i):
        for j in range(self.n):
            while not self.s_full[i].test(j & 1):
                yield
            assert self.qk_done[i] == [j + 1, j + 1], f"softmax {i} reads S of tile {j} before both halves executed"
            for _ in range(self.rng.randrange(3)):
                yield
            self.s_read[i] = j + 1
            if self.var & 4:
                self.s_cons[i].arrive()
            for _ in range(self.rng.randrange(4)):
                yield
            self.p_written[i][0] = j + 1
            if self.var & 2:
                self.p_half[i].arrive()
            for _ in range(self.rng.randrange(4)):
                yield
            self.p_written[i][1] = j + 1
            self.p_ready[i].arrive()
            yield
        while not self.o_final[i].test(0):
            yield
        assert self.pv_done[i] == [self.n, self.n]
        self.finished[i] = True

    def exact(self, bar, done_expected):
        assert bar.done == done_expected, f"parity test passed on a stale phase ({bar.done} != {done_expected})"

    def issuer(self):
        n = self.n
        while not self.kv_full[0].test(0):
            yield
        for i in range(2):
            self.qk(i, 0, (0, 1))
            self.commit(self.s_full[i])
        self.commit(self.kv_empty[0])
        if self.var & 4:
            tile, step = [0, 0], [0, 0]
            v_uses, k_uses = 0, 0

            def stage_ready(t):
                ok = self.kv_full[t % K_STAGES].test((t // K_STAGES) & 1)
                if ok:
                    self.exact(self.kv_full[t % K_STAGES], t // K_STAGES + 1)
                return ok
            while tile[0] < n or tile[1] < n:
                for i in range(2):
                    j = tile[i]
                    if j >= n:
                        continue
                    if step[i] == 0:
                        if j + 1 >= n:
                            step[i] = 1
                        elif self.s_cons[i].test(j & 1) and stage_ready(2 * j + 2):
                            self.exact(self.s_cons[i], j + 1)
                            self.qk(i, j + 1, (1,))
                            step[i] = 1
                    elif step[i] == 1:
                        if self.p_half[i].test(j & 1) and stage_ready(2 * j + 1):
                            self.exact(self.p_half[i], j + 1)
                            self.pv(i, j, 0)
                            step[i] = 2
                    elif self.p_ready[i].test(j & 1):
                        self.exact(self.p_ready[i], j + 1)
                        self.pv(i, j, 1)
                        sh = 4 * (j & 3)
                        v_uses += 1 << sh
                        if (v_uses >> sh) & 15 == 2:
                            v_uses &= ~(15 << sh)
                            self.commit(self.kv_empty[(2 * j + 1) % K_STAGES])
                        if j + 1 < n:
                            self.qk(i, j + 1, (0,))
                            self.commit(self.s_full[i])
                            k_uses += 1 << sh
                            if (k_uses >> sh) & 15 == 2:
                                k_uses &= ~(15 << sh)
                                self.commit(self.kv_empty[(2 * j + 2) % K_STAGES])
                        else:
                            self.commit(self.o_final[i])
                        tile[i], step[i] = j + 1, 0
                yield
            return
        for j in range(n):                     # fixed order (VAR 0 / 1: one hand-off; VAR 2 / 3: two halves)
            tv, tk = 2 * j + 1, 2 * j + 2
            while not self.kv_full[tv % K_STAGES].test((tv // K_STAGES) & 1):
                yield
            for i in range(2):
                if self.var & 2:
                    while not self.p_half[i].test(j & 1):
                        yield
                    self.exact(self.p_half[i], j + 1)
                    self.pv(i, j, 0)
                while not self.p_ready[i].test(j & 1):
                    yield
                self.exact(self.p_ready[i], j + 1)
                if not self.var & 2:
                    self.pv(i, j, 0)
                self.pv(i, j, 1)
                if i == 1:
                    self.commit(self.kv_empty[tv % K_STAGES])
                if j + 1 < n:
                    if i == 0:
                        while not self.kv_full[tk % K_STAGES].test((tk // K_STAGES) & 1):
                            yield
                    self.qk(i, j + 1, (0, 1))
                    self.commit(self.s_full[i])
                    if i == 1:
                        self.commit(self.kv_empty[tk % K_STAGES])
                else:
                    self.commit(self.o_final[i])

    def run(self):
        actors = [self.producer(), self.softmax(0), self.softmax(1), self.issuer()]
        alive = [True] * len(actors)
        idle = 0
        while not all(self.finished):
            k = self.rng.randrange(len(actors) + 1)
            progressed = False
            if k == len(actors):
                progressed = self.pipe_step()
            elif alive[k]:
                before = self.snapshot()
                try:
                    next(actors[k])
                except StopIteration:
                    alive[k] = False
                progressed = before != self.snapshot() or not alive[k]
            idle = 0 if progressed else idle + 1
            assert idle < 5000, f"deadlock: {self.snapshot()}"

    def snapshot(self):
        bars = self.kv_full + self.kv_empty + self.s_full + self.p_half + self.p_ready + self.s_cons + self.o_final
        return (tuple(b.done for b in bars), len(self.pipe), tuple(self.s_read), tuple(map(tuple, self.p_written)),
                tuple(self.stage_content))


@pytest.mark.parametrize("var", [0, 2, 6])
@pytest.mark.parametrize("n", [1, 2, 3, 5, 9])
def test_hand_off_protocol_is_safe_and_live(var, n):
    for seed in range(60):
        Sim(n, var, random.Random(1000 * var + 17 * n + seed)).run()


def test_model_catches_a_broken_protocol():
    """Sanity of the model itself: a producer that ignores the stage-release barriers must be flagged."""
    class NoBackpressure(Sim):
        def producer(self):
            for t in range(2 * self.n):
                s = t % K_STAGES
                assert self.stage_reads[s] == 0, "producer overwrites a stage with MMAs outstanding"
                self.stage_content[s] = t
                self.kv_full[s].arrive()
                yield
    with pytest.raises(AssertionError):
        for seed in range(40):
            NoBackpressure(9, 6, random.Random(seed)).run()
