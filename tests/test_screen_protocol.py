"""Model check of the tcgen05 screen's hand-off protocol (csrc/tensor_search.cu, tensor_screen_kernel): one TMA
producer, ISS MMA-issuing threads, TEAMS epilogue teams, a ring of B stages and NBUF accumulator buffers in tensor
memory, all synchronised by mbarriers that are waited on by PARITY.

What can go wrong is not arithmetic but ordering: a waiter that is one phase ahead of an mbarrier reads the parity of the
phase before as "complete" (round 1's three-team variant deadlocked on exactly that), a stage released too early is
overwritten under a running MMA, an accumulator is refilled before its team has read it.  This test runs the kernel's
loops -- the same index walks, parities and arrival counts, transcribed from the kernel -- as cooperating coroutines
under a randomised scheduler with asynchronous TMA / tensor-pipe completion, and checks that every epilogue read sees
the unit it was waiting for, that no buffer or stage is overwritten while it is still needed, and that everything
terminates.  It also shows that the model has teeth: three issuers with two teams SHARING the per-buffer barriers
(what `NB = lcm(NBUF, TEAMS)` exists to prevent) is caught.

CPU only; no GPU, no library."""
import random

import pytest


class MBar:
    """mbarrier with `count` expected arrivals per phase; wait(parity) as mbarrier.try_wait.parity: true iff the phase
    of that parity is the immediately preceding (completed) one -- on a fresh barrier parity 1 passes, parity 0 waits"""

    def __init__(self, count):
        self.count, self.pending, self.completed = count, count, 0

    def arrive(self):
        self.pending -= 1
        if self.pending == 0:
            self.pending = self.count
            self.completed += 1

    def passed(self, parity):
        return parity != (self.completed & 1)


def lcm(a, b):
    x = a
    while x % b:
        x += a
    return x


class Screen:
    def __init__(self, nt, stages, G, SUB, NBUF, ISS, TEAMS, nb_override=None, seed=0):
        self.nt, self.stages, self.G, self.SUB, self.NBUF, self.ISS, self.TEAMS = nt, stages, G, SUB, NBUF, ISS, TEAMS
        self.nunits = nt * SUB
        self.NB = nb_override or (lcm(NBUF, TEAMS) if ISS > 1 else NBUF)
        self.rng = random.Random(seed)
        self.b_full = [MBar(1) for _ in range(stages)]
        self.b_empty = [MBar(ISS) for _ in range(stages)]
        self.acc_full = [MBar(1) for _ in range(self.NB)]
        self.acc_empty = [MBar(1) for _ in range(self.NB)]  # one arrival per team here (the kernel: one per warp of the team)
        self.stage_holds = [None] * stages     # tile group a stage holds (None while a copy is in flight)
        self.stage_readers = [0] * stages      # MMAs issued but not completed that read the stage
        self.buf_holds = [None] * NBUF         # unit an accumulator holds
        self.buf_busy = [False] * NBUF         # written by an MMA in flight, or complete and not yet read
        self.pipe = []                         # in-order tensor pipe: (kind, payload); kind 'mma' | 'commit'
        self.copies = []                       # TMA copies in flight
        self.reads = []                        # units read by the epilogue, in order per team
        self.errors = []

    # ---- agents (generators: `yield cond` suspends until cond() is true) ----
    def producer(self):
        s, ph = 0, 0
        for g0 in range(0, self.nt, self.G):
            bar = self.b_empty[s]
            yield lambda bar=bar, par=ph ^ 1: bar.passed(par)
            if self.stage_readers[s]:
                self.errors.append(f"stage {s} overwritten under {self.stage_readers[s]} running MMAs (group {g0 // self.G})")
            self.stage_holds[s] = None
            self.copies.append((s, g0 // self.G))
            s += 1
            if s == self.stages:
                s, ph = 0, ph ^ 1

    def issue(self, u, buf, stage, grp, full_bar):
        if self.buf_busy[buf]:
            self.errors.append(f"unit {u} issued into accumulator {buf} that still holds unread unit {self.buf_holds[buf]}")
        self.buf_busy[buf] = True
        self.stage_readers[stage] += 1
        self.pipe.append(("mma", (u, buf, stage, grp)))
        self.pipe.append(("commit", full_bar))

    def issuer_single(self):
        """ISS == 1: the `issue_unit` lambda, units in order, barrier = buffer = u % NBUF"""
        s, ph, j = 0, 0, 0
        for u in range(self.nunits):
            buf, sub = u % self.NBUF, u % self.SUB
            bar = self.acc_empty[buf]
            yield lambda bar=bar, par=((u // self.NBUF) & 1) ^ 1: bar.passed(par)
            if j == 0 and sub == 0:
                bar = self.b_full[s]
                yield lambda bar=bar, par=ph: bar.passed(par)
            self.issue(u, buf, s, (u // self.SUB) // self.G, self.acc_full[buf])
            if sub == self.SUB - 1:
                if j == self.G - 1 or u // self.SUB == self.nt - 1:
                    self.pipe.append(("commit", self.b_empty[s]))
                    j = 0
                    s += 1
                    if s == self.stages:
                        s, ph = 0, ph ^ 1
                else:
                    j += 1

    def issuer_multi(self, ident):
        """ISS > 1: issuer `ident` takes the units u % ISS == ident; barrier u % NB, buffer from the barrier index"""
        NB, NBUF, ISS = self.NB, self.NBUF, self.ISS
        upgrp = self.G * self.SUB
        sg, held, phg = 0, -1, 0
        nb, par = ident % NB, 0
        for u in range(ident, self.nunits, ISS):
            grp = u // upgrp
            if grp != held:
                if held >= 0:
                    self.pipe.append(("commit", self.b_empty[sg]))
                    sg += 1
                    if sg == self.stages:
                        sg, phg = 0, phg ^ 1
                bar = self.b_full[sg]
                yield lambda bar=bar, p=phg: bar.passed(p)
                held = grp
            if NB == NBUF:
                buf, eb, epar = nb, nb, par ^ 1
            else:
                buf = nb - NBUF if nb >= NBUF else nb
                eb = nb - NBUF if nb >= NBUF else nb + NBUF
                epar = par if nb >= NBUF else par ^ 1
            assert nb == u % NB and par == (u // NB) & 1 and buf == u % NBUF   # the incremental walks are the closed forms
            bar = self.acc_empty[eb]
            yield lambda bar=bar, p=epar: bar.passed(p)
            self.issue(u, buf, sg, grp, self.acc_full[nb])
            nb += ISS
            if nb >= NB:
                nb, par = nb - NB, par ^ 1
        if held >= 0:
            self.pipe.append(("commit", self.b_empty[sg]))

    def epilogue(self, team):
        """the F16 epilogue's walk (the FP32 ones use the closed forms with NB == NBUF, which the same walk reproduces)"""
        NB, NBUF, TEAMS = self.NB, self.NBUF, self.TEAMS
        nb, par = team, 0
        for u in range(team, self.nunits, TEAMS):
            bar_i, bpar = nb, par
            buf = nb if NB == NBUF else (nb - NBUF if nb >= NBUF else nb)
            if NB % TEAMS == 0 or True:
                assert buf == u % NBUF
            nb += TEAMS
            if nb >= NB:
                nb, par = nb - NB, par ^ 1
            bar = self.acc_full[bar_i]
            yield lambda bar=bar, p=bpar: bar.passed(p)
            if self.buf_holds[buf] != u:
                self.errors.append(f"team {team} waited for unit {u} and read unit {self.buf_holds[buf]} from accumulator {buf}")
            self.reads.append(u)
            self.buf_busy[buf] = False
            self.acc_empty[bar_i].arrive()

    # ---- asynchronous hardware ----
    def step_hardware(self):
        """complete one TMA copy or retire the head of the tensor pipe (in issue order); returns False if idle"""
        choices = []
        if self.copies:
            choices.append("copy")
        if self.pipe:
            choices.append("pipe")
        if not choices:
            return False
        if self.rng.choice(choices) == "copy":
            s, grp = self.copies.pop(self.rng.randrange(len(self.copies)))
            self.stage_holds[s] = grp
            self.b_full[s].arrive()
        else:
            kind, payload = self.pipe.pop(0)
            if kind == "mma":
                u, buf, stage, grp = payload
                if self.stage_holds[stage] != grp:
                    self.errors.append(f"MMA of unit {u} read stage {stage} holding group {self.stage_holds[stage]}, wanted {grp}")
                self.stage_readers[stage] -= 1
                self.buf_holds[buf] = u
            else:
                payload.arrive()
        return True

    def run(self, max_steps=2_000_000):
        agents = [self.producer()]
        agents += [self.issuer_single()] if self.ISS == 1 else [self.issuer_multi(i) for i in range(self.ISS)]
        agents += [self.epilogue(t) for t in range(self.TEAMS)]
        waiting = {}
        for a in agents:
            try:
                waiting[a] = next(a)
            except StopIteration:
                pass
        for _ in range(max_steps):
            if self.errors:
                return False
            runnable = [a for a, cond in waiting.items() if cond()]
            if not waiting and not self.pipe and not self.copies:
                return sorted(self.reads) == list(range(self.nunits))
            # hardware progresses on its own; agents are picked at random among the runnable ones
            if runnable and (self.rng.random() < 0.6 or not (self.pipe or self.copies)):
                a = self.rng.choice(runnable)
                try:
                    waiting[a] = a.send(None)
                except StopIteration:
                    del waiting[a]
            elif not self.step_hardware():
                self.errors.append(f"deadlock: {len(waiting)} agents wait, nothing in flight, {len(self.reads)} of {self.nunits} units read")
                return False
        self.errors.append("did not terminate")
        return False


# (stages, G, SUB, NBUF, ISS, TEAMS): the instantiations of tensor_screen_dispatch_t
CONFIGS = {
    "f16 short: A in TMEM, 3 buffers, 3 issuers, 3 teams (G=2)": (6, 2, 2, 3, 3, 3),
    "f16 short, k <= 13 (G=4)": (6, 4, 2, 3, 3, 3),
    "bf16 short: 4 buffers, 2 issuers, 2 teams (G=4)": (6, 4, 2, 4, 2, 2),
    "bf16 short (G=2)": (6, 2, 2, 4, 2, 2),
    "SS form, 4 buffers, 4 issuers (NNS_T_ISS_F16=4)": (6, 2, 2, 4, 4, 2),
    "64/80 columns: A in TMEM, 3 buffers, ONE issuer, 2 teams": (8, 1, 2, 3, 1, 2),
    "128/144 columns: 2 buffers of 128 references": (4, 1, 1, 2, 1, 2),
    "3 buffers, 3 issuers, TWO teams (NNS_T_TEAMS_F16=2): barrier per (buffer, team)": (6, 2, 2, 3, 3, 2),
}


@pytest.mark.parametrize("name", list(CONFIGS))
@pytest.mark.parametrize("nt", [1, 2, 3, 5, 12, 13, 37, 64])
def test_hand_off_protocol_is_ordered_and_terminates(name, nt):
    stages, G, SUB, NBUF, ISS, TEAMS = CONFIGS[name]
    for seed in range(12):
        sim = Screen(nt, stages, G, SUB, NBUF, ISS, TEAMS, seed=seed)
        ok = sim.run()
        assert ok and not sim.errors, (name, nt, seed, sim.errors[:3])


def test_the_model_catches_two_teams_sharing_a_barrier_under_several_issuers():
    """3 buffers, 3 issuers, 2 teams with ONE barrier per buffer (NB = NBUF): the teams alternate on a buffer's barrier,
    units complete out of order, and a team that is a phase ahead takes the other team's completion for its own"""
    caught = 0
    for seed in range(40):
        sim = Screen(37, 6, 2, 2, 3, 3, 2, nb_override=3, seed=seed)
        if not sim.run() or sim.errors:
            caught += 1
    assert caught > 0
