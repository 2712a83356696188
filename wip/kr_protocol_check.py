"""Randomised interleaving check of the K-round pipeline's synchronisation (wip/0002-k-round-pipeline.patch).

Not a test of the product: a paper model of the barrier protocol, run on the CPU, that looks for deadlocks
(mbarrier parity waits that can never pass) and TMEM / shared-memory hazards (an epilogue warp reading an
accumulator an MMA may still write, an MMA overwriting an accumulator some warp has not finished reading, an
activation chunk rewritten while an issued MMA may still read it) under random schedules.

Model: 16 epilogue warps (quadrant q = w % 4, column split cs = w // 4), one MMA warp whose MMAs complete
asynchronously and in order (tcgen05.commit fires when everything issued before it has completed), mbarriers with
hardware semantics (a waiter only sees the parity of the current phase).
"""
import random
import sys

NK = [2, 13, 13, 13, 13]  # K steps (16 columns each) of the A operand of layers 0..4 at config 2
NLAYERS = len(NK)
L = NLAYERS - 1           # output layer
STEPS = 4                 # horizon steps simulated per run
TMEM_KSTEPS = 12
NWARPS = 16


def rounds(nk):
    return (nk + 3) // 4


class Bar:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, "more arrivals than the barrier expects in one phase"
        if self.pending == 0:
            self.pending, self.phase = self.count, self.phase + 1

    def passed(self, parity):  # try_wait.parity: true once the phase with this parity is no longer the current one
        return (self.phase & 1) != parity


class Sim:
    def __init__(self, rng):
        self.rng = rng
        self.bar_k = [Bar(NWARPS) for _ in range(4)]
        self.bar_acc = Bar(1)
        self.cta_bar = Bar(NWARPS)          # bar.sync among the epilogue warps (end of step)
        self.inflight = []                  # issued, not yet completed MMAs / commits, in order
        self.acc_tag = [None, None]         # layer (global index g) whose complete result sits in accumulator x
        self.acc_writer = [None, None]      # g of the layer currently accumulating into accumulator x
        self.acc_reads_left = [0, 0]        # chunk reads of accumulator x's current result still to come
        self.act_tag = {}                   # chunk -> g of the layer this activation chunk is the input of
        self.errors = []

    # ---- asynchronous tensor pipe ------------------------------------------------------------------------
    def pipe_step(self):
        if not self.inflight:
            return False
        op = self.inflight.pop(0)
        if op[0] == "mma":
            _, g, kk, last = op
            chunk = kk
            if self.act_tag.get(chunk) != g:
                self.errors.append(f"MMA layer {g} K step {kk} read activation chunk holding {self.act_tag.get(chunk)}")
            if last:
                self.acc_tag[g & 1] = g
                self.acc_writer[g & 1] = None
                self.acc_reads_left[g & 1] = expected_reads(g)
        else:
            self.bar_acc.arrive()
        return True

    def mma_issue(self, g, kk, last):
        x = g & 1
        if kk == 0:
            if self.acc_reads_left[x] != 0:
                self.errors.append(f"layer {g} starts accumulating into accumulator {x} with {self.acc_reads_left[x]} reads of layer {self.acc_tag[x]} outstanding")
            self.acc_writer[x] = g
            self.acc_tag[x] = None
        self.inflight.append(("mma", g, kk, last))

    # ---- epilogue side -----------------------------------------------------------------------------------
    def write_act(self, chunk, g_next):
        for op in self.inflight:
            if op[0] == "mma" and op[2] == chunk:
                self.errors.append(f"activation chunk {chunk} rewritten for layer {g_next} while an MMA of layer {op[1]} may still read it")
        self.act_tag[chunk] = g_next

    def read_acc(self, g):
        x = g & 1
        if self.acc_tag[x] != g:
            self.errors.append(f"epilogue of layer {g} read accumulator {x} holding {self.acc_tag[x]} (writer {self.acc_writer[x]})")
        self.acc_reads_left[x] -= 1


def expected_reads(g):
    l = g % NLAYERS
    if l == L:
        return NWARPS  # every warp reads its output groups once (modelled as one read)
    return sum(1 for w in range(NWARPS) for r in range(rounds(NK[l + 1])) if 4 * r + w // 4 < NK[l + 1])


def mma_warp(sim):
    kpar = [0, 0, 0, 0]
    g = 0
    for _ in range(STEPS):
        for l in range(NLAYERS):
            nk = NK[l]
            for r in range(rounds(nk)):
                while not sim.bar_k[r].passed(kpar[r]):
                    yield "wait"
                kpar[r] ^= 1
                for kk in range(4 * r, min(4 * r + 4, nk)):
                    sim.mma_issue(g, kk, kk == nk - 1)
                    yield "issued"
                if r == rounds(nk) - 1:
                    sim.inflight.append(("commit", g))
            g += 1


def epilogue_warp(sim, w):
    cs = w // 4
    acc_par = 0
    g = 0

    def build_input(g0):
        nk0 = NK[0]
        for c in range(cs, nk0, 4):
            sim.write_act(c, g0)
        for r in range(rounds(nk0)):
            sim.bar_k[r].arrive()

    build_input(0)
    yield "built"
    for _ in range(STEPS):
        for l in range(L):
            while not sim.bar_acc.passed(acc_par):
                yield "wait"
            acc_par ^= 1
            nk_next = NK[l + 1]
            for r in range(rounds(nk_next)):
                c = 4 * r + cs
                if c < nk_next:
                    sim.read_acc(g)
                    yield "read"
                    sim.write_act(c, g + 1)
                sim.bar_k[r].arrive()
                yield "arrived"
            g += 1
        while not sim.bar_acc.passed(acc_par):
            yield "wait"
        acc_par ^= 1
        sim.read_acc(g)
        g += 1
        yield "output"
        phase = sim.cta_bar.phase
        sim.cta_bar.arrive()
        while sim.cta_bar.phase == phase:
            yield "wait"
        build_input(g)
        yield "built"


def run(seed):
    rng = random.Random(seed)
    sim = Sim(rng)
    agents = [mma_warp(sim)] + [epilogue_warp(sim, w) for w in range(NWARPS)]
    alive = list(range(len(agents)))
    idle = 0
    while alive:
        if rng.random() < 0.3 and sim.pipe_step():
            idle = 0
            continue
        a = rng.choice(alive)
        try:
            what = next(agents[a])
        except StopIteration:
            alive.remove(a)
            idle = 0
            continue
        if what == "wait":
            idle += 1
            if idle > 20000 and not sim.inflight:
                return [f"deadlock: agents {alive} all waiting (seed {seed})"]
            if idle > 20000:
                sim.pipe_step()
        else:
            idle = 0
        if sim.errors:
            return sim.errors
    return sim.errors


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    bad = 0
    for seed in range(n):
        errs = run(seed)
        if errs:
            bad += 1
            print("seed", seed, errs[:3])
    print(f"{n} random schedules, {bad} with problems")
