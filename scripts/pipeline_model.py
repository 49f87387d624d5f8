"""Discrete-time model of one CTA's steady state (issuer thread, in-order tensor pipe, softmax warpgroups), fed with
the constants measured on B200 in round 1.  A planning tool: it is calibrated on three measured
periods (D=128 default 3250, D=64 default ~2750, D=64 with separate P columns ~2800 cycles per KV tile pair) and is
then asked what the column-split softmax (csrc/fa_fwd_sm100_colsplit.cuh) should do.  No GPU needed.

What it is good for, and what not (end of round 1): with the segment durations taken from the in-kernel trace it
reproduces the D=128 period (3290 vs 3250 measured); it is NOT accurate for D=64 (3070 / 2530 vs 2750 / 2800 measured
for aliased / separate P): the trace durations include ~250 cycles of tracing overhead per tile and the contention
between the two softmax warps of a sub-partition is cruder in the model than on the machine.  Use it for ORDERING
questions (who waits for whom), not for 5 % decisions.  Its answer for the column-split kernel: ~3190, i.e. softmax
bound at 2 x (ld+max+vote 390 + exponentials 1055 + hand-offs 120) -- a few percent better than the default, not the
tensor-bound 2800 a back-of-envelope estimate suggested, because the per-tile fixed costs no longer overlap with the
other tile's exponentials and cannot be prefetched either (S_t(j+1) completes ~1260 cycles after P_t(j) is handed
over, about one tile-slot later: just in time, never early).

    python scripts/pipeline_model.py

Model.  Time advances in steps of DT cycles.
  * tensor pipe: FIFO of (duration, barriers-to-complete); one batch at a time; a commit becomes visible L_COMMIT
    cycles after the batch ends;
  * issuer: a sequential program of `wait(barrier)` and `issue(batch)`; after a wait is satisfied it needs L_WAKE
    cycles before the next instruction takes effect (try_wait return + fence + descriptor arithmetic + elect);
  * softmax agents: sequential programs of `wait(barrier)`, `work(cycles)` and `arrive(barrier)`; while TWO agents that
    share the SM sub-partitions are both in a `work` phase each advances at rate CONTENTION (two warps per
    sub-partition get 2 x 1/2109 rows per cycle instead of 1/1518 each: 0.72 of the solo rate each).
"""
import collections

DT = 2
L_COMMIT = 30
L_WAKE = 40
CONTENTION = 1518.0 / 2109.0     # applies only while BOTH agents are in an exponential ("exp") phase


class Sim:
    def __init__(self, ntiles):
        self.t = 0
        self.done = collections.defaultdict(lambda: None)     # barrier name -> completion time
        self.queue = collections.deque()                      # tensor FIFO
        self.busy_until = 0
        self.cur = None
        self.pending = []                                     # (time, barrier) commits in flight
        self.tensor_busy = 0
        self.ntiles = ntiles

    def complete(self, name, when):
        if self.done[name] is None:
            self.done[name] = when

    def ready(self, name):
        d = self.done[name]
        return d is not None and d <= self.t


class Agent:
    """prog: list of ('wait', name) | ('work', cycles) | ('arrive', name, count_key) | ('issue', dur, [names])"""

    def __init__(self, sim, prog, is_softmax):
        self.sim, self.prog, self.pc = sim, prog, 0
        self.left = 0.0
        self.hold = 0
        self.is_softmax = is_softmax

    def in_work(self):
        return self.pc < len(self.prog) and self.prog[self.pc][0] == "work"

    def in_exp(self):
        return self.in_work() and len(self.prog[self.pc]) > 2 and self.prog[self.pc][2] == "exp"

    def step(self, rate, arrivals):
        sim = self.sim
        budget = DT
        while self.pc < len(self.prog) and budget > 0:
            op = self.prog[self.pc]
            if self.hold > 0:
                d = min(self.hold, budget)
                self.hold -= d
                budget -= d
                continue
            if op[0] == "wait":
                if sim.ready(op[1]):
                    self.pc += 1
                    self.hold = op[2] if len(op) > 2 else 0
                else:
                    return
            elif op[0] == "work":
                if self.left <= 0:
                    self.left = float(op[1])
                adv = budget * rate
                if adv >= self.left:
                    budget -= self.left / rate
                    self.left = 0
                    self.pc += 1
                else:
                    self.left -= adv
                    budget = 0
            elif op[0] == "arrive":
                arrivals[op[1]] += 1
                if arrivals[op[1]] >= op[2]:
                    sim.complete(op[1], sim.t + 10)            # mbarrier arrive -> waiter sees it
                self.pc += 1
            elif op[0] == "issue":
                sim.queue.append((op[1], op[2]))
                self.pc += 1
                self.hold = 8 * max(1, op[1] // 100)           # a few cycles of issue per MMA
        return


def run(issuer_prog, softmax_progs, horizon=400000):
    sim = Sim(0)
    issuer = Agent(sim, issuer_prog, False)
    sms = [Agent(sim, p, True) for p in softmax_progs]
    arrivals = collections.defaultdict(int)
    while sim.t < horizon:
        # tensor pipe
        if sim.cur is None and sim.queue:
            dur, names = sim.queue.popleft()
            sim.cur = (sim.t + dur, names)
        if sim.cur is not None:
            sim.tensor_busy += DT
            if sim.t >= sim.cur[0]:
                for n in sim.cur[1]:
                    sim.complete(n, sim.t + L_COMMIT)
                sim.cur = None
        nexp = sum(1 for a in sms if a.in_exp())
        rate = CONTENTION if nexp >= 2 else 1.0
        issuer.step(1.0, arrivals)
        for a in sms:
            a.step(rate if a.in_exp() else 1.0, arrivals)
        sim.t += DT
        if issuer.pc >= len(issuer.prog) and all(a.pc >= len(a.prog) for a in sms) and sim.cur is None and not sim.queue:
            break
    return sim


def default_kernel(n, S, PV, ld_max, check, seg, separate_p=False):
    """fa_fwd_sm100_kernel: one warpgroup per Q tile, 3-stage hand-off (PV split 4/2/2 of 8 k-steps)."""
    iss = []
    for t in (0, 1):
        iss.append(("issue", S, [f"s{t}_0"]))
    for j in range(n):
        for t in (0, 1):
            iss.append(("wait", f"ph{t}_{j}", L_WAKE))
            iss.append(("issue", PV // 2, []))
            if separate_p and j + 1 < n:
                iss.append(("issue", S, [f"s{t}_{j + 1}"]))
            iss.append(("wait", f"p3{t}_{j}", L_WAKE))
            iss.append(("issue", PV // 4, []))
            iss.append(("wait", f"pf{t}_{j}", L_WAKE))
            iss.append(("issue", PV // 4, [f"pv{t}_{j}"]))
            if not separate_p and j + 1 < n:
                iss.append(("issue", S, [f"s{t}_{j + 1}"]))
    progs = []
    for t in (0, 1):
        p = []
        for j in range(n):
            p.append(("wait", f"s{t}_{j}", 30))
            p.append(("work", ld_max + check))
            if separate_p and j > 0:
                p.append(("wait", f"pv{t}_{j - 1}", 20))
            p.append(("work", seg[0], "exp"))
            p.append(("arrive", f"ph{t}_{j}", 1))
            p.append(("work", seg[1], "exp"))
            p.append(("arrive", f"p3{t}_{j}", 1))
            p.append(("work", seg[2], "exp"))
            p.append(("arrive", f"pf{t}_{j}", 1))
        progs.append(p)
    return iss, progs


def colsplit_kernel(n, S, PV, ld_max, check, exps_tile, handoff):
    """fa_fwd_sm100_colsplit_kernel: all eight softmax warps on every tile (one agent), hand-off in two halves."""
    iss = []
    for t in (0, 1):
        iss.append(("issue", S, [f"s{t}_0"]))
    for j in range(n):
        for t in (0, 1):
            iss.append(("wait", f"pa{t}_{j}", L_WAKE))
            iss.append(("issue", PV // 2, []))
            iss.append(("wait", f"pb{t}_{j}", L_WAKE))
            iss.append(("issue", PV // 2, []))
            if j + 1 < n:
                iss.append(("issue", S, [f"s{t}_{j + 1}"]))
    p = []
    for j in range(n):
        for t in (0, 1):
            p.append(("wait", f"s{t}_{j}", 30))
            p.append(("work", ld_max + check))
            p.append(("work", exps_tile * 0.5 + handoff, "exp"))
            p.append(("arrive", f"pa{t}_{j}", 1))
            p.append(("work", exps_tile * 0.5 + handoff, "exp"))
            p.append(("arrive", f"pb{t}_{j}", 1))
    return iss, [p]


def period(sim, n):
    # steady-state period: time between s0_j completions in the middle of the run
    a, b = n // 4, 3 * n // 4
    return (sim.done[f"s0_{b}"] - sim.done[f"s0_{a}"]) / (b - a)


if __name__ == "__main__":
    N = 48
    # measured (trace, profiles/r01_trace_cfg3.txt): s_full -> ld+max 490, -> rescale vote 205 (trace-inflated),
    # -> p_half 698, -> p_3q 368, -> p_full 425
    SEG = (698, 368, 425)
    cases = [
        ("D=128 default (measured 3250)", default_kernel(N, 816, 584, 490, 205, SEG)),
        ("D=64  default (measured ~2750)", default_kernel(N, 408, 368, 490, 205, SEG)),
        ("D=64  separate P (measured ~2800)", default_kernel(N, 408, 368, 490, 205, SEG, separate_p=True)),
        # column split: 64 columns per thread: two x32 loads (~240 incl. latency), 32 FMNMX3 (~90), bar.red (~60);
        # exponentials of a 128-key tile with two warps per sub-partition: 2109 / 2 = 1055
        ("D=128 column-split (prediction)", colsplit_kernel(N, 816, 584, 330, 60, 1055, 60)),
        ("D=64  column-split (prediction)", colsplit_kernel(N, 408, 368, 330, 60, 1055, 60)),
    ]
    for name, (iss, progs) in cases:
        sim = run(iss, progs)
        print(f"{name:40s} period {period(sim, N):7.0f} cycles per KV tile pair   tensor busy "
              f"{100.0 * sim.tensor_busy / sim.t:5.1f} %")
