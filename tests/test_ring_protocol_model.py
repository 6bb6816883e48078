"""The ring protocol of the evaluation kernels (vmat_eval_kernels.cuh: ring_push, consume_run / consume_stages)
as a host model, run under random interleavings: one producer refilling `nst` slots in order, W consumer warps
that take a header stage and then the super-slice's data stages in batches of 4 / 2 / 1, phase-parity waits on the
`full` / `empty` barriers with the hardware's semantics (a parity test passes when the barrier's current phase
differs from the tested parity - so a wait that is one whole phase off passes too, which is what a wrong protocol
would trip over), and TMA copies that land in any order.  CPU only.

What it pins: (stage, phase) bookkeeping across ring wraps inside a batch, look-ahead waits of up to three stages,
batches against shallow rings, zero-length super-slices.  What it cannot see: anything about WHEN a load is
performed - the bug this round found (an arrive overtaking header loads still in flight) was
below this model, which is why the model was written first: it cleared the protocol."""
import random

import pytest


def simulate(seed, nst, W, batch, nss=40, broken=None):
    rng = random.Random(seed)
    nchs = [rng.choice([0, 1, 2, 3, 4, 5, 7, 19]) for _ in range(nss)]
    stages = []
    for i, n in enumerate(nchs):
        stages.append(("H", n))
        stages.extend(("D", 0) for _ in range(n))
    stages.append(("E", 0))
    full_c, empty_c, empty_arr = [0] * nst, [0] * nst, [0] * nst      # completed phases / arrivals of the current phase
    content = [None] * nst
    inflight = []

    def passes(completed, parity):                                     # mbarrier.try_wait.parity
        return (completed & 1) != parity

    def advance(o):
        o["stage"] += 1
        if o["stage"] == nst:
            o["stage"] = 0
            o["phase"] ^= 1

    def release(s):
        empty_arr[s] += 1
        if empty_arr[s] == W:
            empty_arr[s] = 0
            empty_c[s] += 1

    prod = dict(i=0, stage=0, phase=0)
    cons = [dict(stage=0, phase=0, idx=0, left=0) for _ in range(W)]
    done = [False] * W
    for _ in range(400_000):
        if all(done):
            return "ok"
        acts = [w for w in range(W) if not done[w]]
        if prod["i"] < len(stages):
            acts.append("P")
        if inflight:
            acts.append("L")
        a = rng.choice(acts)
        if a == "P":                                                   # ring_push: wait empty (parity ^ 1), issue the copy
            s = prod["stage"]
            if passes(empty_c[s], prod["phase"] ^ 1) or broken == "producer_does_not_wait":
                inflight.append((s, prod["i"]))
                prod["i"] += 1
                advance(prod)
        elif a == "L":                                                 # a copy lands (any order), completes the full phase
            s, i = inflight.pop(rng.randrange(len(inflight)))
            content[s] = i
            full_c[s] += 1
        else:
            o = cons[a]
            i = o["idx"]
            kind, n = stages[i]
            if kind in ("H", "E"):
                s = o["stage"]
                if passes(full_c[s], o["phase"]):
                    if content[s] != i:
                        return f"warp {a} read stage {content[s]} where the header {i} should be (slot {s})"
                    if kind == "E":
                        done[a] = True
                        continue
                    release(s)
                    advance(o)
                    o["idx"] += 1
                    o["left"] = n
            else:
                left = o["left"]
                U = 4 if (batch >= 4 and left >= 4) else 2 if (batch >= 2 and left >= 2) else 1
                t = dict(stage=o["stage"], phase=o["phase"])
                st, ph = [], []
                for _u in range(U):
                    st.append(t["stage"])
                    ph.append(o["phase"] if broken == "one_parity_per_batch" else t["phase"])
                    advance(t)
                if all(passes(full_c[st[u]], ph[u]) for u in range(U)):
                    for u in range(U):
                        if content[st[u]] != i + u:
                            return f"warp {a} read stage {content[st[u]]} where data stage {i + u} should be (slot {st[u]})"
                    for u in range(U):
                        release(st[u])
                    o.update(stage=t["stage"], phase=t["phase"], idx=i + U, left=left - U)
    return "no progress"


@pytest.mark.parametrize("nst", [4, 5, 8, 10, 12, 16])
@pytest.mark.parametrize("W", [1, 4])
@pytest.mark.parametrize("batch", [1, 2, 4])
def test_ring_protocol_holds_under_random_interleavings(nst, W, batch):
    for seed in range(12):
        assert simulate(seed, nst, W, batch) == "ok"


def test_model_notices_a_broken_protocol():
    """The same model with a deliberate bug must fail - otherwise the test above proves nothing."""
    outcomes = {simulate(seed, 5, 2, 4, broken="one_parity_per_batch") for seed in range(20)}
    assert outcomes != {"ok"}
    outcomes = {simulate(seed, 8, 4, 2, broken="producer_does_not_wait") for seed in range(20)}
    assert outcomes != {"ok"}
