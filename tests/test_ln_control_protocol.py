"""CPU: the hand-off protocol between the LayerNorm control warp and the four LayerNorm warps of a chained launch
(easy_vitpose_b200/csrc/chain.cuh, `warp == 9 && p.ln_ctl`), restated as a discrete-event model and driven with random timings.

Device side: the control warp keeps a poll cursor A and a publish cursor B over the CTA's job list.  Per loop iteration it
(1) probes the source counter of job A if A <= B + 1 and, when complete, arrives on the "ready" barrier of slot A & 1 and advances
A; (2) probes the "done" mbarrier of slot B & 1 if B < A and, when complete, publishes job B and advances B.  The LayerNorm warps
wait for ready(seq), do their rows, arrive on done(seq), in order.  What must hold for ANY timing:
  * a slot's barrier is never signalled for job i + 2 before every party has consumed job i (no phase overrun);
  * every job is announced once and published once, in order, and the model terminates (no dead wait);
  * a finished job is published even while the NEXT job's source rows are still outstanding (the reason neither probe blocks:
    the first job of LayerNorm stage 1 waits for fc2 tiles that may themselves wait for this CTA's last job of stage 0)."""
import random

import pytest


def simulate(n_jobs, source_ready_at, ln_cost, warps=4, seed=0, source_needs_publish=None):
    """Time advances one control-loop iteration per tick.  source_ready_at[j]: tick from which job j's source counter is complete
    (or None when it depends on a publish, see source_needs_publish = {job: job_that_must_be_published_first})."""
    rng = random.Random(seed)
    a = b = 0
    ready_signalled, published = [], []
    ready_phase = [0, 0]                      # completed phases per "ready" slot
    done_arrivals = [0, 0]                    # arrivals on the current phase of each "done" slot
    done_phase = [0, 0]
    warp_seq = [0] * warps                    # next job each LayerNorm warp will take
    warp_busy_until = [0] * warps
    warp_waiting_ready = [True] * warps
    publish_tick = {}
    tick = 0
    while b < n_jobs:
        tick += 1
        assert tick < 100000, "the protocol does not terminate"
        # ---- LayerNorm warps
        for w in range(warps):
            seq = warp_seq[w]
            if seq >= n_jobs:
                continue
            if warp_waiting_ready[w]:
                if ready_phase[seq & 1] > seq >> 1:                       # phase (seq >> 1) of the slot has completed
                    assert ready_phase[seq & 1] == (seq >> 1) + 1, "ready barrier ran a phase ahead of a waiting warp"
                    warp_waiting_ready[w] = False
                    warp_busy_until[w] = tick + rng.randint(1, ln_cost)
            elif tick >= warp_busy_until[w]:
                slot = seq & 1
                assert done_phase[slot] == seq >> 1, "arrival on a done barrier whose phase was not consumed yet"
                done_arrivals[slot] += 1
                if done_arrivals[slot] == warps:
                    done_arrivals[slot] = 0
                    done_phase[slot] += 1
                warp_seq[w] += 1
                warp_waiting_ready[w] = True
        # ---- control warp, one loop iteration
        if a < n_jobs and a <= b + 1:
            need = (source_needs_publish or {}).get(a)
            src_ok = (need in publish_tick) if need is not None else tick >= source_ready_at[a]
            if src_ok:
                slot = a & 1
                assert ready_phase[slot] == a >> 1, "ready signalled out of order"
                assert a < 2 or (a - 2) in publish_tick, "ready(i + 2) before done(i) was seen"
                ready_phase[slot] += 1
                ready_signalled.append(a)
                a += 1
        if b < a and done_phase[b & 1] > b >> 1:
            assert done_phase[b & 1] == (b >> 1) + 1, "done barrier ran a phase ahead of the control warp"
            published.append(b)
            publish_tick[b] = tick
            b += 1
    assert ready_signalled == list(range(n_jobs)) and published == list(range(n_jobs))
    return publish_tick


@pytest.mark.parametrize("seed", range(25))
def test_random_timings(seed):
    rng = random.Random(1000 + seed)
    n = rng.randint(1, 14)
    t, src = 0, []
    for _ in range(n):
        t += rng.choice([0, 0, 1, 5, 40])
        src.append(t)
    rng.shuffle(src)                                                      # later jobs may become ready before earlier ones
    simulate(n, src, ln_cost=rng.choice([1, 3, 12, 60]), seed=seed)


def test_finished_job_is_published_while_the_next_source_is_outstanding():
    """Job 2's source only completes after job 1 has been published (stage 1's first job behind stage 0's last one): a control
    warp that blocked on the poll of job 2 before publishing job 1 would never finish."""
    ticks = simulate(4, [0, 0, None, 0], ln_cost=5, source_needs_publish={2: 1})
    assert ticks[1] < ticks[2]


def test_single_job_and_no_jobs():
    simulate(1, [3], ln_cost=2)
    assert simulate(0, [], ln_cost=2) == {}
