"""CPU model of the parallel ziggurat chain resolution of csrc/rng.cu (zig_is_head / zig_starts_at / zig_chain0 /
zig_enter), checked against the plain sequential walk on ADVERSARIAL synthetic segments: non-simple densities up to
50 % and attempts consuming up to 33 draws -- patterns the real tables produce once in 10^6 segments or never, so the
GPU-vs-NumPy tests cannot reach them.  The model follows the device functions statement by statement; it guards the
algorithm (window of 32, head detection, merge of the entry-e chain into the entry-0 chain), not the CUDA code."""
import random

SEG = 512          # the device uses 4096; the logic does not depend on it
MAX_CONS = 33      # largest consumption the device lets through (tail loop bound, see ZIG_MAX_CONS)
ENTRIES = 16


def sequential(ns, cons, yes, entry):
    """The reference order of events: walk the attempts one after the other."""
    p, count, starts = entry, 0, []
    while p < SEG:
        starts.append(p)
        if ns[p]:
            count += yes[p]
            p += cons[p]
        else:
            count += 1
            p += 1
    return count, p - SEG, starts


def is_head(ns, cons, q):
    prev = -1
    for r in range(q - 1, max(q - 33, -1), -1):      # the 32 positions before q, nearest first
        if ns[r]:
            if prev < 0:
                prev = r
            if r + cons[r] > q:
                return False, prev
    return True, prev


def next_ns(ns, p):
    while p < SEG and not ns[p]:
        p += 1
    return min(p, SEG)


def starts_at(ns, cons, q):
    h = q
    while True:
        head, prev = is_head(ns, cons, h)
        if head:
            break
        h = prev
    p = h
    while p < q:
        p = next_ns(ns, p + cons[p])
    return p == q


def chain0(ns, cons, yes):
    covered = [False] * SEG
    reach = 0
    for q in range(SEG):
        if ns[q] and starts_at(ns, cons, q):
            for k in range(q + 1, min(q + cons[q], SEG)):
                covered[k] = True
            reach = max(reach, q + cons[q])
    ym = [(not covered[p]) and ((not ns[p]) or yes[p]) for p in range(SEG)]
    return covered, ym, max(reach, SEG) - SEG


def enter(ns, cons, yes, covered, ym, exit0, e):
    p, cnt = e, 0
    ym = list(ym)
    for k in range(e):
        ym[k] = False
    while p < SEG and covered[p]:
        step, y = (cons[p], yes[p]) if ns[p] else (1, 1)
        for k in range(p, min(p + step, SEG)):
            ym[k] = False
        if y:
            ym[p] = True
        cnt += y
        p += step
    if p >= SEG:
        return cnt, p - SEG, ym
    return cnt + sum(ym[p:]), exit0, ym     # device: cnt + total0 - yields0 before p; ym[p:] is untouched there


def random_segment(rng, density, long_share):
    ns = [rng.random() < density for _ in range(SEG)]
    cons = [1] * SEG
    yes = [1] * SEG
    for p in range(SEG):
        if ns[p]:
            cons[p] = rng.choice([3, 5, 9, 17, 33]) if rng.random() < long_share else 2
            yes[p] = 1 if cons[p] > 2 else int(rng.random() < 0.6)   # tails always yield, wedges sometimes
    return ns, cons, yes


def test_parallel_resolution_equals_sequential_walk_on_adversarial_segments():
    rng = random.Random(2024)
    for density, long_share, reps in ((0.012, 0.02, 40), (0.1, 0.2, 60), (0.3, 0.3, 60), (0.5, 0.5, 40), (0.9, 0.1, 20)):
        for _ in range(reps):
            ns, cons, yes = random_segment(rng, density, long_share)
            covered, ym, exit0 = chain0(ns, cons, yes)
            c0, e0, starts0 = sequential(ns, cons, yes, 0)
            assert sum(ym) == c0 and exit0 == e0
            assert [p for p in range(SEG) if not covered[p]] == starts0
            for e in range(1, ENTRIES):
                ce, ee, starts_e = sequential(ns, cons, yes, e)
                cnt, ex, ym_e = enter(ns, cons, yes, covered, ym, exit0, e)
                assert (cnt, ex) == (ce, ee), (density, e)
                want = [False] * SEG
                for p in starts_e:
                    want[p] = (not ns[p]) or bool(yes[p])
                assert ym_e == want, (density, e)
