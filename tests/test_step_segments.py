"""Host bookkeeping of the AdamW step counters (flair_hub/tasks/tasks_module.py StepSegments): after any sequence of steps in
which whole ranges of the parameter arena received no gradient, every arena element's counter must equal what torch's
per-parameter ``state['step']`` would be -- the number of steps in which it DID receive one."""
import numpy as np

from flair_for_aigle_b200.flair_hub.tasks.tasks_module import StepSegments


def test_segments_follow_per_parameter_step_counters():
    rng = np.random.default_rng(0)
    n = 1000
    ranges = [(0, 400), (400, 700), (700, 1000), (100, 250)]          # two "encoders", "the rest", and an odd inner range
    for trial in range(20):
        seg = StepSegments(n)
        want = np.zeros(n, dtype=np.int64)
        for step in range(12):
            skip = [ranges[i] for i in range(len(ranges)) if rng.random() < 0.3]
            todo = seg.advance(skip)
            skipped = np.zeros(n, dtype=bool)
            for lo, hi in skip:
                skipped[lo:hi] = True
            want[~skipped] += 1
            # the segments partition the arena, in order, and every updated range carries its own counter
            assert seg.items[0][0] == 0 and seg.items[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(seg.items, seg.items[1:]))
            got = np.zeros(n, dtype=np.int64)
            for lo, hi, k in seg.items:
                got[lo:hi] = k
            assert np.array_equal(got, want)
            updated = np.zeros(n, dtype=bool)
            for lo, hi, k in todo:
                assert (want[lo:hi] == k).all()
                updated[lo:hi] = True
            assert np.array_equal(updated, ~skipped)
        assert len(seg) <= 2 * len(ranges) + 1


def test_no_skip_keeps_one_segment():
    seg = StepSegments(10)
    for k in range(1, 5):
        assert seg.advance() == [(0, 10, k)]
    assert len(seg) == 1
    seg.split(0)
    seg.split(10)
    assert len(seg) == 1
