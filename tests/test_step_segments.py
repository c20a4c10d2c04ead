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


def _host_adamw(n, seed):
    """The optimizer's host-visible state without its CUDA-only constructor (the kernels are not involved here)."""
    import torch
    from flair_for_aigle_b200.flair_hub.tasks.tasks_module import AdamW
    g = torch.Generator().manual_seed(seed)
    opt = AdamW.__new__(AdamW)
    opt.arena = torch.randn(n, generator=g)
    opt.exp_avg, opt.exp_avg_sq = torch.randn(n, generator=g), torch.rand(n, generator=g)
    opt.step_count, opt.step_dev = 0, torch.zeros(1, dtype=torch.int64)
    opt._segs = StepSegments(n)
    opt.lr, opt.weight_decay, opt.betas, opt.eps = 5e-5, 0.01, (0.9, 0.999), 1e-8
    return opt


def test_optimizer_state_round_trip_for_a_resume():
    """AdamW.state_dict / load_state_dict (what SegmentationTask.save_checkpoint stores beside the weights): moments, the
    per-segment step counters left by modality dropout, hyper-parameters; mismatching arenas / broken segments are refused."""
    import pytest
    import torch
    a = _host_adamw(1000, 1)
    for k in range(5):
        a._segs.advance([(0, 400)] if k % 2 else [])
        a.step_count += 1
    a.lr = 2.5e-5
    state = a.state_dict()
    assert state["segments"] == [[0, 400, 3], [400, 1000, 5]] and state["step_count"] == 5 and state["exp_avg"].device.type == "cpu"
    b = _host_adamw(1000, 2)
    b.load_state_dict(state)
    assert torch.equal(b.exp_avg, a.exp_avg) and torch.equal(b.exp_avg_sq, a.exp_avg_sq) and b.exp_avg is not state["exp_avg"]
    assert b._segs.items == [[0, 400, 3], [400, 1000, 5]] and b.step_count == 5 and int(b.step_dev) == 5 and b.lr == 2.5e-5
    assert b._segs.advance() == [(0, 400, 4), (400, 1000, 6)]            # counting goes on where it stopped
    with pytest.raises(ValueError, match="arena"):
        _host_adamw(999, 3).load_state_dict(state)
    broken = dict(state, segments=[[0, 400, 3], [500, 1000, 5]])
    with pytest.raises(ValueError, match="segments"):
        _host_adamw(1000, 4).load_state_dict(broken)
