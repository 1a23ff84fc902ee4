"""CPU tests of the host step's chunk planner (lv_host_step_plan: pure host logic inside liblivevideo.so, no device touched).
The same function cuts the batch inside lv_step_host / lv_step_host_pictures, so what is checked here is what runs."""
import itertools

import numpy as np
import pytest

MB = 1 << 20


def _cover(plan, ns, nf):
    seen = np.zeros((ns, nf), np.int32)
    for (s0, n_s, t0, n_t) in plan:
        assert n_s >= 1 and n_t >= 1 and s0 >= 0 and t0 >= 0 and s0 + n_s <= ns and t0 + n_t <= nf
        seen[s0:s0 + n_s, t0:t0 + n_t] += 1
    return seen


@pytest.fixture(autouse=True)
def _clean_env(monkeypatch):
    for k in ("LV_HOST_CHUNK_MB", "LV_HOST_SLOT_MB", "LV_HOST_RAMP"):
        monkeypatch.delenv(k, raising=False)


def test_every_frame_exactly_once_in_order(lv, monkeypatch):
    rng = np.random.default_rng(7)
    for _ in range(300):
        w, h = 16 * int(rng.integers(1, 250)), 2 * int(rng.integers(1, 1100))
        ns, nf = int(rng.integers(1, 300)), int(rng.integers(1, 200))
        monkeypatch.setenv("LV_HOST_CHUNK_MB", str(int(rng.choice([1, 2, 16, 64, 1024]))))
        monkeypatch.setenv("LV_HOST_SLOT_MB", str(int(rng.choice([1, 64, 256, 4096]))))
        monkeypatch.setenv("LV_HOST_RAMP", str(int(rng.integers(0, 2))))
        plan = lv.host_step_plan(w, h, ns, nf)
        assert (_cover(plan, ns, nf) == 1).all(), (w, h, ns, nf)
        # a stream's frames go through in order (chunk k's reference is the state chunk k - 1 left behind)
        last = {}
        for (s0, n_s, t0, n_t) in plan:
            for s in range(s0, s0 + n_s):
                assert last.get(s, 0) == t0, (w, h, ns, nf, s)
                last[s] = t0 + n_t
        # a slot holds every chunk: streams x frames of I420 input within LV_HOST_SLOT_MB, or a single frame per stream
        slot = int(__import__("os").environ["LV_HOST_SLOT_MB"]) * MB
        frame = w * h * 3 // 2
        for (s0, n_s, t0, n_t) in plan:
            assert n_s * n_t * frame <= slot or n_t == 1, (w, h, ns, nf)
            assert n_s * frame <= slot or n_s == 1


def test_default_plan_of_the_headline_workload(lv):
    """1080p x 64, one stream: the first chunks ramp up 1, 2, 4 frames, then 5-frame chunks (16 MiB of input)."""
    plan = lv.host_step_plan(1920, 1080, 1, 64)
    nts = [c[3] for c in plan]
    assert nts[:4] == [1, 2, 4, 5] and set(nts[3:-1]) == {5} and sum(nts) == 64
    assert all(c[0] == 0 and c[1] == 1 for c in plan)
    assert [c[2] for c in plan] == list(itertools.accumulate([0] + nts[:-1]))


def test_ramp_off_and_small_pictures_do_not_ramp(lv, monkeypatch):
    monkeypatch.setenv("LV_HOST_RAMP", "0")
    assert [c[3] for c in lv.host_step_plan(1920, 1080, 1, 64)][:3] == [5, 5, 5]
    monkeypatch.delenv("LV_HOST_RAMP")
    # 64 x 40 pictures: a megabyte of input is 273 frames, so a 4-frame step is one chunk (one launch)
    assert lv.host_step_plan(64, 40, 1, 4) == [(0, 1, 0, 4)]
    # CIF: the ramp starts at 7 frames (>= 1 MiB), not at one
    assert lv.host_step_plan(352, 288, 1, 300)[0] == (0, 1, 0, 7)


def test_many_streams_are_cut_by_streams(lv):
    """256 streams of 720p: one frame of all streams is 354 MB, more than a 256 MiB slot, so chunks take 194 streams x 1
    frame; the second stream group does not ramp (the pipeline is full)."""
    plan = lv.host_step_plan(1280, 720, 256, 16)
    frame = 1280 * 720 * 3 // 2
    cs = (256 * MB) // frame
    assert cs == 194 and plan[0] == (0, cs, 0, 1) and plan[16] == (cs, 256 - cs, 0, 1) and len(plan) == 32
    assert (_cover(plan, 256, 16) == 1).all()


def test_decoder_pictures_plan_by_their_own_size(lv):
    """lv_step_host_pictures: 1920 x 1088 16-bit pictures are 6.3 MB each on the host, so a 16 MiB chunk holds 2."""
    samples = 1920 * 1088 * 3 // 2
    plan = lv.host_step_plan(1920, 1080, 1, 9, in_frame_bytes=2 * samples)
    assert [c[3] for c in plan] == [1, 2, 2, 2, 2]


def test_plan_rejects_bad_arguments(lv):
    with pytest.raises(lv.LiveVideoError):
        lv.host_step_plan(1922, 1080, 1, 4)
    with pytest.raises(lv.LiveVideoError):
        lv.host_step_plan(1920, 1080, -1, 4)
    assert lv.host_step_plan(1920, 1080, 0, 4) == [] and lv.host_step_plan(1920, 1080, 3, 0) == []
