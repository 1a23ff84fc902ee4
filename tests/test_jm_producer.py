"""The frame producer of BASELINE.json configs[4]: the reference's bundled JM decoder (compiled unmodified into
oracle/_ref/jm_ldecod) must decode the synthetic I_PCM Annex-B stream back to exactly the generator's frames -- the
planar I420 layout of lib/JM/output.c:427-451 that the GPU path consumes.  CPU only; skipped where the binary is absent."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
JM = os.path.join(ROOT, "oracle", "_ref", "jm_ldecod")


def _decode(stream, w, h):
    r = subprocess.run([JM, stream, "-", str(w), str(h)], capture_output=True, timeout=120)
    assert r.returncode == 0, r.stderr[-300:]
    return np.frombuffer(r.stdout, dtype=np.uint8)


@pytest.fixture(scope="module")
def jm():
    if not os.path.exists(JM):
        if os.path.exists("/root/reference/lib/JM"):
            subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "jm", "build_jm.py")], check=False)
    if not os.path.exists(JM):
        pytest.skip("oracle/_ref/jm_ldecod not built (needs the reference tree)")
    return JM


@pytest.mark.parametrize("w,h,n", [(352, 288, 3), (1920, 1080, 2), (64, 48, 4), (176, 72, 2)])
def test_jm_decodes_ipcm_stream_to_generator_frames(jm, orc, tmp_path, w, h, n):
    import ipcm_stream
    frames = [orc.gen("camera" if min(w, h) >= 48 else "uniform", 1, 0, t, w, h) for t in range(n)]
    path = str(tmp_path / "s.264")
    ipcm_stream.write_stream(path, frames, w, h)
    out = _decode(path, w, h)
    fb = w * h * 3 // 2
    assert out.size == n * fb          # cropped to width x height (1080 is coded as 1088 rows)
    for t in range(n):
        assert np.array_equal(out[t * fb:(t + 1) * fb], frames[t]), t


def test_emulation_prevention(jm, tmp_path):
    """Payloads full of 00 00 0x patterns survive the NAL escaping (lib/JM/nalucommon.c EBSPtoRBSP undoes it)."""
    import ipcm_stream
    w, h = 64, 48
    fb = w * h * 3 // 2
    rng = np.random.default_rng(0)
    frames = [np.zeros(fb, np.uint8), rng.integers(0, 4, fb, dtype=np.uint8), np.full(fb, 3, np.uint8),
              np.tile(np.array([0, 0, 1, 0, 0, 0, 3, 0, 0, 2], np.uint8), fb // 10 + 1)[:fb],
              rng.integers(0, 256, fb, dtype=np.uint8)]
    path = str(tmp_path / "z.264")
    ipcm_stream.write_stream(path, frames, w, h)
    raw = open(path, "rb").read()
    body = raw[4:]
    assert b"\x00\x00\x00" not in body.replace(b"\x00\x00\x00\x01", b"")   # no start-code emulation inside NAL units
    out = _decode(path, w, h)
    for t, fr in enumerate(frames):
        assert np.array_equal(out[t * fb:(t + 1) * fb], fr), t


@pytest.mark.gpu
def test_config5_chain_small(jm, orc):
    """stream -> JM -> pipe -> pinned x2 -> lv_step_host -> outputs equal the oracle on the generator frames"""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import config5_e2e
    rep = config5_e2e.run(352, 288, 11, 4, check=True)
    assert rep["bit_exact_vs_oracle"] is True
    assert rep["gpu_launches"] >= 3


JM_INPROC = os.path.join(ROOT, "oracle", "_ref", "jm_inproc")


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,n,batch", [(352, 288, 7, 3), (1920, 1080, 5, 2), (64, 40, 4, 8)])
def test_inprocess_hook_feeds_real_16bit_pictures(jm, orc, tmp_path, w, h, n, batch):
    """The reference's decoder and liblivevideo.so in ONE process (oracle/_ref/jm_inproc = JM unmodified, its
    write_out_picture replaced at link time by tools/jm_hook/jm_inproc.c): every decoded picture's unsigned-short planes
    (imgpel, lib/JM/global.h:61; 1080p is coded as 1088 rows and cropped, output.c:375-380) are block-copied into a pinned
    batch and go through lv_step_host_pictures.  BGR, counts, map and summaries must equal the oracle on the generator's
    frames, bit for bit."""
    import json
    import ipcm_stream
    if not os.path.exists(JM_INPROC):
        pytest.skip("oracle/_ref/jm_inproc not built (needs the reference tree and liblivevideo.so)")
    frames = np.stack([orc.gen("camera" if min(w, h) >= 48 else "uniform", 3, 0, t, w, h) for t in range(n)])
    path = str(tmp_path / "s.264")
    ipcm_stream.write_stream(path, list(frames), w, h)
    prefix = str(tmp_path / "out")
    r = subprocess.run([JM_INPROC, path, str(w), str(h), str(batch), prefix], capture_output=True, timeout=300)
    assert r.returncode == 0, r.stderr[-500:]
    rep = json.loads(r.stdout.decode().strip().splitlines()[-1])
    assert rep["frames"] == n and rep["imgpel_bytes"] == 2
    assert rep["picture"] == [(w + 15) // 16 * 16, (h + 15) // 16 * 16]        # the decoder's own (uncropped) planes
    # one kernel per batch narrows, crops, converts and differences (macroblock-aligned decoder pictures always take the
    # one-kernel form of lv_step_host_pictures); layouts it cannot take would show two
    assert (n + batch - 1) // batch <= rep["gpu_launches"] <= 2 * ((n + batch - 1) // batch)
    prev = np.zeros((1, w * h), np.uint8)
    ref = orc.convert_diff_batch(frames[None], prev, 1, n, w, h, 20, 0, want_mask=False)
    assert np.array_equal(np.fromfile(prefix + ".bgr", np.uint8), ref["bgr"])
    assert np.array_equal(np.fromfile(prefix + ".cnt", np.uint16), ref["mb_count"])
    assert np.array_equal(np.fromfile(prefix + ".map", np.uint8), ref["mb_map"])
    assert np.array_equal(np.fromfile(prefix + ".sum", np.uint32).reshape(-1, 2), ref["summary"])
