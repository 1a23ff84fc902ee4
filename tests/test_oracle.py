"""The oracle against the reference's golden material (CPU only).

Pins the C restatement (oracle/lv_oracle.c) to (i) the known-answer vectors and digests SURVEY.md 8(c)
recorded from the reference's original object code, (ii) tests/golden/golden.json (same object code, generated
by tests/golden/make_golden.py) and (iii) -- when this host can execute i386 binaries -- the object code itself
on all 2^24 (Y,U,V) triples."""
import hashlib

import numpy as np
import pytest

# (Y,U,V) -> (B,G,R) from the real Convert.obj
SURVEY_KATS = {
    (16, 128, 128): (0, 0, 0), (235, 128, 128): (254, 254, 254), (128, 128, 128): (130, 130, 130),
    (0, 0, 0): (0, 135, 0), (255, 255, 255): (255, 125, 255), (255, 0, 0): (19, 255, 73),
    (0, 255, 255): (237, 0, 183), (41, 240, 110): (254, 0, 0), (100, 100, 100): (41, 131, 52),
    (200, 30, 220): (16, 178, 255), (81, 90, 240): (0, 0, 254), (145, 54, 34): (0, 255, 0),
    (17, 129, 127): (3, 1, 0), (1, 2, 3): (0, 133, 0),
}
# SURVEY.md 8(c) golden digests: uniform generator, seed=1, stream=0
SURVEY_DIGESTS = [
    (352, 288, 0, [0, 74, 125, 0, 97, 148], "529aeddc", "67db954f948a9d8a19b084b6b4ec499ea4e0d2f96ce8ea2e58937840bbbc9f61"),
    (352, 288, 1, [0, 228, 230, 0, 177, 179], "8665fce4", "42e1b331a89b43f7dba54e88a0770ac6507d142e5e92f96ff1f91bfe15b3cc5e"),
    (352, 288, 2, [123, 56, 248, 139, 72, 255], "3dab18f7", "3fb8ef2bf2b89f5c64db03e7f0be68c8ba98a5135aa43408ea52dc09b832f206"),
    (352, 288, 3, [22, 219, 255, 15, 212, 255], "8c83309f", "080bb38dc42b5de73e251bef092b7cb963fd7c18048323da12b382eef89e24cb"),
    (1280, 720, 0, [137, 0, 79, 119, 0, 61], "86eb07bb", "6e657d7920242a2063cb608e04a32e7c800fb13725714ab107edd9ead0c5626c"),
    (1920, 1080, 0, [0, 126, 49, 48, 254, 177], "4f1a3419", "e29984f0bbb661d33b19c509dd6c224a71cb9f8aed64bd2a2eb110af5ab3b41a"),
    (1920, 1080, 1, [226, 0, 0, 255, 103, 62], "3eea6da5", "324724d1d053cf9f8edc7f924aa88bd8cd9731c396da310a753bbb5d6224a9fa"),
    (3840, 2160, 0, [255, 118, 255, 255, 139, 255], "a30359a9", "b0a06c25de4cf57b3e44664d2c712cacdfe363812a2626d701ab2203800f0a60"),
]


def test_lowbias32_and_generator_prefix(orc):
    assert orc.lowbias32(1) == 0x688990C0
    assert orc.lowbias32(0xDEADBEEF) == 0xE628C683
    for (w, h) in [(352, 288), (1280, 720), (1920, 1080)]:
        fr = orc.gen("uniform", 1, 0, 0, w, h)
        px = w * h
        assert fr[:4].tolist() == [214, 96, 123, 232]
        assert fr[px:px + 2].tolist() == [175, 116]
        assert fr[px + px // 4:px + px // 4 + 2].tolist() == [42, 104]


@pytest.mark.parametrize("yuv,bgr", sorted(SURVEY_KATS.items()))
def test_survey_kats(orc, yuv, bgr):
    assert orc.px(*yuv) == bgr


def test_golden_kats(orc, golden):
    for k in golden["kat"]:
        assert list(orc.px(*k["yuv"])) == k["bgr"], k


@pytest.mark.parametrize("w,h,frame,first6,crc,sha", SURVEY_DIGESTS)
def test_survey_digests(orc, w, h, frame, first6, crc, sha):
    out = orc.convert_frame(orc.gen("uniform", 1, 0, frame, w, h), w, h)
    assert out[:6].tolist() == first6
    assert "%08x" % orc.crc32(out) == crc
    assert hashlib.sha256(out.tobytes()).hexdigest() == sha


def test_golden_convert_and_inputs(orc, golden):
    for e, i in zip(golden["convert"], golden["inputs"]):
        fr = orc.gen(e["kind"], e["seed"], e["stream"], e["frame"], e["w"], e["h"])
        assert hashlib.sha256(fr.tobytes()).hexdigest() == i["sha256"], ("input", e)
        out = orc.convert_frame(fr, e["w"], e["h"])
        assert out[:6].tolist() == e["first6"]
        assert hashlib.sha256(out.tobytes()).hexdigest() == e["sha256"], e


def test_layout_bottom_up_and_chroma_replication(orc, golden):
    # 4x4, source rows Y=16,80,160,235: first byte of destination rows 0..3 (254,167,74,0)
    y = np.repeat(np.array([16, 80, 160, 235], dtype=np.uint8), 4)
    out = orc.yv12_to_rgb24(y, np.full(4, 128, np.uint8), np.full(4, 128, np.uint8), 4, 4)
    assert [int(out[12 * r]) for r in range(4)] == [254, 167, 74, 0] == golden["layout_kat"]["first_byte_of_dst_rows"]
    # 4x2, U row = [0,255]: nearest-neighbour 2x2 replication
    out = orc.yv12_to_rgb24(np.full(8, 128, np.uint8), np.array([0, 255], np.uint8), np.full(2, 128, np.uint8), 4, 2)
    assert [int(out[3 * p]) for p in range(4)] == [0, 0, 255, 255] == golden["chroma_kat"]["b_row0"]
    assert [int(out[12 + 3 * p]) for p in range(4)] == [0, 0, 255, 255] == golden["chroma_kat"]["b_row1"]


def test_not_the_textbook_form(orc):
    # the reference's two-stage truncation differs from the textbook >>16 form by +-1 on 12-18 % of triples
    rng = np.random.default_rng(0)
    t = rng.integers(0, 256, size=(20000, 3))
    diff = np.zeros(3)
    for (y, u, v) in t:
        a, b = orc.px(y, u, v), orc.textbook_px(y, u, v)
        for c in range(3):
            assert abs(a[c] - b[c]) <= 1
            diff[c] += a[c] != b[c]
    rate = diff / len(t)
    assert 0.10 < rate[0] < 0.15 and 0.15 < rate[1] < 0.21 and 0.12 < rate[2] < 0.17, rate
    assert orc.px(255, 0, 0) == (19, 255, 73) and orc.textbook_px(255, 0, 0)[0] == 20


def test_exhaustive_against_reference_object_code(orc):
    """All 2^24 triples: 256 frames of 512x512 (Y constant, U = column, V = row) through the ORIGINAL code."""
    if not orc.have_oracle32():
        pytest.skip("this host cannot execute the i386 reference object code (oracle/_ref/oracle32)")
    u = np.repeat(np.arange(256, dtype=np.uint8)[None, :], 256, 0).ravel()
    v = np.repeat(np.arange(256, dtype=np.uint8)[:, None], 256, 1).ravel()
    bad = 0
    for y0 in range(0, 256, 64):
        frames = np.stack([np.concatenate([np.full(512 * 512, y, np.uint8), u, v]) for y in range(y0, y0 + 64)])
        ref = orc.oracle32_convert(frames, 512, 512).reshape(64, -1)
        for i in range(64):
            bad += int((orc.convert_frame(frames[i], 512, 512) != ref[i]).sum())
    assert bad == 0


def test_object_code_on_golden_sizes(orc, golden):
    if not orc.have_oracle32():
        pytest.skip("oracle32 not runnable here")
    for e in golden["convert"][:6]:
        fr = orc.gen(e["kind"], e["seed"], e["stream"], e["frame"], e["w"], e["h"])
        out = orc.oracle32_convert(fr, e["w"], e["h"])
        assert hashlib.sha256(out.tobytes()).hexdigest() == e["sha256"]


# ---- difference / macroblock part ---------------------------------------------------------------------

def test_is_nearly_zero_matches_macro(orc):
    L = orc.lib()
    for tol in (0, 1, 20, 25, 254, 255):
        for a in range(-255, 256):
            macro = ((a >= 0) and (a <= tol)) or ((a < 0) and (a >= -tol))   # lib/JM/image.h:22
            assert bool(L.lvo_is_nearly_zero(a, tol)) == macro == (abs(a) <= tol)


def test_frame_diff_boundaries(orc):
    w, h = 32, 16
    ref = np.full(w * h, 100, np.uint8)
    cur = ref.copy()
    cur[0] = 120          # |d| == tol      -> unchanged
    cur[1] = 121          # |d| == tol + 1  -> changed
    cur[16] = 79          # -21             -> changed (second macroblock)
    cur[17] = 80          # -20             -> unchanged
    mask, cnt, mp, total = orc.frame_diff_mb(cur, ref, w, h, 20, 0)
    assert mask[:2].tolist() == [0, 1] and mask[16:18].tolist() == [1, 0]
    assert cnt.tolist() == [1, 1] and mp.tolist() == [1, 1] and total == 2
    _, cnt, mp, _ = orc.frame_diff_mb(cur, ref, w, h, 20, 1)
    assert mp.tolist() == [0, 0]
    # numpy restatement on random data incl. a partial last macroblock row (1080-like: h % 16 == 8)
    rng = np.random.default_rng(1)
    w, h = 48, 24
    cur = rng.integers(0, 256, w * h, dtype=np.uint8)
    ref = rng.integers(0, 256, w * h, dtype=np.uint8)
    for tol, thr in [(0, 0), (20, 0), (25, 128), (254, 1), (255, 0), (0, 255), (0, 256)]:
        mask, cnt, mp, total = orc.frame_diff_mb(cur, ref, w, h, tol, thr)
        m = (np.abs(cur.astype(int) - ref.astype(int)) > tol).astype(np.uint8).reshape(h, w)
        assert np.array_equal(mask.reshape(h, w), m)
        exp = np.array([[m[16 * by:16 * by + 16, 16 * bx:16 * bx + 16].sum() for bx in range(3)] for by in range(2)])
        assert np.array_equal(cnt.reshape(2, 3), exp)
        assert np.array_equal(mp.reshape(2, 3), (exp > thr).astype(np.uint8))
        assert total == m.sum()


def test_batch_semantics_and_golden_diff(orc, golden):
    for e in golden["diff"]:
        if e["w"] * e["h"] * e["n_streams"] * e["n_frames"] > 8_000_000:
            continue   # the big ones are covered on the GPU box; keep the CPU suite short
        w, h, ns, nf = e["w"], e["h"], e["n_streams"], e["n_frames"]
        frames = orc.gen_batch(e["kind"], e["seed"], ns, nf, w, h)
        prev = np.zeros((ns, w * h), dtype=np.uint8)
        r = orc.convert_diff_batch(frames, prev, ns, nf, w, h, e["tol"], e["mb_thresh"], n_threads=2)
        assert hashlib.sha256(r["mask"].tobytes()).hexdigest() == e["mask"]["sha256"]
        assert hashlib.sha256(r["mb_count"].tobytes()).hexdigest() == e["mb_count"]["sha256"]
        assert hashlib.sha256(r["mb_map"].tobytes()).hexdigest() == e["mb_map"]["sha256"]
        assert hashlib.sha256(r["bgr"].tobytes()).hexdigest() == e["bgr"]["sha256"]
        assert r["summary"].tolist() == e["summary"]
        # frame t is differenced against frame t-1, frame 0 against the (zero) carried state; state <- last luma
        px = w * h
        for s in range(ns):
            assert np.array_equal(prev[s], frames[s, nf - 1, :px])
            for t in range(nf):
                ref = np.zeros(px, np.uint8) if t == 0 else frames[s, t - 1, :px]
                mask, cnt, mp, _ = orc.frame_diff_mb(frames[s, t, :px], ref, w, h, e["tol"], e["mb_thresh"])
                f = s * nf + t
                assert np.array_equal(r["mask"][f * px:(f + 1) * px], mask)


def test_camera_generator_exercises_all_block_classes(orc):
    w, h = 352, 288
    f0, f1 = orc.gen("camera", 1, 0, 0, w, h), orc.gen("camera", 1, 0, 1, w, h)
    mask, cnt, mp, total = orc.frame_diff_mb(f1[:w * h], f0[:w * h], w, h, 20, 0)
    assert (cnt == 0).any() and ((cnt > 0) & (cnt < 256)).any()
    # against the all-zero initial state (calloc'ed prev_frm) most macroblocks are fully changed
    _, cnt0, _, _ = orc.frame_diff_mb(f0[:w * h], np.zeros(w * h, np.uint8), w, h, 20, 0)
    assert (cnt0 == 256).any()
    d = np.abs(f1[:w * h].astype(int) - f0[:w * h].astype(int))
    assert (d == 20).any() and (d == 21).any()       # threshold boundary is hit on both sides


def test_write_bks_and_object_box(orc):
    rng = np.random.default_rng(3)
    w, h = 32, 16
    cur = rng.integers(0, 256, w * h * 3 // 2, dtype=np.uint8)
    bkd = rng.integers(0, 256, w * h * 3 // 2, dtype=np.uint8)
    out = orc.write_bks(cur, bkd, w, h, 25)
    near = np.abs(cur.astype(int) - bkd.astype(int)) <= 25
    exp = cur.copy()
    exp[:w * h][near[:w * h]] = 16
    exp[w * h:][near[w * h:]] = 128
    assert np.array_equal(out, exp)
    mask = np.zeros(w * h, np.uint8)
    assert orc.object_box(mask, w, 2, 20, 1, 10) == (21, 1, 11, 0)     # empty: the reference's initial values
    mask[5 * w + 7] = 1
    mask[9 * w + 3] = 1
    assert orc.object_box(mask, w, 2, 20, 1, 10) == (3, 7, 5, 9)


# ---- the other five cscc.lib converters (Convert.h:25-30) -----------------------------------------------------------

@pytest.mark.parametrize("name", ["rgb24_to_yv12", "yvu9_to_yv12", "yuy2_to_yv12", "yv12_to_yvu9", "yv12_to_yuy2"])
def test_other_converters_against_object_code(orc, name):
    if not orc.have_oracle32():
        pytest.skip("oracle32 not runnable here")
    mode, inb, outb = orc.FORMATS[name]
    rng = np.random.default_rng(mode)
    for (w, h) in [(16, 8), (64, 48), (352, 288), (320, 240), (1920, 1080)]:
        data = rng.integers(0, 256, w * h // 8 * inb, dtype=np.uint8)
        assert np.array_equal(orc.convert_format(name, data, w, h), orc.oracle32_convert(data, w, h, mode=mode)), (name, w, h)


def test_rgb24_to_yv12_exhaustive_against_object_code(orc):
    """All 2^24 (R,G,B) triples, each as the top-left pixel of a 2x2 block (the one that also produces chroma)."""
    if not orc.have_oracle32():
        pytest.skip("oracle32 not runnable here")
    vals = np.arange(1 << 24, dtype=np.uint32)
    bad = 0
    for c in range(16):
        v = vals[c << 20:(c + 1) << 20]
        blk = np.stack([(v >> 16) & 255, (v >> 8) & 255, v & 255], axis=-1).astype(np.uint8).reshape(1024, 1024, 3)
        img = np.zeros((2048, 2048, 3), np.uint8)
        img[0::2, 0::2] = blk
        img[1::2, 0::2] = blk[:, ::-1]
        img[0::2, 1::2] = blk[::-1]
        img[1::2, 1::2] = blk
        data = img.reshape(-1)
        bad += int((orc.convert_format("rgb24_to_yv12", data, 2048, 2048) != orc.oracle32_convert(data, 2048, 2048, mode=1)).sum())
    assert bad == 0


def test_other_converters_known_values(orc):
    # flat colours through RGB24_to_YV12 (values recorded from the object code): bytes are R,G,B
    for rgb, yuv in {(0, 0, 0): (0, 128, 128), (255, 0, 0): (76, 85, 255), (0, 255, 0): (149, 44, 22),
                     (0, 0, 255): (29, 255, 108), (255, 255, 255): (254, 128, 129)}.items():
        out = orc.convert_format("rgb24_to_yv12", np.tile(np.array(rgb, np.uint8), 16 * 8), 16, 8)
        assert (int(out[0]), int(out[128]), int(out[160])) == yuv
    # YUY2 -> YV12 -> YUY2 keeps luma and the odd rows' chroma
    rng = np.random.default_rng(2)
    w, h = 16, 8
    yuy2 = rng.integers(0, 256, 2 * w * h, dtype=np.uint8)
    yv12 = orc.convert_format("yuy2_to_yv12", yuy2, w, h)
    back = orc.convert_format("yv12_to_yuy2", yv12, w, h).reshape(h, w // 2, 4)
    src = yuy2.reshape(h, w // 2, 4)
    assert np.array_equal(back[:, :, 0::2], src[:, :, 0::2])
    assert np.array_equal(back[0::2, :, 1::2], src[1::2, :, 1::2]) and np.array_equal(back[1::2, :, 1::2], src[1::2, :, 1::2])


# ---- the decoder picture -> I420 step (img2buf / write_out_picture, lib/JM/output.c) ------------------------------
def _plane16(n, seed, mask):
    i = np.arange(n, dtype=np.uint64)
    x = (i + seed) & 0xffffffff
    x ^= x >> 16
    x = (x * 0x7feb352d) & 0xffffffff
    x ^= x >> 15
    x = (x * 0x846ca68b) & 0xffffffff
    x ^= x >> 16
    return (x & mask).astype(np.uint16)


def test_img2buf_matches_golden_from_the_reference_code(orc):
    """tests/golden/img2buf.json was produced by the reference's own compiled img2buf (make_img2buf_golden.py)."""
    import json
    import os
    with open(os.path.join(os.path.dirname(__file__), "golden", "img2buf.json")) as f:
        g = json.load(f)
    assert len(g["cases"]) >= 8
    for c in g["cases"]:
        sx, sy, (cl, cr, ct, cb) = c["size_x"], c["size_y"], c["crop"]
        p = _plane16(sx * sy, c["seed"], c["mask"])
        out = orc.img2buf(p, sx, sy, cl, cr, ct, cb)
        assert out.size == c["bytes"] and [int(v) for v in out[:8]] == c["first8"]
        assert hashlib.sha256(out.tobytes()).hexdigest() == c["sha256"], c
        # the semantics in one line: low byte of the sample, crop rectangle cut away
        want = (p.reshape(sy, sx)[ct:sy - cb, cl:sx - cr] & 0xff).astype(np.uint8).ravel()
        assert np.array_equal(out, want)


def test_img2buf_against_the_live_reference_code(orc):
    """Where oracle/_ref/jm_ldecod exists (this container), compare with the reference's function directly."""
    rng = np.random.default_rng(3)
    if orc.jm_img2buf(np.zeros(4, np.uint16), 2, 2, 0, 0, 0, 0) is None:
        pytest.skip("oracle/_ref/jm_ldecod not built here")
    for (sx, sy, cl, cr, ct, cb) in [(32, 16, 0, 0, 0, 0), (40, 24, 8, 4, 2, 6), (1920, 1088, 0, 0, 0, 8), (18, 10, 2, 2, 2, 2)]:
        p = rng.integers(0, 65536, sx * sy, dtype=np.uint16)
        assert np.array_equal(orc.img2buf(p, sx, sy, cl, cr, ct, cb), orc.jm_img2buf(p, sx, sy, cl, cr, ct, cb))


def test_picture16_to_i420_plane_order_and_half_crop(orc):
    """write_out_picture (output.c:362-453): Y with the luma crop, then imgUV[0], imgUV[1] with half of it."""
    sx, sy, cl, cr, ct, cb = 48, 32, 4, 12, 2, 6
    rng = np.random.default_rng(4)
    pic = rng.integers(0, 1024, sx * sy * 3 // 2, dtype=np.uint16)
    out = orc.picture16_to_i420(pic, sx, sy, cl, cr, ct, cb)
    w, h = sx - cl - cr, sy - ct - cb
    y = orc.img2buf(pic[:sx * sy], sx, sy, cl, cr, ct, cb)
    u = orc.img2buf(pic[sx * sy:sx * sy * 5 // 4], sx // 2, sy // 2, cl // 2, cr // 2, ct // 2, cb // 2)
    v = orc.img2buf(pic[sx * sy * 5 // 4:], sx // 2, sy // 2, cl // 2, cr // 2, ct // 2, cb // 2)
    assert out.size == w * h * 3 // 2 and np.array_equal(out, np.concatenate([y, u, v]))


def test_is_nearly_zero_matches_the_reference_macro_as_compiled(orc):
    """tests/golden/nearlyzero.json holds the reference's own macro IsNearlyZero (lib/JM/image.h:22), compiled into
    oracle/_ref/jm_ldecod, evaluated on every difference a in [-255, 255] and every tol in [0, 255]: the oracle's
    restatement -- and with it the rule 'changed <=> |cur - ref| > tol' -- agrees on all 130 816 pairs."""
    import json
    import os
    with open(os.path.join(os.path.dirname(__file__), "golden", "nearlyzero.json")) as f:
        g = json.load(f)
    tab = np.array([[orc.lib().lvo_is_nearly_zero(a, tol) for tol in range(256)] for a in range(-255, 256)], dtype=np.uint8)
    assert int(tab.sum()) == g["ones"]
    assert hashlib.sha256(tab.tobytes()).hexdigest() == g["sha256"]
    for s in g["samples"]:
        assert int(tab[s["a"] + 255, s["tol"]]) == s["nearly_zero"], s
    a = np.arange(-255, 256)[:, None]
    assert np.array_equal(tab, (np.abs(a) <= np.arange(256)[None, :]).astype(np.uint8))     # <=> |a| <= tol
