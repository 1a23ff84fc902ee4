#!/usr/bin/env python3
"""Generate tests/golden/golden.json.  Run in the build container (needs oracle/_ref/oracle32, i.e. the
reference tree, to execute the reference's ORIGINAL object code); the JSON is committed and is all the
GPU box needs.

  conversion entries  : produced by the reference's own machine code (cscc.lib -> oracle32)
  kat entries         : single (Y,U,V) triples through the same machine code
  diff entries        : produced by the C restatement of lib/JM/image.h:22 + image.c:4795-4803/5143-5150
                        (the reference has no binary or test for these: "unpinned by upstream")
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import lvoracle as orc  # noqa: E402


def digest(a):
    a = np.ascontiguousarray(a)
    return {"crc32": "%08x" % orc.crc32(a.view(np.uint8)), "sha256": hashlib.sha256(a.tobytes()).hexdigest(),
            "bytes": int(a.nbytes)}


def main():
    assert orc.have_oracle32(), "oracle32 must run here (reference object code)"
    g = {"generator": "oracle/_ref/oracle32 = /root/reference/cscc.lib:Convert.obj (original i386 object code)",
         "convert": [], "kat": [], "diff": [], "inputs": []}
    kats = [(16, 128, 128), (235, 128, 128), (128, 128, 128), (0, 0, 0), (255, 255, 255), (255, 0, 0), (0, 255, 255),
            (41, 240, 110), (100, 100, 100), (200, 30, 220), (81, 90, 240), (145, 54, 34), (17, 129, 127), (1, 2, 3),
            (0, 0, 255), (0, 255, 0), (255, 0, 255), (255, 255, 0), (16, 0, 0), (16, 255, 255), (235, 16, 240)]
    for (y, u, v) in kats:
        out = orc.oracle32_convert(np.array([y] * 8 + [u] * 2 + [v] * 2, dtype=np.uint8), 4, 2)
        g["kat"].append({"yuv": [y, u, v], "bgr": [int(x) for x in out[:3]]})
    cases = [(352, 288, "uniform", 0), (352, 288, "uniform", 1), (352, 288, "uniform", 2), (352, 288, "uniform", 3),
             (1280, 720, "uniform", 0), (1920, 1080, "uniform", 0), (1920, 1080, "uniform", 1),
             (3840, 2160, "uniform", 0), (352, 288, "camera", 0), (352, 288, "camera", 5), (1920, 1080, "camera", 2),
             (320, 240, "uniform", 0), (16, 2, "uniform", 0), (4, 2, "uniform", 7), (20, 6, "uniform", 1),
             (1280, 720, "camera", 9)]
    for (w, h, kind, frame) in cases:
        fr = orc.gen(kind, 1, 0, frame, w, h)
        out = orc.oracle32_convert(fr, w, h)
        assert np.array_equal(out, orc.convert_frame(fr, w, h))
        e = {"w": w, "h": h, "kind": kind, "seed": 1, "stream": 0, "frame": frame, "first6": [int(x) for x in out[:6]]}
        e.update(digest(out))
        g["convert"].append(e)
        ie = {"w": w, "h": h, "kind": kind, "seed": 1, "stream": 0, "frame": frame}
        ie.update(digest(fr))
        g["inputs"].append(ie)
    # layout / chroma-siting KATs through the object code
    y = np.repeat(np.array([16, 80, 160, 235], dtype=np.uint8), 4)
    out = orc.oracle32_convert(np.concatenate([y, np.full(8, 128, np.uint8)]), 4, 4)
    g["layout_kat"] = {"first_byte_of_dst_rows": [int(out[12 * r]) for r in range(4)]}
    fr = np.concatenate([np.full(8, 128, np.uint8), np.array([0, 255], np.uint8), np.full(2, 128, np.uint8)])
    out = orc.oracle32_convert(fr, 4, 2)
    g["chroma_kat"] = {"b_row0": [int(out[3 * p]) for p in range(4)], "b_row1": [int(out[12 + 3 * p]) for p in range(4)]}
    # difference / macroblock entries (C restatement)
    for (w, h, ns, nf, tol, thr, kind) in [(352, 288, 1, 6, 20, 0, "camera"), (352, 288, 1, 3, 25, 128, "uniform"),
                                           (1920, 1080, 1, 3, 20, 0, "camera"), (1280, 720, 2, 3, 20, 1, "camera"),
                                           (3840, 2160, 1, 2, 20, 0, "camera"), (64, 48, 2, 4, 0, 255, "camera"),
                                           (352, 288, 1, 3, 255, 0, "uniform"), (352, 288, 1, 3, 0, 256, "uniform")]:
        frames = orc.gen_batch(kind, 1, ns, nf, w, h)
        prev = np.zeros((ns, w * h), dtype=np.uint8)
        r = orc.convert_diff_batch(frames, prev, ns, nf, w, h, tol, thr)
        e = {"w": w, "h": h, "n_streams": ns, "n_frames": nf, "tol": tol, "mb_thresh": thr, "kind": kind, "seed": 1,
             "mask": digest(r["mask"]), "mb_count": digest(r["mb_count"]), "mb_map": digest(r["mb_map"]),
             "bgr": digest(r["bgr"]), "summary": r["summary"].tolist(), "state": digest(prev)}
        g["diff"].append(e)
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(g, f, indent=1)
    print("wrote golden.json:", len(g["convert"]), "convert,", len(g["kat"]), "kat,", len(g["diff"]), "diff entries")


if __name__ == "__main__":
    main()
