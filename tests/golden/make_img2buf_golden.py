#!/usr/bin/env python3
"""Generate tests/golden/img2buf.json with the reference's OWN compiled img2buf (lib/JM/output.c:87-178, linked into
oracle/_ref/jm_ldecod from the unmodified JM sources; `jm_ldecod --img2buf`).  Run in the build container (needs the
reference tree); the JSON is committed and is all the GPU box needs.

Input planes are counter based: sample i = lowbias32(seed + i) & mask (mask 0xff: 8-bit content, 0x3ff: 10-bit content
whose upper bits the reference DROPS, 0xffff: every bit pattern)."""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import lvoracle as orc  # noqa: E402

CASES = [  # size_x, size_y, crop_left, crop_right, crop_top, crop_bottom, seed, mask
    (48, 32, 0, 0, 0, 0, 1, 0xff), (48, 32, 4, 2, 2, 6, 2, 0x3ff), (16, 16, 0, 0, 0, 8, 3, 0xffff),
    (352, 288, 0, 0, 0, 0, 4, 0xff), (1920, 1088, 0, 0, 0, 8, 5, 0x3ff), (960, 544, 0, 0, 0, 4, 6, 0xffff),
    (64, 48, 16, 16, 2, 2, 7, 0xffff), (176, 144, 2, 6, 4, 0, 8, 0x1ff),
]


def plane(n, seed, mask):
    i = np.arange(n, dtype=np.uint64)
    x = (i + seed) & 0xffffffff
    x ^= x >> 16; x = (x * 0x7feb352d) & 0xffffffff
    x ^= x >> 15; x = (x * 0x846ca68b) & 0xffffffff
    x ^= x >> 16
    return (x & mask).astype(np.uint16)


def main():
    out = {"generator": "oracle/_ref/jm_ldecod --img2buf = /root/reference/lib/JM/output.c:87-178 compiled unmodified, "
                        "symbol_size_in_bytes = 1, imgpel = unsigned short", "cases": []}
    for (sx, sy, cl, cr, ct, cb, seed, mask) in CASES:
        p = plane(sx * sy, seed, mask)
        assert int(p[0]) == (orc.lowbias32(seed) & mask)
        got = orc.jm_img2buf(p, sx, sy, cl, cr, ct, cb)
        assert got is not None, "oracle/_ref/jm_ldecod must exist (python oracle/jm/build_jm.py)"
        out["cases"].append({"size_x": sx, "size_y": sy, "crop": [cl, cr, ct, cb], "seed": seed, "mask": mask,
                             "bytes": int(got.size), "first8": [int(v) for v in got[:8]],
                             "sha256": hashlib.sha256(got.tobytes()).hexdigest()})
    with open(os.path.join(HERE, "img2buf.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(f"wrote {len(out['cases'])} cases")
    # the difference primitive: the reference's own macro IsNearlyZero (lib/JM/image.h:22) as compiled into jm_ldecod,
    # for every a in [-255, 255] (rows) and tol in [0, 255] (columns)
    import subprocess
    exe = os.path.join(HERE, "..", "..", "oracle", "_ref", "jm_ldecod")
    tab = np.frombuffer(subprocess.run([exe, "--nearlyzero"], capture_output=True, check=True).stdout, dtype=np.uint8)
    assert tab.size == 511 * 256
    t = tab.reshape(511, 256)
    nz = {"generator": "oracle/_ref/jm_ldecod --nearlyzero = the macro IsNearlyZero of /root/reference/lib/JM/image.h:22 as compiled",
          "rows": "a = -255..255", "cols": "tol = 0..255", "sha256": hashlib.sha256(tab.tobytes()).hexdigest(),
          "ones": int(tab.sum()),
          "samples": [{"a": a, "tol": tol, "nearly_zero": int(t[a + 255, tol])}
                      for (a, tol) in [(0, 0), (1, 0), (-1, 0), (20, 20), (21, 20), (-20, 20), (-21, 20), (25, 25), (-26, 25),
                                       (255, 255), (-255, 255), (255, 254), (-255, 254), (128, 127), (-128, 128)]]}
    with open(os.path.join(HERE, "nearlyzero.json"), "w") as f:
        json.dump(nz, f, indent=1)
    print("wrote nearlyzero.json", nz["sha256"][:16], nz["ones"])


if __name__ == "__main__":
    main()
