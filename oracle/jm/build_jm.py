#!/usr/bin/env python3
"""Build oracle/_ref/jm_ldecod: the reference's bundled JM 10.2 decoder (lib/JM/*.c, compiled UNMODIFIED from where
the sources lie under /root/reference) linked with oracle/jm/jm_main.c.  TEST / BENCH INFRASTRUCTURE ONLY: it is the
host-side, untimed producer of planar I420 frames for BASELINE.json configs[4].  Nothing is copied into the repo;
the only output is the binary under oracle/_ref/ (git-ignored, travels to the GPU box)."""
import glob
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.normpath(os.path.join(HERE, "..", "_ref", "jm_ldecod"))
JM_DIR = os.environ.get("LV_REFERENCE_JM", "/root/reference/lib/JM")
# -std=gnu89 -fcommon: the 2005-era sources rely on implicit int / tentative definitions in headers (global.h:146-238
# defines ~90 globals in a header); windows.h is replaced by oracle/jm/shim/windows.h (BYTE, LARGE_INTEGER, max/min)
CFLAGS = ["-std=gnu89", "-w", "-fcommon", "-O2", "-I" + os.path.join(HERE, "shim"), "-I" + JM_DIR]


def build(verbose=True):
    srcs = sorted(glob.glob(os.path.join(JM_DIR, "*.c")))
    if not srcs:
        raise FileNotFoundError(f"{JM_DIR} not present (the GPU box only uses the prebuilt binary)")
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        objs = []
        procs = []
        for s in srcs + [os.path.join(HERE, "jm_main.c")]:
            o = os.path.join(tmp, os.path.basename(s)[:-2] + ".o")
            objs.append(o)
            procs.append(subprocess.Popen(["gcc", *CFLAGS, "-c", s, "-o", o], stderr=subprocess.PIPE))
        for p in procs:
            _, err = p.communicate()
            if p.returncode != 0:
                raise RuntimeError(err.decode()[-2000:])
        subprocess.run(["gcc", "-o", OUT, *objs, "-lm"], check=True)
    if verbose:
        print(f"built {OUT} from {len(srcs)} JM sources")
    return OUT


if __name__ == "__main__":
    try:
        build()
    except Exception as e:  # noqa: BLE001
        print(f"jm_ldecod not built: {e}", file=sys.stderr)
        sys.exit(1)
