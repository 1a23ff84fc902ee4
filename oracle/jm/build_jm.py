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
ROOT = os.path.normpath(os.path.join(HERE, "..", ".."))
OUT = os.path.normpath(os.path.join(HERE, "..", "_ref", "jm_ldecod"))
# the same decoder with the GPU library in ONE process (tools/jm_hook/jm_inproc.c): the reference's output.o is linked with
# its write_out_picture weakened, so the hook's definition receives every picture -- no reference source is edited
OUT_INPROC = os.path.normpath(os.path.join(HERE, "..", "_ref", "jm_inproc"))
JM_DIR = os.environ.get("LV_REFERENCE_JM", "/root/reference/lib/JM")
# -std=gnu89 -fcommon: the 2005-era sources rely on implicit int / tentative definitions in headers (global.h:146-238
# defines ~90 globals in a header); windows.h is replaced by oracle/jm/shim/windows.h (BYTE, LARGE_INTEGER, max/min)
# -fPIC: calls to global functions stay symbolic inside a translation unit (write_picture -> write_out_picture in output.c)
CFLAGS = ["-std=gnu89", "-w", "-fcommon", "-O2", "-fPIC", "-I" + os.path.join(HERE, "shim"), "-I" + JM_DIR]


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
        # in-process harness: JM objects (output.o with a weak write_out_picture) + the hook + liblivevideo.so
        lib_dir = os.path.join(ROOT, "livevideo_b200")
        if os.path.exists(os.path.join(lib_dir, "liblivevideo.so")):
            jm_objs = [o for o in objs if not o.endswith("jm_main.o")]
            out_o = os.path.join(tmp, "output.o")
            subprocess.run(["objcopy", "--weaken-symbol=write_out_picture", out_o], check=True)
            hook_o = os.path.join(tmp, "jm_inproc.o")
            subprocess.run(["gcc", *CFLAGS, "-std=gnu99", "-I" + HERE, "-I" + os.path.join(ROOT, "include"), "-c",
                            os.path.join(ROOT, "tools", "jm_hook", "jm_inproc.c"), "-o", hook_o], check=True)
            subprocess.run(["gcc", "-o", OUT_INPROC, hook_o, *jm_objs, "-L" + lib_dir, "-llivevideo",
                            "-Wl,-rpath,$ORIGIN/../../livevideo_b200", "-lm"], check=True)
            if verbose:
                print(f"built {OUT_INPROC} (decoder + GPU library in one process)")
    return OUT


if __name__ == "__main__":
    try:
        build()
    except Exception as e:  # noqa: BLE001
        print(f"jm_ldecod not built: {e}", file=sys.stderr)
        sys.exit(1)
