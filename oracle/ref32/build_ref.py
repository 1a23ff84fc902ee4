#!/usr/bin/env python3
"""Build oracle/_ref/oracle32 from the reference's ORIGINAL object code.  TEST INFRASTRUCTURE ONLY.

The colour conversion of wonsang/LiveVideo has no source in the reference tree: Convert.h:19-31
only declares class ColorSpaceConversions; the code is the i386 MSVC COFF archive cscc.lib
(member ./Temp/Convert.obj).  This script links that object, from where it lies under
/root/reference, into a freestanding static i386 ELF (oracle/ref32/harness32.c supplies
_start, __ftol and raw syscalls).  Nothing from the reference is copied into the repo; the only
output is oracle/_ref/oracle32 (git-ignored, travels to the GPU box with gpurun).

Steps:
  ar x cscc.lib                      -> Temp/Convert.obj (pe-i386)
  objcopy -I pe-i386 -O elf32-i386   -> conv.o
  patch every R_386_PC32 field in .text from 0 to -4 (PE stores 0 for `call rel32`, ELF expects the
        -4 addend in place); the sites are read from `objdump -r`, not hard-coded
  objcopy --redefine-sym             -> C-linkable names for the MSVC-mangled methods
  gcc -m32 -ffreestanding -c harness32.c ; ld -m elf_i386 -static
"""
import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.normpath(os.path.join(HERE, "..", "_ref"))
REF_LIB = os.environ.get("LV_REFERENCE_LIB", "/root/reference/cscc.lib")

RENAMES = {
    "??0ColorSpaceConversions@@QAE@XZ": "csc_ctor",
    "?YV12_to_RGB24@ColorSpaceConversions@@QAEXPAE000HH@Z": "csc_yv12_rgb24",
    "?RGB24_to_YV12@ColorSpaceConversions@@QAEXPAE0HH@Z": "csc_rgb24_yv12",
    "?YVU9_to_YV12@ColorSpaceConversions@@QAEXPAE0HH@Z": "csc_yvu9_yv12",
    "?YUY2_to_YV12@ColorSpaceConversions@@QAEXPAE0HH@Z": "csc_yuy2_yv12",
    "?YV12_to_YVU9@ColorSpaceConversions@@QAEXPAE0HH@Z": "csc_yv12_yvu9",
    "?YV12_to_YUY2@ColorSpaceConversions@@QAEXPAE0HH@Z": "csc_yv12_yuy2",
}


def run(*cmd, cwd=None):
    return subprocess.run(cmd, cwd=cwd, check=True, capture_output=True, text=True).stdout


def build(verbose=True):
    if not os.path.exists(REF_LIB):
        raise FileNotFoundError(f"{REF_LIB} not present (the GPU box only uses the prebuilt oracle32)")
    os.makedirs(OUT_DIR, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "Temp"))
        shutil.copy(REF_LIB, os.path.join(tmp, "cscc.lib"))
        run("ar", "x", "cscc.lib", cwd=tmp)
        run("objcopy", "-I", "pe-i386", "-O", "elf32-i386", "Temp/Convert.obj", "conv.o", cwd=tmp)
        # locate .text in the file and the PC-relative relocation sites
        text_off = None
        for line in run("objdump", "-h", "conv.o", cwd=tmp).splitlines():
            m = re.match(r"\s*\d+\s+\.text\s+([0-9a-f]+)\s+[0-9a-f]+\s+[0-9a-f]+\s+([0-9a-f]+)", line)
            if m:
                text_off = int(m.group(2), 16)
        assert text_off is not None
        sites, in_text = [], False
        for line in run("objdump", "-r", "conv.o", cwd=tmp).splitlines():
            if line.startswith("RELOCATION RECORDS FOR"):
                in_text = "[.text]" in line
            m = re.match(r"([0-9a-f]{8})\s+R_386_PC32\s+", line)
            if m and in_text:
                sites.append(int(m.group(1), 16))
        path = os.path.join(tmp, "conv.o")
        blob = bytearray(open(path, "rb").read())
        for s in sites:
            assert blob[text_off + s:text_off + s + 4] == b"\0\0\0\0", hex(s)
            blob[text_off + s:text_off + s + 4] = (-4 & 0xFFFFFFFF).to_bytes(4, "little")
        open(path, "wb").write(blob)
        args = []
        for old, new in RENAMES.items():
            args += ["--redefine-sym", f"{old}={new}"]
        run("objcopy", *args, "conv.o", "conv_r.o", cwd=tmp)
        run("gcc", "-m32", "-O1", "-ffreestanding", "-fno-pic", "-fno-stack-protector", "-fno-builtin",
            "-nostdlib", "-static", "-c", os.path.join(HERE, "harness32.c"), "-o", "harness.o", cwd=tmp)
        subprocess.run(["ld", "-m", "elf_i386", "-static", "-z", "noexecstack", "-o",
                        os.path.join(OUT_DIR, "oracle32"), "harness.o", "conv_r.o"],
                       cwd=tmp, check=True, capture_output=True, text=True)
    if verbose:
        print(f"built {os.path.join(OUT_DIR, 'oracle32')} ({len(sites)} PC32 sites patched)")
    return os.path.join(OUT_DIR, "oracle32")


if __name__ == "__main__":
    try:
        build()
    except Exception as e:  # noqa: BLE001
        print(f"oracle32 not built: {e}", file=sys.stderr)
        sys.exit(1)
