#!/usr/bin/env python3
"""Build the CPU checkers under oracle/ - TEST INFRASTRUCTURE ONLY.

  oracle/_ref/libtrb_ref.so   the reference's own our_gl.cpp + tgaimage.cpp compiled WHERE THEY
                              LIE under /root/reference (never copied) together with
                              oracle/ref_harness.cpp.  Only built when /root/reference exists
                              (this container); the prebuilt .so travels to the GPU box.
  oracle/libtrb_port.so       oracle/port_oracle.cpp, the self-contained restatement.

Flags: -O2 -ffp-contract=off (x86-64 SSE2 has no implicit FMA; the reference ships an MSVC
/fp:precise build without /arch:AVX2), so the doubles are bit-reproducible.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("TRB_REFERENCE_DIR", "/root/reference")
CXX = os.environ.get("TRB_CXX", "g++")  # not $CXX: the image points it at another toolchain
FLAGS = ["-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-fvisibility=hidden",
         "-Wall", "-Wno-unused-function", "-Wno-sign-compare"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources if os.path.exists(s))


def build_ref(force=False):
    srcs = [os.path.join(HERE, "ref_harness.cpp"), os.path.join(HERE, "post_restate.inc"),
            os.path.join(HERE, "..", "include", "trb.h")]
    out_dir = os.path.join(HERE, "_ref")
    out = os.path.join(out_dir, "libtrb_ref.so")
    ref_srcs = [os.path.join(REF, "our_gl.cpp"), os.path.join(REF, "tgaimage.cpp")]
    if not all(os.path.exists(s) for s in ref_srcs):
        return out if os.path.exists(out) else None
    os.makedirs(out_dir, exist_ok=True)
    if force or _newer(out, srcs + ref_srcs):
        cmd = [CXX] + FLAGS + ["-w", "-I", REF, srcs[0]] + ref_srcs + ["-o", out]
        subprocess.check_call(cmd)
    return out


def build_port(force=False):
    srcs = [os.path.join(HERE, "port_oracle.cpp"), os.path.join(HERE, "post_restate.inc"),
            os.path.join(HERE, "..", "include", "trb.h")]
    out = os.path.join(HERE, "libtrb_port.so")
    if not os.path.exists(srcs[0]):
        return None
    if force or _newer(out, srcs):
        subprocess.check_call([CXX] + FLAGS + [srcs[0], "-o", out, "-lpthread"])
    return out


def main():
    force = "--force" in sys.argv
    print("ref :", build_ref(force))
    print("port:", build_port(force))


if __name__ == "__main__":
    main()
