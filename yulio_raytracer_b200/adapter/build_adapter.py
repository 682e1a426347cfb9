#!/usr/bin/env python3
"""Builds lib/libdevice_cuda.so (the embree::Device plugin, adapter/device_cuda_plugin.cpp) and lib/plugin_smoke
(a caller that only knows devices/device/device.h). Both compile against the reference's own headers, taken from the
patched scratch overlay oracle/make_overlay.py creates (the unmodified affinespace.h does not parse under gcc, SURVEY
Appendix B #2), so this only runs where /root/reference is mounted; the built files travel to the GPU box."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
REPO = os.path.dirname(PKG)
OVL = os.path.join(REPO, "build", "oracle_overlay")
LIB = os.path.join(PKG, "lib")


def main():
    if not os.path.isdir(os.path.join(OVL, "devices", "device")):
        subprocess.check_call([sys.executable, os.path.join(REPO, "oracle", "make_overlay.py")])
    inc = ["-I" + OVL, "-I" + os.path.join(OVL, "common"), "-I" + os.path.join(OVL, "devices"), "-I" + os.path.join(REPO, "include")]
    base = ["g++", "-std=c++14", "-O2", "-msse4.2", "-fPIC", "-fpermissive", "-w", "-DNDEBUG"] + inc
    os.makedirs(LIB, exist_ok=True)
    subprocess.check_call(base + ["-shared", os.path.join(HERE, "device_cuda_plugin.cpp"), "-o", os.path.join(LIB, "libdevice_cuda.so"),
                                  "-L" + LIB, "-lyrt_device_cuda", "-Wl,-rpath,$ORIGIN", "-Wl,--no-undefined"])
    subprocess.check_call(base + [os.path.join(HERE, "plugin_smoke.cpp"), "-o", os.path.join(LIB, "plugin_smoke"), "-ldl"])
    print("built", os.path.join(LIB, "libdevice_cuda.so"), "and plugin_smoke")
    return 0


if __name__ == "__main__":
    sys.exit(main())
