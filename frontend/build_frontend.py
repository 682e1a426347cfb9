#!/usr/bin/env python3
"""Builds yulio_raytracer_b200/lib/libyulio_rt.so (StartRT/WaitRT/StopRT/GetLastErrorRT/GetCurrentStatusRT, frontend/yulio_rt.cpp)
and lib/rt_test. Links the reference's own scene loaders (devices/device/loaders/*.cpp, from the patched scratch overlay) and its
vendored Assimp 3.2 (frontend/build_assimp.py), all compiled from the sources under /root/reference: needs the mount. The built
files travel to the GPU box."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = os.environ.get("YRT_REFERENCE", "/root/reference")
OVL = os.path.join(REPO, "build", "oracle_overlay")
OBJ = os.path.join(REPO, "build", "frontend_obj")
LIB = os.path.join(REPO, "yulio_raytracer_b200", "lib")
ASSIMP = os.path.join(REF, "3rd party", "assimp-3.2")

REF_SOURCES = ["devices/device/loaders/ColladaLoader.cpp", "devices/device/loaders/loaders.cpp", "devices/device/loaders/obj_loader.cpp",
               "devices/device/loaders/xml_loader.cpp", "devices/device/loaders/xml_parser.cpp", "devices/device/handle.cpp",
               "common/lexers/stringstream.cpp", "common/lexers/tokenstream.cpp", "common/sys/filename.cpp", "common/sys/stl/string.cpp",
               "common/sys/platform.cpp", "common/image/image.cpp", "common/image/ppm.cpp", "common/image/pfm.cpp", "common/image/tga.cpp"]


def main():
    if not os.path.isdir(os.path.join(OVL, "devices", "device")):
        subprocess.check_call([sys.executable, os.path.join(REPO, "oracle", "make_overlay.py")])
    subprocess.check_call([sys.executable, os.path.join(HERE, "build_assimp.py")])
    os.makedirs(OBJ, exist_ok=True); os.makedirs(LIB, exist_ok=True)
    inc = ["-I" + OVL, "-I" + os.path.join(OVL, "common"), "-I" + os.path.join(OVL, "devices"), "-I" + HERE,
           "-I" + os.path.join(ASSIMP, "include")]
    base = ["g++", "-std=c++14", "-O2", "-msse4.2", "-fPIC", "-fpermissive", "-w", "-DNDEBUG", "-pthread"] + inc
    jobs = [os.path.join(OVL, s) for s in REF_SOURCES] + [os.path.join(HERE, "yulio_rt.cpp"), os.path.join(HERE, "codec_stubs.cpp")]

    def cc(src):
        obj = os.path.join(OBJ, os.path.relpath(src, REPO).replace("/", "_") + ".o")
        r = subprocess.run(base + ["-c", src, "-o", obj], capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(f"--- {src}\n{r.stderr[-3000:]}\n"); raise SystemExit(1)
        return obj

    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        objs = list(ex.map(cc, jobs))
    so = os.path.join(LIB, "libyulio_rt.so")
    subprocess.check_call(["g++", "-shared", "-o", so] + objs + [os.path.join(OBJ, "assimp", "libassimp_collada.a"), "-ldl", "-pthread",
                                                                 "-Wl,--no-undefined", "-Wl,--exclude-libs,ALL"])
    subprocess.check_call(["g++", "-std=c++14", "-O2", "-I" + HERE, os.path.join(HERE, "rt_test.cpp"), "-o", os.path.join(LIB, "rt_test"),
                           "-L" + LIB, "-lyulio_rt", "-Wl,-rpath,$ORIGIN", "-pthread"])
    print("built", so, "and rt_test")
    return 0


if __name__ == "__main__":
    sys.exit(main())
