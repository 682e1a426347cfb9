#!/usr/bin/env python3
"""Compiles the reference's vendored (and Yulio-patched) Assimp 3.2 — Collada importer + the post-processing steps of
aiProcessPreset_TargetRealtime_Quality only — from the sources where they lie under /root/reference/3rd party/assimp-3.2,
with g++ directly (not the reference's build system), into build/frontend_obj/assimp/libassimp_collada.a.
Used by frontend/build_frontend.py (the Collada -> Device front end is reused, not rewritten)."""
import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("YRT_REFERENCE", "/root/reference")
ASSIMP = os.path.join(REF, "3rd party", "assimp-3.2")
OBJ = os.path.join(REPO, "build", "frontend_obj", "assimp")
IMPORTERS = """3DS 3D AC ASE ASSBIN B3D BLEND BVH C4D COB CSM DXF FBX HMP IFC IRRMESH IRR LWO LWS MD2 MD3 MD5 MDC MDL MS3D NDO NFF OBJ OFF
OGRE OPENGEX PLY Q3BSP Q3D RAW SMD STL TERRAGEN XGL X""".split()


def main():
    os.makedirs(OBJ, exist_ok=True)
    defs = [f"-DASSIMP_BUILD_NO_{i}_IMPORTER" for i in IMPORTERS] + ["-DASSIMP_BUILD_NO_EXPORT", "-DASSIMP_BUILD_BOOST_WORKAROUND",
                                                                     "-DASSIMP_BUILD_NO_OWN_ZLIB", "-DNDEBUG"]
    inc = ["-I" + os.path.join(ASSIMP, "include"), "-I" + os.path.join(ASSIMP, "code"), "-I" + os.path.join(ASSIMP, "code", "BoostWorkaround"),
           "-I" + os.path.join(ASSIMP, "contrib", "irrXML"), "-I" + os.path.join(ASSIMP, "contrib", "ConvertUTF"), "-I" + os.path.join(ASSIMP, "contrib")]
    srcs = sorted(glob.glob(os.path.join(ASSIMP, "code", "*.cpp"))) + [os.path.join(ASSIMP, "contrib", "irrXML", "irrXML.cpp"),
                                                                       os.path.join(ASSIMP, "contrib", "ConvertUTF", "ConvertUTF.c")]
    skip = ("Exporter", "IFC", "FBX", "Blender", "Ogre", "OpenGEX", "C4D", "Q3BSP", "XGL", "AssbinExporter", "AssxmlExporter", "StepExporter")
    srcs = [s for s in srcs if not os.path.basename(s).startswith(skip)]

    def cc(src):
        obj = os.path.join(OBJ, os.path.basename(src) + ".o")
        if os.path.exists(obj) and os.path.getmtime(obj) >= os.path.getmtime(src):
            return obj, None
        comp = "gcc" if src.endswith(".c") else "g++"
        flags = ["-O2", "-fPIC", "-w", "-fpermissive"] + ([] if src.endswith(".c") else ["-std=c++11"])
        r = subprocess.run([comp] + flags + defs + inc + ["-c", src, "-o", obj], capture_output=True, text=True)
        return (obj, None) if r.returncode == 0 else (None, f"--- {src}\n{r.stderr[-1500:]}")

    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        res = list(ex.map(cc, srcs))
    errs = [e for _, e in res if e]
    if errs:
        sys.stderr.write("\n".join(errs[:5]) + f"\n{len(errs)} file(s) failed\n")
        return 1
    lib = os.path.join(OBJ, "libassimp_collada.a")
    if os.path.exists(lib):
        os.remove(lib)
    subprocess.check_call(["ar", "rcs", lib] + [o for o, _ in res])
    print("built", lib, len(res), "objects")
    return 0


if __name__ == "__main__":
    sys.exit(main())
