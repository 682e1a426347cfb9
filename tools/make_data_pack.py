#!/usr/bin/env python3
"""Copies the REAL texture sets of the two stripped Collada scenes from the reference mount into data/ (committed: test and benchmark data,
not source code). The .dae files themselves are git-lfs-stripped from the mount; their image sets are not:

  data/sponza/         models/Sponza/*.JPG                                  14 files, 0.5 MB   (BASELINE configs[2], C3)
  data/sample_scene/   sample_scene/22 Frederick St. good_tempo/*.{jpg,jpeg,png}   152 files, 26 MB   (BASELINE configs[3], C4: 273 MB as RGBA8,
                                                                             30 RGBA PNGs with alpha, twelve 1x1 JPEGs)

yulio_raytracer_b200/workloads.py reads them through rtNewImageFromFile, the way rtLoadTexture does (devices/device/loaders/loaders.cpp:29-61)."""
import os
import shutil
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("YRT_REFERENCE", "/root/reference")
SETS = {"sponza": (os.path.join(REF, "models", "Sponza"), (".jpg",)),
        "sample_scene": (os.path.join(REF, "sample_scene", "22 Frederick St. good_tempo"), (".jpg", ".jpeg", ".png"))}


def main():
    for name, (src, exts) in SETS.items():
        if not os.path.isdir(src):
            print(f"make_data_pack: {src} not found (reference mount absent): keeping data/{name} as committed", file=sys.stderr)
            continue
        dst = os.path.join(REPO, "data", name)
        os.makedirs(dst, exist_ok=True)
        n = b = 0
        for f in sorted(os.listdir(src)):
            p = os.path.join(src, f)
            if os.path.isfile(p) and f.lower().endswith(exts):
                shutil.copyfile(p, os.path.join(dst, f)); n += 1; b += os.path.getsize(p)
        print(f"data/{name}: {n} files, {b / 1e6:.1f} MB")
    return 0


if __name__ == "__main__":
    sys.exit(main())
