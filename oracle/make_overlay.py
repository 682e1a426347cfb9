#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY (oracle). Never imported by the product path.

Creates a scratch *overlay* copy of the reference's CPU path-tracer sources
(common/, devices/device/, devices/device_singleray/) under build/oracle_overlay/
(git-ignored, gpurun-ignored) and applies the small, mechanical patches needed to
compile the MSVC-only tree with g++ 13 on Linux, plus four documented semantic PINS
that make the oracle deterministic.  Nothing from the reference is committed to this
repository: the overlay is regenerated from /root/reference every build and only the
resulting shared objects land in oracle/_ref/.

Portability patches (no behaviour change), cf. SURVEY.md Appendix B:
  B1 common/simd/sseb.h:108            integer shuffle on a float mask needs casts under gcc
  B2 common/math/affinespace.h:153     `struct ArrayXf : final {`  is not C++
  B3 shapes/trianglemesh_full.cpp:106  `mesh->unlikely(motion.size())` typo
  B4 api/singleray_device.cpp:57       "materials/uber.h" vs Uber.h on a case-sensitive FS
  B5 textures/Bilinear.h, nearestneighbor.h   implicit Color -> Color4 conversion
  B6 common/sys/intrinsics.h           __rdtsc/__rdpmc clash with <x86intrin.h>

Semantic pins (each one is a *stated* deviation from the shipped reference, needed
because the shipped behaviour is either undefined, non-deterministic or hardware
dependent; DESIGN.md "Parity contract" lists them):
  P1 integrators/pathtraceintegrator.cpp:151   libc rand() shared by all threads ->
     counter hash of (pixel.x bits, pixel.y bits, depth, light index)  [SURVEY F6]
  P2 same line: tMaxShadowRay == +inf makes the jitter inf - inf = NaN under IEEE rules, so the shadow ray's tfar
     is NaN. Pinned to exactly that (never to a fast-math reassociation): tfar = NaN, and a NaN interval is
     empty in both the shim and the CUDA traversal, i.e. such shadow rays are never occluded [SURVEY F7]
  P3 common/math/{math.h,vector3f_sse.h,color_sse.h}: rcp/rsqrt built on the
     *approximate* SSE rcpps/rsqrtps (vendor specific bits) -> the exact 1/x and
     1/sqrt(x) the reference itself uses in its non-SSE branch (math.h:65,69) [SURVEY F5]
  P4 renderers/integratorrenderer.cpp: ray counter + timing exported through
     a C hook so bench.py can read Mrays/s without scraping stdout (no behaviour change).
  P7 shapes/disk.h:53-57: the apex vertex (index n) gets no normal / texture coordinate, so shading a disk reads element n of arrays
     of n elements: the apex is given the rim's values, normal (0,0,1) and texcoord (0,0).
  P6 renderers/debugrenderer.cpp:116: `cosineSampleHemisphere(rand.getFloat(), rand.getFloat(), Nf)` leaves the order of the two draws
     to the compiler (unspecified in C++; gcc and MSVC differ in practice): pinned to u first, then v.
  P5 textures/Bilinear.h:31-34: texels x+1 / y+1 are read past the allocation for a 1-pixel-wide / -tall image (the 1x1
     white fallback of a missing texture, and the sample scene's own 1x1 JPEGs): the neighbour index is clamped, which is
     the variant the reference left commented out two lines below (Bilinear.h:36-37). Images >= 2 px per axis are unaffected.
"""
import os
import shutil
import sys

REF = os.environ.get("YRT_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(ROOT), "build", "oracle_overlay")


def patch(rel, old, new, count=1):
    path = os.path.join(OUT, rel)
    with open(path, "r", encoding="latin-1") as f:
        s = f.read()
    n = s.count(old)
    if n < 1 or (count and n != count):
        raise SystemExit(f"overlay patch failed for {rel}: expected {count} match(es) of {old!r}, found {n}")
    s = s.replace(old, new)
    with open(path, "w", encoding="latin-1") as f:
        f.write(s)


def main():
    global OUT
    # --fast: the overlay of the BASELINE build (build_ref.py: -O3 -ffast-math, the reference's release flags, common/cmake/gcc.cmake:27):
    # portability patches and pins P1, P2, P4 only — the reference's own approximate rcp / rsqrt stay (no P3).
    fast = "--fast" in sys.argv
    if fast:
        OUT = OUT + "_fast"
    if not os.path.isdir(REF):
        raise SystemExit(f"reference tree not found at {REF}")
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    os.makedirs(os.path.join(OUT, "devices"))
    shutil.copytree(os.path.join(REF, "common"), os.path.join(OUT, "common"),
                    ignore=shutil.ignore_patterns("freeglut", "*.vcxproj*", "*.lib", "*.dll"))
    shutil.copytree(os.path.join(REF, "devices", "device"), os.path.join(OUT, "devices", "device"),
                    ignore=shutil.ignore_patterns("*.vcxproj*"))
    shutil.copytree(os.path.join(REF, "devices", "device_singleray"),
                    os.path.join(OUT, "devices", "device_singleray"),
                    ignore=shutil.ignore_patterns("*.vcxproj*"))
    for d, _, files in os.walk(OUT):
        os.chmod(d, 0o755)
        for f in files:
            os.chmod(os.path.join(d, f), 0o644)

    # ---- B1..B6 portability -------------------------------------------------
    patch("common/simd/sseb.h",
          "return _mm_shuffle_epi32(a, _MM_SHUFFLE(i3, i2, i1, i0));",
          "return _mm_castsi128_ps(_mm_shuffle_epi32(_mm_castps_si128(a), _MM_SHUFFLE(i3, i2, i1, i0)));")
    patch("common/math/affinespace.h", "struct ArrayXf : final {", "struct ArrayXf {")
    patch("devices/device_singleray/shapes/trianglemesh_full.cpp",
          "if (mesh->unlikely(motion.size())) {", "if (unlikely(mesh->motion.size())) {")
    patch("devices/device_singleray/api/singleray_device.cpp",
          '#include "materials/uber.h"', '#include "materials/Uber.h"')
    patch("devices/device_singleray/textures/Bilinear.h",
          "return invert ? Color4(1.f) - c : c;", "return invert ? Color4((Color4(1.f) - c).m128) : c;")
    # B5 + PIN P5 (clamped neighbour texel; the uncommented statement only — the commented-out variant below it stays as it is)
    patch("devices/device_singleray/textures/Bilinear.h",
          "const Color4 c = (image->get(x, y) * u_opposite + image->get(x + 1, y) * u_ratio) * v_opposite +\n"
          "\t\t\t\t(image->get(x, y + 1) * u_opposite + image->get(x + 1, y + 1) * u_ratio) * v_ratio;",
          "const int x1 = x + 1 > int(image->width) - 1 ? x : x + 1, y1 = y + 1 > int(image->height) - 1 ? y : y + 1; /* PIN P5 */\n"
          "\t\t\tconst Color4 c = Color4((__m128)((image->get(x, y) * u_opposite + image->get(x1, y) * u_ratio) * v_opposite +\n"
          "\t\t\t\t(image->get(x, y1) * u_opposite + image->get(x1, y1) * u_ratio) * v_ratio));")
    patch("devices/device_singleray/textures/nearestneighbor.h",
          "return invert ? Color4(1.f) - c : c;", "return invert ? Color4((Color4(1.f) - c).m128) : c;")
    patch("common/sys/intrinsics.h",
          "__forceinline uint64 __rdtsc()  {", "__forceinline uint64 __yrt_unused_rdtsc()  {")
    patch("common/sys/intrinsics.h", "#include <xmmintrin.h>", "#include <xmmintrin.h>\n#include <x86intrin.h>", count=0)
    patch("common/sys/intrinsics.h",
          "__forceinline uint64 __rdpmc(int i) {\n  uint32 high,low;",
          "__forceinline uint64 __yrt_unused_rdpmc(int i) {\n  uint32 high,low;")

    # ---- P3 exact reciprocal / reciprocal square root ------------------------
    if not fast:
      patch("common/math/math.h",
            "return _mm_cvtss_f32(_mm_sub_ps(_mm_add_ps(r, r), _mm_mul_ps(_mm_mul_ps(r, r), vx)));",
            "(void)r; return 1.0f / x; /* PIN P3 */")
      patch("common/math/math.h",
            "return _mm_cvtss_f32(c);",
            "(void)c; return 1.0f / sqrtf(x); /* PIN P3 */")
      patch("common/math/vector3f_sse.h",
            "return _mm_sub_ps(_mm_add_ps(r, r), _mm_mul_ps(_mm_mul_ps(r, r), a));",
            "(void)r; return _mm_div_ps(_mm_set1_ps(1.0f), a.m128); /* PIN P3 */")
      patch("common/math/vector3f_sse.h",
            "return _mm_add_ps(_mm_mul_ps(_mm_set1_ps(1.5f),r), _mm_mul_ps(_mm_mul_ps(_mm_mul_ps(a, _mm_set1_ps(-0.5f)), r), _mm_mul_ps(r, r)));",
            "(void)r; return _mm_div_ps(_mm_set1_ps(1.0f), _mm_sqrt_ps(a.m128)); /* PIN P3 */")
      patch("common/math/color_sse.h",
            "return _mm_sub_ps(_mm_add_ps(r, r), _mm_mul_ps(_mm_mul_ps(r, r), a));",
            "(void)r; return _mm_div_ps(_mm_set1_ps(1.0f), a.m128); /* PIN P3 */")
      patch("common/math/color_sse.h",
            "return _mm_add_ps(_mm_mul_ps(_mm_set1_ps(1.5f),r), _mm_mul_ps(_mm_mul_ps(_mm_mul_ps(a, _mm_set1_ps(-0.5f)), r), _mm_mul_ps(r, r)));",
            "(void)r; return _mm_div_ps(_mm_set1_ps(1.0f), _mm_sqrt_ps(a.m128)); /* PIN P3 */")

    # ---- P1 + P2 shadow-ray jitter --------------------------------------------
    patch("devices/device_singleray/integrators/pathtraceintegrator.cpp",
          "const float shadowRayJitterLength = 2.f * tMaxShadowRay * tMaxShadowJitter * random<float>() - tMaxShadowRay * tMaxShadowJitter;",
          "const float shadowRayJitterLength = yrt_oracle_shadow_jitter(tMaxShadowRay, tMaxShadowJitter, state.pixel.x, state.pixel.y, (unsigned)lightPath.depth, (unsigned)i); /* PIN P1+P2 */")
    patch("devices/device_singleray/integrators/pathtraceintegrator.cpp",
          '#include "integrators/pathtraceintegrator.h"',
          '#include "integrators/pathtraceintegrator.h"\n#include "yrt_oracle_pins.h"')

    # ---- P6 debug renderer: order of the two random numbers of a diffuse bounce ------------------
    patch("devices/device_singleray/renderers/debugrenderer.cpp",
          "new (&ray) Ray(ray.org+0.999f*ray.tfar*ray.dir,cosineSampleHemisphere(rand.getFloat(),rand.getFloat(),Nf),4.0f*float(ulp)/**hit.error*/);",
          "const float yrt_u = rand.getFloat(); const float yrt_v = rand.getFloat(); /* PIN P6 */\n"
          "\t\t  new (&ray) Ray(ray.org+0.999f*ray.tfar*ray.dir,cosineSampleHemisphere(yrt_u,yrt_v,Nf),4.0f*float(ulp)/**hit.error*/);")

    # ---- P7 disk apex attributes ------------------------------------------------------------------
    patch("devices/device_singleray/shapes/disk.h",
          "position.push_back(P+Vector3f(0,0,h));",
          "position.push_back(P+Vector3f(0,0,h)); normal.push_back(Vector3f(0.0f,0.0f,1.0f)); texcoord.push_back(Vec2f(0.0f,0.0f)); /* PIN P7 */")

    # ---- P4 ray counter / timing hook ------------------------------------------
    patch("devices/device_singleray/renderers/integratorrenderer.cpp",
          "std::cout << stream.str() << std::endl;",
          "yrt_oracle_report_frame(dt, (double)(size_t)atomicNumRays); if (!yrt_oracle_quiet()) std::cout << stream.str() << std::endl; /* PIN P4 */")
    patch("devices/device_singleray/renderers/integratorrenderer.cpp",
          '#include "renderers/integratorrenderer.h"',
          '#include "renderers/integratorrenderer.h"\n#include "yrt_oracle_pins.h"')
    print("overlay ready:", OUT)


if __name__ == "__main__":
    sys.exit(main())
