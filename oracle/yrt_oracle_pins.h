// TEST INFRASTRUCTURE ONLY (oracle). Force-included into the overlay copy of the
// reference's device_singleray by oracle/make_overlay.py; never used by the product.
//
// Holds the stated semantic pins P1/P2/P4 (see make_overlay.py for the list).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>

// ---- P1: counter-based replacement for libc rand() at pathtraceintegrator.cpp:151 ----
// The same function (same constants, same op order) lives on the device in
// yulio_raytracer_b200/csrc/pins.cuh; tests/test_pins.py compares the two.
static inline uint32_t yrt_fmix32(uint32_t h) {
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}
static inline uint32_t yrt_hash4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    uint32_t h = yrt_fmix32(a + 0x9E3779B9u);
    h = yrt_fmix32(h ^ (b + 0x7F4A7C15u));
    h = yrt_fmix32(h ^ (c + 0x94D049BBu));
    h = yrt_fmix32(h ^ (d + 0xBF58476Du));
    return h;
}
static inline float yrt_hash_unit(uint32_t h) {            // [0,1)
    return (float)(h >> 8) * (1.0f / 16777216.0f);
}

// ---- P1 + P2: jitter length of the dome-light shadow ray ----
// reference: 2*tMax*j*rand - tMax*j   (inf - inf = NaN when tMax == +inf, SURVEY F7: pinned to NaN)
static inline float yrt_oracle_shadow_jitter(float tMax, float jitter, float px, float py,
                                             unsigned depth, unsigned light) {
    if (std::isinf(tMax)) return std::numeric_limits<float>::quiet_NaN();   // P2
    uint32_t bx, by;
    std::memcpy(&bx, &px, 4); std::memcpy(&by, &py, 4);
    const float r = yrt_hash_unit(yrt_hash4(bx, by, depth, light));   // P1
    return 2.f * tMax * jitter * r - tMax * jitter;
}

// ---- P4: frame statistics hook (implemented in oracle_capi.cpp) ----
extern "C" void yrt_oracle_report_frame(double seconds, double rays);
extern "C" int  yrt_oracle_quiet();
