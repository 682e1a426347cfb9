// host_math.hpp — host-only helpers completing common.cuh: matrix inverse / normal transform, parsing of the
// 12-float transform, and image texel access with the reference's conversions.
// Reference: common/math/linearspace3.h:60-69,146; common/math/affinespace.h:70-88,166-182;
// common/image/image.h:82-84; common/math/color_sse.h:50-53.
#pragma once
#include <cstring>

#include "common.cuh"
#include "host_objects.hpp"

namespace yrt {

inline float lin3_det(const Lin3& m) { return dot(m.vx, cross(m.vy, m.vz)); }
// inverse().transposed() = rcp(det) * (cross(vy,vz), cross(vz,vx), cross(vx,vy)) as COLUMNS
inline Lin3 lin3_inverse_transposed(const Lin3& m) {
    const float r = rcpf(lin3_det(m));
    Lin3 o; o.vx = r * cross(m.vy, m.vz); o.vy = r * cross(m.vz, m.vx); o.vz = r * cross(m.vx, m.vy);
    return o;
}
inline Lin3 lin3_transposed(const Lin3& m) {
    Lin3 o; o.vx = V3(m.vx.x, m.vy.x, m.vz.x); o.vy = V3(m.vx.y, m.vy.y, m.vz.y); o.vz = V3(m.vx.z, m.vy.z, m.vz.z);
    return o;
}
inline Lin3 lin3_inverse(const Lin3& m) { return lin3_transposed(lin3_inverse_transposed(m)); }
inline V3 xfmNormal(const Aff3& m, V3 n) { return xfmVector(lin3_inverse_transposed(m.l), n); }
inline Aff3 aff3_inverse(const Aff3& a) { Aff3 r; r.l = lin3_inverse(a.l); r.p = -xfmVector(r.l, a.p); return r; }
inline Aff3 aff3_identity() { Aff3 a; a.l = lin3_identity(); a.p = V3(0.f); return a; }
inline bool aff3_is_identity(const Aff3& a) {
    return a.l.vx == V3(1, 0, 0) && a.l.vy == V3(0, 1, 0) && a.l.vz == V3(0, 0, 1) && a.p == V3(0.f);
}
inline Aff3 aff3_from_array(const float* v) {
    Aff3 a;
    if (!v) return aff3_identity();
    a.l.vx = V3(v[0], v[1], v[2]); a.l.vy = V3(v[3], v[4], v[5]); a.l.vz = V3(v[6], v[7], v[8]); a.p = V3(v[9], v[10], v[11]);
    return a;
}
inline void aff3_to_array(const Aff3& a, float* v) {
    v[0] = a.l.vx.x; v[1] = a.l.vx.y; v[2] = a.l.vx.z; v[3] = a.l.vy.x; v[4] = a.l.vy.y; v[5] = a.l.vy.z;
    v[6] = a.l.vz.x; v[7] = a.l.vz.y; v[8] = a.l.vz.z; v[9] = a.p.x; v[10] = a.p.y; v[11] = a.p.z;
}
// AffineSpace::lookAtPoint (affinespace.h:73-78)
inline Aff3 aff3_look_at_point(V3 eye, V3 point, V3 up) {
    const V3 Z = normalize(point - eye), U = normalize(cross(up, Z)), V = normalize(cross(Z, U));
    Aff3 a; a.l.vx = U; a.l.vy = V; a.l.vz = Z; a.p = eye; return a;
}
inline Aff3 aff3_scale(V3 s) { Aff3 a; a.l.vx = V3(s.x, 0, 0); a.l.vy = V3(0, s.y, 0); a.l.vz = V3(0, 0, s.z); a.p = V3(0.f); return a; }

// Image::get as Color (rgb) on the host
inline Col host_texel(const ImageObj& img, long long x, long long y) {
    const size_t i = (size_t)y * (size_t)img.width + (size_t)x;
    const float k = 1.f / 255.f;
    switch (img.format) {
    case TEX_RGB8: { const unsigned char* p = (const unsigned char*)img.pixels + 3 * i; return Col(p[0] * k, p[1] * k, p[2] * k); }
    case TEX_RGBA8: { const unsigned char* p = (const unsigned char*)img.pixels + 4 * i; return Col(p[0] * k, p[1] * k, p[2] * k); }
    case TEX_RGB_F32: { const float* p = (const float*)img.pixels + 3 * i; return Col(p[0], p[1], p[2]); }
    default: { const float* p = (const float*)img.pixels + 4 * i; return Col(p[0], p[1], p[2]); }
    }
}

}  // namespace yrt
