// shading.cuh — device restatement of the reference's per-hit shading pipeline:
// textures, postIntersect, materials -> BRDF lobes, lobe eval/sample, light sampling.
// Each function cites the reference code it restates; arithmetic order follows the reference
// (strict float, see common.cuh) so results are comparable at the bit level up to libm.
#pragma once
#include "common.cuh"

namespace yrt {

struct Col4 { float r, g, b, a; };

// ---- textures ---------------------------------------------------------------------------------
// ImageT::get + Color4(Col3c/Col4c/Col3f/Col4f)   common/image/image.h:82-84, common/math/color_sse.h:50-53
// NB: RGB8 texels get alpha = 1/255 (the reference multiplies the constant 1.0 lane by 1/255 too).
YRT_D Col4 texel(const TextureRec& t, int x, int y) {
    const size_t i = (size_t)y * (size_t)t.width + (size_t)x;
    const float k = 1.f / 255.f;
    Col4 c;
    if (t.format == TEX_RGBA8) { const uchar4 p = ((const uchar4*)t.data)[i]; c.r = p.x * k; c.g = p.y * k; c.b = p.z * k; c.a = p.w * k; }
    else if (t.format == TEX_RGB8) { const uint8_t* p = (const uint8_t*)t.data + 3 * i; c.r = p[0] * k; c.g = p[1] * k; c.b = p[2] * k; c.a = 1.0f * k; }
    else if (t.format == TEX_RGBA_F32) { const float4 p = ((const float4*)t.data)[i]; c.r = p.x; c.g = p.y; c.b = p.z; c.a = p.w; }
    else { const float* p = (const float*)t.data + 3 * i; c.r = p[0]; c.g = p[1]; c.b = p[2]; c.a = 1.0f; }
    return c;
}
YRT_D Col4 c4mul(Col4 c, float s) { Col4 r = {c.r * s, c.g * s, c.b * s, c.a * s}; return r; }
YRT_D Col4 c4add(Col4 a, Col4 b) { Col4 r = {a.r + b.r, a.g + b.g, a.b + b.b, a.a + b.a}; return r; }

// Bilinear::get  devices/device_singleray/textures/Bilinear.h:23-40   (repeat wrap of the base texel only, all
// four lanes filtered, no y flip);  NearestNeighbor::get  textures/nearestneighbor.h:40-47
YRT_D Col4 tex_get(const TextureRec& t, float px, float py) {
    const float s1 = px - floorf(px), t1 = py - floorf(py);
    Col4 c;
    if (t.bilinear) {
        const float u = s1 * t.width - .5f, v = t1 * t.height - .5f;
        const int x = iclamp(int(floorf(u)), 0, t.width - 2), y = iclamp(int(floorf(v)), 0, t.height - 2);
        const float ur = u - x, vr = v - y, uo = 1.f - ur, vo = 1.f - vr;
        // pin P5: a 1-pixel-wide / -tall image (the 1x1 white fallback of a missing texture, singleray_device.cpp:250, and the sample
        // scene's own 1x1 JPEGs) makes the reference read texel x+1 / y+1 past the allocation; the neighbour index is clamped instead
        // (the variant left commented out in Bilinear.h:36-37). Any image of 2 or more pixels per axis is unaffected.
        const int x1 = x + 1 > t.width - 1 ? x : x + 1, y1 = y + 1 > t.height - 1 ? y : y + 1;
        c = c4add(c4mul(c4add(c4mul(texel(t, x, y), uo), c4mul(texel(t, x1, y), ur)), vo),
                  c4mul(c4add(c4mul(texel(t, x, y1), uo), c4mul(texel(t, x1, y1), ur)), vr));
    } else {
        const int si = (int)(s1 * float(t.width)), ti = (int)(t1 * float(t.height));
        c = texel(t, iclamp(si, 0, t.width - 1), iclamp(ti, 0, t.height - 1));
    }
    if (t.invert) { c.r = 1.f - c.r; c.g = 1.f - c.g; c.b = 1.f - c.b; c.a = 1.f - c.a; }
    return c;
}

// ---- differential geometry -----------------------------------------------------------------
struct DG { V3 P, Ng, Ns; float s, t, error; int material, areaLight, illumMask, shadowMask; V3 Tx, Ty; };   // Tx, Ty: only the EXT instantiation

// TriangleMeshFull::postIntersect      shapes/trianglemesh_full.cpp:207-275
// TriangleMeshWithNormals::postIntersect shapes/trianglemesh_normals.cpp:140-162
// Triangle::postIntersect              shapes/triangle.h:84-93
// Tx/Ty (trianglemesh_full.cpp:252-270, trianglemesh_normals.cpp:154-155) are produced by the EXT instantiation only: BrushedMetal and
// a bump-mapped Obj are the materials that read them. Per-vertex tangent arrays ("tangent_x" / "tangent_y") are interpolated unnormalised
// like the reference does (:253-256, :263-266); without them the tangents come from the positions and texture coordinates.
template <bool EXT>
YRT_D void post_intersect(const SceneData& sc, V3 org, V3 dir, float t, float u, float v, int triIdx, float time, DG& dg) {
    // one 80-byte record per leaf-order triangle (bvh_build.cu: write_triangle) replaces geometry record -> indices -> 3 vertices
    const float4* h = sc.triShade + 5ull * (uint32_t)triIdx;
#ifndef YRT_SHADE_REC_EVICT_LAST
#define YRT_SHADE_REC_EVICT_LAST 1
#endif
#if YRT_SHADE_REC_EVICT_LAST
    const uint64_t pol = bvh_policy();                            // shading records are re-used across paths: last to leave L2 (bvh.cuh)
    const float4 r0 = bvh_ld(h, pol), r1 = bvh_ld(h + 1, pol), r2 = bvh_ld(h + 2, pol), r3 = bvh_ld(h + 3, pol), r4 = bvh_ld(h + 4, pol);
#else
    const float4 r0 = __ldg(h), r1 = __ldg(h + 1), r2 = __ldg(h + 2), r3 = __ldg(h + 3), r4 = __ldg(h + 4);
#endif
    const uint32_t flags = __float_as_uint(r0.w);
    const GeomRec& g = sc.geoms[flags >> 2];
    dg.material = g.material; dg.areaLight = g.areaLight; dg.illumMask = g.illumMask; dg.shadowMask = g.shadowMask;
    dg.P = org + t * dir;
    dg.Ng = V3(r0.x, r0.y, r0.z);                               // normalize(ray.Ng), or Triangle::Ng for a triangle shape
    if (EXT) { dg.Tx = V3(0.f); dg.Ty = V3(0.f); }
    // a moving triangle (trianglemesh_full.cpp:211-215): ray.Ng came from the vertices at the ray's time, so does the geometric normal here
    V3 m0(0.f), m1(0.f), m2(0.f); bool moving = false;
    if (sc.hasMotion) {
        const float4* tp = sc.tris + 3ull * (uint32_t)triIdx;
        const float4 q2 = __ldg(tp + 2);
        if (__float_as_uint(q2.w) & 2u) {                       // YRT_TRI_FLAG_MOTION
            moving = true;
            const float4 q0 = __ldg(tp), q1 = __ldg(tp + 1);
            const float4* mp = sc.triMotion + 3ull * (uint32_t)triIdx;
            const float4 d0 = __ldg(mp), d1 = __ldg(mp + 1), d2 = __ldg(mp + 2);
            m0 = V3(q0.x, q0.y, q0.z) + time * V3(d0.x, d0.y, d0.z); m1 = V3(q1.x, q1.y, q1.z) + time * V3(d1.x, d1.y, d1.z); m2 = V3(q2.x, q2.y, q2.z) + time * V3(d2.x, d2.y, d2.z);
            dg.Ng = normalize(cross(m0 - m1, m2 - m0));
        }
    }
    const float w = 1.0f - u - v;
    if (flags & 2u) { dg.s = r1.w * w + r3.w * u + r4.y * v; dg.t = r2.w * w + r4.x * u + r4.z * v; }
    else { dg.s = u; dg.t = v; }
    if (flags & 1u) {
        V3 Ns = w * V3(r1.x, r1.y, r1.z) + u * V3(r2.x, r2.y, r2.z) + v * V3(r3.x, r3.y, r3.z);
        const float len2 = dot(Ns, Ns);
        Ns = len2 > 0 ? Ns * rsqrtf_exact(len2) : dg.Ng;
        if (dot(Ns, dg.Ng) < 0) Ns = -Ns;
        dg.Ns = Ns;
    } else dg.Ns = dg.Ng;
    if (EXT && g.type != MESH_TRIANGLE) {
        const float4* tp = sc.tris + 3ull * (uint32_t)triIdx;
        const float4 q0 = __ldg(tp), q1 = __ldg(tp + 1), q2 = __ldg(tp + 2);
        const V3 p0 = moving ? m0 : V3(q0.x, q0.y, q0.z), dPdu = (moving ? m1 : V3(q1.x, q1.y, q1.z)) - p0, dPdv = (moving ? m2 : V3(q2.x, q2.y, q2.z)) - p0;
        if (g.type == MESH_NORMALS) { dg.Tx = dPdu; dg.Ty = dPdv; }
        else {
            float dsdu = 1.f, dtdu = 0.f, dsdv = 0.f, dtdv = 1.f;
            if (flags & 2u) { dsdu = r3.w - r1.w; dtdu = r4.x - r2.w; dsdv = r4.y - r1.w; dtdv = r4.z - r2.w; }
            int4 vi = make_int4(0, 0, 0, 0);
            if (g.tanXBase != YRT_NO_ATTR || g.tanYBase != YRT_NO_ATTR) vi = sc.indices[g.idxBase + (uint32_t)__float_as_int(r4.w)];
            if (g.tanXBase != YRT_NO_ATTR) {
                const float4 a = sc.tangents[g.tanXBase + vi.x], b = sc.tangents[g.tanXBase + vi.y], c = sc.tangents[g.tanXBase + vi.z];
                dg.Tx = w * V3(a.x, a.y, a.z) + u * V3(b.x, b.y, b.z) + v * V3(c.x, c.y, c.z);
            } else {
                const V3 dPds = normalize(dPdu * dtdv - dPdv * dtdu);
                dg.Tx = normalize(dPds - dot(dPds, dg.Ns) * dg.Ns);
            }
            if (g.tanYBase != YRT_NO_ATTR) {
                const float4 a = sc.tangents[g.tanYBase + vi.x], b = sc.tangents[g.tanYBase + vi.y], c = sc.tangents[g.tanYBase + vi.z];
                dg.Ty = w * V3(a.x, a.y, a.z) + u * V3(b.x, b.y, b.z) + v * V3(c.x, c.y, c.z);
            } else {
                const V3 dPdt = normalize(dPdv * dsdu - dPdu * dsdv);
                dg.Ty = normalize(dPdt - dot(dPdt, dg.Ns) * dg.Ns);
            }
        }
    }
    dg.error = rmax(fabsf(t), reduce_max(vabs(dg.P)));
}

// ---- BRDF lobes -------------------------------------------------------------------------------
// type bits: brdfs/brdf.h:25-45
#define BR_DIFFUSE_REFLECTION 0x00000001u
#define BR_GLOSSY_REFLECTION 0x00000010u
#define BR_SPECULAR_REFLECTION 0x00000100u
#define BR_SPECULAR_TRANSMISSION 0x01000000u
#define BR_DIFFUSE 0x000F000Fu
#define BR_TRANSMISSION 0xFFFF0000u
#define BR_ALL 0xFFFFFFFFu

enum LobeKind { LOBE_LAMBERTIAN, LOBE_TRANSMISSION, LOBE_SPECULAR, LOBE_REFLECTION, LOBE_DIEL_REFL, LOBE_DIEL_TRANS,
                LOBE_THIN_DIEL_TRANS, LOBE_CONST_DIEL_TRANS, LOBE_MICROFACET_UBER,
                // EXT instantiation only (Plastic, Metal, BrushedMetal, MetallicPaint, Velvet):
                LOBE_CONDUCTOR, LOBE_MICROFACET_METAL, LOBE_MICROFACET_ANISO, LOBE_LAYER_LAMBERT, LOBE_LAYER_GLITTER, LOBE_MINNAERT, LOBE_VELVETY };
struct Lobe { int kind; uint32_t type; Col c; float a, b; Col e, k; float x; };   // meaning of c,a,b depends on kind; e,k,x: EXT kinds
// The lobes of one hit and the per-lobe candidates of CompositedBRDF::sample are indexed per lane at run time. As local arrays
// they cost uncoalesced local-memory traffic (ncu r1: 2.7 of 32 bytes used per sector); they live in shared memory instead, word
// w of slot i of thread t at s[(i * WORDS + w) * stride + t]: bank = t, conflict-free for any per-lane i.
#define YRT_MAX_LOBES 3
#define YRT_LOBE_WORDS 7
#define YRT_LOBE_WORDS_EXT 14
#define YRT_CAND_WORDS 9
template <bool EXT> struct LobesT {
    static constexpr int WORDS = EXT ? YRT_LOBE_WORDS_EXT : YRT_LOBE_WORDS;
    float* s; float* cand; int stride; int n;
    YRT_D Lobe get(int i) const {
        const float* p = s + i * WORDS * stride; Lobe l;
        l.kind = __float_as_int(p[0]); l.type = __float_as_uint(p[stride]); l.c = Col(p[2 * stride], p[3 * stride], p[4 * stride]); l.a = p[5 * stride]; l.b = p[6 * stride];
        if (EXT) { l.e = Col(p[7 * stride], p[8 * stride], p[9 * stride]); l.k = Col(p[10 * stride], p[11 * stride], p[12 * stride]); l.x = p[13 * stride]; }
        else { l.e = Col(0.f); l.k = Col(0.f); l.x = 0.f; }
        return l;
    }
    YRT_D uint32_t type(int i) const { return __float_as_uint(s[(i * WORDS + 1) * stride]); }
};
template <bool EXT>
YRT_D void add_lobe(LobesT<EXT>& L, int kind, uint32_t type, Col c, float a = 0.f, float b = 0.f, Col e = Col(0.f), Col k = Col(0.f), float x = 0.f) {
    if (L.n < YRT_MAX_LOBES) {
        float* p = L.s + L.n * LobesT<EXT>::WORDS * L.stride; L.n++;
        p[0] = __int_as_float(kind); p[L.stride] = __uint_as_float(type); p[2 * L.stride] = c.x; p[3 * L.stride] = c.y; p[4 * L.stride] = c.z;
        p[5 * L.stride] = a; p[6 * L.stride] = b;
        if (EXT) { p[7 * L.stride] = e.x; p[8 * L.stride] = e.y; p[9 * L.stride] = e.z; p[10 * L.stride] = k.x; p[11 * L.stride] = k.y; p[12 * L.stride] = k.z; p[13 * L.stride] = x; }
    }
}

struct Sample3 { V3 v; float pdf; };

// brdfs/optics.h:30-39,79-85,101-119
YRT_D V3 reflect_v(V3 V, V3 N, float cosi) { return 2.0f * cosi * N - V; }
YRT_D V3 reflect_v(V3 V, V3 N) { return reflect_v(V, N, dot(V, N)); }
YRT_D float fresnel_diel3(float cosi, float cost, float eta) {
    const float Rper = (eta * cosi - cost) * rcpf(eta * cosi + cost);
    const float Rpar = (cosi - eta * cost) * rcpf(cosi + eta * cost);
    return 0.5f * (Rpar * Rpar + Rper * Rper);
}
YRT_D float fresnel_diel(float cosi, float eta, float* outCosT = nullptr) {
    const float k = 1.0f - eta * eta * (1.0f - cosi * cosi);
    if (k < 0.0f) return 1.0f;
    const float cost = sqrtf(k);
    if (outCosT) *outCosT = cost;
    return fresnel_diel3(cosi, cost, eta);
}
YRT_D Sample3 refract_v(V3 V, V3 N, float eta, float cosi, float& cost) {
    const float k = 1.0f - eta * eta * (1.0f - cosi * cosi);
    Sample3 s;
    if (k < 0.0f) { cost = 0.0f; s.v = V3(0.f); s.pdf = 0.0f; return s; }
    cost = sqrtf(k);
    s.v = eta * (cosi * N - V) - cost * N; s.pdf = eta * eta;
    return s;
}

// samplers/shapesampler.h:95-141
YRT_D Sample3 cosine_sample_hemisphere(float u, float v, V3 N) {
    const float phi = YRT_TWO_PI * u;
    const float cosTheta = sqrtf(v), sinTheta = sqrtf(1.0f - v);
    float sinPhi, cosPhi; YRT_SINCOS(phi, sinPhi, cosPhi);
    Sample3 s; s.v = xfmVector(frame(N), V3(cosPhi * sinTheta, sinPhi * sinTheta, cosTheta)); s.pdf = cosTheta * YRT_ONE_OVER_PI;
    return s;
}
YRT_D Sample3 power_cosine_sample_hemisphere(float u, float v, V3 N, float e) {
    const float phi = YRT_TWO_PI * u;
    const float cosTheta = YRT_POWF(v, rcpf(e + 1));
    const float sinTheta = sqrtf(rmax(0.f, 1.f - cosTheta * cosTheta));
    float sinPhi, cosPhi; YRT_SINCOS(phi, sinPhi, cosPhi);
    Sample3 s; s.v = xfmVector(frame(N), V3(cosPhi * sinTheta, sinPhi * sinTheta, cosTheta));
    s.pdf = (e + 1.0f) * YRT_POWF(cosTheta, e) * YRT_ONE_OVER_TWO_PI;
    return s;
}

// fresnelConductor  brdfs/optics.h:121-130
YRT_D Col fresnel_conductor(float cosi, Col eta, Col k) {
    const Col tmp = eta * eta + k * k;
    const Col dpar = tmp * cosi * cosi + 2.0f * eta * cosi + Col(1.f);
    const Col Rpar = (tmp * cosi * cosi - 2.0f * eta * cosi + Col(1.f)) * Col(rcpf(dpar.x), rcpf(dpar.y), rcpf(dpar.z));
    const Col dper = tmp + 2.0f * eta * cosi + Col(cosi * cosi);
    const Col Rper = (tmp - 2.0f * eta * cosi + Col(cosi * cosi)) * Col(rcpf(dper.x), rcpf(dper.y), rcpf(dper.z));
    return 0.5f * (Rpar + Rper);
}

// AnisotropicPowerCosineDistribution  brdfs/microfacet/anisotropic_power_cosine_distribution.h:28-76 (dx = dg.Tx, dy = dg.Ty, dz = dg.Ns)
YRT_D float aniso_eval(const DG& dg, float nx, float ny, V3 wh) {
    const float norm2 = sqrtf((nx + 2) * (ny + 2)) * YRT_ONE_OVER_TWO_PI;
    const float cosPhiH = dot(wh, dg.Tx), sinPhiH = dot(wh, dg.Ty), cosThetaH = dot(wh, dg.Ns);
    const float R = cosPhiH * cosPhiH + sinPhiH * sinPhiH;
    if (R == 0.0f) return norm2;
    const float n = (nx * (cosPhiH * cosPhiH) + ny * (sinPhiH * sinPhiH)) * rcpf(R);
    return norm2 * YRT_POWF(fabsf(cosThetaH), n);
}
YRT_D Sample3 aniso_sample(const DG& dg, float nx, float ny, float sx, float sy) {
    const float norm1 = sqrtf((nx + 1) * (ny + 1)) * YRT_ONE_OVER_TWO_PI;
    const float phi = YRT_TWO_PI * sx;
    float sinP, cosP; YRT_SINCOS(phi, sinP, cosP);
    const float sinPhi0 = sqrtf(nx + 1) * sinP, cosPhi0 = sqrtf(ny + 1) * cosP;
    const float norm = rsqrtf_exact(sinPhi0 * sinPhi0 + cosPhi0 * cosPhi0);
    const float sinPhi = sinPhi0 * norm, cosPhi = cosPhi0 * norm;
    const float n = nx * (cosPhi * cosPhi) + ny * (sinPhi * sinPhi);
    const float cosTheta = YRT_POWF(sy, rcpf(n + 1));
    const float sinTheta = sqrtf(rmax(0.f, 1.f - cosTheta * cosTheta));
    Sample3 s; s.pdf = norm1 * YRT_POWF(cosTheta, n);
    const V3 h(cosPhi * sinTheta, sinPhi * sinTheta, cosTheta);
    s.v = h.x * dg.Tx + h.y * dg.Ty + h.z * dg.Ns;
    return s;
}

// Microfacet<Fresnel, Distribution>::eval  brdfs/microfacet.h:43-58 with F and D supplied by the caller's lobe kind
template <bool EXT> YRT_D Col lobe_eval(const Lobe& l, V3 wo, const DG& dg, V3 wi);
YRT_D Col microfacet_eval(const Lobe& l, V3 wo, const DG& dg, V3 wi, int fresnelKind /*0 dielectric l.a, 1 conductor l.e,l.k*/, int distKind /*0 power cosine n, 1 anisotropic*/, float n, float ny) {
    if (dot(wi, dg.Ng) <= 0) return Col(0.f);
    const float cosThetaO = dot(wo, dg.Ns), cosThetaI = dot(wi, dg.Ns);
    if (cosThetaI <= 0.0f || cosThetaO <= 0.0f) return Col(0.f);
    const V3 wh = normalize(wi + wo);
    const float cosThetaH = dot(wh, dg.Ns), cosTheta = dot(wi, wh);
    const Col F = fresnelKind == 0 ? Col(fresnel_diel(cosTheta, l.a)) : fresnel_conductor(cosTheta, l.e, l.k);
    const float D = distKind == 0 ? ((n + 2) * YRT_ONE_OVER_TWO_PI) * YRT_POWF(fabsf(dot(wh, dg.Ns)), n) : aniso_eval(dg, n, ny, wh);
    const float rcpCosTheta = rcpf(cosTheta);            // once: the out-of-line reciprocal is opaque to common-subexpression elimination
    const float G = rmin(rmin(1.0f, 2.0f * cosThetaH * cosThetaO * rcpCosTheta), 2.0f * cosThetaH * cosThetaI * rcpCosTheta);
    return l.c * D * G * F * rcpf(4.0f * cosThetaO);
}

// Lambertian::eval lambertian.h:35-37; Specular::eval specular.h:34-38; Microfacet::eval microfacet.h:43-58
// (+ PowerCosineDistribution::eval power_cosine_distribution.h:35-39, FresnelDielectric::eval fresnel.h:80-82);
// EXT: DielectricLayer::eval dielectriclayer.h:44-56, Minnaert::eval minnaert.h:33-37, Velvety::eval velvety.h:33-39;
// all specular lobes evaluate to zero.
template <bool EXT>
YRT_D Col lobe_eval(const Lobe& l, V3 wo, const DG& dg, V3 wi) {
    switch (l.kind) {
    case LOBE_LAMBERTIAN: return l.c * YRT_ONE_OVER_PI * rclamp(dot(wi, dg.Ns));
    case LOBE_SPECULAR: {
        const V3 r = reflect_v(wo, dg.Ns);
        if (dot(r, wi) < 0) return Col(0.f);
        return l.c * (l.a + 2) * (1.0f / (2.0f * YRT_PI)) * YRT_POWF(dot(r, wi), l.a) * rclamp(dot(wi, dg.Ns));
    }
    case LOBE_REFLECTION: return l.c;
    case LOBE_MICROFACET_UBER: return microfacet_eval(l, wo, dg, wi, 0, 0, l.b, 0.f);      // l.a = etai/etat, l.b = n
    default: break;
    }
    if (EXT) switch (l.kind) {
    case LOBE_MICROFACET_METAL: return microfacet_eval(l, wo, dg, wi, 1, 0, l.a, 0.f);    // l.a = n
    case LOBE_MICROFACET_ANISO: return microfacet_eval(l, wo, dg, wi, 1, 1, l.a, l.b);    // l.a = nx, l.b = ny
    case LOBE_MINNAERT: {
        const float cosThetaI = rclamp(dot(wi, dg.Ns));
        const float backScatter = YRT_POWF(rclamp(dot(wo, wi)), l.a);
        return l.c * backScatter * cosThetaI * rcpf(YRT_PI);                             // Color / float = a * rcp(b)  (color_sse.h:162)
    }
    case LOBE_VELVETY: {
        const float cosThetaO = rclamp(dot(wo, dg.Ns)), cosThetaI = rclamp(dot(wi, dg.Ns));
        const float sinThetaO = sqrtf(1.0f - cosThetaO * cosThetaO);
        const float horizonScatter = YRT_POWF(sinThetaO, l.a);
        return l.c * horizonScatter * cosThetaI * rcpf(YRT_PI);
    }
    case LOBE_LAYER_LAMBERT: case LOBE_LAYER_GLITTER: {                                    // l.a = etait, l.b = etati, T = one
        const float cosThetaO = dot(wo, dg.Ns), cosThetaI = dot(wi, dg.Ns);
        if (cosThetaI <= 0.0f || cosThetaO <= 0.0f) return Col(0.f);
        float cosThetaO1, cosThetaI1;
        const V3 wo1 = refract_v(wo, dg.Ns, l.a, cosThetaO, cosThetaO1).v, wi1 = refract_v(wi, dg.Ns, l.a, cosThetaI, cosThetaI1).v;
        const float Fi = 1.0f - fresnel_diel3(cosThetaI, cosThetaI1, l.a);
        Col Fg;
        if (l.kind == LOBE_LAYER_LAMBERT) Fg = l.c * YRT_ONE_OVER_PI * rclamp(dot(-wi1, dg.Ns));
        else Fg = microfacet_eval(l, -wo1, dg, -wi1, 1, 0, l.x, 0.f);                      // glitter: Microfacet<FresnelConductor, PowerCosine(l.x)>
        const float Fo = 1.0f - fresnel_diel3(cosThetaO, cosThetaO1, l.a);
        return Col(Fo) * Fg * Fi;
    }
    default: break;
    }
    return Col(0.f);
}

// The per-lobe sample() methods: lambertian.h:39-41, specular.h:40-42, transmission.h:38-40, reflection.h:40-43,
// dielectric.h:39-45 (DielectricReflection), :80-87 (DielectricTransmission), :122-132 (ThinDielectricTransmission),
// :185-189 (ConstDielectricTransmission), microfacet.h:60-67.
// Microfacet::sample  microfacet.h:60-67 around a sampled half vector
YRT_D Col microfacet_finish(const Lobe& l, V3 wo, const DG& dg, Sample3& wi, Sample3 wh, int fresnelKind, int distKind, float n, float ny) {
    wi.v = reflect_v(wo, wh.v); wi.pdf = wh.pdf * rcpf(4.0f * fabsf(dot(wo, wh.v)));
    if (dot(wi.v, dg.Ns) <= 0.0f) return Col(0.f);
    return microfacet_eval(l, wo, dg, wi.v, fresnelKind, distKind, n, ny);
}
// PowerCosineDistribution::sample  power_cosine_distribution.h:43-51
YRT_D Sample3 power_cosine_half_vector(const DG& dg, float n, float sx, float sy) {
    const float phi = YRT_TWO_PI * sx;
    float sinPhi, cosPhi; YRT_SINCOS(phi, sinPhi, cosPhi);
    const float cosTheta = YRT_POWF(sy, rcpf(n + 1));
    const float sinTheta = sqrtf(rmax(0.f, 1.f - cosTheta * cosTheta));
    Sample3 wh; wh.v = xfmVector(frame(dg.Ns), V3(cosPhi * sinTheta, sinPhi * sinTheta, cosTheta));
    wh.pdf = ((n + 1) * YRT_ONE_OVER_TWO_PI) * YRT_POWF(cosTheta, n);
    return wh;
}

template <bool EXT>
YRT_D Col lobe_sample(const Lobe& l, V3 wo, const DG& dg, Sample3& wi, float sx, float sy) {
    switch (l.kind) {
    case LOBE_LAMBERTIAN: wi = cosine_sample_hemisphere(sx, sy, dg.Ns); return lobe_eval<EXT>(l, wo, dg, wi.v);
    case LOBE_SPECULAR: wi = power_cosine_sample_hemisphere(sx, sy, reflect_v(wo, dg.Ns), l.a); return lobe_eval<EXT>(l, wo, dg, wi.v);
    case LOBE_TRANSMISSION: wi.v = -wo; wi.pdf = 1.0f; return l.c;
    case LOBE_REFLECTION: wi.v = reflect_v(wo, dg.Ns); wi.pdf = 1.0f; return l.c;
    case LOBE_DIEL_REFL: {
        const float cosThetaO = rclamp(dot(wo, dg.Ns));
        wi.v = reflect_v(wo, dg.Ns, cosThetaO); wi.pdf = 1.0f;
        return l.b * Col(fresnel_diel(cosThetaO, l.a));                                // l.b = alpha
    }
    case LOBE_DIEL_TRANS: {
        const float cosThetaO = rclamp(dot(wo, dg.Ns));
        float cosThetaI;
        wi = refract_v(wo, dg.Ns, l.a, cosThetaO, cosThetaI);
        return Col(1.0f - fresnel_diel3(cosThetaO, cosThetaI, l.a));
    }
    case LOBE_THIN_DIEL_TRANS: {
        wi.v = -wo; wi.pdf = 1.0f;
        const float cosTheta = rclamp(dot(wo, dg.Ns));
        if (cosTheta <= 0.0f) return Col(0.f);
        const float alpha = l.b * rcpf(cosTheta);                                      // l.b = thickness, l.c = logT
        const Col e(expf(l.c.x * alpha), expf(l.c.y * alpha), expf(l.c.z * alpha));
        return e * (1.f - fresnel_diel(cosTheta, l.a));
    }
    case LOBE_CONST_DIEL_TRANS: {
        wi.v = -wo; wi.pdf = 1.0f;
        const float cosTheta = rclamp(dot(wo, dg.Ns));
        return cosTheta <= 0.0f ? Col(0.f) : l.c;
    }
    case LOBE_MICROFACET_UBER: {
        wi.v = V3(0.f); wi.pdf = 0.f;
        if (dot(wo, dg.Ns) <= 0.0f) return Col(0.f);
        return microfacet_finish(l, wo, dg, wi, power_cosine_half_vector(dg, l.b, sx, sy), 0, 0, l.b, 0.f);
    }
    default: break;
    }
    wi.v = V3(0.f); wi.pdf = 0.f;
    if (EXT) switch (l.kind) {
    case LOBE_CONDUCTOR: wi.v = reflect_v(wo, dg.Ns); wi.pdf = 1.0f; return l.c * fresnel_conductor(dot(wo, dg.Ns), l.e, l.k);   // conductor.h:41-44
    case LOBE_MICROFACET_METAL:
        if (dot(wo, dg.Ns) <= 0.0f) return Col(0.f);
        return microfacet_finish(l, wo, dg, wi, power_cosine_half_vector(dg, l.a, sx, sy), 1, 0, l.a, 0.f);
    case LOBE_MICROFACET_ANISO:
        if (dot(wo, dg.Ns) <= 0.0f) return Col(0.f);
        return microfacet_finish(l, wo, dg, wi, aniso_sample(dg, l.a, l.b, sx, sy), 1, 1, l.a, l.b);
    case LOBE_MINNAERT: case LOBE_VELVETY: wi = cosine_sample_hemisphere(sx, sy, dg.Ns); return lobe_eval<EXT>(l, wo, dg, wi.v);
    case LOBE_LAYER_LAMBERT: case LOBE_LAYER_GLITTER: {                                    // DielectricLayer::sample  dielectriclayer.h:58-80
        const float cosThetaO = dot(wo, dg.Ns);
        if (cosThetaO <= 0.0f) return Col(0.f);
        float cosThetaO1; const Sample3 wo1 = refract_v(wo, dg.Ns, l.a, cosThetaO, cosThetaO1);
        Sample3 wi1; wi1.v = V3(0.f); wi1.pdf = 0.f; Col Fg;
        const V3 wg = -wo1.v;
        if (l.kind == LOBE_LAYER_LAMBERT) { wi1 = cosine_sample_hemisphere(sx, sy, dg.Ns); Fg = l.c * YRT_ONE_OVER_PI * rclamp(dot(wi1.v, dg.Ns)); }
        else if (dot(wg, dg.Ns) <= 0.0f) Fg = Col(0.f);
        else Fg = microfacet_finish(l, wg, dg, wi1, power_cosine_half_vector(dg, l.x, sx, sy), 1, 0, l.x, 0.f);
        const float cosThetaI1 = dot(wi1.v, dg.Ns);
        if (cosThetaI1 <= 0.0f) return Col(0.f);
        float cosThetaI; const Sample3 wi0 = refract_v(-wi1.v, -dg.Ns, l.b, cosThetaI1, cosThetaI);
        if (wi0.pdf == 0.0f) return Col(0.f);
        wi.v = wi0.v; wi.pdf = wi1.pdf;
        const float Fi = 1.0f - fresnel_diel3(cosThetaI, cosThetaI1, l.a);
        const float Fo = 1.0f - fresnel_diel3(cosThetaO, cosThetaO1, l.a);
        return Col(Fo) * Fg * Fi;
    }
    default: break;
    }
    return Col(0.f);
}

// CompositedBRDF::eval  brdfs/compositedbrdf.h:74-80
template <bool EXT>
YRT_D Col lobes_eval(const LobesT<EXT>& L, V3 wo, const DG& dg, V3 wi, uint32_t typeMask) {
    Col c(0.f);
#pragma unroll 1
    for (int i = 0; i < L.n; i++) { const Lobe l = L.get(i); if (l.type & typeMask) c += lobe_eval<EXT>(l, wo, dg, wi); }
    return c;
}

// CompositedBRDF::sample  brdfs/compositedbrdf.h:119-181
template <bool EXT>
YRT_D Col lobes_sample(const LobesT<EXT>& L, V3 wo, const DG& dg, Sample3& wiOut, uint32_t& typeOut, float sx, float sy, float ss, uint32_t typeMask) {
    float sum = 0.0f; int num = 0;
    const int st = L.stride;
#pragma unroll 1
    for (int i = 0; i < L.n; i++) {
        const Lobe l = L.get(i);
        if (!(l.type & typeMask)) continue;
        Sample3 wi; const Col c = lobe_sample<EXT>(l, wo, dg, wi, sx, sy);
        if (c == Col(0.f) || wi.pdf <= 0.0f) continue;
        const float f = (c.x + c.y + c.z) * rcpf(wi.pdf);
        sum += f;
        float* q = L.cand + num * YRT_CAND_WORDS * st; num++;
        q[0] = f; q[st] = c.x; q[2 * st] = c.y; q[3 * st] = c.z; q[4 * st] = wi.v.x; q[5 * st] = wi.v.y; q[6 * st] = wi.v.z; q[7 * st] = wi.pdf;
        q[8 * st] = __uint_as_float(l.type);
    }
    if (num == 0) { wiOut.v = V3(0.f); wiOut.pdf = 0.f; typeOut = 0; return Col(0.f); }
    // f[i] /= sum; d[0] = f[0]; d[i] = d[i-1] + f[i]; d[num-1] = 1; first i with ss <= d[i]
    int i = 0; float d = 0.f, fi = 0.f;
#pragma unroll 1
    for (;; i++) {
        fi = L.cand[i * YRT_CAND_WORDS * st] / sum;
        d = (i == 0) ? fi : d + fi;
        if (i >= num - 1 || !(ss > d)) break;
    }
    const float* q = L.cand + i * YRT_CAND_WORDS * st;
    wiOut.v = V3(q[4 * st], q[5 * st], q[6 * st]); wiOut.pdf = q[7 * st] * fi;
    typeOut = __float_as_uint(q[8 * st]);
    return Col(q[st], q[2 * st], q[3 * st]);
}

// ---- materials ----------------------------------------------------------------------------------
// Matte matte.h:35-37; Obj obj.h:50-69; Uber Uber.h:34-69; MatteTextured matte_textured.h:39-41;
// Dielectric dielectric.h:57-69; ThinDielectric thindielectric.h:44-60; Mirror mirror.h:36-38
template <bool EXT>
YRT_D void material_shade(const SceneData& sc, const MaterialRec& m, DG& dg, Col mediumT, float mediumEta, LobesT<EXT>& L) {
    L.n = 0;
    switch (m.type) {
    case MAT_MATTE: add_lobe(L, LOBE_LAMBERTIAN, BR_DIFFUSE_REFLECTION, m.c0); break;
    case MAT_MIRROR: add_lobe(L, LOBE_REFLECTION, BR_SPECULAR_REFLECTION, m.c0); break;
    case MAT_MATTE_TEXTURED:
        if (m.tex[0] >= 0) { const Col4 c = tex_get(sc.textures[m.tex[0]], m.dsx * dg.s + m.s0x, m.dsy * dg.t + m.s0y); add_lobe(L, LOBE_LAMBERTIAN, BR_DIFFUSE_REFLECTION, Col(c.r, c.g, c.b)); }
        break;
    case MAT_OBJ: {
        if (EXT && m.tex[4] >= 0) {                                                       // bump map bends the shading normal (obj.h:52-56)
            const Col4 bump = tex_get(sc.textures[m.tex[4]], dg.s, dg.t);
            const V3 b(2.0f * bump.r - 1.0f, 2.0f * bump.g - 1.0f, 2.0f * bump.b - 1.0f);
            dg.Ns = normalize(b.x * dg.Tx + b.y * dg.Ty + b.z * dg.Ns);
        }
        float d = m.f[0]; if (m.tex[0] >= 0) d *= tex_get(sc.textures[m.tex[0]], dg.s, dg.t).r;
        if (d < 1.0f) add_lobe(L, LOBE_TRANSMISSION, BR_SPECULAR_TRANSMISSION, Col(1.0f - d));
        Col Kd = d * m.c0; if (m.tex[1] >= 0) { const Col4 c = tex_get(sc.textures[m.tex[1]], dg.s, dg.t); Kd *= Col(c.r, c.g, c.b); }
        if (Kd != Col(0.f)) add_lobe(L, LOBE_LAMBERTIAN, BR_DIFFUSE_REFLECTION, Kd);
        float Ns = m.f[1]; if (m.tex[3] >= 0) Ns *= tex_get(sc.textures[m.tex[3]], dg.s, dg.t).r;
        Col Ks = d * m.c1; if (m.tex[2] >= 0) { const Col4 c = tex_get(sc.textures[m.tex[2]], dg.s, dg.t); Ks *= Col(c.r, c.g, c.b); }
        if (Ks != Col(0.f)) add_lobe(L, LOBE_SPECULAR, BR_GLOSSY_REFLECTION, Ks, Ns);
        break;
    }
    case MAT_UBER: {
        Col4 dc = {m.c0.x, m.c0.y, m.c0.z, 1.f};
        float alpha = 1.f, opacity = 0.f;
        if (m.tex[0] >= 0) { dc = tex_get(sc.textures[m.tex[0]], m.dsx * dg.s + m.s0x, m.dsy * dg.t + m.s0y); alpha = dc.a; opacity = 1.f - alpha; }
        add_lobe(L, LOBE_LAMBERTIAN, BR_DIFFUSE_REFLECTION, Col(dc.r * alpha, dc.g * alpha, dc.b * alpha));
        if (alpha < 1.f) add_lobe(L, LOBE_CONST_DIEL_TRANS, BR_SPECULAR_TRANSMISSION, Col(opacity));
        const float roughness = m.f[1], reflectivity = m.f[2], rcpRoughness = m.f[3], etait = m.f[4];   // f[4] = 1.f * rcp(eta), computed once at commit
        if (reflectivity > 0.f) add_lobe(L, LOBE_DIEL_REFL, BR_SPECULAR_REFLECTION, Col(0.f), etait, alpha * reflectivity);
        else if (roughness == 0.f) add_lobe(L, LOBE_DIEL_REFL, BR_SPECULAR_REFLECTION, Col(0.f), etait, alpha);
        else add_lobe(L, LOBE_MICROFACET_UBER, BR_GLOSSY_REFLECTION, Col(alpha), etait, rcpRoughness);
        break;
    }
    case MAT_DIELECTRIC: {
        // currentMedium == mediumOutside ? (outside -> inside) : (inside -> outside)   dielectric.h:57-69
        const bool outside = (mediumT == m.tOutside) && (mediumEta == m.etaOutside);
        const float eta = outside ? m.etaOutside * rcpf(m.etaInside) : m.etaInside * rcpf(m.etaOutside);
        add_lobe(L, LOBE_DIEL_REFL, BR_SPECULAR_REFLECTION, Col(0.f), eta, 1.f);
        add_lobe(L, LOBE_DIEL_TRANS, BR_SPECULAR_TRANSMISSION, Col(0.f), eta);
        break;
    }
    case MAT_THIN_DIELECTRIC: {
        const float eta = m.f[0], thickness = m.f[1], transparency = m.f[2];
        add_lobe(L, LOBE_DIEL_REFL, BR_SPECULAR_REFLECTION, Col(0.f), 1.0f * rcpf(eta), 1.f);
        Col4 dc = {m.c0.x, m.c0.y, m.c0.z, 1.f};
        if (m.tex[0] >= 0) dc = tex_get(sc.textures[m.tex[0]], m.dsx * dg.s + m.s0x, m.dsy * dg.t + m.s0y);
        const Col T(dc.r * transparency, dc.g * transparency, dc.b * transparency);
        add_lobe(L, LOBE_THIN_DIEL_TRANS, BR_SPECULAR_TRANSMISSION, Col(logf(T.x), logf(T.y), logf(T.z)), 1.f * rcpf(eta), thickness);
        break;
    }
    default: break;
    }
    if (EXT) switch (m.type) {
    case MAT_PLASTIC: {                                                               // plastic.h:38-47
        const float eta = m.f[0], roughness = m.f[1], rcpRoughness = m.f[3];
        add_lobe(L, LOBE_LAYER_LAMBERT, BR_DIFFUSE_REFLECTION, m.c0, 1.0f * rcpf(eta), eta * rcpf(1.0f));
        if (roughness == 0.0f) add_lobe(L, LOBE_DIEL_REFL, BR_SPECULAR_REFLECTION, Col(0.f), 1.0f * rcpf(eta), 1.f);
        else add_lobe(L, LOBE_MICROFACET_UBER, BR_GLOSSY_REFLECTION, Col(1.f), 1.0f * rcpf(eta), rcpRoughness);
        break;
    }
    case MAT_METAL:                                                                   // metal.h:43-50
        if (m.f[1] == 0.0f) add_lobe(L, LOBE_CONDUCTOR, BR_SPECULAR_REFLECTION, m.c0, 0.f, 0.f, m.c1, m.c2);
        else add_lobe(L, LOBE_MICROFACET_METAL, BR_GLOSSY_REFLECTION, m.c0, m.f[3], 0.f, m.c1, m.c2);
        break;
    case MAT_BRUSHED_METAL:                                                           // brushedmetal.h:46-55
        if (m.f[1] == 0.0f || m.f[2] == 0.0f) add_lobe(L, LOBE_CONDUCTOR, BR_SPECULAR_REFLECTION, m.c0, 0.f, 0.f, m.c1, m.c2);
        else add_lobe(L, LOBE_MICROFACET_ANISO, BR_GLOSSY_REFLECTION, m.c0, m.f[3], m.f[4], m.c1, m.c2);
        break;
    case MAT_METALLIC_PAINT: {                                                        // metallicpaint.h:35-62
        const float eta = m.f[0], glitterSpread = m.f[1];
        add_lobe(L, LOBE_DIEL_REFL, BR_SPECULAR_REFLECTION, Col(0.f), 1.0f * rcpf(eta), 1.f);
        add_lobe(L, LOBE_LAYER_LAMBERT, BR_DIFFUSE_REFLECTION, m.c0, 1.0f * rcpf(eta), eta * rcpf(1.0f));
        if (glitterSpread != 0 && m.c1 != Col(0.f))
            add_lobe(L, LOBE_LAYER_GLITTER, BR_GLOSSY_REFLECTION, m.c1, 1.0f * rcpf(eta), eta * rcpf(1.0f), Col(0.62f), Col(4.8f), rcpf(glitterSpread));
        break;
    }
    case MAT_VELVET:                                                                  // velvet.h:38-41
        add_lobe(L, LOBE_MINNAERT, BR_DIFFUSE_REFLECTION, m.c0, m.f[0]);
        add_lobe(L, LOBE_VELVETY, BR_DIFFUSE_REFLECTION, m.c1, m.f[1]);
        break;
    default: break;
    }
}

// ---- lights -------------------------------------------------------------------------------------
struct LightSampleD { V3 wi; float pdf; float tMax; Col L; };

// samplers/shapesampler.h (uniformSampleTriangle, uniformSampleCone)
YRT_D V3 uniform_sample_triangle(float u, float v, V3 A, V3 B, V3 C) {
    const float su = sqrtf(u);
    return C + (1.0f - su) * (A - C) + (v * su) * (B - C);
}

// HDRILight::Le  lights/hdrilight.cpp:58-86 (lat-long lookup, manual bilinear with x wrap)
YRT_D Col hdri_Le(const SceneData& sc, const LightRec& l, V3 wo) {
    const V3 wi = xfmVector(l.world2local.l, -wo);
    const float theta = acosf(rclamp(wi.y, -1.0f, 1.0f));
    float phi = atan2f(-wi.z, -wi.x);
    if (phi < 0) phi += 2.0f * YRT_PI;
    const float u = 1.0f - (phi * YRT_ONE_OVER_TWO_PI), v = theta * YRT_ONE_OVER_PI;
    const TextureRec& t = sc.textures[l.image];
    const int width = t.width, height = t.height;
    long long x = (long long)(u * width); x = x < 0 ? 0 : (x > width - 1 ? width - 1 : x);
    long long xNext = x + 1; if (xNext == width) xNext = 0;
    const float alpha = u * width - x;
    long long y = (long long)(v * height); y = y < 0 ? 0 : (y > height - 1 ? height - 1 : y);
    long long yNext = y + 1; if (yNext == height) yNext = height - 1;
    const float beta = v * height - y;
    const Col4 a0 = texel(t, (int)x, (int)y), a1 = texel(t, (int)xNext, (int)y), a2 = texel(t, (int)xNext, (int)yNext), a3 = texel(t, (int)x, (int)yNext);
    const Col c0(a0.r, a0.g, a0.b), c1(a1.r, a1.g, a1.b), c2(a2.r, a2.g, a2.b), c3(a3.r, a3.g, a3.b);
    const Col temp0 = beta * c3 + (1 - beta) * c0, temp1 = beta * c2 + (1 - beta) * c1;
    return l.L * (alpha * temp1 + (1 - alpha) * temp0);
}

// EnvironmentLight::Le: AmbientLight ambientlight.h:59-61, DistantLight distantlight.h:62-65, HDRILight
YRT_D Col env_Le(const SceneData& sc, const LightRec& l, V3 wo) {
    if (l.type == LIGHT_AMBIENT) return l.L;
    if (l.type == LIGHT_DISTANT) return dot(-wo, l.v0) >= l.b ? l.L : Col(0.f);
    if (l.type == LIGHT_HDRI) return hdri_Le(sc, l, wo);
    return Col(0.f);
}

// Light::sample for the non-precomputed lights: ambientlight.h:67-80, trianglelight.h:92-100, pointlight.h:52-59,
// spotlight.h:63-75, directionallight.h:52-54, distantlight.h:72-76. tMax is produced for completeness; the
// integrator overrides it (pathtraceintegrator.cpp:151-157).
YRT_D Col light_sample(const LightRec& l, const DG& dg, LightSampleD& ls, float sx, float sy) {
    switch (l.type) {
    case LIGHT_AMBIENT: { const Sample3 s = cosine_sample_hemisphere(sx, sy, dg.Ns); ls.wi = s.v; ls.pdf = s.pdf; return l.L; }
    case LIGHT_TRIANGLE: {
        const V3 d = uniform_sample_triangle(sx, sy, l.v0, l.v1, l.v2) - dg.P;
        ls.tMax = length(d);
        const float dDotNg = dot(d, l.Ng);
        if (dDotNg >= 0) { ls.pdf = 0.f; return Col(0.f); }
        ls.wi = d * rcpf(ls.tMax); ls.pdf = 2.0f * ls.tMax * ls.tMax * ls.tMax * rcpf(fabsf(dDotNg));
        return l.L;
    }
    case LIGHT_POINT: {
        const V3 d = l.v0 - dg.P; const float dist = length(d);
        ls.wi = d / dist; ls.pdf = dist * dist; ls.tMax = dist; return l.L;
    }
    case LIGHT_SPOT: {
        const V3 d = l.v0 - dg.P; const float dist = length(d);
        ls.wi = d * rcpf(dist); ls.pdf = dist * dist; ls.tMax = dist;
        const float cosAngle = dot(ls.wi, l.v1);
        if (l.a != l.b) return l.L * rclamp((cosAngle - l.b) * rcpf(l.a - l.b));
        else if (cosAngle > l.a) return l.L;
        return Col(0.f);
    }
    case LIGHT_DIRECTIONAL: ls.wi = l.v0; ls.pdf = 1.0f; ls.tMax = INFINITY; return l.L;
    case LIGHT_DISTANT: {
        // uniformSampleCone(u,v,angle,N)  shapesampler.h
        const float phi = YRT_TWO_PI * sx;
        const float cosTheta = 1.0f - sy * (1.0f - YRT_COSF(l.a));
        const float sinTheta = sqrtf(rmax(0.f, 1.f - cosTheta * cosTheta));
        float sinPhi, cosPhi; YRT_SINCOS(phi, sinPhi, cosPhi);
        ls.wi = xfmVector(frame(l.v0), V3(cosPhi * sinTheta, sinPhi * sinTheta, cosTheta));
        ls.pdf = rcpf(4.0f * YRT_PI * (YRT_SINF(0.5f * l.a) * YRT_SINF(0.5f * l.a)));
        ls.tMax = INFINITY; return l.L;
    }
    }
    ls.pdf = 0.f; return Col(0.f);
}

}  // namespace yrt
