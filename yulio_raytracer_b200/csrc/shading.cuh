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
        c = c4add(c4mul(c4add(c4mul(texel(t, x, y), uo), c4mul(texel(t, x + 1, y), ur)), vo),
                  c4mul(c4add(c4mul(texel(t, x, y + 1), uo), c4mul(texel(t, x + 1, y + 1), ur)), vr));
    } else {
        const int si = (int)(s1 * float(t.width)), ti = (int)(t1 * float(t.height));
        c = texel(t, iclamp(si, 0, t.width - 1), iclamp(ti, 0, t.height - 1));
    }
    if (t.invert) { c.r = 1.f - c.r; c.g = 1.f - c.g; c.b = 1.f - c.b; c.a = 1.f - c.a; }
    return c;
}

// ---- differential geometry -----------------------------------------------------------------
struct DG { V3 P, Ng, Ns; float s, t, error; int material, areaLight, illumMask, shadowMask; };

// TriangleMeshFull::postIntersect      shapes/trianglemesh_full.cpp:207-275
// TriangleMeshWithNormals::postIntersect shapes/trianglemesh_normals.cpp:140-162
// Triangle::postIntersect              shapes/triangle.h:84-93
// (Tx/Ty are not produced: no material on the hot path reads them — Obj bump maps are rejected at commit.)
YRT_D void post_intersect(const SceneData& sc, V3 org, V3 dir, float t, float u, float v, int geomID, int primID, V3 rayNg, DG& dg) {
    const GeomRec g = sc.geoms[geomID];
    dg.material = g.material; dg.areaLight = g.areaLight; dg.illumMask = g.illumMask; dg.shadowMask = g.shadowMask;
    dg.P = org + t * dir;
    if (g.type == MESH_TRIANGLE) { dg.Ng = g.triNg; dg.Ns = g.triNg; dg.s = u; dg.t = v; }
    else {
        dg.Ng = normalize(rayNg);
        const int4 tri = sc.indices[g.idxBase + primID];
        const float w = 1.0f - u - v;
        if (g.uvBase != YRT_NO_ATTR) {
            const float2 st0 = sc.uvs[g.uvBase + tri.x], st1 = sc.uvs[g.uvBase + tri.y], st2 = sc.uvs[g.uvBase + tri.z];
            dg.s = st0.x * w + st1.x * u + st2.x * v; dg.t = st0.y * w + st1.y * u + st2.y * v;
        } else { dg.s = u; dg.t = v; }
        if (g.nrmBase != YRT_NO_ATTR) {
            const float4 a = sc.normals[g.nrmBase + tri.x], b = sc.normals[g.nrmBase + tri.y], c = sc.normals[g.nrmBase + tri.z];
            V3 Ns = w * V3(a.x, a.y, a.z) + u * V3(b.x, b.y, b.z) + v * V3(c.x, c.y, c.z);
            const float len2 = dot(Ns, Ns);
            Ns = len2 > 0 ? Ns * rsqrtf_exact(len2) : dg.Ng;
            if (dot(Ns, dg.Ng) < 0) Ns = -Ns;
            dg.Ns = Ns;
        } else dg.Ns = dg.Ng;
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
                LOBE_THIN_DIEL_TRANS, LOBE_CONST_DIEL_TRANS, LOBE_MICROFACET_UBER };
struct Lobe { int kind; uint32_t type; Col c; float a, b; };   // meaning of c,a,b depends on kind
// The lobes of one hit and the per-lobe candidates of CompositedBRDF::sample are indexed per lane at run time. As local arrays
// they cost uncoalesced local-memory traffic (ncu r1: 2.7 of 32 bytes used per sector); they live in shared memory instead, word
// w of slot i of thread t at s[(i * WORDS + w) * stride + t]: bank = t, conflict-free for any per-lane i.
#define YRT_MAX_LOBES 3
#define YRT_LOBE_WORDS 7
#define YRT_CAND_WORDS 9
struct Lobes {
    float* s; float* cand; int stride; int n;
    YRT_D Lobe get(int i) const {
        const float* p = s + i * YRT_LOBE_WORDS * stride; Lobe l;
        l.kind = __float_as_int(p[0]); l.type = __float_as_uint(p[stride]); l.c = Col(p[2 * stride], p[3 * stride], p[4 * stride]); l.a = p[5 * stride]; l.b = p[6 * stride];
        return l;
    }
    YRT_D uint32_t type(int i) const { return __float_as_uint(s[(i * YRT_LOBE_WORDS + 1) * stride]); }
};
YRT_D void add_lobe(Lobes& L, int kind, uint32_t type, Col c, float a = 0.f, float b = 0.f) {
    if (L.n < YRT_MAX_LOBES) {
        float* p = L.s + L.n * YRT_LOBE_WORDS * L.stride; L.n++;
        p[0] = __int_as_float(kind); p[L.stride] = __uint_as_float(type); p[2 * L.stride] = c.x; p[3 * L.stride] = c.y; p[4 * L.stride] = c.z;
        p[5 * L.stride] = a; p[6 * L.stride] = b;
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
    Sample3 s; s.v = xfmVector(frame(N), V3(YRT_COSF(phi) * sinTheta, YRT_SINF(phi) * sinTheta, cosTheta)); s.pdf = cosTheta * YRT_ONE_OVER_PI;
    return s;
}
YRT_D Sample3 power_cosine_sample_hemisphere(float u, float v, V3 N, float e) {
    const float phi = YRT_TWO_PI * u;
    const float cosTheta = YRT_POWF(v, rcpf(e + 1));
    const float sinTheta = sqrtf(rmax(0.f, 1.f - cosTheta * cosTheta));
    Sample3 s; s.v = xfmVector(frame(N), V3(YRT_COSF(phi) * sinTheta, YRT_SINF(phi) * sinTheta, cosTheta));
    s.pdf = (e + 1.0f) * YRT_POWF(cosTheta, e) * YRT_ONE_OVER_TWO_PI;
    return s;
}

// Lambertian::eval lambertian.h:35-37; Specular::eval specular.h:34-38; Microfacet::eval microfacet.h:43-58
// (+ PowerCosineDistribution::eval power_cosine_distribution.h:35-39, FresnelDielectric::eval fresnel.h:80-82);
// all specular lobes evaluate to zero.
YRT_D Col lobe_eval(const Lobe& l, V3 wo, const DG& dg, V3 wi) {
    switch (l.kind) {
    case LOBE_LAMBERTIAN: return l.c * YRT_ONE_OVER_PI * rclamp(dot(wi, dg.Ns));
    case LOBE_SPECULAR: {
        const V3 r = reflect_v(wo, dg.Ns);
        if (dot(r, wi) < 0) return Col(0.f);
        return l.c * (l.a + 2) * (1.0f / (2.0f * YRT_PI)) * YRT_POWF(dot(r, wi), l.a) * rclamp(dot(wi, dg.Ns));
    }
    case LOBE_REFLECTION: return l.c;
    case LOBE_MICROFACET_UBER: {
        if (dot(wi, dg.Ng) <= 0) return Col(0.f);
        const float cosThetaO = dot(wo, dg.Ns), cosThetaI = dot(wi, dg.Ns);
        if (cosThetaI <= 0.0f || cosThetaO <= 0.0f) return Col(0.f);
        const V3 wh = normalize(wi + wo);
        const float cosThetaH = dot(wh, dg.Ns), cosTheta = dot(wi, wh);
        const Col F = Col(fresnel_diel(cosTheta, l.a));                               // l.a = etai/etat
        const float D = ((l.b + 2) * YRT_ONE_OVER_TWO_PI) * YRT_POWF(fabsf(dot(wh, dg.Ns)), l.b);   // l.b = n
        const float G = rmin(rmin(1.0f, 2.0f * cosThetaH * cosThetaO * rcpf(cosTheta)), 2.0f * cosThetaH * cosThetaI * rcpf(cosTheta));
        return l.c * D * G * F * rcpf(4.0f * cosThetaO);
    }
    default: return Col(0.f);
    }
}

// The per-lobe sample() methods: lambertian.h:39-41, specular.h:40-42, transmission.h:38-40, reflection.h:40-43,
// dielectric.h:39-45 (DielectricReflection), :80-87 (DielectricTransmission), :122-132 (ThinDielectricTransmission),
// :185-189 (ConstDielectricTransmission), microfacet.h:60-67.
YRT_D Col lobe_sample(const Lobe& l, V3 wo, const DG& dg, Sample3& wi, float sx, float sy) {
    switch (l.kind) {
    case LOBE_LAMBERTIAN: wi = cosine_sample_hemisphere(sx, sy, dg.Ns); return lobe_eval(l, wo, dg, wi.v);
    case LOBE_SPECULAR: wi = power_cosine_sample_hemisphere(sx, sy, reflect_v(wo, dg.Ns), l.a); return lobe_eval(l, wo, dg, wi.v);
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
        // PowerCosineDistribution::sample  power_cosine_distribution.h:43-51
        const float phi = YRT_TWO_PI * sx;
        const float cosPhi = YRT_COSF(phi), sinPhi = YRT_SINF(phi);
        const float cosTheta = YRT_POWF(sy, rcpf(l.b + 1));
        const float sinTheta = sqrtf(rmax(0.f, 1.f - cosTheta * cosTheta));
        const V3 wh = xfmVector(frame(dg.Ns), V3(cosPhi * sinTheta, sinPhi * sinTheta, cosTheta));
        const float whPdf = ((l.b + 1) * YRT_ONE_OVER_TWO_PI) * YRT_POWF(cosTheta, l.b);
        wi.v = reflect_v(wo, wh); wi.pdf = whPdf * rcpf(4.0f * fabsf(dot(wo, wh)));
        if (dot(wi.v, dg.Ns) <= 0.0f) return Col(0.f);
        return lobe_eval(l, wo, dg, wi.v);
    }
    }
    wi.v = V3(0.f); wi.pdf = 0.f;
    return Col(0.f);
}

// CompositedBRDF::eval  brdfs/compositedbrdf.h:74-80
YRT_D Col lobes_eval(const Lobes& L, V3 wo, const DG& dg, V3 wi, uint32_t typeMask) {
    Col c(0.f);
#pragma unroll 1
    for (int i = 0; i < L.n; i++) { const Lobe l = L.get(i); if (l.type & typeMask) c += lobe_eval(l, wo, dg, wi); }
    return c;
}

// CompositedBRDF::sample  brdfs/compositedbrdf.h:119-181
YRT_D Col lobes_sample(const Lobes& L, V3 wo, const DG& dg, Sample3& wiOut, uint32_t& typeOut, float sx, float sy, float ss, uint32_t typeMask) {
    float sum = 0.0f; int num = 0;
    const int st = L.stride;
#pragma unroll 1
    for (int i = 0; i < L.n; i++) {
        const Lobe l = L.get(i);
        if (!(l.type & typeMask)) continue;
        Sample3 wi; const Col c = lobe_sample(l, wo, dg, wi, sx, sy);
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
YRT_D void material_shade(const SceneData& sc, const MaterialRec& m, const DG& dg, Col mediumT, float mediumEta, Lobes& L) {
    L.n = 0;
    switch (m.type) {
    case MAT_MATTE: add_lobe(L, LOBE_LAMBERTIAN, BR_DIFFUSE_REFLECTION, m.c0); break;
    case MAT_MIRROR: add_lobe(L, LOBE_REFLECTION, BR_SPECULAR_REFLECTION, m.c0); break;
    case MAT_MATTE_TEXTURED:
        if (m.tex[0] >= 0) { const Col4 c = tex_get(sc.textures[m.tex[0]], m.dsx * dg.s + m.s0x, m.dsy * dg.t + m.s0y); add_lobe(L, LOBE_LAMBERTIAN, BR_DIFFUSE_REFLECTION, Col(c.r, c.g, c.b)); }
        break;
    case MAT_OBJ: {
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
        const float eta = m.f[0], roughness = m.f[1], reflectivity = m.f[2], rcpRoughness = m.f[3];
        if (reflectivity > 0.f) add_lobe(L, LOBE_DIEL_REFL, BR_SPECULAR_REFLECTION, Col(0.f), 1.f * rcpf(eta), alpha * reflectivity);
        else if (roughness == 0.f) add_lobe(L, LOBE_DIEL_REFL, BR_SPECULAR_REFLECTION, Col(0.f), 1.f * rcpf(eta), alpha);
        else add_lobe(L, LOBE_MICROFACET_UBER, BR_GLOSSY_REFLECTION, Col(alpha), 1.f * rcpf(eta), rcpRoughness);
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
        ls.wi = xfmVector(frame(l.v0), V3(YRT_COSF(phi) * sinTheta, YRT_SINF(phi) * sinTheta, cosTheta));
        ls.pdf = rcpf(4.0f * YRT_PI * (YRT_SINF(0.5f * l.a) * YRT_SINF(0.5f * l.a)));
        ls.tMax = INFINITY; return l.L;
    }
    }
    ls.pdf = 0.f; return Col(0.f);
}

}  // namespace yrt
