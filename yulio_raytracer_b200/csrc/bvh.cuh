// bvh.cuh — compressed 8-wide BVH node layout + the ray/box and ray/triangle tests of device_cuda.
//
// Replaces the reference's calls into the (un-vendored, binary-only) Intel Embree 2.15:
//   rtcIntersect  devices/device_singleray/integrators/pathtraceintegrator.cpp:72
//   rtcOccluded   devices/device_singleray/integrators/pathtraceintegrator.cpp:160
// Node format follows Ylitie/Karras/Laine, "Efficient Incoherent Ray Traversal on GPUs Through
// Compressed Wide BVHs" (HPG 2017): 80 bytes = origin (12) + 3 exponents + imask (4) + child base,
// triangle base (8) + leaf-triangle mask (4) + 4 spare + 6 x 8 quantised planes (48).  Triangles are 48 bytes
// (3 x float4) with geomID / primID / flags in the w lanes.
// Unlike the paper's per-child meta bytes (a shifted unary count per hit child: ~8 ALU instructions per child, 66 per node in the r2
// SASS, on the pipe that bounds the kernel), a node carries ONE mask: bit 8k + s = "slot s is a leaf child with more than k triangles".
// The node test collects one bit per hit slot; inner hits are put into traversal order by a 2 KB table lookup, leaf hits become
// triangle bits with one multiply and one AND: (hits * 0x010101) & triMask. A pending triangle set is (node index, bits); the
// triangle of bit b is triBase + popc(triMask & ((1 << b) - 1)), fetched from the node's second 16 bytes when the triangle is tested
// (1.6 triangle tests against 9.5 node tests per ray on BASELINE config 4).
//
// Arithmetic contract of the triangle test: "YRT-PLUECKER-1" (stated in oracle/embree2_shim.cpp,
// which holds the CPU twin used by the oracle) — bit-exact (t,u,v,geomID,primID).
#pragma once
#include "common.cuh"

namespace yrt {

struct __align__(16) Node8 {
    float px, py, pz;
    uint8_t ex, ey, ez, imask;
    uint32_t childBase, triBase;
    uint32_t triMask, reserved;
    uint8_t qlox[8], qloy[8], qloz[8], qhix[8], qhiy[8], qhiz[8];
};
static_assert(sizeof(Node8) == 80, "Node8 must be 80 bytes");

#define YRT_TRI_FLAG_CULL 1u
#define YRT_TRI_FLAG_MOTION 2u       // the triangle's vertices move linearly over the shutter interval: p(time) = p + time * d (SceneData::triMotion)
#define YRT_STACK_SIZE 96
#ifndef YRT_BOX_PAD
#define YRT_BOX_PAD 9.5367431640625e-07f     // 2^-20
#endif

struct HitRec { float t, u, v; int geomID, primID; V3 Ng; uint32_t tri; };   // tri = leaf-order triangle index

#if defined(__CUDACC__)

// edge(s,e).D with explicit fused multiply-adds (mirrors edgeFn in oracle/embree2_shim.cpp)
YRT_D float edge_fn(V3 s, V3 e, V3 D) {
    const float cx = __fmaf_rn(s.y, e.z, -__fmul_rn(s.z, e.y));
    const float cy = __fmaf_rn(s.z, e.x, -__fmul_rn(s.x, e.z));
    const float cz = __fmaf_rn(s.x, e.y, -__fmul_rn(s.y, e.x));
    return __fmaf_rn(cz, D.z, __fmaf_rn(cy, D.y, __fmul_rn(cx, D.x)));
}

// Returns true and fills (t,u,v,Ng) if the supporting-plane hit lies inside the triangle.
// The caller applies the (tnear, tbest, id tie-break) acceptance rule and the cull filter.
YRT_D bool tri_test(V3 O, V3 D, V3 p0, V3 p1, V3 p2, float& t, float& u, float& v, V3& Ng, float& den) {
    const V3 v0 = p0 - O, v1 = p1 - O, v2 = p2 - O;
    const V3 e0 = v2 - v0, e1 = v0 - v1, e2 = v1 - v2;
    const float U = edge_fn(v2 + v0, e0, D);
    const float V = edge_fn(v0 + v1, e1, D);
    const float W = edge_fn(v1 + v2, e2, D);
    const float mn = fminf(fminf(U, V), W), mx = fmaxf(fmaxf(U, V), W);
    if (!(mn >= 0.f || mx <= 0.f)) return false;
    const float UVW = (U + V) + W;
    if (UVW == 0.f) return false;
    const V3 a = p0 - p1, b = p2 - p0;
    Ng = cross(a, b);                                   // separate mul/sub: the cull filter's arithmetic (trianglemesh_full.cpp:112-116)
    den = dot(Ng, D);
    if (den == 0.f) return false;
    const float T = dot(v0, Ng);
    t = T / den; u = U / UVW; v = V / UVW;
    return true;
}

struct RayPre {                  // per-ray constants of the box test
    V3 O, D, idir;
    uint32_t octinv;             // bit 4/2/1 set where dir.x/y/z >= 0
};

YRT_D RayPre ray_prepare(V3 O, V3 D) {
    RayPre r; r.O = O; r.D = D;
    const float tiny = 1e-18f;
    const float dx = fabsf(D.x) < tiny ? copysignf(tiny, D.x) : D.x;
    const float dy = fabsf(D.y) < tiny ? copysignf(tiny, D.y) : D.y;
    const float dz = fabsf(D.z) < tiny ? copysignf(tiny, D.z) : D.z;
    r.idir = V3(__frcp_rn(dx), __frcp_rn(dy), __frcp_rn(dz));   // correctly rounded reciprocal == 1.0f / x
    r.octinv = (D.x >= 0.f ? 4u : 0u) | (D.y >= 0.f ? 2u : 0u) | (D.z >= 0.f ? 1u : 0u);
    return r;
}

// Tried and dropped (round 1, C3 stand-in): folding the 2^23 bias into the FMA constant (fma(2^23 + q, s, o - 2^23 s)) saves the
// 48 FADDs per node but rounds the constant at half a quantisation step; padding for it (+0.51 step) cost +16 % node visits, and
// even the 2^15-step-bias variant (byte in mantissa bits 8..15, +1/256 step of padding) cost +7 % — bounce rays start exactly on
// box planes of neighbouring axis-aligned geometry, so any extra dilation flips many far-plane-vs-tnear decisions. Net slower.
static __constant__ uint32_t yrt_c_magic = 0x4B000000u;   // a constant-bank operand: ptxas cannot fold it back into an immediate

// exact u8 -> float without I2F: PRMT builds the float 2^23 + byte, one FADD removes the bias. `magic` is 0x4B000000 held in a register
// (node_test reads it from the constant bank): PRMT takes one immediate, and with the constant as the immediate ptxas kept
// re-materialising the four byte selectors into registers — ~45 extra MOV/IMAD per node test in the r1 SASS.
YRT_D float byte_to_float(uint32_t packed, uint32_t magic, int i) {
    uint32_t r;
    switch (i) {       // i is a compile-time constant after unrolling: the selector is an immediate
    case 0: asm("prmt.b32 %0, %1, %2, 0x7650;" : "=r"(r) : "r"(packed), "r"(magic)); break;
    case 1: asm("prmt.b32 %0, %1, %2, 0x7651;" : "=r"(r) : "r"(packed), "r"(magic)); break;
    case 2: asm("prmt.b32 %0, %1, %2, 0x7652;" : "=r"(r) : "r"(packed), "r"(magic)); break;
    default: asm("prmt.b32 %0, %1, %2, 0x7653;" : "=r"(r) : "r"(packed), "r"(magic)); break;
    }
    return __uint_as_float(r) - 8388608.0f;
}

// Packed FP32 (sm_100: FADD2 / FFMA2 work on two floats held in an aligned register pair, each half rounded exactly like the scalar
// instruction): the near and the far plane of one axis share an instruction — {near byte, far byte} in one register pair, the axis
// scale as a scalar (broadcast) operand, {near offset, far offset} as the addend pair — so the bias removal and the six plane
// evaluations of a child take 3 + 3 instructions instead of 6 + 6. The results are bit-identical, only the instruction count changes.
#ifndef YRT_NODE_F32X2
#define YRT_NODE_F32X2 1
#endif
#ifndef YRT_NODE_SIGNBIT
#define YRT_NODE_SIGNBIT 1
#endif
YRT_D uint64_t f2_pack(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
YRT_D uint64_t f2_packu(uint32_t lo, uint32_t hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi)); return r; }
YRT_D void f2_unpack(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
YRT_D uint64_t f2_add(uint64_t a, uint64_t b) { uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
YRT_D uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
// Measured on B200 (tools/mb/pipes.cu): PRMT, SHF and FMNMX share the alu pipe at 2 warp instructions / clock / SM, IDP.4A runs at the same rate
// on another pipe (PRMT + IDP.4A together: 3.9 / clock / SM). The node test is bound by the alu pipe, so the planes selected by
// YRT_NODE_DP4A_MASK (bit 0..5 = near x, far x, near y, far y, near z, far z) are extracted with a dot product against a one-hot
// byte vector — dp4a(word, 1 << 8i, 2^23 as bits) is the same 32-bit value the PRMT builds.
#ifndef YRT_NODE_DP4A_MASK
#define YRT_NODE_DP4A_MASK 0x15
#endif
template <bool DP4A>
YRT_D uint32_t byte_biased(uint32_t packed, uint32_t magic, int i) {       // the float 2^23 + byte i, as bits
    uint32_t r;
    if (DP4A) { asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(packed), "r"(1u << (8 * i)), "r"(magic)); return r; }
    switch (i) {
    case 0: asm("prmt.b32 %0, %1, %2, 0x7650;" : "=r"(r) : "r"(packed), "r"(magic)); break;
    case 1: asm("prmt.b32 %0, %1, %2, 0x7651;" : "=r"(r) : "r"(packed), "r"(magic)); break;
    case 2: asm("prmt.b32 %0, %1, %2, 0x7652;" : "=r"(r) : "r"(packed), "r"(magic)); break;
    default: asm("prmt.b32 %0, %1, %2, 0x7653;" : "=r"(r) : "r"(packed), "r"(magic)); break;
    }
    return r;
}

// Traversal order: the child in slot s is visited at position s ^ octinv (highest first). perm8 moves bit s of an 8-bit slot mask to
// bit s ^ o; the streaming kernels look it up in a 2 KB shared-memory table (perm_lut_fill), the one-ray-per-thread path computes it.
#define YRT_PERM_LUT_BYTES 2048
YRT_D uint32_t perm8(uint32_t x, uint32_t o) {
    if (o & 1u) x = ((x & 0x55u) << 1) | ((x >> 1) & 0x55u);
    if (o & 2u) x = ((x & 0x33u) << 2) | ((x >> 2) & 0x33u);
    if (o & 4u) x = ((x & 0x0fu) << 4) | (x >> 4);
    return x;
}
YRT_D void perm_lut_fill(uint8_t* lut) {            // lut[o * 256 + x] = perm8(x, o); every thread of the CTA calls it, then a barrier
    for (uint32_t e = threadIdx.x; e < YRT_PERM_LUT_BYTES; e += blockDim.x) lut[e] = (uint8_t)perm8(e & 0xffu, e >> 8);
}

// Intersects the 8 quantised child boxes of one node; returns the hit mask: bits 24..31 inner children at position 24 + (slot ^ octinv),
// bits 0..23 leaf triangles in the node's triMask numbering (bit 8k + slot). permLut: perm_lut_fill's table, or nullptr.
YRT_D uint32_t node_test(const uint4 n0, const uint4 n1, const uint4 n2, const uint4 n3, const uint4 n4,
                         const RayPre& r, float tnear, float tbest, const uint8_t* permLut) {
    const V3 p(__uint_as_float(n0.x), __uint_as_float(n0.y), __uint_as_float(n0.z));
    const uint32_t e = n0.w;
    const float sx = __uint_as_float((e & 0xffu) << 23) * r.idir.x;
    const float sy = __uint_as_float(((e >> 8) & 0xffu) << 23) * r.idir.y;
    const float sz = __uint_as_float(((e >> 16) & 0xffu) << 23) * r.idir.z;
    const float ox = (p.x - r.O.x) * r.idir.x, oy = (p.y - r.O.y) * r.idir.y, oz = (p.z - r.O.z) * r.idir.z;
    const float padx = fabsf(ox) * YRT_BOX_PAD, pady = fabsf(oy) * YRT_BOX_PAD, padz = fabsf(oz) * YRT_BOX_PAD;
    const float oxn = ox - padx, oxf = ox + padx, oyn = oy - pady, oyf = oy + pady, ozn = oz - padz, ozf = oz + padz;
    const bool nx = r.idir.x < 0.f, ny = r.idir.y < 0.f, nz = r.idir.z < 0.f;
    const float tfarPadded = tbest * (1.0f + YRT_BOX_PAD);
    const uint32_t magic = yrt_c_magic;
    uint32_t hit8 = 0;
#if YRT_NODE_SIGNBIT
    uint32_t miss8 = 0;           // slot 7 first: after eight shifts slot s sits at bit s
#endif
#if YRT_NODE_F32X2
    const uint64_t unbias = f2_pack(-8388608.0f, -8388608.0f);
    const uint64_t sX2 = f2_pack(sx, sx), sY2 = f2_pack(sy, sy), sZ2 = f2_pack(sz, sz);
    const uint64_t oX2 = f2_pack(oxn, oxf), oY2 = f2_pack(oyn, oyf), oZ2 = f2_pack(ozn, ozf);
#endif
#pragma unroll
    for (int hh = 0; hh < 2; hh++) {
        const int half = YRT_NODE_SIGNBIT ? 1 - hh : hh;
        const uint32_t qlox = half ? n2.y : n2.x, qloy = half ? n2.w : n2.z;
        const uint32_t qloz = half ? n3.y : n3.x, qhix = half ? n3.w : n3.z;
        const uint32_t qhiy = half ? n4.y : n4.x, qhiz = half ? n4.w : n4.z;
        const uint32_t nearx = nx ? qhix : qlox, farx = nx ? qlox : qhix;
        const uint32_t neary = ny ? qhiy : qloy, fary = ny ? qloy : qhiy;
        const uint32_t nearz = nz ? qhiz : qloz, farz = nz ? qloz : qhiz;
#pragma unroll
        for (int ii = 0; ii < 4; ii++) {
            const int i = YRT_NODE_SIGNBIT ? 3 - ii : ii;
            float tnx, tny, tnz, tfx, tfy, tfz;
#if YRT_NODE_F32X2
#define YRT_BB(plane, word) byte_biased<((YRT_NODE_DP4A_MASK >> plane) & 1) != 0>(word, magic, i)
            f2_unpack(f2_fma(f2_add(f2_packu(YRT_BB(0, nearx), YRT_BB(1, farx)), unbias), sX2, oX2), tnx, tfx);
            f2_unpack(f2_fma(f2_add(f2_packu(YRT_BB(2, neary), YRT_BB(3, fary)), unbias), sY2, oY2), tny, tfy);
            f2_unpack(f2_fma(f2_add(f2_packu(YRT_BB(4, nearz), YRT_BB(5, farz)), unbias), sZ2, oZ2), tnz, tfz);
#undef YRT_BB
#else
            tnx = __fmaf_rn(byte_to_float(nearx, magic, i), sx, oxn);
            tny = __fmaf_rn(byte_to_float(neary, magic, i), sy, oyn);
            tnz = __fmaf_rn(byte_to_float(nearz, magic, i), sz, ozn);
            tfx = __fmaf_rn(byte_to_float(farx, magic, i), sx, oxf);
            tfy = __fmaf_rn(byte_to_float(fary, magic, i), sy, oyf);
            tfz = __fmaf_rn(byte_to_float(farz, magic, i), sz, ozf);
#endif
            const float tmin = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, tnear));
            const float tmax = fminf(fminf(tfx, tfy), fminf(tfz, tfarPadded));
#if YRT_NODE_SIGNBIT
            // the sign of tmax - tmin (an FADD on the fma pipe) shifted into the mask (one SHF) instead of FSETP + SEL + ADD on the alu
            // pipe, the pipe that bounds this kernel. It differs from tmin <= tmax only where the box must not be entered anyway or
            // may be entered needlessly: tmax = -0 with tmin = +0 (everything inside lies at t <= 0 <= tnear) and inf - inf = NaN.
            miss8 = __funnelshift_l(__float_as_uint(tmax - tmin), miss8, 1);
#else
            if (tmin <= tmax) hit8 |= 1u << (4 * half + i);
#endif
        }
    }
#if YRT_NODE_SIGNBIT
    hit8 = ~miss8 & 0xffu;
#endif
    // an empty slot (inverted box) that a degenerate node lets through is neither in imask nor in triMask
    const uint32_t imask = e >> 24;
    const uint32_t inner = hit8 & imask;
    const uint32_t ordered = permLut ? (uint32_t)permLut[(r.octinv << 8) + inner] : perm8(inner, r.octinv);
    return (ordered << 24) | (((hit8 & ~imask) * 0x010101u) & n1.z);
}

// leaf-order index of the triangle behind bit `bit` of a pending set of node `nodeIdx` (see the node format above)
YRT_D uint32_t tri_index(const uint4 nodeWord1, uint32_t bit) { return nodeWord1.y + (uint32_t)__popc(nodeWord1.z & ((1u << bit) - 1u)); }

struct TraceCounters { uint32_t nodes, tris; uint32_t overflow; };   // overflow: a traversal stack ran out of its YRT_STACK_SIZE entries

// ---- cache policy of the traversal kernels ----------------------------------------------------------------------------------
// A bounce moves ~2 GB of ray / hit / queue records through the 126 MB L2 exactly once, while every ray re-reads the same few tens of
// MB of nodes and triangles. With default caching the stream evicts the BVH (ncu r1, C2: L2 hit 44 %). So: the streams use
// ld.global.cs / st.global.cs (evict-first, kernels.cu: ClosestIO / ShadowIO), and node / triangle fetches carry an L2 evict_last
// policy (createpolicy.fractional.L2::evict_last + ld.global.nc.L2::cache_hint) so that they are the last lines L2 gives up.
#ifndef YRT_BVH_EVICT_LAST
#define YRT_BVH_EVICT_LAST 0      // measured r2 (profiles/README.md): no effect on C2-C4, and the policy's register pair costs 2.6 % on config 5 -> off
#endif
#ifndef YRT_IDLE_EVERY
#define YRT_IDLE_EVERY 1          // look for idle slots every n-th macro step
#endif
#ifndef YRT_STREAM_HINTS
#define YRT_STREAM_HINTS 1
#endif
YRT_D uint64_t bvh_policy() {
#if YRT_BVH_EVICT_LAST
    uint64_t p; asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p;
#else
    return 0ull;
#endif
}
YRT_D uint4 bvh_ld(const uint4* p, uint64_t pol) {
#if YRT_BVH_EVICT_LAST
    uint4 v; asm("ld.global.nc.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol));
    return v;
#else
    (void)pol; return __ldg(p);
#endif
}
YRT_D float4 bvh_ld(const float4* p, uint64_t pol) {
    const uint4 v = bvh_ld((const uint4*)p, pol);
    return make_float4(__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z), __uint_as_float(v.w));
}

// While-while traversal of the compressed BVH8.  ANY: rtcOccluded semantics (first accepted hit ends).
// Closest-hit acceptance is order independent: (t, geomID, primID) lexicographic minimum.
template <bool ANY, bool COUNT>
YRT_D bool trace_ray(const uint4* __restrict__ nodes, const float4* __restrict__ tris, uint32_t numNodes,
                     V3 O, V3 D, float tnear, float tfar, HitRec& hit, TraceCounters* cnt) {
    hit.geomID = -1; hit.primID = -1; hit.t = tfar; hit.tri = 0xffffffffu;
    if (numNodes == 0 || !(tnear <= tfar)) return false;   // NaN tfar: no hit
    const RayPre r = ray_prepare(O, D);
    uint2 stack[YRT_STACK_SIZE];
    int sp = 0;
    float tbest = tfar;
    bool have = false;
    // root = "child 7^octinv of a virtual parent with imask 0": popc(0) = 0 -> node index 0
    uint2 G = make_uint2(0u, 0x80000000u);
    uint2 T = make_uint2(0u, 0u);
    while (true) {
        {   // invariant: G has at least one pending inner child bit here
            const uint32_t bit = 31u - __clz(G.y);
            G.y &= ~(1u << bit);
            const uint32_t slot = (bit - 24u) ^ r.octinv;
            const uint32_t nodeIdx = G.x + __popc((G.y & 0xffu) & ~(0xffffffffu << slot));
            if (G.y & 0xff000000u) {
                if (sp < YRT_STACK_SIZE) stack[sp++] = G; else cnt->overflow = 1u;     // reported by the caller: never a silent drop
            }
            const uint4* np = nodes + 5ull * nodeIdx;
            const uint4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
            if (COUNT) cnt->nodes++;
            const uint32_t hm = node_test(n0, n1, n2, n3, n4, r, tnear, tbest, nullptr);
            G = make_uint2(n1.x, (hm & 0xff000000u) | (n0.w >> 24));
            T = make_uint2(nodeIdx, hm & 0x00ffffffu);
        }

        while (T.y) {
            const uint32_t bit = 31u - __clz(T.y);
            T.y &= ~(1u << bit);
            const uint32_t triIdx = tri_index(__ldg(nodes + 5ull * T.x + 1), bit);
            const float4* tp = tris + 3ull * triIdx;
            const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
            if (COUNT) cnt->tris++;
            float t, u, v, den; V3 Ng;
            if (!tri_test(O, D, V3(a.x, a.y, a.z), V3(b.x, b.y, b.z), V3(c.x, c.y, c.z), t, u, v, Ng, den)) continue;
            if (!(t > tnear)) continue;
            const int g = __float_as_int(a.w), p = __float_as_int(b.w);
            bool closer = t < tbest;
            if (!ANY && !closer && have && t == tbest) closer = (g < hit.geomID) || (g == hit.geomID && p < hit.primID);
            if (!closer) continue;
            // back-face cull filter (shapes/trianglemesh_full.cpp:101-121): reject if dot(Ng, dir) <= 0
            if ((__float_as_uint(c.w) & YRT_TRI_FLAG_CULL) && den <= 0.f) continue;
            have = true; tbest = t;
            hit.t = t; hit.u = u; hit.v = v; hit.geomID = g; hit.primID = p; hit.Ng = Ng; hit.tri = triIdx;
            if (ANY) return true;
        }
        if ((G.y & 0xff000000u) == 0) {
            if (sp == 0) break;
            G = stack[--sp];
        }
    }
    return have;
}


// ------------------------------------------------------------------------------------------------------------
// trace_stream — persistent, warp-synchronous traversal of a whole ray queue (the production path).
//
// Why not one ray per thread for the life of the thread (trace_ray above): the round-1 ncu capture of that loop on
// incoherent bounce rays showed 8.4 of 32 lanes active per issued instruction (profiles/r1_closest_baseline.md):
// lanes idle (a) while the slowest ray of the warp finishes and (b) while a few lanes test triangles. Here
//   * a warp owns 32 ray slots and refills finished slots from a slice of the queue (one atomic per YRT_SLICE rays),
//   * every macro step the warp votes (two ballots) whether to run the NODE phase (each lane with a pending inner
//     node pops and tests one compressed node) or the TRIANGLE phase (each lane with pending triangles tests one),
//     postponing the minority kind of work on the per-lane stack (Aila/Laine speculative traversal, Ylitie et al.
//     triangle postponing), so both phases run with most lanes participating,
//   * the first YRT_SM_STACK stack entries live in shared memory, deeper ones spill to local memory,
//   * a finished ray keeps its result in registers (sp = -1) until the warp's next refill writes the results of all idle slots at once,
//   * the warp's slice of the queue and the slot-permutation table live in shared memory; the loop itself fits 64 registers
//     (8 CTAs of 128 threads per SM) without a spill in the node phase.
// Results are independent of the schedule: closest-hit acceptance is the (t, geomID, primID) minimum, any-hit is
// an OR over accepted hits, and the triangle arithmetic is YRT-PLUECKER-1 exactly as in trace_ray.
#define YRT_TRACE_THREADS 128
#define YRT_SM_STACK 8
#define YRT_SLICE 256u
#define YRT_REFILL_MIN 8
#define YRT_NO_TRI 0xffffffffu

// refill when >= refillMin slots idle; triangle phase when triNum*nT >= triDen*nN; prefetch (cfg prefetch=1, off by default): after a node test,
// pull every hit inner child and the first pending triangle towards L2. Meant for scenes whose BVH exceeds L2 (BASELINE config 5: a dependent chain of
// ~50 node fetches per ray). Measured r2 (profiles/README.md): slower at every size — 1e6 triangles 7182 -> 1888 Mrays/s, 1e7 2981 -> 2063,
// 1e8 (5.9 GB of BVH) 978 -> 827, C4 3272 -> 2193: the prefetch instructions cost more than the latency they hide
struct TraceTune { int refillMin; int triNum; int triDen; int prefetch; };
YRT_D void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }

// MOTION (scenes with moving meshes only): every ray carries its time (IO::time), a triangle flagged YRT_TRI_FLAG_MOTION is moved to
// p + time * d before the test — a multiply and an add per component, as the oracle's shim does (embree2_shim.cpp) — for the test, the
// cull filter and the tie-break alike. Static triangles and static scenes run the code they always ran.
template <bool ANY, bool COUNT, bool MOTION, class IO>
YRT_D void trace_stream(const uint4* __restrict__ nodes, const float4* __restrict__ tris, const float4* __restrict__ triMotion, uint32_t numNodes,
                        uint32_t n, uint32_t* __restrict__ workCounter, IO io, TraceCounters& cnt, const TraceTune tune) {
    __shared__ uint2 smStack[YRT_SM_STACK * YRT_TRACE_THREADS];
    __shared__ uint8_t smPerm[YRT_PERM_LUT_BYTES];
    uint2 lstack[YRT_STACK_SIZE - YRT_SM_STACK];
    const unsigned FULL = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u, ltmask = (1u << lane) - 1u;
    // the warp's slice of the queue, (next, end), lives in shared memory: it is touched at refills only, and the traversal loop has no register to spare
    __shared__ uint2 smSlice[YRT_TRACE_THREADS / 32];
    const unsigned wid = threadIdx.x >> 5;
    if (lane == 0) smSlice[wid] = make_uint2(0u, 0u);
    bool drained = false;                          // the queue has no unclaimed rays left and the warp's slice is used up
    // rays claimed per atomic: a full YRT_SLICE for long queues, down to one warp's worth for short ones so that a
    // late, nearly empty bounce still spreads over all SMs instead of serialising on a few warps
    const uint32_t totalWarps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t sliceRaw = (n / totalWarps) & ~31u;

    bool active = false, occluded = false;
    RayPre r; r.O = V3(0.f); r.D = V3(0.f); r.idir = V3(0.f); r.octinv = 0;
    float tnear = 0.f, tbest = 0.f, bu = 0.f, bv = 0.f, time = 0.f;      // tbest: the ray's tfar until a hit is accepted, then the hit's t
    uint32_t bestTri = YRT_NO_TRI, tag = 0;
    uint2 G = make_uint2(0u, 0u), T = make_uint2(0u, 0u);
    int sp = 0;
    const uint64_t pol = bvh_policy();
#if YRT_IDLE_EVERY > 1
    uint32_t iterNo = 0;
#endif
    perm_lut_fill(smPerm);
    __syncthreads();

    while (true) {
        // ---- refill idle slots ---------------------------------------------------------------------
#if YRT_IDLE_EVERY > 1
        const unsigned idleMask = ((iterNo++ % YRT_IDLE_EVERY) == 0u) ? __ballot_sync(FULL, !active) : 0u;      // iterNo is warp-uniform
#else
        const unsigned idleMask = __ballot_sync(FULL, !active);
#endif
        if (idleMask) {
            const uint32_t nIdle = __popc(idleMask);
            if (nIdle >= (uint32_t)tune.refillMin && !drained) {
                // results of the rays that finished since the last refill (sp < 0) are written here, refillMin or more slots at a time: writing
                // each one when its ray ends put a store (and the wait for its address) into almost every macro step, for two or three lanes
                if (!active && sp < 0) {
                    if (ANY) io.store_any(tag, occluded);
                    else io.store_hit(tag, tbest, bu, bv, bestTri, tris);     // bestTri == YRT_NO_TRI: miss
                    sp = 0;
                }
                uint2 sl = smSlice[wid];
                if (sl.x >= sl.y) {
                    const uint32_t slice = sliceRaw < 32u ? 32u : (sliceRaw > YRT_SLICE ? YRT_SLICE : sliceRaw);
                    uint32_t s = 0;
                    if (lane == 0) s = atomicAdd(workCounter, slice);
                    s = __shfl_sync(FULL, s, 0);
                    if (s >= n) drained = true;
                    else { sl.x = s; sl.y = (n - s < slice) ? n : s + slice; }
                }
                const uint32_t avail = sl.y > sl.x ? sl.y - sl.x : 0u;
                const uint32_t my = __popc(idleMask & ltmask);
                if (!active && my < avail) {
                    V3 O, D; float tfar;
                    tag = io.load(sl.x + my, O, D, tnear, tfar);
                    if (MOTION) time = io.time(tag);
                    r = ray_prepare(O, D);
                    tbest = tfar; bu = 0.f; bv = 0.f; bestTri = YRT_NO_TRI; occluded = false; sp = 0;
                    T = make_uint2(0u, 0u);
                    // root = "child 7^octinv of a virtual parent with imask 0"; an empty scene or an empty / NaN
                    // interval (pin P2) starts with no work and is retired below as a miss
                    G = make_uint2(0u, (numNodes != 0 && tnear <= tfar) ? 0x80000000u : 0u);
                    active = true;
                }
                sl.x += nIdle < avail ? nIdle : avail;
                __syncwarp();
                if (lane == 0) smSlice[wid] = sl;
                __syncwarp();
            }
            if (!__any_sync(FULL, active)) {
                if (drained) break;
                continue;
            }
        }

        // ---- vote on the phase ------------------------------------------------------------------------
        const bool nodeWork = active && (G.y & 0xff000000u) != 0u;
        const bool triWork = active && T.y != 0u;
        const int nN = __popc(__ballot_sync(FULL, nodeWork)), nT = __popc(__ballot_sync(FULL, triWork));

        if (nT != 0 && (nN == 0 || tune.triNum * nT >= tune.triDen * nN)) {
            // ---- TRIANGLE phase: one triangle per participating lane -------------------------------------
            if (triWork) {
                const uint32_t bit = 31u - __clz(T.y);
                T.y &= ~(1u << bit);
                const uint32_t triIdx = tri_index(bvh_ld(nodes + 5ull * T.x + 1, pol), bit);     // the node's (child base, triangle base, triMask, -)
                const float4* tp = tris + 3ull * triIdx;
                const float4 a = bvh_ld(tp, pol), b = bvh_ld(tp + 1, pol), c = bvh_ld(tp + 2, pol);
                if (COUNT) cnt.tris++;
                V3 p0(a.x, a.y, a.z), p1(b.x, b.y, b.z), p2(c.x, c.y, c.z);
                if (MOTION && (__float_as_uint(c.w) & YRT_TRI_FLAG_MOTION)) {
                    const float4* mp = triMotion + 3ull * triIdx;
                    const float4 d0 = __ldg(mp), d1 = __ldg(mp + 1), d2 = __ldg(mp + 2);
                    p0 = p0 + time * V3(d0.x, d0.y, d0.z); p1 = p1 + time * V3(d1.x, d1.y, d1.z); p2 = p2 + time * V3(d2.x, d2.y, d2.z);
                }
                float t, u, v, den; V3 Ng;
                if (tri_test(r.O, r.D, p0, p1, p2, t, u, v, Ng, den) && t > tnear) {
                    bool closer = t < tbest;
                    if (!ANY && !closer && bestTri != YRT_NO_TRI && t == tbest) {      // tie: (geomID, primID) ascending
                        const int g = __float_as_int(a.w), p = __float_as_int(b.w);
                        const int bg = __float_as_int(__ldg(tris + 3ull * bestTri).w), bp = __float_as_int(__ldg(tris + 3ull * bestTri + 1).w);
                        closer = (g < bg) || (g == bg && p < bp);
                    }
                    // back-face cull filter (shapes/trianglemesh_full.cpp:101-121): reject if dot(Ng, dir) <= 0
                    if (closer && !((__float_as_uint(c.w) & YRT_TRI_FLAG_CULL) && den <= 0.f)) {
                        tbest = t; bu = u; bv = v; bestTri = triIdx;
                        if (ANY) { occluded = true; G.y = 0u; T.y = 0u; sp = 0; }
                    }
                }
            }
        } else if (nN != 0) {
            // ---- NODE phase: one compressed node per participating lane --------------------------------
            if (nodeWork) {
                // the node's five loads go out first; the stack work below does not depend on them and covers part of their latency
                const uint32_t bit = 31u - __clz(G.y);
                G.y &= ~(1u << bit);
                const uint32_t slot = (bit - 24u) ^ r.octinv;
                const uint32_t nodeIdx = G.x + __popc((G.y & 0xffu) & ~(0xffffffffu << slot));
                const uint4* np = nodes + 5ull * nodeIdx;
                const uint4 n0 = bvh_ld(np, pol), n1 = bvh_ld(np + 1, pol), n2 = bvh_ld(np + 2, pol), n3 = bvh_ld(np + 3, pol), n4 = bvh_ld(np + 4, pol);
                if (T.y) {                                       // postpone the pending triangles (below the rest of this node's children)
                    if (sp < YRT_SM_STACK) smStack[sp * YRT_TRACE_THREADS + threadIdx.x] = T;
                    else if (sp < YRT_STACK_SIZE) lstack[sp - YRT_SM_STACK] = T;
                    if (sp < YRT_STACK_SIZE) sp++; else io.overflow();     // reported (wb.stats[7] -> the call fails): never a silent drop
                    T.y = 0u;
                }
                if (G.y & 0xff000000u) {
                    if (sp < YRT_SM_STACK) smStack[sp * YRT_TRACE_THREADS + threadIdx.x] = G;
                    else if (sp < YRT_STACK_SIZE) lstack[sp - YRT_SM_STACK] = G;
                    if (sp < YRT_STACK_SIZE) sp++; else io.overflow();
                }
                if (COUNT) cnt.nodes++;
                const uint32_t hm = node_test(n0, n1, n2, n3, n4, r, tnear, tbest, smPerm);
                G = make_uint2(n1.x, (hm & 0xff000000u) | (n0.w >> 24));
                T = make_uint2(nodeIdx, hm & 0x00ffffffu);
                if (tune.prefetch) {
                    uint32_t m = G.y >> 24;
                    while (m) {
                        const uint32_t b2 = 31u - __clz(m); m &= ~(1u << b2);
                        const uint32_t sl = b2 ^ r.octinv;
                        const char* cp = (const char*)(nodes + 5ull * (G.x + __popc((G.y & 0xffu) & ~(0xffffffffu << sl))));
                        prefetch_l2(cp); prefetch_l2(cp + 64);
                    }
                    if (T.y) prefetch_l2(tris + 3ull * tri_index(n1, 31u - __clz(T.y)));
                }
            }
        }

        // ---- normalise: refill (G, T) from the stack or retire the ray ------------------------------------
        if (active && (G.y & 0xff000000u) == 0u && T.y == 0u) {
            if (sp > 0) {
                sp--;
                const uint2 e = sp < YRT_SM_STACK ? smStack[sp * YRT_TRACE_THREADS + threadIdx.x] : lstack[sp - YRT_SM_STACK];
                if (e.y & 0xff000000u) G = e; else T = e;
            } else {
                sp = -1;                                             // finished: the result waits in registers for the next refill
                active = false;
            }
        }
    }
    if (sp < 0) {                                                    // the rays that finished after the last refill
        if (ANY) io.store_any(tag, occluded);
        else io.store_hit(tag, tbest, bu, bv, bestTri, tris);
    }
}

#endif  // __CUDACC__
}  // namespace yrt
