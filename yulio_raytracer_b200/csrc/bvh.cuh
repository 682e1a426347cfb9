// bvh.cuh — compressed 8-wide BVH node layout + the ray/box and ray/triangle tests of device_cuda.
//
// Replaces the reference's calls into the (un-vendored, binary-only) Intel Embree 2.15:
//   rtcIntersect  devices/device_singleray/integrators/pathtraceintegrator.cpp:72
//   rtcOccluded   devices/device_singleray/integrators/pathtraceintegrator.cpp:160
// Node format follows Ylitie/Karras/Laine, "Efficient Incoherent Ray Traversal on GPUs Through
// Compressed Wide BVHs" (HPG 2017): 80 bytes = origin (12) + 3 exponents + imask (4) + child base,
// triangle base (8) + 8 meta bytes + 6 x 8 quantised planes (48).  Triangles are 48 bytes
// (3 x float4) with geomID / primID / flags in the w lanes.
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
    uint8_t meta[8];
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

// Intersects the 8 quantised child boxes of one node; returns the CWBVH hit mask:
// bits 24..31 inner children at position 24 + (slot ^ octinv), bits 0..23 leaf triangles.
YRT_D uint32_t node_test(const uint4 n0, const uint4 n1, const uint4 n2, const uint4 n3, const uint4 n4,
                         const RayPre& r, float tnear, float tbest) {
    const V3 p(__uint_as_float(n0.x), __uint_as_float(n0.y), __uint_as_float(n0.z));
    const uint32_t e = n0.w;
    const float sx = __uint_as_float((e & 0xffu) << 23) * r.idir.x;
    const float sy = __uint_as_float(((e >> 8) & 0xffu) << 23) * r.idir.y;
    const float sz = __uint_as_float(((e >> 16) & 0xffu) << 23) * r.idir.z;
    const float ox = (p.x - r.O.x) * r.idir.x, oy = (p.y - r.O.y) * r.idir.y, oz = (p.z - r.O.z) * r.idir.z;
    const float padx = fabsf(ox) * YRT_BOX_PAD, pady = fabsf(oy) * YRT_BOX_PAD, padz = fabsf(oz) * YRT_BOX_PAD;
    const float oxn = ox - padx, oxf = ox + padx, oyn = oy - pady, oyf = oy + pady, ozn = oz - padz, ozf = oz + padz;
    const bool nx = r.idir.x < 0.f, ny = r.idir.y < 0.f, nz = r.idir.z < 0.f;
    const uint32_t octinv4 = r.octinv * 0x01010101u;
    const float tfarPadded = tbest * (1.0f + YRT_BOX_PAD);
    const uint32_t magic = yrt_c_magic;
    uint32_t hitmask = 0;
#pragma unroll
    for (int half = 0; half < 2; half++) {
        const uint32_t meta4 = half ? n1.w : n1.z;
        const uint32_t qlox = half ? n2.y : n2.x, qloy = half ? n2.w : n2.z;
        const uint32_t qloz = half ? n3.y : n3.x, qhix = half ? n3.w : n3.z;
        const uint32_t qhiy = half ? n4.y : n4.x, qhiz = half ? n4.w : n4.z;
        const uint32_t nearx = nx ? qhix : qlox, farx = nx ? qlox : qhix;
        const uint32_t neary = ny ? qhiy : qloy, fary = ny ? qloy : qhiy;
        const uint32_t nearz = nz ? qhiz : qloz, farz = nz ? qloz : qhiz;
        const uint32_t isInner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
        const uint32_t innerMask4 = (isInner4 >> 4) * 0xffu;            // 0xff per inner byte
        const uint32_t bitIndex4 = (meta4 ^ (octinv4 & innerMask4)) & 0x1f1f1f1fu;
        const uint32_t childBits4 = (meta4 >> 5) & 0x07070707u;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float tnx = __fmaf_rn(byte_to_float(nearx, magic, i), sx, oxn);
            const float tny = __fmaf_rn(byte_to_float(neary, magic, i), sy, oyn);
            const float tnz = __fmaf_rn(byte_to_float(nearz, magic, i), sz, ozn);
            const float tfx = __fmaf_rn(byte_to_float(farx, magic, i), sx, oxf);
            const float tfy = __fmaf_rn(byte_to_float(fary, magic, i), sy, oyf);
            const float tfz = __fmaf_rn(byte_to_float(farz, magic, i), sz, ozf);
            const float tmin = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, tnear));
            const float tmax = fminf(fminf(tfx, tfy), fminf(tfz, tfarPadded));
            if (tmin <= tmax) {
                const uint32_t bits = (childBits4 >> (8 * i)) & 0xffu;
                const uint32_t idx = (bitIndex4 >> (8 * i)) & 0xffu;
                hitmask |= bits << idx;
            }
        }
    }
    return hitmask;
}

struct TraceCounters { uint32_t nodes, tris; uint32_t overflow; };   // overflow: a traversal stack ran out of its YRT_STACK_SIZE entries

// ---- cache policy of the traversal kernels ----------------------------------------------------------------------------------
// A bounce moves ~2 GB of ray / hit / queue records through the 126 MB L2 exactly once, while every ray re-reads the same few tens of
// MB of nodes and triangles. With default caching the stream evicts the BVH (ncu r1, C2: L2 hit 44 %). So: the streams use
// ld.global.cs / st.global.cs (evict-first, kernels.cu: ClosestIO / ShadowIO), and node / triangle fetches carry an L2 evict_last
// policy (createpolicy.fractional.L2::evict_last + ld.global.nc.L2::cache_hint) so that they are the last lines L2 gives up.
#ifndef YRT_BVH_EVICT_LAST
#define YRT_BVH_EVICT_LAST 1
#endif
#ifndef YRT_STREAM_HINTS
#define YRT_STREAM_HINTS 1
#endif
#ifndef YRT_TRI_BATCH
#define YRT_TRI_BATCH 0
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
            const uint32_t hm = node_test(n0, n1, n2, n3, n4, r, tnear, tbest);
            G = make_uint2(n1.x, (hm & 0xff000000u) | (n0.w >> 24));
            T = make_uint2(n1.y, hm & 0x00ffffffu);
        }

        while (T.y) {
            const uint32_t bit = 31u - __clz(T.y);
            T.y &= ~(1u << bit);
            const uint32_t triIdx = T.x + bit;
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
//   * the first YRT_SM_STACK stack entries live in shared memory, deeper ones spill to local memory.
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
#if YRT_TRI_BATCH
    __shared__ uint32_t smTriList[YRT_TRACE_THREADS / 32][32];      // owner lane << 27 | leaf-order triangle index (< 2^27)
    __shared__ float4 smTriRes[YRT_TRACE_THREADS / 32][32];         // (t, u, v, triangle) of an accepted candidate, t = +inf otherwise
#endif
    uint2 lstack[YRT_STACK_SIZE - YRT_SM_STACK];
    const unsigned FULL = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u, ltmask = (1u << lane) - 1u;
    uint32_t sliceNext = 0, sliceEnd = 0; bool exhausted = false;
    // rays claimed per atomic: a full YRT_SLICE for long queues, down to one warp's worth for short ones so that a
    // late, nearly empty bounce still spreads over all SMs instead of serialising on a few warps
    const uint32_t totalWarps = (gridDim.x * blockDim.x) >> 5;
    uint32_t slice = (n / totalWarps) & ~31u;
    slice = slice < 32u ? 32u : (slice > YRT_SLICE ? YRT_SLICE : slice);

    bool active = false, occluded = false;
    RayPre r; r.O = V3(0.f); r.D = V3(0.f); r.idir = V3(0.f); r.octinv = 0;
    float tnear = 0.f, tbest = 0.f, bt = 0.f, bu = 0.f, bv = 0.f, time = 0.f;
    uint32_t bestTri = YRT_NO_TRI, tag = 0;
    uint2 G = make_uint2(0u, 0u), T = make_uint2(0u, 0u);
    int sp = 0;
    const uint64_t pol = bvh_policy();

    while (true) {
        // ---- refill idle slots ---------------------------------------------------------------------
        const unsigned idleMask = __ballot_sync(FULL, !active);
        if (idleMask) {
            const uint32_t nIdle = __popc(idleMask);
            if (nIdle >= (uint32_t)tune.refillMin && !(exhausted && sliceNext >= sliceEnd)) {
                if (sliceNext >= sliceEnd) {
                    uint32_t s = 0;
                    if (lane == 0) s = atomicAdd(workCounter, slice);
                    s = __shfl_sync(FULL, s, 0);
                    if (s >= n) exhausted = true;
                    else { sliceNext = s; sliceEnd = (n - s < slice) ? n : s + slice; }
                }
                const uint32_t avail = sliceEnd > sliceNext ? sliceEnd - sliceNext : 0u;
                const uint32_t my = __popc(idleMask & ltmask);
                if (!active && my < avail) {
                    V3 O, D; float tfar;
                    tag = io.load(sliceNext + my, O, D, tnear, tfar);
                    if (MOTION) time = io.time(tag);
                    r = ray_prepare(O, D);
                    tbest = tfar; bt = tfar; bu = 0.f; bv = 0.f; bestTri = YRT_NO_TRI; occluded = false; sp = 0;
                    T = make_uint2(0u, 0u);
                    // root = "child 7^octinv of a virtual parent with imask 0"; an empty scene or an empty / NaN
                    // interval (pin P2) starts with no work and is retired below as a miss
                    G = make_uint2(0u, (numNodes != 0 && tnear <= tfar) ? 0x80000000u : 0u);
                    active = true;
                }
                sliceNext += nIdle < avail ? nIdle : avail;
            }
            if (!__any_sync(FULL, active)) {
                if (exhausted && sliceNext >= sliceEnd) break;
                continue;
            }
        }

        // ---- vote on the phase ------------------------------------------------------------------------
        const bool nodeWork = active && (G.y & 0xff000000u) != 0u;
        const bool triWork = active && T.y != 0u;
        const int nN = __popc(__ballot_sync(FULL, nodeWork)), nT = __popc(__ballot_sync(FULL, triWork));

        if (nT != 0 && (nN == 0 || tune.triNum * nT >= tune.triDen * nN)) {
#if YRT_TRI_BATCH
            // ---- TRIANGLE phase, batched across the warp's rays --------------------------------------------
            // The pending triangles of all participating lanes go into a shared-memory work list (prefix sum over the per-lane counts);
            // the 32 lanes then test 32 list entries at a time, each fetching the owner's ray with shuffles, and write (t, u, v, triangle)
            // back; every owner finally folds its own entries in list order. A lane with three pending triangles no longer needs three
            // rounds of this phase with two thirds of the warp idle (ncu r2: the phase ran with 4-9 of 32 lanes). The result does not depend
            // on the order: acceptance is the (t, geomID, primID) minimum / any accepted hit.
            {
                const uint32_t wid = threadIdx.x >> 5;
                const uint32_t pend = triWork ? T.y : 0u;
                const uint32_t myCnt = __popc(pend);
                uint32_t incl = myCnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(FULL, incl, o); if (lane >= (uint32_t)o) incl += v; }
                const uint32_t total = __shfl_sync(FULL, incl, 31), myOff = incl - myCnt;
                for (uint32_t base = 0; base < total; base += 32u) {              // warp-uniform
                    {   // owners publish their entries that fall into [base, base + 32)
                        uint32_t tmp = pend, idx = myOff;
                        while (tmp) {
                            const uint32_t bit = 31u - __clz(tmp); tmp &= ~(1u << bit);
                            if (idx >= base && idx < base + 32u) smTriList[wid][idx - base] = ((T.x + bit) & 0x07ffffffu) | (lane << 27);
                            idx++;
                        }
                    }
                    __syncwarp();
                    const bool mine = base + lane < total;
                    const uint32_t entry = mine ? smTriList[wid][lane] : (lane << 27);
                    const uint32_t owner = entry >> 27, triIdx = entry & 0x07ffffffu;
                    const float ox = __shfl_sync(FULL, r.O.x, owner), oy = __shfl_sync(FULL, r.O.y, owner), oz = __shfl_sync(FULL, r.O.z, owner);
                    const float dx = __shfl_sync(FULL, r.D.x, owner), dy = __shfl_sync(FULL, r.D.y, owner), dz = __shfl_sync(FULL, r.D.z, owner);
                    const float otn = __shfl_sync(FULL, tnear, owner), otb = __shfl_sync(FULL, tbest, owner);
                    float otime = 0.f; if (MOTION) otime = __shfl_sync(FULL, time, owner);
                    float4 res = make_float4(INFINITY, 0.f, 0.f, 0.f);
                    if (mine) {
                        const float4* tp = tris + 3ull * triIdx;
                        const float4 a = bvh_ld(tp, pol), b = bvh_ld(tp + 1, pol), c = bvh_ld(tp + 2, pol);
                        if (COUNT) cnt.tris++;
                        V3 p0(a.x, a.y, a.z), p1(b.x, b.y, b.z), p2(c.x, c.y, c.z);
                        if (MOTION && (__float_as_uint(c.w) & YRT_TRI_FLAG_MOTION)) {
                            const float4* mp = triMotion + 3ull * triIdx;
                            const float4 d0 = __ldg(mp), d1 = __ldg(mp + 1), d2 = __ldg(mp + 2);
                            p0 = p0 + otime * V3(d0.x, d0.y, d0.z); p1 = p1 + otime * V3(d1.x, d1.y, d1.z); p2 = p2 + otime * V3(d2.x, d2.y, d2.z);
                        }
                        float t, u, v, den; V3 Ng;
                        // candidates beyond the owner's current best cannot win (ties, t == best, are decided by the owner)
                        if (tri_test(V3(ox, oy, oz), V3(dx, dy, dz), p0, p1, p2, t, u, v, Ng, den) && t > otn && (ANY ? t < otb : t <= otb) &&
                            !((__float_as_uint(c.w) & YRT_TRI_FLAG_CULL) && den <= 0.f)) res = make_float4(t, u, v, __uint_as_float(triIdx));
                    }
                    smTriRes[wid][lane] = res;
                    __syncwarp();
                    if (myCnt) {                                                   // owners fold their entries of this window, in list order
                        const uint32_t lo = myOff > base ? myOff : base, hi = (myOff + myCnt < base + 32u) ? myOff + myCnt : base + 32u;
                        for (uint32_t i = lo; i < hi; i++) {
                            const float4 q = smTriRes[wid][i - base];
                            if (!(q.x < INFINITY) || (ANY && occluded)) continue;
                            const uint32_t qTri = __float_as_uint(q.w);
                            bool closer = q.x < tbest;
                            if (!ANY && !closer && bestTri != YRT_NO_TRI && q.x == tbest) {      // tie: (geomID, primID) ascending
                                const int g = __float_as_int(__ldg(tris + 3ull * qTri).w), pr = __float_as_int(__ldg(tris + 3ull * qTri + 1).w);
                                const int bg = __float_as_int(__ldg(tris + 3ull * bestTri).w), bp = __float_as_int(__ldg(tris + 3ull * bestTri + 1).w);
                                closer = (g < bg) || (g == bg && pr < bp);
                            }
                            if (ANY) closer = true;                                    // any accepted candidate inside (tnear, tfar] occludes
                            if (closer) {
                                tbest = q.x; bt = q.x; bu = q.y; bv = q.z; bestTri = qTri;
                                if (ANY) { occluded = true; G.y = 0u; sp = 0; }
                            }
                        }
                    }
                    __syncwarp();
                }
                if (triWork) T.y = 0u;
            }
#else
            // ---- TRIANGLE phase: one triangle per participating lane -------------------------------------
            if (triWork) {
                const uint32_t bit = 31u - __clz(T.y);
                T.y &= ~(1u << bit);
                const uint32_t triIdx = T.x + bit;
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
                        tbest = t; bt = t; bu = u; bv = v; bestTri = triIdx;
                        if (ANY) { occluded = true; G.y = 0u; T.y = 0u; sp = 0; }
                    }
                }
            }
#endif
        } else if (nN != 0) {
            // ---- NODE phase: one compressed node per participating lane --------------------------------
            if (nodeWork) {
                if (T.y) {                                       // postpone the pending triangles
                    if (sp < YRT_SM_STACK) smStack[sp * YRT_TRACE_THREADS + threadIdx.x] = T;
                    else if (sp < YRT_STACK_SIZE) lstack[sp - YRT_SM_STACK] = T;
                    if (sp < YRT_STACK_SIZE) sp++; else io.overflow();     // reported (wb.stats[7] -> the call fails): never a silent drop
                    T.y = 0u;
                }
                const uint32_t bit = 31u - __clz(G.y);
                G.y &= ~(1u << bit);
                const uint32_t slot = (bit - 24u) ^ r.octinv;
                const uint32_t nodeIdx = G.x + __popc((G.y & 0xffu) & ~(0xffffffffu << slot));
                if (G.y & 0xff000000u) {
                    if (sp < YRT_SM_STACK) smStack[sp * YRT_TRACE_THREADS + threadIdx.x] = G;
                    else if (sp < YRT_STACK_SIZE) lstack[sp - YRT_SM_STACK] = G;
                    if (sp < YRT_STACK_SIZE) sp++; else io.overflow();
                }
                const uint4* np = nodes + 5ull * nodeIdx;
                const uint4 n0 = bvh_ld(np, pol), n1 = bvh_ld(np + 1, pol), n2 = bvh_ld(np + 2, pol), n3 = bvh_ld(np + 3, pol), n4 = bvh_ld(np + 4, pol);
                if (COUNT) cnt.nodes++;
                const uint32_t hm = node_test(n0, n1, n2, n3, n4, r, tnear, tbest);
                G = make_uint2(n1.x, (hm & 0xff000000u) | (n0.w >> 24));
                T = make_uint2(n1.y, hm & 0x00ffffffu);
                if (tune.prefetch) {
                    uint32_t m = G.y >> 24;
                    while (m) {
                        const uint32_t b2 = 31u - __clz(m); m &= ~(1u << b2);
                        const uint32_t sl = b2 ^ r.octinv;
                        const char* cp = (const char*)(nodes + 5ull * (G.x + __popc((G.y & 0xffu) & ~(0xffffffffu << sl))));
                        prefetch_l2(cp); prefetch_l2(cp + 64);
                    }
                    if (T.y) prefetch_l2(tris + 3ull * (T.x + (31u - __clz(T.y))));
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
                if (ANY) io.store_any(tag, occluded);
                else io.store_hit(tag, bt, bu, bv, bestTri, tris);     // bestTri == YRT_NO_TRI: miss
                active = false;
            }
        }
    }
}

#endif  // __CUDACC__
}  // namespace yrt
