// bvh_build.cu — GPU construction of the compressed 8-wide BVH (replaces rtcCommit,
// reference: devices/device_singleray/api/scene_flat.h:87-112, which rebuilds for every cube face).
//
// Pipeline (all on the device, one stream):
//   1. k_tri_bounds      triangle AABBs + scene bounds (warp-reduced atomics)
//   2. k_morton          63-bit Morton code of the AABB centre (21 bits per axis)
//   3. cub::DeviceRadixSort::SortPairs (key = Morton code, value = triangle reference index)
//   4. binary hierarchy over the Morton-ordered triangles, one of
//      PLOC (default)    Meister & Bittner 2018, parallel locally-ordered clustering: every cluster looks R positions to either
//                        side for the neighbour with the smallest merged surface area, mutual nearest neighbours merge, the
//                        cluster array is compacted (CUB exclusive scan), repeat until one cluster is left; the last <= 1024
//                        clusters are finished by one CTA in shared memory
//      LBVH (cfg bvh=0)  Karras 2012 topology (k_lbvh_topology) + bottom-up AABB fit (k_lbvh_fit)
//   6. k_collapse        level-synchronous top-down collapse of the binary tree into 8-wide nodes
//                        (largest-surface-area child opened first, subtrees of <= 3 triangles become
//                        leaf children), greedy octant slot assignment, conservative 8-bit plane
//                        quantisation, triangles written in leaf order (48 B each)
// CUB is used as a library primitive for the radix sort only (ships with the CUDA toolkit).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>
#include <nvtx3/nvToolsExt.h>

#include "bvh.cuh"
#include "device_internal.hpp"

namespace yrt {

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) throw std::runtime_error(std::string("CUDA error in BVH build: ") + cudaGetErrorString(e_) + " at " #x); } while (0)

struct Box { float lo[3], hi[3]; };

__device__ __forceinline__ uint32_t f2ord(float f) { uint32_t u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float ord2f(uint32_t u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

__global__ void k_init_bounds(uint32_t* sb) {
    if (threadIdx.x < 3) sb[threadIdx.x] = 0xffffffffu;       // min (ordered encoding)
    else if (threadIdx.x < 6) sb[threadIdx.x] = 0u;           // max
}

__global__ void k_tri_bounds(const uint2* __restrict__ refs, uint32_t n, const GeomRec* __restrict__ geoms,
                             const float4* __restrict__ positions, const int4* __restrict__ indices, const float4* __restrict__ motions,
                             float4* __restrict__ boxLo, float4* __restrict__ boxHi, uint32_t* __restrict__ sceneBounds) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    if (i < n) {
        const uint2 r = refs[i];
        const GeomRec g = geoms[r.x];
        const int4 t = indices[g.idxBase + r.y];
        const float4 a = positions[g.vtxBase + t.x], b = positions[g.vtxBase + t.y], c = positions[g.vtxBase + t.z];
        lo[0] = fminf(a.x, fminf(b.x, c.x)); hi[0] = fmaxf(a.x, fmaxf(b.x, c.x));
        lo[1] = fminf(a.y, fminf(b.y, c.y)); hi[1] = fmaxf(a.y, fmaxf(b.y, c.y));
        lo[2] = fminf(a.z, fminf(b.z, c.z)); hi[2] = fmaxf(a.z, fmaxf(b.z, c.z));
        if (motions && g.motBase != YRT_NO_ATTR) {               // a moving triangle is bounded over the whole shutter interval: both end positions
            const int vi[3] = {t.x, t.y, t.z}; const float4 p[3] = {a, b, c};
            for (int k = 0; k < 3; k++) {
                const float4 m = motions[g.motBase + vi[k]];
                const float e[3] = {p[k].x + m.x, p[k].y + m.y, p[k].z + m.z};
                for (int ax = 0; ax < 3; ax++) { lo[ax] = fminf(lo[ax], e[ax]); hi[ax] = fmaxf(hi[ax], e[ax]); }
            }
        }
        boxLo[i] = make_float4(lo[0], lo[1], lo[2], 0.f);
        boxHi[i] = make_float4(hi[0], hi[1], hi[2], 0.f);
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float l = lo[k], h = hi[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { l = fminf(l, __shfl_xor_sync(0xffffffffu, l, o)); h = fmaxf(h, __shfl_xor_sync(0xffffffffu, h, o)); }
        if ((threadIdx.x & 31) == 0 && l <= h) { atomicMin(&sceneBounds[k], f2ord(l)); atomicMax(&sceneBounds[3 + k], f2ord(h)); }
    }
}

__device__ __forceinline__ uint64_t spread21(uint64_t x) {
    x &= 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

__global__ void k_morton(const float4* __restrict__ boxLo, const float4* __restrict__ boxHi, uint32_t n,
                         const uint32_t* __restrict__ sceneBounds, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 lo = boxLo[i], hi = boxHi[i];
    float c[3] = {0.5f * lo.x + 0.5f * hi.x, 0.5f * lo.y + 0.5f * hi.y, 0.5f * lo.z + 0.5f * hi.z};
    uint64_t q[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float mn = ord2f(sceneBounds[k]), mx = ord2f(sceneBounds[3 + k]);
        const float ext = mx - mn;
        float f = ext > 0.f ? (c[k] - mn) / ext : 0.f;
        f = fminf(fmaxf(f, 0.f), 1.f);
        q[k] = (uint64_t)fminf(f * 2097152.0f, 2097151.0f);
    }
    keys[i] = (spread21(q[0]) << 2) | (spread21(q[1]) << 1) | spread21(q[2]);
    vals[i] = i;
}

// ---- Karras LBVH ---------------------------------------------------------------------------
// node reference encoding: bit 31 set = leaf (sorted position in low bits), else internal index
#define LEAF_FLAG 0x80000000u

__device__ __forceinline__ int delta(const uint64_t* __restrict__ keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    const uint64_t a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz((uint32_t)i ^ (uint32_t)j);
    return __clzll((long long)(a ^ b));
}

struct Lbvh {                                               // the binary tree both builders produce
    uint32_t* left; uint32_t* right; uint32_t* count;       // per internal node: children references, triangles below
    float4* lo; float4* hi;                                 // per internal node
    const float4* leafLo; const float4* leafHi;             // per leaf, by sorted position
    uint32_t* parent; uint32_t* visit;                      // LBVH only (parent also per leaf at [n-1 + leaf])
};

__global__ void k_ploc_init(uint32_t n, uint32_t* __restrict__ cid) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) cid[i] = 0x80000000u | i;
}

__global__ void k_gather_leaf_boxes(const uint32_t* __restrict__ sortedIdx, uint32_t n, const float4* __restrict__ boxLo, const float4* __restrict__ boxHi,
                                    float4* __restrict__ leafLo, float4* __restrict__ leafHi) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t s = sortedIdx[i];
    leafLo[i] = boxLo[s]; leafHi[i] = boxHi[s];
}

__global__ void k_lbvh_topology(const uint64_t* __restrict__ keys, int n, Lbvh t) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int s = lmax >> 1; s > 0; s >>= 1)
        if (delta(keys, n, i, i + (l + s) * d) > dmin) l += s;
    const int j = i + l * d;
    const int dnode = delta(keys, n, i, j);
    int s = 0;
    for (int div = 2, tt = (l + div - 1) / div; ; div <<= 1, tt = (l + div - 1) / div) {
        if (delta(keys, n, i, i + (s + tt) * d) > dnode) s += tt;
        if (tt <= 1) break;
    }
    const int gamma = i + s * d + min(d, 0);
    const int first = min(i, j), last = max(i, j);
    const uint32_t L = (first == gamma) ? (LEAF_FLAG | (uint32_t)gamma) : (uint32_t)gamma;
    const uint32_t R = (last == gamma + 1) ? (LEAF_FLAG | (uint32_t)(gamma + 1)) : (uint32_t)(gamma + 1);
    t.left[i] = L; t.right[i] = R; t.count[i] = (uint32_t)(last - first + 1);
    if (L & LEAF_FLAG) t.parent[(n - 1) + gamma] = i; else t.parent[gamma] = i;
    if (R & LEAF_FLAG) t.parent[(n - 1) + gamma + 1] = i; else t.parent[gamma + 1] = i;
    if (i == 0) t.parent[0] = 0xffffffffu;
}

__global__ void k_lbvh_fit(int n, Lbvh t) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t node = t.parent[(n - 1) + i];
    while (node != 0xffffffffu) {
        if (atomicAdd(&t.visit[node], 1u) == 0u) return;     // first arrival waits for the sibling
        __threadfence();
        const uint32_t L = t.left[node], R = t.right[node];
        float4 llo, lhi, rlo, rhi;
        if (L & LEAF_FLAG) { llo = t.leafLo[L & ~LEAF_FLAG]; lhi = t.leafHi[L & ~LEAF_FLAG]; }
        else { llo = __ldcg(&t.lo[L]); lhi = __ldcg(&t.hi[L]); }
        if (R & LEAF_FLAG) { rlo = t.leafLo[R & ~LEAF_FLAG]; rhi = t.leafHi[R & ~LEAF_FLAG]; }
        else { rlo = __ldcg(&t.lo[R]); rhi = __ldcg(&t.hi[R]); }
        t.lo[node] = make_float4(fminf(llo.x, rlo.x), fminf(llo.y, rlo.y), fminf(llo.z, rlo.z), 0.f);
        t.hi[node] = make_float4(fmaxf(lhi.x, rhi.x), fmaxf(lhi.y, rhi.y), fmaxf(lhi.z, rhi.z), 0.f);
        __threadfence();
        node = t.parent[node];
    }
}

// ---- PLOC (parallel locally-ordered clustering) ---------------------------------------------------
// Cluster i of an iteration = (node reference, box) at position i of a compact, Morton-ordered array.
#define PLOC_BLOCK 256
#define PLOC_MAX_RADIUS 32
#define PLOC_FINAL 1024          // cluster count at which one CTA finishes the tree in shared memory

__device__ __forceinline__ float union_area(const float4 alo, const float4 ahi, const float4 blo, const float4 bhi) {
    const float dx = fmaxf(ahi.x, bhi.x) - fminf(alo.x, blo.x), dy = fmaxf(ahi.y, bhi.y) - fminf(alo.y, blo.y), dz = fmaxf(ahi.z, bhi.z) - fminf(alo.z, blo.z);
    return dx * dy + dy * dz + dz * dx;
}

// nearest neighbour within +-radius positions: minimum merged surface area, ties to the smaller position (then the pair with the
// globally smallest area is always mutual, so every iteration merges at least one pair)
__global__ void __launch_bounds__(PLOC_BLOCK) k_ploc_nn(const float4* __restrict__ lo, const float4* __restrict__ hi, uint32_t m, int radius, uint32_t* __restrict__ nn) {
    __shared__ float4 sLo[PLOC_BLOCK + 2 * PLOC_MAX_RADIUS], sHi[PLOC_BLOCK + 2 * PLOC_MAX_RADIUS];
    const int base = (int)(blockIdx.x * PLOC_BLOCK) - radius;
    for (int k = threadIdx.x; k < PLOC_BLOCK + 2 * radius; k += PLOC_BLOCK) {
        const int g = base + k;
        if (g >= 0 && g < (int)m) { sLo[k] = lo[g]; sHi[k] = hi[g]; }
    }
    __syncthreads();
    const int i = (int)(blockIdx.x * PLOC_BLOCK + threadIdx.x);
    if (i >= (int)m) return;
    const float4 alo = sLo[threadIdx.x + radius], ahi = sHi[threadIdx.x + radius];
    float best = INFINITY; int bj = -1;
    for (int d = -radius; d <= radius; d++) {
        const int j = i + d;
        if (d == 0 || j < 0 || j >= (int)m) continue;
        const float a = union_area(alo, ahi, sLo[threadIdx.x + radius + d], sHi[threadIdx.x + radius + d]);
        if (a < best) { best = a; bj = j; }
    }
    nn[i] = (uint32_t)bj;
}

// mutual pairs merge into a new internal node that takes the lower position; flags mark surviving positions
__global__ void __launch_bounds__(PLOC_BLOCK) k_ploc_merge(uint32_t m, const uint32_t* __restrict__ nn, const uint32_t* __restrict__ cid,
                                                           const float4* __restrict__ lo, const float4* __restrict__ hi, Lbvh t, uint32_t* __restrict__ nodeCounter,
                                                           uint32_t* __restrict__ cidTmp, float4* __restrict__ loTmp, float4* __restrict__ hiTmp, uint32_t* __restrict__ flags) {
    const uint32_t i = blockIdx.x * PLOC_BLOCK + threadIdx.x;
    if (i >= m) return;
    const uint32_t j = nn[i];
    uint32_t id = cid[i]; float4 l = lo[i], h = hi[i]; uint32_t keep = 1u;
    if (j < m && nn[j] == i) {
        if (i < j) {
            const uint32_t other = cid[j]; const float4 ol = lo[j], oh = hi[j];
            const uint32_t node = atomicAdd(nodeCounter, 1u);
            t.left[node] = id; t.right[node] = other;
            t.count[node] = ((id & LEAF_FLAG) ? 1u : t.count[id]) + ((other & LEAF_FLAG) ? 1u : t.count[other]);
            l = make_float4(fminf(l.x, ol.x), fminf(l.y, ol.y), fminf(l.z, ol.z), 0.f);
            h = make_float4(fmaxf(h.x, oh.x), fmaxf(h.y, oh.y), fmaxf(h.z, oh.z), 0.f);
            t.lo[node] = l; t.hi[node] = h;
            id = node;
        } else keep = 0u;
    }
    cidTmp[i] = id; loTmp[i] = l; hiTmp[i] = h; flags[i] = keep;
}

__global__ void __launch_bounds__(PLOC_BLOCK) k_ploc_compact(uint32_t m, const uint32_t* __restrict__ flags, const uint32_t* __restrict__ offsets,
                                                             const uint32_t* __restrict__ cidTmp, const float4* __restrict__ loTmp, const float4* __restrict__ hiTmp,
                                                             uint32_t* __restrict__ cid, float4* __restrict__ lo, float4* __restrict__ hi, uint32_t* __restrict__ mOut) {
    const uint32_t i = blockIdx.x * PLOC_BLOCK + threadIdx.x;
    if (i >= m) return;
    if (flags[i]) { const uint32_t o = offsets[i]; cid[o] = cidTmp[i]; lo[o] = loTmp[i]; hi[o] = hiTmp[i]; }
    if (i == m - 1) *mOut = offsets[i] + flags[i];
}

// the last m <= PLOC_FINAL clusters: the same iteration, in one CTA, until the root is left (rootOut = its reference)
__global__ void __launch_bounds__(PLOC_FINAL) k_ploc_final(uint32_t m, int radius, const uint32_t* __restrict__ cid, const float4* __restrict__ lo, const float4* __restrict__ hi,
                                                           Lbvh t, uint32_t* __restrict__ nodeCounter, uint32_t* __restrict__ rootOut) {
    __shared__ float4 sLo[PLOC_FINAL], sHi[PLOC_FINAL];
    __shared__ uint32_t sId[PLOC_FINAL], sNn[PLOC_FINAL], sWarp[32], sTotal;
    const uint32_t i = threadIdx.x, lane = i & 31u, warp = i >> 5;
    if (i < m) { sLo[i] = lo[i]; sHi[i] = hi[i]; sId[i] = cid[i]; }
    __syncthreads();
    while (m > 1) {
        if (i < m) {
            const float4 alo = sLo[i], ahi = sHi[i];
            float best = INFINITY; int bj = -1;
            for (int d = -radius; d <= radius; d++) {
                const int j = (int)i + d;
                if (d == 0 || j < 0 || j >= (int)m) continue;
                const float a = union_area(alo, ahi, sLo[j], sHi[j]);
                if (a < best) { best = a; bj = j; }
            }
            sNn[i] = (uint32_t)bj;
        }
        __syncthreads();
        uint32_t id = 0, keep = 0; float4 l = make_float4(0, 0, 0, 0), h = l;
        if (i < m) {
            const uint32_t j = sNn[i];
            id = sId[i]; l = sLo[i]; h = sHi[i]; keep = 1u;
            if (j < m && sNn[j] == i) {
                if (i < j) {
                    const uint32_t other = sId[j]; const float4 ol = sLo[j], oh = sHi[j];
                    const uint32_t node = atomicAdd(nodeCounter, 1u);
                    t.left[node] = id; t.right[node] = other;
                    t.count[node] = ((id & LEAF_FLAG) ? 1u : t.count[id]) + ((other & LEAF_FLAG) ? 1u : t.count[other]);
                    l = make_float4(fminf(l.x, ol.x), fminf(l.y, ol.y), fminf(l.z, ol.z), 0.f);
                    h = make_float4(fmaxf(h.x, oh.x), fmaxf(h.y, oh.y), fmaxf(h.z, oh.z), 0.f);
                    t.lo[node] = l; t.hi[node] = h;
                    id = node;
                } else keep = 0u;
            }
        }
        // block-wide exclusive scan of keep
        uint32_t incl = keep;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (uint32_t)o) incl += v; }
        if (lane == 31) sWarp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = sWarp[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= (uint32_t)o) wi += v; }
            sWarp[lane] = wi - w;
            if (lane == 31) sTotal = wi;
        }
        __syncthreads();
        const uint32_t pos = sWarp[warp] + incl - keep, total = sTotal;      // every read of this round's clusters happened before the barriers above
        if (keep) { sId[pos] = id; sLo[pos] = l; sHi[pos] = h; }
        __threadfence_block();
        __syncthreads();
        m = total;
    }
    if (i == 0) *rootOut = sId[0];
}

// ---- collapse to compressed BVH8 -------------------------------------------------------------
struct CollapseArgs {
    Lbvh t; int n;
    const uint32_t* sortedIdx;
    const uint2* refs; const GeomRec* geoms; const float4* positions; const int4* indices; const float4* normals; const float2* uvs;
    const float4* motions; float4* triMotion;
    Node8* nodes; float4* tris; float4* triShade;
    uint32_t* counters;          // [0] node counter, [1] triangle counter, [2] next-level task count
    const uint2* tasksIn; uint2* tasksOut; uint32_t numTasks; int splitLeaves;
    const float* dpF; const uint8_t* dpDec;      // SAH-optimal collapse (k_collapse_dp): NULL = greedy opening
};

__device__ __forceinline__ void ref_box(const CollapseArgs& a, uint32_t ref, float lo[3], float hi[3]) {
    float4 l, h;
    if (ref & LEAF_FLAG) { l = a.t.leafLo[ref & ~LEAF_FLAG]; h = a.t.leafHi[ref & ~LEAF_FLAG]; }
    else { l = a.t.lo[ref]; h = a.t.hi[ref]; }
    lo[0] = l.x; lo[1] = l.y; lo[2] = l.z; hi[0] = h.x; hi[1] = h.y; hi[2] = h.z;
}
__device__ __forceinline__ uint32_t ref_count(const CollapseArgs& a, uint32_t ref) { return (ref & LEAF_FLAG) ? 1u : a.t.count[ref]; }
// sorted positions of the (<= 3) triangles below `ref`, left to right
__device__ __forceinline__ uint32_t ref_leaves(const CollapseArgs& a, uint32_t ref, uint32_t out[3]) {
    uint32_t stack[3]; int sp = 0; uint32_t n = 0;
    stack[sp++] = ref;
    while (sp > 0 && n < 3u) {
        const uint32_t r = stack[--sp];
        if (r & LEAF_FLAG) out[n++] = r & ~LEAF_FLAG;
        else { stack[sp++] = a.t.right[r]; if (sp < 3) stack[sp++] = a.t.left[r]; }
    }
    return n;
}

// Leaf-order triangle (48 B, traversal) + its shading record (80 B): everything postIntersect reads, gathered once at build time so that
// the shading kernel reaches it with ONE dependent load from the hit (triangle index) instead of geometry record -> index triple ->
// three vertices. r0 = (Ng | geomID << 2 | hasNormals | hasUV << 1), r1..r3 = (n_k | uv pieces), r4 = (uv1.y, uv2.x, uv2.y, primID).
// Ng = normalize(cross(p0 - p1, p2 - p0)): the arithmetic of the hit record + TriangleMeshFull::postIntersect (trianglemesh_full.cpp:221).
__device__ void write_triangle(const CollapseArgs& a, uint32_t sortedPos, uint32_t outIdx) {
    const uint2 r = a.refs[a.sortedIdx[sortedPos]];
    const GeomRec g = a.geoms[r.x];
    const int4 t = a.indices[g.idxBase + r.y];
    const float4 p0 = a.positions[g.vtxBase + t.x], p1 = a.positions[g.vtxBase + t.y], p2 = a.positions[g.vtxBase + t.z];
    float4* o = a.tris + 3ull * outIdx;
    o[0] = make_float4(p0.x, p0.y, p0.z, __int_as_float((int)r.x));
    o[1] = make_float4(p1.x, p1.y, p1.z, __int_as_float((int)r.y));
    const bool moving = a.motions && g.motBase != YRT_NO_ATTR;
    o[2] = make_float4(p2.x, p2.y, p2.z, __uint_as_float((g.cull ? YRT_TRI_FLAG_CULL : 0u) | (moving ? YRT_TRI_FLAG_MOTION : 0u)));
    if (a.triMotion) {                                          // d = vertex(t = 1) - vertex(t = 0) with vertex(t = 1) = p + motion rounded first,
        float4* mo = a.triMotion + 3ull * outIdx;               // exactly what the reference hands to the second vertex buffer (trianglemesh_full.cpp:157-160)
        const int vi[3] = {t.x, t.y, t.z}; const float4 pv[3] = {p0, p1, p2};
        for (int k = 0; k < 3; k++) {
            float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
            if (moving) { const float4 m = a.motions[g.motBase + vi[k]]; d = make_float4((pv[k].x + m.x) - pv[k].x, (pv[k].y + m.y) - pv[k].y, (pv[k].z + m.z) - pv[k].z, 0.f); }
            mo[k] = d;
        }
    }
    const V3 q0(p0.x, p0.y, p0.z), q1(p1.x, p1.y, p1.z), q2(p2.x, p2.y, p2.z);
    V3 Ng = g.type == MESH_TRIANGLE ? g.triNg : normalize(cross(q0 - q1, q2 - q0));
    uint32_t flags = (uint32_t)r.x << 2;
    float4 n0 = make_float4(0, 0, 0, 0), n1 = n0, n2 = n0; float2 s0 = make_float2(0, 0), s1 = s0, s2 = s0;
    if (g.type != MESH_TRIANGLE) {
        if (g.nrmBase != YRT_NO_ATTR) { flags |= 1u; n0 = a.normals[g.nrmBase + t.x]; n1 = a.normals[g.nrmBase + t.y]; n2 = a.normals[g.nrmBase + t.z]; }
        if (g.uvBase != YRT_NO_ATTR) { flags |= 2u; s0 = a.uvs[g.uvBase + t.x]; s1 = a.uvs[g.uvBase + t.y]; s2 = a.uvs[g.uvBase + t.z]; }
    }
    float4* h = a.triShade + 5ull * outIdx;
    h[0] = make_float4(Ng.x, Ng.y, Ng.z, __uint_as_float(flags));
    h[1] = make_float4(n0.x, n0.y, n0.z, s0.x);
    h[2] = make_float4(n1.x, n1.y, n1.z, s0.y);
    h[3] = make_float4(n2.x, n2.y, n2.z, s1.x);
    h[4] = make_float4(s1.y, s2.x, s2.y, __int_as_float((int)r.y));
}


// ---- SAH-optimal collapse: dynamic programming over the binary tree (Ylitie, Karras, Laine 2017, section 4) -----------------------
// The greedy opening below fills the 8 slots of a node with the largest subtrees and leaves whatever has 4..7 triangles as an inner child
// of its own: on the C4 office 85 % of the BVH8 nodes sit at the bottom of the tree with 3.8 of 8 slots used. The DP chooses, for every
// binary subtree n and every slot budget j = 1..8, the cheapest way to hand n to a parent node as at most j entries:
//     F(n, 1) = min( Leaf(n) = A(n) T(n) c_tri   if T(n) <= 3,      Inner(n) = A(n) c_node + S(n, 8) )
//     S(n, j) = min over k = 1..j-1 of F(left, k) + F(right, j - k)                          (n's two children share j entries)
//     F(n, j) = min( F(n, j - 1), S(n, j) )   for j >= 2;       F(triangle, j) = A c_tri
// with A = surface area of the subtree's box (probability that a ray visits it), T = triangles below. Decisions are kept per node:
// byte 0 = flags (bit 0: F(n,1) is a leaf; bit j-1: F(n,j) takes the split S(n,j) rather than F(n,j-1)), bytes 1..7 = argmin k of S(n, 2..8).
// Bottom-up over parent pointers with one arrival counter per node (the second child to arrive computes the parent).
struct DpArgs { Lbvh t; uint32_t n; uint32_t rootRef; float* F; uint8_t* dec; float cNode, cTri; };

__global__ void k_bvh2_parents(Lbvh t, uint32_t ni, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ni) return;
    const uint32_t l = t.left[i], r = t.right[i];
    t.parent[(l & LEAF_FLAG) ? ni + (l & ~LEAF_FLAG) : l] = i;
    t.parent[(r & LEAF_FLAG) ? ni + (r & ~LEAF_FLAG) : r] = i;
}

__device__ __forceinline__ float box_area(float4 lo, float4 hi) {
    const float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z;
    return dx * dy + dy * dz + dz * dx;
}

__global__ void k_collapse_dp(DpArgs a) {
    const uint32_t leaf = blockIdx.x * blockDim.x + threadIdx.x;
    if (leaf >= a.n) return;
    const uint32_t ni = a.n - 1;
    uint32_t cur = a.t.parent[ni + leaf];
    while (true) {
        __threadfence();                                          // this thread's tables of the child are visible before it reports arrival
        if (atomicAdd(&a.t.visit[cur], 1u) == 0u) return;         // the sibling subtree is not finished: its thread computes this node
        float Fc[2][8];
#pragma unroll
        for (int c = 0; c < 2; c++) {
            const uint32_t r = c ? a.t.right[cur] : a.t.left[cur];
            if (r & LEAF_FLAG) {
                const float v = box_area(a.t.leafLo[r & ~LEAF_FLAG], a.t.leafHi[r & ~LEAF_FLAG]) * a.cTri;
#pragma unroll
                for (int j = 0; j < 8; j++) Fc[c][j] = v;
            } else {
#pragma unroll
                for (int j = 0; j < 8; j++) Fc[c][j] = __ldcg(&a.F[8ull * r + j]);          // written by another SM: bypass L1
            }
        }
        const float A = box_area(a.t.lo[cur], a.t.hi[cur]);
        const uint32_t T = a.t.count[cur];
        float S[9]; uint8_t K[9];
#pragma unroll
        for (int j = 2; j <= 8; j++) {
            float best = INFINITY; int bk = 1;
#pragma unroll
            for (int k = 1; k < j; k++) { const float v = Fc[0][k - 1] + Fc[1][j - k - 1]; if (v < best) { best = v; bk = k; } }
            S[j] = best; K[j] = (uint8_t)bk;
        }
        const float inner = A * a.cNode + S[8], leafCost = T <= 3u ? A * (float)T * a.cTri : INFINITY;
        float F[9]; uint32_t flags = 0;
        F[1] = fminf(leafCost, inner); if (leafCost <= inner) flags |= 1u;
#pragma unroll
        for (int j = 2; j <= 8; j++) { if (S[j] < F[j - 1]) { F[j] = S[j]; flags |= 1u << (j - 1); } else F[j] = F[j - 1]; }
#pragma unroll
        for (int j = 1; j <= 8; j++) a.F[8ull * cur + j - 1] = F[j];
        uint8_t* d = a.dec + 8ull * cur;
        d[0] = (uint8_t)flags;
#pragma unroll
        for (int j = 2; j <= 8; j++) d[j - 1] = K[j];
        if (cur == a.rootRef) return;
        cur = a.t.parent[cur];
    }
}

// The entries (<= 8) of the BVH8 node made from binary node `root`, following the DP's decisions. leafMask: bit c set = entry c is a leaf child.
__device__ int dp_expand(const CollapseArgs& a, uint32_t root, uint32_t cand[8], uint32_t& leafMask) {
    struct Item { uint32_t ref; uint8_t j, split; };
    Item stack[16]; int sp = 0, count = 0;
    leafMask = 0;
    stack[sp++] = Item{root, 8, 1};
    while (sp > 0) {
        const Item it = stack[--sp];
        if (it.ref & LEAF_FLAG) { leafMask |= 1u << count; cand[count++] = it.ref; continue; }
        const uint8_t* d = a.dpDec + 8ull * it.ref;
        int j = it.j;
        if (!it.split) {
            while (j > 1 && !((d[0] >> (j - 1)) & 1u)) j--;      // F(n, j) fell back to F(n, j - 1)
            if (j == 1) { if (d[0] & 1u) leafMask |= 1u << count; cand[count++] = it.ref; continue; }
        }
        const int k = d[j - 1];                                   // left takes k entries, right j - k; left is expanded first (pushed last)
        stack[sp++] = Item{a.t.right[it.ref], (uint8_t)(j - k), 0};
        stack[sp++] = Item{a.t.left[it.ref], (uint8_t)k, 0};
    }
    return count;
}

__global__ void k_collapse(CollapseArgs a) {
    const uint32_t ti = blockIdx.x * blockDim.x + threadIdx.x;
    if (ti >= a.numTasks) return;
    const uint2 task = a.tasksIn[ti];                 // x = BVH2 reference, y = output node index
    uint32_t cand[8]; int count;
    uint32_t leafMask = 0; bool useMask = false;             // DP mode: which entries are leaf children (a subtree of <= 3 triangles may also become a node)
    float nlo[3], nhi[3];
    ref_box(a, task.x, nlo, nhi);
    if (task.x & LEAF_FLAG) { cand[0] = task.x; count = 1; }    // single-triangle scene
    else if (a.dpDec) { count = dp_expand(a, task.x, cand, leafMask); useMask = true; }
    else {
        cand[0] = a.t.left[task.x]; cand[1] = a.t.right[task.x]; count = 2;
        // pass 0 opens subtrees of more than 3 triangles, largest surface area first; pass 1 uses slots that are still free to split
        // the small ones too (a leaf child of 3 triangles becomes 2 + 1 or 1 + 1 + 1 with tighter boxes): the node test costs the
        // same for 3 or 8 occupied slots, every triangle test avoided is a gain
        for (int pass = 0; pass < (a.splitLeaves ? 2 : 1); pass++)
        while (count < 8) {
            int best = -1; float bestArea = -1.f;
            for (int c = 0; c < count; c++) {
                if ((cand[c] & LEAF_FLAG) || (pass == 0 && ref_count(a, cand[c]) <= 3u)) continue;
                float lo[3], hi[3]; ref_box(a, cand[c], lo, hi);
                const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
                const float area = dx * dy + dy * dz + dz * dx;
                if (area > bestArea) { bestArea = area; best = c; }
            }
            if (best < 0) break;
            const uint32_t open = cand[best];
            cand[best] = a.t.left[open];
            cand[count++] = a.t.right[open];
        }
    }
    // greedy octant slot assignment (slot bit 4/2/1 = child lies towards +x/+y/+z of the node centre)
    float cx[8], cy[8], cz[8];
    const float ncx = 0.5f * (nlo[0] + nhi[0]), ncy = 0.5f * (nlo[1] + nhi[1]), ncz = 0.5f * (nlo[2] + nhi[2]);
    for (int c = 0; c < count; c++) {
        float lo[3], hi[3]; ref_box(a, cand[c], lo, hi);
        cx[c] = 0.5f * (lo[0] + hi[0]) - ncx; cy[c] = 0.5f * (lo[1] + hi[1]) - ncy; cz[c] = 0.5f * (lo[2] + hi[2]) - ncz;
    }
    int slotOf[8]; uint32_t slotUsed = 0, childDone = 0;
    for (int c = 0; c < 8; c++) slotOf[c] = -1;
    for (int it = 0; it < count; it++) {
        float bestCost = -INFINITY; int bc = -1, bs = -1;
        for (int c = 0; c < count; c++) {
            if (childDone & (1u << c)) continue;
            for (int s = 0; s < 8; s++) {
                if (slotUsed & (1u << s)) continue;
                const float cost = ((s & 4) ? cx[c] : -cx[c]) + ((s & 2) ? cy[c] : -cy[c]) + ((s & 1) ? cz[c] : -cz[c]);
                if (cost > bestCost) { bestCost = cost; bc = c; bs = s; }
            }
        }
        slotOf[bc] = bs; slotUsed |= 1u << bs; childDone |= 1u << bc;
    }
    uint32_t slotRef[8]; bool slotValid[8]; uint32_t slotLeaf = 0;
    for (int s = 0; s < 8; s++) slotValid[s] = false;
    for (int c = 0; c < count; c++) {
        slotRef[slotOf[c]] = cand[c]; slotValid[slotOf[c]] = true;
        const bool leafChild = useMask ? ((leafMask >> c) & 1u) != 0u : ((cand[c] & LEAF_FLAG) || ref_count(a, cand[c]) <= 3u);
        if (leafChild) slotLeaf |= 1u << slotOf[c];
    }

    // quantisation frame
    Node8 nd;
    nd.px = nlo[0]; nd.py = nlo[1]; nd.pz = nlo[2];
    float scale[3]; uint8_t ebyte[3];
    for (int k = 0; k < 3; k++) {
        const float ext = nhi[k] - nlo[k];
        int e;
        if (!(ext > 0.f)) e = -120;
        else {
            int fe; frexpf(ext / 255.0f, &fe);       // ext/255 = m * 2^fe, m in [0.5,1)  ->  2^fe >= ext/255
            e = fe;
            while (ldexpf(255.0f, e) < ext * 1.000001f) e++;
        }
        e = max(-126, min(127, e));
        ebyte[k] = (uint8_t)(e + 127);
        scale[k] = __uint_as_float((uint32_t)ebyte[k] << 23);
    }
    nd.ex = ebyte[0]; nd.ey = ebyte[1]; nd.ez = ebyte[2];

    // leaf triangles: bit 8k + s of triMask = "slot s is a leaf child with more than k triangles" (k < 3); the node's triangles are stored
    // in the order of those bits, so the triangle of bit b is triBase + popc(triMask & ((1 << b) - 1)) (bvh.cuh)
    uint32_t nInner = 0, nTris = 0, imask = 0, triMask = 0;
    for (int s = 0; s < 8; s++) {
        if (!slotValid[s]) continue;
        const uint32_t r = slotRef[s];
        const uint32_t cnt = ref_count(a, r);
        if (!((slotLeaf >> s) & 1u)) { nInner++; imask |= 1u << s; }
        else { nTris += cnt; for (uint32_t k = 0; k < cnt; k++) triMask |= 1u << (8u * k + (uint32_t)s); }
    }
    nd.imask = (uint8_t)imask;
    const uint32_t childBase = nInner ? atomicAdd(&a.counters[0], nInner) : 0u;
    const uint32_t triBase = nTris ? atomicAdd(&a.counters[1], nTris) : 0u;
    const uint32_t taskBase = nInner ? atomicAdd(&a.counters[2], nInner) : 0u;
    nd.childBase = childBase; nd.triBase = triBase; nd.triMask = triMask; nd.reserved = 0u;

    uint32_t innerSeen = 0;
    uint8_t* qplanes[6] = {nd.qlox, nd.qloy, nd.qloz, nd.qhix, nd.qhiy, nd.qhiz};
    const float org[3] = {nlo[0], nlo[1], nlo[2]};
    for (int s = 0; s < 8; s++) {
        if (!slotValid[s]) {
            for (int k = 0; k < 3; k++) { qplanes[k][s] = 255; qplanes[3 + k][s] = 0; }   // inverted box: never hit
            continue;
        }
        const uint32_t r = slotRef[s];
        float lo[3], hi[3]; ref_box(a, r, lo, hi);
        for (int k = 0; k < 3; k++) {
            int ql = (int)floorf((lo[k] - org[k]) / scale[k]);
            ql = max(0, min(255, ql));
            while (ql > 0 && fmaf((float)ql, scale[k], org[k]) > lo[k]) ql--;
            int qh = (int)ceilf((hi[k] - org[k]) / scale[k]);
            qh = max(0, min(255, qh));
            while (qh < 255 && fmaf((float)qh, scale[k], org[k]) < hi[k]) qh++;
            qplanes[k][s] = (uint8_t)ql; qplanes[3 + k][s] = (uint8_t)qh;
        }
        if (imask & (1u << s)) {
            a.tasksOut[taskBase + innerSeen] = make_uint2(r, childBase + innerSeen);
            innerSeen++;
        } else {
            uint32_t leaves[3]; const uint32_t nl = ref_leaves(a, r, leaves);
            for (uint32_t k = 0; k < nl; k++) write_triangle(a, leaves[k], triBase + (uint32_t)__popc(triMask & ((1u << (8u * k + (uint32_t)s)) - 1u)));
        }
    }
    a.nodes[task.y] = nd;
}

// ---------------------------------------------------------------------------------------------
// Scratch comes from the stream-ordered pool (cudaMallocAsync): a rebuild per cube face must not pay
// cudaMalloc/cudaFree round trips (measured: 30 ms ... 1.4 s of wall clock per build with the synchronous allocator).
// Every scratch block and both events belong to a BuildScratch: whatever is still registered when build_bvh leaves — normally or
// through an exception (out of memory, a failed launch) — is released by its destructor.
struct BuildScratch {
    cudaStream_t stream; std::vector<void*> live; cudaEvent_t e0 = nullptr, e1 = nullptr;
    explicit BuildScratch(cudaStream_t s) : stream(s) {}
    void forget(void* p) { for (size_t i = live.size(); i-- > 0;) if (live[i] == p) { live.erase(live.begin() + (long)i); return; } }
    ~BuildScratch() { for (void* p : live) cudaFreeAsync(p, stream); if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); }
};
static thread_local BuildScratch* g_scratch = nullptr;      // per host thread: a group device builds on several GPUs at once (group_api.cu)
template <typename T> static T* dalloc(size_t n) {
    T* p = nullptr; CK(cudaMallocAsync((void**)&p, (n ? n : 1) * sizeof(T), g_scratch->stream));
    g_scratch->live.push_back(p);
    return p;
}
static void dfree(void* p) { if (p) { g_scratch->forget(p); cudaFreeAsync(p, g_scratch->stream); } }

void build_bvh(const BvhBuildInput& in, BvhResult& out, cudaStream_t stream) {
    out.nodes = nullptr; out.tris = nullptr; out.triShade = nullptr; out.triMotion = nullptr; out.numNodes = 0; out.numTris = 0; out.buildMs = 0.f; out.launches = 0;
    const uint32_t n = in.numRefs;
    if (n == 0) return;
    nvtxRangePushA("device_cuda: BVH8 build");
    struct Pop { ~Pop() { nvtxRangePop(); } } nvtxPop;
    BuildScratch scratch(stream); g_scratch = &scratch;
    CK(cudaEventCreate(&scratch.e0)); CK(cudaEventCreate(&scratch.e1));
    const cudaEvent_t e0 = scratch.e0, e1 = scratch.e1;
    CK(cudaEventRecord(e0, stream));
    const int B = 256; const uint32_t G = (n + B - 1) / B;

    float4* boxLo = dalloc<float4>(n); float4* boxHi = dalloc<float4>(n);
    uint32_t* sceneBounds = dalloc<uint32_t>(8);
    uint64_t* keys = dalloc<uint64_t>(n); uint64_t* keysSorted = dalloc<uint64_t>(n);
    uint32_t* vals = dalloc<uint32_t>(n); uint32_t* sortedIdx = dalloc<uint32_t>(n);
    k_init_bounds<<<1, 32, 0, stream>>>(sceneBounds);
    k_tri_bounds<<<G, B, 0, stream>>>(in.refs, n, in.geoms, in.positions, in.indices, in.motions, boxLo, boxHi, sceneBounds);
    k_morton<<<G, B, 0, stream>>>(boxLo, boxHi, n, sceneBounds, keys, vals);
    out.launches += 3;
    size_t tmpBytes = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, tmpBytes, keys, keysSorted, vals, sortedIdx, (int)n, 0, 63, stream));
    void* tmp = dalloc<char>(tmpBytes);
    CK(cub::DeviceRadixSort::SortPairs(tmp, tmpBytes, keys, keysSorted, vals, sortedIdx, (int)n, 0, 63, stream));
    out.launches += 8;   // CUB onesweep passes (library kernels, not counted as ours elsewhere)

    Lbvh t{};
    const uint32_t ni = n > 1 ? n - 1 : 1;
    float4* leafLo = dalloc<float4>(n); float4* leafHi = dalloc<float4>(n);
    k_gather_leaf_boxes<<<G, B, 0, stream>>>(sortedIdx, n, boxLo, boxHi, leafLo, leafHi); out.launches++;
    dfree(boxLo); dfree(boxHi); boxLo = boxHi = nullptr;      // stream-ordered: reusable by the allocations below
    t.left = dalloc<uint32_t>(ni); t.right = dalloc<uint32_t>(ni); t.count = dalloc<uint32_t>(ni);
    t.lo = dalloc<float4>(ni); t.hi = dalloc<float4>(ni); t.leafLo = leafLo; t.leafHi = leafHi;
    uint32_t rootRef = n > 1 ? 0u : (LEAF_FLAG | 0u);
    if (n > 1 && !in.ploc) {
        t.parent = dalloc<uint32_t>((size_t)ni + n); t.visit = dalloc<uint32_t>(ni);
        CK(cudaMemsetAsync(t.visit, 0, ni * sizeof(uint32_t), stream));
        k_lbvh_topology<<<(n - 1 + B - 1) / B, B, 0, stream>>>(keysSorted, (int)n, t);
        k_lbvh_fit<<<G, B, 0, stream>>>((int)n, t);
        out.launches += 2;
    } else if (n > 1) {
        // cluster arrays: (reference, box) by position; [0] = current, [1] = merge output before compaction
        const int radius = in.plocRadius < 1 ? 1 : (in.plocRadius > PLOC_MAX_RADIUS ? PLOC_MAX_RADIUS : in.plocRadius);
        uint32_t* cid[2] = {dalloc<uint32_t>(n), dalloc<uint32_t>(n)};
        float4* clo[2] = {dalloc<float4>(n), dalloc<float4>(n)}; float4* chi[2] = {dalloc<float4>(n), dalloc<float4>(n)};
        uint32_t* nn = dalloc<uint32_t>(n); uint32_t* flags = dalloc<uint32_t>(n); uint32_t* offsets = dalloc<uint32_t>(n);
        uint32_t* pc = dalloc<uint32_t>(4);                        // [0] node counter, [1] cluster count, [2] root reference
        CK(cudaMemsetAsync(pc, 0, 4 * sizeof(uint32_t), stream));
        size_t scanBytes = 0; CK(cub::DeviceScan::ExclusiveSum(nullptr, scanBytes, flags, offsets, (int)n, stream));
        void* scanTmp = dalloc<char>(scanBytes);
        k_ploc_init<<<G, B, 0, stream>>>(n, cid[0]);
        CK(cudaMemcpyAsync(clo[0], leafLo, (size_t)n * sizeof(float4), cudaMemcpyDeviceToDevice, stream));
        CK(cudaMemcpyAsync(chi[0], leafHi, (size_t)n * sizeof(float4), cudaMemcpyDeviceToDevice, stream));
        out.launches++;
        uint32_t m = n; uint32_t* hostM = in.hostWord;               // the device's pinned read-back word
        while (m > PLOC_FINAL) {
            const uint32_t g = (m + PLOC_BLOCK - 1) / PLOC_BLOCK;
            k_ploc_nn<<<g, PLOC_BLOCK, 0, stream>>>(clo[0], chi[0], m, radius, nn);
            k_ploc_merge<<<g, PLOC_BLOCK, 0, stream>>>(m, nn, cid[0], clo[0], chi[0], t, pc, cid[1], clo[1], chi[1], flags);
            CK(cub::DeviceScan::ExclusiveSum(scanTmp, scanBytes, flags, offsets, (int)m, stream));
            k_ploc_compact<<<g, PLOC_BLOCK, 0, stream>>>(m, flags, offsets, cid[1], clo[1], chi[1], cid[0], clo[0], chi[0], pc + 1);
            out.launches += 5;
            CK(cudaMemcpyAsync(hostM, pc + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
            CK(cudaStreamSynchronize(stream));
            m = *hostM;
        }
        k_ploc_final<<<1, PLOC_FINAL, 0, stream>>>(m, radius, cid[0], clo[0], chi[0], t, pc, pc + 2); out.launches++;
        CK(cudaMemcpyAsync(hostM, pc + 2, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        rootRef = *hostM;
        dfree(cid[0]); dfree(cid[1]); dfree(clo[0]); dfree(clo[1]); dfree(chi[0]); dfree(chi[1]); dfree(nn); dfree(flags); dfree(offsets); dfree(pc); dfree(scanTmp);
    }

    // SAH-optimal collapse: cost tables and decisions per binary node, bottom-up
    float* dpF = nullptr; uint8_t* dpDec = nullptr;
    if (n > 1 && in.dpCollapse) {
        if (!t.parent) t.parent = dalloc<uint32_t>((size_t)ni + n);
        if (!t.visit) t.visit = dalloc<uint32_t>(ni);
        dpF = dalloc<float>(8ull * ni); dpDec = dalloc<uint8_t>(8ull * ni);
        k_bvh2_parents<<<(ni + B - 1) / B, B, 0, stream>>>(t, ni, n);
        CK(cudaMemsetAsync(t.visit, 0, ni * sizeof(uint32_t), stream));
        DpArgs da{t, n, rootRef, dpF, dpDec, 1.0f, in.cTri};
        k_collapse_dp<<<G, B, 0, stream>>>(da);
        out.launches += 2;
    }

    // worst case one BVH8 node per BVH2 internal node; shrunk to the exact size afterwards
    Node8* nodesTmp = dalloc<Node8>((size_t)ni + 1);
    float4* tris = dalloc<float4>(3ull * n); float4* triShade = dalloc<float4>(5ull * n);
    float4* triMotion = in.motions ? dalloc<float4>(3ull * n) : nullptr;
    uint32_t* counters = dalloc<uint32_t>(4);
    uint2* tasksA = dalloc<uint2>(ni + 1); uint2* tasksB = dalloc<uint2>(ni + 1);
    const uint32_t initCounters[4] = {1u, 0u, 0u, 0u};
    CK(cudaMemcpyAsync(counters, initCounters, sizeof(initCounters), cudaMemcpyHostToDevice, stream));
    const uint2 rootTask = make_uint2(rootRef, 0u);
    CK(cudaMemcpyAsync(tasksA, &rootTask, sizeof(rootTask), cudaMemcpyHostToDevice, stream));

    CollapseArgs ca;
    ca.t = t; ca.n = (int)n; ca.sortedIdx = sortedIdx;
    ca.refs = in.refs; ca.geoms = in.geoms; ca.positions = in.positions; ca.indices = in.indices; ca.normals = in.normals; ca.uvs = in.uvs;
    ca.motions = in.motions; ca.triMotion = triMotion;
    ca.nodes = nodesTmp; ca.tris = tris; ca.triShade = triShade; ca.counters = counters; ca.splitLeaves = in.splitLeaves;
    ca.dpF = dpF; ca.dpDec = dpDec;
    uint32_t numTasks = 1; uint2* tin = tasksA; uint2* tout = tasksB;
    uint32_t hostCounters[4];
    while (numTasks) {
        ca.tasksIn = tin; ca.tasksOut = tout; ca.numTasks = numTasks;
        k_collapse<<<(numTasks + 63) / 64, 64, 0, stream>>>(ca);
        out.launches++;
        CK(cudaMemcpyAsync(hostCounters, counters, sizeof(hostCounters), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        numTasks = hostCounters[2];
        const uint32_t zero = 0;
        CK(cudaMemcpyAsync(counters + 2, &zero, sizeof(zero), cudaMemcpyHostToDevice, stream));
        std::swap(tin, tout);
    }
    out.numNodes = hostCounters[0]; out.numTris = hostCounters[1];
    Node8* nodes = dalloc<Node8>(out.numNodes);
    CK(cudaMemcpyAsync(nodes, nodesTmp, (size_t)out.numNodes * sizeof(Node8), cudaMemcpyDeviceToDevice, stream));
    CK(cudaEventRecord(e1, stream));
    CK(cudaStreamSynchronize(stream));
    CK(cudaEventElapsedTime(&out.buildMs, e0, e1));
    out.nodes = nodes; out.tris = tris; out.triShade = triShade; out.triMotion = triMotion;
    scratch.forget(nodes); scratch.forget(tris); scratch.forget(triShade); if (triMotion) scratch.forget(triMotion);      // the results outlive the build (SceneHandle::releaseDevice)

    dfree(nodesTmp); dfree(counters); dfree(tasksA); dfree(tasksB); dfree(dpF); dfree(dpDec);
    dfree(t.left); dfree(t.right); dfree(t.parent); dfree(t.count);
    dfree(t.lo); dfree(t.hi); dfree(t.visit); dfree(leafLo); dfree(leafHi);
    dfree(tmp); dfree(keys); dfree(keysSorted); dfree(vals); dfree(sortedIdx);
    dfree(boxLo); dfree(boxHi); dfree(sceneBounds);
    g_scratch = nullptr;
}

}  // namespace yrt
