// kernels.cu — the wavefront path tracer of device_cuda (hand-written CUDA for sm_100a).
//
// One frame = for each chunk of pixels:  raygen -> [ trace_closest -> shade -> trace_shadow -> resolve ] x maxDepth -> film.
// It restates, stage by stage, the reference's per-pixel recursion
//   IntegratorRenderer::RenderJob::renderTile   devices/device_singleray/renderers/integratorrenderer.cpp:118-185
//   PathTraceIntegrator::Li                     devices/device_singleray/integrators/pathtraceintegrator.cpp:50-217
// as data-parallel stages over SoA queues (16-byte lanes). All kernels are persistent-style:
// the grid is sized to the SM count and strides over a queue whose length lives in device memory,
// so a whole frame is enqueued without host synchronisation.
//
// Determinism: radiance is accumulated per path in the reference's order (emission, then the
// lights in order, bounce by bounce) and per pixel in sample order, so the result does not depend
// on scheduling, queue order or the number of GPUs.
#include "bvh.cuh"
#include "device_internal.hpp"
#include "camera.cuh"
#include "shading.cuh"

namespace yrt {

#define FLAG_IGNORE_VISIBLE_LIGHTS 1u
#define FLAG_UNBENT 2u

__device__ __forceinline__ V3 f4v(float4 a) { return V3(a.x, a.y, a.z); }

// raster row of buffer row b / number of buffer rows  (api/swapchain.h:57-70)
__host__ __device__ __forceinline__ int buffer2raster(int b, int serverID, int serverCount) { return 4 * ((b >> 2) * serverCount + serverID) + (b & 3); }

// ---- per-pixel sample-set choice --------------------------------------------------------------
// integratorrenderer.cpp:132-149: one LCG per 16x16 tile seeded tile_x*91711 + tile_y*81551 + 3433*firstActiveLine,
// one getInt(sets) per rendered pixel in dy,dx order. LCG = common/math/random.h:32-68.
struct DevLcg {
    int seed, state, table[32];
    __device__ void step() { const int k = seed / 127773; seed = 16807 * (seed - k * 127773) - 2836 * k; if (seed < 0) seed += 2147483647; }
    __device__ void init(int s) {
        seed = (s == 0) ? 1 : (s < 0 ? -s : s);
        for (int j = 32 + 7; j >= 0; j--) { step(); if (j < 32) table[j] = seed; }
        state = table[0];
    }
    __device__ int next() { step(); const int j = state / (1 + (2147483647 - 1) / 32); state = table[j]; table[j] = seed; return state; }
};

__global__ void k_pixel_sets(FrameConst fc, uint8_t* __restrict__ pixelSet, int sets) {
    const int numTilesX = (fc.width + 15) / 16, numTilesY = (fc.height + 15) / 16;
    const int tile = blockIdx.x * blockDim.x + threadIdx.x;
    if (tile >= numTilesX * numTilesY) return;
    const int tile_x = (tile % numTilesX) * 16, tile_y = (tile / numTilesX) * 16;
    DevLcg rng; rng.init(tile_x * 91711 + tile_y * 81551 + 3433 * fc.serverID);
    for (int dy = 0; dy < 16; dy++) {
        const int y = tile_y + dy;
        if (y >= fc.height) continue;
        const int row = y >> 2;
        if (((row - fc.serverID) % fc.serverCount) != 0) continue;
        const int by = 4 * ((y >> 2) / fc.serverCount) + (y & 3);
        for (int dx = 0; dx < 16; dx++) {
            const int x = tile_x + dx;
            if (x >= fc.width) continue;
            pixelSet[(size_t)by * fc.width + x] = (uint8_t)(rng.next() % sets);
        }
    }
}
void launch_pixel_sets(const FrameConst& fc, uint8_t* pixelSet, int sets, LaunchCfg lc) {
    const int tiles = ((fc.width + 15) / 16) * ((fc.height + 15) / 16);
    k_pixel_sets<<<(tiles + 63) / 64, 64, 0, lc.stream>>>(fc, pixelSet, sets);
}

// virtual pixel (pixelBegin + p) of a render call -> face, buffer pixel within the face
struct PixelCoord { int face, bx, by; uint32_t bp; };
__device__ __forceinline__ PixelCoord pixel_coord(const FrameConst& fc, uint32_t vp) {
    PixelCoord c;
    c.face = fc.numFaces > 1 ? (int)(vp / fc.pixelsPerFace) : 0;
    c.bp = vp - (uint32_t)c.face * fc.pixelsPerFace;
    c.by = (int)(c.bp / (uint32_t)fc.width); c.bx = (int)(c.bp % (uint32_t)fc.width);
    return c;
}
// path p of a chunk -> (face, buffer pixel, sample); raster coordinates
struct PathCoord { int face, bx, by, x, y, s; uint32_t bp; };
__device__ __forceinline__ PathCoord path_coord(const FrameConst& fc, uint32_t pixelBegin, uint32_t pid) {
    PathCoord c; const uint32_t spp = (uint32_t)fc.integ.spp;
    const PixelCoord q = pixel_coord(fc, pixelBegin + pid / spp);
    c.face = q.face; c.bp = q.bp; c.s = (int)(pid % spp); c.by = q.by; c.bx = q.bx;
    c.x = c.bx; c.y = buffer2raster(c.by, fc.serverID, fc.serverCount);
    return c;
}
__device__ __forceinline__ const float* sample_rec(const FrameConst& fc, const uint8_t* pixelSet, const PathCoord& c) {
    return fc.sampleTable + ((size_t)pixelSet[c.bp] * fc.integ.spp + c.s) * fc.integ.recFloats;
}

__global__ void __launch_bounds__(256) k_raygen(FrameConst fc, const __grid_constant__ FrameCameras cams, WavefrontBuffers wb, uint32_t pixelBegin, uint32_t numPaths) {
    for (uint32_t pid = blockIdx.x * blockDim.x + threadIdx.x; pid < numPaths; pid += gridDim.x * blockDim.x) {
        const PathCoord c = path_coord(fc, pixelBegin, pid);
        const float* rec = sample_rec(fc, wb.pixelSet, c);
        const float fx = (float(c.x) + rec[0]) * fc.rcpWidth, fy = (float(c.y) + rec[1]) * fc.rcpHeight;   // integratorrenderer.cpp:156-157
        V3 org, dir; camera_ray(cams.cam[c.face], fx, fy, rec[3], rec[4], org, dir);
        wb.rayO[pid] = make_float4(org.x, org.y, org.z, 0.f);
        wb.rayD[pid] = make_float4(dir.x, dir.y, dir.z, INFINITY);
        if (fc.scene.hasMotion) wb.hitA[pid] = make_float4(rec[2], 0.f, 0.f, 0.f);     // the ray's time (integratorrenderer.cpp:160) for the traversal
        // throughput (1, unbent), radiance (0) and medium (vacuum) of a fresh path are implied: k_shade(depth 0) does not read them
        wb.queueA[pid] = pid;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { wb.counters[0] = numPaths; wb.counters[1] = 0; wb.counters[2] = 0; wb.counters[4] = 0; wb.counters[5] = 0; }
}
void launch_raygen(const FrameConst& fc, const FrameCameras& cams, const WavefrontBuffers& wb, uint32_t pixelBegin, uint32_t numPixels, LaunchCfg lc) {
    k_raygen<<<lc.blocks, 256, 0, lc.stream>>>(fc, cams, wb, pixelBegin, numPixels * (uint32_t)fc.integ.spp);
}

__global__ void k_export_primary(FrameConst fc, WavefrontBuffers wb, uint32_t numPaths, float* __restrict__ out) {
    for (uint32_t pid = blockIdx.x * blockDim.x + threadIdx.x; pid < numPaths; pid += gridDim.x * blockDim.x) {
        ((float4*)out)[2 * (size_t)pid] = wb.rayO[pid]; ((float4*)out)[2 * (size_t)pid + 1] = wb.rayD[pid];
    }
}
void launch_export_primary(const FrameConst& fc, const WavefrontBuffers& wb, uint32_t pixelBegin, uint32_t numPixels, float* out, LaunchCfg lc) {
    k_export_primary<<<lc.blocks, 256, 0, lc.stream>>>(fc, wb, numPixels * (uint32_t)fc.integ.spp, out);
}

// ---- traversal kernels (persistent warps over the ray queues, bvh.cuh: trace_stream) -----------------------
// counters[4] / [5] / [6]: next unclaimed queue position of the closest-hit / any-hit / user launch
#if YRT_STREAM_HINTS
#define YRT_LD_STREAM(p) __ldcs(p)
#define YRT_ST_STREAM(p, v) __stcs(p, v)
#else
#define YRT_LD_STREAM(p) (*(p))
#define YRT_ST_STREAM(p, v) (*(p) = (v))
#endif
struct ClosestIO {
    WavefrontBuffers wb; const uint32_t* __restrict__ queue;
    __device__ __forceinline__ uint32_t load(uint32_t i, V3& O, V3& D, float& tnear, float& tfar) const {
        const uint32_t pid = YRT_LD_STREAM(&queue[i]);
        const float4 o = YRT_LD_STREAM(&wb.rayO[pid]), d = YRT_LD_STREAM(&wb.rayD[pid]);
        O = f4v(o); D = f4v(d); tnear = o.w; tfar = d.w;
        return pid;
    }
    // the wavefront's hit record is (t, u, v, leaf-order triangle index): k_shade finds everything else in SceneData::triShade
    __device__ __forceinline__ void store_hit(uint32_t pid, float t, float u, float v, uint32_t tri, const float4*) const {
        YRT_ST_STREAM(&wb.hitA[pid], make_float4(t, u, v, __int_as_float(tri == YRT_NO_TRI ? -1 : (int)tri)));
    }
    __device__ __forceinline__ void store_any(uint32_t, bool) const {}
    // motion blur: the path's time travels in the hit record's first lane until the traversal overwrites it (k_raygen / k_shade put it there)
    __device__ __forceinline__ float time(uint32_t pid) const { return wb.hitA[pid].x; }
    __device__ __forceinline__ void overflow() const { atomicOr(&wb.stats[7], 1ull); }     // traversal stack exhausted: the render call throws
};
struct ShadowIO {
    WavefrontBuffers wb;
    __device__ __forceinline__ uint32_t load(uint32_t i, V3& O, V3& D, float& tnear, float& tfar) const {
        const float4 o = YRT_LD_STREAM(&wb.shO[i]), d = YRT_LD_STREAM(&wb.shD[i]);
        O = f4v(o); D = f4v(d); tnear = o.w; tfar = d.w;
        return i;
    }
    __device__ __forceinline__ void store_hit(uint32_t, float, float, float, uint32_t, const float4*) const {}
    __device__ __forceinline__ void store_any(uint32_t i, bool occluded) const { wb.shC[i].w = occluded ? 1.f : 0.f; }
    // motion blur: a shadow ray's time arrives in the lane the occlusion flag is written to
    __device__ __forceinline__ float time(uint32_t i) const { return wb.shC[i].w; }
    __device__ __forceinline__ void overflow() const { atomicOr(&wb.stats[7], 1ull); }
};
struct UserIO {
    const float4* __restrict__ rays; float4* __restrict__ hits; unsigned long long* errFlags;
    __device__ __forceinline__ uint32_t load(uint32_t i, V3& O, V3& D, float& tnear, float& tfar) const {
        const float4 o = __ldg(&rays[2ull * i]), d = __ldg(&rays[2ull * i + 1]);
        O = f4v(o); D = f4v(d); tnear = o.w; tfar = d.w;
        return i;
    }
    // the API's hit record: (t, u, v, geomID | primID, Ng) with Ng = cross(p0 - p1, p2 - p0) unnormalised (RTCRay, rtcore_ray.h:28-55)
    __device__ __forceinline__ void store_hit(uint32_t i, float t, float u, float v, uint32_t tri, const float4* __restrict__ tris) const {
        if (tri == YRT_NO_TRI) { hits[2ull * i] = make_float4(t, 0.f, 0.f, __int_as_float(-1)); hits[2ull * i + 1] = make_float4(__int_as_float(-1), 0.f, 0.f, 0.f); return; }
        const float4* tp = tris + 3ull * tri;
        const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
        const V3 p0(a.x, a.y, a.z), p1(b.x, b.y, b.z), p2(c.x, c.y, c.z), Ng = cross(p0 - p1, p2 - p0);
        hits[2ull * i] = make_float4(t, u, v, a.w);
        hits[2ull * i + 1] = make_float4(b.w, Ng.x, Ng.y, Ng.z);
    }
    __device__ __forceinline__ void store_any(uint32_t i, bool occluded) const {
        float4 a = hits[2ull * i]; a.w = __int_as_float(occluded ? 0 : -1); hits[2ull * i] = a;
    }
    __device__ __forceinline__ float time(uint32_t) const { return 0.f; }     // the API's ray record has no time: shutter open (t = 0)
    __device__ __forceinline__ void overflow() const { if (errFlags) atomicOr(errFlags, 1ull); }
};

// ---- A/B baseline: one ray per thread for the life of the thread (bvh.cuh: trace_ray), cfg trav=0 ------------
template <bool ANY, class IO>
__global__ void __launch_bounds__(YRT_TRACE_THREADS) k_trace_simple(SceneData sc, IO io, const uint32_t* __restrict__ nPtr, uint32_t nImm, unsigned long long* errFlags) {
    const uint32_t n = nPtr ? *nPtr : nImm;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        V3 O, D; float tnear, tfar;
        const uint32_t tag = io.load(i, O, D, tnear, tfar);
        HitRec h; TraceCounters cnt = {0, 0, 0};
        const bool hit = trace_ray<ANY, false>((const uint4*)sc.nodes, sc.tris, sc.numNodes, O, D, tnear, tfar, h, &cnt);
        if (cnt.overflow) atomicOr(errFlags, 1ull);
        if (ANY) io.store_any(tag, hit);
        else io.store_hit(tag, h.t, hit ? h.u : 0.f, hit ? h.v : 0.f, hit ? h.tri : YRT_NO_TRI, sc.tris);
    }
}

#ifndef YRT_CLOSEST_MINBLOCKS
#define YRT_CLOSEST_MINBLOCKS 8
#endif
template <bool COUNT, bool MOTION>
__global__ void __launch_bounds__(YRT_TRACE_THREADS, YRT_CLOSEST_MINBLOCKS) k_trace_closest(SceneData sc, WavefrontBuffers wb, int queueSel) {
    const uint32_t n = wb.counters[queueSel];
    TraceCounters cnt = {0, 0, 0};
    ClosestIO io{wb, queueSel ? wb.queueB : wb.queueA};
    trace_stream<false, COUNT, MOTION>((const uint4*)sc.nodes, sc.tris, sc.triMotion, sc.numNodes, n, &wb.counters[4], io, cnt, TraceTune{sc.tuneRefillMin, sc.tuneTriNum, sc.tuneTriDen, sc.tunePrefetch});
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&wb.stats[0], (unsigned long long)n);
    if (COUNT) { atomicAdd(&wb.stats[2], (unsigned long long)cnt.nodes); atomicAdd(&wb.stats[3], (unsigned long long)cnt.tris); }
}
__global__ void k_count_closest(WavefrontBuffers wb, int queueSel) { atomicAdd(&wb.stats[0], (unsigned long long)wb.counters[queueSel]); }
void launch_trace_closest(const FrameConst& fc, const WavefrontBuffers& wb, int queueSel, LaunchCfg lc) {
    if (fc.scene.tuneSimple && !fc.scene.hasMotion) {
        ClosestIO io{wb, queueSel ? wb.queueB : wb.queueA};
        k_trace_simple<false><<<lc.blocks, YRT_TRACE_THREADS, 0, lc.stream>>>(fc.scene, io, &wb.counters[queueSel], 0u, &wb.stats[7]);
        k_count_closest<<<1, 1, 0, lc.stream>>>(wb, queueSel);
        return;
    }
    if (fc.scene.hasMotion) {
        if (fc.countStats) k_trace_closest<true, true><<<lc.blocks, YRT_TRACE_THREADS, 0, lc.stream>>>(fc.scene, wb, queueSel);
        else k_trace_closest<false, true><<<lc.blocks, YRT_TRACE_THREADS, 0, lc.stream>>>(fc.scene, wb, queueSel);
    } else if (fc.countStats) k_trace_closest<true, false><<<lc.blocks, YRT_TRACE_THREADS, 0, lc.stream>>>(fc.scene, wb, queueSel);
    else k_trace_closest<false, false><<<lc.blocks, YRT_TRACE_THREADS, 0, lc.stream>>>(fc.scene, wb, queueSel);
}

#ifndef YRT_SHADOW_MINBLOCKS
#define YRT_SHADOW_MINBLOCKS 8      // 64 registers, no spills; without the bound ptxas takes 72 since the node loads moved up: 7 CTAs/SM
#endif
#define YRT_SHADOW_BOUNDS __launch_bounds__(YRT_TRACE_THREADS, YRT_SHADOW_MINBLOCKS)
template <bool COUNT, bool MOTION>
__global__ void YRT_SHADOW_BOUNDS k_trace_shadow(SceneData sc, WavefrontBuffers wb) {
    const uint32_t n = wb.counters[2];
    TraceCounters cnt = {0, 0, 0};
    ShadowIO io{wb};
    trace_stream<true, COUNT, MOTION>((const uint4*)sc.nodes, sc.tris, sc.triMotion, sc.numNodes, n, &wb.counters[5], io, cnt, TraceTune{sc.tuneRefillMin, sc.tuneTriNum, sc.tuneTriDen, sc.tunePrefetch});
    if (COUNT) { atomicAdd(&wb.stats[4], (unsigned long long)cnt.nodes); atomicAdd(&wb.stats[5], (unsigned long long)cnt.tris); }
}
void launch_trace_shadow(const FrameConst& fc, const WavefrontBuffers& wb, LaunchCfg lc) {
    if (fc.scene.tuneSimple && !fc.scene.hasMotion) { ShadowIO io{wb}; k_trace_simple<true><<<lc.blocks, YRT_TRACE_THREADS, 0, lc.stream>>>(fc.scene, io, &wb.counters[2], 0u, &wb.stats[7]); return; }
    if (fc.scene.hasMotion) {
        if (fc.countStats) k_trace_shadow<true, true><<<lc.blocks, YRT_TRACE_THREADS, 0, lc.stream>>>(fc.scene, wb);
        else k_trace_shadow<false, true><<<lc.blocks, YRT_TRACE_THREADS, 0, lc.stream>>>(fc.scene, wb);
    } else if (fc.countStats) k_trace_shadow<true, false><<<lc.blocks, YRT_TRACE_THREADS, 0, lc.stream>>>(fc.scene, wb);
    else k_trace_shadow<false, false><<<lc.blocks, YRT_TRACE_THREADS, 0, lc.stream>>>(fc.scene, wb);
}

#ifndef YRT_USER_MINBLOCKS
#define YRT_USER_MINBLOCKS 8      // 64 registers: 8 CTAs/SM like the wavefront kernels (r2 A/B on the config-5 soups: +1 % at 1e6 / 1e7 triangles, +4 % at 1e8 over 7 CTAs/SM)
#endif
#define YRT_USER_BOUNDS __launch_bounds__(YRT_TRACE_THREADS, YRT_USER_MINBLOCKS)
template <bool ANY, bool COUNT>
__global__ void YRT_USER_BOUNDS k_trace_user(SceneData sc, const float4* __restrict__ rays, float4* __restrict__ hits, uint32_t n,
                                                                  uint32_t* workCounter, unsigned long long* stats) {
    TraceCounters cnt = {0, 0, 0};
    UserIO io{rays, hits, stats ? stats + 7 : nullptr};
    trace_stream<ANY, COUNT, false>((const uint4*)sc.nodes, sc.tris, nullptr, sc.numNodes, n, workCounter, io, cnt, TraceTune{sc.tuneRefillMin, sc.tuneTriNum, sc.tuneTriDen, sc.tunePrefetch});
    if (COUNT && stats) { atomicAdd(&stats[2], (unsigned long long)cnt.nodes); atomicAdd(&stats[3], (unsigned long long)cnt.tris); }
}
void launch_trace_user(const SceneData& sc, const float* rays, float* hits, size_t n, int closest, int countStats,
                       unsigned long long* stats, uint32_t* workCounter, LaunchCfg lc) {
    const float4* r = (const float4*)rays; float4* h = (float4*)hits; const uint32_t m = (uint32_t)n;
    if (sc.tuneSimple) {
        UserIO io{r, h, stats ? stats + 7 : nullptr};
        if (closest) k_trace_simple<false><<<lc.blocks, YRT_TRACE_THREADS, 0, lc.stream>>>(sc, io, nullptr, m, stats + 7);
        else k_trace_simple<true><<<lc.blocks, YRT_TRACE_THREADS, 0, lc.stream>>>(sc, io, nullptr, m, stats + 7);
        return;
    }
    if (closest) {
        if (countStats) k_trace_user<false, true><<<lc.blocks, YRT_TRACE_THREADS, 0, lc.stream>>>(sc, r, h, m, workCounter, stats);
        else k_trace_user<false, false><<<lc.blocks, YRT_TRACE_THREADS, 0, lc.stream>>>(sc, r, h, m, workCounter, stats);
    } else {
        if (countStats) k_trace_user<true, true><<<lc.blocks, YRT_TRACE_THREADS, 0, lc.stream>>>(sc, r, h, m, workCounter, stats);
        else k_trace_user<true, false><<<lc.blocks, YRT_TRACE_THREADS, 0, lc.stream>>>(sc, r, h, m, workCounter, stats);
    }
}

// ---- shading -------------------------------------------------------------------------------------
// Warp-aggregated append. Must be reached by all 32 lanes of the warp (the shading loop below is
// written so that every lane executes every iteration); returns the first slot of this lane's
// `count` consecutive elements (count may be 0).
__device__ __forceinline__ uint32_t warp_alloc(uint32_t* counter, uint32_t count) {
    const uint32_t lane = threadIdx.x & 31;
    uint32_t incl = count;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (uint32_t)o) incl += v; }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    uint32_t base = 0;
    if (lane == 31 && total) base = atomicAdd(counter, total);
    base = __shfl_sync(0xffffffffu, base, 31);
    return base + incl - count;
}

#ifndef YRT_SHADE_THREADS
#define YRT_SHADE_THREADS 128
#endif
// Cache policy of the shading kernel: the path state (ray, hit, throughput, queues, shadow-ray slots) is read once and written once per
// bounce — 6.7 GB per launch on C4 — while the shading records (96 MB) and the texture set (273 MB) are re-used across paths; with default
// caching the stream evicts them from L2 (ncu r2: L2 hit 53 %). The streams use ld.global.cs / st.global.cs (evict-first).
#ifndef YRT_SHADE_STREAM
#define YRT_SHADE_STREAM 1
#endif
#if YRT_SHADE_STREAM
#define SH_LD(p) __ldcs(p)
#define SH_ST(p, v) __stcs(p, v)
#else
#define SH_LD(p) (*(p))
#define SH_ST(p, v) (*(p) = (v))
#endif
// The shading kernel is bound by instruction fetch; barriers keep the warps of a CTA in the same code region so that they share
// fetched lines (measured at 1024^2: -4 % shade time on the C3 stand-in, +4 % on C2; one barrier per iteration is the
// default, YRT_SHADE_SYNC=2 adds two more inside the iteration). Every thread executes every iteration, so the barriers are uniform.
// Measured and left off (round 1): regrouping the CTA's queue entries by shading class before shading (-DYRT_SHADE_REGROUP=1) costs three
// barriers, two dependent loads and the warp-level coalescing of the path state: shade time +8 % on the C3 stand-in, +7 % on C2.
#ifndef YRT_SHADE_REGROUP
#define YRT_SHADE_REGROUP 0
#endif
#ifndef YRT_SHADE_PREFETCH
#define YRT_SHADE_PREFETCH 0
#endif
#ifndef YRT_SHADE_SYNC
#define YRT_SHADE_SYNC 1
#endif
#if YRT_SHADE_SYNC
#define SHADE_BARRIER() __syncthreads()
#else
#define SHADE_BARRIER()
#endif
#if YRT_SHADE_SYNC >= 2
#define SHADE_BARRIER2() __syncthreads()
#else
#define SHADE_BARRIER2()
#endif
template <bool EXT>
__global__ void __launch_bounds__(YRT_SHADE_THREADS, EXT ? 4 : YRT_SHADE_MINBLOCKS) k_shade(FrameConst fc, WavefrontBuffers wb, int queueSel, uint32_t pixelBegin, int depth) {
    __shared__ float smLobes[YRT_MAX_LOBES * LobesT<EXT>::WORDS * YRT_SHADE_THREADS], smCand[YRT_MAX_LOBES * YRT_CAND_WORDS * YRT_SHADE_THREADS];
#if YRT_SHADE_REGROUP
    __shared__ uint32_t smHist[16], smPerm[YRT_SHADE_THREADS];
#endif
    const uint32_t* __restrict__ queue = queueSel ? wb.queueB : wb.queueA;
    uint32_t* __restrict__ nextQueue = queueSel ? wb.queueA : wb.queueB;
    const uint32_t n = wb.counters[queueSel];
    const SceneData& sc = fc.scene; const IntegratorData& ig = fc.integ;
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t nIter = (n + stride - 1) / stride;
    uint32_t shadowRays = 0;
    for (uint32_t it = 0; it < nIter; it++) {
        uint32_t i = it * stride + blockIdx.x * blockDim.x + threadIdx.x;
        bool valid = i < n;
        SHADE_BARRIER();
        uint32_t pid = 0;
        if (valid) pid = SH_LD(&queue[i]);
#if YRT_SHADE_REGROUP
        // Regroup the CTA's entries by shading class (counting sort over <= 16 classes in shared memory) so that the lanes of a warp
        // run the same material code: bounce rays hit unrelated surfaces, and the heaviest paths (glossy lobes, texture fetches) ran
        // with ~10 of 32 lanes active. The permutation stays inside the CTA's window of the queue, so the path-state accesses keep
        // their cache-line locality. Primary hits (bounce 0) are coherent already.
        if (depth > 0) {
            uint32_t cls = 15u;                                  // entries past the end of the queue sort last
            if (valid) { const int tr = __float_as_int(wb.hitA[pid].w); cls = tr < 0 ? 0u : (uint32_t)sc.geoms[__float_as_uint(sc.triShade[5ull * tr].w) >> 2].shadeClass; }
            if (threadIdx.x < 16) smHist[threadIdx.x] = 0u;
            __syncthreads();
            // warp-aggregated histogram: one shared-memory atomic per (warp, class) instead of one per thread
            const unsigned peers = __match_any_sync(0xffffffffu, cls);
            const unsigned lane = threadIdx.x & 31u;
            const int leader = __ffs(peers) - 1;
            uint32_t wbase = 0;
            if ((int)lane == leader) wbase = atomicAdd(&smHist[cls], (uint32_t)__popc(peers));
            const uint32_t rank = __shfl_sync(0xffffffffu, wbase, leader) + (uint32_t)__popc(peers & ((1u << lane) - 1u));
            __syncthreads();
            uint32_t base = 0;
            for (uint32_t c = 0; c < cls; c++) base += smHist[c];
            smPerm[base + rank] = valid ? pid : 0xffffffffu;
            __syncthreads();
            pid = smPerm[threadIdx.x];
            valid = pid != 0xffffffffu;
        }
#endif
        bool alive = false, needLights = false;
        uint32_t flags = 0;
        DG dg; LobesT<EXT> lobes; lobes.s = &smLobes[threadIdx.x]; lobes.cand = &smCand[threadIdx.x]; lobes.stride = YRT_SHADE_THREADS; lobes.n = 0; V3 wo(0.f); Col thr(0.f); const float* rec = nullptr; float fx = 0, fy = 0;
        float4 d4 = make_float4(0, 0, 0, 0), m4 = make_float4(1, 1, 1, 1); float hitT = 0.f;
        if (valid) {
            const float4 o4 = SH_LD(&wb.rayO[pid]); d4 = SH_LD(&wb.rayD[pid]);
            const float4 hA = SH_LD(&wb.hitA[pid]);
            // radiance so far: read (and written back) only by the vertices that add to it — misses that see the environment and hits
            // on visible emitters; k_resolve adds the light samples. Bounce 0 initialises it.
            Col L(0.f); bool Ltouched = depth == 0;
            auto loadL = [&]() { if (!Ltouched) { const float4 L4 = wb.Lacc[pid]; L = Col(L4.x, L4.y, L4.z); Ltouched = true; } };
            if (depth == 0) { thr = Col(1.f); flags = FLAG_UNBENT; }            // LightPath(ray): pathtraceintegrator.h:41-43
            else {
                const float4 t4 = SH_LD(&wb.thr[pid]); thr = Col(t4.x, t4.y, t4.z); flags = __float_as_uint(t4.w) >> 16;
                if (sc.hasMedia) m4 = wb.medium[pid];                           // only Dielectric materials change the medium
            }
            hitT = hA.x;
            const PathCoord pc = path_coord(fc, pixelBegin, pid);
            rec = sample_rec(fc, wb.pixelSet, pc);
            fx = (float(pc.x) + rec[0]) * fc.rcpWidth; fy = (float(pc.y) + rec[1]) * fc.rcpHeight;
            const V3 org = f4v(o4), dir = f4v(d4);
            wo = -dir;
            const int triIdx = __float_as_int(hA.w);
            if (triIdx < 0) {
                // environment shading (pathtraceintegrator.cpp:79-92)
                if (ig.backplateTex >= 0 && (flags & FLAG_UNBENT)) {
                    const TextureRec& bp = sc.textures[ig.backplateTex];
                    const int x = iclamp(int(fx * bp.width), 0, bp.width - 1), y = iclamp(int(fy * bp.height), 0, bp.height - 1);
                    const Col4 c = texel(bp, x, y);
                    loadL(); L += thr * Col(c.r, c.g, c.b);
                } else if (!(flags & FLAG_IGNORE_VISIBLE_LIGHTS) && sc.numEnvLights > 0) {
                    loadL();
                    for (int e = 0; e < sc.numEnvLights; e++) L += thr * env_Le(sc, sc.lights[sc.envLightIdx[e]], wo);
                }
            } else {
                post_intersect<EXT>(sc, org, dir, hA.x, hA.y, hA.z, triIdx, sc.hasMotion ? rec[2] : 0.f, dg);
                bool backfacing = false;
                if (dot(dg.Ng, dir) > 0.f) { backfacing = true; dg.Ng = -dg.Ng; dg.Ns = -dg.Ns; }   // :95-98
                if (dg.material >= 0) material_shade<EXT>(sc, sc.materials[dg.material], dg, Col(m4.x, m4.y, m4.z), m4.w, lobes);
                if (!(flags & FLAG_IGNORE_VISIBLE_LIGHTS) && dg.areaLight >= 0 && !backfacing) { loadL(); L += thr * sc.lights[dg.areaLight].L; }  // :114-115
                for (int k = 0; k < lobes.n; k++) needLights |= (lobes.type(k) & BR_DIFFUSE) != 0;
                alive = true;
            }
            if (Ltouched) wb.Lacc[pid] = make_float4(L.x, L.y, L.z, 0.f);
        }
        // ---- direct lighting: one slot per light, in light order (pathtraceintegrator.cpp:124-166).
        // Skipped lights leave an invalid ray (tfar < tnear) so the per-path span stays contiguous and the
        // resolve kernel can add the contributions in the reference's order.
        SHADE_BARRIER2();
        const uint32_t nl = (valid && alive && needLights) ? (uint32_t)sc.numLights : 0u;
        const uint32_t base = warp_alloc(&wb.counters[2], nl);
        for (uint32_t li = 0; li < nl; li++) {
            const uint32_t slot = base + li;
            if (slot >= wb.shadowCapacity) break;
            const LightRec& lt = sc.lights[li];
            bool ok = (lt.illumMask & dg.illumMask) != 0;
            LightSampleD ls; ls.wi = V3(0.f); ls.pdf = 0.f; ls.tMax = 0.f; Col lL(0.f), brdf(0.f);
            if (ok) {
                if (lt.precomputedId >= 0) {
                    const float* p = rec + ig.offLight + 8 * lt.precomputedId;
                    ls.wi = V3(p[0], p[1], p[2]); ls.pdf = p[3]; lL = Col(p[4], p[5], p[6]);
                } else lL = light_sample(lt, dg, ls, rec[ig.off2D + 2 * ig.lightSampleID], rec[ig.off2D + 2 * ig.lightSampleID + 1]);
                ok = !(lL == Col(0.f) || ls.pdf == 0.f);
            }
            if (ok) { brdf = lobes_eval(lobes, wo, dg, ls.wi, BR_DIFFUSE); ok = !(brdf == Col(0.f)); }
            if (!ok) {
                SH_ST(&wb.shO[slot], make_float4(0.f, 0.f, 0.f, 1.f)); SH_ST(&wb.shD[slot], make_float4(0.f, 0.f, 1.f, -1.f));
                SH_ST(&wb.shC[slot], make_float4(0.f, 0.f, 0.f, 1.f));   // empty interval: retired as "not occluded" -> w = 0 with a zero contribution
                continue;
            }
            // dome-light shadow-ray length (pathtraceintegrator.cpp:147-158) with the stated pins P1/P2
            // P2: tMaxShadowRay == +inf makes the reference's expression inf - inf = NaN -> tfar = NaN -> never occluded
            float tMax;
            if (isinf(ig.tMaxShadowRay)) tMax = __int_as_float(0x7fc00000);
            else {
                const float r = hash_unit(hash4(__float_as_uint(fx), __float_as_uint(fy), (uint32_t)depth, li));
                const float jit = 2.f * ig.tMaxShadowRay * ig.tMaxShadowJitter * r - ig.tMaxShadowRay * ig.tMaxShadowJitter;
                tMax = ig.tMaxShadowRay + jit;
                const float dp = dot(ls.wi, ig.up);
                if (dp <= 0.f) tMax += ig.tMaxShadowRay * 100.f * smoothstepf(0.f, 1.f, fabsf(dp));
            }
            const float eps = dg.error * ig.epsilon;
            const Col contrib = thr * lL * brdf * rcpf(ls.pdf);
            SH_ST(&wb.shO[slot], make_float4(dg.P.x, dg.P.y, dg.P.z, eps));
            SH_ST(&wb.shD[slot], make_float4(ls.wi.x, ls.wi.y, ls.wi.z, tMax - eps));
            SH_ST(&wb.shC[slot], make_float4(contrib.x, contrib.y, contrib.z, sc.hasMotion ? rec[2] : 1.f));   // w: the ray's time in, the occlusion flag out
            shadowRays++;
        }
        if (nl) SH_ST(&wb.shadowPid[base / nl], pid);                // slots are claimed in groups of numLights: base is a multiple of nl

        // ---- path continuation (pathtraceintegrator.cpp:169-213)
        SHADE_BARRIER2();
        bool cont = false;
        if (valid && alive) {
            cont = depth < ig.maxDepth - 1;
            if (cont && depth >= ig.rrDepth - 1) {
                const float q = rmin(reduce_max(thr) * 1.f * 1.f, .95f);       // eta stays 1: CompositedBRDF::sample drops Sample::eta
                if (rec[ig.off1D + ig.firstScatterTypeSampleID + depth] >= q) cont = false;
            }
            if (cont) {
                Sample3 smp; uint32_t type;
                const float* s2 = rec + ig.off2D + 2 * (ig.firstScatterSampleID + depth);
                const float ss = rec[ig.off1D + ig.firstScatterTypeSampleID + depth];
                Col c = lobes_sample(lobes, wo, dg, smp, type, s2[0], s2[1], ss, BR_ALL);
                if (c == Col(0.f) || smp.pdf <= 0.f) cont = false;
                else {
                    const Col tr(m4.x, m4.y, m4.z);
                    if (tr != Col(1.f)) c *= Col(YRT_POWF(tr.x, hitT), YRT_POWF(tr.y, hitT), YRT_POWF(tr.z, hitT));   // :198-201
                    if (type & BR_TRANSMISSION) {
                        const MaterialRec& m = sc.materials[dg.material];
                        if (m.isMediaInterface) {   // Material::nextMedium  materials/material.h:49-52
                            const bool inside = (tr == m.tInside) && (m4.w == m.etaInside);
                            m4 = inside ? make_float4(m.tOutside.x, m.tOutside.y, m.tOutside.z, m.etaOutside)
                                        : make_float4(m.tInside.x, m.tInside.y, m.tInside.z, m.etaInside);
                        }
                    }
                    const Col nthr = thr * c * rcpf(smp.pdf);
                    uint32_t nflags = 0;
                    if (type & BR_DIFFUSE) nflags |= FLAG_IGNORE_VISIBLE_LIGHTS;
                    if ((flags & FLAG_UNBENT) && smp.v == f4v(d4)) nflags |= FLAG_UNBENT;
                    SH_ST(&wb.rayO[pid], make_float4(dg.P.x, dg.P.y, dg.P.z, dg.error * ig.epsilon));
                    SH_ST(&wb.rayD[pid], make_float4(smp.v.x, smp.v.y, smp.v.z, INFINITY));
                    SH_ST(&wb.thr[pid], make_float4(nthr.x, nthr.y, nthr.z, __uint_as_float(nflags << 16)));
                    if (sc.hasMedia) wb.medium[pid] = m4;
                    if (sc.hasMotion) wb.hitA[pid] = make_float4(rec[2], 0.f, 0.f, 0.f);      // lastRay.time rides on (pathtraceintegrator.cpp:210)
                    if (reduce_max(nthr) < ig.minContribution) cont = false;     // loop-top test of the next bounce (:66)
                }
            }
        }
        const uint32_t slot = warp_alloc(&wb.counters[queueSel ^ 1], cont ? 1u : 0u);
        if (cont) SH_ST(&nextQueue[slot], pid);
    }
    // rtcOccluded-equivalent ray count (pathtraceintegrator.cpp:161): valid shadow rays only
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) shadowRays += __shfl_xor_sync(0xffffffffu, shadowRays, o);
    if ((threadIdx.x & 31) == 0 && shadowRays) atomicAdd(&wb.stats[1], (unsigned long long)shadowRays);
}
void launch_shade(const FrameConst& fc, const WavefrontBuffers& wb, int queueSel, uint32_t pixelBegin, int depth, LaunchCfg lc) {
    // scenes with Plastic / Metal / BrushedMetal / MetallicPaint / Velvet take the EXT instantiation (wider lobe records, tangents)
    if (fc.scene.hasExtMaterials) k_shade<true><<<lc.blocks, YRT_SHADE_THREADS, 0, lc.stream>>>(fc, wb, queueSel, pixelBegin, depth);
    else k_shade<false><<<lc.blocks, YRT_SHADE_THREADS, 0, lc.stream>>>(fc, wb, queueSel, pixelBegin, depth);
}

// adds the unoccluded light contributions of this bounce in light order (one thread per path that sampled lights: its
// numLights consecutive shadow slots), then the counters are reset
__global__ void __launch_bounds__(256) k_resolve(WavefrontBuffers wb, uint32_t numLights) {
    const uint32_t n = wb.counters[2] / numLights;
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < n; g += gridDim.x * blockDim.x) {
        const uint32_t pid = wb.shadowPid[g];
        float4 L4 = wb.Lacc[pid]; Col L(L4.x, L4.y, L4.z);
        for (uint32_t k = 0; k < numLights; k++) {
            const uint32_t slot = g * numLights + k;
            if (slot >= wb.shadowCapacity) break;
            const float4 c = wb.shC[slot];
            if (c.w == 0.f) L += Col(c.x, c.y, c.z);
        }
        wb.Lacc[pid] = make_float4(L.x, L.y, L.z, 0.f);
    }
}
__global__ void k_reset_counters(WavefrontBuffers wb, int queueSel) { wb.counters[queueSel] = 0; wb.counters[2] = 0; wb.counters[4] = 0; wb.counters[5] = 0; }
void launch_resolve(const FrameConst& fc, const WavefrontBuffers& wb, int queueSel, LaunchCfg lc) {
    if (fc.scene.numLights > 0) k_resolve<<<lc.blocks, 256, 0, lc.stream>>>(wb, (uint32_t)fc.scene.numLights);
    k_reset_counters<<<1, 1, 0, lc.stream>>>(wb, queueSel);
}

// ---- film: per-pixel sample sum, accumulation buffer, tone mapping, packing -------------------
// SwapChain::update -> AccuBuffer::update api/framebuffer.h:289-304; DefaultToneMapper::eval
// tonemappers/defaulttonemapper.h:38-51; FrameBufferRGB8/RGBA8/RGBFloat32::set api/framebuffer.h:127-129,171-178,220-226
__device__ __forceinline__ void pack_pixel(const FilmParams& fp, void* fb, int bx, int by, Col c) {
    if (fp.format == 0) {
        float* o = (float*)((char*)fb + (size_t)by * fp.fbStrideBytes) + 3 * bx;
        o[0] = c.x; o[1] = c.y; o[2] = c.z;
    } else {
        const int bpp = fp.format == 1 ? 4 : 3;
        unsigned char* o = (unsigned char*)fb + (size_t)by * fp.fbStrideBytes + bpp * bx;
        o[0] = (unsigned char)rclamp(c.x * 255.0f, 0.0f, 255.0f);
        o[1] = (unsigned char)rclamp(c.y * 255.0f, 0.0f, 255.0f);
        o[2] = (unsigned char)rclamp(c.z * 255.0f, 0.0f, 255.0f);
        if (bpp == 4) o[3] = 0;
    }
}

__global__ void __launch_bounds__(256) k_film(FrameConst fc, const __grid_constant__ FilmParams fp, WavefrontBuffers wb, uint32_t pixelBegin, uint32_t numPixels) {
    const int spp = fc.integ.spp;
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < numPixels; p += gridDim.x * blockDim.x) {
        Col L(0.f);
        for (int s = 0; s < spp; s++) { const float4 l = __ldcs(&wb.Lacc[(size_t)p * spp + s]); L += Col(l.x, l.y, l.z); }
        const PixelCoord pc = pixel_coord(fc, pixelBegin + p);
        const uint32_t bp = pc.bp;
        const int bx = pc.bx, by = pc.by;
        const int y = buffer2raster(by, fc.serverID, fc.serverCount);
        const float weight = (float)spp;
        float4* __restrict__ accum = fp.face[pc.face].accum;
        Col L0;
        if (fp.accumulate) {
            const float4 cur = accum[bp];
            const float4 next = make_float4(cur.x + L.x, cur.y + L.y, cur.z + L.z, cur.w + weight);
            accum[bp] = next;
            L0 = Col(next.x, next.y, next.z) * rcpf(next.w);
        } else {
            accum[bp] = make_float4(L.x, L.y, L.z, weight);
            L0 = L * rcpf(weight);
        }
        Col c = L0;
        if (fp.gamma != 1.0f) c = Col(powf(c.x, fp.rcpGamma), powf(c.y, fp.rcpGamma), powf(c.z, fp.rcpGamma));
        if (fp.vignetting) {
            const float ddx = (float(bx) - 0.5f * float(fc.width)) * rcpf(0.5f * float(fc.width));
            const float ddy = (float(y) - 0.5f * float(fc.height)) * rcpf(0.5f * float(fc.width));
            const float d = sqrtf(ddx * ddx + ddy * ddy);
            c *= powf(cosf(d * 0.5f), 3.0f);
        }
        pack_pixel(fp, fp.face[pc.face].fb, bx, by, c);
    }
}
void launch_film(const FrameConst& fc, const WavefrontBuffers& wb, const FilmParams& fp, uint32_t pixelBegin, uint32_t numPixels, LaunchCfg lc) {
    k_film<<<lc.blocks, 256, 0, lc.stream>>>(fc, fp, wb, pixelBegin, numPixels);
}


// ---- debug renderer (renderers/debugrenderer.cpp:66-148, maxDepth 1) -----------------------------
// Primary rays through the pixel corners with the fixed lens sample (0.5, 0.5); colour = hash of geomID + primID,
// written without tone mapping or accumulation.
__global__ void __launch_bounds__(128) k_debug(FrameConst fc, const __grid_constant__ FrameCameras cams, const __grid_constant__ FilmParams fp, WavefrontBuffers wb, uint32_t numPixels) {
    const SceneData& sc = fc.scene;
    unsigned long long rays = 0;
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < numPixels; p += gridDim.x * blockDim.x) {
        const PixelCoord pc = pixel_coord(fc, p);
        const int bx = pc.bx, by = pc.by;
        const int y = buffer2raster(by, fc.serverID, fc.serverCount);
        const float fx = float(bx) * fc.rcpWidth, fy = float(y) * fc.rcpHeight;
        for (int i = 0; i < fc.integ.spp; i++) {
            V3 org, dir; camera_ray(cams.cam[pc.face], fx, fy, 0.5f, 0.5f, org, dir);
            HitRec h; TraceCounters cnt = {0, 0, 0};
            if (fc.integ.maxDepth > 0) { trace_ray<false, false>((const uint4*)sc.nodes, sc.tris, sc.numNodes, org, dir, 0.f, INFINITY, h, &cnt); rays++; if (cnt.overflow) atomicOr(&wb.stats[7], 1ull); }
            else { h.geomID = -1; h.primID = -1; }
            Col c(1.f);
            if (h.geomID >= 0) {
                const int id = h.geomID + h.primID;
                c = Col(float((3434553u * ((unsigned)(id + 3243))) % 255u) / 255.0f,
                        float((7342453u * ((unsigned)(id + 8237))) % 255u) / 255.0f,
                        float((9234454u * ((unsigned)(id + 2343))) % 255u) / 255.0f);
            }
            pack_pixel(fp, fp.face[pc.face].fb, bx, by, c);
        }
    }
    if (rays) atomicAdd(&wb.stats[0], rays);
}
// The same renderer with maxDepth > 1 (debugrenderer.cpp:104-121): after every hit but the last a diffuse bounce leaves the hit point,
// ray = (org + 0.999 t dir, cosineSampleHemisphere(u, v, Nf), tnear 4 ulp), with u, v from ONE Random per 16x16 tile seeded tile * 1024 and
// drawn in pixel / sample / bounce order — so a pixel's numbers depend on how many bounces the pixels before it in its tile took. One
// thread walks one tile in that order (a debugging aid: correctness first). u is drawn before v (pin P6: the reference leaves the order
// of its two getFloat() arguments to the compiler). The pixel shows the ID colour of the last hit, or white once a ray escapes.
__global__ void __launch_bounds__(64) k_debug_bounces(FrameConst fc, const __grid_constant__ FrameCameras cams, const __grid_constant__ FilmParams fp, WavefrontBuffers wb) {
    const SceneData& sc = fc.scene;
    const int numTilesX = (fc.width + 15) / 16, numTilesY = (fc.height + 15) / 16;
    const int job = blockIdx.x * blockDim.x + threadIdx.x;
    if (job >= numTilesX * numTilesY * fc.numFaces) return;
    const int face = job / (numTilesX * numTilesY), tile = job % (numTilesX * numTilesY);
    const int x0 = (tile % numTilesX) * 16, y0 = (tile / numTilesX) * 16;
    DevLcg rng; rng.init(tile * 1024);
    unsigned long long rays = 0; bool overflow = false;
    for (int dy = 0; dy < 16; dy++) {
        const int y = y0 + dy;
        if (y >= fc.height || (((y >> 2) - fc.serverID) % fc.serverCount) != 0) continue;
        const int by = 4 * ((y >> 2) / fc.serverCount) + (y & 3);
        const float fy = float(y) * fc.rcpHeight;
        for (int dx = 0; dx < 16; dx++) {
            const int x = x0 + dx;
            if (x >= fc.width) continue;
            const float fx = float(x) * fc.rcpWidth;
            for (int i = 0; i < fc.integ.spp; i++) {
                V3 org, dir; camera_ray(cams.cam[face], fx, fy, 0.5f, 0.5f, org, dir);
                float tnear = 0.f; HitRec h; h.geomID = -1; h.primID = -1;
                for (int depth = 0; depth < fc.integ.maxDepth; depth++) {
                    TraceCounters cnt = {0, 0, 0};
                    trace_ray<false, false>((const uint4*)sc.nodes, sc.tris, sc.numNodes, org, dir, tnear, INFINITY, h, &cnt); rays++;
                    overflow |= cnt.overflow != 0;
                    if (h.geomID < 0) break;
                    if (depth + 1 < fc.integ.maxDepth) {
                        V3 Nf = normalize(h.Ng);
                        if (dot(-dir, Nf) < 0.f) Nf = -Nf;
                        const float u = rmin(float(rng.next()) / 2147483647.0f, 1.0f - YRT_ULP), v = rmin(float(rng.next()) / 2147483647.0f, 1.0f - YRT_ULP);
                        const V3 nd = cosine_sample_hemisphere(u, v, Nf).v;
                        org = org + 0.999f * h.t * dir; dir = nd; tnear = 4.0f * YRT_ULP;
                    }
                }
                Col c(1.f);
                if (h.geomID >= 0) {
                    const int id = h.geomID + h.primID;
                    c = Col(float((3434553u * ((unsigned)(id + 3243))) % 255u) / 255.0f,
                            float((7342453u * ((unsigned)(id + 8237))) % 255u) / 255.0f,
                            float((9234454u * ((unsigned)(id + 2343))) % 255u) / 255.0f);
                }
                pack_pixel(fp, fp.face[face].fb, x, by, c);
            }
        }
    }
    if (rays) atomicAdd(&wb.stats[0], rays);
    if (overflow) atomicOr(&wb.stats[7], 1ull);
}
void launch_debug(const FrameConst& fc, const FrameCameras& cams, const WavefrontBuffers& wb, const FilmParams& fp, uint32_t numPixels, LaunchCfg lc) {
    if (fc.integ.maxDepth > 1) {
        const int jobs = ((fc.width + 15) / 16) * ((fc.height + 15) / 16) * fc.numFaces;
        k_debug_bounces<<<(jobs + 63) / 64, 64, 0, lc.stream>>>(fc, cams, fp, wb);
        return;
    }
    k_debug<<<lc.blocks, 128, 0, lc.stream>>>(fc, cams, fp, wb, numPixels);
}

}  // namespace yrt
