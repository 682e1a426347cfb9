// device_internal.hpp — interfaces between the host object model (device_host.cu), the BVH
// builder (bvh_build.cu) and the wavefront kernels (kernels.cu). Internal to libyrt_device_cuda.so.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "common.cuh"

namespace yrt {

// ---- BVH build -------------------------------------------------------------------------------
struct BvhBuildInput {
    const uint2* refs;            // device: (geomID, primID) of every valid triangle
    uint32_t numRefs;
    const GeomRec* geoms;         // device
    const float4* positions;      // device
    const int4* indices;          // device
    const float4* normals;        // device (shading normals, GeomRec::nrmBase)
    const float2* uvs;            // device (texture coordinates, GeomRec::uvBase)
    const float4* motions;        // device (motion vectors, GeomRec::motBase), NULL when the scene has no moving mesh
    uint32_t* hostWord;           // pinned host word for the builder's read-backs (owned by the device)
    int ploc = 1;                 // 1: PLOC hierarchy (default), 0: Karras LBVH
    int splitLeaves = 1;          // BVH8 collapse: use free child slots to split leaf children of 2-3 triangles
    int plocRadius = 8;           // PLOC neighbour search radius (positions to either side)
    int dpCollapse = 1;           // BVH8 collapse: 1 = SAH-optimal cut by dynamic programming (k_collapse_dp), 0 = greedy largest-area opening
    float cTri = 0.6f;            // DP cost of one triangle test relative to one node visit
};
struct BvhResult {
    void* nodes; float4* tris; float4* triShade; float4* triMotion; uint32_t numNodes, numTris; float buildMs; uint32_t launches;
};
void build_bvh(const BvhBuildInput& in, BvhResult& out, cudaStream_t stream);

// ---- wavefront state (SoA, 16-byte vector lanes) ---------------------------------------------
struct WavefrontBuffers {
    uint32_t capacity;            // paths per chunk
    uint32_t shadowCapacity;      // shadow rays per chunk and bounce
    float4* rayO; float4* rayD;   // org.xyz|tnear, dir.xyz|tfar
    float4* hitA;                 // t,u,v | leaf-order triangle index (-1 = miss)
    float4* thr;                  // throughput.rgb | (depth | flags << 16)
    float4* Lacc;                 // radiance accumulated along the path
    float4* medium;               // transmission.rgb | eta
    uint32_t* shadowPid;          // per group of numLights shadow slots: the path that owns it
    float4* shO; float4* shD;     // shadow ray queue
    float4* shC;                  // contribution.rgb | occluded flag (written by the any-hit kernel)
    uint32_t* queueA; uint32_t* queueB;
    uint32_t* counters;           // [0] |queueA|, [1] |queueB|, [2] shadow rays this bounce, [4]/[5]/[6] next unclaimed ray of the closest / any-hit / user traversal launch
    unsigned long long* stats;    // [0] closest rays, [1] shadow rays, [2] node visits, [3] triangle tests
    uint8_t* pixelSet;            // per pixel of the frame: sample set index
    uint32_t* queueS;             // queue re-ordered by the ray sort (sort.cu)
    uint32_t* sortKeys; uint32_t* sortKeysOut;   // sort keys (origin cell Morton code | direction octant)
};

// A render call covers numFaces frames of one size (1 for rtRenderFrame, the 12 stereo cube cameras of a viewpoint for
// yrtxRenderCubeMap) as ONE wavefront: the frames are stacked into a virtual pixel range [0, numFaces * pixelsPerFace), every queue,
// counter and launch spans all of them, so bounce queues are numFaces times longer and launches numFaces times fewer.
#define YRT_MAX_FACES 12
struct FrameCameras { CameraData cam[YRT_MAX_FACES]; };     // by value to the ray-generation kernels only (4.9 KB of kernel parameters)
struct FaceTarget { float4* accum; void* fb; };             // per face: accumulation buffer and packed output

struct FrameConst {               // everything a frame's kernels need, passed by value
    SceneData scene;
    IntegratorData integ;
    const float* sampleTable;     // device
    int width, height;            // raster size of one face
    int numFaces; uint32_t pixelsPerFace;   // active (this server's) pixels per face
    int serverID, serverCount;    // row-band interleave of the reference's network device (api/swapchain.h:57-70)
    float rcpWidth, rcpHeight;
    int debugRenderer;            // 1: renderers/debugrenderer.cpp semantics
    int countStats;
};

struct FilmParams {
    FaceTarget face[YRT_MAX_FACES];   // accum: per pixel (buffer coordinates) sum L.rgb | sum weight; fb: packed output in the framebuffer's format
    int format;                   // 0 RGB_FLOAT32, 1 RGBA8, 2 RGB8
    int fbStrideBytes;
    int accumulate;
    float gamma, rcpGamma; int vignetting;
};

#ifndef YRT_SHADE_MINBLOCKS
#define YRT_SHADE_MINBLOCKS 6
#endif
struct LaunchCfg { int blocks; int threads; cudaStream_t stream; };

// pixels [pixelBegin, pixelBegin+numPixels) of the *active-row* enumeration of the frame
void launch_pixel_sets(const FrameConst& fc, uint8_t* pixelSet, int sets, LaunchCfg lc);
void launch_raygen(const FrameConst& fc, const FrameCameras& cams, const WavefrontBuffers& wb, uint32_t pixelBegin, uint32_t numPixels, LaunchCfg lc);
void launch_trace_closest(const FrameConst& fc, const WavefrontBuffers& wb, int queueSel, LaunchCfg lc);
void launch_shade(const FrameConst& fc, const WavefrontBuffers& wb, int queueSel, uint32_t pixelBegin, int depth, LaunchCfg lc);
void launch_trace_shadow(const FrameConst& fc, const WavefrontBuffers& wb, LaunchCfg lc);
void launch_resolve(const FrameConst& fc, const WavefrontBuffers& wb, int queueSel, LaunchCfg lc);
void launch_film(const FrameConst& fc, const WavefrontBuffers& wb, const FilmParams& fp, uint32_t pixelBegin, uint32_t numPixels, LaunchCfg lc);
// renderers/debugrenderer.cpp:66-148 (maxDepth 1): primary-hit ID image written straight into the framebuffer
void launch_debug(const FrameConst& fc, const FrameCameras& cams, const WavefrontBuffers& wb, const FilmParams& fp, uint32_t numPixels, LaunchCfg lc);
void launch_export_primary(const FrameConst& fc, const WavefrontBuffers& wb, uint32_t pixelBegin, uint32_t numPixels, float* out, LaunchCfg lc);
// ray sort (sort.cu): keys from the extension rays of queue `queueSel`, CUB radix sort of (key, path id) into wb.queueS
void launch_sort_keys(const FrameConst& fc, const WavefrontBuffers& wb, int queueSel, uint32_t n, LaunchCfg lc);
size_t sort_temp_bytes(uint32_t capacity);
void sort_queue(const WavefrontBuffers& wb, int queueSel, uint32_t n, void* temp, size_t tempBytes, cudaStream_t stream);
#define YRT_SORT_KEY_BITS 21
// yrtxTraceRays: rays/hits on the device, 8 floats each
void launch_trace_user(const SceneData& sc, const float* rays, float* hits, size_t n, int closest, int countStats,
                       unsigned long long* stats, uint32_t* workCounter, LaunchCfg lc);

}  // namespace yrt
