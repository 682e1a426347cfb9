// device_impl.hpp — the device object behind the C-ABI (include/yrt_device.h) and the two handle kinds
// that own GPU memory (scene, framebuffer). Internal to libyrt_device_cuda.so.
//
// Mirrors the role (not the code) of SingleRayDevice (reference: devices/device_singleray/api/
// singleray_device.cpp:155-168), BackendSceneFlat::Handle (api/scene_flat.h:58-118) and SwapChain
// (api/swapchain.h:29-125).
#pragma once
#include <cuda_runtime.h>

#include <mutex>
#include <string>
#include <vector>

#include "../../include/yrt_device.h"
#include "device_internal.hpp"
#include "host_objects.hpp"
#include "sampler_host.hpp"

namespace yrt {

#define YRT_CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " #x); } while (0)

template <typename T> struct DevBuf {          // owning device array
    T* p = nullptr; size_t n = 0;
    void alloc(size_t count) {
        if (count <= n && p) return;
        release();
        YRT_CK(cudaMalloc((void**)&p, (count ? count : 1) * sizeof(T))); n = count;
    }
    void upload(const std::vector<T>& v, cudaStream_t s) {
        alloc(v.size());
        if (!v.empty()) YRT_CK(cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s));
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    ~DevBuf() { release(); }
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete; DevBuf& operator=(const DevBuf&) = delete;
};

// One slot of a scene: what BackendSceneFlat::Primitive holds after setPrimitive baked the transform
// (api/scene_flat.h:63-73).
struct ScenePrim {
    std::shared_ptr<ShapeObj> shape;           // transformed copy (or the original for an identity transform)
    std::shared_ptr<LightObj> light;           // transformed copy
    std::shared_ptr<MaterialObj> material;
    int illumMask = -1, shadowMask = -1;
};

struct HdriHost {                              // host-side sampling state of one HDRILight (lights/hdrilight.cpp:39-56)
    std::shared_ptr<ImageObj> image; Dist2D dist; Col L; Aff3 local2world;
};

struct SceneHandle : Handle {
    SceneHandle() : Handle(HK_SCENE) {}
    ~SceneHandle() override;
    std::string accelTy = "default", builderTy = "default", traverserTy = "default";
    std::vector<std::shared_ptr<ScenePrim>> prims;
    bool dirty = true;                         // prims changed since the last commit
    bool committed = false;
    // Incremental commit (the per-cube-face billboard update, renderer.cpp:551-559, only moves vertices): a slot whose new
    // primitive has the signature of the committed one is patched in place; anything else re-flattens the scene.
    struct SlotLayout { const void* material = nullptr; int type = -1, geomID = -1, illumMask = 0, shadowMask = 0; bool hasLight = false, cull = false, allFinite = false;
                        size_t nv = 0, nn = 0, nuv = 0, nt = 0; bool hasTangents = false; uint32_t vtxBase = 0, nrmBase = 0, uvBase = 0, idxBase = 0; };
    std::vector<SlotLayout> layout; std::vector<size_t> patchSlots; bool structureDirty = true; uint32_t numRefs = 0;
    std::vector<float4> hostPositions, hostNormals; std::vector<float2> hostUvs; std::vector<int4> hostIndices;
    // ---- committed (device) state
    SceneData data{};                          // pointers below
    DevBuf<GeomRec> geoms; DevBuf<float4> positions, normals, tangents, motions; DevBuf<float2> uvs; DevBuf<int4> indices;
    DevBuf<MaterialRec> materials; DevBuf<TextureRec> textures; DevBuf<LightRec> lights; DevBuf<uint2> refsBuf;
    void* nodes = nullptr; float4* tris = nullptr; float4* triShade = nullptr; float4* triMotion = nullptr;
    std::vector<std::shared_ptr<ImageObj>> imagesInUse;   // keeps device pixel storage alive
    std::vector<std::shared_ptr<ImageObj>> extraImages;   // images addressable by FrameConst (backplate) appended lazily
    std::vector<TextureRec> hostTextures;
    std::vector<HdriHost> hdri;                // precomputed lights, in precomputedId order
    std::vector<int> geomOfSlot;               // slot -> geomID or -1
    float buildMs = 0.f; uint32_t buildLaunches = 0; uint64_t commitCount = 0, rebuildCount = 0;
    V3 bboxLo = V3(INFINITY), bboxHi = V3(-INFINITY);
    void releaseDevice();
};

struct FrameBufferHandle : Handle {            // SwapChain + AccuBuffer
    FrameBufferHandle() : Handle(HK_FRAMEBUFFER) { constant = true; }
    ~FrameBufferHandle() override;
    int format = 2;                            // 0 RGB_FLOAT32, 1 RGBA8, 2 RGB8
    size_t width = 0, height = 0, depth = 1, cur = 0, strideBytes = 0;
    std::vector<void*> host; std::vector<bool> owned;    // host buffers (cudaHostAlloc'ed when owned)
    void* devPacked = nullptr;                 // stride*height bytes
    float4* accum = nullptr;                   // width*height
    int pendingBuf = -1;                       // host buffer that still needs the D2H copy (yrtxSetReadback(0))
    size_t bytes() const { return strideBytes * height; }
};

struct TableKey {                              // identity of the uploaded sample table
    int spp = -1, sets = -1, maxDepth = -1, filter = -1, iteration = -1, numPre = -1; uint64_t sceneKey = 0;
    bool operator==(const TableKey& o) const {
        return spp == o.spp && sets == o.sets && maxDepth == o.maxDepth && filter == o.filter && iteration == o.iteration &&
               numPre == o.numPre && sceneKey == o.sceneKey;
    }
};

struct WavefrontStorage {
    WavefrontBuffers wb{}; size_t pixelSetCapacity = 0; void* sortTemp = nullptr; size_t sortTempBytes = 0;
    void ensure(uint32_t capacity, uint32_t shadowCapacity, size_t pixels);
    void release();
};

struct FrameTimers {                           // CUDA-event pairs collected during a frame, summed after the final sync
    std::vector<cudaEvent_t> pool; size_t used = 0;
    struct Span { int kind; cudaEvent_t a, b; };
    std::vector<Span> spans;
    cudaEvent_t get();
    void begin(int kind, cudaStream_t s); void end(cudaStream_t s);
    void reset() { used = 0; spans.clear(); }
    void release();
};

}  // namespace yrt

namespace yrt {
// the stereo cube-map strip being assembled on the device (image_codecs.cu; devices/renderer/renderer.cpp:665-725)
double microbench(yrt_device* dev, int kind, size_t bytes);   // microbench.cu
struct CubeStrip { unsigned char* dev = nullptr; size_t w = 0, h = 0; uchar4* wm = nullptr; int wmW = 0, wmH = 0; int facesAdded = 0; };
}
struct yrt_device {
    std::vector<yrt_device*> members;          // non-empty: a group device (cfg gpus=N, group_api.cu); nothing below is used then
    std::mutex mutex;                          // RT_COMMAND_HEADER (api/singleray_device.cpp:97)
    int gpu = 0; int numSMs = 148; cudaStream_t stream = nullptr;
    int serverID = 0, serverCount = 1;         // g_serverID / g_serverCount (api/singleray_device.cpp:109-110)
    uint32_t chunkPaths = 1u << 26;      // paths per wavefront pass: whole faces where memory allows (launch tails dominate small chunks)
    int countStats = 0, verbose = 0, alwaysRebuild = 0, useTimers = 1;
    // node / triangle phase vote (bvh.cuh: TraceTune): triangle phase when tuneTriNum * nT >= tuneTriDen * nN. Measured with the r2 node format: the path
    // tracer's bounce rays do best at 2:1 (C4 +1.5 %, C3 +1.9 %, C2 -1.7 % against 3:1), the raw ray API on the config-5 soup at 3:1 (2:1 costs 5 %)
    int tuneRefillMin = 8, tuneTriNum = 2, tuneTriDen = 1, tuneUserTriNum = 3, tuneSimple = 0;
    int tunePrefetch = 0;                      // cfg prefetch=0|1 (bvh.cuh: TraceTune; measured slower, off)
    size_t l2Bytes = 0;
    int shadeCtas = 6, traceCtas = 8;
    int bvhCollapseDp = 1; float bvhCTri = 0.6f;              // cfg collapse=0|1, ctri=<percent>: SAH-optimal BVH8 collapse and its triangle cost
    int bvhPloc = 1, plocRadius = 8, splitLeaves = 1;           // cfg bvh=0 selects the Karras LBVH hierarchy (A/B), plocr the PLOC search radius
    uint32_t syncMinPaths = 1u << 20;          // cfg syncmin=: per-bounce queue-length read-back only while at least this many paths are alive
    int sortRays = 0; uint32_t sortMin = 1u << 16;   // cfg sort=0|1: re-order bounce queues of at least sortMin rays (sort.cu); measured slower, off
    uint32_t* hostCounters = nullptr;          // pinned: queue lengths read back once per bounce   // cfg refill=,trinum=,triden= (bvh.cuh: TraceTune)
    bool readback = true;                      // copy the frame to the host buffer inside yrtRenderFrame (yrtxSetReadback)
    // Two chunk lanes (cfg lanes=2, off by default): consecutive chunks of a render call run on two streams with their own wavefront
    // state, each kernel launched with half the CTAs, so that a latency-bound shading kernel of one chunk shares the SMs with an
    // issue-bound traversal kernel of the other (ncu r2: k_shade issues 44 % of the time, the traversal kernels 77 %). Measured r2
    // (profiles/README.md): 3-6 % SLOWER on C2 / C3 / C4 for every CTA split tried — co-resident kernels take each other's L1 and
    // warp slots and neither gains issue slots — so one lane stays the default; the frames are bit-identical either way.
    int lanes = 1; cudaStream_t stream1 = nullptr;
    yrt::WavefrontStorage wf, wf1;
    yrt::FrameTimers timers;
    yrt::DevBuf<float> sampleTable; yrt::TableKey tableKey; int tableSpp = 1, tableN1 = 0, tableN2 = 0, tableRec = 0;
    yrt::PixelFilter filters[3]; bool filterReady[3] = {false, false, false};
    yrtx_frame_stats stats{};
    yrt::CubeStrip strip;
    void bind() const;
};
