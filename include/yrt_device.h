/* yrt_device.h — C-ABI of device_cuda, the B200-native render device.
 *
 * This is the drop-in boundary: one `extern "C"` entry point per virtual of the reference's
 * `embree::Device` plugin interface (reference: devices/device/device.h:126-329), plus the
 * plugin factory (reference: devices/device/device.cpp:24-35, exported as `create` by every
 * back end, e.g. devices/device_singleray/api/singleray_device.cpp:105-107).
 * Plain pointers and sizes only; no C++ or torch types cross this line.  The C++ adapter
 * `CudaDevice : embree::Device` that forwards to these functions (what a maintainer of the
 * reference adds, ~150 lines) is in yulio_raytracer_b200/adapter/ and INTEGRATION.md.
 *
 * Conventions (mirroring the reference's semantics):
 *  - Every object is an opaque, reference-counted handle created with count 1
 *    (reference: devices/device_singleray/api/handle.h:29-31). yrtDecRef destroys at 0.
 *  - yrtSet* buffer a parameter inside the handle; yrtCommit (re)creates the object from the
 *    buffered parameters only if something changed (api/handle.h:99-103,129-133).
 *    Unknown property names are ignored; wrong-typed values fall back to the default.
 *  - Errors: the reference throws std::runtime_error across the boundary
 *    (api/singleray_device.cpp:190..435). Here: handle-returning calls return NULL and
 *    void calls return a non-zero yrt_status; yrtGetLastError gives the message (per calling
 *    thread). The adapter turns these back into std::runtime_error.
 *  - Setters with handle == NULL silently succeed; property == NULL is an error
 *    (api/singleray_device.cpp:474-479).
 *  - All calls on one device are serialised by a per-device mutex
 *    (api/singleray_device.cpp:97); yrtRenderFrame is synchronous.
 *  - There is NO CPU fallback: yrtCreateDevice fails if no sm_100-class CUDA device is usable.
 */
#ifndef YRT_DEVICE_H
#define YRT_DEVICE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define YRT_API __attribute__((visibility("default")))
#else
#define YRT_API
#endif

typedef struct yrt_device yrt_device; /* device_cuda instance (one CUDA context / GPU set) */
typedef void* yrt_handle;             /* == embree::Device::RTHandle and all derived handle types */

typedef int yrt_status;               /* 0 = ok */
#define YRT_OK 0
#define YRT_ERROR 1                   /* std::runtime_error in the reference */

/* ---- device life cycle ------------------------------------------------------------ */
/* reference: Device::rtCreateDevice -> create(parms, numThreads, threadsPriority, rtcore_cfg)
 * (devices/device/device.cpp:24-48). `cfg` is the free-form "k=v,k=v" string the front end
 * passes as -rtcore (devices/renderer/renderer.cpp:922-937); keys understood:
 *   gpu=I (CUDA ordinal, default 0), chunk=P (paths per wavefront chunk, default 2^26, clamped to 40 % of the free device memory), stats=1 (count node
 *   visits / triangle tests), timers=0|1 (per-stage CUDA events, default 1), rebuild=1 (rebuild the BVH on every
 *   scene commit like the reference does), serverID=I,serverCount=N (row-band interleave, see yrtSetInt1(NULL,..)),
 *   sort=0|1 (re-order bounce queues by origin cell + direction octant; default 0: measured slower, profiles/README.md),
 *   verbose=0|1.   numThreads / threadsPriority are accepted and ignored (no CPU workers). */
YRT_API yrt_device* yrtCreateDevice(const char* parms, size_t numThreads, int threadsPriority, const char* cfg);
YRT_API void        yrtDestroyDevice(yrt_device* dev);                 /* reference: virtual ~Device, device.h:54 */
YRT_API const char* yrtGetLastError(void);                             /* message of the calling thread's last failed call */

/* ---- creation of objects (device.h:126-223) ----------------------------------------- */
YRT_API yrt_handle yrtNewCamera(yrt_device*, const char* type);                                   /* device.h:126 */
/* type "immutable": the bytes are copied. type "immutable_managed": the library takes ownership of `data`, which must be a block
 * returned by malloc(); it is released with free(). (The reference's callers allocate managed blocks with embree::alignedMalloc, an
 * interior pointer: the C++ plugin adapter/device_cuda_plugin.cpp copies those and releases them with the alignedFree convention.) */
YRT_API yrt_handle yrtNewData(yrt_device*, const char* type, size_t bytes, const void* data);     /* device.h:134 */
YRT_API yrt_handle yrtNewDataFromFile(yrt_device*, const char* type, const char* file, size_t offset, size_t bytes); /* device.h:144 */
YRT_API yrt_handle yrtNewImage(yrt_device*, const char* type, size_t width, size_t height, const void* data, int copy); /* device.h:151 */
YRT_API yrt_handle yrtNewImageFromFile(yrt_device*, const char* file);                            /* device.h:157 */
YRT_API yrt_handle yrtNewTexture(yrt_device*, const char* type);                                  /* device.h:161 */
YRT_API yrt_handle yrtNewMaterial(yrt_device*, const char* type);                                 /* device.h:167 */
YRT_API yrt_handle yrtNewShape(yrt_device*, const char* type);                                    /* device.h:172 */
YRT_API yrt_handle yrtNewLight(yrt_device*, const char* type);                                    /* device.h:178 */
/* transform: 12 floats, column-major vx,vy,vz,p (common/math/affinespace.h:166-182) or NULL = identity */
YRT_API yrt_handle yrtNewShapePrimitive(yrt_device*, yrt_handle shape, yrt_handle material, const float* transform, int faceCamera); /* device.h:185 */
YRT_API yrt_handle yrtNewLightPrimitive(yrt_device*, yrt_handle light, yrt_handle material, const float* transform);               /* device.h:191 */
YRT_API yrt_handle yrtTransformPrimitive(yrt_device*, yrt_handle prim, const float* transform);   /* device.h:197 */
YRT_API yrt_handle yrtNewScene(yrt_device*, const char* type);                                    /* device.h:200 */
YRT_API yrt_status yrtSetPrimitive(yrt_device*, yrt_handle scene, size_t slot, yrt_handle prim);  /* device.h:204 */
YRT_API yrt_status yrtUpdatePrimitive(yrt_device*, yrt_handle scene, size_t slot, yrt_handle prim,
                                      const float camPos[3], const float camUp[3]);               /* device.h:207 */
YRT_API yrt_handle yrtNewToneMapper(yrt_device*, const char* type);                               /* device.h:210 */
YRT_API yrt_handle yrtNewRenderer(yrt_device*, const char* type);                                 /* device.h:214 */
YRT_API yrt_handle yrtNewFrameBuffer(yrt_device*, const char* type, size_t width, size_t height, size_t buffers, void** ptrs); /* device.h:223 */
YRT_API void*      yrtMapFrameBuffer(yrt_device*, yrt_handle frameBuffer, int bufID);             /* device.h:227 */
YRT_API yrt_status yrtUnmapFrameBuffer(yrt_device*, yrt_handle frameBuffer, int bufID);           /* device.h:231 */
YRT_API yrt_status yrtSwapBuffers(yrt_device*, yrt_handle frameBuffer);                           /* device.h:234 */
YRT_API yrt_status yrtIncRef(yrt_device*, yrt_handle handle);                                     /* device.h:237 */
YRT_API yrt_status yrtDecRef(yrt_device*, yrt_handle handle);                                     /* device.h:241 */

/* ---- setting of parameters (device.h:248-312) --------------------------------------- */
YRT_API yrt_status yrtSetBool1(yrt_device*, yrt_handle, const char* property, int x);
YRT_API yrt_status yrtSetBool2(yrt_device*, yrt_handle, const char* property, int x, int y);
YRT_API yrt_status yrtSetBool3(yrt_device*, yrt_handle, const char* property, int x, int y, int z);
YRT_API yrt_status yrtSetBool4(yrt_device*, yrt_handle, const char* property, int x, int y, int z, int w);
YRT_API yrt_status yrtSetInt1(yrt_device*, yrt_handle, const char* property, int x);
YRT_API yrt_status yrtSetInt2(yrt_device*, yrt_handle, const char* property, int x, int y);
YRT_API yrt_status yrtSetInt3(yrt_device*, yrt_handle, const char* property, int x, int y, int z);
YRT_API yrt_status yrtSetInt4(yrt_device*, yrt_handle, const char* property, int x, int y, int z, int w);
YRT_API yrt_status yrtSetPointer(yrt_device*, yrt_handle, const char* property, void* p);
YRT_API yrt_status yrtSetFloat1(yrt_device*, yrt_handle, const char* property, float x);
YRT_API yrt_status yrtGetFloat1(yrt_device*, yrt_handle, const char* property, float* x);         /* device.h:276 */
YRT_API yrt_status yrtSetFloat2(yrt_device*, yrt_handle, const char* property, float x, float y);
YRT_API yrt_status yrtSetFloat3(yrt_device*, yrt_handle, const char* property, float x, float y, float z);
YRT_API yrt_status yrtGetFloat3(yrt_device*, yrt_handle, const char* property, float* x, float* y, float* z); /* device.h:283 */
YRT_API yrt_status yrtSetFloat4(yrt_device*, yrt_handle, const char* property, float x, float y, float z, float w);
/* type in {"bool1".."bool4","int1".."int4","float1".."float4"}; stride == (size_t)-1 means packed (device.h:289) */
YRT_API yrt_status yrtSetArray(yrt_device*, yrt_handle, const char* property, const char* type, yrt_handle data,
                               size_t size, size_t stride, size_t ofs);
YRT_API yrt_status yrtSetString(yrt_device*, yrt_handle, const char* property, const char* str);
/* writes at most bufBytes-1 characters + NUL; returns YRT_OK (device.h:293) */
YRT_API yrt_status yrtGetString(yrt_device*, yrt_handle, const char* property, char* buf, size_t bufBytes);
YRT_API yrt_status yrtSetImage(yrt_device*, yrt_handle, const char* property, yrt_handle image);
YRT_API yrt_status yrtSetTexture(yrt_device*, yrt_handle, const char* property, yrt_handle texture);
YRT_API yrt_status yrtSetTransform(yrt_device*, yrt_handle, const char* property, const float* transform12);
YRT_API yrt_status yrtGetTransform(yrt_device*, yrt_handle, const char* property, float* transform12); /* device.h:303 */
YRT_API yrt_status yrtClear(yrt_device*, yrt_handle);                                             /* device.h:308 */
YRT_API yrt_status yrtCommit(yrt_device*, yrt_handle);                                            /* device.h:312 */

/* ---- render calls (device.h:322-329) ------------------------------------------------ */
YRT_API yrt_status yrtRenderFrame(yrt_device*, yrt_handle renderer, yrt_handle camera, yrt_handle scene,
                                  yrt_handle tonemapper, yrt_handle frameBuffer, int accumulate); /* device.h:322 */
/* returns 1 if a point was picked, 0 if not, -1 on error (device.h:329) */
YRT_API int        yrtPick(yrt_device*, yrt_handle camera, float x, float y, yrt_handle scene, float* px, float* py, float* pz);

/* =====================================================================================
 * Extensions (prefix yrtx): measurement and parity hooks. They have no counterpart in
 * devices/device/device.h; the reference reports the same quantities on stdout
 * ("render  F fps, T ms, R mrps", devices/device_singleray/renderers/integratorrenderer.cpp:96-111).
 * ===================================================================================== */
typedef struct yrtx_frame_stats {
    double   render_ms;        /* CUDA-event time of the wavefront loop of the last yrtRenderFrame */
    double   build_ms;         /* CUDA-event time of the last BVH build (scene commit that had to rebuild) */
    double   host_ms;          /* wall-clock of the whole last yrtRenderFrame call (incl. the D2H copy of the frame) */
    uint64_t rays_closest;     /* rtcIntersect-equivalent rays (pathtraceintegrator.cpp:74) */
    uint64_t rays_shadow;      /* rtcOccluded-equivalent rays (pathtraceintegrator.cpp:161) */
    uint64_t kernel_launches;  /* device_cuda kernels launched by the last yrtRenderFrame */
    double   trace_ms;         /* CUDA-event time spent in the two traversal kernels */
    uint64_t node_visits;      /* only when cfg has stats=1: BVH8 nodes fetched */
    uint64_t tri_tests;        /* only when cfg has stats=1: triangles tested */
    uint64_t num_triangles;    /* triangles in the committed scene */
    uint64_t num_nodes;        /* BVH8 nodes in the committed scene */
    uint32_t num_gpus;
    uint32_t reserved;
    double   closest_ms;       /* CUDA-event time of the closest-hit traversal launches */
    double   shadow_ms;        /* CUDA-event time of the any-hit traversal launches */
    double   shade_ms;         /* CUDA-event time of the shading + resolve launches */
    double   raygen_film_ms;   /* CUDA-event time of ray generation + film launches */
    uint64_t closest_launches; /* number of closest-hit traversal launches */
    uint64_t shadow_launches;  /* number of any-hit traversal launches */
    uint64_t h2d_bytes;        /* host->device bytes moved by the last yrtRenderFrame (sample table, constants) */
    uint64_t d2h_bytes;        /* device->host bytes moved by the last yrtRenderFrame (the frame) */
    double   sort_ms;          /* CUDA-event time of the ray-sort launches (key generation + radix sort) */
    uint64_t bvh_builds;       /* BVH builds since the scene handle was created (F8: commits that did not change geometry reuse it) */
    double   resolve_ms;       /* CUDA-event time of the resolve + counter-reset launches (shade_ms is the shading kernel alone) */
    uint64_t node_visits_shadow; /* stats=1: the any-hit kernel's share of node_visits */
    uint64_t tri_tests_shadow;   /* stats=1: the any-hit kernel's share of tri_tests */
    uint64_t path_vertices;    /* queue entries processed by the shading kernel (hits + misses) */
    uint64_t shade_launches;   /* number of shading-kernel launches */
    double   miss_ms;          /* CUDA-event time of the miss / environment kernel launches */
    uint64_t errors;           /* device error flags raised during the call (bit 0: traversal stack overflow); non-zero fails the call */
} yrtx_frame_stats;
YRT_API yrt_status yrtxGetFrameStats(yrt_device*, yrtx_frame_stats* out);

/* One ray = 8 floats {org.xyz, tnear, dir.xyz, tfar} (the RTRay layout of device.h:106-112).
 * One hit = 8 x 32 bit {t, u, v, geomID(int), primID(int), Ng.x, Ng.y, Ng.z}; geomID = -1 on miss.
 * rays/hits are HOST pointers unless `onDevice` is non-zero. Returns the CUDA-event time of the
 * traversal kernel in *ms (may be NULL). closest != 0: rtcIntersect semantics, else rtcOccluded
 * (hit[3] = 0 if occluded else -1, other fields untouched). */
YRT_API yrt_status yrtxTraceRays(yrt_device*, yrt_handle scene, size_t n, const float* rays, void* hits,
                                 int closest, int onDevice, float* ms);
/* Primary rays of one frame exactly as yrtRenderFrame would generate them: for pixel (x,y),
 * sample s: index ((y*width + x)*spp + s), 8 floats as above; also the sample-set index chosen
 * per pixel in sets[y*width+x] (may be NULL). Host pointers. */
YRT_API yrt_status yrtxPrimaryRays(yrt_device*, yrt_handle renderer, yrt_handle camera, yrt_handle frameBuffer,
                                   float* rays, int* sets);
/* Precomputed sample table of a renderer for iteration `iteration` (A2 in SURVEY §8a): returns the
 * number of floats per (set, sample) record in *recordFloats and fills table[sets*spp*recordFloats]
 * (may be NULL to query sizes): {pixel.x, pixel.y, time, lens.x, lens.y, 1D[n1], 2D[2*n2]}. */
YRT_API yrt_status yrtxSampleTable(yrt_device*, yrt_handle renderer, yrt_handle scene, int iteration,
                                   int* sets, int* spp, int* n1, int* n2, float* table);
/* Same table without a device or handles (host-only code; usable on a box without a GPU): filter in {"bspline","box","none"}.
 * With table == NULL only the sizes are returned. Record layout as yrtxSampleTable. */
YRT_API yrt_status yrtxHostSampleTable(const char* filter, int spp, int sets, int maxDepth, int iteration,
                                       int* outSpp, int* n1, int* n2, float* table);
/* Device-resident copy of the current buffer of a framebuffer (what yrtRenderFrame wrote before the D2H copy):
 * row stride as in the reference (api/framebuffer.h:106,146,195). Used by bench.py to gather row bands
 * between ranks over NCCL without a host round trip. */
YRT_API yrt_status yrtxFrameBufferDevice(yrt_device*, yrt_handle frameBuffer, void** devPtr, size_t* bytes, size_t* strideBytes);
/* readbackEachFrame == 0: yrtRenderFrame leaves the frame on the GPU and yrtMapFrameBuffer copies it on demand
 * (default 1: the frame is in the host buffer when yrtRenderFrame returns, as in the reference). */
YRT_API yrt_status yrtxSetReadback(yrt_device*, int readbackEachFrame);
/* The cube-face loop of the reference's front end (devices/renderer/renderer.cpp:543-632) as ONE call: the numFaces (1..12) cameras of a
 * viewpoint — same renderer, scene, tone mapper, and one frame buffer per face, all of one size and format — are rendered as one wavefront
 * (12 times longer ray queues, 12 times fewer kernel launches than 12 yrtRenderFrame calls). Each frame is what yrtRenderFrame would have
 * produced for that camera (same sample tables, same per-tile sample sets: bit-identical), so the loop
 *     for (i = 0; i < 12; i++) { rtUpdatePrimitive x prims; rtCommit(scene); rtRenderFrame(.., cameras[i], .., frameBuffer, 0); ... }
 * becomes  rtUpdatePrimitive x prims; rtCommit(scene); yrtxRenderCubeMap(.., cameras, 12, .., frameBuffers, 0)  — the 12 cameras of a
 * viewpoint share their "origin" (ColladaLoader.cpp:470-505), so the camera-aligned primitives turn the same way for every face. */
YRT_API yrt_status yrtxRenderCubeMap(yrt_device*, yrt_handle renderer, const yrt_handle* cameras, size_t numFaces, yrt_handle scene,
                                     yrt_handle tonemapper, const yrt_handle* frameBuffers, int accumulate);

/* Changes one of the integer cfg keys of yrtCreateDevice on a live device (measurement / A-B only): "lanes" (1 | 2 concurrent chunk streams),
 * "tracectas", "shadectas" (persistent CTAs per SM), "syncmin", "timers", "verbose", "rebuild" (every scene commit re-flattens and rebuilds). Unknown keys are an error. */
YRT_API yrt_status yrtxSetOption(yrt_device*, const char* key, long value);
/* The roofline denominators SURVEY §8d asks to be measured on the box itself (csrc/microbench.cu; not on the render path):
 * kind 0: FP32 FMA issue, *result in TFLOP/s; kind 1: read bandwidth in GB/s of a `bytes` working set re-read with L1 bypassed (L2 read
 * bandwidth below the L2 size, HBM read bandwidth far above it); kind 2: warp-instruction issue rate in 1e9 warp instructions / s. */
YRT_API yrt_status yrtxMicrobench(yrt_device*, int kind, size_t bytes, double* result);

/* ---- the image readers behind yrtNewImageFromFile (common/image/image.cpp:26-58), exposed for tests -----
 * format: 0 RGB8, 1 RGBA8, 2 RGB_FLOAT32, 3 RGBA_FLOAT32; pixels may be NULL (query the size first). File images are RGBA8 with row 0
 * = bottom scanline of the picture, as the reference's JPEG / FreeImage loaders store them. yrtxDecodePNGFile needs no device. */
YRT_API yrt_status yrtxReadImage(yrt_device*, yrt_handle image, int* width, int* height, int* format, void* pixels);
YRT_API yrt_status yrtxDecodePNGFile(const char* file, int flipVertical, int flipHorizontal, int* width, int* height, void* rgba);

/* ---- the stereo cube-map strip, assembled and JPEG-encoded on the device ------------------------------
 * What the reference's front end does on the host after every rtRenderFrame of a viewpoint (devices/renderer/renderer.cpp:620-725):
 * map the frame, blend the 100x100 watermark into the centre of faces 0-3 (:637-654), copy the 12 frames into one 12W x H image in
 * the order Left Right Up Down Back Front, cameras 6-11 first (:665-711), store it as JPEG at jpegQuality (:713-718, common/image/
 * jpeg.cpp:207-250). Here the frames stay on the GPU: yrtxStripAddFace copies the RGB8 frame just rendered into its segment, and
 * yrtxStripEncodeJPEG encodes with nvJPEG. cubeFaceIndex = the camera's "cubeFaceIndex" (0..11). */
YRT_API yrt_status yrtxStripBegin(yrt_device*, size_t faceWidth, size_t faceHeight);
/* PNG file, loaded as the reference loads its resource (flipped vertically and horizontally, renderer.cpp:84); NULL or "" removes it */
YRT_API yrt_status yrtxStripSetWatermark(yrt_device*, const char* pngFile);
YRT_API yrt_status yrtxStripAddFace(yrt_device*, yrt_handle frameBuffer, int cubeFaceIndex, int watermark);
/* copies the 12W x H x 3 strip to the host (tests, uncompressed output) */
YRT_API yrt_status yrtxStripRead(yrt_device*, void* rgb);
/* cubeFaceIndex < 0: the whole strip; 0..11: that face's segment (the reference's per-face debug image, :656-660) */
YRT_API yrt_status yrtxStripEncodeJPEG(yrt_device*, int cubeFaceIndex, int quality, const char* file);

#ifdef __cplusplus
}
#endif
#endif /* YRT_DEVICE_H */
