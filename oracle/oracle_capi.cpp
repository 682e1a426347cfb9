// TEST INFRASTRUCTURE ONLY (oracle). Compiled into oracle/_ref/liboracle_singleray.so together
// with the overlay copy of the reference's device_singleray and the embree2 shim.
//
// Exposes the reference CPU device through the SAME C symbols as the product's C-ABI
// (include/yrt_device.h), so the parity tests drive both implementations with one script.
// Every function below is a one-line forward into the unmodified virtual interface
// embree::Device (reference: devices/device/device.h:126-329) of the object returned by the
// reference's own factory `create` (devices/device_singleray/api/singleray_device.cpp:105-107).
#include <chrono>
#include <thread>
#include <vector>
#include <algorithm>
#include <cstring>
#include <stdexcept>
#include <string>

#include "device/device.h"
#include "api/handle.h"
#include "api/scene.h"
#include "api/swapchain.h"
#include "renderers/integratorrenderer.h"
#include "samplers/sampler.h"
#include "filters/boxfilter.h"
#include "filters/bsplinefilter.h"
#include "embree2/rtcore.h"
#include "embree2/rtcore_ray.h"
#include "../../include/yrt_device.h"

namespace embree { extern "C" Device* create(const char* parms, size_t numThreads, int threadsPriority, const char* rtcore_cfg); }

using embree::Device;

struct yrt_device { Device* d; };

static thread_local std::string g_err;
static double g_lastSeconds = 0, g_lastRays = 0, g_lastHostMs = 0;
static int g_quiet = 1;

extern "C" void yrt_oracle_report_frame(double seconds, double rays) { g_lastSeconds = seconds; g_lastRays = rays; }
extern "C" int yrt_oracle_quiet() { return g_quiet; }
extern "C" void yrt_shim_stats(int enable, uint64_t* nodeVisits, uint64_t* triTests);

#define TRY_H(expr) try { return (yrt_handle)(expr); } catch (const std::exception& e) { g_err = e.what(); return nullptr; }
#define TRY_S(stmt) try { stmt; return YRT_OK; } catch (const std::exception& e) { g_err = e.what(); return YRT_ERROR; }
#define D (dev->d)
typedef Device::RTHandle H;

extern "C" {

yrt_device* yrtCreateDevice(const char* parms, size_t numThreads, int threadsPriority, const char* cfg) {
    try {
        if (cfg && strstr(cfg, "verbose=1")) g_quiet = 0;
        Device* d = embree::create(parms ? parms : "", numThreads, threadsPriority, cfg ? cfg : "");
        if (!d) throw std::runtime_error("device creation failed");
        return new yrt_device{d};
    } catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
void yrtDestroyDevice(yrt_device* dev) { if (dev) { delete dev->d; delete dev; } }
const char* yrtGetLastError(void) { return g_err.c_str(); }

yrt_handle yrtNewCamera(yrt_device* dev, const char* type) { TRY_H(D->rtNewCamera(type)) }
yrt_handle yrtNewData(yrt_device* dev, const char* type, size_t bytes, const void* data) { TRY_H(D->rtNewData(type, bytes, data)) }
yrt_handle yrtNewDataFromFile(yrt_device* dev, const char* type, const char* file, size_t offset, size_t bytes) { TRY_H(D->rtNewDataFromFile(type, file, offset, bytes)) }
yrt_handle yrtNewImage(yrt_device* dev, const char* type, size_t w, size_t h, const void* data, int copy) { TRY_H(D->rtNewImage(type, w, h, data, copy != 0)) }
yrt_handle yrtNewImageFromFile(yrt_device* dev, const char* file) { TRY_H(D->rtNewImageFromFile(file)) }
yrt_handle yrtNewTexture(yrt_device* dev, const char* type) { TRY_H(D->rtNewTexture(type)) }
yrt_handle yrtNewMaterial(yrt_device* dev, const char* type) { TRY_H(D->rtNewMaterial(type)) }
yrt_handle yrtNewShape(yrt_device* dev, const char* type) { TRY_H(D->rtNewShape(type)) }
yrt_handle yrtNewLight(yrt_device* dev, const char* type) { TRY_H(D->rtNewLight(type)) }
yrt_handle yrtNewShapePrimitive(yrt_device* dev, yrt_handle shape, yrt_handle material, const float* xfm, int faceCamera) {
    TRY_H(D->rtNewShapePrimitive((Device::RTShape)shape, (Device::RTMaterial)material, xfm, faceCamera != 0)) }
yrt_handle yrtNewLightPrimitive(yrt_device* dev, yrt_handle light, yrt_handle material, const float* xfm) {
    TRY_H(D->rtNewLightPrimitive((Device::RTLight)light, (Device::RTMaterial)material, xfm)) }
yrt_handle yrtTransformPrimitive(yrt_device* dev, yrt_handle prim, const float* xfm) { TRY_H(D->rtTransformPrimitive((Device::RTPrimitive)prim, xfm)) }
yrt_handle yrtNewScene(yrt_device* dev, const char* type) { TRY_H(D->rtNewScene(type)) }
yrt_status yrtSetPrimitive(yrt_device* dev, yrt_handle scene, size_t slot, yrt_handle prim) { TRY_S(D->rtSetPrimitive((Device::RTScene)scene, slot, (Device::RTPrimitive)prim)) }
yrt_status yrtUpdatePrimitive(yrt_device* dev, yrt_handle scene, size_t slot, yrt_handle prim, const float p[3], const float u[3]) {
    TRY_S(D->rtUpdatePrimitive((Device::RTScene)scene, slot, (Device::RTPrimitive)prim, embree::Vector3f(p[0], p[1], p[2]), embree::Vector3f(u[0], u[1], u[2]))) }
yrt_handle yrtNewToneMapper(yrt_device* dev, const char* type) { TRY_H(D->rtNewToneMapper(type)) }
yrt_handle yrtNewRenderer(yrt_device* dev, const char* type) { TRY_H(D->rtNewRenderer(type)) }
yrt_handle yrtNewFrameBuffer(yrt_device* dev, const char* type, size_t w, size_t h, size_t buffers, void** ptrs) { TRY_H(D->rtNewFrameBuffer(type, w, h, buffers, ptrs)) }
void* yrtMapFrameBuffer(yrt_device* dev, yrt_handle fb, int bufID) { TRY_H(D->rtMapFrameBuffer((Device::RTFrameBuffer)fb, bufID)) }
yrt_status yrtUnmapFrameBuffer(yrt_device* dev, yrt_handle fb, int bufID) { TRY_S(D->rtUnmapFrameBuffer((Device::RTFrameBuffer)fb, bufID)) }
yrt_status yrtSwapBuffers(yrt_device* dev, yrt_handle fb) { TRY_S(D->rtSwapBuffers((Device::RTFrameBuffer)fb)) }
yrt_status yrtIncRef(yrt_device* dev, yrt_handle h) { TRY_S(D->rtIncRef((H)h)) }
yrt_status yrtDecRef(yrt_device* dev, yrt_handle h) { TRY_S(D->rtDecRef((H)h)) }

yrt_status yrtSetBool1(yrt_device* dev, yrt_handle h, const char* p, int x) { TRY_S(D->rtSetBool1((H)h, p, x != 0)) }
yrt_status yrtSetBool2(yrt_device* dev, yrt_handle h, const char* p, int x, int y) { TRY_S(D->rtSetBool2((H)h, p, x != 0, y != 0)) }
yrt_status yrtSetBool3(yrt_device* dev, yrt_handle h, const char* p, int x, int y, int z) { TRY_S(D->rtSetBool3((H)h, p, x != 0, y != 0, z != 0)) }
yrt_status yrtSetBool4(yrt_device* dev, yrt_handle h, const char* p, int x, int y, int z, int w) { TRY_S(D->rtSetBool4((H)h, p, x != 0, y != 0, z != 0, w != 0)) }
yrt_status yrtSetInt1(yrt_device* dev, yrt_handle h, const char* p, int x) { TRY_S(D->rtSetInt1((H)h, p, x)) }
yrt_status yrtSetInt2(yrt_device* dev, yrt_handle h, const char* p, int x, int y) { TRY_S(D->rtSetInt2((H)h, p, x, y)) }
yrt_status yrtSetInt3(yrt_device* dev, yrt_handle h, const char* p, int x, int y, int z) { TRY_S(D->rtSetInt3((H)h, p, x, y, z)) }
yrt_status yrtSetInt4(yrt_device* dev, yrt_handle h, const char* p, int x, int y, int z, int w) { TRY_S(D->rtSetInt4((H)h, p, x, y, z, w)) }
yrt_status yrtSetPointer(yrt_device* dev, yrt_handle h, const char* p, void* ptr) { TRY_S(D->rtSetPointer((H)h, p, ptr)) }
yrt_status yrtSetFloat1(yrt_device* dev, yrt_handle h, const char* p, float x) { TRY_S(D->rtSetFloat1((H)h, p, x)) }
yrt_status yrtGetFloat1(yrt_device* dev, yrt_handle h, const char* p, float* x) { TRY_S(D->rtGetFloat1((H)h, p, *x)) }
yrt_status yrtSetFloat2(yrt_device* dev, yrt_handle h, const char* p, float x, float y) { TRY_S(D->rtSetFloat2((H)h, p, x, y)) }
yrt_status yrtSetFloat3(yrt_device* dev, yrt_handle h, const char* p, float x, float y, float z) { TRY_S(D->rtSetFloat3((H)h, p, x, y, z)) }
yrt_status yrtGetFloat3(yrt_device* dev, yrt_handle h, const char* p, float* x, float* y, float* z) { TRY_S(D->rtGetFloat3((H)h, p, *x, *y, *z)) }
yrt_status yrtSetFloat4(yrt_device* dev, yrt_handle h, const char* p, float x, float y, float z, float w) { TRY_S(D->rtSetFloat4((H)h, p, x, y, z, w)) }
yrt_status yrtSetArray(yrt_device* dev, yrt_handle h, const char* p, const char* type, yrt_handle data, size_t size, size_t stride, size_t ofs) {
    TRY_S(D->rtSetArray((H)h, p, type, (Device::RTData)data, size, stride, ofs)) }
yrt_status yrtSetString(yrt_device* dev, yrt_handle h, const char* p, const char* s) { TRY_S(D->rtSetString((H)h, p, s)) }
yrt_status yrtGetString(yrt_device* dev, yrt_handle h, const char* p, char* buf, size_t n) {
    try { std::string s; D->rtGetString((H)h, p, s); if (n) { strncpy(buf, s.c_str(), n - 1); buf[n - 1] = 0; } return YRT_OK; }
    catch (const std::exception& e) { g_err = e.what(); return YRT_ERROR; } }
yrt_status yrtSetImage(yrt_device* dev, yrt_handle h, const char* p, yrt_handle img) { TRY_S(D->rtSetImage((H)h, p, (Device::RTImage)img)) }
yrt_status yrtSetTexture(yrt_device* dev, yrt_handle h, const char* p, yrt_handle tex) { TRY_S(D->rtSetTexture((H)h, p, (Device::RTTexture)tex)) }
yrt_status yrtSetTransform(yrt_device* dev, yrt_handle h, const char* p, const float* x) { TRY_S(D->rtSetTransform((H)h, p, x)) }
yrt_status yrtGetTransform(yrt_device* dev, yrt_handle h, const char* p, float* out) {
    try { embree::AffineSpace3f a; D->rtGetTransform((H)h, p, &a);
          const embree::Vector3f c[4] = {a.l.vx, a.l.vy, a.l.vz, a.p};
          for (int i = 0; i < 4; i++) { out[3 * i] = c[i].x; out[3 * i + 1] = c[i].y; out[3 * i + 2] = c[i].z; }
          return YRT_OK; }
    catch (const std::exception& e) { g_err = e.what(); return YRT_ERROR; } }
yrt_status yrtClear(yrt_device* dev, yrt_handle h) { TRY_S(D->rtClear((H)h)) }
yrt_status yrtCommit(yrt_device* dev, yrt_handle h) { TRY_S(D->rtCommit((H)h)) }

yrt_status yrtRenderFrame(yrt_device* dev, yrt_handle renderer, yrt_handle camera, yrt_handle scene, yrt_handle tonemapper, yrt_handle fb, int accumulate) {
    try {
        auto t0 = std::chrono::steady_clock::now();
        D->rtRenderFrame((Device::RTRenderer)renderer, (Device::RTCamera)camera, (Device::RTScene)scene, (Device::RTToneMapper)tonemapper, (Device::RTFrameBuffer)fb, accumulate);
        g_lastHostMs = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        return YRT_OK;
    } catch (const std::exception& e) { g_err = e.what(); return YRT_ERROR; }
}
int yrtPick(yrt_device* dev, yrt_handle camera, float x, float y, yrt_handle scene, float* px, float* py, float* pz) {
    try { return D->rtPick((Device::RTCamera)camera, x, y, (Device::RTScene)scene, *px, *py, *pz) ? 1 : 0; }
    catch (const std::exception& e) { g_err = e.what(); return -1; } }

// ---- extensions ---------------------------------------------------------------------
yrt_status yrtxGetFrameStats(yrt_device*, yrtx_frame_stats* out) {
    memset(out, 0, sizeof(*out));
    out->render_ms = g_lastSeconds * 1e3;       // the reference's own timer span (integratorrenderer.cpp:99-110,122)
    out->host_ms = g_lastHostMs;
    out->rays_closest = (uint64_t)g_lastRays;   // the reference counts both kinds in one counter
    yrt_shim_stats(1, &out->node_visits, &out->tri_tests);
    return YRT_OK;
}

yrt_status yrtxTraceRays(yrt_device*, yrt_handle scene, size_t n, const float* rays, void* hits, int closest, int onDevice, float* ms) {
    try {
        if (onDevice) throw std::runtime_error("oracle has no device memory");
        auto* sh = dynamic_cast<embree::InstanceHandle<embree::BackendScene>*>((embree::_RTHandle*)scene);
        if (!sh || !sh->getInstance()) throw std::runtime_error("invalid scene handle");
        RTCScene rs = sh->getInstance()->scene;
        auto t0 = std::chrono::steady_clock::now();
        // large batches (the at-size parity tests: millions of primary rays per cube face) are split over the host threads; the shim's
        // rtcIntersect / rtcOccluded are re-entrant (the reference calls them from all its render threads)
        const size_t nThreads = n >= 65536 ? std::max<size_t>(1, std::thread::hardware_concurrency()) : 1;
        auto work = [&](size_t begin, size_t end) {
        for (size_t i = begin; i < end; i++) {
            RTCRay r; memset(&r, 0, sizeof(r));
            const float* q = rays + 8 * i;
            r.org[0] = q[0]; r.org[1] = q[1]; r.org[2] = q[2]; r.tnear = q[3];
            r.dir[0] = q[4]; r.dir[1] = q[5]; r.dir[2] = q[6]; r.tfar = q[7];
            r.mask = 0xffffffffu; r.geomID = r.primID = r.instID = RTC_INVALID_GEOMETRY_ID;
            float* hf = (float*)hits + 8 * i; int32_t* hi = (int32_t*)hf;
            if (closest) {
                rtcIntersect(rs, r);
                hf[0] = r.tfar; hf[1] = r.u; hf[2] = r.v; hi[3] = (int32_t)r.geomID; hi[4] = (int32_t)r.primID;
                hf[5] = r.Ng[0]; hf[6] = r.Ng[1]; hf[7] = r.Ng[2];
            } else {
                rtcOccluded(rs, r);
                hi[3] = r.geomID == RTC_INVALID_GEOMETRY_ID ? -1 : 0;
            }
        }
        };
        if (nThreads == 1) work(0, n);
        else {
            std::vector<std::thread> th;
            const size_t per = (n + nThreads - 1) / nThreads;
            for (size_t k = 0; k < nThreads; k++) { const size_t b = std::min(n, k * per), e = std::min(n, b + per); if (b < e) th.emplace_back(work, b, e); }
            for (auto& t : th) t.join();
        }
        if (ms) *ms = (float)std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        return YRT_OK;
    } catch (const std::exception& e) { g_err = e.what(); return YRT_ERROR; }
}

// The reference's own SamplerFactory, driven the way RenderJob drives it (integratorrenderer.cpp:86-88 +
// PathTraceIntegrator::requestSamples pathtraceintegrator.cpp:35-47, no precomputed lights).
yrt_status yrtxHostSampleTable(const char* filter, int spp, int sets, int maxDepth, int iteration, int* outSpp, int* n1, int* n2, float* table) {
    try {
        const std::string f(filter ? filter : "bspline");
        embree::Ref<embree::Filter> flt;
        if (f == "box") flt = new embree::BoxFilter; else if (f == "bspline") flt = new embree::BSplineFilter;
        else if (f != "none") throw std::runtime_error("unknown filter type: " + f);
        embree::Ref<embree::SamplerFactory> sf = new embree::SamplerFactory((unsigned)spp, (unsigned)sets);
        sf->request2D(); sf->request2D(maxDepth); sf->request1D(maxDepth);
        sf->init(iteration, flt);
        const int a = sf->numSamples1D, b = sf->numSamples2D, rec = 5 + a + 2 * b;
        if (outSpp) *outSpp = sf->samplesPerPixel; if (n1) *n1 = a; if (n2) *n2 = b;
        if (table)
            for (int set = 0; set < sf->sampleSets; set++)
                for (int s = 0; s < sf->samplesPerPixel; s++) {
                    const embree::PrecomputedSample& p = sf->samples[set][s];
                    float* o = table + ((size_t)set * sf->samplesPerPixel + s) * rec;
                    o[0] = p.pixel.x; o[1] = p.pixel.y; o[2] = p.time; o[3] = p.lens.x; o[4] = p.lens.y;
                    for (int d = 0; d < a; d++) o[5 + d] = p.samples1D[d];
                    for (int d = 0; d < b; d++) { o[5 + a + 2 * d] = p.samples2D[d].x; o[5 + a + 2 * d + 1] = p.samples2D[d].y; }
                }
        return YRT_OK;
    } catch (const std::exception& e) { g_err = e.what(); return YRT_ERROR; }
}

}  // extern "C"
