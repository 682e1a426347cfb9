// yulio_rt.cpp — Linux re-host of the Yulio front end above the device boundary:
// the asynchronous entry points StartRT / WaitRT / StopRT / GetLastErrorRT / GetCurrentStatusRT and the per-viewpoint
// stereo-cube-map loop, written against the reference's embree::Device interface only.
//
// What is REUSED from the reference, compiled from its own sources (frontend/build_frontend.py): the scene loaders
// (devices/device/loaders/ColladaLoader.cpp, obj_loader.cpp, xml_loader.cpp -> rtLoadScene) and the vendored, Yulio-patched
// Assimp 3.2. What is NEW here (the reference's devices/renderer/renderer.cpp is Win32-bound: <windows.h>, GetModuleHandle,
// resources): the control flow of StartRT/workerThreadRT (renderer.cpp:1490-1610), the ParamsRT -> renderer/light/framebuffer
// set-up that the reference does through a synthetic command line (renderer.cpp:1557-1587 + parseCommandLine :974-1403),
// outputMode's cube-face loop and 12W x H strip assembly (renderer.cpp:508-737), and the status tracker (:99-233).
// Output files: <dae dir>/<dae name>_<camera name>.jpg as in the reference when the back end is device_cuda: the 12W x H strip is
// assembled, watermarked and JPEG-encoded on the GPU through device_cuda's optional hooks (include/yrt_device.h yrtxStrip*, SURVEY
// §8f-2), no frame is mapped to the host. Any other back end (the CPU reference in the tests) gets the host assembly below and a
// .ppm (the reference's FreeImage / libjpeg-turbo are Windows binaries in the mount). The watermark is a Win32 resource in the
// reference; here it is the PNG named by YULIO_RT_WATERMARK, or watermark.png next to this library.
#include <dlfcn.h>

#include <atomic>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "device/device.h"
#include "device/handle.h"
#include "device/loaders/loaders.h"
#include "sys/filename.h"

#include "YulioRT.h"

namespace embree {

// ---- back-end selection: what devices/device/device.cpp:24-48 does, plus the "cuda" line a maintainer adds ----------------
typedef Device* (*create_device_func)(const char* parms, size_t numThreads, int threadsPriority, const char* rtcore_cfg);

// device_cuda's optional strip hooks (NULL with any other back end)
struct StripHooks {
    void* native = nullptr;
    int (*setReadback)(void*, int) = nullptr;
    int (*begin)(void*, size_t, size_t) = nullptr;
    int (*setWatermark)(void*, const char*) = nullptr;
    int (*addFace)(void*, void*, int, int) = nullptr;
    int (*encode)(void*, int, int, const char*) = nullptr;
    // yrtxRenderCubeMap: the 12 cameras of a viewpoint as one render call (include/yrt_device.h)
    int (*renderCubeMap)(void*, void*, void* const*, size_t, void*, void*, void* const*, int) = nullptr;
    const char* (*lastError)() = nullptr;
    bool ok() const { return native && begin && addFace && encode; }
};
static StripHooks g_hooks;

static std::string ownDirectory() {
    Dl_info info;
    if (dladdr((void*)&ownDirectory, &info) && info.dli_fname) {
        const std::string p(info.dli_fname);
        const size_t k = p.find_last_of('/');
        return k == std::string::npos ? "." : p.substr(0, k);
    }
    return ".";
}

Device* Device::rtCreateDevice(const char* type, size_t numThreads, int threadsPriority, const char* rtcore_cfg) {
    std::string file;
    const char* forced = getenv("YULIO_RT_DEVICE_LIB");           // tests: run the same front end on the CPU reference back end
    if (forced && *forced) file = forced;
    else if (!strcmp(type, "default") || !strcmp(type, "cuda")) file = ownDirectory() + "/libdevice_cuda.so";
    else if (!strcmp(type, "singleray")) file = ownDirectory() + "/libdevice_singleray.so";
    else throw std::runtime_error("unknown device: " + std::string(type));
    void* lib = dlopen(file.c_str(), RTLD_NOW);
    if (!lib) throw std::runtime_error("failed loading library \"" + file + "\": " + dlerror());
    create_device_func f = (create_device_func)dlsym(lib, "create");
    if (!f) throw std::runtime_error("invalid device library");
    Device* dev = f("", numThreads, threadsPriority, rtcore_cfg);
    if (!dev) throw std::runtime_error("device creation failed");
    g_hooks = StripHooks();
    if (void* (*nat)(Device*) = (void* (*)(Device*))dlsym(lib, "device_cuda_native")) {
        g_hooks.native = nat(dev);
        g_hooks.setReadback = (int (*)(void*, int))dlsym(lib, "yrtxSetReadback");
        g_hooks.begin = (int (*)(void*, size_t, size_t))dlsym(lib, "yrtxStripBegin");
        g_hooks.setWatermark = (int (*)(void*, const char*))dlsym(lib, "yrtxStripSetWatermark");
        g_hooks.addFace = (int (*)(void*, void*, int, int))dlsym(lib, "yrtxStripAddFace");
        g_hooks.encode = (int (*)(void*, int, int, const char*))dlsym(lib, "yrtxStripEncodeJPEG");
        g_hooks.renderCubeMap = (int (*)(void*, void*, void* const*, size_t, void*, void*, void* const*, int))dlsym(lib, "yrtxRenderCubeMap");
        g_hooks.lastError = (const char* (*)())dlsym(lib, "yrtGetLastError");
    }
    return dev;
}

}  // namespace embree

namespace Yulio {
using namespace embree;

// ---- status tracker (renderer.cpp:99-233): state machine + progress = (stage + tile fraction) / stages ----------------------
class StatusTracker {
    std::mutex m; int stages = 0, stage = 0; StatusRT st{Inactive, 0.f, NoError};
public:
    void reset() { std::lock_guard<std::mutex> l(m); st = StatusRT{Inactive, 0.f, NoError}; stages = stage = 0; }
    void init(int n) { std::lock_guard<std::mutex> l(m); stages = n; stage = 0; }
    void setState(StateRT s) { std::lock_guard<std::mutex> l(m); st.state = s; if (s == Stopped || s == Done) st.progress = 1.f; if (s == Inactive) { st.progress = 0.f; st.lastError = NoError; } }
    void setStage(int s) { std::lock_guard<std::mutex> l(m); if (s < stages) stage = s; }
    void stageProgress(float p) { std::lock_guard<std::mutex> l(m); if (stages > 0) st.progress = float(stage) / stages + p / float(stages); }
    void addError(ErrorCodeRT e) { std::lock_guard<std::mutex> l(m); st.lastError = e; }
    StatusRT get() { std::lock_guard<std::mutex> l(m); return st; }
};
static StatusTracker g_status;
static std::atomic<bool> g_running{false}, g_stop{false}, g_keep{false};
static std::thread g_worker;

static std::atomic<int> g_stageSpan{1};            // stages (cube faces) covered by the render call in flight
static void rendererStatus(const RendererStatus& s) { g_status.stageProgress(s.progress * float(g_stageSpan.load())); }   // rsc, renderer.cpp:231-233

struct Session {                                   // the g_* globals of renderer.cpp:243-300, per StartRT call
    Device* device = nullptr;
    std::string sceneFile, faceCullingMode; ParamsRT p;
    std::vector<Handle<Device::RTPrimitive>> prims; std::vector<Handle<Device::RTCamera>> cameras;
    Handle<Device::RTRenderer> renderer = nullptr; Handle<Device::RTToneMapper> tonemapper = nullptr; Handle<Device::RTFrameBuffer> frameBuffer = nullptr;
    float sceneScale = 1.f;
};

static void writePPM(const std::string& file, const unsigned char* rgb, size_t w, size_t h) {
    FILE* f = fopen(file.c_str(), "wb");
    if (!f) throw std::runtime_error("cannot open " + file);
    fprintf(f, "P6\n%zu %zu\n255\n", w, h);
    fwrite(rgb, 3, w * h, f);
    fclose(f);
}

static const char* faceNames[6] = {"front_image_", "right_image_", "back_image_", "left_image_", "top_image_", "bottom_image_"};

// outputMode's stereo branch (renderer.cpp:518-737)
static void renderCubeMaps(Session& S) {
    Device* dev = S.device;
    g_status.setState(Rendering);
    Handle<Device::RTScene> scene = dev->rtNewScene("default");                      // createScene, renderer.cpp:334-343
    dev->rtSetString(scene, "accel", "default"); dev->rtSetString(scene, "builder", "default"); dev->rtSetString(scene, "traverser", "default");
    for (size_t i = 0; i < S.prims.size(); i++) dev->rtSetPrimitive(scene, i, S.prims[i]);
    dev->rtCommit(scene);
    g_status.init((int)S.cameras.size());
    const size_t W = (size_t)S.p.size, H = (size_t)S.p.size;
    const FileName fn(S.sceneFile);
    const std::string base = std::string(fn.path()) + "/" + fn.name() + "_";
    std::vector<std::vector<unsigned char>> faces;
    std::vector<std::string> saved;
    const Vector3f camUp(0.f, 1.f, 0.f);                                              // g_camUp default, renderer.cpp:246
    const bool gpuStrip = g_hooks.ok() && !getenv("YULIO_RT_HOST_STRIP");
    if (gpuStrip) {
        if (g_hooks.setReadback) g_hooks.setReadback(g_hooks.native, 0);             // frames stay on the GPU
        if (S.p.waterMark && g_hooks.setWatermark) {
            const char* wm = getenv("YULIO_RT_WATERMARK");
            const std::string file = wm && *wm ? wm : ownDirectory() + "/watermark.png";
            if (g_hooks.setWatermark(g_hooks.native, file.c_str()) != 0) printf("yulio_rt: no watermark (%s)\n", g_hooks.lastError ? g_hooks.lastError() : file.c_str());
        }
    } else if (S.p.waterMark) printf("yulio_rt: the watermark is applied by device_cuda's strip path only\n");
    // device_cuda: the 12 stereo cube cameras of a viewpoint share their origin (ColladaLoader.cpp:470-505), so the camera-aligned
    // primitives turn once per viewpoint and all 12 faces are rendered by ONE call, yrtxRenderCubeMap, as one wavefront; the frames go
    // from their frame buffers into the strip on the device. YULIO_RT_PER_FACE=1 keeps the reference's literal per-face loop below.
    const bool batched = gpuStrip && g_hooks.renderCubeMap && S.cameras.size() % 12 == 0 && !getenv("YULIO_RT_PER_FACE");
    if (batched) {
        std::vector<Handle<Device::RTFrameBuffer>> fbs;
        fbs.push_back(S.frameBuffer);
        for (int f = 1; f < 12; f++) fbs.push_back(dev->rtNewFrameBuffer("RGB8", W, H, 1));
        auto hook = [&](int rc) { if (rc != 0) throw std::runtime_error(g_hooks.lastError ? g_hooks.lastError() : "device_cuda hook failed"); };
        for (size_t v = 0; v + 12 <= S.cameras.size() && !g_stop; v += 12) {
            g_status.setStage((int)v); g_stageSpan = 12;
            Vector3f camPos;
            dev->rtGetFloat3(S.cameras[v], "origin", camPos.x, camPos.y, camPos.z);
            for (size_t j = 0; j < S.prims.size(); ++j) dev->rtUpdatePrimitive(scene, j, S.prims[j], camPos, camUp);
            dev->rtCommit(scene);
            void* cams[12]; void* bufs[12];
            for (int f = 0; f < 12; f++) {
                if (S.p.toeIn) { dev->rtSetBool1(S.cameras[v + f], "toeIn", true); dev->rtCommit(S.cameras[v + f]); }
                cams[f] = (void*)(Device::RTCamera)S.cameras[v + f]; bufs[f] = (void*)(Device::RTFrameBuffer)fbs[f];
            }
            hook(g_hooks.renderCubeMap(g_hooks.native, (void*)(Device::RTRenderer)S.renderer, cams, 12, (void*)(Device::RTScene)scene,
                                       (void*)(Device::RTToneMapper)S.tonemapper, bufs, 0));
            g_stageSpan = 1;
            if (g_stop) { if (!g_keep) for (const auto& f : saved) remove(f.c_str()); break; }
            std::string cameraName;
            dev->rtGetString(S.cameras[v], "name", cameraName);
            hook(g_hooks.begin(g_hooks.native, W, H));
            for (int f = 0; f < 12; f++) {
                dev->rtSwapBuffers(fbs[f]);
                hook(g_hooks.addFace(g_hooks.native, bufs[f], f, S.p.waterMark ? 1 : 0));
                if (S.p.debug) {
                    const std::string file = base + cameraName + "_" + faceNames[f % 6] + (f < 6 ? "left" : "right") + ".jpg";
                    hook(g_hooks.encode(g_hooks.native, f, S.p.jpegQuality, file.c_str())); saved.push_back(file);
                }
            }
            const std::string file = base + cameraName + ".jpg";
            hook(g_hooks.encode(g_hooks.native, -1, S.p.jpegQuality, file.c_str())); saved.push_back(file);
            printf("Generated stereoscopic cube map #%zu in file %s\n", v / 12 + 1, file.c_str());
        }
        g_status.setState(g_stop ? Stopped : Done);
        return;
    }
    for (size_t i = 0; i < S.cameras.size() && !g_stop; ++i) {
        g_status.setStage((int)i);
        const Handle<Device::RTCamera>& cam = S.cameras[i];
        Vector3f camPos;
        dev->rtGetFloat3(cam, "origin", camPos.x, camPos.y, camPos.z);               // dynamic geometry, renderer.cpp:551-559
        for (size_t j = 0; j < S.prims.size(); ++j) dev->rtUpdatePrimitive(scene, j, S.prims[j], camPos, camUp);
        dev->rtCommit(scene);
        std::string cameraName;
        dev->rtGetString(cam, "name", cameraName);
        const size_t face = i % 12;
        if (face == 0) faces.clear();
        if (S.p.toeIn) { dev->rtSetBool1(cam, "toeIn", true); dev->rtCommit(cam); }  // renderer.cpp:571-576
        dev->rtRenderFrame(S.renderer, cam, scene, S.tonemapper, S.frameBuffer, 0);
        if (g_stop) {
            // Stopped inside this face: it is not mapped or saved. The reference goes on to rtMapFrameBuffer here (renderer.cpp:620-626), and
            // with its CPU device that can wait forever: when every worker sees the stop flag at the top of its tile loop nobody calls
            // finishTile(forceFinish) and FrameBuffer::wait() never wakes (integratorrenderer.cpp:126,176; api/framebuffer.h:61-77).
            if (!g_keep) for (const auto& f : saved) remove(f.c_str());
            break;
        }
        dev->rtSwapBuffers(S.frameBuffer);
        if (gpuStrip) {
            // frame -> its strip segment on the device (watermark on faces 0-3), JPEG from the device: renderer.cpp:620-718
            auto hook = [&](int rc) { if (rc != 0) throw std::runtime_error(g_hooks.lastError ? g_hooks.lastError() : "device_cuda strip hook failed"); };
            if (face == 0) hook(g_hooks.begin(g_hooks.native, W, H));
            hook(g_hooks.addFace(g_hooks.native, (void*)(Device::RTFrameBuffer)S.frameBuffer, (int)face, S.p.waterMark ? 1 : 0));
            if (S.p.debug) {
                const std::string f = base + cameraName + "_" + faceNames[face % 6] + (face < 6 ? "left" : "right") + ".jpg";
                hook(g_hooks.encode(g_hooks.native, (int)face, S.p.jpegQuality, f.c_str())); saved.push_back(f);
            }
            if (face == 11) {
                const std::string f = base + cameraName + ".jpg";
                hook(g_hooks.encode(g_hooks.native, -1, S.p.jpegQuality, f.c_str())); saved.push_back(f);
                printf("Generated stereoscopic cube map #%zu in file %s\n", i / 12 + 1, f.c_str());
            }
        } else {
        const unsigned char* px = (const unsigned char*)dev->rtMapFrameBuffer(S.frameBuffer);
        const size_t stride = (3 * W + 3) / 4 * 4;                                    // api/framebuffer.h:195
        std::vector<unsigned char> img(W * H * 3);
        for (size_t y = 0; y < H; y++) memcpy(&img[y * W * 3], px + y * stride, W * 3);
        dev->rtUnmapFrameBuffer(S.frameBuffer);
        if (S.p.debug) {
            const std::string f = base + cameraName + "_" + faceNames[face % 6] + (face < 6 ? "left" : "right") + ".ppm";
            writePPM(f, img.data(), W, H); saved.push_back(f);
        }
        faces.push_back(std::move(img));
        if (face == 11) {
            // 12W x H strip: Left Right Up Down Back Front; segments 0-5 take cameras 6-11, segments 6-11 cameras 0-5 (renderer.cpp:677-710)
            static const size_t order[6] = {3, 1, 4, 5, 2, 0};
            std::vector<unsigned char> strip(12 * W * H * 3);
            for (size_t seg = 0; seg < 12; seg++) {
                const size_t src = 6 * (seg / 6 == 0 ? 1 : 0) + order[seg % 6];
                for (size_t y = 0; y < H; y++) memcpy(&strip[(y * 12 * W + seg * W) * 3], &faces[src][y * W * 3], W * 3);
            }
            const std::string f = base + cameraName + ".ppm";
            writePPM(f, strip.data(), 12 * W, H); saved.push_back(f);
            printf("Generated stereoscopic cube map #%zu in file %s\n", i / 12 + 1, f.c_str());
        }
        }
        if (g_stop) {                                                                 // renderer.cpp:728-736
            if (!g_keep) for (const auto& f : saved) remove(f.c_str());
            break;
        }
    }
    g_status.setState(g_stop ? Stopped : Done);
}

static void worker(Session* Sp) {
    Session& S = *Sp;
    Device* dev = S.device;
    try {
        {   // workerThreadRT, renderer.cpp:1490-1521
            std::vector<Handle<Device::RTPrimitive>> prims = rtLoadScene(S.sceneFile, &S.cameras, S.faceCullingMode);
            S.prims.insert(S.prims.end(), prims.begin(), prims.end());
        }
        if (S.cameras.empty()) g_status.addError(InvalidColladaFormat);
        else {
            dev->rtGetFloat1(S.cameras[0], "sceneScale", S.sceneScale);
            // the synthetic command line of StartRT, in its order (renderer.cpp:1557-1587): -renderer, -spp, -size, -depth,
            // -tMaxShadowRay (x scene scale, :1237-1240), -ambientlight (:1026-1032)
            const std::string r = S.p.renderer ? S.p.renderer : "pathtracer";
            if (r != "pt" && r != "pathtracer") throw std::runtime_error("(when parsing -renderer) : unknown renderer: " + r);
            S.renderer = dev->rtNewRenderer("pathtracer");                            // parsePathTracer, renderer.cpp:414-442
            dev->rtSetFloat1(S.renderer, "tMaxShadowRay", std::numeric_limits<float>::infinity());
            dev->rtSetInt1(S.renderer, "sampler.spp", 1);
            dev->rtSetPointer(S.renderer, "stopFlag", &g_stop);
            dev->rtSetPointer(S.renderer, "statusCallback", (void*)&rendererStatus);
            dev->rtCommit(S.renderer);
            dev->rtSetInt1(S.renderer, "sampler.spp", S.p.spp); dev->rtCommit(S.renderer);
            S.frameBuffer = dev->rtNewFrameBuffer("RGB8", (size_t)S.p.size, (size_t)S.p.size, 1);
            dev->rtSetInt1(S.renderer, "maxDepth", S.p.depth); dev->rtCommit(S.renderer);
            dev->rtSetFloat1(S.renderer, "tMaxShadowRay", S.p.tMaxShadowRay * S.sceneScale); dev->rtCommit(S.renderer);
            {
                Handle<Device::RTLight> light = dev->rtNewLight("ambientlight");
                dev->rtSetFloat3(light, "L", S.p.ambientlight[0], S.p.ambientlight[1], S.p.ambientlight[2]);
                dev->rtCommit(light);
                S.prims.push_back(dev->rtNewLightPrimitive(light, nullptr, nullptr));
            }
            renderCubeMaps(S);
        }
    } catch (const std::exception& e) {
        fprintf(stderr, "yulio_rt: %s\n", e.what());
        g_status.addError(UnknownError);
        g_status.setState(Stopped);
    }
    // clearGlobalObjects (renderer.cpp:371-387): handles first, then the device
    S.prims.clear(); S.cameras.clear(); S.renderer = nullptr; S.tonemapper = nullptr; S.frameBuffer = nullptr;
    rtClearTextureCache(); rtClearImageCache();
    delete dev; g_device = nullptr;
    delete Sp;
}

DllApi bool StartRT(const char* colladaFile, const ParamsRT* params) {
    if (g_running) { g_status.addError(RenderingIsInProgress); return false; }       // renderer.cpp:1524-1527
    g_status.reset();
    if (!colladaFile) { g_status.addError(MissingColladaFile); return false; }
    g_status.setState(Initialiazing);
    const FileName fn(colladaFile);
    if (fn.ext() != "dae") { g_status.addError(MissingColladaFile); return false; }  // renderer.cpp:1542-1547
    Session* S = new Session();
    S->sceneFile = colladaFile;
    if (params) S->p = *params;
    S->faceCullingMode = S->p.faceCullingMode ? S->p.faceCullingMode : "default";
    try {
        const char* type = getenv("YULIO_RT_DEVICE");
        const char* cfg = getenv("YULIO_RT_CFG");
        if (!g_device) g_device = Device::rtCreateDevice(type && *type ? type : "default", 0, S->p.threadsPriority, cfg ? cfg : "");
        S->device = g_device;
        S->tonemapper = g_device->rtNewToneMapper("default");                         // createGlobalObjects, renderer.cpp:352-369
        g_device->rtSetFloat1(S->tonemapper, "gamma", 1.0f); g_device->rtSetBool1(S->tonemapper, "vignetting", false);
        g_device->rtCommit(S->tonemapper);
    } catch (const std::exception& e) {
        fprintf(stderr, "yulio_rt: %s\n", e.what());
        g_status.addError(UnknownError); delete S; return false;
    }
    g_stop = false;
    g_worker = std::thread(worker, S);
    g_running = g_worker.joinable();
    return g_running;
}

DllApi bool WaitRT() {
    if (!g_running) return false;
    g_worker.join(); g_running = false; g_stop = false;
    return true;
}

DllApi bool StopRT(bool keepResults) {
    if (!g_running) return false;
    g_keep = keepResults; g_stop = true;
    g_worker.join(); g_running = false; g_stop = false;
    return true;
}

DllApi ErrorCodeRT GetLastErrorRT() { return g_status.get().lastError; }
DllApi void GetCurrentStatusRT(StatusRT* status) { if (status) *status = g_status.get(); }

}  // namespace Yulio
