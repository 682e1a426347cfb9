// group_api.cu — device_cuda on all the GPUs of one box, in ONE process, behind the unchanged Device API.
//
// `yrtCreateDevice(..., cfg "gpus=N")` (rtcore_cfg of the reference's rtCreateDevice, devices/renderer/renderer.cpp:922-937) returns a group
// device: N ordinary device_cuda instances, one per GPU, created with the reference's own distributed partition
// serverID = i, serverCount = N (4-row bands dealt round-robin, devices/device_singleray/api/swapchain.h:57-70). It is the in-process,
// NVLink-era counterpart of the reference's TCP `device_network`: every API call is replayed on every member
// (devices/device_network/network_device.cpp:125-164), so the scene is replicated per GPU; rtRenderFrame and rtCommit(scene) run on all
// members concurrently (one host thread each); rtMapFrameBuffer collects the members' bands into the caller's frame
// (network_device.cpp:235-310, 647-662). No reduction: the bands are disjoint. The frame of a group of N equals the frame of the
// reference's network device with N servers (per-tile sample-set LCGs are seeded with the server id, integratorrenderer.cpp:134), which
// tests/test_gpu_group.py checks bit for bit against N separately created members.
//
// The public entry points (include/yrt_device.h) are generated (tools/gen_group_api.py -> gen/group_wrappers.inc): a plain device goes
// straight to the single-GPU implementation yrtX_core (host_api.cu), a group either runs the call on every member or lands in grp:: here.
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../include/yrt_device.h"
#include "device_impl.hpp"
#include "gen/core_decls.inc"

namespace yrt { extern thread_local std::string g_lastError;
void strip_add_face_device(yrt_device* dev, const unsigned char* devRgb, size_t strideBytes, size_t w, size_t h, int cubeFaceIndex, int watermark); }

namespace grp {

static const uint32_t GROUP_MAGIC = 0x59475250u;      // "YGRP"

struct Handle {                                        // one API object on every member
    uint32_t magic = GROUP_MAGIC; std::atomic<int> refs{1}; std::vector<yrt_handle> m;
    // frame buffers only
    bool isFrameBuffer = false; int format = 2; size_t width = 0, height = 0, depth = 1, cur = 0, strideBytes = 0;
    std::vector<void*> host; std::vector<bool> owned;
    // the assembled frame lives on member 0: `staging` receives every member's compacted bands over NVLink (one cudaMemcpyPeerAsync
    // per member), k_interleave_bands writes them to their raster rows in `full`
    int gpu0 = 0; unsigned char* staging = nullptr; unsigned char* full = nullptr; size_t memberBytes = 0; uint64_t hostCopies = 0, peerCopies = 0;
    ~Handle() {
        for (size_t i = 0; i < host.size(); i++) if (owned[i]) cudaFreeHost(host[i]);
        if (staging || full) { cudaSetDevice(gpu0); if (staging) cudaFree(staging); if (full) cudaFree(full); }
    }
};

inline bool is(const yrt_device* d) { return d && !d->members.empty(); }
inline Handle* gh(yrt_handle h) {
    Handle* g = (Handle*)h;
    if (!g || g->magic != GROUP_MAGIC) throw std::runtime_error("invalid handle");
    return g;
}
inline yrt_handle un(yrt_handle h, int i) { return h ? gh(h)->m[(size_t)i] : nullptr; }

static yrt_status fail(const std::exception& e) { yrt::g_lastError = e.what(); return YRT_ERROR; }

// runs f(member, index) on every member, in order; first failure wins
template <class F> yrt_status each(yrt_device* dev, F f) {
    try {
        std::lock_guard<std::mutex> lock(dev->mutex);
        for (size_t i = 0; i < dev->members.size(); i++)
            if (f(dev->members[i], (int)i) != YRT_OK) return YRT_ERROR;          // the member left its message in g_lastError (same thread)
        return YRT_OK;
    } catch (const std::exception& e) { return fail(e); }
}
// One persistent host thread per member: render, scene commit and frame read-back run on all GPUs at once, and a cube face of a small
// scene is only a few milliseconds of GPU work — spawning threads per call would show (3 x N thread starts per face).
struct Workers {
    struct Slot { std::thread th; std::mutex m; std::condition_variable cv; std::function<void()> job; bool stop = false; };
    std::vector<std::unique_ptr<Slot>> slots; std::mutex doneM; std::condition_variable doneCv; size_t pending = 0;
    explicit Workers(size_t n) {
        for (size_t i = 0; i < n; i++) {
            slots.emplace_back(new Slot());
            Slot* s = slots.back().get();
            s->th = std::thread([this, s] {
                for (;;) {
                    std::function<void()> job;
                    { std::unique_lock<std::mutex> l(s->m); s->cv.wait(l, [s] { return s->stop || s->job; }); if (s->stop) return; job.swap(s->job); }
                    job();
                    { std::lock_guard<std::mutex> l(doneM); pending--; }
                    doneCv.notify_all();
                }
            });
        }
    }
    ~Workers() {
        for (auto& s : slots) { { std::lock_guard<std::mutex> l(s->m); s->stop = true; } s->cv.notify_all(); }
        for (auto& s : slots) s->th.join();
    }
    void run(const std::function<void(size_t)>& f) {       // f(i) on worker i for every i; returns when all are done
        { std::lock_guard<std::mutex> l(doneM); pending = slots.size(); }
        for (size_t i = 0; i < slots.size(); i++) { { std::lock_guard<std::mutex> l(slots[i]->m); slots[i]->job = [&f, i] { f(i); }; } slots[i]->cv.notify_all(); }
        std::unique_lock<std::mutex> l(doneM); doneCv.wait(l, [this] { return pending == 0; });
    }
};
static std::mutex g_workersM; static std::vector<std::pair<yrt_device*, std::unique_ptr<Workers>>> g_workers;
static Workers& workers(yrt_device* dev) {
    std::lock_guard<std::mutex> l(g_workersM);
    for (auto& w : g_workers) if (w.first == dev) return *w.second;
    g_workers.emplace_back(dev, std::unique_ptr<Workers>(new Workers(dev->members.size())));
    return *g_workers.back().second;
}
static void drop_workers(yrt_device* dev) {
    std::lock_guard<std::mutex> l(g_workersM);
    for (size_t i = 0; i < g_workers.size(); i++) if (g_workers[i].first == dev) { g_workers.erase(g_workers.begin() + (long)i); return; }
}

template <class F> yrt_status each_parallel(yrt_device* dev, F f) {
    const size_t n = dev->members.size();
    std::vector<std::string> err(n); std::vector<int> rc(n, YRT_OK);
    workers(dev).run([&](size_t i) {
        try { rc[i] = f(dev->members[i], (int)i); if (rc[i] != YRT_OK) err[i] = yrtGetLastError_core(); }
        catch (const std::exception& e) { rc[i] = YRT_ERROR; err[i] = e.what(); }
    });
    for (size_t i = 0; i < n; i++) if (rc[i] != YRT_OK) { yrt::g_lastError = err[i]; return YRT_ERROR; }
    return YRT_OK;
}
template <class F> yrt_handle make(yrt_device* dev, F f) {
    try {
        std::lock_guard<std::mutex> lock(dev->mutex);
        std::unique_ptr<Handle> g(new Handle());
        for (size_t i = 0; i < dev->members.size(); i++) {
            yrt_handle h = f(dev->members[i], (int)i);
            if (!h) { for (size_t k = 0; k < g->m.size(); k++) yrtDecRef_core(dev->members[k], g->m[k]); return nullptr; }
            g->m.push_back(h);
        }
        return g.release();
    } catch (const std::exception& e) { fail(e); return nullptr; }
}

// ---- device ----------------------------------------------------------------------------------------------------------------
static long cfg_int(const std::string& cfg, const char* key, long def) {
    size_t pos = 0; const std::string k = std::string(key) + "=";
    while (pos < cfg.size()) {
        size_t end = cfg.find(',', pos); if (end == std::string::npos) end = cfg.size();
        std::string item = cfg.substr(pos, end - pos);
        while (!item.empty() && item[0] == ' ') item.erase(0, 1);
        if (item.compare(0, k.size(), k) == 0) return strtol(item.c_str() + k.size(), nullptr, 10);
        pos = end + 1;
    }
    return def;
}

// cfg without the keys the group assigns per member (the first match of a key wins in cfg_int: a user's "gpu=" or "serverID=" must not
// reach the members, or all of them would land on one GPU / render the same bands)
static std::string cfg_without_member_keys(const std::string& cfg) {
    std::string out; size_t pos = 0;
    while (pos < cfg.size()) {
        size_t end = cfg.find(',', pos); if (end == std::string::npos) end = cfg.size();
        std::string item = cfg.substr(pos, end - pos);
        std::string key = item.substr(0, item.find('='));
        while (!key.empty() && key[0] == ' ') key.erase(0, 1);
        if (!item.empty() && key != "gpus" && key != "gpu" && key != "serverID" && key != "serverCount") { if (!out.empty()) out += ","; out += item; }
        pos = end + 1;
    }
    return out;
}

yrt_device* yrtCreateDevice(const char* parms, size_t numThreads, int threadsPriority, const char* cfg) {
    const std::string c(cfg ? cfg : "");
    const long n = cfg_int(c, "gpus", 1);
    if (n <= 1) return yrtCreateDevice_core(parms, numThreads, threadsPriority, cfg);
    try {
        int have = 0; cudaGetDeviceCount(&have);
        if (n > have) throw std::runtime_error("device_cuda: cfg gpus=" + std::to_string(n) + " but only " + std::to_string(have) + " CUDA device(s) are visible");
        const long first = cfg_int(c, "gpu", 0);
        if (first < 0 || first + n > have) throw std::runtime_error("device_cuda: cfg gpu=" + std::to_string(first) + ",gpus=" + std::to_string(n) + " exceeds the " + std::to_string(have) + " visible CUDA device(s)");
        const std::string rest = cfg_without_member_keys(c);
        std::unique_ptr<yrt_device> g(new yrt_device());
        // the members' CUDA contexts are created concurrently (about 1.5 s each)
        std::vector<yrt_device*> ms((size_t)n, nullptr); std::vector<std::string> err((size_t)n); std::vector<std::thread> th;
        for (long i = 0; i < n; i++)
            th.emplace_back([&, i] {
                const std::string mc = "gpus=1,gpu=" + std::to_string(first + i) + ",serverID=" + std::to_string(i) + ",serverCount=" + std::to_string(n) +
                                       (rest.empty() ? "" : "," + rest);
                ms[(size_t)i] = yrtCreateDevice_core(parms, numThreads, threadsPriority, mc.c_str());
                if (!ms[(size_t)i]) err[(size_t)i] = yrtGetLastError_core();
            });
        for (auto& t : th) t.join();
        for (long i = 0; i < n; i++)
            if (!ms[(size_t)i]) {
                for (yrt_device* d : ms) if (d) yrtDestroyDevice_core(d);
                throw std::runtime_error(err[(size_t)i]);
            }
        g->members = ms;
        // frames are assembled on member 0 over NVLink: peer access from GPU 0 to the others, and no per-member host copies
        cudaSetDevice(ms[0]->gpu);
        for (long i = 1; i < n; i++) {
            int can = 0; cudaDeviceCanAccessPeer(&can, ms[0]->gpu, ms[(size_t)i]->gpu);
            if (can) { const cudaError_t e = cudaDeviceEnablePeerAccess(ms[(size_t)i]->gpu, 0); if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError(); else cudaGetLastError(); }
        }
        for (yrt_device* m : ms) m->readback = false;
        return g.release();
    } catch (const std::exception& e) { fail(e); return nullptr; }
}
void yrtDestroyDevice(yrt_device* dev) {
    drop_workers(dev);
    std::vector<std::thread> th;                       // releasing ~11 GB of wavefront state per GPU takes over a second: all members at once
    for (yrt_device* m : dev->members) th.emplace_back([m] { yrtDestroyDevice_core(m); });
    for (auto& t : th) t.join();
    dev->members.clear();
    delete dev;
}

// ---- objects with special ownership --------------------------------------------------------------------------------------------
yrt_handle yrtNewData(yrt_device* dev, const char* type, size_t bytes, const void* data) {
    // "immutable_managed" hands the caller's allocation over (api/data.h:40-45): member 0 takes it (and frees it with its handle, which is
    // released together with the others'), members 1.. copy from it
    const bool managed = type && !strcasecmp(type, "immutable_managed");
    return make(dev, [&](yrt_device* m, int i) { return yrtNewData_core(m, (managed && i > 0) ? "immutable" : type, bytes, data); });
}

yrt_handle yrtNewFrameBuffer(yrt_device* dev, const char* type, size_t width, size_t height, size_t buffers, void** ptrs) {
    yrt_handle h = make(dev, [&](yrt_device* m, int) { return yrtNewFrameBuffer_core(m, type, width, height, buffers, nullptr); });
    if (!h) return nullptr;
    Handle* g = gh(h);
    g->isFrameBuffer = true; g->width = width; g->height = height; g->depth = buffers ? buffers : 1;
    const std::string t(type ? type : "");
    if (!strcasecmp(t.c_str(), "RGB_FLOAT32")) { g->format = 0; g->strideBytes = width * 12; }          // api/framebuffer.h:106,146,195
    else if (!strcasecmp(t.c_str(), "RGBA8")) { g->format = 1; g->strideBytes = width * 4; }
    else { g->format = 2; g->strideBytes = (3 * width + 3) / 4 * 4; }
    for (size_t i = 0; i < g->depth; i++) {
        void* p = ptrs ? ptrs[i] : nullptr; bool own = false;
        const size_t bytes = g->strideBytes * height;
        if (!p) { YRT_CK(cudaHostAlloc(&p, bytes != 0 ? bytes : 1, cudaHostAllocPortable)); memset(p, 0, bytes); own = true; }
        g->host.push_back(p); g->owned.push_back(own);
    }
    return h;
}

// raster row y of the assembled frame <- buffer row 4 * ((y >> 2) / n) + (y & 3) of member (y >> 2) % n  (api/swapchain.h:57-70)
__global__ void k_interleave_bands(const uint32_t* __restrict__ staging, size_t memberWords, int n, int height, int strideWords, uint32_t* __restrict__ full) {
    const size_t total = (size_t)height * strideWords;
    for (size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += (size_t)gridDim.x * blockDim.x) {
        const int y = (int)(w / strideWords), x = (int)(w % strideWords);
        const int member = (y >> 2) % n, row = 4 * ((y >> 2) / n) + (y & 3);
        full[w] = staging[(size_t)member * memberWords + (size_t)row * strideWords + x];
    }
}

// The members' bands -> the full frame on member 0, GPU to GPU: one cudaMemcpyPeerAsync per member over NVLink, one interleave kernel.
// The members have finished rendering (yrtRenderFrame / yrtxRenderCubeMap return after their streams are idle). Caller holds dev->mutex.
static const unsigned char* assemble_on_member0(yrt_device* dev, Handle* g) {
    const size_t n = dev->members.size();
    yrt_device* m0 = dev->members[0];
    m0->bind();
    const size_t maxRows = 4 * ((((g->height + 3) >> 2) + n - 1) / n);
    g->memberBytes = maxRows * g->strideBytes; g->gpu0 = m0->gpu;
    if (!g->staging) { YRT_CK(cudaMalloc((void**)&g->staging, std::max<size_t>(16, n * g->memberBytes))); YRT_CK(cudaMalloc((void**)&g->full, std::max<size_t>(16, g->height * g->strideBytes))); }
    for (size_t i = 0; i < n; i++) {
        void* src = nullptr; size_t bytes = 0, stride = 0;
        if (yrtxFrameBufferDevice_core(dev->members[i], g->m[i], &src, &bytes, &stride) != YRT_OK) throw std::runtime_error(yrtGetLastError_core());
        m0->bind();
        size_t rows = 0; for (size_t y = 0; y < g->height; y++) if ((((y >> 2) + n - i) % n) == 0) rows++;
        YRT_CK(cudaMemcpyPeerAsync(g->staging + i * g->memberBytes, m0->gpu, src, dev->members[i]->gpu, rows * g->strideBytes, m0->stream));
        g->peerCopies++;
    }
    const size_t words = g->height * g->strideBytes / 4;
    k_interleave_bands<<<(unsigned)std::min<size_t>(4096, (words + 255) / 256 + 1), 256, 0, m0->stream>>>((const uint32_t*)g->staging, g->memberBytes / 4, (int)n, (int)g->height,
                                                                                                          (int)(g->strideBytes / 4), (uint32_t*)g->full);
    YRT_CK(cudaGetLastError());
    return g->full;
}

void* yrtMapFrameBuffer(yrt_device* dev, yrt_handle fb, int bufID) {
    try {
        Handle* g = gh(fb);
        if (!g->isFrameBuffer) throw std::runtime_error("invalid framebuffer handle");
        std::lock_guard<std::mutex> lock(dev->mutex);
        const size_t buf = bufID < 0 ? g->cur : (size_t)bufID % g->depth;
        unsigned char* out = (unsigned char*)g->host[buf];
        const unsigned char* full = assemble_on_member0(dev, g);
        yrt_device* m0 = dev->members[0];
        YRT_CK(cudaMemcpyAsync(out, full, g->height * g->strideBytes, cudaMemcpyDeviceToHost, m0->stream));    // the ONE host copy of the frame
        YRT_CK(cudaStreamSynchronize(m0->stream));
        g->hostCopies++; dev->stats.d2h_bytes += g->height * g->strideBytes;
        return out;
    } catch (const std::exception& e) { fail(e); return nullptr; }
}
yrt_status yrtUnmapFrameBuffer(yrt_device*, yrt_handle fb, int) { try { gh(fb); return YRT_OK; } catch (const std::exception& e) { return fail(e); } }
yrt_status yrtSwapBuffers(yrt_device* dev, yrt_handle fb) {
    try { Handle* g = gh(fb); g->cur = (g->cur + 1) % g->depth; } catch (const std::exception& e) { return fail(e); }
    return each(dev, [&](yrt_device* m, int i) { return yrtSwapBuffers_core(m, un(fb, i)); });
}

yrt_status yrtIncRef(yrt_device* dev, yrt_handle h) {
    try { gh(h)->refs++; } catch (const std::exception& e) { return fail(e); }
    return each(dev, [&](yrt_device* m, int i) { return yrtIncRef_core(m, un(h, i)); });
}
yrt_status yrtDecRef(yrt_device* dev, yrt_handle h) {
    Handle* g; try { g = gh(h); } catch (const std::exception& e) { return fail(e); }
    const yrt_status rc = each(dev, [&](yrt_device* m, int i) { return yrtDecRef_core(m, g->m[(size_t)i]); });
    if (g->refs.fetch_sub(1) == 1) { g->magic = 0; delete g; }
    return rc;
}

// ---- queries: every member holds the same parameters, member 0 answers ------------------------------------------------------------
static yrt_device* m0(yrt_device* dev) { return dev->members[0]; }
yrt_status yrtGetFloat1(yrt_device* dev, yrt_handle h, const char* p, float* x) { try { return yrtGetFloat1_core(m0(dev), un(h, 0), p, x); } catch (const std::exception& e) { return fail(e); } }
yrt_status yrtGetFloat3(yrt_device* dev, yrt_handle h, const char* p, float* x, float* y, float* z) { try { return yrtGetFloat3_core(m0(dev), un(h, 0), p, x, y, z); } catch (const std::exception& e) { return fail(e); } }
yrt_status yrtGetString(yrt_device* dev, yrt_handle h, const char* p, char* buf, size_t n) { try { return yrtGetString_core(m0(dev), un(h, 0), p, buf, n); } catch (const std::exception& e) { return fail(e); } }
yrt_status yrtGetTransform(yrt_device* dev, yrt_handle h, const char* p, float* t) { try { return yrtGetTransform_core(m0(dev), un(h, 0), p, t); } catch (const std::exception& e) { return fail(e); } }
int yrtPick(yrt_device* dev, yrt_handle cam, float x, float y, yrt_handle scene, float* px, float* py, float* pz) {
    try { return yrtPick_core(m0(dev), un(cam, 0), x, y, un(scene, 0), px, py, pz); } catch (const std::exception& e) { fail(e); return -1; }
}

// serverID / serverCount are the group's own business; the status callback is made by member 0 only (its fraction of paths retired is the
// group's, the bands are balanced), the stop flag is polled by every member
yrt_status yrtSetInt1(yrt_device* dev, yrt_handle h, const char* p, int x) {
    if (!h) return YRT_OK;
    return each(dev, [&](yrt_device* m, int i) { return yrtSetInt1_core(m, un(h, i), p, x); });
}
yrt_status yrtSetPointer(yrt_device* dev, yrt_handle h, const char* p, void* ptr) {
    const bool cb = p && !strcmp(p, "statusCallback");
    return each(dev, [&](yrt_device* m, int i) { return yrtSetPointer_core(m, un(h, i), p, (cb && i > 0) ? nullptr : ptr); });
}

// ---- the two calls that carry the work: all members at once -------------------------------------------------------------------------
yrt_status yrtCommit(yrt_device* dev, yrt_handle h) {
    try { gh(h); std::lock_guard<std::mutex> lock(dev->mutex); return each_parallel(dev, [&](yrt_device* m, int i) { return yrtCommit_core(m, un(h, i)); }); }
    catch (const std::exception& e) { return fail(e); }
}
yrt_status yrtRenderFrame(yrt_device* dev, yrt_handle renderer, yrt_handle camera, yrt_handle scene, yrt_handle tonemapper, yrt_handle fb, int accumulate) {
    try {
        gh(renderer); gh(camera); gh(scene); gh(tonemapper); gh(fb);
        std::lock_guard<std::mutex> lock(dev->mutex);
        dev->stats.d2h_bytes = 0;
        return each_parallel(dev, [&](yrt_device* m, int i) {
            return yrtRenderFrame_core(m, un(renderer, i), un(camera, i), un(scene, i), un(tonemapper, i), un(fb, i), accumulate); });
    } catch (const std::exception& e) { return fail(e); }
}

yrt_status yrtxRenderCubeMap(yrt_device* dev, yrt_handle renderer, const yrt_handle* cameras, size_t numFaces, yrt_handle scene, yrt_handle tonemapper,
                             const yrt_handle* fbs, int accumulate) {
    try {
        if (!cameras || !fbs || numFaces < 1 || numFaces > YRT_MAX_FACES) throw std::runtime_error("device_cuda: yrtxRenderCubeMap takes 1..12 cameras and frame buffers");
        gh(renderer); gh(scene); gh(tonemapper);
        for (size_t f = 0; f < numFaces; f++) { gh(cameras[f]); gh(fbs[f]); }
        std::lock_guard<std::mutex> lock(dev->mutex);
        dev->stats.d2h_bytes = 0;
        return each_parallel(dev, [&](yrt_device* m, int i) {
            yrt_handle c[YRT_MAX_FACES], b[YRT_MAX_FACES];
            for (size_t f = 0; f < numFaces; f++) { c[f] = un(cameras[f], i); b[f] = un(fbs[f], i); }
            return yrtxRenderCubeMap_core(m, un(renderer, i), c, numFaces, un(scene, i), un(tonemapper, i), b, accumulate); });
    } catch (const std::exception& e) { return fail(e); }
}

yrt_status yrtxMicrobench(yrt_device* dev, int kind, size_t bytes, double* result) { return yrtxMicrobench_core(dev->members[0], kind, bytes, result); }

// ---- extensions -----------------------------------------------------------------------------------------------------------------
yrt_status yrtxGetFrameStats(yrt_device* dev, yrtx_frame_stats* out) {
    if (!out) return YRT_ERROR;
    yrtx_frame_stats a{}; bool first = true;
    for (yrt_device* m : dev->members) {
        yrtx_frame_stats s{};
        if (yrtxGetFrameStats_core(m, &s) != YRT_OK) return YRT_ERROR;
        if (first) { a = s; first = false; continue; }
        // time-like fields: the slowest member; counters: the sum
        a.render_ms = std::max(a.render_ms, s.render_ms); a.build_ms = std::max(a.build_ms, s.build_ms); a.host_ms = std::max(a.host_ms, s.host_ms);
        a.trace_ms = std::max(a.trace_ms, s.trace_ms); a.closest_ms = std::max(a.closest_ms, s.closest_ms); a.shadow_ms = std::max(a.shadow_ms, s.shadow_ms);
        a.shade_ms = std::max(a.shade_ms, s.shade_ms); a.raygen_film_ms = std::max(a.raygen_film_ms, s.raygen_film_ms); a.sort_ms = std::max(a.sort_ms, s.sort_ms);
        a.resolve_ms = std::max(a.resolve_ms, s.resolve_ms); a.miss_ms = std::max(a.miss_ms, s.miss_ms);
        a.node_visits_shadow += s.node_visits_shadow; a.tri_tests_shadow += s.tri_tests_shadow; a.path_vertices += s.path_vertices; a.shade_launches += s.shade_launches; a.errors |= s.errors;
        a.rays_closest += s.rays_closest; a.rays_shadow += s.rays_shadow; a.kernel_launches += s.kernel_launches; a.node_visits += s.node_visits;
        a.tri_tests += s.tri_tests; a.closest_launches += s.closest_launches; a.shadow_launches += s.shadow_launches; a.h2d_bytes += s.h2d_bytes; a.d2h_bytes += s.d2h_bytes;
    }
    a.num_gpus = (uint32_t)dev->members.size();
    a.d2h_bytes += dev->stats.d2h_bytes;               // frames the group itself copied to the host (yrtMapFrameBuffer) since the last render call
    *out = a;
    return YRT_OK;
}
yrt_status yrtxTraceRays(yrt_device* dev, yrt_handle scene, size_t n, const float* rays, void* hits, int closest, int onDevice, float* ms) {
    try { return yrtxTraceRays_core(m0(dev), un(scene, 0), n, rays, hits, closest, onDevice, ms); } catch (const std::exception& e) { return fail(e); }
}
yrt_status yrtxPrimaryRays(yrt_device* dev, yrt_handle r, yrt_handle c, yrt_handle fb, float* rays, int* sets) {
    try { return yrtxPrimaryRays_core(m0(dev), un(r, 0), un(c, 0), un(fb, 0), rays, sets); } catch (const std::exception& e) { return fail(e); }
}
yrt_status yrtxSampleTable(yrt_device* dev, yrt_handle r, yrt_handle s, int it, int* sets, int* spp, int* n1, int* n2, float* table) {
    try { return yrtxSampleTable_core(m0(dev), un(r, 0), un(s, 0), it, sets, spp, n1, n2, table); } catch (const std::exception& e) { return fail(e); }
}
// the assembled frame on member 0 (bands gathered over NVLink); valid until the next render into this frame buffer
yrt_status yrtxFrameBufferDevice(yrt_device* dev, yrt_handle fb, void** devPtr, size_t* bytes, size_t* strideBytes) {
    try {
        Handle* g = gh(fb);
        if (!g->isFrameBuffer) throw std::runtime_error("invalid framebuffer handle");
        std::lock_guard<std::mutex> lock(dev->mutex);
        const unsigned char* full = assemble_on_member0(dev, g);
        YRT_CK(cudaStreamSynchronize(dev->members[0]->stream));
        if (devPtr) *devPtr = (void*)full; if (bytes) *bytes = g->height * g->strideBytes; if (strideBytes) *strideBytes = g->strideBytes;
        return YRT_OK;
    } catch (const std::exception& e) { return fail(e); }
}
// a group's frames reach the host through yrtMapFrameBuffer only (assembled on member 0 first): the members never copy their bands out
yrt_status yrtxSetReadback(yrt_device* dev, int) { for (yrt_device* m : dev->members) m->readback = false; return YRT_OK; }
yrt_status yrtxReadImage(yrt_device* dev, yrt_handle img, int* w, int* h, int* f, void* px) {
    try { return yrtxReadImage_core(m0(dev), un(img, 0), w, h, f, px); } catch (const std::exception& e) { return fail(e); }
}
// the cube-map strip lives on member 0; a face reaches it through the collected host frame
yrt_status yrtxStripBegin(yrt_device* dev, size_t w, size_t h) { return yrtxStripBegin_core(m0(dev), w, h); }
yrt_status yrtxStripSetWatermark(yrt_device* dev, const char* f) { return yrtxStripSetWatermark_core(m0(dev), f); }
yrt_status yrtxStripRead(yrt_device* dev, void* rgb) { return yrtxStripRead_core(m0(dev), rgb); }
yrt_status yrtxStripEncodeJPEG(yrt_device* dev, int face, int q, const char* f) { return yrtxStripEncodeJPEG_core(m0(dev), face, q, f); }
yrt_status yrtxStripAddFace(yrt_device* dev, yrt_handle fb, int face, int watermark) {
    try {
        Handle* g = gh(fb);
        if (g->format != 2) throw std::runtime_error("device_cuda: the strip takes RGB8 frames");
        // bands -> member 0 over NVLink -> strip segment: the frame never touches the host
        std::lock_guard<std::mutex> lock(dev->mutex);
        const unsigned char* full = assemble_on_member0(dev, g);
        std::lock_guard<std::mutex> lock0(m0(dev)->mutex); m0(dev)->bind();
        yrt::strip_add_face_device(m0(dev), full, g->strideBytes, g->width, g->height, face, watermark);
        return YRT_OK;
    } catch (const std::exception& e) { return fail(e); }
}

}  // namespace grp

#include "gen/group_wrappers.inc"
