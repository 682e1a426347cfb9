// device_cuda_plugin.cpp — the C++ side of the drop-in boundary: `libdevice_cuda.so`, a back end of the reference's
// plugin interface. It is compiled against the reference's UNMODIFIED devices/device/device.h and exports the same
// factory symbol every reference back end exports:
//     extern "C" embree::Device* create(const char* parms, size_t numThreads, int threadsPriority, const char* rtcore_cfg);
// (reference: devices/device/device.cpp:24-35 looks it up with getSymbol(lib, "create");
//  devices/device_singleray/api/singleray_device.cpp:105-107 is the CPU back end's definition).
// Every virtual of embree::Device (device.h:126-329) forwards to the C-ABI of include/yrt_device.h; failed calls
// are turned back into the std::runtime_error the reference's callers expect (api/singleray_device.cpp:190..435).
// Nothing else of the reference is linked: the plugin depends only on libyrt_device_cuda.so.
#include <stdexcept>
#include <strings.h>
#include <cstdlib>
#include <string>

#include "device/device.h"
#include "yrt_device.h"

namespace embree {

class CudaDevice : public Device {
    yrt_device* d;
    static void fail(const char* what) { throw std::runtime_error(std::string(what) + ": " + yrtGetLastError()); }
    template <typename T> static T ok(T h, const char* what) { if (!h) fail(what); return h; }
    static void ok(yrt_status s, const char* what) { if (s != YRT_OK) fail(what); }

public:
    explicit CudaDevice(yrt_device* dev) : d(dev) {}
    yrt_device* native() const { return d; }
    ~CudaDevice() override { yrtDestroyDevice(d); }

    RTCamera rtNewCamera(const char* type) override { return (RTCamera)ok(yrtNewCamera(d, type), "rtNewCamera"); }
    // "immutable_managed" hands over a block the reference's callers allocate with embree::alignedMalloc (xml_loader.cpp:227-267,
    // network_server.cpp:138) — an INTERIOR pointer whose malloc base lies ((int*)ptr)[-1] bytes before it (common/sys/platform.cpp:171-188).
    // The C-ABI's managed type owns a plain malloc() block, so here the data is copied and the caller's block released the reference's way.
    RTData rtNewData(const char* type, size_t bytes, const void* data) override {
        if (type && !strcasecmp(type, "immutable_managed")) {
            RTData h = (RTData)ok(yrtNewData(d, "immutable", bytes, data), "rtNewData");
            if (data) { const int ofs = ((const int*)data)[-1]; free((char*)data - ofs); }     // embree::alignedFree
            return h;
        }
        return (RTData)ok(yrtNewData(d, type, bytes, data), "rtNewData");
    }
    RTData rtNewDataFromFile(const char* type, const char* file, size_t offset, size_t bytes) override { return (RTData)ok(yrtNewDataFromFile(d, type, file, offset, bytes), "rtNewDataFromFile"); }
    RTImage rtNewImage(const char* type, size_t width, size_t height, const void* data, const bool copy) override { return (RTImage)ok(yrtNewImage(d, type, width, height, data, copy ? 1 : 0), "rtNewImage"); }
    RTImage rtNewImageFromFile(const char* file) override { return (RTImage)ok(yrtNewImageFromFile(d, file), "rtNewImageFromFile"); }
    RTTexture rtNewTexture(const char* type) override { return (RTTexture)ok(yrtNewTexture(d, type), "rtNewTexture"); }
    RTMaterial rtNewMaterial(const char* type) override { return (RTMaterial)ok(yrtNewMaterial(d, type), "rtNewMaterial"); }
    RTShape rtNewShape(const char* type) override { return (RTShape)ok(yrtNewShape(d, type), "rtNewShape"); }
    RTLight rtNewLight(const char* type) override { return (RTLight)ok(yrtNewLight(d, type), "rtNewLight"); }
    RTPrimitive rtNewShapePrimitive(RTShape shape, RTMaterial material, const float* transform, bool faceCamera) override {
        return (RTPrimitive)ok(yrtNewShapePrimitive(d, shape, material, transform, faceCamera ? 1 : 0), "rtNewShapePrimitive"); }
    RTPrimitive rtNewLightPrimitive(RTLight light, RTMaterial material, const float* transform) override {
        return (RTPrimitive)ok(yrtNewLightPrimitive(d, light, material, transform), "rtNewLightPrimitive"); }
    RTPrimitive rtTransformPrimitive(RTPrimitive prim, const float* transform) override { return (RTPrimitive)ok(yrtTransformPrimitive(d, prim, transform), "rtTransformPrimitive"); }
    RTScene rtNewScene(const char* type) override { return (RTScene)ok(yrtNewScene(d, type), "rtNewScene"); }
    void rtSetPrimitive(RTScene scene, size_t slot, RTPrimitive prim) override { ok(yrtSetPrimitive(d, scene, slot, prim), "rtSetPrimitive"); }
    void rtUpdatePrimitive(RTScene scene, size_t slot, RTPrimitive prim, const Vector3f& camPos, const Vector3f& camUp) override {
        const float p[3] = {camPos.x, camPos.y, camPos.z}, u[3] = {camUp.x, camUp.y, camUp.z};
        ok(yrtUpdatePrimitive(d, scene, slot, prim, p, u), "rtUpdatePrimitive"); }
    RTToneMapper rtNewToneMapper(const char* type) override { return (RTToneMapper)ok(yrtNewToneMapper(d, type), "rtNewToneMapper"); }
    RTRenderer rtNewRenderer(const char* type) override { return (RTRenderer)ok(yrtNewRenderer(d, type), "rtNewRenderer"); }
    RTFrameBuffer rtNewFrameBuffer(const char* type, size_t width, size_t height, size_t buffers, void** ptrs) override {
        return (RTFrameBuffer)ok(yrtNewFrameBuffer(d, type, width, height, buffers, ptrs), "rtNewFrameBuffer"); }
    void* rtMapFrameBuffer(RTFrameBuffer fb, int bufID) override { return ok(yrtMapFrameBuffer(d, fb, bufID), "rtMapFrameBuffer"); }
    void rtUnmapFrameBuffer(RTFrameBuffer fb, int bufID) override { ok(yrtUnmapFrameBuffer(d, fb, bufID), "rtUnmapFrameBuffer"); }
    void rtSwapBuffers(RTFrameBuffer fb) override { ok(yrtSwapBuffers(d, fb), "rtSwapBuffers"); }
    void rtIncRef(RTHandle h) override { ok(yrtIncRef(d, h), "rtIncRef"); }
    void rtDecRef(RTHandle h) override { ok(yrtDecRef(d, h), "rtDecRef"); }

    void rtSetBool1(RTHandle h, const char* p, bool x) override { ok(yrtSetBool1(d, h, p, x), "rtSetBool1"); }
    void rtSetBool2(RTHandle h, const char* p, bool x, bool y) override { ok(yrtSetBool2(d, h, p, x, y), "rtSetBool2"); }
    void rtSetBool3(RTHandle h, const char* p, bool x, bool y, bool z) override { ok(yrtSetBool3(d, h, p, x, y, z), "rtSetBool3"); }
    void rtSetBool4(RTHandle h, const char* p, bool x, bool y, bool z, bool w) override { ok(yrtSetBool4(d, h, p, x, y, z, w), "rtSetBool4"); }
    void rtSetInt1(RTHandle h, const char* p, int x) override { ok(yrtSetInt1(d, h, p, x), "rtSetInt1"); }
    void rtSetInt2(RTHandle h, const char* p, int x, int y) override { ok(yrtSetInt2(d, h, p, x, y), "rtSetInt2"); }
    void rtSetInt3(RTHandle h, const char* p, int x, int y, int z) override { ok(yrtSetInt3(d, h, p, x, y, z), "rtSetInt3"); }
    void rtSetInt4(RTHandle h, const char* p, int x, int y, int z, int w) override { ok(yrtSetInt4(d, h, p, x, y, z, w), "rtSetInt4"); }
    void rtSetPointer(RTHandle h, const char* p, void* ptr) override { ok(yrtSetPointer(d, h, p, ptr), "rtSetPointer"); }
    void rtSetFloat1(RTHandle h, const char* p, float x) override { ok(yrtSetFloat1(d, h, p, x), "rtSetFloat1"); }
    void rtGetFloat1(RTHandle h, const char* p, float& x) override { ok(yrtGetFloat1(d, h, p, &x), "rtGetFloat1"); }
    void rtSetFloat2(RTHandle h, const char* p, float x, float y) override { ok(yrtSetFloat2(d, h, p, x, y), "rtSetFloat2"); }
    void rtSetFloat3(RTHandle h, const char* p, float x, float y, float z) override { ok(yrtSetFloat3(d, h, p, x, y, z), "rtSetFloat3"); }
    void rtGetFloat3(RTHandle h, const char* p, float& x, float& y, float& z) override { ok(yrtGetFloat3(d, h, p, &x, &y, &z), "rtGetFloat3"); }
    void rtSetFloat4(RTHandle h, const char* p, float x, float y, float z, float w) override { ok(yrtSetFloat4(d, h, p, x, y, z, w), "rtSetFloat4"); }
    void rtSetArray(RTHandle h, const char* p, const char* type, RTData data, size_t size, size_t stride, size_t ofs) override {
        ok(yrtSetArray(d, h, p, type, data, size, stride, ofs), "rtSetArray"); }
    void rtSetString(RTHandle h, const char* p, const char* str) override { ok(yrtSetString(d, h, p, str), "rtSetString"); }
    void rtGetString(RTHandle h, const std::string& p, std::string& str) override {
        char buf[4096]; buf[0] = 0; ok(yrtGetString(d, h, p.c_str(), buf, sizeof(buf)), "rtGetString"); str = buf; }
    void rtSetImage(RTHandle h, const char* p, RTImage img) override { ok(yrtSetImage(d, h, p, img), "rtSetImage"); }
    void rtSetTexture(RTHandle h, const char* p, RTTexture tex) override { ok(yrtSetTexture(d, h, p, tex), "rtSetTexture"); }
    void rtSetTransform(RTHandle h, const char* p, const float* t) override { ok(yrtSetTransform(d, h, p, t), "rtSetTransform"); }
    void rtGetTransform(RTHandle h, const char* p, AffineSpace3f* t) override {
        float v[12]; ok(yrtGetTransform(d, h, p, v), "rtGetTransform");
        if (t) *t = AffineSpace3f(Vector3f(v[0], v[1], v[2]), Vector3f(v[3], v[4], v[5]), Vector3f(v[6], v[7], v[8]), Vector3f(v[9], v[10], v[11])); }
    void rtClear(RTHandle h) override { ok(yrtClear(d, h), "rtClear"); }
    void rtCommit(RTHandle h) override { ok(yrtCommit(d, h), "rtCommit"); }

    void rtRenderFrame(RTRenderer renderer, RTCamera camera, RTScene scene, RTToneMapper tonemapper, RTFrameBuffer fb, int accumulate) override {
        ok(yrtRenderFrame(d, renderer, camera, scene, tonemapper, fb, accumulate), "rtRenderFrame"); }
    bool rtPick(RTCamera camera, float x, float y, RTScene scene, float& px, float& py, float& pz) override {
        const int r = yrtPick(d, camera, x, y, scene, &px, &py, &pz);
        if (r < 0) fail("rtPick");
        return r == 1; }
};

extern "C" __attribute__((visibility("default"))) Device* create(const char* parms, size_t numThreads, int threadsPriority, const char* rtcore_cfg) {
    yrt_device* dev = yrtCreateDevice(parms, numThreads, threadsPriority, rtcore_cfg);
    if (!dev) throw std::runtime_error(std::string("device_cuda: ") + yrtGetLastError());
    return new CudaDevice(dev);
}

// Optional hook for callers that know device_cuda's extensions (include/yrt_device.h, yrtx*): the C handle behind a Device created by
// this plugin, or NULL for any other Device. Object handles (RTFrameBuffer, ...) returned by this plugin ARE the C-ABI handles.
extern "C" __attribute__((visibility("default"))) yrt_device* device_cuda_native(Device* device) {
    CudaDevice* c = dynamic_cast<CudaDevice*>(device);
    return c ? c->native() : nullptr;
}

}  // namespace embree
