// plugin_smoke.cpp — a minimal caller on the reference's side of the boundary: it loads a back end exactly the way
// Device::rtCreateDevice does (devices/device/device.cpp:24-35: dlopen + dlsym("create")) and drives it through the
// embree::Device virtual interface only (devices/device/device.h:126-329), like the reference's loaders and front end.
// Usage: plugin_smoke <path/to/libdevice_XXX.so> [size]   -> prints "mean r g b" of a small Cornell-like frame.
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <vector>

#include "device/device.h"

using namespace embree;
typedef Device* (*create_device_func)(const char*, size_t, int, const char*);

int main(int argc, char** argv) try {
    if (argc < 2) { fprintf(stderr, "usage: %s <libdevice_*.so> [size]\n", argv[0]); return 2; }
    const size_t size = argc > 2 ? (size_t)atoi(argv[2]) : 64;
    void* lib = dlopen(argv[1], RTLD_NOW);
    if (!lib) throw std::runtime_error(std::string("failed loading library: ") + dlerror());
    create_device_func f = (create_device_func)dlsym(lib, "create");
    if (!f) throw std::runtime_error("invalid device library");
    Device* dev = f("", 1, 0, "");
    if (!dev) throw std::runtime_error("device creation failed");

    // a white floor quad + a red back wall, a quad light above (two triangle lights), pinhole camera
    const float pos[] = {-1, 0, -1, 1, 0, -1, 1, 0, 1, -1, 0, 1,   -1, 0, 1, 1, 0, 1, 1, 2, 1, -1, 2, 1};
    const int idx[] = {0, 2, 1, 0, 3, 2};
    std::vector<Device::RTPrimitive> prims;
    const float colors[2][3] = {{0.8f, 0.8f, 0.8f}, {0.8f, 0.1f, 0.1f}};
    for (int m = 0; m < 2; m++) {
        // the second mesh hands its positions over as "immutable_managed", allocated the way the reference's loaders do it: an interior
        // pointer from embree::alignedMalloc (xml_loader.cpp:227-267, common/sys/platform.cpp:171-181). The device must release it with the
        // alignedFree convention — plain free() on this pointer corrupts the heap (glibc aborts) — and must have copied it by then.
        Device::RTData dp;
        if (m == 1) {
            const size_t bytes = 48, align = 64;
            char* base = (char*)malloc(bytes + align + sizeof(int));
            char* unaligned = base + sizeof(int);
            char* aligned = unaligned + align - ((size_t)unaligned & (align - 1));
            ((int*)aligned)[-1] = (int)((size_t)aligned - (size_t)base);
            memcpy(aligned, pos + 12, bytes);
            dp = dev->rtNewData("immutable_managed", bytes, aligned);
        } else dp = dev->rtNewData("immutable", 48, pos);
        Device::RTData di = dev->rtNewData("immutable", sizeof(idx), idx);
        Device::RTShape mesh = dev->rtNewShape("trianglemesh");
        dev->rtSetArray(mesh, "positions", "float3", dp, 4, 12, 0);
        dev->rtSetArray(mesh, "indices", "int3", di, 2, 12, 0);
        dev->rtCommit(mesh);
        Device::RTMaterial mat = dev->rtNewMaterial("matte");
        dev->rtSetFloat3(mat, "reflectance", colors[m][0], colors[m][1], colors[m][2]);
        dev->rtCommit(mat);
        prims.push_back(dev->rtNewShapePrimitive(mesh, mat, NULL));
        dev->rtDecRef(dp); dev->rtDecRef(di);
    }
    const float lv[2][9] = {{0.5f, 1.9f, 0.5f, 0.5f, 1.9f, -0.5f, -0.5f, 1.9f, -0.5f}, {0.5f, 1.9f, 0.5f, -0.5f, 1.9f, -0.5f, -0.5f, 1.9f, 0.5f}};
    for (int l = 0; l < 2; l++) {
        Device::RTLight light = dev->rtNewLight("trianglelight");
        dev->rtSetFloat3(light, "v0", lv[l][0], lv[l][1], lv[l][2]); dev->rtSetFloat3(light, "v1", lv[l][3], lv[l][4], lv[l][5]);
        dev->rtSetFloat3(light, "v2", lv[l][6], lv[l][7], lv[l][8]); dev->rtSetFloat3(light, "L", 10, 10, 10);
        dev->rtCommit(light);
        prims.push_back(dev->rtNewLightPrimitive(light, NULL, NULL));
    }
    Device::RTScene scene = dev->rtNewScene("default");
    for (size_t i = 0; i < prims.size(); i++) dev->rtSetPrimitive(scene, i, prims[i]);
    dev->rtCommit(scene);

    const AffineSpace3f space = AffineSpace3f::lookAtPoint(Vector3f(0, 1, -3), Vector3f(0, 0.8f, 0), Vector3f(0, 1, 0));
    const float xfm[12] = {space.l.vx.x, space.l.vx.y, space.l.vx.z, space.l.vy.x, space.l.vy.y, space.l.vy.z,
                           space.l.vz.x, space.l.vz.y, space.l.vz.z, space.p.x, space.p.y, space.p.z};
    Device::RTCamera cam = dev->rtNewCamera("pinhole");
    dev->rtSetTransform(cam, "local2world", xfm); dev->rtSetFloat1(cam, "angle", 60.0f); dev->rtSetFloat1(cam, "aspectRatio", 1.0f);
    dev->rtCommit(cam);
    Device::RTRenderer renderer = dev->rtNewRenderer("pathtracer");
    dev->rtSetInt1(renderer, "maxDepth", 3); dev->rtSetInt1(renderer, "sampler.spp", 4);
    dev->rtCommit(renderer);
    Device::RTToneMapper tm = dev->rtNewToneMapper("default");
    dev->rtCommit(tm);
    Device::RTFrameBuffer fb = dev->rtNewFrameBuffer("RGB_FLOAT32", size, size, 1);
    dev->rtRenderFrame(renderer, cam, scene, tm, fb, 0);
    dev->rtSwapBuffers(fb);
    const float* px = (const float*)dev->rtMapFrameBuffer(fb);
    double sum[3] = {0, 0, 0};
    for (size_t i = 0; i < size * size; i++) for (int c = 0; c < 3; c++) sum[c] += px[3 * i + c];
    dev->rtUnmapFrameBuffer(fb);
    printf("mean %.6f %.6f %.6f\n", sum[0] / (size * size), sum[1] / (size * size), sum[2] / (size * size));
    bool threw = false;
    try { dev->rtNewMaterial("lava"); } catch (const std::runtime_error&) { threw = true; }
    if (!threw) throw std::runtime_error("unknown material type did not throw");
    delete dev;
    return 0;
} catch (const std::exception& e) { fprintf(stderr, "plugin_smoke: %s\n", e.what()); return 1; }
