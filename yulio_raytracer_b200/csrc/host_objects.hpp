// host_objects.hpp — host-side object model of device_cuda: handles, buffered parameters and the
// immutable objects they create. Mirrors the semantics (not the code) of the reference's handle
// system: devices/device_singleray/api/handle.h:24-162, api/parms.h:31-133, api/variant.h,
// api/data.h, api/datastream.h, api/instance.h:29-108 (SURVEY §8b "Handles").
#pragma once
#include <atomic>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "common.cuh"

namespace yrt {

struct DataObj {                       // rtNewData (api/data.h:33-45)
    char* ptr = nullptr; size_t bytes = 0; bool owned = true;
    ~DataObj() { if (ptr) free(ptr); }  // "immutable_managed" takes ownership too
};

struct ImageObj {                      // rtNewImage (common/image/image.h:46-111)
    int width = 0, height = 0; int format = TEX_RGB8;
    const void* pixels = nullptr; std::vector<unsigned char> storage;   // storage empty when aliasing caller memory (copy=false)
    void* devPixels = nullptr;         // uploaded lazily at scene commit
    size_t bytes() const { const size_t bpp = format == TEX_RGB8 ? 3 : format == TEX_RGBA8 ? 4 : format == TEX_RGB_F32 ? 12 : 16; return (size_t)width * height * bpp; }
    ~ImageObj();
};

struct TextureObj { std::shared_ptr<ImageObj> image; bool bilinear = true; bool invert = false; };

struct MaterialObj {
    MaterialRec rec;                   // tex[] filled at scene commit from `textures`
    std::shared_ptr<TextureObj> textures[5];
};

struct ShapeObj {
    int type = MESH_FULL;              // MeshType
    std::vector<V3> position, normal, tangentX, tangentY, motion; std::vector<float2> texcoord; std::vector<int4> triangles;   // motion: per-vertex displacement over the shutter interval ("motions", sphere "dPdt")
    bool cullBackFaces = false;
    V3 v0, v1, v2, triNg;              // MESH_TRIANGLE
    bool triNgValid = false;           // the Parms constructor of Triangle leaves Ng unset (shapes/triangle.h:33-38)
};

struct LightObj {
    int type = LIGHT_AMBIENT;
    Col L; V3 v0, v1, v2; float a = 0, b = 0; Aff3 local2world; std::shared_ptr<ImageObj> image;
};

struct ToneMapperObj { float gamma = 1.f; bool vignetting = false; };

struct RendererObj {
    bool debug = false;
    int maxDepth = 10, rrDepth = 5; float minContribution = .02f, epsilon = 32.f * YRT_ULP;
    float tMaxShadowRay = INFINITY, tMaxShadowJitter = .15f; V3 up = V3(0.f, 1.f, 0.f);
    std::shared_ptr<ImageObj> backplate;
    int spp = 1, sets = 64; int filter = 2;   // FilterKind
    float gamma = 1.f; int showProgress = 0;
    void* stopFlag = nullptr; void* statusCallback = nullptr;
    int iteration = 0;                 // IntegratorRenderer::iteration (integratorrenderer.cpp:67-69)
};

// ---- variants / parameter maps ----------------------------------------------------------------
struct Variant {
    enum Type { EMPTY, BOOL1, BOOL2, BOOL3, BOOL4, INT1, INT2, INT3, INT4, FLOAT1, FLOAT2, FLOAT3, FLOAT4,
                STRING, IMAGE, TEXTURE, TRANSFORM, POINTER };
    Type type = EMPTY;
    bool b[4] = {false, false, false, false}; int i[4] = {0, 0, 0, 0}; float f[12] = {0};
    std::string str; void* ptr = nullptr;
    std::shared_ptr<ImageObj> image; std::shared_ptr<TextureObj> texture;
    // typed array view (rtSetArray): `type` is the element type
    std::shared_ptr<DataObj> data; size_t size = 0, stride = 0, ofs = 0; bool isArray = false;
    const char* elem(size_t k) const { return data->ptr + k * stride + ofs; }
};

struct Parms {
    std::map<std::string, Variant> m;
    const Variant* find(const char* n, Variant::Type t) const { auto it = m.find(n); return (it == m.end() || it->second.type != t || it->second.isArray) ? nullptr : &it->second; }
    bool getBool(const char* n, bool d = false) const { auto v = find(n, Variant::BOOL1); return v ? v->b[0] : d; }
    int getInt(const char* n, int d = 0) const { auto v = find(n, Variant::INT1); return v ? v->i[0] : d; }
    float getFloat(const char* n, float d = 0.f) const { auto v = find(n, Variant::FLOAT1); return v ? v->f[0] : d; }
    void getVec2(const char* n, float& x, float& y, float dx, float dy) const { auto v = find(n, Variant::FLOAT2); x = v ? v->f[0] : dx; y = v ? v->f[1] : dy; }
    V3 getV3(const char* n, V3 d = V3(0.f)) const { auto v = find(n, Variant::FLOAT3); return v ? V3(v->f[0], v->f[1], v->f[2]) : d; }
    std::string getString(const char* n, const std::string& d = "") const { auto v = find(n, Variant::STRING); return v ? v->str : d; }
    void* getPointer(const char* n) const { auto v = find(n, Variant::POINTER); return v ? v->ptr : nullptr; }
    std::shared_ptr<ImageObj> getImage(const char* n) const { auto v = find(n, Variant::IMAGE); return v ? v->image : nullptr; }
    std::shared_ptr<TextureObj> getTexture(const char* n) const { auto v = find(n, Variant::TEXTURE); return v ? v->texture : nullptr; }
    Aff3 getTransform(const char* n) const {
        Aff3 a; a.l = lin3_identity(); a.p = V3(0.f);
        if (auto v = find(n, Variant::TRANSFORM)) { a.l.vx = V3(v->f[0], v->f[1], v->f[2]); a.l.vy = V3(v->f[3], v->f[4], v->f[5]); a.l.vz = V3(v->f[6], v->f[7], v->f[8]); a.p = V3(v->f[9], v->f[10], v->f[11]); }
        return a;
    }
    const Variant* getArray(const char* n) const { auto it = m.find(n); return (it == m.end() || it->second.type == Variant::EMPTY) ? nullptr : &it->second; }
};

// ---- handles -----------------------------------------------------------------------------------
enum HandleKind { HK_CAMERA, HK_DATA, HK_IMAGE, HK_TEXTURE, HK_MATERIAL, HK_SHAPE, HK_LIGHT, HK_PRIMITIVE, HK_SCENE,
                  HK_TONEMAPPER, HK_RENDERER, HK_FRAMEBUFFER };
static const uint32_t HANDLE_MAGIC = 0x59525448u;   // 'YRTH'

struct Handle {
    uint32_t magic = HANDLE_MAGIC;
    std::atomic<int> refs{1};          // created owned by the application (api/handle.h:29-31)
    HandleKind kind;
    std::string type;                  // lower-cased creation type string
    Parms parms; bool modified = true; bool constant = false;
    explicit Handle(HandleKind k) : kind(k) {}
    virtual ~Handle() { magic = 0; }
};

struct CameraHandle : Handle { CameraHandle() : Handle(HK_CAMERA) {} std::shared_ptr<CameraData> inst; };
struct DataHandle : Handle { DataHandle() : Handle(HK_DATA) { constant = true; } std::shared_ptr<DataObj> inst; };
struct ImageHandle : Handle { ImageHandle() : Handle(HK_IMAGE) { constant = true; } std::shared_ptr<ImageObj> inst; };
struct TextureHandle : Handle { TextureHandle() : Handle(HK_TEXTURE) {} std::shared_ptr<TextureObj> inst; };
struct MaterialHandle : Handle { MaterialHandle() : Handle(HK_MATERIAL) {} std::shared_ptr<MaterialObj> inst; };
struct ShapeHandle : Handle { ShapeHandle() : Handle(HK_SHAPE) {} std::shared_ptr<ShapeObj> inst; };
struct LightHandle : Handle { LightHandle() : Handle(HK_LIGHT) {} std::shared_ptr<LightObj> inst; };
struct ToneMapperHandle : Handle { ToneMapperHandle() : Handle(HK_TONEMAPPER) {} std::shared_ptr<ToneMapperObj> inst; };
struct RendererHandle : Handle { RendererHandle() : Handle(HK_RENDERER) {} std::shared_ptr<RendererObj> inst; };

struct PrimHandle : Handle {           // api/instance.h:29-108
    PrimHandle() : Handle(HK_PRIMITIVE) {}
    ShapeHandle* shape = nullptr; LightHandle* light = nullptr; MaterialHandle* material = nullptr;   // ref-counted
    Aff3 transform; int illumMask = -1, shadowMask = -1; bool bstatic = false, faceCamera = false;
    ~PrimHandle() override;
};

}  // namespace yrt
