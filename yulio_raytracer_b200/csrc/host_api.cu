// host_api.cu — the C-ABI of device_cuda (include/yrt_device.h): device life cycle, handles, buffered
// parameters and the construction of the immutable host objects from them.
//
// One entry point per virtual of the reference's embree::Device (devices/device/device.h:126-329); the
// semantics restated here are those of devices/device_singleray/api/singleray_device.cpp:105-728 and the
// object constructors it dispatches to (cameras/*.h, shapes/*.h, lights/*.h, materials/*.h, textures/*.h,
// tonemappers/defaulttonemapper.h, renderers/integratorrenderer.cpp:30-61, integrators/pathtraceintegrator.cpp:21-33).
// There is no CPU fallback: creation fails without a usable CUDA device.
#include <strings.h>

#include <cstdio>
#include <cstdlib>

#include "../../include/yrt_device.h"
#include "gen/core_decls.inc"
#include "gen/core_names.inc"      // the entry points below are the single-GPU implementation: yrtX -> yrtX_core (group_api.cu owns the public names)
#include "device_impl.hpp"
#include "camera.cuh"
#include "host_math.hpp"

using namespace yrt;

namespace yrt {

thread_local std::string g_lastError;

// implemented in host_render.cu
void scene_set_primitive(yrt_device* dev, SceneHandle* sc, size_t slot, PrimHandle* prim, const Aff3* overrideXfm);
void scene_commit(yrt_device* dev, SceneHandle* sc);
void render_frame(yrt_device* dev, RendererHandle* r, CameraHandle* c, SceneHandle* s, ToneMapperHandle* t, FrameBufferHandle* f, int accumulate);
void render_frames(yrt_device* dev, RendererHandle* r, size_t numFaces, CameraHandle* const* c, SceneHandle* s, ToneMapperHandle* t, FrameBufferHandle* const* f, int accumulate);
FrameBufferHandle* framebuffer_create(yrt_device* dev, const char* type, size_t w, size_t h, size_t buffers, void** ptrs);
void* framebuffer_map(yrt_device* dev, FrameBufferHandle* fb, int bufID);
void trace_rays(yrt_device* dev, SceneHandle* sc, size_t n, const float* rays, void* hits, int closest, int onDevice, float* ms);
void primary_rays(yrt_device* dev, RendererHandle* r, CameraHandle* c, FrameBufferHandle* f, float* rays, int* sets);
void sample_table(yrt_device* dev, RendererHandle* r, SceneHandle* s, int iteration, int* sets, int* spp, int* n1, int* n2, float* table);
std::shared_ptr<ImageObj> load_image_file(const char* file);
// implemented in image_codecs.cu
void strip_begin(yrt_device* dev, size_t faceW, size_t faceH);
void strip_set_watermark(yrt_device* dev, const char* pngFile);
void strip_add_face(yrt_device* dev, FrameBufferHandle* fb, int cubeFaceIndex, int watermark);
void strip_read(yrt_device* dev, void* rgb);
void strip_encode_jpeg(yrt_device* dev, int cubeFaceIndex, int quality, const char* file);
void strip_release(yrt_device* dev);
std::shared_ptr<ImageObj> decode_png_file(const char* file, bool flipVertical, bool flipHorizontal);

ImageObj::~ImageObj() { if (devPixels) cudaFree(devPixels); }
PrimHandle::~PrimHandle() {
    auto drop = [](Handle* h) { if (h && h->refs.fetch_sub(1) == 1) delete h; };
    drop(shape); drop(light); drop(material);
}

static std::string lower(const char* s) { std::string r(s ? s : ""); for (auto& c : r) c = (char)tolower((unsigned char)c); return r; }

template <typename T> static T* cast(yrt_handle h, HandleKind kind, const char* name) {
    Handle* b = (Handle*)h;
    if (!b || b->magic != HANDLE_MAGIC || b->kind != kind) throw std::runtime_error(std::string("invalid ") + name + " handle");
    return static_cast<T*>(b);
}
static Handle* anyHandle(yrt_handle h) {
    Handle* b = (Handle*)h;
    if (!b || b->magic != HANDLE_MAGIC) throw std::runtime_error("invalid handle");
    return b;
}

// ------------------------------------------------------------------------------------------------
// object construction from parameter maps
// ------------------------------------------------------------------------------------------------
static std::shared_ptr<CameraData> make_camera(const std::string& type, const Parms& p) {
    auto c = std::make_shared<CameraData>();
    memset(c.get(), 0, sizeof(CameraData));
    const Aff3 l2w = p.getTransform("local2world");
    c->local2world = l2w;
    if (type == "pinhole" || type == "depthoffield") {          // cameras/pinholecamera.h:30-36
        const float angle = p.getFloat("angle", 64.0f), aspect = p.getFloat("aspectRatio", 1.0f);
        const V3 W = xfmVector(l2w, V3(-0.5f * aspect, -0.5f, 0.5f * rcpf(tanf(deg2rad(0.5f * angle)))));
        Aff3 p2w; p2w.l.vx = aspect * l2w.l.vx; p2w.l.vy = l2w.l.vy; p2w.l.vz = W; p2w.p = l2w.p;
        c->type = CAM_PINHOLE; c->p2w[0] = p2w;
        if (type == "depthoffield") {                           // cameras/depthoffieldcamera.h:30-34
            c->type = CAM_DOF;
            c->lensRadius = p.getFloat("lensRadius", 0.0f);
            c->focalDistance = p.getFloat("focalDistance") / length(0.5f * p2w.l.vx + 0.5f * p2w.l.vy + p2w.l.vz);
        }
        return c;
    }
    // cameras/StereoCubeCamera.h:16-66
    c->type = CAM_STEREO;
    const float angle = 90.f, aspect = 1.f;
    c->cubeFaceIndex = p.getInt("cubeFaceIndex", 0);
    c->origin = p.getV3("origin", l2w.p);
    const V3 lookAt = p.getV3("lookAt", V3(0.f, 0.f, -1.f));
    c->up = p.getV3("up", V3(0.f, 1.f, 0.f));
    const V3 right = cross(normalize(c->up), normalize(lookAt - c->origin));
    const float sceneScale = p.getFloat("sceneScale", 1.f);
    const float EYE_SEPARATION = 6.35f * 0.393701f;
    c->eyeSeparation = p.getFloat("eyeSeparation", EYE_SEPARATION) * sceneScale;
    const float zeroParallax = p.getFloat("zeroParallaxDistance", EYE_SEPARATION * 30.f) * sceneScale;
    if (zeroParallax != 0.f) { c->rcpZeroParallax = 1.f / zeroParallax; c->toeIn = p.getBool("toeIn", false) ? 1 : 0; }
    else { c->rcpZeroParallax = 0.f; c->toeIn = 0; }
    c->falloffAngle = rclamp(p.getFloat("stereFalloffAngle", 30.f), 0.f, 90.f);
    const V3 W = xfmVector(l2w, V3(-.5f * aspect, -.5f, .5f * rcpf(tanf(deg2rad(.5f * angle)))));
    Aff3 f; f.l.vx = aspect * l2w.l.vx; f.l.vy = l2w.l.vy; f.l.vz = W; f.p = l2w.p;
    c->p2w[0] = f;
    c->xyzStraight = normalize(.5f * f.l.vx + .5f * f.l.vy + f.l.vz);
    c->p2w[1] = mul(aff3_rotate_about(c->origin, c->up, deg2rad(90.f)), f);
    c->p2w[2] = mul(aff3_rotate_about(c->origin, c->up, deg2rad(180.f)), f);
    c->p2w[3] = mul(aff3_rotate_about(c->origin, c->up, deg2rad(-90.f)), f);
    c->p2w[4] = mul(aff3_rotate_about(c->origin, right, deg2rad(-90.f)), f);
    c->p2w[4] = mul(aff3_rotate_about(c->origin, c->up, deg2rad(180.f)), c->p2w[4]);
    c->p2w[5] = mul(aff3_rotate_about(c->origin, right, deg2rad(90.f)), f);
    c->p2w[5] = mul(aff3_rotate_about(c->origin, c->up, deg2rad(180.f)), c->p2w[5]);
    return c;
}

static void read_v3_array(const Parms& p, const char* name, const char* err, std::vector<V3>& out, bool& present) {
    const Variant* v = p.getArray(name);
    present = v != nullptr;
    if (!v) return;
    if (!v->isArray || !v->data || v->type != Variant::FLOAT3) throw std::runtime_error(err);
    out.resize(v->size);
    for (size_t i = 0; i < v->size; i++) { float f[3]; memcpy(f, v->elem(i), 12); out[i] = V3(f[0], f[1], f[2]); }
}

static std::shared_ptr<ShapeObj> make_shape(const std::string& type, const Parms& p) {
    auto s = std::make_shared<ShapeObj>();
    if (type == "trianglemesh") {                               // shapes/trianglemesh.h:29-41, trianglemesh_full.cpp:21-66
        bool hasP, hasM, hasN, hasTx, hasTy;
        std::vector<V3>& motion = s->motion; std::vector<V3>& tx = s->tangentX; std::vector<V3>& ty = s->tangentY;
        read_v3_array(p, "positions", "wrong position format", s->position, hasP);
        read_v3_array(p, "motions", "wrong motion vector format", motion, hasM);
        read_v3_array(p, "normals", "wrong normal format", s->normal, hasN);
        read_v3_array(p, "tangent_x", "wrong tangent format", tx, hasTx);
        read_v3_array(p, "tangent_y", "wrong tangent format", ty, hasTy);
        bool hasUV = false;
        for (const char* key : {"texcoords", "texcoords0"}) {
            const Variant* v = p.getArray(key);
            if (!v) continue;
            hasUV = true;
            if (!v->isArray || !v->data || v->type != Variant::FLOAT2) throw std::runtime_error("wrong texcoords0 format");
            s->texcoord.resize(v->size);
            for (size_t i = 0; i < v->size; i++) { float f[2]; memcpy(f, v->elem(i), 8); s->texcoord[i] = make_float2(f[0], f[1]); }
        }
        if (const Variant* v = p.getArray("indices")) {
            if (!v->isArray || !v->data || v->type != Variant::INT3) throw std::runtime_error("wrong triangle format");
            s->triangles.resize(v->size);
            for (size_t i = 0; i < v->size; i++) { int t[3]; memcpy(t, v->elem(i), 12); s->triangles[i] = make_int4(t[0], t[1], t[2], 0); }
        }
        s->cullBackFaces = p.getBool("cullBackFaces", false);
        const bool withNormals = hasP && !hasM && hasN && !hasTx && !hasTy && !hasUV;
        s->type = withNormals ? MESH_NORMALS : MESH_FULL;
        if (withNormals && s->normal.size() != s->position.size()) {
            // TriangleMeshWithNormals keeps one vertex array sized by the last of positions/normals (trianglemesh_normals.cpp:24-33)
            const size_t n = s->normal.size(); s->position.resize(n, V3(0.f));
        }
        if (!s->normal.empty() && s->normal.size() < s->position.size()) s->normal.resize(s->position.size(), V3(0.f));
        if (!s->texcoord.empty() && s->texcoord.size() < s->position.size()) s->texcoord.resize(s->position.size(), make_float2(0.f, 0.f));
        if (!motion.empty() && motion.size() < s->position.size()) motion.resize(s->position.size(), V3(0.f));
        if (!tx.empty() && tx.size() < s->position.size()) tx.resize(s->position.size(), V3(0.f));
        if (!ty.empty() && ty.size() < s->position.size()) ty.resize(s->position.size(), V3(0.f));
        return s;
    }
    if (type == "triangle") {                                   // shapes/triangle.h:33-38
        s->type = MESH_TRIANGLE;
        s->v0 = p.getV3("v0"); s->v1 = p.getV3("v1"); s->v2 = p.getV3("v2");
        return s;
    }
    if (type == "sphere") {                                     // shapes/sphere.h:34-81
        s->type = MESH_FULL;
        const V3 P = p.getV3("P"), dPdt = p.getV3("dPdt");
        const float r = p.getFloat("r");
        const size_t numTheta = (size_t)p.getInt("numTheta"), numPhi = (size_t)p.getInt("numPhi");
        auto eval = [](float theta, float phi) { return V3(sinf(theta) * cosf(phi), cosf(theta), sinf(theta) * sinf(phi)); };
        for (size_t theta = 0; theta <= numTheta; theta++) {
            const float rcpNumTheta = rcpf(float(numTheta));
            for (size_t phi = 0; phi < numPhi; phi++) {
                const float rcpNumPhi = rcpf(float(numPhi));
                V3 pt = eval(theta * YRT_PI * rcpNumTheta, phi * 2.0f * YRT_PI * rcpNumPhi);
                const V3 dpdu = eval((theta + 0.001f) * YRT_PI * rcpNumTheta, phi * 2.0f * YRT_PI * rcpNumPhi) - pt;
                const V3 dpdv = eval(theta * YRT_PI * rcpNumTheta, (phi + 0.001f) * 2.0f * YRT_PI * rcpNumPhi) - pt;
                pt = r * pt + P;
                s->position.push_back(pt);
                if (dPdt != V3(0.f)) s->motion.push_back(dPdt);                      // sphere.h:62
                s->normal.push_back(normalize(cross(dpdv, dpdu)));
                s->texcoord.push_back(make_float2(phi * rcpNumPhi, theta * rcpNumTheta));
            }
            if (theta == 0) continue;
            for (size_t phi = 1; phi <= numPhi; phi++) {
                const size_t p00 = (theta - 1) * numPhi + phi - 1, p01 = (theta - 1) * numPhi + phi % numPhi;
                const size_t p10 = theta * numPhi + phi - 1, p11 = theta * numPhi + phi % numPhi;
                if (theta > 1) s->triangles.push_back(make_int4((int)p10, (int)p00, (int)p01, 0));
                if (theta < numTheta) s->triangles.push_back(make_int4((int)p11, (int)p10, (int)p01, 0));
            }
        }
        return s;
    }
    // disk: shapes/disk.h:34-67
    s->type = MESH_FULL;
    const V3 P = p.getV3("P"); const float h = p.getFloat("h"), r = p.getFloat("r");
    const size_t n = (size_t)p.getInt("numTriangles");
    const float rcpN = rcpf(float(n));
    for (size_t phi = 0; phi < n; phi++) {
        const V3 d(sinf(phi * 2.0f * YRT_PI * rcpN), cosf(phi * 2.0f * YRT_PI * rcpN), 0.0f);
        s->position.push_back(P + r * d);
        s->normal.push_back(V3(0.0f, 0.0f, 1.0f));
        s->texcoord.push_back(make_float2(0.0f, 0.0f));
    }
    s->position.push_back(P + V3(0, 0, h));
    // pin P7: the reference gives the apex no normal / texture coordinate and reads element n of arrays of n (disk.h:53-57); the apex gets
    // the rim's values, (0,0,1) and (0,0), here and in the oracle overlay
    s->normal.push_back(V3(0.0f, 0.0f, 1.0f)); s->texcoord.push_back(make_float2(0.f, 0.f));
    for (size_t phi = 0; phi < n; phi++) {
        const size_t p0 = n, p1 = (phi + 0) % n, p2 = (phi + 1) % n;
        switch (phi % 3) {
        case 0: s->triangles.push_back(make_int4((int)p0, (int)p2, (int)p1, 0)); break;
        case 1: s->triangles.push_back(make_int4((int)p1, (int)p0, (int)p2, 0)); break;
        case 2: s->triangles.push_back(make_int4((int)p2, (int)p1, (int)p0, 0)); break;
        }
    }
    return s;
}

static std::shared_ptr<LightObj> make_light(const std::string& type, const Parms& p) {
    auto l = std::make_shared<LightObj>();
    l->L = Col(0.f); l->v0 = l->v1 = l->v2 = V3(0.f); l->local2world = aff3_identity();
    if (type == "ambientlight") { l->type = LIGHT_AMBIENT; l->L = p.getV3("L"); }                         // ambientlight.h:39-41
    else if (type == "pointlight") { l->type = LIGHT_POINT; l->v0 = p.getV3("P"); l->L = p.getV3("I"); }   // pointlight.h:40-43
    else if (type == "spotlight") {                                                                        // spotlight.h:41-47
        l->type = LIGHT_SPOT; l->v0 = p.getV3("P"); l->v1 = -normalize(p.getV3("D")); l->L = p.getV3("I");
        l->a = cosf(0.5f * deg2rad(p.getFloat("angleMin"))); l->b = cosf(0.5f * deg2rad(p.getFloat("angleMax")));
    } else if (type == "directionallight") { l->type = LIGHT_DIRECTIONAL; l->v0 = -normalize(p.getV3("D")); l->L = p.getV3("E"); }  // directionallight.h:39-42
    else if (type == "distantlight") {                                                                     // distantlight.h:46-51
        l->type = LIGHT_DISTANT; l->v0 = -normalize(p.getV3("D")); l->L = p.getV3("L");
        l->a = deg2rad(p.getFloat("halfAngle")); l->b = cosf(l->a);
    } else if (type == "hdrilight") {                                                                      // hdrilight.cpp:39-56
        l->type = LIGHT_HDRI; l->local2world = p.getTransform("local2world"); l->L = p.getV3("L", V3(1.f));
        l->image = p.getImage("image");
        if (!l->image) {
            auto img = std::make_shared<ImageObj>(); img->width = 5; img->height = 5; img->format = TEX_RGB_F32;
            img->storage.resize(5 * 5 * 12); float one = 1.f;
            for (int i = 0; i < 75; i++) memcpy(&img->storage[4 * i], &one, 4);
            img->pixels = img->storage.data(); l->image = img;
        }
    } else {                                                                                               // trianglelight.h:46-55
        l->type = LIGHT_TRIANGLE; l->v0 = p.getV3("v0"); l->v1 = p.getV3("v1"); l->v2 = p.getV3("v2"); l->L = p.getV3("L");
    }
    return l;
}

static std::shared_ptr<MaterialObj> make_material(const std::string& type, const Parms& p) {
    auto m = std::make_shared<MaterialObj>();
    MaterialRec& r = m->rec; memset(&r, 0, sizeof(r));
    for (int i = 0; i < 5; i++) r.tex[i] = -1;
    r.tOutside = r.tInside = Col(1.f); r.etaOutside = r.etaInside = 1.f;     // Medium::Vacuum (materials/material.h:35-38)
    auto st = [&]() { p.getVec2("s0", r.s0x, r.s0y, 0.f, 0.f); p.getVec2("ds", r.dsx, r.dsy, 1.f, 1.f); };
    if (type == "matte") { r.type = MAT_MATTE; r.c0 = p.getV3("reflectance", V3(1.f)); }                  // matte.h:31-33
    else if (type == "mirror") { r.type = MAT_MIRROR; r.c0 = p.getV3("reflectance", V3(1.f)); }            // mirror.h:32-34
    else if (type == "mattetextured") { r.type = MAT_MATTE_TEXTURED; m->textures[0] = p.getTexture("Kd"); st(); }   // matte_textured.h:32-37
    else if (type == "uber") {                                                                             // Uber.h:20-30
        r.type = MAT_UBER; m->textures[0] = p.getTexture("Kd"); r.c0 = p.getV3("diffuse", V3(0.f)); st();
        r.f[0] = p.getFloat("eta", 1.4f); r.f[1] = p.getFloat("roughness", .9f); r.f[2] = p.getFloat("reflectivity", .0f);
        r.f[3] = rcpf(r.f[1]); r.f[4] = 1.f * rcpf(r.f[0]);
    } else if (type == "dielectric" || type == "glass") {                                                  // dielectric.h:31-45
        r.type = MAT_DIELECTRIC; r.etaOutside = p.getFloat("etaOutside", 1.0f); r.etaInside = p.getFloat("etaInside", 1.4f);
        r.tOutside = p.getV3("transmissionOutside", V3(1.f)); r.tInside = p.getV3("transmission", V3(1.f)); r.isMediaInterface = 1;
    } else if (type == "thindielectric" || type == "thinglass") {                                          // thindielectric.h:33-41
        r.type = MAT_THIN_DIELECTRIC; m->textures[0] = p.getTexture("Kd"); st();
        r.c0 = p.getV3("transmission", V3(1.f)); r.f[0] = p.getFloat("eta", 1.4f); r.f[1] = p.getFloat("thickness", .1f);
        r.f[2] = p.getFloat("transparency", 1.f);
    } else if (type == "obj") {                                                                            // obj.h:32-48
        r.type = MAT_OBJ;
        m->textures[0] = p.getTexture("map_d"); r.f[0] = p.getFloat("d", 1.0f);
        m->textures[1] = p.getTexture("map_Kd"); r.c0 = p.getV3("Kd", V3(1.f));
        m->textures[2] = p.getTexture("map_Ks"); r.c1 = p.getV3("Ks", V3(0.f));
        m->textures[3] = p.getTexture("map_Ns"); r.f[1] = p.getFloat("Ns", 10.0f);
        m->textures[4] = p.getTexture("map_Bump");
    } else if (type == "plastic") {                                                                        // plastic.h:31-36
        r.type = MAT_PLASTIC; r.c0 = p.getV3("pigmentColor", V3(1.f)); r.f[0] = p.getFloat("eta", 1.4f); r.f[1] = p.getFloat("roughness", 0.01f); r.f[3] = rcpf(r.f[1]);
    } else if (type == "metal") {                                                                          // metal.h:35-41
        r.type = MAT_METAL; r.c0 = p.getV3("reflectance", V3(1.f)); r.c1 = p.getV3("eta", V3(1.4f)); r.c2 = p.getV3("k", V3(0.f));
        r.f[1] = p.getFloat("roughness", 0.01f); r.f[3] = rcpf(r.f[1]);
    } else if (type == "brushedmetal") {                                                                   // brushedmetal.h:36-44
        r.type = MAT_BRUSHED_METAL; r.c0 = p.getV3("reflectance", V3(1.f)); r.c1 = p.getV3("eta", V3(1.4f)); r.c2 = p.getV3("k", V3(0.f));
        r.f[1] = p.getFloat("roughnessX", 0.01f); r.f[2] = p.getFloat("roughnessY", 0.01f); r.f[3] = rcpf(r.f[1]); r.f[4] = rcpf(r.f[2]);
    } else if (type == "metallicpaint") {                                                                  // metallicpaint.h:35-44
        r.type = MAT_METALLIC_PAINT; r.c0 = p.getV3("shadeColor", V3(1.f)); r.c1 = p.getV3("glitterColor", V3(0.f));
        r.f[1] = p.getFloat("glitterSpread", 1.0f); r.f[0] = p.getFloat("eta", 1.4f);
    } else if (type == "velvet") {                                                                         // velvet.h:31-36
        r.type = MAT_VELVET; r.c0 = p.getV3("reflectance", V3(1.f)); r.f[0] = p.getFloat("backScattering", 0.f);
        r.c1 = p.getV3("horizonScatteringColor", V3(1.f)); r.f[1] = p.getFloat("horizonScatteringFallOff", 0.f);
    } else throw std::runtime_error("unknown material type: " + type);                                    // singleray_device.cpp:280
    return m;
}

static std::shared_ptr<RendererObj> make_renderer(const std::string& type, const Parms& p) {
    auto r = std::make_shared<RendererObj>();
    if (type == "debug") {                                                 // renderers/debugrenderer.cpp:22-26
        r->debug = true; r->maxDepth = p.getInt("maxDepth", 1); r->spp = p.getInt("sampler.spp", 1);
        return r;
    }
    const std::string integ = p.getString("integrator", "pathtracer");    // integratorrenderer.cpp:33-36
    if (integ != "pathtracer") throw std::runtime_error("unknown integrator type: " + integ);
    r->maxDepth = p.getInt("maxDepth", 10); r->rrDepth = p.getInt("rrDepth", 5);   // pathtraceintegrator.cpp:24-32
    r->minContribution = p.getFloat("minContribution", .02f);
    r->epsilon = p.getFloat("epsilon", 32.f) * YRT_ULP;
    r->tMaxShadowRay = p.getFloat("tMaxShadowRay", INFINITY);
    r->tMaxShadowJitter = p.getFloat("tMaxShadowJitter", .15f);
    r->up = p.getV3("up", V3(0.f, 1.f, 0.f));
    r->backplate = p.getImage("backplate");
    const std::string smp = p.getString("sampler", "multijittered");      // integratorrenderer.cpp:38-41, sampler.cpp:27-30
    if (smp != "multijittered") throw std::runtime_error("unknown sampler type: " + smp);
    r->spp = p.getInt("sampler.spp", 1); r->sets = p.getInt("sampler.sets", 64);
    const std::string flt = p.getString("filter", "bspline");             // integratorrenderer.cpp:43-48
    if (flt == "none") r->filter = FILTER_NONE; else if (flt == "box") r->filter = FILTER_BOX; else if (flt == "bspline") r->filter = FILTER_BSPLINE;
    else throw std::runtime_error("unknown filter type: " + flt);
    r->gamma = p.getFloat("gamma", 1.0f); r->showProgress = p.getInt("showprogress", 0);
    r->stopFlag = p.getPointer("stopFlag"); r->statusCallback = p.getPointer("statusCallback");
    if (r->maxDepth < 0) r->maxDepth = 0;
    if (r->spp < 1) r->spp = 1;
    if (r->sets < 1 || r->sets > 255) throw std::runtime_error("device_cuda: sampler.sets must be in [1,255]");
    return r;
}

static void commit_handle(yrt_device* dev, Handle* h) {
    if (h->constant) throw std::runtime_error("cannot modify constant handle");   // api/handle.h:58
    switch (h->kind) {
    case HK_PRIMITIVE: return;                                                    // api/instance.h:74
    case HK_SCENE: scene_commit(dev, static_cast<SceneHandle*>(h)); return;
    default: break;
    }
    if (!h->modified) return;                                                     // api/handle.h:99-103
    switch (h->kind) {
    case HK_CAMERA: static_cast<CameraHandle*>(h)->inst = make_camera(h->type, h->parms); break;
    case HK_TEXTURE: {                                                            // textures/Bilinear.h:18-21, nearestneighbor.h
        auto t = std::make_shared<TextureObj>();
        t->image = h->parms.getImage("image"); t->invert = h->parms.getBool("invert", false); t->bilinear = h->type == "bilinear";
        if (!t->image) {   // both texture constructors fall back to a 1x1 white image... keep a defined value
            auto img = std::make_shared<ImageObj>(); img->width = 1; img->height = 1; img->format = TEX_RGB8;
            img->storage.assign(3, 255); img->pixels = img->storage.data(); t->image = img;
        }
        static_cast<TextureHandle*>(h)->inst = t; break;
    }
    case HK_MATERIAL: static_cast<MaterialHandle*>(h)->inst = make_material(h->type, h->parms); break;
    case HK_SHAPE: static_cast<ShapeHandle*>(h)->inst = make_shape(h->type, h->parms); break;
    case HK_LIGHT: static_cast<LightHandle*>(h)->inst = make_light(h->type, h->parms); break;
    case HK_TONEMAPPER: {
        auto t = std::make_shared<ToneMapperObj>(); t->gamma = h->parms.getFloat("gamma", 1.0f); t->vignetting = h->parms.getBool("vignetting", false);
        static_cast<ToneMapperHandle*>(h)->inst = t; break;
    }
    case HK_RENDERER: static_cast<RendererHandle*>(h)->inst = make_renderer(h->type, h->parms); break;
    default: break;
    }
    h->modified = false;
}

static void set_variant(yrt_device* dev, yrt_handle handle, const char* property, Variant&& v) {
    if (!property) throw std::runtime_error("invalid property");                  // singleray_device.cpp:476
    if (!handle) {
        if (v.type == Variant::INT1 && !v.isArray) {                               // singleray_device.cpp:505-508
            if (!strcmp(property, "serverID")) dev->serverID = v.i[0];
            else if (!strcmp(property, "serverCount")) dev->serverCount = v.i[0] < 1 ? 1 : v.i[0];
        }
        return;
    }
    Handle* h = anyHandle(handle);
    if (h->constant) throw std::runtime_error("cannot modify constant handle");
    if (h->kind == HK_PRIMITIVE) {                                                 // api/instance.h:55-61
        PrimHandle* p = static_cast<PrimHandle*>(h);
        const std::string k(property);
        if (k == "illumMask") p->illumMask = v.i[0]; else if (k == "shadowMask") p->shadowMask = v.i[0];
        else if (k == "static") p->bstatic = v.b[0]; else if (k == "faceCamera") p->faceCamera = v.b[0];
        return;
    }
    if (h->kind == HK_SCENE) {                                                     // api/scene.h:42-47
        SceneHandle* s = static_cast<SceneHandle*>(h);
        const std::string k(property);
        if (k == "accel") s->accelTy = v.str; else if (k == "builder") s->builderTy = v.str; else if (k == "traverser") s->traverserTy = v.str;
        return;
    }
    h->parms.m[property] = std::move(v);
    h->modified = true;
}

static Variant get_variant(yrt_handle handle, const char* property) {
    Handle* h = anyHandle(handle);
    if (h->constant) throw std::runtime_error("nothing to get");                  // api/handle.h:63-65
    const std::string k(property ? property : "");
    Variant v;
    if (h->kind == HK_PRIMITIVE) {
        PrimHandle* p = static_cast<PrimHandle*>(h);
        if (k == "illumMask") { v.type = Variant::INT1; v.i[0] = p->illumMask; } else if (k == "shadowMask") { v.type = Variant::INT1; v.i[0] = p->shadowMask; }
        else if (k == "static") { v.type = Variant::BOOL1; v.b[0] = p->bstatic; } else if (k == "faceCamera") { v.type = Variant::BOOL1; v.b[0] = p->faceCamera; }
        return v;
    }
    if (h->kind == HK_SCENE) {
        SceneHandle* s = static_cast<SceneHandle*>(h);
        v.type = Variant::STRING;
        if (k == "accel") v.str = s->accelTy; else if (k == "builder") v.str = s->builderTy; else if (k == "traverser") v.str = s->traverserTy; else v.type = Variant::EMPTY;
        return v;
    }
    auto it = h->parms.m.find(k);
    return it == h->parms.m.end() ? v : it->second;
}

}  // namespace yrt

void yrt_device::bind() const { YRT_CK(cudaSetDevice(gpu)); }

// ------------------------------------------------------------------------------------------------
// exported C symbols
// ------------------------------------------------------------------------------------------------
#define GUARD_H(...) try { if (!dev) throw std::runtime_error("invalid device"); std::lock_guard<std::mutex> lock(dev->mutex); dev->bind(); __VA_ARGS__ } \
    catch (const std::exception& e) { g_lastError = e.what(); return nullptr; }
#define GUARD_S(...) try { if (!dev) throw std::runtime_error("invalid device"); std::lock_guard<std::mutex> lock(dev->mutex); dev->bind(); __VA_ARGS__; return YRT_OK; } \
    catch (const std::exception& e) { g_lastError = e.what(); return YRT_ERROR; }

static long cfg_int(const std::string& cfg, const char* key, long def) {
    size_t pos = 0; const std::string k = std::string(key) + "=";
    while (pos < cfg.size()) {
        size_t end = cfg.find(',', pos); if (end == std::string::npos) end = cfg.size();
        std::string item = cfg.substr(pos, end - pos);
        while (!item.empty() && item[0] == ' ') item.erase(0, 1);
        if (item.compare(0, k.size(), k) == 0) return strtol(item.c_str() + k.size(), nullptr, 0);
        pos = end + 1;
    }
    return def;
}

extern "C" {

yrt_device* yrtCreateDevice(const char* parms, size_t numThreads, int threadsPriority, const char* cfg_) {
    (void)parms; (void)numThreads; (void)threadsPriority;
    try {
        const std::string cfg(cfg_ ? cfg_ : "");
        int count = 0;
        cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count == 0)
            throw std::runtime_error(std::string("device_cuda needs a CUDA device and has no CPU fallback: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "no device found"));
        const int gpu = (int)cfg_int(cfg, "gpu", 0);
        if (gpu < 0 || gpu >= count) throw std::runtime_error("device_cuda: gpu ordinal out of range");
        cudaDeviceProp prop; YRT_CK(cudaGetDeviceProperties(&prop, gpu));
        if (prop.major < 10) throw std::runtime_error(std::string("device_cuda is built for sm_100a only; found ") + prop.name);
        auto* dev = new yrt_device();
        dev->gpu = gpu; dev->numSMs = prop.multiProcessorCount; dev->l2Bytes = (size_t)prop.l2CacheSize;
        YRT_CK(cudaSetDevice(gpu));
        YRT_CK(cudaStreamCreateWithFlags(&dev->stream, cudaStreamNonBlocking));
        YRT_CK(cudaStreamCreateWithFlags(&dev->stream1, cudaStreamNonBlocking));
        dev->lanes = (int)cfg_int(cfg, "lanes", 1) >= 2 ? 2 : 1;
        {   // keep freed scratch in the stream-ordered pool instead of returning it to the OS at every synchronisation
            cudaMemPool_t pool; YRT_CK(cudaDeviceGetDefaultMemPool(&pool, gpu));
            uint64_t keep = ~0ull; YRT_CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        }
        dev->chunkPaths = (uint32_t)cfg_int(cfg, "chunk", 1l << 26);
        if (dev->chunkPaths < 1024) dev->chunkPaths = 1024;
        dev->countStats = (int)cfg_int(cfg, "stats", 0);
        dev->verbose = (int)cfg_int(cfg, "verbose", 0);
        dev->alwaysRebuild = (int)cfg_int(cfg, "rebuild", 0);
        dev->useTimers = (int)cfg_int(cfg, "timers", 1);
        dev->serverID = (int)cfg_int(cfg, "serverID", 0);
        dev->serverCount = (int)cfg_int(cfg, "serverCount", 1);
        dev->tuneRefillMin = (int)cfg_int(cfg, "refill", dev->tuneRefillMin);
        dev->tuneTriNum = (int)cfg_int(cfg, "trinum", dev->tuneTriNum); dev->tuneUserTriNum = (int)cfg_int(cfg, "trinum", dev->tuneUserTriNum); dev->tuneTriDen = (int)cfg_int(cfg, "triden", dev->tuneTriDen); dev->tuneSimple = cfg_int(cfg, "trav", 1) == 0;
        dev->shadeCtas = (int)cfg_int(cfg, "shadectas", YRT_SHADE_MINBLOCKS); dev->traceCtas = (int)cfg_int(cfg, "tracectas", 8);
        dev->syncMinPaths = (uint32_t)cfg_int(cfg, "syncmin", dev->syncMinPaths);
        dev->bvhPloc = (int)cfg_int(cfg, "bvh", 1); dev->plocRadius = (int)cfg_int(cfg, "plocr", dev->plocRadius); dev->splitLeaves = (int)cfg_int(cfg, "splitleaves", 1);
        dev->tunePrefetch = (int)cfg_int(cfg, "prefetch", 0);
        dev->bvhCollapseDp = (int)cfg_int(cfg, "collapse", 1); dev->bvhCTri = 0.01f * (float)cfg_int(cfg, "ctri", 60);
        dev->sortRays = (int)cfg_int(cfg, "sort", 0); dev->sortMin = (uint32_t)cfg_int(cfg, "sortmin", 1l << 16);
        YRT_CK(cudaHostAlloc((void**)&dev->hostCounters, 16 * sizeof(uint32_t), cudaHostAllocDefault));
        if (dev->tuneRefillMin < 1) dev->tuneRefillMin = 1; if (dev->tuneRefillMin > 32) dev->tuneRefillMin = 32;
        dev->stats.num_gpus = 1;
        return dev;
    } catch (const std::exception& e) { g_lastError = e.what(); return nullptr; }
}

void yrtDestroyDevice(yrt_device* dev) {
    if (!dev) return;
    cudaSetDevice(dev->gpu);
    cudaStreamSynchronize(dev->stream); cudaStreamSynchronize(dev->stream1);
    dev->wf.release(); dev->wf1.release(); dev->timers.release(); dev->sampleTable.release(); strip_release(dev);
    if (dev->hostCounters) cudaFreeHost(dev->hostCounters);
    cudaStreamDestroy(dev->stream); cudaStreamDestroy(dev->stream1);
    delete dev;
}

const char* yrtGetLastError(void) { return g_lastError.c_str(); }

// ---- creation --------------------------------------------------------------------------------------
#define NEW_SIMPLE(fn, HT, what, ...)                                                                         \
    yrt_handle fn(yrt_device* dev, const char* type) {                                                        \
        GUARD_H(const std::string t = lower(type); static const char* ok[] = {__VA_ARGS__, nullptr};          \
                bool found = false; for (const char** k = ok; *k; k++) found |= t == *k;                     \
                if (!found) throw std::runtime_error(std::string(what) + std::string(type ? type : ""));     \
                auto* h = new HT(); h->type = t; return h;)                                                   \
    }
NEW_SIMPLE(yrtNewCamera, CameraHandle, "unknown camera type: ", "pinhole", "depthoffield", "stereo")
NEW_SIMPLE(yrtNewMaterial, MaterialHandle, "unknown material type: ", "matte", "plastic", "dielectric", "glass", "thindielectric", "thinglass",
           "mirror", "metal", "brushedmetal", "metallicpaint", "mattetextured", "uber", "obj", "velvet")
NEW_SIMPLE(yrtNewShape, ShapeHandle, "unknown shape type: ", "trianglemesh", "triangle", "sphere", "disk")
NEW_SIMPLE(yrtNewLight, LightHandle, "unknown light type: ", "ambientlight", "pointlight", "spotlight", "directionallight", "distantlight",
           "hdrilight", "trianglelight")
NEW_SIMPLE(yrtNewToneMapper, ToneMapperHandle, "unknown tonemapper type: ", "default")
NEW_SIMPLE(yrtNewRenderer, RendererHandle, "unknown renderer type: ", "debug", "pathtracer")

yrt_handle yrtNewTexture(yrt_device* dev, const char* type) {
    GUARD_H(std::string t = lower(type);
            if (t == "image") t = "nearest";
            if (t != "bilinear" && t != "nearest") throw std::runtime_error("unsupported texture type: " + std::string(type ? type : ""));
            auto* h = new TextureHandle(); h->type = t; return h;)
}

yrt_handle yrtNewScene(yrt_device* dev, const char* type) {
    GUARD_H(if (!type || (strcmp(type, "default") && strcmp(type, "flat"))) throw std::runtime_error("unknown scene type: " + std::string(type ? type : ""));
            auto* h = new SceneHandle(); h->type = type; return h;)
}

yrt_handle yrtNewData(yrt_device* dev, const char* type, size_t bytes, const void* data) {
    GUARD_H(const std::string t = lower(type);
            auto d = std::make_shared<DataObj>(); d->bytes = bytes;
            if (t == "immutable") {                                    // api/data.h:33-38 (copy)
                d->ptr = (char*)malloc(bytes ? bytes : 1);
                if (!d->ptr) throw std::runtime_error("memory allocation failed");
                if (bytes && data) memcpy(d->ptr, data, bytes);
            } else if (t == "immutable_managed") d->ptr = (char*)data;  // takes ownership (freed with the handle)
            else throw std::runtime_error("unknown data buffer type: " + std::string(type ? type : ""));
            auto* h = new DataHandle(); h->type = t; h->inst = d; return h;)
}

yrt_handle yrtNewDataFromFile(yrt_device* dev, const char* type, const char* file, size_t offset, size_t bytes) {
    GUARD_H(if (file && !strncmp(file, "server:", 7)) file += 7;
            if (lower(type) != "immutable") throw std::runtime_error("unknown data buffer type: " + std::string(type ? type : ""));
            FILE* f = fopen(file ? file : "", "rb");
            if (!f) throw std::runtime_error("cannot open file " + std::string(file ? file : ""));
            auto d = std::make_shared<DataObj>(); d->bytes = bytes; d->ptr = (char*)malloc(bytes ? bytes : 1);
            if (!d->ptr) { fclose(f); throw std::runtime_error("memory allocation failed"); }
            fseek(f, (long)offset, SEEK_SET);
            const size_t got = fread(d->ptr, 1, bytes, f); fclose(f);
            if (got != bytes) throw std::runtime_error("error filling data buffer from file");
            auto* h = new DataHandle(); h->type = "immutable"; h->inst = d; return h;)
}

yrt_handle yrtNewImage(yrt_device* dev, const char* type, size_t width, size_t height, const void* data, int copy) {
    GUARD_H(const std::string t = lower(type);
            auto img = std::make_shared<ImageObj>(); img->width = (int)width; img->height = (int)height;
            if (t == "rgb8") img->format = TEX_RGB8; else if (t == "rgba8") img->format = TEX_RGBA8;
            else if (t == "rgb_float32") img->format = TEX_RGB_F32; else if (t == "rgba_float32") img->format = TEX_RGBA_F32;
            else throw std::runtime_error("unknown image type: " + std::string(type ? type : ""));
            if (copy || !data) { img->storage.assign(img->bytes(), 0); if (data) memcpy(img->storage.data(), data, img->bytes()); img->pixels = img->storage.data(); }
            else img->pixels = data;                                   // aliases caller memory (common/image/image.h:64-74)
            auto* h = new ImageHandle(); h->type = t; h->inst = img; return h;)
}

yrt_handle yrtNewImageFromFile(yrt_device* dev, const char* file) {
    GUARD_H(if (file && !strncmp(file, "server:", 7)) file += 7;
            auto img = load_image_file(file);
            if (!img) {                                                // 1x1 white fallback (singleray_device.cpp:250)
                img = std::make_shared<ImageObj>(); img->width = img->height = 1; img->format = TEX_RGB8;
                img->storage.assign(3, 255); img->pixels = img->storage.data();
            }
            auto* h = new ImageHandle(); h->type = "file"; h->inst = img; return h;)
}

yrt_handle yrtNewShapePrimitive(yrt_device* dev, yrt_handle shape, yrt_handle material, const float* transform, int faceCamera) {
    GUARD_H(auto* s = cast<ShapeHandle>(shape, HK_SHAPE, "shape"); auto* m = cast<MaterialHandle>(material, HK_MATERIAL, "material");
            auto* p = new PrimHandle(); p->type = "shape"; p->shape = s; p->material = m; s->refs++; m->refs++;
            p->transform = aff3_from_array(transform); p->faceCamera = faceCamera != 0; return p;)
}

yrt_handle yrtNewLightPrimitive(yrt_device* dev, yrt_handle light, yrt_handle material, const float* transform) {
    GUARD_H(auto* l = cast<LightHandle>(light, HK_LIGHT, "light");
            MaterialHandle* m = material ? cast<MaterialHandle>(material, HK_MATERIAL, "material") : nullptr;
            auto* p = new PrimHandle(); p->type = "light"; p->light = l; p->material = m; l->refs++; if (m) m->refs++;
            p->transform = aff3_from_array(transform); return p;)
}

yrt_handle yrtTransformPrimitive(yrt_device* dev, yrt_handle prim, const float* transform) {
    GUARD_H(auto* o = cast<PrimHandle>(prim, HK_PRIMITIVE, "primitive");
            auto* p = new PrimHandle(); p->type = o->type;
            p->shape = o->shape; p->light = o->light; p->material = o->material;
            if (p->shape) p->shape->refs++; if (p->light) p->light->refs++; if (p->material) p->material->refs++;
            p->transform = mul(aff3_from_array(transform), o->transform);      // api/instance.h:46-50
            p->illumMask = o->illumMask; p->shadowMask = o->shadowMask; p->bstatic = o->bstatic; p->faceCamera = o->faceCamera;
            return p;)
}

yrt_status yrtSetPrimitive(yrt_device* dev, yrt_handle scene, size_t slot, yrt_handle prim) {
    GUARD_S(auto* s = cast<SceneHandle>(scene, HK_SCENE, "scene");
            PrimHandle* p = prim ? cast<PrimHandle>(prim, HK_PRIMITIVE, "primitive") : nullptr;
            scene_set_primitive(dev, s, slot, p, nullptr))
}

yrt_status yrtUpdatePrimitive(yrt_device* dev, yrt_handle scene, size_t slot, yrt_handle prim, const float camPos[3], const float camUp[3]) {
    GUARD_S(auto* s = cast<SceneHandle>(scene, HK_SCENE, "scene");
            if (!prim) { scene_set_primitive(dev, s, slot, nullptr, nullptr); return YRT_OK; }
            auto* p = cast<PrimHandle>(prim, HK_PRIMITIVE, "primitive");
            if (!p->faceCamera) return YRT_OK;                                 // singleray_device.cpp:360
            // billboard transform (singleray_device.cpp:363-396)
            const V3 cp(camPos[0], camPos[1], camPos[2]), cu(camUp[0], camUp[1], camUp[2]);
            const V3 primPos = p->transform.p;
            V3 toEye = cp - primPos; toEye.y = 0.f; toEye = normalize(toEye);
            const Aff3 lookAt = aff3_look_at_point(V3(0.f), toEye, cu);
            V3 right = cross(cu, V3(0.f, 0.f, 1.f));
            if (right == V3(0.f)) right = cross(cu, V3(0.f, 1.f, 0.f));
            if (right == V3(0.f)) right = cross(cu, V3(1.f, 0.f, 0.f));
            const Aff3 vertical = aff3_rotate_about(V3(0.f), right, deg2rad(-90.f));
            Aff3 m = mul(mul(aff3_translate(primPos), lookAt), vertical);
            {   // glm::decompose scale of a TRS matrix: column lengths, negated when the basis is left-handed
                const Lin3& l = p->transform.l;
                V3 sc(length(l.vx), length(l.vy), length(l.vz));
                if (sc.x != 0.f && sc.y != 0.f && sc.z != 0.f) {
                    if (dot(l.vx, cross(l.vy, l.vz)) < 0.f) sc = -sc;
                    m = mul(m, aff3_scale(sc));
                }
            }
            scene_set_primitive(dev, s, slot, p, &m))
}

yrt_handle yrtNewFrameBuffer(yrt_device* dev, const char* type, size_t width, size_t height, size_t buffers, void** ptrs) {
    GUARD_H(return framebuffer_create(dev, type, width, height, buffers, ptrs);)
}
void* yrtMapFrameBuffer(yrt_device* dev, yrt_handle fb, int bufID) {
    GUARD_H(return framebuffer_map(dev, cast<FrameBufferHandle>(fb, HK_FRAMEBUFFER, "framebuffer"), bufID);)
}
yrt_status yrtUnmapFrameBuffer(yrt_device* dev, yrt_handle fb, int bufID) { (void)bufID; GUARD_S(cast<FrameBufferHandle>(fb, HK_FRAMEBUFFER, "framebuffer")) }
yrt_status yrtSwapBuffers(yrt_device* dev, yrt_handle fb) {
    GUARD_S(auto* f = cast<FrameBufferHandle>(fb, HK_FRAMEBUFFER, "framebuffer"); f->cur = (f->cur + 1) % f->depth)
}
yrt_status yrtIncRef(yrt_device* dev, yrt_handle h) { (void)dev; try { anyHandle(h)->refs++; return YRT_OK; } catch (const std::exception& e) { g_lastError = e.what(); return YRT_ERROR; } }
yrt_status yrtDecRef(yrt_device* dev, yrt_handle h) {
    GUARD_S(Handle* b = anyHandle(h); if (b->refs.fetch_sub(1) == 1) delete b)
}

// ---- parameters ------------------------------------------------------------------------------------
static Variant vb(int n, int x, int y = 0, int z = 0, int w = 0) { Variant v; v.type = (Variant::Type)(Variant::BOOL1 + n - 1); v.b[0] = x != 0; v.b[1] = y != 0; v.b[2] = z != 0; v.b[3] = w != 0; return v; }
static Variant vi(int n, int x, int y = 0, int z = 0, int w = 0) { Variant v; v.type = (Variant::Type)(Variant::INT1 + n - 1); v.i[0] = x; v.i[1] = y; v.i[2] = z; v.i[3] = w; return v; }
static Variant vf(int n, float x, float y = 0, float z = 0, float w = 0) { Variant v; v.type = (Variant::Type)(Variant::FLOAT1 + n - 1); v.f[0] = x; v.f[1] = y; v.f[2] = z; v.f[3] = w; return v; }

yrt_status yrtSetBool1(yrt_device* dev, yrt_handle h, const char* p, int x) { GUARD_S(set_variant(dev, h, p, vb(1, x))) }
yrt_status yrtSetBool2(yrt_device* dev, yrt_handle h, const char* p, int x, int y) { GUARD_S(set_variant(dev, h, p, vb(2, x, y))) }
yrt_status yrtSetBool3(yrt_device* dev, yrt_handle h, const char* p, int x, int y, int z) { GUARD_S(set_variant(dev, h, p, vb(3, x, y, z))) }
yrt_status yrtSetBool4(yrt_device* dev, yrt_handle h, const char* p, int x, int y, int z, int w) { GUARD_S(set_variant(dev, h, p, vb(4, x, y, z, w))) }
yrt_status yrtSetInt1(yrt_device* dev, yrt_handle h, const char* p, int x) { GUARD_S(set_variant(dev, h, p, vi(1, x))) }
yrt_status yrtSetInt2(yrt_device* dev, yrt_handle h, const char* p, int x, int y) { GUARD_S(set_variant(dev, h, p, vi(2, x, y))) }
yrt_status yrtSetInt3(yrt_device* dev, yrt_handle h, const char* p, int x, int y, int z) { GUARD_S(set_variant(dev, h, p, vi(3, x, y, z))) }
yrt_status yrtSetInt4(yrt_device* dev, yrt_handle h, const char* p, int x, int y, int z, int w) { GUARD_S(set_variant(dev, h, p, vi(4, x, y, z, w))) }
yrt_status yrtSetPointer(yrt_device* dev, yrt_handle h, const char* p, void* ptr) { GUARD_S(Variant v; v.type = Variant::POINTER; v.ptr = ptr; set_variant(dev, h, p, std::move(v))) }
yrt_status yrtSetFloat1(yrt_device* dev, yrt_handle h, const char* p, float x) { GUARD_S(set_variant(dev, h, p, vf(1, x))) }
yrt_status yrtSetFloat2(yrt_device* dev, yrt_handle h, const char* p, float x, float y) { GUARD_S(set_variant(dev, h, p, vf(2, x, y))) }
yrt_status yrtSetFloat3(yrt_device* dev, yrt_handle h, const char* p, float x, float y, float z) { GUARD_S(set_variant(dev, h, p, vf(3, x, y, z))) }
yrt_status yrtSetFloat4(yrt_device* dev, yrt_handle h, const char* p, float x, float y, float z, float w) { GUARD_S(set_variant(dev, h, p, vf(4, x, y, z, w))) }

yrt_status yrtGetFloat1(yrt_device* dev, yrt_handle h, const char* p, float* x) { GUARD_S(if (!h) return YRT_OK; const Variant v = get_variant(h, p); *x = v.f[0]) }
yrt_status yrtGetFloat3(yrt_device* dev, yrt_handle h, const char* p, float* x, float* y, float* z) {
    GUARD_S(if (!h) return YRT_OK; const Variant v = get_variant(h, p); *x = v.f[0]; *y = v.f[1]; *z = v.f[2])
}

yrt_status yrtSetArray(yrt_device* dev, yrt_handle h, const char* p, const char* type, yrt_handle data, size_t size, size_t stride, size_t ofs) {
    GUARD_S(auto* d = cast<DataHandle>(data, HK_DATA, "data");
            if (!p) throw std::runtime_error("invalid property");
            if (!h) return YRT_OK;
            static const struct { const char* name; Variant::Type t; size_t bytes; } types[] = {
                {"bool1", Variant::BOOL1, 1}, {"bool2", Variant::BOOL2, 2}, {"bool3", Variant::BOOL3, 3}, {"bool4", Variant::BOOL4, 4},
                {"int1", Variant::INT1, 4}, {"int2", Variant::INT2, 8}, {"int3", Variant::INT3, 12}, {"int4", Variant::INT4, 16},
                {"float1", Variant::FLOAT1, 4}, {"float2", Variant::FLOAT2, 8}, {"float3", Variant::FLOAT3, 12}, {"float4", Variant::FLOAT4, 16}};
            Variant v; bool found = false;
            for (auto& t : types) if (!strcasecmp(type ? type : "", t.name)) {
                v.type = t.t; v.isArray = true; v.data = d->inst; v.size = size; v.stride = stride == (size_t)-1 ? t.bytes : stride; v.ofs = ofs; found = true;
                if (size && (size - 1) * v.stride + ofs + t.bytes > d->inst->bytes) throw std::runtime_error("array view exceeds the data buffer");
            }
            if (!found) throw std::runtime_error("unknown array type: " + std::string(type ? type : ""));
            set_variant(dev, h, p, std::move(v)))
}

yrt_status yrtSetString(yrt_device* dev, yrt_handle h, const char* p, const char* str) { GUARD_S(Variant v; v.type = Variant::STRING; v.str = str ? str : ""; set_variant(dev, h, p, std::move(v))) }
yrt_status yrtGetString(yrt_device* dev, yrt_handle h, const char* p, char* buf, size_t n) {
    GUARD_S(if (!h) return YRT_OK; const Variant v = get_variant(h, p); if (n) { strncpy(buf, v.str.c_str(), n - 1); buf[n - 1] = 0; })
}
yrt_status yrtSetImage(yrt_device* dev, yrt_handle h, const char* p, yrt_handle img) {
    GUARD_S(if (!p) throw std::runtime_error("invalid property"); if (!h) return YRT_OK;
            Handle* ih = (Handle*)img;
            if (!ih || ih->magic != HANDLE_MAGIC || ih->kind != HK_IMAGE) throw std::runtime_error("invalid image handle");
            Variant v; v.type = Variant::IMAGE; v.image = static_cast<ImageHandle*>(ih)->inst;
            if (!v.image) throw std::runtime_error("invalid image value");
            set_variant(dev, h, p, std::move(v)))
}
yrt_status yrtSetTexture(yrt_device* dev, yrt_handle h, const char* p, yrt_handle tex) {
    GUARD_S(if (!p) throw std::runtime_error("invalid property"); if (!h) return YRT_OK;
            auto* t = cast<TextureHandle>(tex, HK_TEXTURE, "texture");
            Variant v; v.type = Variant::TEXTURE; v.texture = t->inst;      // the instance at the time of the call (may be null: texture ignored)
            set_variant(dev, h, p, std::move(v)))
}
yrt_status yrtSetTransform(yrt_device* dev, yrt_handle h, const char* p, const float* t) {
    GUARD_S(Variant v; v.type = Variant::TRANSFORM; if (t) memcpy(v.f, t, 48); else { const Aff3 i = aff3_identity(); aff3_to_array(i, v.f); } set_variant(dev, h, p, std::move(v)))
}
yrt_status yrtGetTransform(yrt_device* dev, yrt_handle h, const char* p, float* out) {
    GUARD_S(if (!h) return YRT_OK; const Variant v = get_variant(h, p); memcpy(out, v.f, 48))
}
yrt_status yrtClear(yrt_device* dev, yrt_handle h) {
    GUARD_S(if (!h) throw std::runtime_error("invalid handle"); Handle* b = anyHandle(h);
            if (b->kind == HK_SHAPE && b->type == "trianglemesh") { b->parms.m.clear(); b->modified = true; })   // only CreateHandle clears (api/handle.h:136-139)
}
yrt_status yrtCommit(yrt_device* dev, yrt_handle h) { GUARD_S(if (!h) throw std::runtime_error("invalid handle"); commit_handle(dev, anyHandle(h))) }

// ---- render calls ------------------------------------------------------------------------------------
yrt_status yrtRenderFrame(yrt_device* dev, yrt_handle renderer, yrt_handle camera, yrt_handle scene, yrt_handle tonemapper, yrt_handle fb, int accumulate) {
    GUARD_S(render_frame(dev, cast<RendererHandle>(renderer, HK_RENDERER, "renderer"), cast<CameraHandle>(camera, HK_CAMERA, "camera"),
                         cast<SceneHandle>(scene, HK_SCENE, "scene"), cast<ToneMapperHandle>(tonemapper, HK_TONEMAPPER, "tonemapper"),
                         cast<FrameBufferHandle>(fb, HK_FRAMEBUFFER, "framebuffer"), accumulate))
}

int yrtPick(yrt_device* dev, yrt_handle camera, float x, float y, yrt_handle scene, float* px, float* py, float* pz) {
    try {
        if (!dev) throw std::runtime_error("invalid device");
        std::lock_guard<std::mutex> lock(dev->mutex); dev->bind();
        auto* c = cast<CameraHandle>(camera, HK_CAMERA, "camera"); auto* s = cast<SceneHandle>(scene, HK_SCENE, "scene");
        if (!c->inst) throw std::runtime_error("invalid camera value");
        // Camera::ray(Vec2f(x,y), Vec2f(0.5,0.5))  (singleray_device.cpp:692-708), any camera model, evaluated on the host
        V3 org, dir; camera_ray(*c->inst, x, y, 0.5f, 0.5f, org, dir);
        Aff3 m; m.p = org;
        float ray[8] = {m.p.x, m.p.y, m.p.z, 0.f, dir.x, dir.y, dir.z, INFINITY};
        float hit[8]; int32_t* hi = (int32_t*)hit; hi[3] = -1;
        trace_rays(dev, s, 1, ray, hit, 1, 0, nullptr);
        const V3 P = m.p + hit[0] * dir;
        *px = P.x; *py = P.y; *pz = P.z;
        return hi[3] >= 0 ? 1 : 0;
    } catch (const std::exception& e) { g_lastError = e.what(); return -1; }
}

// ---- extensions --------------------------------------------------------------------------------------
yrt_status yrtxGetFrameStats(yrt_device* dev, yrtx_frame_stats* out) { GUARD_S(*out = dev->stats) }
yrt_status yrtxTraceRays(yrt_device* dev, yrt_handle scene, size_t n, const float* rays, void* hits, int closest, int onDevice, float* ms) {
    GUARD_S(trace_rays(dev, cast<SceneHandle>(scene, HK_SCENE, "scene"), n, rays, hits, closest, onDevice, ms))
}
yrt_status yrtxPrimaryRays(yrt_device* dev, yrt_handle renderer, yrt_handle camera, yrt_handle fb, float* rays, int* sets) {
    GUARD_S(primary_rays(dev, cast<RendererHandle>(renderer, HK_RENDERER, "renderer"), cast<CameraHandle>(camera, HK_CAMERA, "camera"),
                         cast<FrameBufferHandle>(fb, HK_FRAMEBUFFER, "framebuffer"), rays, sets))
}
yrt_status yrtxSampleTable(yrt_device* dev, yrt_handle renderer, yrt_handle scene, int iteration, int* sets, int* spp, int* n1, int* n2, float* table) {
    GUARD_S(sample_table(dev, cast<RendererHandle>(renderer, HK_RENDERER, "renderer"), scene ? cast<SceneHandle>(scene, HK_SCENE, "scene") : nullptr,
                         iteration, sets, spp, n1, n2, table))
}
yrt_status yrtxHostSampleTable(const char* filter, int spp, int sets, int maxDepth, int iteration, int* outSpp, int* n1, int* n2, float* table) {
    try {
        const std::string f(filter ? filter : "bspline");
        const int kind = f == "none" ? FILTER_NONE : f == "box" ? FILTER_BOX : f == "bspline" ? FILTER_BSPLINE : -1;
        if (kind < 0) throw std::runtime_error("unknown filter type: " + f);
        if (spp < 1 || sets < 1 || maxDepth < 0) throw std::runtime_error("invalid sample table request");
        PixelFilter pf; if (kind != FILTER_NONE) pf.init((FilterKind)kind);
        const int a = maxDepth, b = 1 + maxDepth;                       // pathtraceintegrator.cpp:39-46
        const SampleTable t = buildSampleTable(spp, sets, a, b, iteration, kind == FILTER_NONE ? nullptr : &pf);
        if (outSpp) *outSpp = t.spp; if (n1) *n1 = a; if (n2) *n2 = b;
        if (table) memcpy(table, t.rec.data(), t.rec.size() * sizeof(float));
        return YRT_OK;
    } catch (const std::exception& e) { g_lastError = e.what(); return YRT_ERROR; }
}
yrt_status yrtxFrameBufferDevice(yrt_device* dev, yrt_handle fb, void** devPtr, size_t* bytes, size_t* strideBytes) {
    GUARD_S(auto* f = cast<FrameBufferHandle>(fb, HK_FRAMEBUFFER, "framebuffer");
            if (devPtr) *devPtr = f->devPacked; if (bytes) *bytes = f->bytes(); if (strideBytes) *strideBytes = f->strideBytes)
}
yrt_status yrtxSetReadback(yrt_device* dev, int readbackEachFrame) { GUARD_S(dev->readback = readbackEachFrame != 0) }
yrt_status yrtxSetOption(yrt_device* dev, const char* key, long value) {
    GUARD_S(const std::string k(key ? key : "");
            if (k == "lanes") dev->lanes = value >= 2 ? 2 : 1;
            else if (k == "tracectas") dev->traceCtas = (int)std::max(1l, std::min(16l, value));
            else if (k == "shadectas") dev->shadeCtas = (int)std::max(1l, std::min(16l, value));
            else if (k == "syncmin") dev->syncMinPaths = (uint32_t)std::max(0l, value);
            else if (k == "timers") dev->useTimers = value != 0;
            else if (k == "verbose") dev->verbose = (int)value;
            else if (k == "rebuild") dev->alwaysRebuild = value != 0;
            else throw std::runtime_error("device_cuda: unknown option " + k))
}
yrt_status yrtxMicrobench(yrt_device* dev, int kind, size_t bytes, double* result) { GUARD_S(const double v = yrt::microbench(dev, kind, bytes); if (result) *result = v) }
yrt_status yrtxRenderCubeMap(yrt_device* dev, yrt_handle renderer, const yrt_handle* cameras, size_t numFaces, yrt_handle scene, yrt_handle tonemapper,
                             const yrt_handle* frameBuffers, int accumulate) {
    GUARD_S(if (!cameras || !frameBuffers || numFaces < 1 || numFaces > YRT_MAX_FACES) throw std::runtime_error("device_cuda: yrtxRenderCubeMap takes 1..12 cameras and frame buffers");
            CameraHandle* c[YRT_MAX_FACES]; FrameBufferHandle* f[YRT_MAX_FACES];
            for (size_t i = 0; i < numFaces; i++) { c[i] = cast<CameraHandle>(cameras[i], HK_CAMERA, "camera"); f[i] = cast<FrameBufferHandle>(frameBuffers[i], HK_FRAMEBUFFER, "framebuffer"); }
            render_frames(dev, cast<RendererHandle>(renderer, HK_RENDERER, "renderer"), numFaces, c, cast<SceneHandle>(scene, HK_SCENE, "scene"),
                          cast<ToneMapperHandle>(tonemapper, HK_TONEMAPPER, "tonemapper"), f, accumulate))
}

// ---- decoded images (tests of the format readers) ----------------------------------------------------
yrt_status yrtxReadImage(yrt_device* dev, yrt_handle image, int* width, int* height, int* format, void* pixels) {
    GUARD_S(auto* h = cast<ImageHandle>(image, HK_IMAGE, "image");
            if (!h->inst) throw std::runtime_error("invalid image value");
            if (width) *width = h->inst->width; if (height) *height = h->inst->height; if (format) *format = h->inst->format;
            if (pixels && h->inst->pixels) memcpy(pixels, h->inst->pixels, h->inst->bytes());)
}
yrt_status yrtxDecodePNGFile(const char* file, int flipVertical, int flipHorizontal, int* width, int* height, void* rgba) {
    try {
        auto img = decode_png_file(file ? file : "", flipVertical != 0, flipHorizontal != 0);
        if (!img) throw std::runtime_error(std::string("cannot decode ") + (file ? file : "(null)"));
        if (width) *width = img->width; if (height) *height = img->height;
        if (rgba) memcpy(rgba, img->storage.data(), img->storage.size());
        return YRT_OK;
    } catch (const std::exception& e) { g_lastError = e.what(); return YRT_ERROR; }
}

// ---- stereo cube-map strip on the device (image_codecs.cu) -------------------------------------------
yrt_status yrtxStripBegin(yrt_device* dev, size_t faceWidth, size_t faceHeight) { GUARD_S(strip_begin(dev, faceWidth, faceHeight)) }
yrt_status yrtxStripSetWatermark(yrt_device* dev, const char* pngFile) { GUARD_S(strip_set_watermark(dev, pngFile)) }
yrt_status yrtxStripAddFace(yrt_device* dev, yrt_handle fb, int cubeFaceIndex, int watermark) {
    GUARD_S(strip_add_face(dev, cast<FrameBufferHandle>(fb, HK_FRAMEBUFFER, "framebuffer"), cubeFaceIndex, watermark))
}
yrt_status yrtxStripRead(yrt_device* dev, void* rgb) { GUARD_S(if (!rgb) throw std::runtime_error("invalid buffer"); strip_read(dev, rgb)) }
yrt_status yrtxStripEncodeJPEG(yrt_device* dev, int cubeFaceIndex, int quality, const char* file) {
    GUARD_S(if (!file) throw std::runtime_error("invalid file name"); strip_encode_jpeg(dev, cubeFaceIndex, quality, file))
}

}  // extern "C"
