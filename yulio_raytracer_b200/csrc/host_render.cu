// host_render.cu — host side of the render path of device_cuda: scene flattening + upload, framebuffers,
// the per-frame wavefront schedule, and the measurement/parity extensions.
//
// Restates the host-side control flow of
//   BackendSceneFlat::Handle::setPrimitive/create     devices/device_singleray/api/scene_flat.h:63-112
//   BackendSceneFlat ctor (lights / env lights)        api/scene_flat.h:120-136, api/scene.h:76-86
//   SwapChain / FrameBuffer                            api/swapchain.h:29-125, api/framebuffer.h:97-230
//   IntegratorRenderer::renderFrame / RenderJob        renderers/integratorrenderer.cpp:63-116
//   PathTraceIntegrator::requestSamples                integrators/pathtraceintegrator.cpp:35-47
//   SamplerFactory::init (light samples)               samplers/sampler.cpp:141-150, lights/hdrilight.cpp:92-102
// around the CUDA stages in kernels.cu / bvh_build.cu. Nothing here computes radiance on the CPU.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <map>

#include <nvtx3/nvToolsExt.h>

#include "device_impl.hpp"
#include "host_math.hpp"
#include "bvh.cuh"

namespace yrt {

// ------------------------------------------------------------------------------------------------
// image files: PPM / PFM as the reference reads them (common/image/ppm.cpp:40-110, pfm.cpp:40-95).
// JPEG/PNG decoding is done by the front end and arrives through rtNewImage.
// ------------------------------------------------------------------------------------------------
static void skip_header_space(FILE* f) {
    for (;;) {
        const int c = fgetc(f);
        if (c == '#') { char line[1024]; if (!fgets(line, sizeof(line), f)) return; continue; }
        if (c != EOF && isspace(c)) continue;
        if (c != EOF) ungetc(c, f);
        return;
    }
}
static unsigned char quant8(float v) {              // Color4::set(Col4c&): clamp(c)*255 truncated (color_sse.h:68)
    const float c = rclamp(v, 0.f, 1.f) * 255.0f;
    return (unsigned char)(int)c;
}
std::shared_ptr<ImageObj> decode_jpeg_file(const char* file);                     // image_codecs.cu
std::shared_ptr<ImageObj> decode_png_file(const char* file, bool flipVertical, bool flipHorizontal);
std::shared_ptr<ImageObj> load_image_file(const char* file) {
    if (!file) return nullptr;
    const std::string name(file);
    const size_t dot = name.rfind('.');
    std::string ext = dot == std::string::npos ? "" : name.substr(dot + 1);
    for (auto& c : ext) c = (char)tolower((unsigned char)c);
    if (ext == "jpg" || ext == "jpeg") return decode_jpeg_file(file);             // common/image/image.cpp:41-45 (libjpeg-turbo in the reference)
    if (ext == "png") return decode_png_file(file, false, false);                 // common/image/image.cpp:50 (FreeImage in the reference)
    if (ext != "ppm" && ext != "pfm") { printf("cannot read file %s: image format %s not supported\n", file, ext.c_str()); return nullptr; }
    FILE* f = fopen(file, "rb");
    if (!f) { printf("cannot read file %s: cannot open\n", file); return nullptr; }
    auto img = std::make_shared<ImageObj>();
    char magic[8] = {0};
    bool ok = fscanf(f, "%7s", magic) == 1;
    if (ok) skip_header_space(f);
    if (ok && ext == "ppm") {
        int w = 0, h = 0, maxc = 0;
        ok = fscanf(f, "%i %i %i", &w, &h, &maxc) == 3 && w > 0 && h > 0 && maxc > 0;
        if (ok) {
            fgetc(f);
            const float rcpMax = 1.0f / float(maxc);
            img->width = w; img->height = h; img->format = TEX_RGBA8; img->storage.assign((size_t)w * h * 4, 0);
            unsigned char* o = img->storage.data();
            const bool text = !strcmp(magic, "P3"), bin = !strcmp(magic, "P6");
            ok = text || bin;
            for (size_t i = 0; ok && i < (size_t)w * h; i++) {
                float rgb[3];
                if (text) { int r, g, b; ok = fscanf(f, "%i %i %i", &r, &g, &b) == 3; rgb[0] = float(r); rgb[1] = float(g); rgb[2] = float(b); }
                else if (maxc <= 255) { unsigned char c[3]; ok = fread(c, 3, 1, f) == 1; rgb[0] = c[0]; rgb[1] = c[1]; rgb[2] = c[2]; }
                else { unsigned short c[3]; ok = fread(c, 6, 1, f) == 1; rgb[0] = c[0]; rgb[1] = c[1]; rgb[2] = c[2]; }
                for (int k = 0; k < 3; k++) o[4 * i + k] = quant8(rgb[k] * rcpMax);
                o[4 * i + 3] = 255;
            }
        }
    } else if (ok) {
        int w = 0, h = 0; float maxc = 0.f;
        ok = fscanf(f, "%i %i %f", &w, &h, &maxc) == 3 && w > 0 && h > 0 && !(maxc > 0.0f) && !strcmp(magic, "PF");
        if (ok) {
            fgetc(f);
            const float rcpMax = -1.0f / maxc;
            img->width = w; img->height = h; img->format = TEX_RGB_F32; img->storage.assign((size_t)w * h * 12, 0);
            float* o = (float*)img->storage.data();
            ok = fread(o, 12, (size_t)w * h, f) == (size_t)w * h;
            for (size_t i = 0; ok && i < (size_t)w * h * 3; i++) o[i] = o[i] * rcpMax;
        }
    }
    fclose(f);
    if (!ok) { printf("cannot read file %s: error reading\n", file); return nullptr; }
    img->pixels = img->storage.data();
    return img;
}

static void image_upload(ImageObj& img, cudaStream_t s) {
    if (img.devPixels) return;
    YRT_CK(cudaMalloc(&img.devPixels, img.bytes() ? img.bytes() : 1));
    if (img.bytes()) YRT_CK(cudaMemcpyAsync(img.devPixels, img.pixels, img.bytes(), cudaMemcpyHostToDevice, s));
}

// ------------------------------------------------------------------------------------------------
// scene
// ------------------------------------------------------------------------------------------------
SceneHandle::~SceneHandle() { releaseDevice(); }
void SceneHandle::releaseDevice() {
    if (nodes) cudaFreeAsync(nodes, nullptr); if (tris) cudaFreeAsync(tris, nullptr); if (triShade) cudaFreeAsync(triShade, nullptr);   // allocated from the stream-ordered pool (bvh_build.cu)
    if (triMotion) cudaFreeAsync(triMotion, nullptr);
    nodes = nullptr; tris = nullptr; triShade = nullptr; triMotion = nullptr;
}

// Shape::transform  shapes/trianglemesh_full.cpp:68-90, trianglemesh_normals.cpp:42-56, triangle.h:47-49
static std::shared_ptr<ShapeObj> transform_shape(const std::shared_ptr<ShapeObj>& s, const Aff3& xfm) {
    if (s->type == MESH_TRIANGLE) {
        auto t = std::make_shared<ShapeObj>(); t->type = MESH_TRIANGLE;
        t->v0 = xfmPoint(xfm, s->v0); t->v1 = xfmPoint(xfm, s->v1); t->v2 = xfmPoint(xfm, s->v2);
        t->triNg = normalize(cross(t->v2 - t->v0, t->v1 - t->v0)); t->triNgValid = true;
        return t;
    }
    if (aff3_is_identity(xfm)) return s;
    auto t = std::make_shared<ShapeObj>();
    t->type = s->type; t->cullBackFaces = s->cullBackFaces; t->texcoord = s->texcoord; t->triangles = s->triangles;
    t->position.resize(s->position.size());
    for (size_t i = 0; i < s->position.size(); i++) t->position[i] = xfmPoint(xfm, s->position[i]);
    t->normal.resize(s->normal.size());
    if (!s->normal.empty()) {
        const Lin3 it = lin3_inverse_transposed(xfm.l);
        for (size_t i = 0; i < s->normal.size(); i++) t->normal[i] = xfmVector(it, s->normal[i]);
    }
    t->motion.resize(s->motion.size());                                                      // xfmVector (trianglemesh_full.cpp:80-81)
    for (size_t i = 0; i < s->motion.size(); i++) t->motion[i] = xfmVector(xfm, s->motion[i]);
    t->tangentX.resize(s->tangentX.size()); t->tangentY.resize(s->tangentY.size());          // xfmVector (trianglemesh_full.cpp:84-87)
    for (size_t i = 0; i < s->tangentX.size(); i++) t->tangentX[i] = xfmVector(xfm, s->tangentX[i]);
    for (size_t i = 0; i < s->tangentY.size(); i++) t->tangentY[i] = xfmVector(xfm, s->tangentY[i]);
    return t;
}

// Light::transform  lights/*.h (ambient :53, directional :42, distant :51, hdri :44, point :44, spot :48, triangle :58)
static std::shared_ptr<LightObj> transform_light(const std::shared_ptr<LightObj>& l, const Aff3& xfm) {
    auto t = std::make_shared<LightObj>(*l);
    switch (l->type) {
    case LIGHT_POINT: t->v0 = xfmPoint(xfm, l->v0); break;
    case LIGHT_SPOT: t->v0 = xfmPoint(xfm, l->v0); t->v1 = xfmVector(xfm, l->v1); break;
    case LIGHT_DIRECTIONAL: case LIGHT_DISTANT: t->v0 = xfmVector(xfm, l->v0); break;
    case LIGHT_HDRI: t->local2world = mul(xfm, l->local2world); break;
    case LIGHT_TRIANGLE: t->v0 = xfmPoint(xfm, l->v0); t->v1 = xfmPoint(xfm, l->v1); t->v2 = xfmPoint(xfm, l->v2); break;
    default: break;
    }
    return t;
}

void scene_set_primitive(yrt_device* dev, SceneHandle* sc, size_t slot, PrimHandle* prim, const Aff3* overrideXfm) {
    (void)dev;
    if (slot >= sc->prims.size()) { sc->prims.resize(slot + 1); sc->structureDirty = true; }
    sc->dirty = true;
    if (!prim) { sc->prims[slot] = nullptr; sc->structureDirty = true; return; }
    const Aff3 xfm = overrideXfm ? *overrideXfm : prim->transform;
    auto sp = std::make_shared<ScenePrim>();
    std::shared_ptr<ShapeObj> shape = prim->shape ? prim->shape->inst : nullptr;
    std::shared_ptr<LightObj> light = prim->light ? prim->light->inst : nullptr;
    if (light) {                                              // light->shape(): only the triangle light has one
        shape = nullptr;
        if (light->type == LIGHT_TRIANGLE) {
            shape = std::make_shared<ShapeObj>(); shape->type = MESH_TRIANGLE;
            shape->v0 = light->v0; shape->v1 = light->v1; shape->v2 = light->v2;
        }
    }
    if (shape) shape = transform_shape(shape, xfm);
    if (light) light = transform_light(light, xfm);
    sp->shape = shape; sp->light = light;
    sp->material = prim->material ? prim->material->inst : nullptr;
    sp->illumMask = prim->illumMask; sp->shadowMask = prim->shadowMask;
    sc->prims[slot] = sp;
    sc->patchSlots.push_back(slot);
}

// scene bounds for the ray-sort keys (sort.cu)
static void scene_bounds(SceneHandle* sc) {
    V3 lo(INFINITY), hi(-INFINITY);
    for (const float4& q : sc->hostPositions) {
        if (!(std::isfinite(q.x) && std::isfinite(q.y) && std::isfinite(q.z))) continue;
        lo = V3(fminf(lo.x, q.x), fminf(lo.y, q.y), fminf(lo.z, q.z)); hi = V3(fmaxf(hi.x, q.x), fmaxf(hi.y, q.y), fmaxf(hi.z, q.z));
    }
    if (!(lo.x <= hi.x)) { lo = V3(0.f); hi = V3(1.f); }
    sc->bboxLo = lo; sc->bboxHi = hi;
    const V3 e = hi - lo;
    sc->data.bboxLo = lo; sc->data.bboxRcpExtent = V3(e.x > 0.f ? 1.f / e.x : 0.f, e.y > 0.f ? 1.f / e.y : 0.f, e.z > 0.f ? 1.f / e.z : 0.f);
}

static bool slot_matches(const SceneHandle::SlotLayout& L, const ScenePrim& p) {
    if (!p.shape || p.light || L.hasLight || L.geomID < 0 || !L.allFinite) return false;
    const ShapeObj& s = *p.shape;
    if (s.type != L.type || s.type == MESH_TRIANGLE || p.material.get() != L.material || s.cullBackFaces != L.cull) return false;
    if (L.hasTangents || !s.tangentX.empty() || !s.tangentY.empty() || !s.motion.empty()) return false;   // tangent / motion arrays are re-flattened, not patched
    if (p.illumMask != L.illumMask || p.shadowMask != L.shadowMask) return false;
    if (s.position.size() != L.nv || s.normal.size() != L.nn || s.texcoord.size() != L.nuv || s.triangles.size() != L.nt) return false;
    for (const V3& q : s.position) if (!(std::isfinite(q.x) && std::isfinite(q.y) && std::isfinite(q.z))) return false;
    return true;
}

// Vertices of the patched slots are rewritten in the committed arrays and only the BVH is rebuilt. Returns 0 when a
// slot changed more than its vertex data (then the caller re-flattens everything), 1 when vertices moved, 2 when every
// re-set slot carries exactly the vertices already committed — the 12 stereo cube cameras of a viewpoint share their
// origin, so the front end's per-face rtUpdatePrimitive + rtCommit (renderer.cpp:551-559) reproduces the committed scene
// for 11 of 12 faces: nothing to rebuild.
static int scene_patch(yrt_device* dev, SceneHandle* sc) {
    if (!sc->committed || sc->structureDirty || sc->layout.size() != sc->prims.size()) return 0;
    for (size_t slot : sc->patchSlots) {
        if (slot >= sc->prims.size() || !sc->prims[slot] || !slot_matches(sc->layout[slot], *sc->prims[slot])) return 0;
        const SceneHandle::SlotLayout& L = sc->layout[slot];
        const ShapeObj& s = *sc->prims[slot]->shape;      // same index list and texture coordinates as committed?
        if (L.nt && memcmp(s.triangles.data(), &sc->hostIndices[L.idxBase], L.nt * sizeof(int4)) != 0) return 0;
        if (L.nuv && memcmp(s.texcoord.data(), &sc->hostUvs[L.uvBase], L.nuv * sizeof(float2)) != 0) return 0;
    }
    cudaStream_t st = dev->stream;
    bool moved = false;
    for (size_t slot : sc->patchSlots) {
        const SceneHandle::SlotLayout& L = sc->layout[slot];
        const ShapeObj& s = *sc->prims[slot]->shape;
        bool slotMoved = false;
        auto put = [&](float4& dst, const V3& q) {
            if (memcmp(&dst, &q, 3 * sizeof(float)) != 0) { dst = make_float4(q.x, q.y, q.z, 0.f); slotMoved = true; }
        };
        for (size_t i = 0; i < L.nv; i++) put(sc->hostPositions[L.vtxBase + i], s.position[i]);
        for (size_t i = 0; i < L.nn; i++) put(sc->hostNormals[L.nrmBase + i], s.normal[i]);
        if (!slotMoved) continue;
        moved = true;
        if (L.nv) YRT_CK(cudaMemcpyAsync(sc->positions.p + L.vtxBase, &sc->hostPositions[L.vtxBase], L.nv * sizeof(float4), cudaMemcpyHostToDevice, st));
        if (L.nn) YRT_CK(cudaMemcpyAsync(sc->normals.p + L.nrmBase, &sc->hostNormals[L.nrmBase], L.nn * sizeof(float4), cudaMemcpyHostToDevice, st));
    }
    return moved ? 1 : 2;
}

struct NvtxRange { explicit NvtxRange(const char* name) { nvtxRangePushA(name); } ~NvtxRange() { nvtxRangePop(); } };   // timeline ranges (nsys / ncu --nvtx)

void scene_commit(yrt_device* dev, SceneHandle* sc) {
    NvtxRange nvtxCommit("device_cuda: rtCommit(scene)");
    sc->commitCount++;
    // The reference rebuilds the Embree scene on every commit (scene_flat.h:87-112); the result only
    // depends on the slots, so an unchanged scene keeps its BVH unless cfg rebuild=1 asks for the reference's cost.
    if (sc->committed && !sc->dirty && !dev->alwaysRebuild) return;
    cudaStream_t st = dev->stream;
    const auto tc0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (dev->verbose) printf("device_cuda: commit %-10s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tc0).count());
    };

    const int patched = dev->alwaysRebuild ? 0 : scene_patch(dev, sc);
    if (patched == 2) { lap("unchanged"); sc->patchSlots.clear(); sc->dirty = false; return; }
    if (patched == 1) {                                        // vertices of a few slots moved: rebuild the BVH over the patched arrays
        lap("patch");
        sc->patchSlots.clear();
        // the old BVH is released only once the new one exists; a failed build leaves the scene uncommitted (never dangling pointers)
        BvhBuildInput in{sc->refsBuf.p, sc->numRefs, sc->geoms.p, sc->positions.p, sc->indices.p, sc->normals.p, sc->uvs.p, sc->data.hasMotion ? sc->motions.p : nullptr,
                         dev->hostCounters + 8, dev->bvhPloc, dev->splitLeaves, dev->plocRadius, dev->bvhCollapseDp, dev->bvhCTri};
        BvhResult out{};
        try { build_bvh(in, out, st); }
        catch (...) { sc->releaseDevice(); sc->committed = false; sc->structureDirty = true; sc->data.nodes = nullptr; sc->data.tris = nullptr; sc->data.triShade = nullptr; sc->data.triMotion = nullptr; sc->data.numNodes = sc->data.numTris = 0; throw; }
        sc->releaseDevice();
        lap("bvh");
        sc->nodes = out.nodes; sc->tris = out.tris; sc->triShade = out.triShade; sc->triMotion = out.triMotion; sc->buildMs = out.buildMs; sc->buildLaunches = out.launches; sc->rebuildCount++;
        sc->data.nodes = sc->nodes; sc->data.tris = sc->tris; sc->data.triShade = sc->triShade; sc->data.triMotion = sc->triMotion; sc->data.numNodes = out.numNodes; sc->data.numTris = out.numTris;
        if (dev->sortRays) scene_bounds(sc);
        sc->dirty = false;
        dev->stats.build_ms = out.buildMs; dev->stats.num_triangles = out.numTris; dev->stats.num_nodes = out.numNodes; dev->stats.bvh_builds = sc->rebuildCount;
        if (dev->verbose) printf("device_cuda: BVH8 rebuild (patched) %u triangles -> %u nodes, %.3f ms\n", out.numTris, out.numNodes, out.buildMs);
        return;
    }
    sc->patchSlots.clear(); sc->structureDirty = false;
    std::vector<GeomRec> geoms; std::vector<float4>& positions = sc->hostPositions; std::vector<float4>& normals = sc->hostNormals;
    std::vector<float2>& uvs = sc->hostUvs; std::vector<int4>& indices = sc->hostIndices;
    positions.clear(); normals.clear(); uvs.clear(); indices.clear();
    sc->layout.assign(sc->prims.size(), SceneHandle::SlotLayout());
    std::vector<uint2> refs; std::vector<MaterialRec> materials; std::vector<LightRec> lights; std::vector<float4> tangents, motions;
    std::map<const MaterialObj*, int> matIndex; std::map<const TextureObj*, int> texIndex;
    sc->hostTextures.clear(); sc->imagesInUse.clear(); sc->hdri.clear(); sc->geomOfSlot.assign(sc->prims.size(), -1);

    auto texture_index = [&](const std::shared_ptr<ImageObj>& img, bool bilinear, bool invert) {
        image_upload(*img, st); sc->imagesInUse.push_back(img);
        TextureRec t; t.data = img->devPixels; t.width = img->width; t.height = img->height; t.format = img->format;
        t.bilinear = bilinear ? 1 : 0; t.invert = invert ? 1 : 0; t.pad = 0;
        sc->hostTextures.push_back(t);
        return (int)sc->hostTextures.size() - 1;
    };
    auto material_index = [&](const std::shared_ptr<MaterialObj>& m) {
        if (!m) return -1;
        auto it = matIndex.find(m.get());
        if (it != matIndex.end()) return it->second;
        MaterialRec r = m->rec;
        for (int k = 0; k < 5; k++) {
            r.tex[k] = -1;
            const auto& t = m->textures[k];
            if (!t || !t->image) continue;
            auto ti = texIndex.find(t.get());
            if (ti == texIndex.end()) ti = texIndex.emplace(t.get(), texture_index(t->image, t->bilinear, t->invert)).first;
            r.tex[k] = ti->second;
        }
        materials.push_back(r);
        return matIndex[m.get()] = (int)materials.size() - 1;
    };

    {   // one pass to size the arrays: a rebuild per cube face must not pay for vector growth
        size_t nv = 0, nt = 0;
        for (const auto& p : sc->prims) if (p && p->shape) { nv += p->shape->type == MESH_TRIANGLE ? 3 : p->shape->position.size(); nt += p->shape->type == MESH_TRIANGLE ? 1 : p->shape->triangles.size(); }
        positions.reserve(nv); normals.reserve(nv); uvs.reserve(nv); indices.reserve(nt); refs.reserve(nt); geoms.reserve(sc->prims.size());
    }
    SceneData d{};
    d.numEnvLights = 0; d.numPrecomputed = 0;
    for (size_t slot = 0; slot < sc->prims.size(); slot++) {
        const auto& p = sc->prims[slot];
        if (!p) continue;
        int lightIdx = -1;
        if (p->light) {                                        // BackendScene::add(light)  api/scene.h:76-79
            const LightObj& l = *p->light;
            LightRec r; memset(&r, 0, sizeof(r));
            r.type = l.type; r.illumMask = p->illumMask; r.shadowMask = p->shadowMask; r.precomputedId = -1;
            r.L = l.L; r.v0 = l.v0; r.v1 = l.v1; r.v2 = l.v2; r.a = l.a; r.b = l.b; r.image = -1;
            r.Ng = cross(l.v0 - l.v1, l.v2 - l.v0);           // trianglelight.h:39
            r.world2local.l = lin3_identity(); r.world2local.p = V3(0.f);
            r.isEnv = (l.type == LIGHT_AMBIENT || l.type == LIGHT_DISTANT || l.type == LIGHT_HDRI) ? 1 : 0;
            if (l.type == LIGHT_HDRI) {
                r.world2local = aff3_inverse(l.local2world);
                r.image = texture_index(l.image, false, false);
                r.precomputedId = d.numPrecomputed++;
                HdriHost h; h.image = l.image; h.L = l.L; h.local2world = l.local2world;
                const size_t w = (size_t)l.image->width, hgt = (size_t)l.image->height;   // hdrilight.cpp:48-55
                std::vector<std::vector<float>> imp(hgt, std::vector<float>(w));
                for (size_t y = 0; y < hgt; y++)
                    for (size_t x = 0; x < w; x++)
                        imp[y][x] = sinf(YRT_PI * (y + 0.5f) * rcpf(float(hgt))) * reduce_add(host_texel(*l.image, (long long)x, (long long)y));
                h.dist.init(imp, w, hgt);
                sc->hdri.push_back(std::move(h));
            }
            lightIdx = (int)lights.size();
            if (r.isEnv) {
                if (d.numEnvLights >= 8) throw std::runtime_error("device_cuda: more than 8 environment lights");
                d.envLightIdx[d.numEnvLights++] = lightIdx;
            }
            lights.push_back(r);
        }
        if (!p->shape) continue;
        const ShapeObj& s = *p->shape;
        GeomRec g; memset(&g, 0, sizeof(g));
        const int geomID = (int)geoms.size();
        const size_t refsBefore = refs.size();
        sc->geomOfSlot[slot] = geomID;
        g.type = s.type; g.material = material_index(p->material); g.areaLight = lightIdx;
        g.cull = s.cullBackFaces ? 1 : 0; g.illumMask = p->illumMask; g.shadowMask = p->shadowMask;
        g.vtxBase = (uint32_t)positions.size(); g.idxBase = (uint32_t)indices.size();
        g.nrmBase = YRT_NO_ATTR; g.uvBase = YRT_NO_ATTR; g.tanXBase = YRT_NO_ATTR; g.tanYBase = YRT_NO_ATTR; g.motBase = YRT_NO_ATTR; g.triNg = V3(0.f);
        if (s.type == MESH_TRIANGLE) {
            g.triNg = s.triNg;
            const V3 v[3] = {s.v0, s.v1, s.v2};
            for (int k = 0; k < 3; k++) positions.push_back(make_float4(v[k].x, v[k].y, v[k].z, 0.f));
            indices.push_back(make_int4(0, 1, 2, 0));
            bool fin = true; for (int k = 0; k < 3; k++) fin &= std::isfinite(v[k].x) && std::isfinite(v[k].y) && std::isfinite(v[k].z);
            if (fin) refs.push_back(make_uint2((uint32_t)geomID, 0u));
        } else {
            const size_t nv = s.position.size();
            for (const V3& q : s.position) positions.push_back(make_float4(q.x, q.y, q.z, 0.f));
            if (!s.normal.empty()) { g.nrmBase = (uint32_t)normals.size(); for (const V3& q : s.normal) normals.push_back(make_float4(q.x, q.y, q.z, 0.f)); }
            if (!s.texcoord.empty()) { g.uvBase = (uint32_t)uvs.size(); for (const float2& q : s.texcoord) uvs.push_back(q); }
            if (!s.motion.empty()) { g.motBase = (uint32_t)motions.size(); for (const V3& q : s.motion) motions.push_back(make_float4(q.x, q.y, q.z, 0.f)); }
            if (!s.tangentX.empty()) { g.tanXBase = (uint32_t)tangents.size(); for (const V3& q : s.tangentX) tangents.push_back(make_float4(q.x, q.y, q.z, 0.f)); }
            if (!s.tangentY.empty()) { g.tanYBase = (uint32_t)tangents.size(); for (const V3& q : s.tangentY) tangents.push_back(make_float4(q.x, q.y, q.z, 0.f)); }
            for (size_t t = 0; t < s.triangles.size(); t++) {
                const int4 tri = s.triangles[t];
                indices.push_back(tri);
                bool okTri = tri.x >= 0 && tri.y >= 0 && tri.z >= 0 && (size_t)tri.x < nv && (size_t)tri.y < nv && (size_t)tri.z < nv;
                if (okTri) for (int k : {tri.x, tri.y, tri.z}) { const V3& q = s.position[k]; okTri &= std::isfinite(q.x) && std::isfinite(q.y) && std::isfinite(q.z); }
                if (okTri) refs.push_back(make_uint2((uint32_t)geomID, (uint32_t)t));   // invalid triangles are never hit
            }
        }
        geoms.push_back(g);
        SceneHandle::SlotLayout& L = sc->layout[slot];
        L.material = p->material.get(); L.type = s.type; L.geomID = geomID; L.illumMask = p->illumMask; L.shadowMask = p->shadowMask;
        L.hasLight = (bool)p->light; L.cull = s.cullBackFaces; L.vtxBase = g.vtxBase; L.nrmBase = g.nrmBase; L.uvBase = g.uvBase; L.idxBase = g.idxBase;
        if (s.type != MESH_TRIANGLE) { L.nv = s.position.size(); L.nn = s.normal.size(); L.nuv = s.texcoord.size(); L.nt = s.triangles.size(); }
        L.allFinite = s.type != MESH_TRIANGLE && refs.size() - refsBefore == s.triangles.size();
        L.hasTangents = !s.tangentX.empty() || !s.tangentY.empty() || !s.motion.empty();
    }

    lap("flatten");
    sc->geoms.upload(geoms, st); sc->positions.upload(positions, st); sc->normals.upload(normals, st); sc->uvs.upload(uvs, st);
    sc->indices.upload(indices, st); sc->materials.upload(materials, st); sc->lights.upload(lights, st); sc->tangents.upload(tangents, st); sc->motions.upload(motions, st);
    sc->textures.upload(sc->hostTextures, st);
    DevBuf<uint2>& dRefs = sc->refsBuf; dRefs.upload(refs, st);      // kept across commits: no cudaMalloc/cudaFree per cube face
    sc->numRefs = (uint32_t)refs.size();

    lap("upload");
    sc->committed = false;                                    // until the new structure exists (the arrays above were replaced already)
    sc->releaseDevice();
    sc->data.nodes = nullptr; sc->data.tris = nullptr; sc->data.triShade = nullptr; sc->data.numNodes = sc->data.numTris = 0;
    BvhBuildInput in{dRefs.p, (uint32_t)refs.size(), sc->geoms.p, sc->positions.p, sc->indices.p, sc->normals.p, sc->uvs.p, motions.empty() ? nullptr : sc->motions.p,
                     dev->hostCounters + 8, dev->bvhPloc, dev->splitLeaves, dev->plocRadius, dev->bvhCollapseDp, dev->bvhCTri};
    BvhResult out{};
    sc->structureDirty = true;                                // a throw below leaves a scene that re-flattens on the next commit
    build_bvh(in, out, st);
    YRT_CK(cudaStreamSynchronize(st));
    sc->structureDirty = false;
    lap("bvh");
    sc->nodes = out.nodes; sc->tris = out.tris; sc->triShade = out.triShade; sc->triMotion = out.triMotion; sc->buildMs = out.buildMs; sc->buildLaunches = out.launches; sc->rebuildCount++;

    d.nodes = sc->nodes; d.tris = sc->tris; d.triShade = sc->triShade; d.triMotion = sc->triMotion; d.hasMotion = motions.empty() ? 0 : 1; d.motions = sc->motions.p;
    d.numNodes = out.numNodes; d.numTris = out.numTris;
    d.geoms = sc->geoms.p; d.positions = sc->positions.p; d.normals = sc->normals.p; d.uvs = sc->uvs.p; d.indices = sc->indices.p; d.tangents = sc->tangents.p;
    d.materials = sc->materials.p; d.textures = sc->textures.p; d.lights = sc->lights.p;
    d.numGeoms = (int)geoms.size(); d.numLights = (int)lights.size();
    {   // shading classes for the CTA-local regrouping of k_shade: same material kind / textured-ness / lobe set -> same class
        std::map<uint64_t, int> classOf;
        for (GeomRec& g : geoms) {
            uint64_t sig = 0;
            if (g.material >= 0) {
                const MaterialRec& m = materials[g.material];
                sig = (uint64_t)m.type | ((uint64_t)(m.tex[0] >= 0) << 8) | ((uint64_t)(m.tex[1] >= 0) << 9);
                if (m.type == MAT_UBER) sig |= (uint64_t)(m.f[2] > 0.f ? 1 : (m.f[1] == 0.f ? 2 : 3)) << 10;
                if (m.type == MAT_UBER && m.tex[0] >= 0 && sc->hostTextures[m.tex[0]].format != TEX_RGBA8 && sc->hostTextures[m.tex[0]].format != TEX_RGBA_F32) sig |= 1ull << 12;  // no alpha lobe
            }
            sig |= (uint64_t)(g.areaLight >= 0) << 16;
            auto it = classOf.find(sig);
            if (it == classOf.end()) it = classOf.emplace(sig, 1 + (int)(classOf.size() % 14)).first;
            g.shadeClass = it->second;
        }
    }
    d.hasMedia = 0; for (const MaterialRec& m : materials) d.hasMedia |= m.isMediaInterface;
    // the EXT shading kernel also serves Obj materials with a bump map (they need the tangent frame, materials/obj.h:52-56)
    d.hasExtMaterials = 0; for (const MaterialRec& m : materials) d.hasExtMaterials |= (m.type >= MAT_PLASTIC || (m.type == MAT_OBJ && m.tex[4] >= 0)) ? 1 : 0;
    sc->data = d;
    scene_bounds(sc);
    sc->committed = true; sc->dirty = false;
    dev->stats.build_ms = out.buildMs; dev->stats.num_triangles = out.numTris; dev->stats.num_nodes = out.numNodes;
    dev->stats.bvh_builds = sc->rebuildCount;
    lap("done");
    if (dev->verbose) printf("device_cuda: BVH8 build %u triangles -> %u nodes, %.3f ms\n", out.numTris, out.numNodes, out.buildMs);
}

// ------------------------------------------------------------------------------------------------
// framebuffers
// ------------------------------------------------------------------------------------------------
FrameBufferHandle::~FrameBufferHandle() {
    for (size_t i = 0; i < host.size(); i++) if (owned[i] && host[i]) cudaFreeHost(host[i]);
    if (devPacked) cudaFree(devPacked); if (accum) cudaFree(accum);
}

FrameBufferHandle* framebuffer_create(yrt_device* dev, const char* type, size_t w, size_t h, size_t buffers, void** ptrs) {
    (void)dev;
    auto fb = std::unique_ptr<FrameBufferHandle>(new FrameBufferHandle());
    const std::string t(type ? type : "");
    if (!strcasecmp(t.c_str(), "RGB_FLOAT32")) { fb->format = 0; fb->strideBytes = w * 12; }          // framebuffer.h:106
    else if (!strcasecmp(t.c_str(), "RGBA8")) { fb->format = 1; fb->strideBytes = w * 4; }             // framebuffer.h:146
    else if (!strcasecmp(t.c_str(), "RGB8")) { fb->format = 2; fb->strideBytes = (3 * w + 3) / 4 * 4; } // framebuffer.h:195
    else throw std::runtime_error("unknown framebuffer type: " + t);
    fb->type = t; fb->width = w; fb->height = h; fb->depth = buffers ? buffers : 1;
    for (size_t i = 0; i < fb->depth; i++) {
        void* p = ptrs ? ptrs[i] : nullptr; bool own = false;
        if (!p) { YRT_CK(cudaHostAlloc(&p, fb->bytes() ? fb->bytes() : 1, cudaHostAllocDefault)); own = true; }
        memset(p, 0, fb->bytes());
        fb->host.push_back(p); fb->owned.push_back(own);
    }
    YRT_CK(cudaMalloc(&fb->devPacked, fb->bytes() ? fb->bytes() : 1));
    YRT_CK(cudaMemset(fb->devPacked, 0, fb->bytes()));
    YRT_CK(cudaMalloc((void**)&fb->accum, (w * h ? w * h : 1) * sizeof(float4)));
    YRT_CK(cudaMemset(fb->accum, 0, w * h * sizeof(float4)));
    return fb.release();
}

void* framebuffer_map(yrt_device* dev, FrameBufferHandle* fb, int bufID) {
    if (bufID < 0) bufID = (int)fb->cur;
    if ((size_t)bufID >= fb->depth) throw std::runtime_error("invalid framebuffer index");
    if (!dev->readback && fb->pendingBuf == bufID) {
        YRT_CK(cudaMemcpyAsync(fb->host[bufID], fb->devPacked, fb->bytes(), cudaMemcpyDeviceToHost, dev->stream));
        YRT_CK(cudaStreamSynchronize(dev->stream));
        fb->pendingBuf = -1;
    }
    return fb->host[bufID];
}

// ------------------------------------------------------------------------------------------------
// wavefront storage / timers
// ------------------------------------------------------------------------------------------------
template <typename T> static void dev_realloc(T*& p, size_t n) { if (p) cudaFree(p); p = nullptr; YRT_CK(cudaMalloc((void**)&p, (n ? n : 1) * sizeof(T))); }

void WavefrontStorage::ensure(uint32_t capacity, uint32_t shadowCapacity, size_t pixels) {
    if (capacity > wb.capacity) {
        dev_realloc(wb.rayO, capacity); dev_realloc(wb.rayD, capacity); dev_realloc(wb.hitA, capacity);
        dev_realloc(wb.thr, capacity); dev_realloc(wb.Lacc, capacity); dev_realloc(wb.medium, capacity);
        dev_realloc(wb.shadowPid, capacity); dev_realloc(wb.queueA, capacity); dev_realloc(wb.queueB, capacity);
        dev_realloc(wb.queueS, capacity); dev_realloc(wb.sortKeys, capacity); dev_realloc(wb.sortKeysOut, capacity);
        if (sortTemp) cudaFree(sortTemp);
        sortTempBytes = sort_temp_bytes(capacity); sortTemp = nullptr; YRT_CK(cudaMalloc(&sortTemp, sortTempBytes ? sortTempBytes : 1));
        wb.capacity = capacity;
    }
    if (shadowCapacity > wb.shadowCapacity) {
        dev_realloc(wb.shO, shadowCapacity); dev_realloc(wb.shD, shadowCapacity); dev_realloc(wb.shC, shadowCapacity);
        wb.shadowCapacity = shadowCapacity;
    }
    if (!wb.counters) { dev_realloc(wb.counters, 16); YRT_CK(cudaMemset(wb.counters, 0, 16 * sizeof(uint32_t))); }
    if (!wb.stats) { dev_realloc(wb.stats, 8); YRT_CK(cudaMemset(wb.stats, 0, 8 * sizeof(unsigned long long))); }
    if (pixels > pixelSetCapacity) { dev_realloc(wb.pixelSet, pixels); pixelSetCapacity = pixels; }
}
void WavefrontStorage::release() {
    void* ps[] = {wb.rayO, wb.rayD, wb.hitA, wb.thr, wb.Lacc, wb.medium, wb.shadowPid, wb.queueA, wb.queueB,
                  wb.shO, wb.shD, wb.shC, wb.counters, wb.stats, wb.pixelSet, wb.queueS, wb.sortKeys, wb.sortKeysOut, sortTemp};
    for (void* p : ps) if (p) cudaFree(p);
    wb = WavefrontBuffers{}; pixelSetCapacity = 0; sortTemp = nullptr; sortTempBytes = 0;
}

cudaEvent_t FrameTimers::get() {
    if (used == pool.size()) { cudaEvent_t e; YRT_CK(cudaEventCreate(&e)); pool.push_back(e); }
    return pool[used++];
}
void FrameTimers::begin(int kind, cudaStream_t s) { Span sp; sp.kind = kind; sp.a = get(); sp.b = nullptr; YRT_CK(cudaEventRecord(sp.a, s)); spans.push_back(sp); }
void FrameTimers::end(cudaStream_t s) { Span& sp = spans.back(); sp.b = get(); YRT_CK(cudaEventRecord(sp.b, s)); }
void FrameTimers::release() { for (auto e : pool) cudaEventDestroy(e); pool.clear(); used = 0; spans.clear(); }

enum { TK_RAYGEN_FILM = 0, TK_CLOSEST = 1, TK_SHADE = 2, TK_SHADOW = 3, TK_SORT = 4, TK_RESOLVE = 5, TK_MISS = 6 };

// ------------------------------------------------------------------------------------------------
// per-frame setup shared by render_frame / primary_rays / sample_table
// ------------------------------------------------------------------------------------------------
struct FrameSetup {
    FrameConst fc{}; int sets = 64; size_t bufferRows = 0; size_t tableBytes = 0; bool tableUploaded = false;
};

static const PixelFilter* device_filter(yrt_device* dev, int kind) {
    if (kind == FILTER_NONE) return nullptr;
    if (!dev->filterReady[kind]) { dev->filters[kind].init((FilterKind)kind); dev->filterReady[kind] = true; }
    return &dev->filters[kind];
}

// SamplerFactory::init + the precomputed light samples; record = {pixel, time, lens, 1D[n1], 2D[n2], light[numPre] x 8}
static std::vector<float> build_table(yrt_device* dev, RendererObj& R, const SceneHandle* sc, int iteration, int& spp, int& n1, int& n2, int& recFloats) {
    n1 = R.maxDepth; n2 = 1 + R.maxDepth;                     // requestSamples: pathtraceintegrator.cpp:39-46
    const SampleTable t = buildSampleTable(R.spp, R.sets, n1, n2, iteration, device_filter(dev, R.filter));
    spp = t.spp;
    const int numPre = sc ? (int)sc->hdri.size() : 0;
    const int base = t.recFloats();
    recFloats = base + 8 * numPre;
    std::vector<float> out((size_t)R.sets * spp * recFloats, 0.f);
    for (int set = 0; set < R.sets; set++)
        for (int s = 0; s < spp; s++) {
            float* o = &out[((size_t)set * spp + s) * recFloats];
            memcpy(o, t.at(set, s), base * sizeof(float));
            const float sx = o[5 + n1 + 0], sy = o[5 + n1 + 1];   // samples2D[lightSampleID = 0]
            for (int k = 0; k < numPre; k++) {                    // HDRILight::sample  lights/hdrilight.cpp:92-102
                const HdriHost& h = sc->hdri[k];
                float px, py, pdf; h.dist.sample(sx, sy, px, py, pdf);
                const float W = float(h.image->width), H = float(h.image->height);
                const float theta = YRT_PI * py * rcpf(H);
                const float phi = YRT_TWO_PI * (1.0f - px * rcpf(W));
                const V3 wl(-sinf(theta) * cosf(phi), cosf(theta), -sinf(theta) * sinf(phi));
                const V3 wi = xfmVector(h.local2world.l, wl);
                const float wpdf = pdf * rcpf(YRT_TWO_PI * YRT_PI * sinf(theta));
                long long ix = (long long)px, iy = (long long)py;
                ix = ix < 0 ? 0 : (ix > h.image->width - 1 ? h.image->width - 1 : ix);
                iy = iy < 0 ? 0 : (iy > h.image->height - 1 ? h.image->height - 1 : iy);
                const Col L = h.L * host_texel(*h.image, ix, iy);
                float* q = o + base + 8 * k;
                q[0] = wi.x; q[1] = wi.y; q[2] = wi.z; q[3] = wpdf; q[4] = L.x; q[5] = L.y; q[6] = L.z; q[7] = 0.f;
            }
        }
    return out;
}

static size_t active_rows(int height, int serverID, int serverCount) {
    size_t n = 0;
    for (int y = 0; y < height; y++) if ((((y >> 2) - serverID) % serverCount) == 0) n++;
    return n;
}

static FrameSetup setup_frame(yrt_device* dev, RendererObj& R, SceneHandle* sc, FrameBufferHandle* fb, int iteration, int numFaces = 1) {
    FrameSetup fs;
    FrameConst& fc = fs.fc;
    if (sc) fc.scene = sc->data;
    fc.scene.tuneRefillMin = dev->tuneRefillMin; fc.scene.tuneTriNum = dev->tuneTriNum; fc.scene.tuneTriDen = dev->tuneTriDen; fc.scene.tuneSimple = dev->tuneSimple;
    fc.scene.tunePrefetch = dev->tunePrefetch > 0 ? 1 : 0;
    fc.width = (int)fb->width; fc.height = (int)fb->height; fc.numFaces = numFaces;
    fc.serverID = dev->serverID; fc.serverCount = dev->serverCount < 1 ? 1 : dev->serverCount;
    fc.rcpWidth = rcpf(float(fb->width)); fc.rcpHeight = rcpf(float(fb->height));
    fc.debugRenderer = R.debug ? 1 : 0; fc.countStats = dev->countStats;
    fs.bufferRows = active_rows(fc.height, fc.serverID, fc.serverCount);
    fc.pixelsPerFace = (uint32_t)(fs.bufferRows * fb->width);
    IntegratorData& ig = fc.integ;
    ig.maxDepth = R.maxDepth; ig.rrDepth = R.rrDepth; ig.minContribution = R.minContribution; ig.epsilon = R.epsilon;
    ig.tMaxShadowRay = R.tMaxShadowRay; ig.tMaxShadowJitter = R.tMaxShadowJitter; ig.up = R.up;
    ig.backplateTex = -1; ig.lightSampleID = 0; ig.firstScatterSampleID = 1; ig.firstScatterTypeSampleID = 0;
    ig.sets = R.sets; fs.sets = R.sets;
    if (R.debug) { ig.spp = R.spp; ig.recFloats = 0; return fs; }

    // sample table (cached while renderer parameters, iteration and the scene's precomputed lights are unchanged)
    // the table depends on the scene only through the precomputed (HDRI) light samples: without them a rebuilt scene — every cube face
    // of a scene with billboards — keeps the table (building + uploading it was ~5 ms per face at 64 spp, depth 10)
    const uint64_t sceneKey = (sc && !sc->hdri.empty()) ? sc->rebuildCount * 1315423911ull + (uint64_t)(uintptr_t)sc : 0;
    TableKey key{R.spp, R.sets, R.maxDepth, R.filter, iteration, sc ? (int)sc->hdri.size() : 0, sceneKey};
    if (!(key == dev->tableKey) || !dev->sampleTable.p) {
        int spp, n1, n2, rec;
        const std::vector<float> tab = build_table(dev, R, sc, iteration, spp, n1, n2, rec);
        dev->sampleTable.upload(tab, dev->stream);
        YRT_CK(cudaStreamSynchronize(dev->stream));          // `tab` is pageable and dies here
        dev->tableKey = key; dev->tableSpp = spp; dev->tableN1 = n1; dev->tableN2 = n2; dev->tableRec = rec;
        fs.tableBytes = tab.size() * sizeof(float); fs.tableUploaded = true;
    }
    R.spp = dev->tableSpp;                                    // SamplerFactory::init rounds samplesPerPixel up for good (sampler.cpp:91)
    ig.spp = dev->tableSpp; ig.recFloats = dev->tableRec; ig.off1D = 5; ig.off2D = 5 + dev->tableN1;
    ig.offLight = 5 + dev->tableN1 + 2 * dev->tableN2;
    fc.sampleTable = dev->sampleTable.p;
    return fs;
}

// ------------------------------------------------------------------------------------------------
// rtRenderFrame
// ------------------------------------------------------------------------------------------------
typedef void (*StatusFn)(const void*);
struct StatusRec { int state; float progress; };               // RendererStatus  devices/device/device.h:341-344

// rtRenderFrame (numFaces = 1) and yrtxRenderCubeMap (the faces of one viewpoint, same scene / renderer / tone mapper / frame size):
// all faces are rendered as one wavefront (device_internal.hpp: FrameConst::numFaces).
void render_frames(yrt_device* dev, RendererHandle* rh, size_t numFaces, CameraHandle* const* chs, SceneHandle* sc, ToneMapperHandle* th,
                   FrameBufferHandle* const* fbs, int accumulate) {
    NvtxRange nvtxFrame(numFaces > 1 ? "device_cuda: yrtxRenderCubeMap" : "device_cuda: rtRenderFrame");
    const auto tHost0 = std::chrono::steady_clock::now();
    auto hostLap = [&](const char* what) { if (dev->verbose >= 3) printf("  host %-10s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tHost0).count()); };
    if (!rh->inst) throw std::runtime_error("invalid renderer value");
    if (numFaces < 1 || numFaces > YRT_MAX_FACES) throw std::runtime_error("device_cuda: a render call takes 1.." + std::to_string(YRT_MAX_FACES) + " faces");
    for (size_t f = 0; f < numFaces; f++) {
        if (!chs[f]->inst) throw std::runtime_error("invalid camera value");
        if (fbs[f]->width != fbs[0]->width || fbs[f]->height != fbs[0]->height || fbs[f]->format != fbs[0]->format)
            throw std::runtime_error("device_cuda: the frame buffers of one render call must have the same size and format");
        for (size_t g = 0; g < f; g++) if (fbs[g] == fbs[f]) throw std::runtime_error("device_cuda: every face of a render call needs its own frame buffer");
    }
    FrameBufferHandle* fb = fbs[0];
    if (!th->inst) throw std::runtime_error("invalid tonemapper value");
    if (!sc->committed) throw std::runtime_error("invalid scene value");
    RendererObj& R = *rh->inst;
    cudaStream_t st = dev->stream;
    StatusRec status{1, 0.f};                                  // updateStatus(Rendering)  integratorrenderer.cpp:65
    StatusFn statusFn = (StatusFn)R.statusCallback;
    if (statusFn) statusFn(&status);
    if (accumulate == 0) R.iteration = 0;
    const int iteration = R.iteration++;
    volatile bool* stopFlag = (volatile bool*)R.stopFlag;      // std::atomic<bool>* in the caller (integratorrenderer.h:100-101)
    auto stopRequested = [&]() { return stopFlag && *stopFlag; };

    FrameSetup fs = setup_frame(dev, R, sc, fb, iteration, (int)numFaces);
    FrameConst& fc = fs.fc;
    FrameCameras cams; memset(&cams, 0, sizeof(cams));
    for (size_t f = 0; f < numFaces; f++) cams.cam[f] = *chs[f]->inst;
    if (R.backplate && !R.debug) {                             // backplate image joins the scene's texture table on first use
        int idx = -1;
        for (size_t i = 0; i < sc->hostTextures.size(); i++) if (sc->hostTextures[i].data == R.backplate->devPixels && R.backplate->devPixels) idx = (int)i;
        if (idx < 0) {
            image_upload(*R.backplate, st); sc->extraImages.push_back(R.backplate);
            TextureRec t; t.data = R.backplate->devPixels; t.width = R.backplate->width; t.height = R.backplate->height;
            t.format = R.backplate->format; t.bilinear = 0; t.invert = 0; t.pad = 0;
            sc->hostTextures.push_back(t); idx = (int)sc->hostTextures.size() - 1;
            sc->textures.release(); sc->textures.upload(sc->hostTextures, st);
            sc->data.textures = sc->textures.p; fc.scene.textures = sc->textures.p;
        }
        fc.integ.backplateTex = idx;
    }

    const int spp = fc.integ.spp;
    const size_t facePixels = fs.bufferRows * fb->width;
    const size_t numPixels = facePixels * numFaces;          // the virtual pixel range of the call
    if ((uint64_t)numPixels >= 0xfff00000ull) throw std::runtime_error("device_cuda: render call exceeds 2^32 pixels");
    const int nl = fc.scene.numLights > 0 ? fc.scene.numLights : 1;
    // chunk size: paths per wavefront pass, bounded so that one shadow-ray slot per (path, light) fits
    uint64_t capacity = dev->chunkPaths;
    const uint64_t totalPaths = (uint64_t)numPixels * spp;
    if (capacity > totalPaths) capacity = totalPaths;
    // two lanes (device_impl.hpp): a call of at least 2^22 paths is cut into at least two chunks so that both streams have work
    const int nLanes = (dev->lanes >= 2 && !R.debug && !dev->sortRays && totalPaths >= (1ull << 22)) ? 2 : 1;
    if (nLanes == 2 && capacity > (totalPaths + 1) / 2) capacity = (totalPaths + 1) / 2;
    WavefrontStorage* const W[2] = {&dev->wf, &dev->wf1};
    const cudaStream_t LS[2] = {dev->stream, dev->stream1};
    {
        bool grow = false;
        for (int l = 0; l < nLanes; l++) grow |= capacity > W[l]->wb.capacity || capacity * nl > W[l]->wb.shadowCapacity;
        if (grow) {
            // growing: 112 B of path state + 48 B per (path, light) shadow slot per lane; stay within 40 % of what is free right now
            size_t freeB = 0, totalB = 0; YRT_CK(cudaMemGetInfo(&freeB, &totalB));
            uint64_t have = 0;
            for (int l = 0; l < nLanes; l++) have += (uint64_t)W[l]->wb.capacity * 112ull + (uint64_t)W[l]->wb.shadowCapacity * 48ull;   // already ours
            const uint64_t budget = (uint64_t)(0.4 * (double)freeB) + have;
            while (nLanes * capacity * (112ull + 48ull * nl) > budget && capacity > 65536) capacity >>= 1;
        }
    }
    while (capacity * nl > 0xfff00000ull && capacity > 65536) capacity >>= 1;          // 32-bit shadow-slot indices
    if (capacity < (uint64_t)spp) capacity = spp;
    const uint32_t pixelsPerChunk = (uint32_t)((capacity + spp - 1) / spp);
    capacity = (uint64_t)pixelsPerChunk * spp;
    for (int l = 0; l < nLanes; l++) W[l]->ensure((uint32_t)capacity, (uint32_t)(capacity * nl), facePixels);

    hostLap("setup");
    const int traceCtas = std::max(1, dev->traceCtas / nLanes), shadeCtas = std::max(1, dev->shadeCtas / nLanes), streamCtas = 8 / nLanes;
    FrameTimers& tm = dev->timers; tm.reset();
    const bool timers = dev->useTimers != 0;
    uint64_t launches = 0, closestLaunches = 0, shadowLaunches = 0, shadeLaunches = 0;
    cudaEvent_t evStart = tm.get(), evStop = tm.get();
    YRT_CK(cudaEventRecord(evStart, st));
    for (int l = 0; l < nLanes; l++) {
        if (l > 0) YRT_CK(cudaStreamWaitEvent(LS[l], evStart, 0));          // the timed span starts on the device's first stream
        YRT_CK(cudaMemsetAsync(W[l]->wb.stats, 0, 8 * sizeof(unsigned long long), LS[l]));
    }

    FilmParams fp; memset(&fp, 0, sizeof(fp)); fp.format = fb->format; fp.fbStrideBytes = (int)fb->strideBytes;
    for (size_t f = 0; f < numFaces; f++) { fp.face[f].accum = fbs[f]->accum; fp.face[f].fb = fbs[f]->devPacked; }
    fp.accumulate = accumulate ? 1 : 0; fp.gamma = th->inst->gamma; fp.rcpGamma = rcpf(th->inst->gamma); fp.vignetting = th->inst->vignetting ? 1 : 0;

    bool stopped = false;
    if (R.debug) {
        launch_debug(fc, cams, W[0]->wb, fp, (uint32_t)numPixels, LaunchCfg{dev->numSMs * 8, 256, st}); launches++;
    } else if (numPixels) {
        for (int l = 0; l < nLanes; l++) {
            launch_pixel_sets(fc, W[l]->wb.pixelSet, fs.sets, LaunchCfg{dev->numSMs * streamCtas, 256, LS[l]}); launches++;
            if (fc.integ.maxDepth <= 0) YRT_CK(cudaMemsetAsync(W[l]->wb.Lacc, 0, (size_t)capacity * sizeof(float4), LS[l]));   // no bounce writes the radiance
        }
        std::vector<cudaEvent_t> chunkDone;
        size_t chunkIdx = 0;
        for (size_t pixelBegin = 0; pixelBegin < numPixels; pixelBegin += pixelsPerChunk, chunkIdx++) {
            const int lane = (int)(chunkIdx % (size_t)nLanes);
            const cudaStream_t ls = LS[lane];
            const WavefrontBuffers& wb = W[lane]->wb;
            const LaunchCfg lcTrace{dev->numSMs * traceCtas, 128, ls}, lcStream{dev->numSMs * streamCtas, 256, ls}, lcShade{dev->numSMs * shadeCtas, 128, ls};
            // keep two chunks per lane in flight so that stopFlag / statusCallback stay responsive (integratorrenderer.cpp:125,178)
            if (chunkIdx >= (size_t)(2 * nLanes)) {
                YRT_CK(cudaEventSynchronize(chunkDone[chunkIdx - 2 * nLanes]));
                if (statusFn) { status.progress = float(pixelBegin - (size_t)(2 * nLanes - 1) * pixelsPerChunk) / float(numPixels); statusFn(&status); }
            }
            if (stopRequested()) { stopped = true; break; }
            const uint32_t np = (uint32_t)std::min<size_t>(pixelsPerChunk, numPixels - pixelBegin);
            NvtxRange nvtxChunk(lane ? "wavefront chunk (lane 1): enqueue" : "wavefront chunk (lane 0): enqueue");
            if (timers) tm.begin(TK_RAYGEN_FILM, ls);
            launch_raygen(fc, cams, wb, (uint32_t)pixelBegin, np, lcStream); launches++;
            if (timers) tm.end(ls);
            int q = 0;
            uint32_t alive = np * (uint32_t)spp;                 // length of the current queue, known on the host
            for (int depth = 0; depth < fc.integ.maxDepth && alive; depth++, q ^= 1) {
                NvtxRange nvtxBounce("bounce: closest / shade / shadow / resolve");
                // bounce queues are re-ordered by origin cell + direction octant (sort.cu); the sorted ids live in queueS
                WavefrontBuffers wq = wb;
                if (depth > 0 && dev->sortRays && alive >= dev->sortMin) {
                    if (timers) tm.begin(TK_SORT, ls);
                    launch_sort_keys(fc, wb, q, alive, lcStream); launches++;
                    sort_queue(wb, q, alive, W[lane]->sortTemp, W[lane]->sortTempBytes, ls); launches += 4;
                    if (timers) tm.end(ls);
                    (q ? wq.queueB : wq.queueA) = wb.queueS;
                }
                if (timers) tm.begin(TK_CLOSEST, ls);
                launch_trace_closest(fc, wq, q, lcTrace); launches++; closestLaunches++;
                if (timers) { tm.end(ls); tm.begin(TK_SHADE, ls); }
                launch_shade(fc, wq, q, (uint32_t)pixelBegin, depth, lcShade); launches++; shadeLaunches++;
                if (timers) tm.end(ls);
                // queue lengths of the next bounce: one small read-back per bounce buys the early exit and the sort size. Below
                // syncMinPaths the round trip costs more than the (short, self-terminating) launches it could save: `alive` then
                // stays an upper bound and the remaining bounces are enqueued without waiting. With two lanes the host never waits
                // inside a chunk (it would stall the other lane's enqueue): every bounce is enqueued, empty ones end at once.
                const bool readCounters = (nLanes == 1 && alive >= dev->syncMinPaths) || (dev->sortRays && alive >= dev->sortMin);
                cudaEvent_t evCnt = nullptr;
                if (readCounters) {
                    YRT_CK(cudaMemcpyAsync(dev->hostCounters, wb.counters, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, ls));
                    evCnt = tm.get(); YRT_CK(cudaEventRecord(evCnt, ls));
                }
                if (fc.scene.numLights > 0) {
                    if (timers) tm.begin(TK_SHADOW, ls);
                    launch_trace_shadow(fc, wq, lcTrace); launches++; shadowLaunches++;
                    if (timers) { tm.end(ls); tm.begin(TK_RESOLVE, ls); }
                }
                else if (timers) tm.begin(TK_RESOLVE, ls);
                launch_resolve(fc, wq, q, lcStream); launches += 2;
                if (timers) tm.end(ls);
                if (readCounters) { YRT_CK(cudaEventSynchronize(evCnt)); alive = dev->hostCounters[q ^ 1]; }
            }
            if (timers) tm.begin(TK_RAYGEN_FILM, ls);
            launch_film(fc, wb, fp, (uint32_t)pixelBegin, np, lcStream); launches++;
            if (timers) tm.end(ls);
            cudaEvent_t e = tm.get(); YRT_CK(cudaEventRecord(e, ls)); chunkDone.push_back(e);
        }
    }
    for (int l = 1; l < nLanes; l++) {                          // the span ends when every lane is done
        cudaEvent_t e = tm.get(); YRT_CK(cudaEventRecord(e, LS[l])); YRT_CK(cudaStreamWaitEvent(st, e, 0));
    }
    YRT_CK(cudaEventRecord(evStop, st));
    hostLap("enqueued");
    uint64_t d2h = 0;
    for (size_t f = 0; f < numFaces; f++) {
        FrameBufferHandle* b = fbs[f];
        if (dev->readback) {
            YRT_CK(cudaMemcpyAsync(b->host[b->cur], b->devPacked, b->bytes(), cudaMemcpyDeviceToHost, st));
            d2h += b->bytes(); b->pendingBuf = -1;
        } else b->pendingBuf = (int)b->cur;
    }
    unsigned long long hstats[8] = {0, 0, 0, 0, 0, 0, 0, 0}, lstats[2][8];
    for (int l = 0; l < nLanes; l++) YRT_CK(cudaMemcpyAsync(lstats[l], W[l]->wb.stats, sizeof(lstats[l]), cudaMemcpyDeviceToHost, st));   // after the lanes joined `st`
    YRT_CK(cudaStreamSynchronize(st));
    YRT_CK(cudaGetLastError());
    for (int l = 0; l < nLanes; l++) for (int k = 0; k < 8; k++) { if (k == 7) hstats[k] |= lstats[l][k]; else hstats[k] += lstats[l][k]; }
    hostLap("synced");

    yrtx_frame_stats& S = dev->stats;
    float ms = 0.f; YRT_CK(cudaEventElapsedTime(&ms, evStart, evStop));
    S.render_ms = ms; S.rays_closest = hstats[0]; S.rays_shadow = hstats[1]; S.node_visits = hstats[2] + hstats[4]; S.tri_tests = hstats[3] + hstats[5];
    S.node_visits_shadow = hstats[4]; S.tri_tests_shadow = hstats[5]; S.path_vertices = R.debug ? 0 : hstats[0]; S.errors = hstats[7];
    S.kernel_launches = launches; S.closest_launches = closestLaunches; S.shadow_launches = shadowLaunches; S.shade_launches = shadeLaunches;
    S.closest_ms = S.shadow_ms = S.shade_ms = S.raygen_film_ms = S.sort_ms = S.resolve_ms = S.miss_ms = 0.0;
    for (const auto& sp : tm.spans) {
        if (!sp.b) continue;
        float t = 0.f; YRT_CK(cudaEventElapsedTime(&t, sp.a, sp.b));
        if (sp.kind == TK_CLOSEST) S.closest_ms += t; else if (sp.kind == TK_SHADOW) S.shadow_ms += t;
        else if (sp.kind == TK_SHADE) S.shade_ms += t; else if (sp.kind == TK_SORT) S.sort_ms += t;
        else if (sp.kind == TK_RESOLVE) S.resolve_ms += t; else if (sp.kind == TK_MISS) S.miss_ms += t; else S.raygen_film_ms += t;
    }
    if (dev->verbose >= 2) {   // stage times in launch order (first 160 spans)
        static const char* names[] = {"raygen/film", "closest", "shade", "shadow", "sort", "resolve", "miss"};
        size_t shown = 0;
        for (const auto& sp : tm.spans) {
            if (!sp.b || shown++ >= 160) break;
            float t = 0.f; cudaEventElapsedTime(&t, sp.a, sp.b);
            printf("  stage %3zu %-12s %9.3f ms\n", shown - 1, names[sp.kind], t);
        }
    }
    S.trace_ms = S.closest_ms + S.shadow_ms;
    S.h2d_bytes = fs.tableUploaded ? fs.tableBytes : 0; S.d2h_bytes = d2h;
    S.num_triangles = fc.scene.numTris; S.num_nodes = fc.scene.numNodes; S.build_ms = sc->buildMs; S.bvh_builds = sc->rebuildCount;
    hostLap("stats");
    S.host_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tHost0).count();
    if (dev->verbose) {   // the reference's line (integratorrenderer.cpp:101-111), fed from CUDA events and device counters
        const double dt = ms * 1e-3;
        printf("render  %.2f fps, %.0f ms, %.3f mrps\n", 1.0 / dt, dt * 1000.0, double(hstats[0] + hstats[1]) / dt * 1e-6);
    }
    if (statusFn) { status.state = 2; status.progress = 1.f; statusFn(&status); }   // updateStatus(Done)
    (void)stopped;
    if (hstats[7] & 1ull) throw std::runtime_error("device_cuda: BVH traversal stack overflow (more than 96 postponed entries on one ray): the frame is incomplete");
}
void render_frame(yrt_device* dev, RendererHandle* rh, CameraHandle* ch, SceneHandle* sc, ToneMapperHandle* th, FrameBufferHandle* fb, int accumulate) {
    render_frames(dev, rh, 1, &ch, sc, th, &fb, accumulate);
}

// ------------------------------------------------------------------------------------------------
// extensions
// ------------------------------------------------------------------------------------------------
void trace_rays(yrt_device* dev, SceneHandle* sc, size_t n, const float* rays, void* hits, int closest, int onDevice, float* ms) {
    if (!sc->committed) throw std::runtime_error("invalid scene value");
    if (ms) *ms = 0.f;
    if (!n) return;
    cudaStream_t st = dev->stream;
    DevBuf<float> dRays, dHits;
    const float* rp = rays; float* hp = (float*)hits;
    if (!onDevice) {
        dRays.alloc(8 * n); dHits.alloc(8 * n);
        YRT_CK(cudaMemcpyAsync(dRays.p, rays, 32 * n, cudaMemcpyHostToDevice, st));
        YRT_CK(cudaMemcpyAsync(dHits.p, hits, 32 * n, cudaMemcpyHostToDevice, st));
        rp = dRays.p; hp = dHits.p;
    }
    dev->wf.ensure(0, 0, 0);
    YRT_CK(cudaMemsetAsync(dev->wf.wb.stats, 0, 8 * sizeof(unsigned long long), st));
    cudaEvent_t a, b; YRT_CK(cudaEventCreate(&a)); YRT_CK(cudaEventCreate(&b));
    LaunchCfg lc{dev->numSMs * 8, 128, st};
    YRT_CK(cudaEventRecord(a, st));
    if (n > 0xfffffff0ull) throw std::runtime_error("device_cuda: yrtxTraceRays is limited to 2^32 - 16 rays per call");
    YRT_CK(cudaMemsetAsync(dev->wf.wb.counters + 6, 0, sizeof(uint32_t), st));
    SceneData sd = sc->data; sd.tuneRefillMin = dev->tuneRefillMin; sd.tuneTriNum = dev->tuneUserTriNum; sd.tuneTriDen = dev->tuneTriDen; sd.tuneSimple = dev->tuneSimple;
    sd.tunePrefetch = dev->tunePrefetch > 0 ? 1 : 0;
    launch_trace_user(sd, rp, hp, n, closest, dev->countStats, dev->wf.wb.stats, dev->wf.wb.counters + 6, lc);
    YRT_CK(cudaEventRecord(b, st));
    if (!onDevice) YRT_CK(cudaMemcpyAsync(hits, dHits.p, 32 * n, cudaMemcpyDeviceToHost, st));
    unsigned long long hstats[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    YRT_CK(cudaMemcpyAsync(hstats, dev->wf.wb.stats, sizeof(hstats), cudaMemcpyDeviceToHost, st));
    YRT_CK(cudaStreamSynchronize(st));
    YRT_CK(cudaGetLastError());
    if (hstats[7] & 1ull) throw std::runtime_error("device_cuda: BVH traversal stack overflow in yrtxTraceRays");
    float t = 0.f; YRT_CK(cudaEventElapsedTime(&t, a, b));
    cudaEventDestroy(a); cudaEventDestroy(b);
    if (ms) *ms = t;
    yrtx_frame_stats& S = dev->stats;
    S.trace_ms = t; S.kernel_launches = 1; S.node_visits = hstats[2]; S.tri_tests = hstats[3];
    if (closest) { S.rays_closest = n; S.rays_shadow = 0; S.closest_ms = t; S.closest_launches = 1; S.shadow_launches = 0; }
    else { S.rays_closest = 0; S.rays_shadow = n; S.shadow_ms = t; S.shadow_launches = 1; S.closest_launches = 0; }
    S.num_triangles = sc->data.numTris; S.num_nodes = sc->data.numNodes;
}

void primary_rays(yrt_device* dev, RendererHandle* rh, CameraHandle* ch, FrameBufferHandle* fb, float* rays, int* sets) {
    if (!rh->inst || !ch->inst) throw std::runtime_error("invalid renderer or camera value");
    RendererObj& R = *rh->inst;
    if (R.debug) throw std::runtime_error("device_cuda: yrtxPrimaryRays needs the pathtracer renderer");
    cudaStream_t st = dev->stream;
    FrameSetup fs = setup_frame(dev, R, nullptr, fb, 0);
    FrameConst& fc = fs.fc;
    FrameCameras cams; memset(&cams, 0, sizeof(cams)); cams.cam[0] = *ch->inst;
    const int spp = fc.integ.spp;
    const size_t numPixels = fs.bufferRows * fb->width;
    uint32_t pixelsPerChunk = (uint32_t)std::max<size_t>(1, std::min<size_t>(numPixels, (1u << 22) / (size_t)spp));
    dev->wf.ensure(pixelsPerChunk * (uint32_t)spp, 1, numPixels);
    const WavefrontBuffers& wb = dev->wf.wb;
    LaunchCfg lc{dev->numSMs * 8, 256, st};
    launch_pixel_sets(fc, wb.pixelSet, fs.sets, lc);
    DevBuf<float> out; out.alloc((size_t)pixelsPerChunk * spp * 8);
    for (size_t pixelBegin = 0; pixelBegin < numPixels; pixelBegin += pixelsPerChunk) {
        const uint32_t np = (uint32_t)std::min<size_t>(pixelsPerChunk, numPixels - pixelBegin);
        launch_raygen(fc, cams, wb, (uint32_t)pixelBegin, np, lc);
        launch_export_primary(fc, wb, (uint32_t)pixelBegin, np, out.p, lc);
        YRT_CK(cudaMemcpyAsync(rays + pixelBegin * spp * 8, out.p, (size_t)np * spp * 32, cudaMemcpyDeviceToHost, st));
        YRT_CK(cudaStreamSynchronize(st));
    }
    if (sets) {
        std::vector<uint8_t> h(numPixels);
        YRT_CK(cudaMemcpyAsync(h.data(), wb.pixelSet, numPixels, cudaMemcpyDeviceToHost, st));
        YRT_CK(cudaStreamSynchronize(st));
        for (size_t i = 0; i < numPixels; i++) sets[i] = h[i];
    }
    YRT_CK(cudaGetLastError());
}

void sample_table(yrt_device* dev, RendererHandle* rh, SceneHandle* sc, int iteration, int* sets, int* spp, int* n1, int* n2, float* table) {
    if (!rh->inst) throw std::runtime_error("invalid renderer value");
    RendererObj R = *rh->inst;                                // by value: querying must not advance the renderer
    if (R.debug) throw std::runtime_error("device_cuda: the debug renderer has no sample table");
    int s, a, b, rec;
    const std::vector<float> tab = build_table(dev, R, (sc && sc->committed) ? sc : nullptr, iteration, s, a, b, rec);
    if (sets) *sets = R.sets; if (spp) *spp = s; if (n1) *n1 = a; if (n2) *n2 = b;
    if (table) {
        const int base = 5 + a + 2 * b;                      // the documented record excludes the light samples
        for (size_t i = 0; i < (size_t)R.sets * s; i++) memcpy(table + i * base, tab.data() + i * rec, base * sizeof(float));
    }
}

}  // namespace yrt
