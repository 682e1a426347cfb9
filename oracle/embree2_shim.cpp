// TEST INFRASTRUCTURE ONLY (oracle). Not part of the product; never linked into
// libyrt_device_cuda.so.
//
// embree2_shim — a CPU restatement of the slice of the Intel Embree v2.15.0 API that the
// reference's live singleray path calls. Intel Embree is an
// un-vendored, binary-only, Windows-only dependency of the reference (Embree.props:10-14;
// the DLL itself is stripped from the mount), so the reference's integrator/shading code
// is linked against this shim instead. "parity unpinned" at this boundary: the reference
// holds no golden vectors for rtcIntersect/rtcOccluded; what IS pinned by reference source
// is restated here and cited:
//   * RTCRay field contract                   3rd party/Embree v2.15.0 x64/include/embree2/rtcore_ray.h:28-55
//   * Ng = (v0-v1) x (v2-v0), unnormalised    devices/device_singleray/lights/trianglelight.h:70-80,
//                                             shapes/trianglemesh_full.cpp:112-113
//   * u weights v1, v weights v2              shapes/trianglemesh_full.cpp:216,229,242
//   * geomID = creation order, primID = tri   shapes/trianglemesh_full.cpp:132-133
//   * filter: candidate written into the ray, rejected by geomID = -1, previous hit restored
//                                             3rd party/Embree v2.15.0 x64/doc/README.md:514-518,2122-2124
//   * scene flags ask for ROBUST traversal    api/scene_flat.h:90-96  -> watertight Pluecker edge test
//
// Arithmetic contract "YRT-PLUECKER-1" (mirrored op-for-op by the CUDA traversal kernel
// in yulio_raytracer_b200/csrc/traverse.cuh so that (t,u,v,geomID,primID) are bit-exact):
//   v_i = p_i - O;  e0 = v2-v0, e1 = v0-v1, e2 = v1-v2;  s0 = v2+v0, s1 = v0+v1, s2 = v1+v2
//   edge(s,e) = fma(c.z,D.z, fma(c.y,D.y, c.x*D.x)),  c = (fma(s.y,e.z,-(s.z*e.y)), fma(s.z,e.x,-(s.x*e.z)), fma(s.x,e.y,-(s.y*e.x)))
//   U,V,W = edge(s0,e0), edge(s1,e1), edge(s2,e2);  reject unless min>=0 or max<=0;  UVW=(U+V)+W != 0
//   Ng = cross(p0-p1, p2-p0) with separate mul/sub (exactly the reference cull filter's arithmetic)
//   den = (Ng.x*D.x + Ng.y*D.y) + Ng.z*D.z != 0;  T = (v0.x*Ng.x + v0.y*Ng.y) + v0.z*Ng.z;  t = T/den
//   accept iff tnear < t and (t < tbest or (t == tbest and (geomID,primID) < best ids));  u = U/UVW, v = V/UVW
// Closest-hit result is therefore independent of traversal order.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "embree2/rtcore.h"
#include "embree2/rtcore_ray.h"

namespace {

struct V3 { float x, y, z; };
static inline V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 add(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }

struct Mesh {
    std::vector<int32_t> idx;     // 3 per triangle (RTCTriangle layout, stride 12)
    std::vector<float> vtx[2];    // x,y,z,pad per vertex (stride 16); [1] = motion end position
    size_t numTris = 0, numVerts = 0, numTimeSteps = 1;
    RTCFilterFunc isectFilter = nullptr, occlFilter = nullptr;
    void* userData = nullptr;
};

struct TriRef { uint32_t geom, prim; };

struct Node {                      // BVH2, 32 bytes
    float lo[3]; uint32_t left;    // inner: index of left child (right = left+1); leaf: first TriRef
    float hi[3]; uint32_t count;   // 0 = inner node, else number of triangles
};

struct Scene {
    std::vector<Mesh*> meshes;
    std::vector<Node> nodes;
    std::vector<TriRef> refs;
    std::vector<float> tri;        // 9 floats per ref (time step 0), gathered for locality
    std::vector<float> triD;       // motion blur: 9 floats per ref, vertex(time step 1) - vertex(time step 0); empty if no mesh moves
    std::vector<uint8_t> moving;   // per ref: its mesh has two time steps
    bool committed = false;
    ~Scene() { for (auto* m : meshes) delete m; }
};

// ------------------------------------------------------------------ build (binned SAH)
struct BuildPrim { float lo[3], hi[3], c[3]; TriRef ref; };

static inline float halfArea(const float* lo, const float* hi) {
    float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    return dx * dy + dy * dz + dz * dx;
}

struct Builder {
    std::vector<BuildPrim>& prims;
    std::vector<Node>& nodes;
    explicit Builder(std::vector<BuildPrim>& p, std::vector<Node>& n) : prims(p), nodes(n) {}

    void bounds(size_t b, size_t e, float* lo, float* hi, float* clo, float* chi) {
        for (int k = 0; k < 3; k++) { lo[k] = clo[k] = INFINITY; hi[k] = chi[k] = -INFINITY; }
        for (size_t i = b; i < e; i++) for (int k = 0; k < 3; k++) {
            lo[k] = std::min(lo[k], prims[i].lo[k]); hi[k] = std::max(hi[k], prims[i].hi[k]);
            clo[k] = std::min(clo[k], prims[i].c[k]); chi[k] = std::max(chi[k], prims[i].c[k]);
        }
    }

    void build(uint32_t nodeIdx, size_t b, size_t e) {
        float lo[3], hi[3], clo[3], chi[3];
        bounds(b, e, lo, hi, clo, chi);
        Node nd;
        for (int k = 0; k < 3; k++) { nd.lo[k] = lo[k]; nd.hi[k] = hi[k]; }
        const size_t n = e - b;
        auto makeLeaf = [&]() { nd.left = (uint32_t)b; nd.count = (uint32_t)n; nodes[nodeIdx] = nd; };
        if (n <= 2) { makeLeaf(); return; }

        constexpr int NB = 16;
        int bestAxis = -1, bestBin = -1; float bestCost = INFINITY;
        for (int ax = 0; ax < 3; ax++) {
            const float ext = chi[ax] - clo[ax];
            if (!(ext > 0.f)) continue;
            float blo[NB][3], bhi[NB][3]; size_t cnt[NB];
            for (int i = 0; i < NB; i++) { cnt[i] = 0; for (int k = 0; k < 3; k++) { blo[i][k] = INFINITY; bhi[i][k] = -INFINITY; } }
            const float scale = NB * (1.f - 1e-6f) / ext;
            for (size_t i = b; i < e; i++) {
                int bi = std::min(NB - 1, std::max(0, (int)((prims[i].c[ax] - clo[ax]) * scale)));
                cnt[bi]++;
                for (int k = 0; k < 3; k++) { blo[bi][k] = std::min(blo[bi][k], prims[i].lo[k]); bhi[bi][k] = std::max(bhi[bi][k], prims[i].hi[k]); }
            }
            float rA[NB]; size_t rN[NB];
            float alo[3] = {INFINITY, INFINITY, INFINITY}, ahi[3] = {-INFINITY, -INFINITY, -INFINITY}; size_t an = 0;
            for (int i = NB - 1; i > 0; i--) {
                for (int k = 0; k < 3; k++) { alo[k] = std::min(alo[k], blo[i][k]); ahi[k] = std::max(ahi[k], bhi[i][k]); }
                an += cnt[i]; rA[i] = an ? halfArea(alo, ahi) : 0.f; rN[i] = an;
            }
            for (int k = 0; k < 3; k++) { alo[k] = INFINITY; ahi[k] = -INFINITY; }
            an = 0;
            for (int i = 0; i < NB - 1; i++) {
                for (int k = 0; k < 3; k++) { alo[k] = std::min(alo[k], blo[i][k]); ahi[k] = std::max(ahi[k], bhi[i][k]); }
                an += cnt[i];
                if (an == 0 || rN[i + 1] == 0) continue;
                float cost = halfArea(alo, ahi) * (float)an + rA[i + 1] * (float)rN[i + 1];
                if (cost < bestCost) { bestCost = cost; bestAxis = ax; bestBin = i; }
            }
        }
        size_t mid;
        if (bestAxis < 0) {
            if (n <= 4) { makeLeaf(); return; }
            mid = b + n / 2;                                   // all centroids coincide: split by count
        } else {
            const float leafCost = halfArea(lo, hi) * (float)n;
            if (n <= 4 && leafCost <= bestCost + halfArea(lo, hi)) { makeLeaf(); return; }
            const float ext = chi[bestAxis] - clo[bestAxis];
            const float scale = NB * (1.f - 1e-6f) / ext;
            const float c0 = clo[bestAxis];
            const int ax = bestAxis, sb = bestBin;
            auto it = std::partition(prims.begin() + b, prims.begin() + e, [&](const BuildPrim& p) {
                int bi = std::min(NB - 1, std::max(0, (int)((p.c[ax] - c0) * scale)));
                return bi <= sb;
            });
            mid = (size_t)(it - prims.begin());
            if (mid == b || mid == e) mid = b + n / 2;
        }
        const uint32_t left = (uint32_t)nodes.size();
        nodes.push_back(Node()); nodes.push_back(Node());
        nd.left = left; nd.count = 0; nodes[nodeIdx] = nd;
        build(left, b, mid);
        build(left + 1, mid, e);
    }
};

static void commitScene(Scene* s) {
    std::vector<BuildPrim> prims;
    size_t total = 0;
    for (auto* m : s->meshes) total += m->numTris;
    prims.reserve(total);
    for (size_t g = 0; g < s->meshes.size(); g++) {
        Mesh* m = s->meshes[g];
        for (size_t t = 0; t < m->numTris; t++) {
            BuildPrim p; p.ref = {(uint32_t)g, (uint32_t)t};
            for (int k = 0; k < 3; k++) { p.lo[k] = INFINITY; p.hi[k] = -INFINITY; }
            bool ok = true;
            for (int c = 0; c < 3; c++) {
                const int32_t vi = m->idx[3 * t + c];
                if (vi < 0 || (size_t)vi >= m->numVerts) { ok = false; break; }
                for (size_t ts = 0; ts < m->numTimeSteps; ts++) for (int k = 0; k < 3; k++) {
                    const float x = m->vtx[ts][4 * (size_t)vi + k];
                    if (!std::isfinite(x)) ok = false;
                    p.lo[k] = std::min(p.lo[k], x); p.hi[k] = std::max(p.hi[k], x);
                }
            }
            if (!ok) continue;                                  // invalid triangles are never hit
            for (int k = 0; k < 3; k++) p.c[k] = 0.5f * p.lo[k] + 0.5f * p.hi[k];
            prims.push_back(p);
        }
    }
    s->nodes.clear(); s->refs.clear(); s->tri.clear();
    s->nodes.reserve(prims.size() * 2 + 2);
    s->nodes.push_back(Node());
    if (prims.empty()) {
        Node nd; for (int k = 0; k < 3; k++) { nd.lo[k] = INFINITY; nd.hi[k] = -INFINITY; } nd.left = 0; nd.count = 0;
        s->nodes[0] = nd; s->nodes[0].count = 0; s->nodes[0].left = 0xffffffffu;   // empty marker
    } else {
        Builder b(prims, s->nodes);
        b.build(0, 0, prims.size());
    }
    s->refs.resize(prims.size()); s->tri.resize(prims.size() * 9);
    bool anyMotion = false;
    for (auto* m : s->meshes) anyMotion |= m->numTimeSteps > 1;
    s->triD.clear(); s->moving.clear();
    if (anyMotion) { s->triD.assign(prims.size() * 9, 0.f); s->moving.assign(prims.size(), 0); }
    for (size_t i = 0; i < prims.size(); i++) {
        s->refs[i] = prims[i].ref;
        Mesh* m = s->meshes[prims[i].ref.geom];
        for (int c = 0; c < 3; c++) {
            const int32_t vi = m->idx[3 * (size_t)prims[i].ref.prim + c];
            for (int k = 0; k < 3; k++) s->tri[9 * i + 3 * c + k] = m->vtx[0][4 * (size_t)vi + k];
            if (m->numTimeSteps > 1) {                          // contract: vertex(time) = v0 + time * (v1 - v0), mul then add
                s->moving[i] = 1;
                for (int k = 0; k < 3; k++) s->triD[9 * i + 3 * c + k] = m->vtx[1][4 * (size_t)vi + k] - m->vtx[0][4 * (size_t)vi + k];
            }
        }
    }
    s->committed = true;
}

// ------------------------------------------------------------------ YRT-PLUECKER-1
static inline float edgeFn(V3 s, V3 e, V3 D) {
    const float cx = __builtin_fmaf(s.y, e.z, -(s.z * e.y));
    const float cy = __builtin_fmaf(s.z, e.x, -(s.x * e.z));
    const float cz = __builtin_fmaf(s.x, e.y, -(s.y * e.x));
    return __builtin_fmaf(cz, D.z, __builtin_fmaf(cy, D.y, cx * D.x));
}

struct Cand { float t, u, v; V3 Ng; };

static inline bool triTest(V3 O, V3 D, V3 p0, V3 p1, V3 p2, Cand& c) {
    const V3 v0 = sub(p0, O), v1 = sub(p1, O), v2 = sub(p2, O);
    const V3 e0 = sub(v2, v0), e1 = sub(v0, v1), e2 = sub(v1, v2);
    const float U = edgeFn(add(v2, v0), e0, D);
    const float V = edgeFn(add(v0, v1), e1, D);
    const float W = edgeFn(add(v1, v2), e2, D);
    const float mn = std::fmin(std::fmin(U, V), W), mx = std::fmax(std::fmax(U, V), W);
    if (!(mn >= 0.f || mx <= 0.f)) return false;
    const float UVW = (U + V) + W;
    if (UVW == 0.f) return false;
    const V3 a = sub(p0, p1), b = sub(p2, p0);
    const V3 Ng = {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
    const float den = (Ng.x * D.x + Ng.y * D.y) + Ng.z * D.z;
    if (den == 0.f) return false;
    const float T = (v0.x * Ng.x + v0.y * Ng.y) + v0.z * Ng.z;
    c.t = T / den; c.u = U / UVW; c.v = V / UVW; c.Ng = Ng;
    return true;
}

static inline bool boxTest(const Node& n, V3 O, V3 rd, float tnear, float tfar, float& tEntry) {
    float t0 = (n.lo[0] - O.x) * rd.x, t1 = (n.hi[0] - O.x) * rd.x;
    float tN = std::fmax(std::fmin(t0, t1), tnear), tF = std::fmin(std::fmax(t0, t1), tfar);
    t0 = (n.lo[1] - O.y) * rd.y; t1 = (n.hi[1] - O.y) * rd.y;
    tN = std::fmax(std::fmin(t0, t1), tN); tF = std::fmin(std::fmax(t0, t1), tF);
    t0 = (n.lo[2] - O.z) * rd.z; t1 = (n.hi[2] - O.z) * rd.z;
    tN = std::fmax(std::fmin(t0, t1), tN); tF = std::fmin(std::fmax(t0, t1), tF);
    tEntry = tN;
    return tN * 0.9999995f <= tF * 1.0000005f;                 // conservative (robust-mode style padding)
}

// ------------------------------------------------------------------ optional ray log (tests)
struct TraceRec {          // 64 bytes; layout mirrored in tests/oracle_api.py
    float org[3], tnear, dir[3], tfar;     // as passed in
    float t, u, v; int32_t geomID, primID; // result (closest) / geomID only (any-hit)
    int32_t kind;                          // 0 = rtcIntersect, 1 = rtcOccluded
    float Ng_x, Ng_y;
};
std::mutex g_traceMutex;
TraceRec* g_trace = nullptr; size_t g_traceCap = 0; std::atomic<size_t> g_traceN{0};
std::atomic<uint64_t> g_nodeVisits{0}, g_triTests{0}; bool g_countStats = false;

template <bool ANY>
static void traverse(Scene* s, RTCRay& ray) {
    if (!s->committed || s->nodes[0].left == 0xffffffffu) return;
    if (!(ray.tnear <= ray.tfar)) return;                      // empty or NaN interval (pin P2): nothing can be accepted below
    const V3 O = {ray.org[0], ray.org[1], ray.org[2]}, D = {ray.dir[0], ray.dir[1], ray.dir[2]};
    const V3 rd = {1.0f / D.x, 1.0f / D.y, 1.0f / D.z};
    const float tnear = ray.tnear;
    float tbest = ray.tfar;
    bool have = false; uint32_t bg = 0, bp = 0;
    uint32_t stack[128]; int sp = 0; stack[sp++] = 0;
    uint64_t nv = 0, nt = 0;
    while (sp) {
        const Node& n = s->nodes[stack[--sp]];
        float te; nv++;
        if (!boxTest(n, O, rd, tnear, tbest, te)) continue;
        if (n.count == 0) {
            const Node& l = s->nodes[n.left]; const Node& r = s->nodes[n.left + 1];
            // near child first (by the sign of the direction along the largest split extent)
            float cl = 0, cr = 0;
            for (int k = 0; k < 3; k++) { const float d = (&D.x)[k]; cl += d * (l.lo[k] + l.hi[k]); cr += d * (r.lo[k] + r.hi[k]); }
            if (cl <= cr) { stack[sp++] = n.left + 1; stack[sp++] = n.left; }
            else          { stack[sp++] = n.left; stack[sp++] = n.left + 1; }
            continue;
        }
        for (uint32_t i = n.left; i < n.left + n.count; i++) {
            const float* q = &s->tri[9 * (size_t)i];
            float moved[9];
            if (!s->moving.empty() && s->moving[i]) {           // linear vertex motion over the shutter interval (rtcore_ray.h: RTCRay::time)
                const float* d = &s->triD[9 * (size_t)i];
                for (int k = 0; k < 9; k++) { const float step = ray.time * d[k]; moved[k] = q[k] + step; }
                q = moved;
            }
            Cand c; nt++;
            if (!triTest(O, D, {q[0], q[1], q[2]}, {q[3], q[4], q[5]}, {q[6], q[7], q[8]}, c)) continue;
            if (!(c.t > tnear)) continue;
            const TriRef ref = s->refs[i];
            bool closer = c.t < tbest;
            if (!ANY && !closer && have && c.t == tbest)
                closer = (ref.geom < bg) || (ref.geom == bg && ref.prim < bp);
            if (!closer) continue;
            Mesh* m = s->meshes[ref.geom];
            RTCFilterFunc f = ANY ? m->occlFilter : m->isectFilter;
            if (f) {
                RTCRay saved = ray;                             // EMBREE_INTERSECTION_FILTER_RESTORE
                ray.tfar = c.t; ray.u = c.u; ray.v = c.v; ray.geomID = ref.geom; ray.primID = ref.prim;
                ray.Ng[0] = c.Ng.x; ray.Ng[1] = c.Ng.y; ray.Ng[2] = c.Ng.z;
                f(m->userData, ray);
                const bool rejected = ray.geomID == RTC_INVALID_GEOMETRY_ID;
                ray = saved;
                if (rejected) continue;
            }
            if (ANY) { ray.geomID = 0; if (g_countStats) { g_nodeVisits += nv; g_triTests += nt; } return; }
            have = true; bg = ref.geom; bp = ref.prim; tbest = c.t;
            ray.tfar = c.t; ray.u = c.u; ray.v = c.v; ray.geomID = ref.geom; ray.primID = ref.prim;
            ray.Ng[0] = c.Ng.x; ray.Ng[1] = c.Ng.y; ray.Ng[2] = c.Ng.z;
        }
    }
    if (g_countStats) { g_nodeVisits += nv; g_triTests += nt; }
}

static void logRay(const RTCRay& in, const RTCRay& out, int kind) {
    std::lock_guard<std::mutex> lock(g_traceMutex);
    const size_t i = g_traceN.load();
    if (!g_trace || i >= g_traceCap) { g_traceN = i + 1; return; }
    TraceRec& r = g_trace[i];
    for (int k = 0; k < 3; k++) { r.org[k] = in.org[k]; r.dir[k] = in.dir[k]; }
    r.tnear = in.tnear; r.tfar = in.tfar;
    r.t = out.tfar; r.u = out.u; r.v = out.v; r.geomID = (int32_t)out.geomID; r.primID = (int32_t)out.primID;
    r.kind = kind; r.Ng_x = out.Ng[0]; r.Ng_y = out.Ng[1];
    g_traceN = i + 1;
}

}  // namespace

// ---------------------------------------------------------------------- exported C API
extern "C" {

void rtcInit(const char*) {}
void rtcExit() {}
void rtcDebug() {}

RTCScene rtcNewScene(RTCSceneFlags, RTCAlgorithmFlags) { return (RTCScene) new Scene(); }
void rtcDeleteScene(RTCScene scene) { delete (Scene*)scene; }
void rtcCommit(RTCScene scene) { commitScene((Scene*)scene); }

unsigned rtcNewTriangleMesh(RTCScene scene, RTCGeometryFlags, size_t numTriangles, size_t numVertices, size_t numTimeSteps) {
    Scene* s = (Scene*)scene;
    Mesh* m = new Mesh();
    m->numTris = numTriangles; m->numVerts = numVertices; m->numTimeSteps = numTimeSteps < 1 ? 1 : (numTimeSteps > 2 ? 2 : numTimeSteps);
    m->idx.assign(3 * numTriangles, 0);
    for (size_t t = 0; t < m->numTimeSteps; t++) m->vtx[t].assign(4 * numVertices + 4, 0.f);
    s->meshes.push_back(m);
    return (unsigned)(s->meshes.size() - 1);
}

void* rtcMapBuffer(RTCScene scene, unsigned geomID, RTCBufferType type) {
    Mesh* m = ((Scene*)scene)->meshes.at(geomID);
    if (type == RTC_INDEX_BUFFER) return m->idx.data();
    if (type == RTC_VERTEX_BUFFER0) return m->vtx[0].data();
    if (type == RTC_VERTEX_BUFFER1 && m->numTimeSteps > 1) return m->vtx[1].data();
    return nullptr;
}
void rtcUnmapBuffer(RTCScene, unsigned, RTCBufferType) {}

void rtcSetIntersectionFilterFunction(RTCScene scene, unsigned geomID, RTCFilterFunc func) { ((Scene*)scene)->meshes.at(geomID)->isectFilter = func; }
void rtcSetOcclusionFilterFunction(RTCScene scene, unsigned geomID, RTCFilterFunc func) { ((Scene*)scene)->meshes.at(geomID)->occlFilter = func; }
void rtcSetUserData(RTCScene scene, unsigned geomID, void* ptr) { ((Scene*)scene)->meshes.at(geomID)->userData = ptr; }

void rtcIntersect(RTCScene scene, RTCRay& ray) {
    if (g_trace) { RTCRay in = ray; traverse<false>((Scene*)scene, ray); logRay(in, ray, 0); }
    else traverse<false>((Scene*)scene, ray);
}
void rtcOccluded(RTCScene scene, RTCRay& ray) {
    if (g_trace) { RTCRay in = ray; traverse<true>((Scene*)scene, ray); logRay(in, ray, 1); }
    else traverse<true>((Scene*)scene, ray);
}

// ---- test hooks (not Embree API) ----
void yrt_shim_trace_begin(void* buffer, size_t capacityRecords) {
    std::lock_guard<std::mutex> lock(g_traceMutex);
    g_trace = (TraceRec*)buffer; g_traceCap = capacityRecords; g_traceN = 0;
}
size_t yrt_shim_trace_end() {
    std::lock_guard<std::mutex> lock(g_traceMutex);
    g_trace = nullptr; g_traceCap = 0; return g_traceN.load();
}
void yrt_shim_stats(int enable, uint64_t* nodeVisits, uint64_t* triTests) {
    if (nodeVisits) *nodeVisits = g_nodeVisits.load();
    if (triTests) *triTests = g_triTests.load();
    g_countStats = enable != 0; if (enable == 2) { g_nodeVisits = 0; g_triTests = 0; }
}

}  // extern "C"
