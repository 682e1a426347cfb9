// common.cuh — shared host/device types and the "strict float" vector math of device_cuda.
//
// Numeric contract: the whole library is compiled with --fmad=false (device) and
// -ffp-contract=off (host), so every a*b+c below is two correctly rounded IEEE operations, in the
// order written — which is the order the reference's SSE code performs them
// (common/math/vector3f_sse.h:155-238, color_sse.h, linearspace3.h, affinespace.h). Fused
// multiply-adds appear only where they are written explicitly (traverse.cuh).  With the oracle
// pinned to exact 1/x and 1/sqrt(x) (oracle/make_overlay.py P3) this makes most of the path
// bit-reproducible against the CPU reference; the libm transcendentals (sinf, cosf, acosf, powf,
// expf, logf, atanf, tanf) are the remaining source of last-bit differences.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define YRT_HD __host__ __device__ __forceinline__
#define YRT_D __device__ __forceinline__
#else
#define YRT_HD inline
#define YRT_D inline
#endif

namespace yrt {

struct V3 {
    float x, y, z;
    YRT_HD V3() {}
    YRT_HD V3(float a) : x(a), y(a), z(a) {}
    YRT_HD V3(float a, float b, float c) : x(a), y(b), z(c) {}
};
typedef V3 Col;  // Color (r,g,b) — same lane arithmetic as the reference's SSE Color

YRT_HD V3 operator+(V3 a, V3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
YRT_HD V3 operator-(V3 a, V3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
YRT_HD V3 operator*(V3 a, V3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
YRT_HD V3 operator*(V3 a, float b) { return V3(a.x * b, a.y * b, a.z * b); }
YRT_HD V3 operator*(float a, V3 b) { return V3(a * b.x, a * b.y, a * b.z); }
YRT_HD V3 operator/(V3 a, float b) { return V3(a.x / b, a.y / b, a.z / b); }  // _mm_div_ps (vector3f_sse.h:163)
YRT_HD V3 operator-(V3 a) { return V3(-a.x, -a.y, -a.z); }
YRT_HD V3& operator+=(V3& a, V3 b) { a = a + b; return a; }
YRT_HD V3& operator*=(V3& a, V3 b) { a = a * b; return a; }
YRT_HD V3& operator*=(V3& a, float b) { a = a * b; return a; }
YRT_HD bool operator==(V3 a, V3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
YRT_HD bool operator!=(V3 a, V3 b) { return a.x != b.x || a.y != b.y || a.z != b.z; }

// dpps 0x7F: (x*x' + y*y') + (z*z' + 0)   (vector3f_sse.h:207-209)
YRT_HD float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
// shuffle(a*b.yzx - a.yzx*b)   (vector3f_sse.h:226-233)
YRT_HD V3 cross(V3 a, V3 b) { return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
// Device code calls the IEEE division / square root / libm expansions through shared out-of-line copies: the shading
// kernel is bound by instruction fetch (ncu: stall_no_instruction ~ 1/3 of samples, ~55 KB of hot SASS), and every inlined
// 1/x, 1/sqrt(x), sinf, cosf, powf costs 10..100 instructions per call site. Same arithmetic, same results.
#if defined(__CUDA_ARCH__) && !defined(YRT_INLINE_MATH)
static __device__ __noinline__ float yrt_rcp_ool(float x) { return 1.0f / x; }
static __device__ __noinline__ float yrt_rsqrt_ool(float x) { return 1.0f / sqrtf(x); }
static __device__ __noinline__ float yrt_sin_ool(float x) { return sinf(x); }
static __device__ __noinline__ float yrt_cos_ool(float x) { return cosf(x); }
// x^y for the BRDF exponents (x in [0,1] or a transmission colour, y > 0): exp2(y * log2 x) on the 1-2 ulp CUDA exp2f / log2f — relative
// error ~1e-7 * (1 + |y log2 x|) against powf's ~1e-7, at about a seventh of powf's instructions (ncu r1: powf was 13.5 % of the
// shading kernel's executed instructions at 10 of 32 lanes). CUDA's powf differs from glibc's in the last bits anyway (common.cuh
// header); the image parity bound of the tests covers both. -DYRT_EXACT_POW restores powf.
#if defined(YRT_EXACT_POW)
static __device__ __noinline__ float yrt_pow_ool(float x, float y) { return powf(x, y); }
#else
static __device__ __noinline__ float yrt_pow_ool(float x, float y) { return y == 0.0f ? 1.0f : exp2f(y * log2f(x)); }
#endif
static __device__ __noinline__ void yrt_sincos_ool(float x, float* s, float* c) { sincosf(x, s, c); }
#define YRT_RCP(x) yrt_rcp_ool(x)
#define YRT_RSQRT(x) yrt_rsqrt_ool(x)
#define YRT_SINF(x) yrt_sin_ool(x)
#define YRT_COSF(x) yrt_cos_ool(x)
#define YRT_POWF(x, y) yrt_pow_ool(x, y)
#define YRT_SINCOS(x, s, c) yrt_sincos_ool(x, &(s), &(c))
#else
#define YRT_RCP(x) (1.0f / (x))
#define YRT_RSQRT(x) (1.0f / sqrtf(x))
#define YRT_SINF(x) sinf(x)
#define YRT_COSF(x) cosf(x)
#define YRT_POWF(x, y) powf(x, y)
#define YRT_SINCOS(x, s, c) do { (s) = sinf(x); (c) = cosf(x); } while (0)
#endif
YRT_HD float rcpf(float x) { return YRT_RCP(x); }             // pinned P3 (math.h:65)
YRT_HD float rsqrtf_exact(float x) { return YRT_RSQRT(x); }  // pinned P3 (math.h:69)
YRT_HD V3 normalize(V3 a) { return a * rsqrtf_exact(dot(a, a)); }  // vector3f_sse.h:238
YRT_HD float length(V3 a) { return sqrtf(dot(a, a)); }
YRT_HD float reduce_max(V3 a) { float m = a.x < a.y ? a.y : a.x; return m < a.z ? a.z : m; }  // max(max(x,y),z), math.h:117
YRT_HD float reduce_add(V3 a) { return a.x + a.y + a.z; }
YRT_HD V3 vabs(V3 a) { return V3(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }
// the reference's min/max/clamp templates (math.h:107-120): a<b ? a : b  — NaN behaviour matters
YRT_HD float rmin(float a, float b) { return a < b ? a : b; }
YRT_HD float rmax(float a, float b) { return a < b ? b : a; }
YRT_HD float rclamp(float x, float lo = 0.f, float hi = 1.f) { return rmax(lo, rmin(x, hi)); }
YRT_HD int iclamp(int x, int lo, int hi) { int m = x < hi ? x : hi; return lo < m ? m : lo; }
YRT_HD float signf_(float x) { return x < 0 ? -1.0f : 1.0f; }
YRT_HD float deg2rad(float x) { return x * 1.74532925199432957692e-2f; }
YRT_HD float rad2deg(float x) { return x * 5.72957795130823208768e1f; }
YRT_HD float smoothstepf(float e0, float e1, float x) {  // math.h:127-133
    x = rclamp((x - e0) / (e1 - e0), 0.f, 1.f);
    return x * x * (3 - 2 * x);
}
#define YRT_PI 3.14159265358979323846f
#define YRT_TWO_PI 6.283185307179586232f
#define YRT_ONE_OVER_PI 0.31830988618379069122f
#define YRT_ONE_OVER_TWO_PI 0.15915494309189534561f
#define YRT_ULP 1.1920928955078125e-07f  // std::numeric_limits<float>::epsilon()

struct Lin3 { V3 vx, vy, vz; };           // column vectors (linearspace3.h)
struct Aff3 { Lin3 l; V3 p; };            // affinespace.h
YRT_HD V3 xfmVector(const Lin3& s, V3 a) { return a.x * s.vx + a.y * s.vy + a.z * s.vz; }  // linearspace3.h:144-145
YRT_HD V3 xfmPoint(const Aff3& m, V3 p) { return xfmVector(m.l, p) + m.p; }
YRT_HD V3 xfmVector(const Aff3& m, V3 v) { return xfmVector(m.l, v); }
YRT_HD Lin3 mul(const Lin3& a, const Lin3& b) { Lin3 r; r.vx = xfmVector(a, b.vx); r.vy = xfmVector(a, b.vy); r.vz = xfmVector(a, b.vz); return r; }
YRT_HD Aff3 mul(const Aff3& a, const Aff3& b) { Aff3 r; r.l = mul(a.l, b.l); r.p = xfmVector(a.l, b.p) + a.p; return r; }  // affinespace.h:98
YRT_HD Lin3 lin3_identity() { Lin3 r; r.vx = V3(1, 0, 0); r.vy = V3(0, 1, 0); r.vz = V3(0, 0, 1); return r; }
YRT_HD Aff3 aff3_translate(V3 p) { Aff3 r; r.l = lin3_identity(); r.p = p; return r; }
// LinearSpace3::rotate(u, r)  (linearspace3.h:95-101): constructor takes ROW-major scalars
YRT_HD Lin3 lin3_rotate(V3 _u, float r) {
    V3 u = normalize(_u);
    float s = sinf(r), c = cosf(r);
    const float m00 = u.x * u.x + (1 - u.x * u.x) * c, m01 = u.x * u.y * (1 - c) - u.z * s, m02 = u.x * u.z * (1 - c) + u.y * s;
    const float m10 = u.x * u.y * (1 - c) + u.z * s, m11 = u.y * u.y + (1 - u.y * u.y) * c, m12 = u.y * u.z * (1 - c) - u.x * s;
    const float m20 = u.x * u.z * (1 - c) - u.y * s, m21 = u.y * u.z * (1 - c) + u.x * s, m22 = u.z * u.z + (1 - u.z * u.z) * c;
    Lin3 R; R.vx = V3(m00, m10, m20); R.vy = V3(m01, m11, m21); R.vz = V3(m02, m12, m22);
    return R;
}
// AffineSpace::rotate(p, u, r) = translate(+p) * rotate(u, r) * translate(-p)   (affinespace.h:70)
YRT_HD Aff3 aff3_rotate_about(V3 p, V3 u, float r) {
    Aff3 R; R.l = lin3_rotate(u, r); R.p = V3(0.f);
    return mul(mul(aff3_translate(p), R), aff3_translate(-p));
}
// frame(N)  (linearspace3.h:118-124)
YRT_HD Lin3 frame(V3 N) {
    V3 dx0 = cross(V3(1, 0, 0), N), dx1 = cross(V3(0, 1, 0), N);
    V3 dx = normalize(dot(dx0, dx0) > dot(dx1, dx1) ? dx0 : dx1);
    V3 dy = normalize(cross(N, dx));
    Lin3 f; f.vx = dx; f.vy = dy; f.vz = N;
    return f;
}

// ---- pin P1 (oracle/yrt_oracle_pins.h holds the CPU twin) -------------------------------------
YRT_HD uint32_t fmix32(uint32_t h) { h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16; return h; }
YRT_HD uint32_t hash4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    uint32_t h = fmix32(a + 0x9E3779B9u);
    h = fmix32(h ^ (b + 0x7F4A7C15u));
    h = fmix32(h ^ (c + 0x94D049BBu));
    h = fmix32(h ^ (d + 0xBF58476Du));
    return h;
}
YRT_HD float hash_unit(uint32_t h) { return (float)(h >> 8) * (1.0f / 16777216.0f); }

// ---- device-side scene description --------------------------------------------------------
enum CameraType { CAM_PINHOLE = 0, CAM_STEREO = 1, CAM_DOF = 2 };
struct CameraData {
    int type; int cubeFaceIndex; int toeIn; int pad0;
    Aff3 p2w[6];                 // pixel2world (pinhole: [0] only)
    V3 origin, up, xyzStraight;
    float eyeSeparation, rcpZeroParallax, falloffAngle;
    float lensRadius, focalDistance;   // depth of field (cameras/depthoffieldcamera.h)
    Aff3 local2world;
};

#define YRT_NO_ATTR 0xffffffffu
enum MeshType { MESH_FULL = 0, MESH_NORMALS = 1, MESH_TRIANGLE = 2 };
struct GeomRec {                 // one per geomID (= one committed shape primitive)
    int type, material, areaLight, cull;
    int illumMask, shadowMask;
    uint32_t vtxBase, idxBase;   // into positions and the int4 index array
    uint32_t nrmBase, uvBase;    // into normals / uvs, YRT_NO_ATTR when the mesh has none
    V3 triNg;                    // MESH_TRIANGLE: normalize(cross(v2-v0, v1-v0))  (shapes/triangle.h:43)
    int shadeClass;              // 1..14: hits whose shading follows the same code path (material kind, textured or not, lobe set)
    uint32_t motBase;            // into SceneData::motions (per-vertex motion vectors, "motions" / sphere "dPdt"), YRT_NO_ATTR when static
    uint32_t tanXBase, tanYBase; // into SceneData::tangents (per-vertex tangent_x / tangent_y arrays), YRT_NO_ATTR when the mesh has none
};

enum MaterialType { MAT_NONE = 0, MAT_MATTE, MAT_OBJ, MAT_UBER, MAT_MATTE_TEXTURED, MAT_DIELECTRIC, MAT_THIN_DIELECTRIC, MAT_MIRROR,
                    MAT_PLASTIC, MAT_METAL, MAT_BRUSHED_METAL, MAT_METALLIC_PAINT, MAT_VELVET };   // from MAT_PLASTIC on: the EXT shading kernel
struct MaterialRec {
    int type;
    int tex[5];                  // texture table indices or -1 (Obj: map_d, map_Kd, map_Ks, map_Ns, map_Bump; others: [0] = Kd)
    float s0x, s0y, dsx, dsy;
    Col c0, c1, c2;              // Matte/Mirror: reflectance; Obj: Kd, Ks; Uber: diffuse; ThinDielectric: transmission; Metal: reflectance, eta, k
    float f[6];                  // Obj: d, Ns; Uber: eta, roughness, reflectivity, rcpRoughness; Thin: eta, thickness, transparency
    // Dielectric media (materials/dielectric.h:31-45)
    Col tOutside, tInside; float etaOutside, etaInside;
    int isMediaInterface;
};

enum TexFormat { TEX_RGB8 = 0, TEX_RGBA8 = 1, TEX_RGB_F32 = 2, TEX_RGBA_F32 = 3 };
struct TextureRec {
    const void* data; int width, height; int format; int bilinear; int invert; int pad;
};

enum LightType { LIGHT_AMBIENT = 0, LIGHT_TRIANGLE, LIGHT_POINT, LIGHT_SPOT, LIGHT_DIRECTIONAL, LIGHT_DISTANT, LIGHT_HDRI };
struct LightRec {
    int type, illumMask, shadowMask, precomputedId;   // precomputedId >= 0: light samples come from the table
    Col L;                       // L / I / E
    V3 v0, v1, v2, Ng;           // triangle (Ng = cross(v0-v1, v2-v0)); point/spot: v0 = P; dir lights: v0 = _wo; spot: v1 = _D
    float a, b;                  // spot: cosAngleMin, cosAngleMax; distant: halfAngle, cosHalfAngle
    Aff3 world2local;            // hdri
    int image;                   // hdri: texture table index of the lat-long image (nearest, no invert)
    int isEnv;
};

struct SceneData {
    int tuneRefillMin, tuneTriNum, tuneTriDen, tuneSimple, tunePrefetch;
    V3 bboxLo, bboxRcpExtent;    // scene bounds (ray-sort keys): cell = (P - bboxLo) * bboxRcpExtent in [0,1]^3   // traversal scheduling knobs (bvh.cuh: TraceTune)
    // acceleration structure
    const void* nodes;           // BVH8 nodes, 80 B each (bvh.cuh)
    const float4* tris;          // 3 x float4 per triangle in leaf order (p0|geomID, p1|primID, p2|cull)
    const float4* triShade;      // 5 x float4 per triangle in leaf order: what postIntersect needs (bvh_build.cu: write_triangle)
    const float4* triMotion;     // motion blur: 3 x float4 per triangle in leaf order, vertex(t = 1) - vertex(t = 0); NULL when nothing moves
    int hasMotion;               // some mesh carries motion vectors: rays carry their time, moving triangles are interpolated per ray
    uint32_t numNodes, numTris;
    // shading data
    const GeomRec* geoms;
    const float4* positions; const float4* normals; const float2* uvs; const int4* indices;
    const float4* motions;       // per-vertex motion vectors of the moving meshes (GeomRec::motBase)
    const float4* tangents;      // per-vertex tangent arrays of the meshes that carry them (GeomRec::tanXBase / tanYBase)
    const MaterialRec* materials; const TextureRec* textures; const LightRec* lights;
    int numGeoms, numLights, numEnvLights, numPrecomputed;
    int envLightIdx[8];
    int hasMedia;                // some material is a medium interface (Dielectric): path media must be tracked
    int hasExtMaterials;         // some material needs the EXT shading kernel (MAT_PLASTIC and later)
};

struct IntegratorData {          // integrators/pathtraceintegrator.cpp:21-33
    int maxDepth, rrDepth; float minContribution, epsilon, tMaxShadowRay, tMaxShadowJitter; V3 up;
    int backplateTex;            // -1 = none
    int lightSampleID, firstScatterSampleID, firstScatterTypeSampleID;
    // sample table layout (floats per record and offsets)
    int recFloats, off1D, off2D, offLight; int spp, sets;
};

}  // namespace yrt
