/* YulioRT.h — the asynchronous C entry points of the Yulio front end, re-hosted on Linux above device_cuda.
 *
 * Interface contract restated from the reference's devices/renderer/YulioRT.h:11-57 (enumerators, struct layouts and
 * the five entry points are the ABI that rt_test_dll/rt_test_dll.cpp:12-44 drives); the only change is the export
 * macro (the reference's is __declspec). Default values are the reference's (YulioRT.h:37-50). */
#pragma once

#if defined(_WIN32)
#define DllApi extern "C" __declspec(dllexport)
#else
#define DllApi extern "C" __attribute__((visibility("default")))
#endif

namespace Yulio {

enum ErrorCodeRT { NoError = 0, RenderingIsInProgress, MissingColladaFile, InvalidColladaFormat, UnitializedRenderer, FailedToPopulateStatus, UnknownError = 1000 };
enum StateRT { Inactive, Initialiazing, Rendering, Stopped, Done };

struct StatusRT {
    StateRT state;
    float progress;          /* relative progress in [0, 1] */
    ErrorCodeRT lastError;
};

struct ParamsRT {
    const char* renderer = "pathtracer";   /* "pathtracer" | "pt" */
    int size = 1536;                       /* cube face resolution */
    int depth = 10;                        /* max path depth */
    float tMaxShadowRay = 120.f;           /* shadow-ray length (scaled by the scene scale) */
    int spp = 256;                         /* samples per pixel (rounded up to a power of two by the sampler) */
    float ambientlight[3] = {.83f, .95f, .98f};
    float eyeSeparation = 2.5f;            /* parsed, not applied to Collada cameras (as in the reference) */
    bool toeIn = true;
    float zeroParallax = 75.f;             /* parsed, not applied to Collada cameras (as in the reference) */
    int jpegQuality = 90;
    bool debug = false;                    /* keep the 12 intermediate face images */
    int threadsPriority = 0;               /* accepted, unused (no CPU workers) */
    bool waterMark = false;
    const char* faceCullingMode = "default"; /* "default" | "forcesingle" | "forcedouble" */
};

DllApi bool StartRT(const char* colladaFile, const ParamsRT* params);
DllApi bool WaitRT();
DllApi bool StopRT(bool keepResults);
DllApi ErrorCodeRT GetLastErrorRT();
DllApi void GetCurrentStatusRT(StatusRT* status);

}  // namespace Yulio
