// sort.cu — ray sorting for the incoherent bounces of the wavefront path tracer.
//
// After a diffuse bounce neighbouring queue entries point anywhere; a warp of the traversal kernel then walks 32
// unrelated parts of the BVH (L1 hit rate, lanes per node phase) and a warp of the shading kernel gathers 32 unrelated
// materials / textures. Between k_shade(d) and k_trace_closest(d+1) the queue of path ids is therefore re-ordered by
//   key = Morton code of the ray origin's cell in the scene bounds (6 bits per axis) << 3 | direction octant
// with CUB's radix sort (library primitive, 21 key bits = 3 passes). Path state stays where it is (SoA indexed by path id),
// only the 4-byte ids move; results do not depend on queue order, so images stay bit-identical. The reference has no
// counterpart (it walks pixels in tile order on the CPU).
//
// MEASURED (round 1, one B200, C3 stand-in 1024x1024x64 spp): OFF by default. Sorting ids without moving the path state makes
// every state access of k_shade / the ray loads uncoalesced (queue order no longer follows path-id order): shade 71 -> 114 ms per
// face, closest-hit only 74.7 -> 72.0 ms, plus 13 ms of sorting. Kept behind cfg sort=1 for the scene sizes where traversal
// misses L2 (config 5), where the trade may flip.
#include <cub/device/device_radix_sort.cuh>
#include <stdexcept>
#include <string>

#include "device_internal.hpp"

namespace yrt {

__device__ __forceinline__ uint32_t spread6(uint32_t x) {        // 6 bits -> every third bit
    x &= 0x3fu;
    x = (x | (x << 8)) & 0x0300fu;
    x = (x | (x << 4)) & 0x030c3u;
    x = (x | (x << 2)) & 0x09249u;
    return x;
}

__global__ void __launch_bounds__(256) k_sort_keys(SceneData sc, WavefrontBuffers wb, int queueSel, uint32_t n) {
    const uint32_t* __restrict__ queue = queueSel ? wb.queueB : wb.queueA;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t pid = queue[i];
        const float4 o = wb.rayO[pid], d = wb.rayD[pid];
        const float cx = fminf(fmaxf((o.x - sc.bboxLo.x) * sc.bboxRcpExtent.x, 0.f), 1.f);
        const float cy = fminf(fmaxf((o.y - sc.bboxLo.y) * sc.bboxRcpExtent.y, 0.f), 1.f);
        const float cz = fminf(fmaxf((o.z - sc.bboxLo.z) * sc.bboxRcpExtent.z, 0.f), 1.f);
        const uint32_t qx = min(63u, (uint32_t)(cx * 64.f)), qy = min(63u, (uint32_t)(cy * 64.f)), qz = min(63u, (uint32_t)(cz * 64.f));
        const uint32_t oct = (d.x >= 0.f ? 4u : 0u) | (d.y >= 0.f ? 2u : 0u) | (d.z >= 0.f ? 1u : 0u);
        wb.sortKeys[i] = (((spread6(qx) << 2) | (spread6(qy) << 1) | spread6(qz)) << 3) | oct;
    }
}
void launch_sort_keys(const FrameConst& fc, const WavefrontBuffers& wb, int queueSel, uint32_t n, LaunchCfg lc) {
    k_sort_keys<<<lc.blocks, 256, 0, lc.stream>>>(fc.scene, wb, queueSel, n);
}

size_t sort_temp_bytes(uint32_t capacity) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                    (int)capacity, 0, YRT_SORT_KEY_BITS, (cudaStream_t)0);
    return bytes;
}

void sort_queue(const WavefrontBuffers& wb, int queueSel, uint32_t n, void* temp, size_t tempBytes, cudaStream_t stream) {
    const cudaError_t e = cub::DeviceRadixSort::SortPairs(temp, tempBytes, (const uint32_t*)wb.sortKeys, wb.sortKeysOut,
                                                          (const uint32_t*)(queueSel ? wb.queueB : wb.queueA), wb.queueS, (int)n, 0, YRT_SORT_KEY_BITS, stream);
    if (e != cudaSuccess) throw std::runtime_error(std::string("CUDA error in the ray sort: ") + cudaGetErrorString(e));
}

}  // namespace yrt
