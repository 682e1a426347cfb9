// sampler_host.hpp — host-side construction of the per-frame precomputed sample table.
//
// The reference precomputes, once per frame, `sets` x `spp` high-dimensional samples with a
// Park-Miller / Bays-Durham LCG and multi-jittered patterns and warps the pixel sample through a
// tabulated filter distribution. The table is tiny (KBs..MBs), strictly sequential, and must be
// bit-identical for the images to agree, so device_cuda builds it on the host with the same
// arithmetic and uploads it:
//   SamplerFactory::init     devices/device_singleray/samplers/sampler.cpp:85-158
//   jittered/multiJittered   devices/device_singleray/samplers/patterns.h:28-68
//   Random                   common/math/random.h:24-78
//   Permutation              common/math/permutation.h:42-48
//   vector_t::shuffle        common/sys/stl/vector.h:129-133
//   Filter::init/sample      devices/device_singleray/filters/filter.cpp:22-43, bsplinefilter.h:25-43, boxfilter.h:25-42
//   Distribution1D/2D        devices/device_singleray/samplers/distribution1d.cpp:42-74, distribution2d.cpp:34-68
// Compiled with -ffp-contract=off: every float expression below is evaluated as written.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <utility>
#include <vector>

namespace yrt {

// Minimal-standard generator with a 32-entry shuffle table.
class Lcg {
public:
    explicit Lcg(int s = 27) { seedWith(s); }
    void seedWith(int s) {
        seed_ = (s == 0) ? 1 : (s < 0 ? -s : s);
        for (int j = 32 + 7; j >= 0; j--) { step(); if (j < 32) table_[j] = seed_; }
        state_ = table_[0];
    }
    int nextInt() {
        step();
        const int j = state_ / (1 + (2147483647 - 1) / 32);
        state_ = table_[j];
        table_[j] = seed_;
        return state_;
    }
    int nextInt(int limit) { return nextInt() % limit; }
    float nextFloat() {
        const float v = nextInt() / 2147483647.0f, cap = 1.0f - 1.1920928955078125e-07f;
        return v < cap ? v : cap;
    }
private:
    void step() {
        const int k = seed_ / 127773;
        seed_ = 16807 * (seed_ - k * 127773) - 2836 * k;
        if (seed_ < 0) seed_ += 2147483647;
    }
    int seed_, state_, table_[32];
};

inline std::vector<int> randomPermutation(int n, Lcg& rng) {
    std::vector<int> p(n);
    for (int i = 0; i < n; i++) p[i] = i;
    for (int i = 0; i < n; i++) std::swap(p[i], p[rng.nextInt(n)]);
    return p;
}

inline void jittered1D(float* out, uint32_t n, Lcg& rng) {
    const float scale = 1.0f / n;
    const std::vector<int> perm = randomPermutation((int)n, rng);
    for (uint32_t i = 0; i < n; i++) out[perm[i]] = (float(i) + rng.nextFloat()) * scale;
}

struct F2 { float x, y; };

inline void multiJittered2D(F2* out, uint32_t N, Lcg& rng) {
    uint32_t b = (uint32_t)sqrtf(float(N));
    if (b * b < N) b++;
    std::vector<F2> grid((size_t)b * b);              // grid[i*b + j]
    std::vector<uint32_t> numbers(b);
    for (uint32_t i = 0; i < b; i++) numbers[i] = i;
    auto shuffle = [&]() { for (size_t i = 0; i < numbers.size(); i++) std::swap(numbers[i], numbers[rng.nextInt((int)numbers.size())]); };
    for (uint32_t i = 0; i < b; i++) {
        shuffle();
        for (uint32_t j = 0; j < b; j++)
            grid[(size_t)i * b + j].x = float(i) / float(b) + (numbers[j] + rng.nextFloat()) / float(b * b);
    }
    for (uint32_t i = 0; i < b; i++) {
        shuffle();
        for (uint32_t j = 0; j < b; j++)
            grid[(size_t)j * b + i].y = float(i) / float(b) + (numbers[j] + rng.nextFloat()) / float(b * b);
    }
    const std::vector<int> perm = randomPermutation((int)N, rng);
    for (uint32_t n = 0; n < N; n++) { const uint32_t np = (uint32_t)perm[n]; out[n] = grid[(size_t)(np / b) * b + np % b]; }
}

// Piecewise-constant 1-D distribution.
struct Dist1D {
    std::vector<float> pdf, cdf;
    void init(const float* f, size_t n) {
        pdf.assign(n, 0.f); cdf.assign(n + 1, 0.f);
        for (size_t i = 1; i < n + 1; i++) cdf[i] = cdf[i - 1] + f[i - 1];
        const float rcpSum = cdf[n] == 0.0f ? 0.0f : 1.0f / cdf[n];
        for (size_t i = 1; i < n + 1; i++) { pdf[i - 1] = f[i - 1] * rcpSum * n; cdf[i] *= rcpSum; }
        cdf[n] = 1.0f;
    }
    // returns (index + fraction, pdf)
    std::pair<float, float> sample(float u) const {
        const size_t n = pdf.size();
        const float* p = std::upper_bound(cdf.data(), cdf.data() + n, u);
        int index = int(p - cdf.data() - 1);
        index = std::max(0, std::min(index, int(n) - 1));
        const float fraction = (u - cdf[index]) * (1.0f / (cdf[index + 1] - cdf[index]));
        return {float(index) + fraction, pdf[index]};
    }
};

struct Dist2D {
    size_t width = 0, height = 0;
    std::vector<Dist1D> rows; Dist1D col;
    // f[row][column], `height` rows of `width` entries
    void init(const std::vector<std::vector<float>>& f, size_t w, size_t h) {
        width = w; height = h; rows.assign(h, Dist1D());
        std::vector<float> fy(h);
        for (size_t y = 0; y < h; y++) {
            fy[y] = 0.0f;
            for (size_t x = 0; x < w; x++) fy[y] += f[y][x];
            rows[y].init(f[y].data(), w);
        }
        col.init(fy.data(), h);
    }
    // returns value (x = position inside the row, y = row position) and pdf
    void sample(float ux, float uy, float& vx, float& vy, float& pdf) const {
        const auto sy = col.sample(uy);
        int y = int(sy.first); y = std::max(0, std::min(y, int(height) - 1));
        const auto sx = rows[y].sample(ux);
        vx = sx.first; vy = sy.first; pdf = sx.second * sy.second;
    }
};

enum FilterKind { FILTER_NONE = 0, FILTER_BOX = 1, FILTER_BSPLINE = 2 };

struct PixelFilter {
    FilterKind kind = FILTER_BSPLINE;
    float width = 4.f, height = 4.f;
    uint32_t tableSize = 256;
    Dist2D dist;
    float eval(float dx, float dy) const {
        if (kind == FILTER_BOX) return (fabsf(dx) <= 0.5f && fabsf(dy) <= 0.5f) ? 1.0f : 0.0f;
        const float d = sqrtf(dx * dx + dy * dy);
        if (d > 2.0f) return 0.0f;
        if (d < 1.0f) { const float t = 1.0f - d; return ((((-3.0f * t) + 3.0f) * t + 3.0f) * t + 1.0f) / 6.0f; }
        const float t = 2.0f - d; return t * t * t / 6.0f;
    }
    void init(FilterKind k) {
        kind = k;
        if (k == FILTER_NONE) return;
        if (k == FILTER_BOX) { width = 2.0f * 0.5f; height = 2.0f * 0.5f; } else { width = 4.0f; height = 4.0f; }
        const float inv = 1.0f / tableSize;
        std::vector<std::vector<float>> a(tableSize, std::vector<float>(tableSize));
        for (uint32_t x = 0; x < tableSize; ++x)
            for (uint32_t y = 0; y < tableSize; ++y)
                a[x][y] = fabsf(eval((x + 0.5f) * inv * width - width * 0.5f, (y + 0.5f) * inv * height - height * 0.5f));
        dist.init(a, tableSize, tableSize);
    }
    F2 sample(F2 uv) const {
        float vx, vy, pdf; dist.sample(uv.x, uv.y, vx, vy, pdf);
        F2 r; r.x = vx / tableSize * width - width * 0.5f; r.y = vy / tableSize * height - height * 0.5f;
        return r;
    }
};

inline uint32_t roundUpPow2(uint32_t v) { v--; v |= v >> 1; v |= v >> 2; v |= v >> 4; v |= v >> 8; v |= v >> 16; return v + 1; }

// One record per (set, sample): {pixel.x, pixel.y, time, lens.x, lens.y, 1D[n1], 2D[2*n2]}; the caller
// appends precomputed light samples (HDRI) behind it.
struct SampleTable {
    int spp = 1, sets = 64, n1 = 0, n2 = 0;
    std::vector<float> rec;     // sets * spp * (5 + n1 + 2*n2)
    int recFloats() const { return 5 + n1 + 2 * n2; }
    const float* at(int set, int s) const { return rec.data() + ((size_t)set * spp + s) * recFloats(); }
};

inline SampleTable buildSampleTable(int sppRequested, int sets, int n1, int n2, int iteration, const PixelFilter* filter) {
    SampleTable t;
    t.sets = sets; t.n1 = n1; t.n2 = n2;
    t.spp = (int)roundUpPow2((uint32_t)sppRequested);
    const int spp = t.spp;
    const int chunk = std::max(spp, 64);
    const int currentChunk = int(iteration * spp) / chunk;
    const int offset = (iteration * spp) % chunk;
    Lcg rng; rng.seedWith(currentChunk * 5897);
    std::vector<F2> pixel(chunk), lens(chunk), s2(chunk);
    std::vector<float> time(chunk), s1(chunk);
    const int R = t.recFloats();
    t.rec.assign((size_t)sets * spp * R, 0.f);
    for (int set = 0; set < sets; set++) {
        multiJittered2D(pixel.data(), chunk, rng);
        jittered1D(time.data(), chunk, rng);
        multiJittered2D(lens.data(), chunk, rng);
        for (int s = 0; s < spp; s++) {
            float* r = &t.rec[((size_t)set * spp + s) * R];
            F2 p = pixel[offset + s];
            if (filter && filter->kind != FILTER_NONE) { const F2 f = filter->sample(p); p.x = f.x + 0.5f; p.y = f.y + 0.5f; }
            r[0] = p.x; r[1] = p.y; r[2] = time[offset + s]; r[3] = lens[offset + s].x; r[4] = lens[offset + s].y;
        }
        for (int d = 0; d < n1; d++) {
            jittered1D(s1.data(), chunk, rng);
            for (int s = 0; s < spp; s++) t.rec[((size_t)set * spp + s) * R + 5 + d] = s1[offset + s];
        }
        for (int d = 0; d < n2; d++) {
            multiJittered2D(s2.data(), chunk, rng);
            for (int s = 0; s < spp; s++) {
                float* r = &t.rec[((size_t)set * spp + s) * R + 5 + n1 + 2 * d];
                r[0] = s2[offset + s].x; r[1] = s2[offset + s].y;
            }
        }
    }
    return t;
}

}  // namespace yrt
