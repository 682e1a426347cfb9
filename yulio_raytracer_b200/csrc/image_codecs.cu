// image_codecs.cu — the data formats either side of the render path:
//
//   decode  rtNewImageFromFile for .jpg/.jpeg (nvJPEG, decoded on the GPU) and .png (zlib inflate + PNG unfiltering on the host),
//           producing what the reference's loaders produce: an RGBA8 image whose row 0 is the BOTTOM scanline of the picture
//           (common/image/jpeg.cpp:53-63 stores libjpeg-turbo rows with yFlip; common/image/freeimage.cpp:40-78 copies FreeImage's
//           bottom-up rows), alpha 255 where the file has none. The reference uses libjpeg-turbo / FreeImage (Windows binaries in
//           the mount); JPEG decoders are not bit-identical to each other, the tests state the tolerance against libjpeg-turbo.
//   strip   the 12W x H stereo cube-map strip assembled on the device from the frames as they are rendered (segment order Left Right
//           Up Down Back Front, cameras 6-11 first: devices/renderer/renderer.cpp:665-718), the centred watermark blend on faces
//           0-3 (renderer.cpp:637-654) and the JPEG encode at `jpegQuality` (common/image/jpeg.cpp:207-250: 4:2:0, baseline) with
//           nvJPEG — the frames never travel to the host one by one.
//
// nvJPEG is a library primitive (like CUB's radix sort in bvh_build.cu). It is opened with dlopen on first use so that the render
// path does not depend on it; a missing library is a clear error from the JPEG entry points only.
#include <dlfcn.h>
#include <zlib.h>

#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include <nvjpeg.h>

#include "device_impl.hpp"

namespace yrt {

// ------------------------------------------------------------------------------------------------
// nvJPEG through dlopen
// ------------------------------------------------------------------------------------------------
struct NvJpeg {
    void* lib = nullptr; nvjpegHandle_t handle = nullptr;
    decltype(&nvjpegCreateSimple) CreateSimple; decltype(&nvjpegCreateEx) CreateEx; decltype(&nvjpegJpegStateCreate) JpegStateCreate; decltype(&nvjpegJpegStateDestroy) JpegStateDestroy;
    decltype(&nvjpegGetImageInfo) GetImageInfo; decltype(&nvjpegDecode) Decode;
    decltype(&nvjpegEncoderStateCreate) EncoderStateCreate; decltype(&nvjpegEncoderStateDestroy) EncoderStateDestroy;
    decltype(&nvjpegEncoderParamsCreate) EncoderParamsCreate; decltype(&nvjpegEncoderParamsDestroy) EncoderParamsDestroy;
    decltype(&nvjpegEncoderParamsSetQuality) EncoderParamsSetQuality; decltype(&nvjpegEncoderParamsSetSamplingFactors) EncoderParamsSetSamplingFactors;
    decltype(&nvjpegEncoderParamsSetOptimizedHuffman) EncoderParamsSetOptimizedHuffman;
    decltype(&nvjpegEncodeImage) EncodeImage; decltype(&nvjpegEncodeRetrieveBitstream) EncodeRetrieveBitstream;
};

// one library handle per CUDA device (a group device decodes the same texture on every member's GPU)
static NvJpeg& nvjpeg() {
    static NvJpeg base; static std::once_flag once; static std::string err; static std::mutex mtx; static NvJpeg perGpu[64]; static bool ready[64];
    std::call_once(once, [] {
        for (const char* name : {"libnvjpeg.so.12", "libnvjpeg.so", "/usr/local/cuda/lib64/libnvjpeg.so.12"}) { base.lib = dlopen(name, RTLD_NOW | RTLD_LOCAL); if (base.lib) break; }
        if (!base.lib) { err = std::string("device_cuda: nvJPEG not available (") + dlerror() + ")"; return; }
        NvJpeg& nj = base;
#define NJ_SYM(field, sym) nj.field = (decltype(nj.field))dlsym(nj.lib, #sym); if (!nj.field) { err = "device_cuda: nvJPEG symbol missing: " #sym; return; }
        NJ_SYM(CreateSimple, nvjpegCreateSimple) NJ_SYM(CreateEx, nvjpegCreateEx) NJ_SYM(JpegStateCreate, nvjpegJpegStateCreate) NJ_SYM(JpegStateDestroy, nvjpegJpegStateDestroy)
        NJ_SYM(GetImageInfo, nvjpegGetImageInfo) NJ_SYM(Decode, nvjpegDecode)
        NJ_SYM(EncoderStateCreate, nvjpegEncoderStateCreate) NJ_SYM(EncoderStateDestroy, nvjpegEncoderStateDestroy)
        NJ_SYM(EncoderParamsCreate, nvjpegEncoderParamsCreate) NJ_SYM(EncoderParamsDestroy, nvjpegEncoderParamsDestroy)
        NJ_SYM(EncoderParamsSetQuality, nvjpegEncoderParamsSetQuality) NJ_SYM(EncoderParamsSetSamplingFactors, nvjpegEncoderParamsSetSamplingFactors)
        NJ_SYM(EncoderParamsSetOptimizedHuffman, nvjpegEncoderParamsSetOptimizedHuffman)
        NJ_SYM(EncodeImage, nvjpegEncodeImage) NJ_SYM(EncodeRetrieveBitstream, nvjpegEncodeRetrieveBitstream)
#undef NJ_SYM
    });
    if (!err.empty() || !base.EncodeRetrieveBitstream) throw std::runtime_error(err.empty() ? "device_cuda: nvJPEG not initialised" : err);
    int gpu = 0; cudaGetDevice(&gpu);
    if (gpu < 0 || gpu >= 64) throw std::runtime_error("device_cuda: unexpected CUDA device index");
    std::lock_guard<std::mutex> lock(mtx);
    if (!ready[gpu]) {
        NvJpeg nj = base;
        // chroma planes are up-sampled with interpolation, as libjpeg-turbo's default "fancy upsampling" does in the reference's reader
        if (nj.CreateEx(NVJPEG_BACKEND_DEFAULT, nullptr, nullptr, NVJPEG_FLAGS_UPSAMPLING_WITH_INTERPOLATION, &nj.handle) != NVJPEG_STATUS_SUCCESS) {
            nj.handle = nullptr;
            if (nj.CreateSimple(&nj.handle) != NVJPEG_STATUS_SUCCESS) throw std::runtime_error("device_cuda: nvjpegCreate failed");
        }
        perGpu[gpu] = nj; ready[gpu] = true;
    }
    return perGpu[gpu];
}
#define NJ_CK(x) do { const nvjpegStatus_t s_ = (x); if (s_ != NVJPEG_STATUS_SUCCESS) throw std::runtime_error(std::string("nvJPEG error ") + std::to_string((int)s_) + " in " #x); } while (0)

static bool read_file(const char* file, std::vector<unsigned char>& out) {
    FILE* f = fopen(file, "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END); const long n = ftell(f); fseek(f, 0, SEEK_SET);
    out.resize(n > 0 ? (size_t)n : 0);
    const bool ok = n > 0 && fread(out.data(), 1, (size_t)n, f) == (size_t)n;
    fclose(f);
    return ok;
}

// ------------------------------------------------------------------------------------------------
// JPEG decode
// ------------------------------------------------------------------------------------------------
// interleaved RGB (top-down) -> RGBA8, rows flipped; the byte passes through Color4 and back exactly as in jpeg.cpp:55-62
// (b * (1/255) then clamp * 255 truncated: the identity for every byte, checked in tests)
__global__ void k_rgb_to_rgba_flip(const unsigned char* __restrict__ rgb, int pitch, int w, int h, uchar4* __restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= w) return;
    for (int y = blockIdx.y; y < h; y += gridDim.y) {                // grid.y is capped at 65535: taller images take several rows per block
        const unsigned char* p = rgb + (size_t)y * pitch + 3 * x;
        out[(size_t)(h - 1 - y) * w + x] = make_uchar4(p[0], p[1], p[2], 255);
    }
}
static unsigned rows_grid(size_t h) { return (unsigned)(h < 1 ? 1 : (h > 65535 ? 65535 : h)); }
struct DevFree { void operator()(void* p) const { if (p) cudaFree(p); } };      // device scratch released on every exit path

std::shared_ptr<ImageObj> decode_jpeg_file(const char* file) {
    std::vector<unsigned char> data;
    if (!read_file(file, data)) { printf("cannot read file %s: cannot open\n", file); return nullptr; }
    try {
        NvJpeg& nj = nvjpeg();
        int comps = 0, widths[NVJPEG_MAX_COMPONENT] = {0}, heights[NVJPEG_MAX_COMPONENT] = {0}; nvjpegChromaSubsampling_t css;
        NJ_CK(nj.GetImageInfo(nj.handle, data.data(), data.size(), &comps, &css, widths, heights));
        const int w = widths[0], h = heights[0];
        if (w <= 0 || h <= 0) throw std::runtime_error("improper dimensions");
        unsigned char* rgbRaw = nullptr; YRT_CK(cudaMalloc((void**)&rgbRaw, (size_t)w * h * 3));
        std::unique_ptr<unsigned char, DevFree> rgb(rgbRaw);
        nvjpegJpegState_t state; NJ_CK(nj.JpegStateCreate(nj.handle, &state));
        nvjpegImage_t dst; memset(&dst, 0, sizeof(dst)); dst.channel[0] = rgb.get(); dst.pitch[0] = (size_t)w * 3;
        const nvjpegStatus_t s = nj.Decode(nj.handle, state, data.data(), data.size(), NVJPEG_OUTPUT_RGBI, &dst, nullptr);
        nj.JpegStateDestroy(state);
        if (s != NVJPEG_STATUS_SUCCESS) throw std::runtime_error("nvjpegDecode failed with status " + std::to_string((int)s));
        YRT_CK(cudaStreamSynchronize(nullptr));                       // the decode ran on the legacy stream: surface its errors here
        auto img = std::make_shared<ImageObj>();
        img->width = w; img->height = h; img->format = TEX_RGBA8;
        YRT_CK(cudaMalloc(&img->devPixels, (size_t)w * h * 4));       // owned (and released) by the ImageObj
        k_rgb_to_rgba_flip<<<dim3((w + 127) / 128, rows_grid((size_t)h)), 128>>>(rgb.get(), w * 3, w, h, (uchar4*)img->devPixels);
        YRT_CK(cudaGetLastError());
        img->storage.resize((size_t)w * h * 4);                       // host mirror: HDRI importance tables and backplates read texels on the host
        YRT_CK(cudaMemcpy(img->storage.data(), img->devPixels, img->storage.size(), cudaMemcpyDeviceToHost));
        img->pixels = img->storage.data();
        return img;
    } catch (const std::exception& e) { printf("cannot read file %s: %s\n", file, e.what()); return nullptr; }
}

// ------------------------------------------------------------------------------------------------
// PNG decode (non-interlaced; grey / RGB / palette / grey+alpha / RGBA at 1..16 bits per sample)
// ------------------------------------------------------------------------------------------------
static uint32_t be32(const unsigned char* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

// Decodes into RGBA8, top-down. FreeImage hands the reference 24/32-bit bitmaps only for RGB / RGBA files; palette, grey and 16-bit files
// leave the reference's image unset (freeimage.cpp:57-84 handles bpp 24 and 32 only) — here they are expanded to RGBA8 (stated deviation).
bool decode_png(const std::vector<unsigned char>& file, int& w, int& h, std::vector<unsigned char>& rgba, std::string& err) {
    static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (file.size() < 8 + 25 || memcmp(file.data(), sig, 8) != 0) { err = "not a PNG file"; return false; }
    size_t pos = 8; int depth = 0, ctype = 0, interlace = 0; bool haveHdr = false;
    std::vector<unsigned char> idat, plte, trns;
    while (pos + 12 <= file.size()) {
        const uint32_t len = be32(&file[pos]); const unsigned char* type = &file[pos + 4];
        if (pos + 12 + (size_t)len > file.size()) { err = "truncated chunk"; return false; }
        const unsigned char* d = &file[pos + 8];
        if (!memcmp(type, "IHDR", 4) && len >= 13) { w = (int)be32(d); h = (int)be32(d + 4); depth = d[8]; ctype = d[9]; interlace = d[12]; haveHdr = true; }
        else if (!memcmp(type, "PLTE", 4)) plte.assign(d, d + len);
        else if (!memcmp(type, "tRNS", 4)) trns.assign(d, d + len);
        else if (!memcmp(type, "IDAT", 4)) idat.insert(idat.end(), d, d + len);
        else if (!memcmp(type, "IEND", 4)) break;
        pos += 12 + (size_t)len;
    }
    if (!haveHdr || w <= 0 || h <= 0) { err = "missing IHDR"; return false; }
    if (interlace) { err = "interlaced PNG not supported"; return false; }
    const int channels = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
    if (!channels || !(depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)) { err = "unsupported colour type / bit depth"; return false; }
    const size_t bpp = (size_t)(channels * depth + 7) / 8, rowBytes = ((size_t)w * channels * depth + 7) / 8;
    std::vector<unsigned char> raw((rowBytes + 1) * (size_t)h);
    uLongf rawLen = (uLongf)raw.size();
    if (uncompress(raw.data(), &rawLen, idat.data(), (uLong)idat.size()) != Z_OK || rawLen != raw.size()) { err = "zlib inflate failed"; return false; }
    // unfilter in place (PNG spec 9.2)
    std::vector<unsigned char> zero(rowBytes, 0);
    for (int y = 0; y < h; y++) {
        unsigned char* row = &raw[(rowBytes + 1) * (size_t)y]; const int ft = row[0]; unsigned char* cur = row + 1;
        const unsigned char* up = y ? &raw[(rowBytes + 1) * (size_t)(y - 1) + 1] : zero.data();
        for (size_t i = 0; i < rowBytes; i++) {
            const int a = i >= bpp ? cur[i - bpp] : 0, b = up[i], c = i >= bpp ? up[i - bpp] : 0;
            int pred = 0;
            switch (ft) {
            case 0: pred = 0; break; case 1: pred = a; break; case 2: pred = b; break; case 3: pred = (a + b) >> 1; break;
            case 4: { const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c); pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); break; }
            default: err = "bad filter type"; return false;
            }
            cur[i] = (unsigned char)(cur[i] + pred);
        }
    }
    rgba.assign((size_t)w * h * 4, 255);
    auto sample = [&](const unsigned char* row, size_t idx) -> unsigned {      // idx-th sample of the row, scaled to 8 bits
        if (depth == 8) return row[idx];
        if (depth == 16) return row[2 * idx];
        const unsigned perByte = 8 / depth, shift = (unsigned)((perByte - 1 - idx % perByte) * depth);
        const unsigned v = (row[idx / perByte] >> shift) & ((1u << depth) - 1u);
        return ctype == 3 ? v : v * 255u / ((1u << depth) - 1u);
    };
    for (int y = 0; y < h; y++) {
        const unsigned char* row = &raw[(rowBytes + 1) * (size_t)y + 1];
        unsigned char* o = &rgba[(size_t)y * w * 4];
        for (int x = 0; x < w; x++, o += 4) {
            switch (ctype) {
            case 0: { const unsigned g = sample(row, x); o[0] = o[1] = o[2] = (unsigned char)g; break; }
            case 2: o[0] = (unsigned char)sample(row, 3 * (size_t)x); o[1] = (unsigned char)sample(row, 3 * (size_t)x + 1); o[2] = (unsigned char)sample(row, 3 * (size_t)x + 2); break;
            case 3: { const unsigned i = sample(row, x); if (3 * i + 2 < plte.size()) { o[0] = plte[3 * i]; o[1] = plte[3 * i + 1]; o[2] = plte[3 * i + 2]; } if (i < trns.size()) o[3] = trns[i]; break; }
            case 4: { const unsigned g = sample(row, 2 * (size_t)x); o[0] = o[1] = o[2] = (unsigned char)g; o[3] = (unsigned char)sample(row, 2 * (size_t)x + 1); break; }
            case 6: for (int k = 0; k < 4; k++) o[k] = (unsigned char)sample(row, 4 * (size_t)x + k); break;
            }
        }
    }
    return true;
}

std::shared_ptr<ImageObj> decode_png_file(const char* file, bool flipVertical, bool flipHorizontal) {
    std::vector<unsigned char> data, rgba; int w = 0, h = 0; std::string err;
    if (!read_file(file, data)) { printf("cannot read file %s: cannot open\n", file); return nullptr; }
    if (!decode_png(data, w, h, rgba, err)) { printf("cannot read file %s: %s\n", file, err.c_str()); return nullptr; }
    auto img = std::make_shared<ImageObj>();
    img->width = w; img->height = h; img->format = TEX_RGBA8; img->storage.resize(rgba.size());
    // FreeImage rows are bottom-up: image row y = picture row h-1-y, unless flipVertical (freeimage.cpp:34-36)
    for (int y = 0; y < h; y++) {
        const unsigned char* src = &rgba[(size_t)(flipVertical ? y : h - 1 - y) * w * 4]; unsigned char* dst = &img->storage[(size_t)y * w * 4];
        if (!flipHorizontal) memcpy(dst, src, (size_t)w * 4);
        else for (int x = 0; x < w; x++) memcpy(dst + 4 * x, src + 4 * (size_t)(w - 1 - x), 4);
    }
    img->pixels = img->storage.data();
    return img;
}

// ------------------------------------------------------------------------------------------------
// stereo cube-map strip on the device
// ------------------------------------------------------------------------------------------------
// face c -> strip segment: segments 0-5 = Left Right Up Down Back Front of cameras 6-11, segments 6-11 = the same of cameras 0-5
// (renderer.cpp:677-710: eyeIndex = segment / 6 == 0 ? 1 : 0; faces 3,1,4,5,2,0)
static int strip_segment(int cubeFaceIndex) {
    static const int inv[6] = {5, 1, 4, 0, 2, 3};
    const int c = ((cubeFaceIndex % 12) + 12) % 12;
    return (c >= 6 ? 0 : 6) + inv[c % 6];
}

// copies one RGB8 frame (row stride fbStride) into its segment, blending the watermark (RGBA8, top-down, already mirrored as the
// reference loads it) centred on the face: blended = (1 - a) * ic + a * wc on Color4 lanes, stored as (uchar)(clamp(c) * 255)
__global__ void k_strip_face(const unsigned char* __restrict__ fb, int fbStride, int w, int h, unsigned char* __restrict__ strip, int stripStride, int xOfs,
                             const uchar4* __restrict__ wm, int wmW, int wmH, int wmX0, int wmY0) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= w) return;
    for (int y = blockIdx.y; y < h; y += gridDim.y) {
        const unsigned char* p = fb + (size_t)y * fbStride + 3 * x;
        float c[3] = {(float)p[0], (float)p[1], (float)p[2]};
        unsigned char o[3] = {p[0], p[1], p[2]};
        if (wm) {
            const int wx = x - wmX0, wy = y - wmY0;
            if (wx >= 0 && wx < wmW && wy >= 0 && wy < wmH) {
                const uchar4 t = wm[(size_t)wy * wmW + wx];
                const float k = 1.f / 255.f, a = t.w * k, wc[3] = {t.x * k, t.y * k, t.z * k};
                for (int i = 0; i < 3; i++) { const float b = (1.f - a) * (c[i] * k) + a * wc[i]; o[i] = (unsigned char)(rclamp(b) * 255.0f); }
            }
        }
        unsigned char* q = strip + (size_t)y * stripStride + 3 * (xOfs + x);
        q[0] = o[0]; q[1] = o[1]; q[2] = o[2];
    }
}

void strip_begin(yrt_device* dev, size_t faceW, size_t faceH) {
    CubeStrip& s = dev->strip;
    if (s.w != faceW || s.h != faceH || !s.dev) {
        if (s.dev) cudaFree(s.dev);
        s.w = faceW; s.h = faceH; s.dev = nullptr;
        const size_t bytes = 12 * faceW * faceH * 3;
        YRT_CK(cudaMalloc((void**)&s.dev, bytes != 0 ? bytes : 1));
    }
    YRT_CK(cudaMemsetAsync(s.dev, 0, 12 * faceW * faceH * 3, dev->stream));
    s.facesAdded = 0;
}

void strip_set_watermark(yrt_device* dev, const char* pngFile) {
    CubeStrip& s = dev->strip;
    if (s.wm) { cudaFree(s.wm); s.wm = nullptr; }
    s.wmW = s.wmH = 0;
    if (!pngFile || !*pngFile) return;
    auto img = decode_png_file(pngFile, true, true);                // LoadWaterMark: loadFreeImage(data, size, scale, true, true)  renderer.cpp:84
    if (!img) throw std::runtime_error(std::string("device_cuda: cannot load watermark ") + pngFile);
    s.wmW = img->width; s.wmH = img->height;
    YRT_CK(cudaMalloc((void**)&s.wm, img->storage.size()));
    YRT_CK(cudaMemcpy(s.wm, img->storage.data(), img->storage.size(), cudaMemcpyHostToDevice));
}

void strip_add_face_device(yrt_device* dev, const unsigned char* devRgb, size_t strideBytes, size_t w, size_t h, int cubeFaceIndex, int watermark);
void strip_add_face(yrt_device* dev, FrameBufferHandle* fb, int cubeFaceIndex, int watermark) {
    if (fb->format != 2) throw std::runtime_error("device_cuda: the strip takes RGB8 frames of the size given to yrtxStripBegin");
    strip_add_face_device(dev, (const unsigned char*)fb->devPacked, fb->strideBytes, fb->width, fb->height, cubeFaceIndex, watermark);
}
// the same for any RGB8 frame resident on this device (a group device assembles its members' bands on member 0, group_api.cu)
void strip_add_face_device(yrt_device* dev, const unsigned char* devRgb, size_t strideBytes, size_t w, size_t h, int cubeFaceIndex, int watermark) {
    CubeStrip& s = dev->strip;
    if (!s.dev) throw std::runtime_error("device_cuda: yrtxStripBegin was not called");
    if (w != s.w || h != s.h) throw std::runtime_error("device_cuda: the strip takes RGB8 frames of the size given to yrtxStripBegin");
    struct { const void* devPacked; size_t strideBytes; } fbv{devRgb, strideBytes}; auto* fb = &fbv;
    const int seg = strip_segment(cubeFaceIndex);
    const bool wm = watermark && s.wm && (((cubeFaceIndex % 12) + 12) % 6) < 4;             // front, right, back, left only (renderer.cpp:637)
    // xDst = x + (W - w) * .5f truncated (renderer.cpp:642-643)
    const int x0 = (int)(0 + ((float)s.w - (float)s.wmW) * .5f), y0 = (int)(0 + ((float)s.h - (float)s.wmH) * .5f);
    k_strip_face<<<dim3((unsigned)((s.w + 127) / 128), rows_grid(s.h)), 128, 0, dev->stream>>>((const unsigned char*)fb->devPacked, (int)fb->strideBytes, (int)s.w, (int)s.h,
                                                                                                s.dev, (int)(12 * s.w * 3), seg * (int)s.w, wm ? s.wm : nullptr, s.wmW, s.wmH, x0, y0);
    YRT_CK(cudaGetLastError());
    s.facesAdded++;
}


void strip_read(yrt_device* dev, void* rgb) {
    CubeStrip& s = dev->strip;
    if (!s.dev) throw std::runtime_error("device_cuda: no strip");
    YRT_CK(cudaMemcpyAsync(rgb, s.dev, 12 * s.w * s.h * 3, cudaMemcpyDeviceToHost, dev->stream));
    YRT_CK(cudaStreamSynchronize(dev->stream));
}

// RGB8 interleaved device image -> baseline JPEG, 4:2:0 (jpeg_set_defaults), quality as given (jpeg_set_quality)
static void encode_jpeg(yrt_device* dev, const unsigned char* devRgb, size_t pitch, int w, int h, int quality, const char* file) {
    NvJpeg& nj = nvjpeg();
    nvjpegEncoderState_t st; nvjpegEncoderParams_t pr;
    NJ_CK(nj.EncoderStateCreate(nj.handle, &st, dev->stream));
    NJ_CK(nj.EncoderParamsCreate(nj.handle, &pr, dev->stream));
    try {
        NJ_CK(nj.EncoderParamsSetQuality(pr, quality < 1 ? 1 : (quality > 100 ? 100 : quality), dev->stream));
        NJ_CK(nj.EncoderParamsSetSamplingFactors(pr, NVJPEG_CSS_420, dev->stream));
        NJ_CK(nj.EncoderParamsSetOptimizedHuffman(pr, 0, dev->stream));
        nvjpegImage_t src; memset(&src, 0, sizeof(src)); src.channel[0] = (unsigned char*)devRgb; src.pitch[0] = pitch;
        NJ_CK(nj.EncodeImage(nj.handle, st, pr, &src, NVJPEG_INPUT_RGBI, w, h, dev->stream));
        size_t len = 0;
        NJ_CK(nj.EncodeRetrieveBitstream(nj.handle, st, nullptr, &len, dev->stream));
        YRT_CK(cudaStreamSynchronize(dev->stream));
        std::vector<unsigned char> out(len);
        NJ_CK(nj.EncodeRetrieveBitstream(nj.handle, st, out.data(), &len, dev->stream));
        YRT_CK(cudaStreamSynchronize(dev->stream));
        FILE* f = fopen(file, "wb");
        if (!f) throw std::runtime_error(std::string("Unable to open \"") + file + "\".");
        const bool ok = fwrite(out.data(), 1, len, f) == len; fclose(f);
        if (!ok) throw std::runtime_error(std::string("Unable to write \"") + file + "\".");
    } catch (...) { nj.EncoderParamsDestroy(pr); nj.EncoderStateDestroy(st); throw; }
    nj.EncoderParamsDestroy(pr); nj.EncoderStateDestroy(st);
}

// cubeFaceIndex < 0: the whole 12W x H strip (renderer.cpp:713-718); 0..11: that face's segment — the reference's per-face debug
// image, which carries the watermark too (renderer.cpp:656-660)
void strip_encode_jpeg(yrt_device* dev, int cubeFaceIndex, int quality, const char* file) {
    CubeStrip& s = dev->strip;
    if (!s.dev) throw std::runtime_error("device_cuda: no strip");
    if (cubeFaceIndex < 0) encode_jpeg(dev, s.dev, 12 * s.w * 3, (int)(12 * s.w), (int)s.h, quality, file);
    else encode_jpeg(dev, s.dev + (size_t)strip_segment(cubeFaceIndex) * s.w * 3, 12 * s.w * 3, (int)s.w, (int)s.h, quality, file);
}

void strip_release(yrt_device* dev) {
    CubeStrip& s = dev->strip;
    if (s.dev) cudaFree(s.dev); if (s.wm) cudaFree(s.wm);
    s = CubeStrip();
}

}  // namespace yrt
