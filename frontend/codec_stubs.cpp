// Link-time stand-ins for the reference's binary-only image codecs (FreeImage / LodePNG glue are Windows libraries in the mount) inside
// libyulio_rt.so. The re-hosted front end never decodes or stores images itself: textures go through the device's rtNewImageFromFile
// (nvJPEG / the PNG reader of csrc/image_codecs.cu) and the strip is encoded on the GPU; the reference's loaders only need these symbols
// to link (common/image/image.cpp:26-74).
#include "image/image.h"
namespace embree {
Ref<Image> loadFreeImage(const FileName&, float, bool) { return null; }
bool storeFreeImage(const Ref<Image>&, const FileName&, int) { throw std::runtime_error("FreeImage is not part of the Linux front end"); }
void storePNG(const Ref<Image>&, const FileName&) { throw std::runtime_error("the PNG writer is not part of the Linux front end"); }
}
