// TEST INFRASTRUCTURE ONLY (oracle). Link-time stand-ins for the reference's binary-only
// image codecs (FreeImage / libjpeg-turbo / LodePNG glue, Windows libs in the mount): the parity
// tests hand decoded pixels to rtNewImage, so the oracle never decodes files other than PPM/PFM.
#include "image/image.h"
namespace embree {
Ref<Image> loadFreeImage(const FileName&, float, bool) { return null; }
bool storeFreeImage(const Ref<Image>&, const FileName&, int) { throw std::runtime_error("FreeImage not available in the oracle build"); }
void storePNG(const Ref<Image>&, const FileName&) { throw std::runtime_error("PNG writer not available in the oracle build"); }
}
