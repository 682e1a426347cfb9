"""Data formats either side of the render path.

CPU: the PNG reader behind yrtNewImageFromFile (csrc/image_codecs.cu) against PIL for every colour type / bit depth / filter it
meets, in the row order of the reference's FreeImage loader (common/image/freeimage.cpp:40-78: bottom-up) and of its watermark load
(devices/renderer/renderer.cpp:84: flipped vertically and horizontally).
GPU: JPEG decode (nvJPEG) against libjpeg-turbo (PIL) within the stated tolerance, PNG / JPEG through rtNewImageFromFile, and the
stereo cube-map strip assembled, watermarked and JPEG-encoded on the device against a host restatement of renderer.cpp:620-725."""
import os

import numpy as np
import pytest
from PIL import Image

from tests import scenes
from yulio_raytracer_b200 import devapi


def _pngs(tmp):
    rng = np.random.default_rng(11)
    a = rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)
    smooth = np.clip(np.add.outer(np.arange(64), np.arange(80)) * 2, 0, 255).astype(np.uint8)
    cases = {"rgb": Image.fromarray(a), "rgba": Image.fromarray(np.dstack([a, rng.integers(0, 256, (37, 53), dtype=np.uint8)])),
             "gray": Image.fromarray(a[..., 0]), "pal17": Image.fromarray(a).quantize(17), "pal4": Image.fromarray(a).quantize(4),
             "la": Image.fromarray(np.dstack([a[..., 0], a[..., 1]]), "LA"), "bilevel": Image.fromarray(a[..., 0] > 128),
             "smooth": Image.fromarray(np.dstack([smooth, smooth // 2, 255 - smooth]))}
    out = {}
    for k, im in cases.items():
        f = os.path.join(tmp, k + ".png"); im.save(f, optimize=(k == "smooth")); out[k] = f
    return out


def test_png_reader_matches_pil(tmp_path):
    for name, f in _pngs(str(tmp_path)).items():
        ref = np.asarray(Image.open(f).convert("RGBA"))
        assert np.array_equal(devapi.decode_png_file(devapi.CUDA_LIB, f, True, False), ref), name          # top-down
        assert np.array_equal(devapi.decode_png_file(devapi.CUDA_LIB, f, False, False), ref[::-1]), name   # FreeImage order: bottom-up
        assert np.array_equal(devapi.decode_png_file(devapi.CUDA_LIB, f, True, True), ref[:, ::-1]), name  # the watermark load


def test_png_reader_16bit_and_errors(tmp_path):
    rng = np.random.default_rng(5)
    v = rng.integers(0, 65536, (20, 30)).astype(np.uint16)
    f = str(tmp_path / "g16.png"); Image.fromarray(v).save(f)
    assert np.array_equal(devapi.decode_png_file(devapi.CUDA_LIB, f, True)[..., 0], (v >> 8).astype(np.uint8))
    bad = str(tmp_path / "bad.png"); open(bad, "wb").write(b"not a png at all")
    with pytest.raises(RuntimeError):
        devapi.decode_png_file(devapi.CUDA_LIB, bad)
    inter = str(tmp_path / "interlaced.png")
    try:
        Image.fromarray(rng.integers(0, 256, (8, 8, 3), dtype=np.uint8)).save(inter, interlace=True)   # PIL may ignore the flag
    except Exception:
        return
    raw = open(inter, "rb").read()
    if raw[28] == 1:                                          # IHDR interlace byte
        with pytest.raises(RuntimeError):
            devapi.decode_png_file(devapi.CUDA_LIB, inter)


def test_byte_colour_round_trip_is_identity():
    """Image4c::get -> Color4 -> set (common/math/color_sse.h:50-68): b * (1/255) * 255 truncated == b for every byte, which is why
    the strip assembly and the JPEG reader may copy bytes."""
    b = np.arange(256, dtype=np.float32)
    assert np.array_equal(((b * (np.float32(1) / np.float32(255))) * np.float32(255)).astype(np.int32), np.arange(256))
    assert np.array_equal(((b / np.float32(255)) * np.float32(255)).astype(np.int32), np.arange(256))


# ------------------------------------------------------------------------------------------------------------------------------
def _picture(w, h, seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([127 + 120 * np.sin(xx / 9.0 + yy / 23.0), 127 + 120 * np.cos(xx / 31.0 - yy / 7.0), (xx * 255 // max(1, w - 1))], -1)
    return np.clip(img + rng.normal(0, 6, img.shape), 0, 255).astype(np.uint8)


@pytest.mark.gpu
@pytest.mark.parametrize("subsampling,tol_mean,tol_max", [(0, 0.6, 4), (2, 1.0, 8)])
def test_jpeg_reader_against_libjpeg_turbo(cuda_dev, tmp_path, subsampling, tol_mean, tol_max):
    """nvJPEG vs libjpeg-turbo (PIL, ISLOW DCT = the reference's TJFLAG_ACCURATEDCT). JPEG decoders are not bit-identical: 4:4:4 agrees
    to the IDCT rounding; with 4:2:0 the chroma up-sampling filters differ at sharp chroma edges. Row 0 = bottom scanline, alpha 255."""
    f = str(tmp_path / f"pic{subsampling}.jpg")
    Image.fromarray(_picture(203, 117, 3)).save(f, quality=92, subsampling=subsampling)
    img = cuda_dev.rtNewImageFromFile(f)
    got = cuda_dev.read_image(img)
    ref = np.asarray(Image.open(f).convert("RGB"))[::-1]
    assert got.shape == (117, 203, 4) and (got[..., 3] == 255).all()
    d = np.abs(got[..., :3].astype(np.int32) - ref.astype(np.int32))
    print(f"subsampling {subsampling}: mean |diff| {d.mean():.3f}, max {d.max()}")
    assert d.mean() <= tol_mean and d.max() <= tol_max, (d.mean(), d.max())


@pytest.mark.gpu
def test_png_and_missing_files_through_the_device(cuda_dev, tmp_path):
    files = _pngs(str(tmp_path))
    got = cuda_dev.read_image(cuda_dev.rtNewImageFromFile(files["rgba"]))
    assert np.array_equal(got, np.asarray(Image.open(files["rgba"]).convert("RGBA"))[::-1])
    white = cuda_dev.read_image(cuda_dev.rtNewImageFromFile(str(tmp_path / "missing.jpg")))      # singleray_device.cpp:250: 1x1 white
    assert white.shape[:2] == (1, 1) and (white == 255).all()


def _psnr(a, b):
    return 10 * np.log10(255.0 ** 2 / max(1e-9, ((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2).mean()))


def _libjpeg_psnr(rgb, quality, scratch):
    """What libjpeg(-turbo) itself loses on this image with the reference's settings (jpeg_set_defaults: 4:2:0; jpeg_set_quality)."""
    Image.fromarray(np.asarray(rgb, np.uint8)).save(scratch, quality=quality, subsampling=2)
    return _psnr(np.asarray(Image.open(scratch).convert("RGB")), rgb)


def _host_strip(faces, wm=None):
    """renderer.cpp:637-711 on the host: watermark into the centre of faces 0-3 (either eye), then Left Right Up Down Back Front of
    cameras 6-11 followed by the same of cameras 0-5."""
    H, W, _ = faces[0].shape
    out = []
    for c, f in enumerate(faces):
        f = f.copy()
        if wm is not None and c % 6 < 4:
            h, w, _ = wm.shape
            x0, y0 = int(np.float32(W - w) * np.float32(.5)), int(np.float32(H - h) * np.float32(.5))
            k = np.float32(1) / np.float32(255)
            a = wm[..., 3:4].astype(np.float32) * k
            ic = f[y0:y0 + h, x0:x0 + w].astype(np.float32) * k
            bl = (np.float32(1) - a) * ic + a * (wm[..., :3].astype(np.float32) * k)
            f[y0:y0 + h, x0:x0 + w] = (np.clip(bl, 0, 1) * np.float32(255)).astype(np.uint8)
        out.append(f)
    order = [3, 1, 4, 5, 2, 0]
    return np.concatenate([out[(6 if seg < 6 else 0) + order[seg % 6]] for seg in range(12)], axis=1)


@pytest.mark.gpu
def test_cube_map_strip_on_the_device(cuda_dev, tmp_path):
    W = 64
    s = scenes.atrium(cuda_dev, W, W, 4, 4, face=0, detail=4, fmt="RGB8", tex_size=32)
    rng = np.random.default_rng(2)
    wm = rng.integers(0, 256, (20, 28, 4), dtype=np.uint8); wm[:5] = 0; wm[5:9, :, 3] = 255
    wmf = str(tmp_path / "wm.png"); Image.fromarray(wm).save(wmf)
    wm_loaded = wm[:, ::-1]                                                        # flipped vertically (top-down) and horizontally (renderer.cpp:84)
    for use_wm in (False, True):
        cuda_dev.strip_set_watermark(wmf if use_wm else None)
        cuda_dev.strip_begin(W, W)
        faces = []
        for i, _ in scenes.render_cube_map(cuda_dev, s):
            cuda_dev.strip_add_face(s.framebuffer, i, use_wm)
            faces.append(cuda_dev.read_framebuffer(s.framebuffer, "RGB8", W, W))
        strip = cuda_dev.strip_read(W, W)
        assert np.array_equal(strip, _host_strip(faces, wm_loaded if use_wm else None))
    f = str(tmp_path / "strip.jpg")
    cuda_dev.strip_encode_jpeg(f, 90)
    dec = np.asarray(Image.open(f).convert("RGB")).astype(np.float64)
    assert dec.shape == strip.shape
    assert _psnr(dec, strip) >= _libjpeg_psnr(strip, 90, str(tmp_path / "ref.jpg")) - 1.5                # quality 90, 4:2:0 (jpeg.cpp:226-229)
    ff = str(tmp_path / "face7.jpg")
    cuda_dev.strip_encode_jpeg(ff, 90, cube_face_index=7)
    assert np.asarray(Image.open(ff)).shape == (W, W, 3)
    cuda_dev.strip_set_watermark(None)
