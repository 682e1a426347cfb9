"""The group device: device_cuda on N GPUs in one process behind the unchanged API (cfg "gpus=N", csrc/group_api.cu) — the in-process
counterpart of the reference's TCP device_network (devices/device_network/network_device.cpp). Needs >= 2 GPUs
(`gpurun --gpus 2 -- python -m pytest tests/test_gpu_group.py -m gpu`); skipped on a one-GPU box.

A group of N must render exactly what N separately created devices with serverID = i, serverCount = N render (the reference's own
row-band partition, api/swapchain.h:57-70, already checked against the reference's servers in test_gpu_golden.py), interleaved."""
import os
import subprocess

import numpy as np
import pytest

from tests import dae_scene, scenes

pytestmark = pytest.mark.gpu


def _gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


needs2 = pytest.mark.skipif(_gpus() < 2, reason="needs two GPUs")


def interleave(parts, height):
    n = len(parts)
    out = np.zeros((height,) + parts[0].shape[1:], parts[0].dtype)
    for i, p in enumerate(parts):
        rows = [y for y in range(height) if ((y >> 2) - i) % n == 0]
        out[rows] = p[: len(rows)]
    return out


@needs2
@pytest.mark.parametrize("fmt", ["RGB_FLOAT32", "RGB8"])
def test_group_equals_its_members(fmt):
    from yulio_raytracer_b200 import Device
    W, H, N = 48, 44, 2
    grp = Device.cuda(cfg=f"gpus={N}")
    s = scenes.atrium(grp, W, H, 8, 6, face=3, detail=4, fmt=fmt, tex_size=32)
    frames = []
    for i, _ in scenes.render_cube_map(grp, s, faces=[3, 8]):
        frames.append(grp.read_framebuffer(s.framebuffer, fmt, W, H))
    st = grp.frame_stats()
    assert st.num_gpus == N
    rays = 0
    parts = [[], []]
    for k in range(N):
        d = Device.cuda(cfg=f"gpu={k},serverID={k},serverCount={N}")
        sk = scenes.atrium(d, W, H, 8, 6, face=3, detail=4, fmt=fmt, tex_size=32)
        for j, (i, _) in enumerate(scenes.render_cube_map(d, sk, faces=[3, 8])):
            parts[j].append(d.read_framebuffer(sk.framebuffer, fmt, W, H))
        fs = d.frame_stats(); rays += fs.rays_closest + fs.rays_shadow
        d.close()
    for j in range(2):
        ref = interleave(parts[j], H)
        assert np.array_equal(frames[j].view(np.uint8), ref.view(np.uint8)), f"face {j}"
    assert st.rays_closest + st.rays_shadow == rays
    grp.close()


@needs2
def test_group_strip_and_pick(tmp_path):
    from PIL import Image
    from yulio_raytracer_b200 import Device
    W = 32
    grp = Device.cuda(cfg="gpus=2")
    s = scenes.atrium(grp, W, W, 4, 4, face=0, detail=4, fmt="RGB8", tex_size=32)
    grp.strip_begin(W, W)
    faces = []
    for i, _ in scenes.render_cube_map(grp, s):
        grp.strip_add_face(s.framebuffer, i)
        faces.append(grp.read_framebuffer(s.framebuffer, "RGB8", W, W))
    strip = grp.strip_read(W, W)
    order = [3, 1, 4, 5, 2, 0]
    ref = np.concatenate([faces[(6 if seg < 6 else 0) + order[seg % 6]] for seg in range(12)], axis=1)
    assert np.array_equal(strip, ref)
    f = str(tmp_path / "strip.jpg"); grp.strip_encode_jpeg(f, 90)
    assert np.asarray(Image.open(f)).shape == (W, 12 * W, 3)
    hit, p = grp.rtPick(s.camera, .5, .5, s.scene)
    assert hit and np.isfinite(p).all()
    grp.close()


@needs2
def test_front_end_on_two_gpus(tmp_path):
    """StartRT .. WaitRT with YULIO_RT_CFG=gpus=2: the C++ front end, the reference's loader and the group device together."""
    from PIL import Image
    from tests.test_frontend import RT_TEST, need_frontend
    need_frontend()
    d1, d2 = tmp_path / "one", tmp_path / "two"
    dae1, dae2 = dae_scene.write_scene(str(d1), "room"), dae_scene.write_scene(str(d2), "room")
    env = dict(os.environ); env.pop("YULIO_RT_DEVICE_LIB", None)
    r1 = subprocess.run([RT_TEST, dae1, "64", "64", "6"], env=env, capture_output=True, text=True, timeout=600)
    env["YULIO_RT_CFG"] = "gpus=2"
    r2 = subprocess.run([RT_TEST, dae2, "64", "64", "6"], env=env, capture_output=True, text=True, timeout=600)
    assert r1.returncode == 0 and r2.returncode == 0, r1.stdout + r1.stderr + r2.stdout + r2.stderr
    a = np.asarray(Image.open(str(d1 / "room_A.jpg")).convert("RGB")).astype(np.float64)
    b = np.asarray(Image.open(str(d2 / "room_A.jpg")).convert("RGB")).astype(np.float64)
    assert a.shape == b.shape == (64, 12 * 64, 3)
    # same scene, same spp, different sample-set assignment per pixel (the per-tile LCG is seeded with the server id): two noisy
    # estimates of the same image
    assert abs(a.mean() - b.mean()) <= 1.5 and np.abs(a - b).mean() <= 12.0, (a.mean(), b.mean(), np.abs(a - b).mean())


@needs2
def test_group_cube_map_and_device_side_assembly():
    """yrtxRenderCubeMap on a group: every member renders its bands of all faces as one wavefront; the strip takes the frames from member 0's
    assembled device copy (peer copies over NVLink) — no frame crosses to the host until a frame buffer is mapped."""
    from yulio_raytracer_b200 import Device
    W, N = 40, 2
    grp = Device.cuda(cfg=f"gpus={N}")
    s = scenes.atrium(grp, W, W, 4, 5, face=0, detail=4, fmt="RGB8", tex_size=32)
    cams = scenes.cube_cameras(grp, s)
    fbs = [grp.rtNewFrameBuffer("RGB8", W, W, 1) for _ in cams]
    scenes.render_cube_map_batched(grp, s, cams, fbs)
    grp.strip_begin(W, W)
    for i, fb in enumerate(fbs):
        grp.strip_add_face(fb, i)
    assert grp.frame_stats().d2h_bytes == 0, "a member copied its bands to the host"
    strip = grp.strip_read(W, W)
    faces = [grp.read_framebuffer(fb, "RGB8", W, W) for fb in fbs]
    order = [3, 1, 4, 5, 2, 0]
    ref = np.concatenate([faces[(6 if seg < 6 else 0) + order[seg % 6]] for seg in range(12)], axis=1)
    assert np.array_equal(strip, ref)
    # per-face loop on the same group: identical frames
    for i, _ in scenes.render_cube_map(grp, s, faces=[0, 7]):
        assert np.array_equal(grp.read_framebuffer(s.framebuffer, "RGB8", W, W), faces[i])
    grp.close()


@needs2
def test_group_ignores_member_keys_in_the_user_cfg():
    """cfg "gpus=2,gpu=0,serverID=0,serverCount=1" must still put the members on two GPUs with bands 0/2 and 1/2 (the first
    match of a key wins in the cfg parser, so the group strips the keys it assigns itself). Same frame as a plain "gpus=2" group, and the
    wavefront state (about 170 MB per member for this frame) shows up on BOTH GPUs."""
    import torch
    from yulio_raytracer_b200 import Device
    W = 256
    frames = []
    for cfg in ("gpus=2", "gpus=2,gpu=0,serverID=0,serverCount=1"):
        for i in range(2):
            torch.cuda.synchronize(i)
        free_before = [torch.cuda.mem_get_info(i)[0] for i in range(2)]
        d = Device.cuda(cfg=cfg)
        s = scenes.atrium(d, W, W, 32, 4, face=1, detail=4, fmt="RGB8", tex_size=32)
        d.rtRenderFrame(s.renderer, s.camera, s.scene, s.tonemapper, s.framebuffer, 0)
        frames.append(d.read_framebuffer(s.framebuffer, "RGB8", W, W))
        used = [free_before[i] - torch.cuda.mem_get_info(i)[0] for i in range(2)]
        assert all(u > (100 << 20) for u in used), f"{cfg}: a member is missing from one of the GPUs (bytes taken per GPU: {used})"
        d.close()
    assert np.array_equal(frames[0], frames[1])
