// camera.cuh — primary-ray generation of the three camera models, shared by the ray-generation kernels and the host (rtPick).
#pragma once
#include "common.cuh"

namespace yrt {

// ---- cameras ------------------------------------------------------------------------------------
// PinHoleCamera::ray cameras/pinholecamera.h:38-40; StereoCubeCamera::ray cameras/StereoCubeCamera.h:68-161;
// DepthOfFieldCamera::ray cameras/depthoffieldcamera.h:36-42
YRT_HD void camera_ray(const CameraData& cam, float px, float py, float lx, float ly, V3& org, V3& dir) {
    if (cam.type == CAM_PINHOLE) {
        const Aff3& m = cam.p2w[0];
        org = m.p; dir = normalize(px * m.l.vx + (1.0f - py) * m.l.vy + m.l.vz);
        return;
    }
    if (cam.type == CAM_DOF) {
        const Aff3& m = cam.p2w[0];
        const float r = sqrtf(lx), theta = YRT_TWO_PI * ly;
        const V3 begin = xfmPoint(cam.local2world, V3(cam.lensRadius * r * cosf(theta), cam.lensRadius * r * sinf(theta), 0.0f));
        const V3 end = m.p + cam.focalDistance * (px * m.l.vx + (1.0f - py) * m.l.vy + m.l.vz);
        org = begin; dir = normalize(end - begin);
        return;
    }
    const Aff3& f = cam.p2w[0];
    const int face = cam.cubeFaceIndex % 6;
    const float yPixel = 1.0f - py;
    Aff3 p2w = cam.p2w[face];
    float theta = 0.f, absoluteVerticalAngle = 0.f;
    if (face < 4) {
        const V3 xDir = normalize(px * f.l.vx + .5f * f.l.vy + f.l.vz);
        theta = acosf(rclamp(dot(xDir, cam.xyzStraight), -1.f, 1.f)) * signf_(px - .5f);
        const V3 yDir = normalize(.5f * f.l.vx + yPixel * f.l.vy + f.l.vz);
        const float yAngle = rad2deg(acosf(rclamp(dot(yDir, cam.xyzStraight), -1.f, 1.f))) * signf_(yPixel - .5f);
        absoluteVerticalAngle = fabsf(yAngle);
    } else {
        const V3 xyDirNorm = normalize(V3(px - .5f, yPixel - .5f, 0.f));
        const V3 xyUp = face == 4 ? V3(0.f, -1.f, 0.f) : V3(0.f, 1.f, 0.f);
        theta = acosf(rclamp(dot(xyDirNorm, xyUp), -1.f, 1.f)) * signf_(px - .5f);
        const V3 xyzDir = normalize(px * f.l.vx + yPixel * f.l.vy + f.l.vz);
        const float xyzAngle = rad2deg(acosf(rclamp(dot(xyzDir, cam.xyzStraight), -1.f, 1.f)));
        absoluteVerticalAngle = 90.f - fabsf(xyzAngle);
    }
    float eyeOffset = cam.eyeSeparation * (cam.cubeFaceIndex < 6 ? -.5f : .5f);
    if (absoluteVerticalAngle > cam.falloffAngle)
        eyeOffset *= 1.f - smoothstepf(0.f, 1.f, smoothstepf(cam.falloffAngle, 90.f, absoluteVerticalAngle));
    p2w = mul(p2w, aff3_translate(V3(eyeOffset, 0.f, 0.f)));
    const Aff3 rot = aff3_rotate_about(cam.origin, cam.up, theta);
    const V3 rayOrigin = mul(rot, p2w).p;
    if (cam.toeIn) {
        const float corr = -atanf(eyeOffset * cam.rcpZeroParallax);
        p2w = mul(aff3_rotate_about(rayOrigin, cam.up, corr), p2w);
    }
    org = rayOrigin; dir = normalize(px * p2w.l.vx + yPixel * p2w.l.vy + p2w.l.vz);
}


}  // namespace yrt
