// host_helpers.cpp — the two small host-side callers of the PathTrace boundary that every C/C++ host needs and that live in the reference's
// C++ headers / host code: the camera bridge (BridgeCamera, Rtxpt/Shaders/PathTracer/PathTracerShared.h:109-141) and the reference-mode
// defaults of the per-frame constants (Sample::UpdatePathTracerConstants, Rtxpt/Sample.cpp:1464-1556 with the SampleUI.h:150-220 defaults).
// Written against the C ABI only (include/rtxpt_b200.h); needs no CUDA device.
#include <cmath>
#include <cstring>
#include "../../include/rtxpt_b200.h"

namespace {
struct V3 { float x, y, z; };
inline V3 cross(V3 a, V3 b) { return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; }
inline V3 scale(V3 a, float s) { return { a.x * s, a.y * s, a.z * s }; }
inline V3 normalize(V3 a) { const float l = std::sqrt((a.x * a.x + a.y * a.y) + a.z * a.z); return { a.x / l, a.y / l, a.z / l }; }
inline void store(float* d, V3 v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; }
}

extern "C" {

RTXPT_API int rtxpt_b200_bridge_camera(uint32_t viewportWidth, uint32_t viewportHeight, const float camPos[3], const float camDir[3], const float camUp[3],
                                       float fovY, float nearZ, float farZ, float focalDistance, float apertureRadius, const float jitter[2], RtxptCameraData* out)
{
    if (!camPos || !camDir || !camUp || !out || !viewportWidth || !viewportHeight) return RTXPT_ERR_INVALID_ARGUMENT;
    memset(out, 0, sizeof(*out));
    const float aspectRatio = float(viewportWidth) / float(viewportHeight);
    out->FocalDistance = focalDistance; out->NearZ = nearZ; out->FarZ = farZ; out->AspectRatio = aspectRatio;
    out->PosW[0] = camPos[0]; out->PosW[1] = camPos[1]; out->PosW[2] = camPos[2];
    out->ViewportSize[0] = viewportWidth; out->ViewportSize[1] = viewportHeight;
    const V3 dir = normalize(V3{ camDir[0], camDir[1], camDir[2] }), up = { camUp[0], camUp[1], camUp[2] };
    const V3 W = scale(dir, focalDistance);
    V3 U = normalize(cross(W, up));
    V3 V = normalize(cross(U, W));
    const float ulen = focalDistance * std::tan(fovY * 0.5f) * aspectRatio, vlen = focalDistance * std::tan(fovY * 0.5f);
    store(out->DirectionW, dir); store(out->CameraW, W); store(out->CameraU, scale(U, ulen)); store(out->CameraV, scale(V, vlen));
    out->ApertureRadius = apertureRadius;
    out->PixelConeSpreadAngle = std::atan(2.0f * std::tan(fovY * 0.5f) / float(viewportHeight));     // the whole (not half) cone angle
    out->Jitter[0] = jitter ? jitter[0] : 0.0f; out->Jitter[1] = jitter ? -jitter[1] : -0.0f;
    return RTXPT_OK;
}

RTXPT_API int rtxpt_b200_default_constants(const RtxptCameraData* camera, int envMapPresent, RtxptPathTracerConstants* out)
{
    if (!camera || !out) return RTXPT_ERR_INVALID_ARGUMENT;
    memset(out, 0, sizeof(*out));
    out->imageWidth = camera->ViewportSize[0]; out->imageHeight = camera->ViewportSize[1];
    out->sampleBaseIndex = 0;
    out->perPixelJitterAAScale = 1.0f;                      // reference AA mode (Sample.cpp:1501)
    out->bounceCount = 20; out->diffuseBounceCount = 2;     // SampleUI.h:170-171
    out->EnvironmentMapDiffuseSampleMIPLevel = 2.0f;        // SampleUI.h:109 (set by the scene in the reference; 2 is its working default)
    out->texLODBias = -1.0f;                                // SampleUI.h:180
    out->fireflyFilterThreshold = 5.0f * 1e3f;              // ReferenceFireflyFilterThreshold 5 x sqrt(preExposedGray = 1) x 1e3 (Sample.cpp:1522, SampleUI.h:213)
    out->NEEEnabled = 1; out->NEEType = 2; out->NEECandidateSamples = 5; out->NEEFullSamples = 1;       // SampleUI.h:153-155
    out->enableRussianRoulette = 1; out->enableLDSamplerForBSDF = 1; out->nestedDielectricsQuality = 1; // SampleUI.h:220,183,181
    out->camera = *camera;
    static const float ident[12] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0 };
    memcpy(out->envMap.Transform, ident, sizeof(ident)); memcpy(out->envMap.InvTransform, ident, sizeof(ident));
    out->envMap.ColorMultiplier[0] = out->envMap.ColorMultiplier[1] = out->envMap.ColorMultiplier[2] = 1.0f;
    out->envMap.Enabled = envMapPresent ? 1.0f : 0.0f;
    out->distantVsLocalImportance = 1.0f;                   // NEEAT_Distant_vs_Local_Importance (SampleUI.h:161)
    return RTXPT_OK;
}

} // extern "C"

// ---- GenericTS addressing of the stable-plane buffer (Rtxpt/Shaders/PathTracer/Utils/Utils.hlsli:262-269 Morton16BitEncode, :320-352): 8x8 tiles,
// Morton order inside a tile, tiles row-major, planes one after the other ------------------------------------------------------------------------
static uint32_t morton16(uint32_t x, uint32_t y)
{
    uint32_t t = (x & 0xff) | ((y & 0xff) << 16);
    t = (t ^ (t << 4)) & 0x0f0f0f0f; t = (t ^ (t << 2)) & 0x33333333; t = (t ^ (t << 1)) & 0x55555555;
    return ((t >> 15) | t) & 0xffff;
}
extern "C" RTXPT_API uint32_t rtxpt_b200_generic_ts_line_stride(uint32_t width, uint32_t) { return ((width + 7u) / 8u) * 8u; }
extern "C" RTXPT_API uint32_t rtxpt_b200_generic_ts_plane_stride(uint32_t width, uint32_t height) { return rtxpt_b200_generic_ts_line_stride(width, height) * ((height + 7u) / 8u) * 8u; }
extern "C" RTXPT_API uint32_t rtxpt_b200_generic_ts_address(uint32_t x, uint32_t y, uint32_t plane, uint32_t lineStride, uint32_t planeStride)
{
    const uint32_t xi = x % 8u, yi = y % 8u;
    return (x - xi) * 8u + (y - yi) * lineStride + morton16(xi, yi) + plane * planeStride;
}

// planar view of a BridgeCamera block: view axes = normalised CameraU / V / W, field of view from their lengths, reverse-free D3D projection (z in [0, 1]); computed in double,
// stored as float - the arithmetic of rtxpt_b200/scene_builder.py (world_to_view, view_to_clip, world_to_clip), which the parity tests feed to both sides
extern "C" RTXPT_API int rtxpt_b200_camera_matrices(const RtxptCameraData* cam, float* outWorldToView, float* outViewToClip, float* outWorldToClip)
{
    if (!cam) return RTXPT_ERR_INVALID_ARGUMENT;
    auto len = [](const float* v) { return std::sqrt(double(v[0]) * v[0] + double(v[1]) * v[1] + double(v[2]) * v[2]); };
    const double lu = len(cam->CameraU), lv = len(cam->CameraV), lw = len(cam->CameraW);
    double right[3], up[3], fwd[3];
    for (int k = 0; k < 3; k++) { right[k] = double(float(cam->CameraU[k] / float(lu))); up[k] = double(float(cam->CameraV[k] / float(lv))); fwd[k] = double(float(cam->CameraW[k] / float(lw))); }
    auto dotp = [&](const double* a) { return double(float(cam->PosW[0])) * a[0] + double(float(cam->PosW[1])) * a[1] + double(float(cam->PosW[2])) * a[2]; };
    double view[16] = { right[0], up[0], fwd[0], 0, right[1], up[1], fwd[1], 0, right[2], up[2], fwd[2], 0, -dotp(right), -dotp(up), -dotp(fwd), 1 };
    const double tanX = double(float(float(lu) / float(lw))), tanY = double(float(float(lv) / float(lw))), n = cam->NearZ, f = cam->FarZ;
    double proj[16] = { 1.0 / tanX, 0, 0, 0, 0, 1.0 / tanY, 0, 0, 0, 0, f / (f - n), 1.0, 0, 0, -n * f / (f - n), 0 };
    if (outWorldToView) for (int i = 0; i < 16; i++) outWorldToView[i] = float(view[i]);
    if (outViewToClip) for (int i = 0; i < 16; i++) outViewToClip[i] = float(proj[i]);
    if (outWorldToClip) for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) { double a = 0; for (int k = 0; k < 4; k++) a += view[r * 4 + k] * proj[k * 4 + c]; outWorldToClip[r * 4 + c] = float(a); }
    return RTXPT_OK;
}
