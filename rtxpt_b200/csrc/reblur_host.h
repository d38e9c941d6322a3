// reblur_host.h - host half of the ReBLUR denoiser: nrd::ReblurSettings as RTXPT sets them (Rtxpt/NRD/NrdConfig.cpp:49-61), nrd::CommonSettings (NrdIntegration.cpp:375-408) and the
// per-frame constants NRD derives from the matrices (External/Nrd/Source/InstanceImpl.cpp:331-451, Reblur.cpp:280-392): camera-relative matrices, frustum, unproject, kernel
// rotators.  Plain C++ on top of rb::Params; used by api.cu and by the test-only host build of the passes (tests/emu/reblur_host_emu.cu).
#pragma once
#include "reblur.cuh"
#include "../../include/rtxpt_b200.h"
#include <algorithm>
#include <cmath>
#include <cstring>

namespace pt { namespace rb {

inline Rotator rbRotator(float angle) { const float ca = cosf(angle), sa = sinf(angle); return { ca, sa, -sa, ca }; }
inline Rotator rbCombine(Rotator a, Rotator b) { return { a.x * b.x + a.z * b.y, a.y * b.x + a.w * b.y, a.x * b.z + a.z * b.w, a.y * b.z + a.w * b.w }; }
inline float rbWeyl(float p, uint32_t n) { const float v = p + float(n) * 0.6180339887498948f; return v - floorf(v); }
inline void rbFrustum(const float* viewToClip, float* f) { const float P00 = viewToClip[0], P11 = viewToClip[5], P20 = viewToClip[8], P21 = viewToClip[9]; f[0] = (-1.0f - P20) / P00; f[2] = 2.0f / P00; f[1] = (1.0f - P21) / P11; f[3] = -2.0f / P11; }
inline void rbCameraPosition(const float* wv, float* o) { for (int k = 0; k < 3; k++) o[k] = -(wv[12] * wv[k * 4] + wv[13] * wv[k * 4 + 1] + wv[14] * wv[k * 4 + 2]); }
// world (relative to the current camera) -> clip of the camera that sits at camRel; optionally the 3x4 world -> view of that camera
inline void rbRelativeWorldToClip(const float* wv, const float* camRel, const float* vc, float* outM, float* outAffine)
{
    float R[9]; for (int r = 0; r < 3; r++) for (int k = 0; k < 3; k++) R[r * 3 + k] = wv[k * 4 + r];
    const float t[3] = { -(R[0] * camRel[0] + R[1] * camRel[1] + R[2] * camRel[2]), -(R[3] * camRel[0] + R[4] * camRel[1] + R[5] * camRel[2]), -(R[6] * camRel[0] + R[7] * camRel[1] + R[8] * camRel[2]) };
    if (outAffine) for (int r = 0; r < 3; r++) { outAffine[r * 4] = R[r * 3]; outAffine[r * 4 + 1] = R[r * 3 + 1]; outAffine[r * 4 + 2] = R[r * 3 + 2]; outAffine[r * 4 + 3] = t[r]; }
    const float wvRel[16] = { R[0], R[3], R[6], 0, R[1], R[4], R[7], 0, R[2], R[5], R[8], 0, t[0], t[1], t[2], 1 };
    for (int r = 0; r < 4; r++) for (int k = 0; k < 4; k++) { float a = 0; for (int j = 0; j < 4; j++) a += wvRel[r * 4 + j] * vc[j * 4 + k]; outM[r * 4 + k] = a; }
}


// everything of rb::Params except the resource pointers; historyValid = this plane's instance has denoised a frame since its creation / last resize
inline void fillFrameParams(Params& p, uint32_t W, uint32_t H, const RtxptReblurFrame* f, bool historyValid)
{
    p.W = W; p.H = H; p.tilesW = (W + 15) / 16; p.frameIndex = f->frameIndex;
    for (int r = 0; r < 3; r++) for (int k = 0; k < 3; k++) p.viewToWorld[r * 3 + k] = f->matWorldToView[k * 4 + r];
    memcpy(p.viewToClip, f->matViewToClip, 64);
    rbFrustum(f->matViewToClip, p.frustum); rbFrustum(f->prevMatViewToClip, p.frustumPrev);
    float pos[3], posPrev[3]; rbCameraPosition(f->matWorldToView, pos); rbCameraPosition(f->prevMatWorldToView, posPrev);
    for (int k = 0; k < 3; k++) p.cameraDelta[k] = posPrev[k] - pos[k];
    const float zero[3] = { 0, 0, 0 };
    rbRelativeWorldToClip(f->matWorldToView, zero, f->matViewToClip, p.worldToClip, nullptr);
    rbRelativeWorldToClip(f->prevMatWorldToView, p.cameraDelta, f->prevMatViewToClip, p.worldToClipPrev, p.worldToViewPrev);
    p.unproject = 1.0f / (0.5f * float(H) * f->matViewToClip[5]); p.minRectDimMulUnproject = float(std::min(W, H)) * p.unproject;
    const float rad90 = 1.5707963267948966f, rad360 = 6.283185307179586f;
    p.rotatorPre = rbRotator(rbWeyl(0.5f, f->frameIndex) * rad90);
    p.rotator = rbCombine(rbRotator(rbWeyl(0.0f, f->frameIndex * 2) * rad90), rbRotator(float((f->frameIndex * 2) & 15u) / 16.0f * rad360));
    p.rotatorPost = rbCombine(rbRotator(rbWeyl(0.0f, f->frameIndex * 2 + 1) * rad90), rbRotator(float((f->frameIndex * 2 + 1) & 15u) / 16.0f * rad360));
    // settings: RTXPT's ReBLUR configuration
    const bool reset = f->resetHistory != 0 || !historyValid;
    const float hitDist[4] = { 3.0f, 0.1f, 20.0f, -25.0f }; memcpy(p.hitDistParams, hitDist, 16);
    p.maxAccumulatedFrameNum = reset ? 0.0f : 50.0f; p.maxFastAccumulatedFrameNum = reset ? 0.0f : 6.0f; p.historyFixFrameNum = 3.0f; p.historyFixBasePixelStride = 14.0f;
    p.diffPrepassBlurRadius = 15.0f; p.specPrepassBlurRadius = 40.0f; p.minBlurRadius = 1.0f; p.maxBlurRadius = 30.0f; p.lobeAngleFraction = 0.15f * 0.15f; p.roughnessFraction = 0.15f;
    p.planeDistSensitivity = 0.02f; p.minHitDistanceWeight = 0.1f; p.minMaterialDiff = 4.0f; p.minMaterialSpec = 4.0f; p.denoisingRange = 100000.0f; p.viewZScale = 1.0f;
    const float thresholdBonus = 1.0f / float(H);           // ( 1 + jitterDelta ) / rectH, Reblur.cpp:293; un-jittered matrices
    p.disocclusionThreshold = (f->disocclusionThreshold > 0.0f ? f->disocclusionThreshold : 0.03f) + thresholdBonus;
    p.disocclusionThresholdAlternate = (f->disocclusionThresholdAlternate > 0.0f ? f->disocclusionThresholdAlternate : 0.2f) + thresholdBonus;
    p.framerateScale = f->frameTimeMs > 0.0f ? std::max(33.333f / f->frameTimeMs, 1.0f) : 2.0f;          // max( 33.333 ms / frame time, 1 ), Reblur.cpp:296
    p.fireflySuppressorMinRelativeScale = 2.0f; p.responsiveAccumulationRoughnessThreshold = 0.0f;
    p.antilagSigmaScale = 4.0f; p.antilagSensitivity = 3.0f; p.stabilizationStrength = reset ? 0.0f : 63.0f / 64.0f;
    p.antiFirefly = 1; p.usePrepassOnlyForSpecularMotionEstimation = 0;
}

} } // namespace pt::rb
