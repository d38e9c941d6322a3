// neeat_host.h - host half of NEE-AT's temporal feedback: LightsBaker's per-frame state (update counter, R2 tile jitter, which feedback is available) and the control values
// of a frame (Rtxpt/Lighting/LightsBaker.cpp:943-962, :985-1070; defaults LightsBaker.h:62-63, :240-255).  Plain C++ over neeat::Params; used by api.cu and the test-only host
// build of the passes (tests/emu).
#pragma once
#include "neeat.cuh"
#include <algorithm>
#include <cmath>

namespace pt { namespace neeat {

struct Settings
{
    float globalTemporalFeedbackWeight = 0.75f, localToGlobalSampleRatio = 0.65f;
    float reservoirHistoryDropoff = 0.005f, depthDisocclusionThreshold = 1.5f, screenSpaceVsWorldSpaceThreshold = 0.3f;
    bool preFilter = true, enableMotionReprojection = true;
    float importanceBoostFrustumMul = 8.0f, importanceBoostFrustumFadeDistance = 5.0f, importanceBoostIntensityDeltaMul = 64.0f;       // LightsBaker.h:245-249
};
struct HostState
{
    Settings settings;
    uint32_t W = 0, H = 0, updateCounter = 0, jitter[2] = { 0, 0 }, jitterPrev[2] = { 0, 0 }, historicLightCount = 0; float jitterF[2] = { 0, 0 };
    bool feedbackBufferFilled = false, lastFrameTemporalFeedbackAvailable = false, lastFrameLocalSamplesAvailable = false;
    void reset(uint32_t w, uint32_t h) { *this = HostState(); W = w; H = h; }
    static uint32_t tilesX(uint32_t w) { return (w + kTileSize - 1) / kTileSize + 1; }      // + 1: border for the jitter offset
    static uint32_t tilesY(uint32_t h) { return (h + kTileSize - 1) / kTileSize + 1; }
};

// LightsBaker::UpdateBegin's bookkeeping: advances the jitter and the counter, decides what of last frame is usable, fills the control part of `p` (not the pointers)
// LightsBaker::UpdateFrustumConsts (LightsBaker.cpp:884-924): left, right, top, bottom, near planes of view.matWorldToClip (row-major, row vector x matrix), normalised
inline void frustumPlanes(const float* M, float planes[5][4])
{
    const int colOf[5] = { 0, 0, 1, 1, 2 }; const float sign[5] = { 1.0f, -1.0f, -1.0f, 1.0f, -1.0f };
    for (int i = 0; i < 5; i++)
    {
        float pl[4] = { M[0 * 4 + 3] + sign[i] * M[0 * 4 + colOf[i]], M[1 * 4 + 3] + sign[i] * M[1 * 4 + colOf[i]], M[2 * 4 + 3] + sign[i] * M[2 * 4 + colOf[i]], -(M[3 * 4 + 3] + sign[i] * M[3 * 4 + colOf[i]]) };
        const float lengthSq = pl[0] * pl[0] + pl[1] * pl[1] + pl[2] * pl[2], scale = lengthSq > 0.f ? 1.0f / sqrtf(lengthSq) : 0.0f;
        for (int k = 0; k < 4; k++) planes[i][k] = pl[k] * scale;
    }
}
// boostFlags: RtxptPathTracerConstants::NEEATImportanceBoost; worldToClip: null when no view has been set (the frustum booster is then off)
inline void beginFrame(HostState& s, Params& p, uint32_t neeType, uint32_t lightCount, float weightsSum, uint32_t boostFlags = 0, const float* worldToClip = nullptr)
{
    s.jitterPrev[0] = s.jitter[0]; s.jitterPrev[1] = s.jitter[1];
    if ((s.updateCounter % 1024) == 0) { s.jitterF[0] = 0; s.jitterF[1] = 0; }
    const float g = 1.32471795724474602596f, a1 = 1.0f / g, a2 = 1.0f / (g * g);
    s.jitterF[0] = fmodf(s.jitterF[0] + a1, 1.0f); s.jitterF[1] = fmodf(s.jitterF[1] + a2, 1.0f);
    for (int k = 0; k < 2; k++) s.jitter[k] = std::min(uint32_t(s.jitterF[k] * float(kTileSize)), kTileSize - 1);
    s.updateCounter++;
    const bool lastFrameLocalSamplesAvailable = s.lastFrameTemporalFeedbackAvailable;
    const bool available = s.feedbackBufferFilled && neeType == 2;
    s.lastFrameTemporalFeedbackAvailable = available; s.lastFrameLocalSamplesAvailable = lastFrameLocalSamplesAvailable && available;
    p.W = s.W; p.H = s.H; p.blendedW = (s.W + 1) / 2; p.blendedH = (s.H + 1) / 2; p.tilesX = HostState::tilesX(s.W); p.tilesY = HostState::tilesY(s.H);
    p.lightCount = lightCount; p.historicLightCount = s.historicLightCount; p.updateCounter = s.updateCounter; p.neeType = neeType;
    p.jitterX = s.jitter[0]; p.jitterY = s.jitter[1]; p.jitterPrevX = s.jitterPrev[0]; p.jitterPrevY = s.jitterPrev[1];
    p.lastFrameFeedbackAvailable = available ? 1u : 0u; p.lastFrameLocalSamplesAvailable = s.lastFrameLocalSamplesAvailable ? 1u : 0u;
    p.temporalFeedbackRequired = neeType == 2 ? 1u : 0u; p.enableMotionReprojection = s.settings.enableMotionReprojection ? 1u : 0u;
    p.reservoirHistoryDropoff = s.settings.reservoirHistoryDropoff; p.depthDisocclusionThreshold = s.settings.depthDisocclusionThreshold;
    p.globalFeedbackUseWeight = available ? std::min(std::max(s.settings.globalTemporalFeedbackWeight, 0.0f), 0.95f) : 0.0f;
    p.localToGlobalSampleRatio = available ? std::min(std::max(s.settings.localToGlobalSampleRatio, 0.0f), 1.0f) : 0.0f;
    p.screenSpaceVsWorldSpaceThreshold = s.settings.screenSpaceVsWorldSpaceThreshold; p.weightsSum = weightsSum;
    p.boostFlags = (boostFlags & 2u) | ((boostFlags & 1u) && worldToClip && s.settings.importanceBoostFrustumMul > 0 ? 1u : 0u);
    p.boostFrustumMul = s.settings.importanceBoostFrustumMul; p.boostFrustumFadeDistance = s.settings.importanceBoostFrustumFadeDistance; p.boostIntensityDeltaMul = s.settings.importanceBoostIntensityDeltaMul;
    if (p.boostFlags & 1u) frustumPlanes(worldToClip, p.frustumPlanes);
    s.historicLightCount = lightCount;
}
// after UpdateEnd's passes: ClearFeedbackHistory ran, so the path tracer's feedback of this frame is what next frame processes
inline void endFrame(HostState& s, const Params& p) { if (p.temporalFeedbackRequired) s.feedbackBufferFilled = true; }

} } // namespace pt::neeat
