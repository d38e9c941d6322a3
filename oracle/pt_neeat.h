// ORACLE — test infrastructure only (see pt_math.h).
// pt_neeat.h: NEE-AT temporal feedback: the per-pixel light feedback reservoirs the path tracer fills, their processing into (a) usage-based
// weights of the global proxy table and (b) the per-tile local samplers of the next frame, and the sampler-side functions that read them.  Restated from
//   Rtxpt/Shaders/PathTracer/Lighting/LightingTypes.hlsli:146-163 (candidate counts), :170-177 (mini-list packing, tile address), :180-291 (LightFeedbackReservoir)
//   Rtxpt/Shaders/PathTracer/Lighting/LightSampler.hlsli:45-96 (tile position, SSC heuristic), :120-180 (SampleLocal, SampleLocalPDF), :182-199 (InsertFeedbackFromNEE)
//   Rtxpt/Shaders/PathTracer/Lighting/LightingAlgorithms.hlsli:654-682 (LocalLightBinarySearch), Rtxpt/Shaders/Libraries/MicroRng.hlsli:12-60
//   Rtxpt/Lighting/LightsBaker.hlsl:774-823 (ClearFeedbackHistory), :880-948 (ComputeProxyCounts with feedback), :1068-1093 (RemapPastToCurrent), :1095-1181 (PreFilter),
//     :1186-1318 (P0), :1321-1377 (SampleLightGlobal, MirrorCoord, SampleLightLocalHistoric, Reproject), :1380-1452 (P1a), :1456-1528 (P1b), :1531-1610 (FillTile / P2),
//     :1745-1850 (P3: bitonic sort + duplicate counting)
//   Rtxpt/Lighting/LightsBaker.cpp:943-962 (R2 tile jitter), :985-1070 (control data of a frame), :1203-1225 (PreFilter, P0 order), :1331-1418 (UpdateEnd: P1a, P1b, P2, P3, clear)
//   Rtxpt/Lighting/LightsBaker.h:62-63, :240-255 (defaults), Rtxpt/Shaders/PathTracer/Lighting/LightingConfig.h (tile 8, window 8, 128 local proxies, early-feedback tile 2)
// Pinned: the sampler side by tests/golden/sampler_golden.npz; P0, P1a, P1b, P2, P3, ClearFeedbackHistory and ComputeProxyCounts by tests/golden/baker_golden.npz
// (LightSampler.hlsli / LightsBaker.hlsl compiled in place).  PreFilter, the importance boosters, the proxy fill and the light-list tracking are restatements.
// Scope: light lists may change between frames (NeeatTrackLightList builds the past <-> current index tables; an unchanged list is the identity remap with its bounds checks), importance boosters off except the pre-filter merge (default on),
// debug views omitted.  ProcessFeedbackHistoryPreFilter is racy across thread groups in the reference (a group's margin may read texels another group has already
// rewritten); here every pixel reads the reservoirs as they were before the pass.
#pragma once
#include "pt_lights.h"
#include <vector>
#include <cmath>

namespace orc {

static const uint NEEAT_TILE_SIZE = 8, NEEAT_WINDOW_SIZE = 8, NEEAT_LOCAL_PROXY_COUNT = 128, NEEAT_BINARY_SEARCH_STEPS = 8, NEEAT_EARLY_FEEDBACK_TILE_SIZE = 2;
static const uint NEEAT_TOP_UP_SAMPLES = NEEAT_LOCAL_PROXY_COUNT - NEEAT_WINDOW_SIZE * NEEAT_WINDOW_SIZE;
static const uint LFR_SCREEN_SPACE_COHERENT_FLAG = 0x80000000u;
static const float LFR_MAX_WEIGHT = 1e12f;

// MicroRng (Libraries/MicroRng.hlsli): pt_scene.h

inline uint PackMiniListLightAndCount(uint lightIndex, uint counter) { return ((lightIndex & 0x007FFFFFu) << 9) | ((counter - 1) & 0x1FFu); }
inline uint UnpackMiniListLight(uint v) { return v >> 9; }
inline uint UnpackMiniListCount(uint v) { return (v & 0x1FFu) + 1; }
inline uint ComputeCandidateSampleLocalCount(float localToGlobalRatio, uint totalCandidateSamples) { return uint(float(totalCandidateSamples - 1) * localToGlobalRatio + 0.75f); }

struct NeeatSettings
{
    float globalTemporalFeedbackWeight = 0.75f, localToGlobalSampleRatio = 0.65f;          // LightsBaker::BakeSettings
    float reservoirHistoryDropoff = 0.005f, depthDisocclusionThreshold = 1.5f, screenSpaceVsWorldSpaceThreshold = 0.3f;
    bool preFilter = true, enableMotionReprojection = true;
    // importance boosters (LightsBaker.h:245-249; both on by default in RTXPT, selected here by RtxptPathTracerConstants::NEEATImportanceBoost: bit 0 frustum, bit 1 intensity delta)
    float importanceBoostFrustumMul = 8.0f, importanceBoostFrustumFadeDistance = 5.0f, importanceBoostIntensityDeltaMul = 64.0f;
};

// one reservoir image pair (RWTexture2D<float> total weight + RWTexture2D<uint> candidate)
struct FeedbackImage
{
    uint W = 0, H = 0; std::vector<float> weight; std::vector<uint> candidate;
    void init(uint w, uint h) { W = w; H = h; weight.assign(size_t(w) * h, 0.0f); candidate.assign(size_t(w) * h, RTXPT_INVALID_LIGHT_INDEX); }
};
struct LightFeedbackReservoir
{
    FeedbackImage* img; size_t at;
    static LightFeedbackReservoir make(FeedbackImage& im, int x, int y) { LightFeedbackReservoir r; r.img = &im; r.at = size_t(y) * im.W + x; return r; }
    float GetTotalWeight() const { return img->weight[at]; }
    void SetTotalWeight(float w) { img->weight[at] = std::min(LFR_MAX_WEIGHT, w); }
    uint GetCandidateRaw() const { return img->candidate[at]; }
    void SetCandidateRaw(uint c) { img->candidate[at] = c; }
    bool IsEmpty() const { return GetTotalWeight() == 0; }
    void Clear() { SetTotalWeight(0); SetCandidateRaw(RTXPT_INVALID_LIGHT_INDEX | 0u); }      // SetCandidate( INVALID, false )
    void CloneFrom(const LightFeedbackReservoir& o, float scale) { if (o.GetTotalWeight() > 0) { SetTotalWeight(o.GetTotalWeight() * scale); SetCandidateRaw(o.GetCandidateRaw()); } else Clear(); }
    void Add(float rnd, uint candidateIndex, float candidateWeight, bool ssc)
    {
        candidateWeight = std::min(LFR_MAX_WEIGHT, candidateWeight);
        float total = GetTotalWeight(); total += candidateWeight; SetTotalWeight(total);
        const float threshold = saturate(candidateWeight / total);
        if (ssc) candidateIndex |= LFR_SCREEN_SPACE_COHERENT_FLAG;
        if (rnd < threshold) SetCandidateRaw(candidateIndex);
    }
    void Merge(float rnd, const LightFeedbackReservoir& other, float otherScale)
    {
        const float otherTotal = std::min(LFR_MAX_WEIGHT, other.GetTotalWeight() * otherScale);
        if (otherTotal > 0)
        {
            uint lightIndex = other.GetCandidateRaw();
            if (lightIndex != RTXPT_INVALID_LIGHT_INDEX) { const bool ssc = (lightIndex & LFR_SCREEN_SPACE_COHERENT_FLAG) != 0; Add(rnd, lightIndex & ~LFR_SCREEN_SPACE_COHERENT_FLAG, otherTotal, ssc); }
        }
    }
    void GetCandidate(uint& index, bool& ssc) const
    {
        index = RTXPT_INVALID_LIGHT_INDEX; ssc = false;
        if (IsEmpty()) return;
        index = GetCandidateRaw();
        if (index != RTXPT_INVALID_LIGHT_INDEX) { ssc = (index & LFR_SCREEN_SPACE_COHERENT_FLAG) != 0; index &= ~LFR_SCREEN_SPACE_COHERENT_FLAG; }
    }
};

struct NeeatState
{
    NeeatSettings settings;
    uint W = 0, H = 0, tilesX = 0, tilesY = 0;
    FeedbackImage feedback, scratch, blended; std::vector<float> historyDepth;
    std::vector<uint> localSamplingBuffer;              // tilesX * tilesY * 128 packed (light, count) tuples, sorted by light inside a tile
    std::vector<uint> feedbackCounters;                 // per light: how many reservoirs of the last frame hold it; [lightCount]: how many hold none
    // LightsBaker's frame state
    uint updateCounter = 0; float jitterF[2] = { 0, 0 }; uint jitter[2] = { 0, 0 }, jitterPrev[2] = { 0, 0 };
    bool feedbackBufferFilled = false;
    // LightingControlData of the current frame
    bool lastFrameTemporalFeedbackAvailable = false, lastFrameLocalSamplesAvailable = false, temporalFeedbackRequired = true;
    float globalFeedbackUseWeight = 0, localToGlobalSampleRatio = 0; uint historicTotalLightCount = 0, validFeedbackCount = 0;
    // boosted light weights of this frame (what ComputeWeights stores) and of the last one (u_lightWeights' historic half), their sum in ComputeWeights' order
    std::vector<float> currentWeights, historicWeights; float currentWeightsSum = 0; float frustumPlanes[5][4] = {};
    // dynamic light lists (LightsBaker.cpp:1086-1225, LightsBaker.hlsl u_historyRemapPastToCurrent / CurrentToPast): what last frame's feedback was indexed by, and this frame's
    // index tables (empty = the list did not change: identity)
    struct ListSnapshot { bool valid = false, envEnabled = false; uint analyticCount = 0, triangleCount = 0; std::vector<uint> envNodes, envLookupMap; } past;
    std::vector<uint> pastToCurrent, currentToPast;

    void init(uint w, uint h)
    {
        W = w; H = h; tilesX = (w + NEEAT_TILE_SIZE - 1) / NEEAT_TILE_SIZE + 1; tilesY = (h + NEEAT_TILE_SIZE - 1) / NEEAT_TILE_SIZE + 1;       // + 1: border for the jitter offset
        feedback.init(w, h); scratch.init(w, h); blended.init((w + 1) / 2, (h + 1) / 2); historyDepth.assign(size_t(w) * h, 0.0f);
        localSamplingBuffer.assign(size_t(tilesX) * tilesY * NEEAT_LOCAL_PROXY_COUNT, 0u);
        currentWeights.clear(); historicWeights.clear(); currentWeightsSum = 0;
        updateCounter = 0; jitterF[0] = jitterF[1] = 0; jitter[0] = jitter[1] = jitterPrev[0] = jitterPrev[1] = 0; feedbackBufferFilled = false;
        lastFrameTemporalFeedbackAvailable = lastFrameLocalSamplesAvailable = false; globalFeedbackUseWeight = localToGlobalSampleRatio = 0; historicTotalLightCount = 0;
        past = ListSnapshot(); pastToCurrent.clear(); currentToPast.clear();
    }
    uint tileBaseAddress(uint tx, uint ty) const { return (tx + ty * tilesX) * NEEAT_LOCAL_PROXY_COUNT; }     // LLSB_ComputeBaseAddress
};

// ---- sampler side (LightSampler.hlsli) ----------------------------------------------------------------------------------------------------------------------
inline uint LocalSamplingTilePos(const NeeatState& s, uint px, uint py) { return s.tileBaseAddress((px + s.jitter[0]) / NEEAT_TILE_SIZE, (py + s.jitter[1]) / NEEAT_TILE_SIZE); }
inline uint SampleLocal(const NeeatState& s, uint tileAddress, float rnd, float& pdf)
{
    const uint indexInIndex = std::min(uint(rnd * float(NEEAT_LOCAL_PROXY_COUNT)), NEEAT_LOCAL_PROXY_COUNT - 1);
    const uint v = s.localSamplingBuffer[tileAddress + indexInIndex];
    pdf = float(UnpackMiniListCount(v)) / float(NEEAT_LOCAL_PROXY_COUNT);
    return UnpackMiniListLight(v);
}
inline float SampleLocalPDF(const NeeatState& s, uint tileAddress, uint lightIndex)
{   // LocalLightBinarySearch (LightingAlgorithms.hlsli:654-682) over the sorted tile list: exactly 8 steps and NO empty-range test.  A light below every key of the tile walks
    // left for 7 steps; step 8 then reads the word just BEFORE the tile - the previous tile's last entry (and returns that tile's count if it happens to hold the light), or for
    // tile 0 address 0x7FFFFFFF, which a D3D typed buffer reads as 0 = "light 0, count 1".  Pinned by tests/golden/sampler_golden.npz; reproduced, not corrected
    uint left = tileAddress, right = tileAddress + NEEAT_LOCAL_PROXY_COUNT - 1;
    for (uint i = 0; i < NEEAT_BINARY_SEARCH_STEPS; i++)
    {
        const uint mid = (left + right) >> 1; const uint v = mid < s.localSamplingBuffer.size() ? s.localSamplingBuffer[mid] : 0u, key = UnpackMiniListLight(v);
        if (key < lightIndex) left = mid + 1;
        else if (key > lightIndex) right = mid - 1;
        else return float(UnpackMiniListCount(v)) / float(NEEAT_LOCAL_PROXY_COUNT);
    }
    return 0.0f;
}
inline void InsertFeedbackFromNEE(NeeatState& s, const LightTable& lt, uint px, uint py, bool ssc, uint lightIndex, float pixelRadianceContributionAvg, float rnd)
{
    float w = pixelRadianceContributionAvg;
    w /= powf(float(lt.proxyCounters[lightIndex]) / float(lt.samplingProxyCount), 0.65f);
    if (ssc) w *= 1.0f;                                                   // RTXPT_LIGHTING_SCREEN_SPACE_COHERENT_FEEDBACK_BIAS
    LightFeedbackReservoir::make(s.feedback, int(px), int(py)).Add(rnd, lightIndex, w, ssc);
}

// ---- baker side (LightsBaker.hlsl) ---------------------------------------------------------------------------------------------------------------------------------
inline uint RemapPastToCurrent(const NeeatState& s, uint totalLightCount, uint historic)
{
    if (historic == RTXPT_INVALID_LIGHT_INDEX) return RTXPT_INVALID_LIGHT_INDEX;
    uint idx = historic < s.historicTotalLightCount ? (s.pastToCurrent.empty() ? historic : s.pastToCurrent[historic]) : RTXPT_INVALID_LIGHT_INDEX;        // LightsBaker.hlsl:1068-1093; no table = unchanged list
    if (idx != RTXPT_INVALID_LIGHT_INDEX && idx >= totalLightCount) idx = RTXPT_INVALID_LIGHT_INDEX;
    return idx;
}
inline void UpdateLocalJitter(NeeatState& s)
{   // R2 sequence, LightsBaker.cpp:943-962
    s.jitterPrev[0] = s.jitter[0]; s.jitterPrev[1] = s.jitter[1];
    if ((s.updateCounter % 1024) == 0) { s.jitterF[0] = 0; s.jitterF[1] = 0; }
    const float g = 1.32471795724474602596f, a1 = 1.0f / g, a2 = 1.0f / (g * g);
    s.jitterF[0] = fmodf(s.jitterF[0] + a1, 1.0f); s.jitterF[1] = fmodf(s.jitterF[1] + a2, 1.0f);
    for (int k = 0; k < 2; k++) s.jitter[k] = std::min(uint(s.jitterF[k] * float(NEEAT_TILE_SIZE)), NEEAT_TILE_SIZE - 1);
}

inline void ProcessFeedbackHistoryPreFilter(NeeatState& s)
{
    const FeedbackImage src = s.feedback;                                        // see the header note on the race
    const int W = int(s.W), H = int(s.H);
    auto load = [&](int x, int y, uint& indexRaw, float& totalWeight) {
        const size_t at = size_t(std::min(std::max(y, 0), H - 1)) * W + std::min(std::max(x, 0), W - 1);
        indexRaw = src.candidate[at]; totalWeight = src.weight[at]; if (indexRaw == RTXPT_INVALID_LIGHT_INDEX) totalWeight = 0; };
    auto isSSC = [](uint indexRaw) { return indexRaw != RTXPT_INVALID_LIGHT_INDEX && (indexRaw & LFR_SCREEN_SPACE_COHERENT_FLAG) != 0; };
    for (int y = 0; y < H; y++) for (int x = 0; x < W; x++)
    {
        uint kIndex[9]; float kWeight[9], cdf[9]; uint cIndex; float cWeight; load(x, y, cIndex, cWeight);
        const bool centerIsSSC = isSSC(cIndex), centerIsNotEmpty = cIndex != RTXPT_INVALID_LIGHT_INDEX;
        float total = 0; int n = 0;
        for (int dy = -1; dy <= 1; dy++) for (int dx = -1; dx <= 1; dx++)
        {
            load(x + dx, y + dy, kIndex[n], kWeight[n]);
            float mul = (dx == 0 && dy == 0) ? 48.0f : 1.0f;
            mul *= (centerIsSSC == isSSC(kIndex[n]) && centerIsNotEmpty) ? 128.0f : 1.0f;
            total += kWeight[n] * mul; cdf[n] = total; n++;
        }
        MicroRng rng = MicroRng::make(uint(x), uint(y), s.updateCounter, 7);
        const float rnd = rng.NextFloat();
        int pick = 8;
        for (int i = 0; i < 8; i++) if (rnd < cdf[i] / total) { pick = i; break; }
        LightFeedbackReservoir r = LightFeedbackReservoir::make(s.feedback, x, y);
        r.SetCandidateRaw(kIndex[pick]); r.SetTotalWeight(kWeight[pick]);
    }
}

// remaps candidates to this frame's light list, counts how often every light was the favourite (global feedback), strips world-space-coherent candidates
inline void ProcessFeedbackHistoryP0(NeeatState& s, uint totalLightCount)
{
    s.feedbackCounters.assign(size_t(totalLightCount) + 1, 0u);
    uint valid = 0;
    for (uint y = 0; y < s.H; y++) for (uint x = 0; x < s.W; x++)
    {
        LightFeedbackReservoir r = LightFeedbackReservoir::make(s.feedback, int(x), int(y));
        uint lightIndexAll = RTXPT_INVALID_LIGHT_INDEX;
        if (!r.IsEmpty())
        {
            uint candidate; bool ssc; r.GetCandidate(candidate, ssc);
            candidate = RemapPastToCurrent(s, totalLightCount, candidate);
            lightIndexAll = candidate;
            if (!ssc) candidate = RTXPT_INVALID_LIGHT_INDEX;
            r.SetCandidateRaw(candidate | (ssc ? LFR_SCREEN_SPACE_COHERENT_FLAG : 0u));
            if (candidate == RTXPT_INVALID_LIGHT_INDEX) r.Clear();
        }
        s.feedbackCounters[std::min(lightIndexAll, totalLightCount)]++;
        if (lightIndexAll != RTXPT_INVALID_LIGHT_INDEX) valid++;
    }
    s.validFeedbackCount = valid;                   // = TotalMaxFeedbackCount - counters[ TotalLightCount ] (threads of the padded dispatch count as invalid)
}

// LightsBaker::UpdateFrustumConsts (LightsBaker.cpp:884-924): the five bounding planes of the view frustum (left, right, top, bottom, near) from view.matWorldToClip
// (row-major, row vector x matrix), normalised; dist = dot( p, plane.xyz ) - plane.w is positive inside
inline void ComputeFrustumPlanes(const float* M, float planes[5][4])
{
    auto vp = [&](int row, int col) { return M[row * 4 + col]; };
    const int colOf[5] = { 0, 0, 1, 1, 2 }; const float sign[5] = { 1.0f, -1.0f, -1.0f, 1.0f, -1.0f };
    for (int i = 0; i < 5; i++)
    {
        float pl[4] = { vp(0, 3) + sign[i] * vp(0, colOf[i]), vp(1, 3) + sign[i] * vp(1, colOf[i]), vp(2, 3) + sign[i] * vp(2, colOf[i]), -(vp(3, 3) + sign[i] * vp(3, colOf[i])) };
        const float lengthSq = pl[0] * pl[0] + pl[1] * pl[1] + pl[2] * pl[2], scale = lengthSq > 0.f ? 1.0f / sqrtf(lengthSq) : 0.0f;
        for (int k = 0; k < 4; k++) planes[i][k] = pl[k] * scale;
    }
}
// ComputeWeights (LightsBaker.hlsl:835-877) with ImportanceBooster (:118-165): lights in or near the view frustum count up to 9x (the environment 5x), lights that got
// brighter than 1.1x their last weight count 64x the increase.  Sum order: 32-light blocks, 128-block groups, groups in index order (pt_lights.h).
inline void ComputeBoostedWeights(NeeatState& s, const LightTable& lt, uint boostFlags, const float* worldToClip)
{
    const uint n = uint(lt.lights.size());
    s.historicWeights.swap(s.currentWeights); s.currentWeights.assign(n, 0.0f);
    const bool frustum = (boostFlags & 1u) != 0 && worldToClip != nullptr && s.settings.importanceBoostFrustumMul > 0;
    const bool delta = (boostFlags & 2u) != 0 && s.lastFrameTemporalFeedbackAvailable && s.settings.importanceBoostIntensityDeltaMul > 0;
    if (frustum) ComputeFrustumPlanes(worldToClip, s.frustumPlanes);
    for (uint i = 0; i < n; i++)
    {
        float w = lt.weights[i];
        if (frustum)
        {
            float boostK;
            if (LightType(lt.lights[i]) == kLightTypeEnvironmentQuad) boostK = 0.5f;
            else
            {
                float distMin = 0;
                for (int k = 0; k < 5; k++) distMin = std::min(distMin, (lt.lights[i].Center[0] * s.frustumPlanes[k][0] + lt.lights[i].Center[1] * s.frustumPlanes[k][1] + lt.lights[i].Center[2] * s.frustumPlanes[k][2]) - s.frustumPlanes[k][3]);
                boostK = saturate(1 - std::max(0.0f, -distMin) / std::max(1e-5f, s.settings.importanceBoostFrustumFadeDistance));
            }
            w *= 1 + s.settings.importanceBoostFrustumMul * boostK;
        }
        if (delta)
        {
            float historic = 0.0f;                                                                       // ImportanceBooster, LightsBaker.hlsl:139-141
            if (s.currentToPast.empty()) historic = i < s.historicWeights.size() ? s.historicWeights[i] : 0.0f;
            else { const uint h = s.currentToPast[i]; if (h != RTXPT_INVALID_LIGHT_INDEX && h < s.historicWeights.size()) historic = s.historicWeights[h]; }
            const float d = w - historic * 1.1f;
            if (d > 0) w += s.settings.importanceBoostIntensityDeltaMul * d;
        }
        s.currentWeights[i] = w;
    }
    float total = 0;
    for (uint g0 = 0; g0 < n; g0 += 32 * 128)
    {
        float groupSum = 0;
        for (uint b0 = g0; b0 < std::min(n, g0 + 32 * 128); b0 += 32) { float blockSum = 0; for (uint i = b0; i < std::min(n, b0 + 32); i++) blockSum += s.currentWeights[i]; groupSum += blockSum; }
        total += groupSum;
    }
    s.currentWeightsSum = total;
}

// ComputeProxyCounts + proxy fill with the usage feedback blended into the power-based weights
inline void RebuildGlobalProxies(const NeeatState& s, LightTable& lt, uint neeType)
{
    const uint n = uint(lt.lights.size());
    const uint budget = LIGHTING_PROXY_RATIO * std::max(n, LIGHTING_MAX_LIGHTS / 10);
    lt.proxyIndices.clear();
    for (uint i = 0; i < n; i++)
    {
        float lightWeight = s.currentWeights[i];
        if (s.lastFrameTemporalFeedbackAvailable)
        {
            const float feedbackWeight = float(s.feedbackCounters[i]) * s.currentWeightsSum / std::max(1.0f, float(s.validFeedbackCount));
            lightWeight = lerp(lightWeight, feedbackWeight, s.globalFeedbackUseWeight);
        }
        uint proxies = 0;
        if (lightWeight > 0) proxies = (neeType == 0) ? 1u : uint(ceilf((float(budget - n) * lightWeight) / s.currentWeightsSum));
        proxies = std::min(proxies, LIGHTING_MAX_PROXIES_PER_LIGHT - 1);
        lt.proxyCounters[i] = proxies;
        lt.proxyIndices.insert(lt.proxyIndices.end(), proxies, i);
    }
    lt.samplingProxyCount = uint(lt.proxyIndices.size());
}

inline uint SampleLightGlobal(const LightTable& lt, MicroRng& rng)
{
    const float rnd = rng.NextFloat(); const uint M = lt.samplingProxyCount;
    return lt.proxyIndices[std::min(uint(rnd * float(M)), M - 1)];
}
inline void MirrorCoord(int& x, int& y, int W, int H)
{
    auto m = [](int v, int maxRes) { int r = v >= 0 ? v : -v; r = r < maxRes ? r : 2 * maxRes - 2 - r; return std::min(std::max(r, 0), maxRes - 1); };
    x = m(x, W); y = m(y, H);
}
// motion: RGBA16F screen motion (pixels) of the depth buffer's frame or null; returns false when disoccluded
inline bool Reproject(const NeeatState& s, const float* depth, const uint16_t* motion, int px, int py, int& hx, int& hy)
{
    if (!s.settings.enableMotionReprojection) { hx = px; hy = py; return true; }
    float mx = 0, my = 0;
    if (motion) { const size_t at = (size_t(py) * s.W + px) * 4; mx = f16tof32(motion[at]); my = f16tof32(motion[at + 1]); }
    // ConvertMotionVectorToPixelSpace with PrevOverCurrentViewportSize = 1
    const float cx = float(px) + 0.5f, cy = float(py) + 0.5f;
    mx = (cx + mx) * 1.0f - cx; my = (cy + my) * 1.0f - cy;
    hx = int(float(px) + mx + 0.5f); hy = int(float(py) + my + 0.5f);
    bool disocclusion = false;
    if (!(hx >= 0 && hy >= 0 && hx < int(s.W) && hy < int(s.H))) disocclusion = true;
    else
    {
        const float historic = s.historyDepth[size_t(hy) * s.W + hx], current = depth[size_t(py) * s.W + px];
        disocclusion = std::max(historic / current, current / historic) > s.settings.depthDisocclusionThreshold;       // HLSL max: a NaN operand yields the other one
    }
    if (disocclusion) { hx = px; hy = py; }
    return !disocclusion;
}

// low-resolution "blended" reservoirs: every 2x2 block (+1 pixel margin) of reprojected reservoirs merged into one
inline void ProcessFeedbackHistoryP1a(NeeatState& s, const LightTable& lt, const float* depth, const uint16_t* motion)
{
    const int T = int(NEEAT_EARLY_FEEDBACK_TILE_SIZE);
    for (uint ly = 0; ly < s.blended.H; ly++) for (uint lx = 0; lx < s.blended.W; lx++)
    {
        MicroRng rng = MicroRng::make(lx, ly, s.updateCounter, 3);
        LightFeedbackReservoir out = LightFeedbackReservoir::make(s.blended, int(lx), int(ly));
        out.Clear();
        if (s.lastFrameTemporalFeedbackAvailable)
            for (int x = -1; x < T + 1; x++) for (int y = -1; y < T + 1; y++)
            {
                const int px = std::min(std::max(int(lx) * T + x, 0), int(s.W) - 1), py = std::min(std::max(int(ly) * T + y, 0), int(s.H) - 1);
                const float baseWeight = (x < 0 || y < 0 || x >= T || y >= T) ? s.settings.reservoirHistoryDropoff : 1.0f;
                int hx, hy;
                if (Reproject(s, depth, motion, px, py, hx, hy))
                {
                    const LightFeedbackReservoir src = LightFeedbackReservoir::make(s.feedback, hx, hy);
                    if (!src.IsEmpty()) out.Merge(rng.NextFloat(), src, baseWeight);
                }
            }
        if (out.GetCandidateRaw() == RTXPT_INVALID_LIGHT_INDEX) out.SetCandidateRaw(SampleLightGlobal(lt, rng));      // always a valid light, even when empty
    }
}

// full-resolution pass: reprojected reservoir + a share of the blended one; holes filled from last frame's tile or the global table
inline void ProcessFeedbackHistoryP1b(NeeatState& s, const LightTable& lt, const float* depth, const uint16_t* motion)
{
    const uint totalLightCount = uint(lt.lights.size());
    for (uint y = 0; y < s.H; y++) for (uint x = 0; x < s.W; x++)
    {
        MicroRng rng = MicroRng::make(x, y, s.updateCounter, 4);
        int hx, hy; const bool reprojectionValid = Reproject(s, depth, motion, int(x), int(y), hx, hy);
        LightFeedbackReservoir target = LightFeedbackReservoir::make(s.scratch, int(x), int(y));
        if (!s.lastFrameTemporalFeedbackAvailable) { target.Clear(); target.SetCandidateRaw(SampleLightGlobal(lt, rng)); continue; }
        target.CloneFrom(LightFeedbackReservoir::make(s.feedback, hx, hy), reprojectionValid ? 1.0f : 0.0f);
        const LightFeedbackReservoir src = LightFeedbackReservoir::make(s.blended, int(x / NEEAT_EARLY_FEEDBACK_TILE_SIZE), int(y / NEEAT_EARLY_FEEDBACK_TILE_SIZE));
        if (!src.IsEmpty()) target.Merge(rng.NextFloat(), src, s.settings.reservoirHistoryDropoff);
        uint res = target.GetCandidateRaw();
        if (res == RTXPT_INVALID_LIGHT_INDEX)
        {
            if (reprojectionValid && s.lastFrameLocalSamplesAvailable)
            {   // SampleLightLocalHistoric: a random entry of the tile the pixel belonged to last frame
                const uint tx = (uint(hx) + s.jitterPrev[0]) / NEEAT_TILE_SIZE, ty = (uint(hy) + s.jitterPrev[1]) / NEEAT_TILE_SIZE;
                const uint indexInIndex = rng.Next() % NEEAT_LOCAL_PROXY_COUNT;
                res = RemapPastToCurrent(s, totalLightCount, UnpackMiniListLight(s.localSamplingBuffer[s.tileBaseAddress(tx, ty) + indexInIndex]));
            }
            if (res == RTXPT_INVALID_LIGHT_INDEX) res = SampleLightGlobal(lt, rng);
            target.SetCandidateRaw(res);
        }
    }
}

// P2 (FillTile): the 8x8 window of full-resolution candidates + 64 top-up picks from the blended image around the tile; P3: sort by light, merge duplicates into counts
// FillTile (LightsBaker.hlsl:1531-1598): the 128 candidates of one tile in the order the reference writes them, before P3 sorts them
inline void FillTile(const NeeatState& s, uint tx, uint ty, uint list[NEEAT_LOCAL_PROXY_COUNT])
{
    const int W = int(s.W), H = int(s.H);
    {
        uint n = 0;
        const int margin = int(NEEAT_WINDOW_SIZE - NEEAT_TILE_SIZE) / 2;
        const int cellX = int(tx * NEEAT_TILE_SIZE) - int(s.jitter[0]), cellY = int(ty * NEEAT_TILE_SIZE) - int(s.jitter[1]);
        for (int x = 0; x < int(NEEAT_WINDOW_SIZE); x++) for (int y = 0; y < int(NEEAT_WINDOW_SIZE); y++)
        {
            int px = cellX - margin + x, py = cellY - margin + y; MirrorCoord(px, py, W, H);
            list[n++] = s.scratch.candidate[size_t(py) * W + px];
        }
        MicroRng rng = MicroRng::make(tx, ty, s.updateCounter, 5);
        const float centerX = float(cellX) + float(NEEAT_TILE_SIZE) * 0.5f, centerY = float(cellY) + float(NEEAT_TILE_SIZE) * 0.5f, radius = float(NEEAT_WINDOW_SIZE) * 4.0f;
        for (uint i = 0; i < NEEAT_TOP_UP_SAMPLES; i++)
        {
            const float ox = (rng.NextFloat() - 0.5f) * radius, oy = (rng.NextFloat() - 0.5f) * radius;
            int px = int(centerX + ox + 0.5f), py = int(centerY + oy + 0.5f); MirrorCoord(px, py, W, H);
            list[n++] = s.blended.candidate[size_t(py / int(NEEAT_EARLY_FEEDBACK_TILE_SIZE)) * s.blended.W + px / int(NEEAT_EARLY_FEEDBACK_TILE_SIZE)];
        }
    }
}
inline void ProcessFeedbackHistoryP2P3(NeeatState& s)
{
    for (uint ty = 0; ty < s.tilesY; ty++) for (uint tx = 0; tx < s.tilesX; tx++)
    {
        uint list[NEEAT_LOCAL_PROXY_COUNT]; FillTile(s, tx, ty, list);
        // P3: keys are the 23-bit light indices the tuples carry (PackMiniListLightAndCount masks the index); ascending sort, then every entry gets the run length of its key
        for (uint i = 0; i < NEEAT_LOCAL_PROXY_COUNT; i++) list[i] = UnpackMiniListLight(PackMiniListLightAndCount(list[i], 1));
        std::sort(list, list + NEEAT_LOCAL_PROXY_COUNT);
        const uint base = s.tileBaseAddress(tx, ty);
        for (uint i = 0; i < NEEAT_LOCAL_PROXY_COUNT;)
        {
            uint j = i; while (j < NEEAT_LOCAL_PROXY_COUNT && list[j] == list[i]) j++;
            for (uint k = i; k < j; k++) s.localSamplingBuffer[base + k] = PackMiniListLightAndCount(list[i], j - i);
            i = j;
        }
    }
}

// seeds the new frame's reservoirs with a faded copy of the processed history (own pixel + 4 neighbours) and snapshots the depth for next frame's reprojection
inline void ClearFeedbackHistory(NeeatState& s, const float* depth)
{
    const float dropOff = s.settings.reservoirHistoryDropoff;
    const int ox[4] = { -1, 1, 0, 0 }, oy[4] = { 0, 0, -1, 1 };
    for (uint y = 0; y < s.H; y++) for (uint x = 0; x < s.W; x++)
    {
        s.historyDepth[size_t(y) * s.W + x] = depth[size_t(y) * s.W + x];
        LightFeedbackReservoir r = LightFeedbackReservoir::make(s.feedback, int(x), int(y));
        if (s.lastFrameTemporalFeedbackAvailable)
        {
            r.CloneFrom(LightFeedbackReservoir::make(s.scratch, int(x), int(y)), dropOff);
            MicroRng rng = MicroRng::make(x, y, s.updateCounter, 6);
            for (int i = 0; i < 4; i++)
            {
                const int sx = std::min(std::max(int(x) + ox[i], 0), int(s.W) - 1), sy = std::min(std::max(int(y) + oy[i], 0), int(s.H) - 1);
                const LightFeedbackReservoir src = LightFeedbackReservoir::make(s.scratch, sx, sy);
                if (!src.IsEmpty()) r.Merge(rng.NextFloat(), src, dropOff * dropOff);
            }
            if (r.GetTotalWeight() < 1e-12f) r.Clear();
        }
        else r.Clear();
    }
}

// ---- one frame of LightsBaker around the path tracer -----------------------------------------------------------------------------------------------------------------------------
// UpdateBegin: before anything of the frame is traced.  Rebuilds lt's global proxy table from the base weights and last frame's feedback.
// The light list against the one last frame's feedback was indexed by: environment quad-tree nodes map through the importance-map lookups (a past node -> the current node that
// holds its corner texel, EnvLightsMapPastToCurrent LightsBaker.hlsl:515-538; a current node -> the past node at its corner, :452-463), analytic lights keep their place in the
// scene's light array (the reference matches them by a hash of the scene object), emissive triangles keep their order behind them (block offsets, :700-711)
inline void NeeatTrackLightList(NeeatState& s, const LightTable& lt)
{
    NeeatState::ListSnapshot now; now.valid = lt.lights.size() >= ENVQT_TOTAL; now.envEnabled = lt.envEnabled; now.analyticCount = lt.analyticLightCount; now.triangleCount = lt.triangleLightCount;
    now.envNodes.resize(size_t(ENVQT_TOTAL) * 2);
    for (uint i = 0; i < ENVQT_TOTAL && now.valid; i++) { now.envNodes[2 * i] = lt.lights[i].Direction1; now.envNodes[2 * i + 1] = lt.lights[i].Direction2; }
    now.envLookupMap = lt.envLookupMap;
    const NeeatState::ListSnapshot& past = s.past;
    s.pastToCurrent.clear(); s.currentToPast.clear();
    const bool changed = past.valid && (past.envEnabled != now.envEnabled || past.analyticCount != now.analyticCount || past.triangleCount != now.triangleCount || past.envNodes != now.envNodes);
    if (changed && s.feedbackBufferFilled)
    {
        const uint E = ENVQT_TOTAL, nPast = past.analyticCount, nCur = now.analyticCount, tPast = past.triangleCount, tCur = now.triangleCount, dimMap = 1024;      // EMISB_IMPORTANCE_MAP_DIM
        s.pastToCurrent.assign(size_t(E) + nPast + tPast, RTXPT_INVALID_LIGHT_INDEX); s.currentToPast.assign(size_t(E) + nCur + tCur, RTXPT_INVALID_LIGHT_INDEX);
        if (past.envEnabled && now.envEnabled && past.envLookupMap.size() == size_t(dimMap) * dimMap && now.envLookupMap.size() == past.envLookupMap.size())
            for (uint i = 0; i < E; i++)
            {
                uint dim = past.envNodes[2 * i + 1] >> 16, x = past.envNodes[2 * i] >> 16, y = past.envNodes[2 * i] & 0xFFFFu;
                if (dim) { const uint ds = dimMap / dim; s.pastToCurrent[i] = now.envLookupMap[size_t(y * ds) * dimMap + x * ds]; }
                dim = now.envNodes[2 * i + 1] >> 16; x = now.envNodes[2 * i] >> 16; y = now.envNodes[2 * i] & 0xFFFFu;
                if (dim) { const uint ds = dimMap / dim; s.currentToPast[i] = past.envLookupMap[size_t(y * ds) * dimMap + x * ds]; }
            }
        for (uint k = 0; k < nPast; k++) s.pastToCurrent[E + k] = k < nCur ? E + k : RTXPT_INVALID_LIGHT_INDEX;
        for (uint k = 0; k < nCur; k++) s.currentToPast[E + k] = k < nPast ? E + k : RTXPT_INVALID_LIGHT_INDEX;
        for (uint t = 0; t < tPast; t++) s.pastToCurrent[E + nPast + t] = t < tCur ? E + nCur + t : RTXPT_INVALID_LIGHT_INDEX;
        for (uint t = 0; t < tCur; t++) s.currentToPast[E + nCur + t] = t < tPast ? E + nPast + t : RTXPT_INVALID_LIGHT_INDEX;
    }
    s.past = std::move(now);
}
inline void NeeatUpdateBegin(NeeatState& s, LightTable& lt, uint neeType, uint boostFlags = 0, const float* worldToClip = nullptr)
{
    NeeatTrackLightList(s, lt);
    UpdateLocalJitter(s);
    s.updateCounter++;
    const bool lastFrameLocalSamplesAvailable = s.lastFrameTemporalFeedbackAvailable;        // last frame's control data
    const bool available = s.feedbackBufferFilled && neeType == 2;
    s.lastFrameTemporalFeedbackAvailable = available;
    s.lastFrameLocalSamplesAvailable = lastFrameLocalSamplesAvailable && available;
    s.globalFeedbackUseWeight = available ? std::min(std::max(s.settings.globalTemporalFeedbackWeight, 0.0f), 0.95f) : 0.0f;
    s.localToGlobalSampleRatio = available ? std::min(std::max(s.settings.localToGlobalSampleRatio, 0.0f), 1.0f) : 0.0f;
    s.temporalFeedbackRequired = neeType == 2;
    const uint n = uint(lt.lights.size());
    s.feedbackCounters.assign(size_t(n) + 1, 0u);
    if (available)
    {
        if (s.settings.preFilter) ProcessFeedbackHistoryPreFilter(s);
        ProcessFeedbackHistoryP0(s, n);
    }
    ComputeBoostedWeights(s, lt, boostFlags, worldToClip);
    RebuildGlobalProxies(s, lt, neeType);
    s.historicTotalLightCount = n;                   // next frame's HistoricTotalLightCount
}
// UpdateEnd: after the BUILD pass in realtime mode (depth / motion of this frame), before the radiance pass of the frame
inline void NeeatUpdateEnd(NeeatState& s, const LightTable& lt, const float* depth, const uint16_t* motion)
{
    ProcessFeedbackHistoryP1a(s, lt, depth, motion);
    ProcessFeedbackHistoryP1b(s, lt, depth, motion);
    ProcessFeedbackHistoryP2P3(s);
    if (s.temporalFeedbackRequired) { ClearFeedbackHistory(s, depth); s.feedbackBufferFilled = true; }
}

} // namespace orc
