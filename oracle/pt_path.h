// ORACLE — test infrastructure only (see pt_math.h).
// pt_path.h: one reference-mode path per pixel, restated from
//   Rtxpt/Shaders/PathTracerSample.hlsl:115-166 (nextHit), :201-256 (raygen loop)
//   Rtxpt/Shaders/PathTracer/PathTracer.hlsli:40-45, :47-91, :139-175, :182-208, :217-380, :382-404, :407-503, :505-762
//   Rtxpt/Shaders/PathTracer/PathTracerNEE.hlsli:41-346
//   Rtxpt/Shaders/PathTracer/PathState.hlsli:83-268 (fp16 payload packing), PathTracerTypes.hlsli:97-200
//   Rtxpt/Shaders/PathTracer/PathTracerHelpers.hlsli:126-153 (thin lens), :195-219 (firefly filter), :29-42
//   Rtxpt/Shaders/PathTracer/PathTracerNestedDielectrics.hlsli:24-131, Rendering/Materials/InteriorList.hlsli:28-246
//   Rtxpt/Shaders/PathTracerBridgeDonut.hlsli:543-564 (camera ray)
//   Rtxpt/Shaders/PathTracer/Lighting/LightSampler.hlsli (global sampling + MIS), Lighting/EnvMap.hlsli:84-87
// Pinned: HandleHitSurface, HandleMiss, EmptyPathInitialize, FirstHitFromVBuffer and postProcessHit reproduce tests/golden/hit_golden.npz - whole calls of the
// reference's PathTracer::HandleHit / HandleMiss and driver steps compiled in place for the reference, BUILD and FILL passes - bit for bit.  The hooks of PathTracerCtx
// (visibilityOverride, cameraRayOverride, envEvalOverride) exist for that comparison: they stand where the golden's stub bridge supplies data.
#pragma once
#include "pt_scene.h"
#include "pt_bvh.h"
#include "pt_lights.h"
#include "pt_neeat.h"
#include "pt_rng.h"
#include "pt_stable_planes.h"

namespace orc {

enum PathFlags : uint {
    PF_active = 1u << 0, PF_hit = 1u << 1, PF_transmission = 1u << 2, PF_specular = 1u << 3, PF_delta = 1u << 4,
    PF_insideDielectricVolume = 1u << 5, PF_terminateAtNextBounce = 1u << 6, PF_enableThreadReorder = 1u << 9,
    PF_deltaTransmissionPath = 1u << 11, PF_deltaOnlyPath = 1u << 12,
    // realtime mode (PathState.hlsli:58-64); bits 14-15 hold the stable plane index
    PF_stablePlaneOnPlane = 1u << 16, PF_stablePlaneOnBranch = 1u << 17, PF_stablePlaneBaseScatterDiff = 1u << 18, PF_exportSpecHitTQueued = 1u << 19,
    PF_stablePlaneOnDominantBranch = 1u << 20
};
enum PathTracerMode : uint { MODE_REFERENCE = 0, MODE_BUILD_STABLE_PLANES = 1, MODE_FILL_STABLE_PLANES = 2 };      // Config.h:56-59
static const uint kVertexIndexBitCount = 10, kVertexIndexBitMask = (1u << kVertexIndexBitCount) - 1u;
static const uint kStablePlaneIndexBitOffset = 14 + kVertexIndexBitCount, kStablePlaneIndexBitMask = 3u << kStablePlaneIndexBitOffset;
enum PackedCounter { CTR_DiffuseBounces = 0, CTR_RejectedHits = 1, CTR_BouncesFromStablePlane = 2 };

struct InteriorList     // InteriorList.hlsli (2 slots)
{
    uint slots[2] = { 0, 0 };
    static const uint kNoMaterial = 0xffffffffu, kMaterialMask = (1u << 28) - 1u, kMaxNestedPriority = 15;
    bool isEmpty() const { return slots[0] == 0; }
    uint getTopNestedPriority() const { return slots[0] >> 28; }
    uint getTopMaterialID() const { return slots[0] != 0 ? (slots[0] & kMaterialMask) : kNoMaterial; }
    uint getNextMaterialID() const { return slots[1] != 0 ? (slots[1] & kMaterialMask) : kNoMaterial; }
    bool isTrueIntersection(uint nestedPriority) const { return nestedPriority == 0 || nestedPriority >= getTopNestedPriority(); }
    void handleIntersection(uint materialID, uint nestedPriority, bool entering)
    {
        if (nestedPriority == 0) nestedPriority = kMaxNestedPriority;
        uint slot = (nestedPriority << 28) | (materialID & kMaterialMask);
        if (entering && slots[0] == 0) slots[0] = slot;
        else if (!entering && slots[0] != 0 && (slots[0] & kMaterialMask) == materialID) slots[0] = 0;
        else if (entering && slots[1] == 0) slots[1] = slot;
        else if (!entering && slots[1] != 0 && (slots[1] & kMaterialMask) == materialID) slots[1] = 0;
        if (slots[0] < slots[1]) std::swap(slots[0], slots[1]);
    }
};

struct NEEBSDFMISInfo   // PathTracerTypes.hlsli:97-160 (PT_USE_RESTIR_DI = 0)
{
    bool LightSamplingEnabled = false, LightSamplingIsSSC = false; uint CandidateSamples = 0, FullSamples = 0;
    static NEEBSDFMISInfo Unpack16bit(uint p) { NEEBSDFMISInfo r; r.LightSamplingEnabled = (p & (1 << 15)) != 0; r.LightSamplingIsSSC = (p & (1 << 13)) != 0; r.CandidateSamples = (p >> 6) & 0x3F; r.FullSamples = p & 0x3F; return r; }
    uint Pack16bit() const { return ((LightSamplingEnabled ? 1u : 0u) << 15) | ((LightSamplingIsSSC ? 1u : 0u) << 13) | ((CandidateSamples & 0x3F) << 6) | (FullSamples & 0x3F); }
};

struct PathState
{
    float3 origin = f3(0), dir = f3(0); uint id = 0; float sceneLength = 0;
    uint pack23[2] = { 0, 0 };      // thp (fp16 x4)
    uint pack45[2] = { 0, 0 };      // L   (fp16 x4)
    InteriorList interiorList;
    uint packedCounters = 0;
    RayCone rayCone;
    uint pack0 = 0, pack1 = 0, flagsAndVertexIndex = 0;
    uint stableBranchID = 0;        // PathState.hlsli:102; in the BUILD pass pack45 holds imageXformPacked and pack0 the motion-vector scene length (:91-92, :151-154)

    mat3 GetImageXform() const { return UnpackOrthoMatrix(pack45); }
    void SetImageXform(const mat3& m) { PackOrthoMatrix(m, pack45); }
    void SetMotionVectorSceneLength(float l) { pack0 = asuint(l); }
    float GetMotionVectorSceneLength() const { return asfloat(pack0); }
    uint getStablePlaneIndex() const { return (flagsAndVertexIndex & kStablePlaneIndexBitMask) >> kStablePlaneIndexBitOffset; }
    void setStablePlaneIndex(uint index) { flagsAndVertexIndex &= ~kStablePlaneIndexBitMask; flagsAndVertexIndex |= index << kStablePlaneIndexBitOffset; }
    void setCounter(uint type, uint v) { const uint shift = type << 3; packedCounters = (packedCounters & ~(0xffu << shift)) | ((v & 0xff) << shift); }
    void setVertexIndex(uint index) { flagsAndVertexIndex &= ~kVertexIndexBitMask; flagsAndVertexIndex |= index; }

    void SetThp(float3 t) { t = clamp3(t, 0, HLF_MAX); pack23[0] = Fp32ToFp16NoClamp(f2(t.x, t.y)); pack23[1] = Fp32ToFp16NoClamp(f2(t.z, 0)); }
    float3 GetThp() const { float2 a = Fp16ToFp32(pack23[0]), b = Fp16ToFp32(pack23[1]); return f3(a.x, a.y, b.x); }
    void SetL(float4 l) { l = f4(clampf(l.x, 0, HLF_MAX), clampf(l.y, 0, HLF_MAX), clampf(l.z, 0, HLF_MAX), clampf(l.w, 0, HLF_MAX)); pack45[0] = Fp32ToFp16NoClamp(f2(l.x, l.y)); pack45[1] = Fp32ToFp16NoClamp(f2(l.z, l.w)); }
    float4 GetL() const { float2 a = Fp16ToFp32(pack45[0]), b = Fp16ToFp32(pack45[1]); return f4(a.x, a.y, b.x, b.y); }
    void SetFireflyFilterK_BsdfScatterPdf(float k, float pdf) { pack0 = (f32tof16(clampf(k, 0, HLF_MAX)) << 16) | f32tof16(clampf(pdf, 0, HLF_MAX)); }
    float GetFireflyFilterK() const { return f16tof32(pack0 >> 16); }
    float GetBsdfScatterPdf() const { return f16tof32(pack0 & 0xFFFF); }
    void SetPackedMISInfo_ThpRuRuCorrection(uint mis, float c) { pack1 = (mis << 16) | f32tof16(clampf(c, 0, HLF_MAX)); }
    uint GetPackedMISInfo() const { return pack1 >> 16; }
    float GetThpRuRuCorrection() const { return f16tof32(pack1 & 0xFFFF); }

    bool hasFlag(uint f) const { return (flagsAndVertexIndex & (f << kVertexIndexBitCount)) != 0; }
    void setFlag(uint f, bool v = true) { uint bit = f << kVertexIndexBitCount; if (v) flagsAndVertexIndex |= bit; else flagsAndVertexIndex &= ~bit; }
    bool isActive() const { return hasFlag(PF_active); }
    void terminate() { setFlag(PF_active, false); }
    uint getVertexIndex() const { return flagsAndVertexIndex & kVertexIndexBitMask; }
    uint getCounter(uint type) const { return (packedCounters >> (type << 3)) & 0xff; }
    void incrementCounter(uint type) { packedCounters += (1u << (type << 3)); }
    void clearScatterEventFlags() { flagsAndVertexIndex &= ~((PF_transmission | PF_specular | PF_delta) << kVertexIndexBitCount); }
};

struct RenderStats { uint64_t scatterRays = 0, shadowRays = 0, nodeVisits = 0, triTests = 0; };

struct PathTracerCtx
{
    const Scene* scene; const Bvh2* bvh; const LightTable* lights; const RtxptPathTracerConstants* c;
    uint sampleIndex;       // Bridge::getSampleIndex() = sampleBaseIndex + subSampleIndex (BridgeDonut:510-513)
    RenderStats* stats;
    // guide export of the reference-mode path (Bridge::ExportSurfaceInit / ExportSurface / ExportNonSurface, BridgeDonut:1096-1153)
    const float* worldToClip = nullptr;     // view.matWorldToClip, row-major, row vector x matrix
    struct GuideOut* guide = nullptr;
    // realtime mode (PATH_TRACER_MODE_BUILD_STABLE_PLANES / _FILL_STABLE_PLANES)
    uint mode = MODE_REFERENCE;
    const RealtimeTargets* sp = nullptr;
    // NEE-AT temporal feedback + local samplers (pt_neeat.h); null = the global-table-only tier (RtxptPathTracerConstants::NEEATFeedback == 0)
    NeeatState* neeat = nullptr;
    // test hook (oracle.cpp's known-answer mirrors): answers a shadow ray instead of the BVH, as the stub bridge of oracle/ref_bridge_stub.h does for the reference's code
    bool (*visibilityOverride)(float3 origin, float3 dir, float tMax, void* user) = nullptr; void* visibilityUser = nullptr;
    void (*cameraRayOverride)(uint px, uint py, float3& origin, float3& dir, void* user) = nullptr; void* cameraRayUser = nullptr;
    float3 (*envEvalOverride)(float3 localDir, float lod) = nullptr;       // stands in for the environment cube map (before the colour multiplier)
    float noisyRadianceAttenuationOverride = 0.0f;
    float noisyRadianceAttenuation() const { return noisyRadianceAttenuationOverride != 0.0f ? noisyRadianceAttenuationOverride : 1.0f / float(sp->rt->subSampleCount); }      // Bridge::getNoisyRadianceAttenuation = invSubSampleCount (BridgeDonut:515-523)
};
struct GuideOut { float depth; uint throughput; float motion[3]; };

inline uint Pack_R11G11B10_FLOAT(float3 rgb)       // Utils/Packing.hlsli:175-184
{
    const float top = asfloat(0x477C0000u);
    rgb = f3(std::min(rgb.x, top), std::min(rgb.y, top), std::min(rgb.z, top));
    uint r = ((f32tof16(rgb.x) + 8) >> 4) & 0x000007FF;
    uint g = ((f32tof16(rgb.y) + 8) << 7) & 0x003FF800;
    uint b = ((f32tof16(rgb.z) + 16) << 17) & 0xFFC00000;
    return r | g | b;
}
inline float clipDepth(const float* M, float3 p)
{
    const float z = p.x * M[2] + p.y * M[6] + p.z * M[10] + M[14], w = p.x * M[3] + p.y * M[7] + p.z * M[11] + M[15];
    return z / w;
}

inline bool HasFinishedSurfaceBounces(const PathTracerCtx& x, uint vertexIndex, uint diffuseBounces)
{
    if (x.c->bounceCount < vertexIndex) return true;
    return diffuseBounces > x.c->diffuseBounceCount;
}

// ---- env map (Lighting/EnvMap.hlsli) ----------------------------------------------------------------------------------
inline float3 envToLocal(const PathTracerCtx& x, float3 d) { return mul_vec_33of34(d, x.c->envMap.InvTransform); }
inline float3 envToWorld(const PathTracerCtx& x, float3 d) { return mul_vec_33of34(d, x.c->envMap.Transform); }
inline float3 envEvalLocal(const PathTracerCtx& x, float3 localDir, float lod)
{
    if (x.envEvalOverride) return x.envEvalOverride(localDir, lod) * f3(x.c->envMap.ColorMultiplier[0], x.c->envMap.ColorMultiplier[1], x.c->envMap.ColorMultiplier[2]);
    return x.scene->env.sampleLevel(localDir, lod) * f3(x.c->envMap.ColorMultiplier[0], x.c->envMap.ColorMultiplier[1], x.c->envMap.ColorMultiplier[2]);
}

// ---- light sampler (global table only) ------------------------------------------------------------------------------------
inline float SampleGlobalPDF(const LightTable& lt, uint lightIndex) { return float(lt.proxyCounters[lightIndex]) / float(lt.samplingProxyCount); }
inline float EvalMISBalance(float n0, float p0, float n1, float p1) { float q0 = n0 * p0, q1 = n1 * p1; return saturate(q0 / (q0 + q1)); }   // Utils/Utils.hlsli:407-437
// LightSampler::SampleGlobal (LightSampler.hlsli:112-122)
inline uint SampleGlobal(const LightTable& lt, float rnd, float& pdf)
{
    const uint M = lt.samplingProxyCount;
    const uint lightIndex = lt.proxyIndices[std::min(uint(rnd * float(M)), M - 1)];
    pdf = float(lt.proxyCounters[lightIndex]) / float(M);
    return lightIndex;
}
// LightSampler::IsScreenSpaceCoherentHeuristic (LightSampler.hlsli:45-49)
inline bool IsScreenSpaceCoherentHeuristic(float threshold, float rayConeWidth, float totalPathLength) { return (rayConeWidth / totalPathLength) < threshold; }
// neeat / pixel / misInfo: the local sampler the previous vertex drew from (LightSampler.hlsli:318-333: localCount > 0 only for screen-space-coherent vertices)
inline float ComputeLightVsBSDF_MIS_ForBSDF(const LightTable& lt, uint lightIndex, float bsdfPdf, float solidAnglePdf, uint fullSampleCount, const NeeatState* neeat = nullptr, uint pathId = 0,
                                            bool isSSC = false, uint candidateSampleCount = 0)
{
    float globalPdf = SampleGlobalPDF(lt, lightIndex);
    float localPdf = 0;
    if (neeat && isSSC && ComputeCandidateSampleLocalCount(neeat->localToGlobalSampleRatio, candidateSampleCount) > 0)
        localPdf = SampleLocalPDF(*neeat, LocalSamplingTilePos(*neeat, pathId >> 16, pathId & 0xFFFFu), lightIndex);
    float lightAvgPdf = (localPdf + globalPdf) * float(fullSampleCount);
    return EvalMISBalance(1, bsdfPdf, 1, lightAvgPdf * solidAnglePdf);
}

struct LightSample { float3 Li = f3(0); float Distance = 0; float3 Direction = f3(0); uint LightIndex = 0xFFFFFFFFu; float SelectionPdf = 0, SolidAnglePdf = 0; bool LightSampleableByBSDF = false, FromLocalDistribution = false;
                     bool Valid() const { return Li.x > 0 || Li.y > 0 || Li.z > 0; } };
// LightSampler::ComputeLightSelectionPdfs (LightSampler.hlsli:245-275): the pdf the OTHER sampler (global vs the pixel's tile) would have picked this light with
inline void ComputeLightSelectionPdfs(const LightTable& lt, const NeeatState* neeat, uint tileAddress, const LightSample& s, uint localCount, uint globalCount, float& thisPdf, float& otherPdf,
                                      float& thisCount, float& otherCount)
{
    thisPdf = s.SelectionPdf;
    if (s.FromLocalDistribution) { otherPdf = SampleGlobalPDF(lt, s.LightIndex); thisCount = float(localCount); otherCount = float(globalCount); }
    else
    {
        thisCount = float(globalCount);
        if (localCount != 0) { otherPdf = SampleLocalPDF(*neeat, tileAddress, s.LightIndex); otherCount = float(localCount); }
        else { otherPdf = 0; otherCount = 0; }
    }
}
// LightSampler::ComputeLightVsBSDF_MIS_ForLight (LightSampler.hlsli:282-314; LightSamplingMISBoost() == 1)
inline float ComputeLightVsBSDF_MIS_ForLight(const LightSample& s, float thisPdf, float otherPdf, uint fullSampleCount, float bsdfPdf)
{
    const float lightAvgPdf = (thisPdf + otherPdf) * float(fullSampleCount);
    return EvalMISBalance(1, lightAvgPdf * s.SolidAnglePdf, 1, s.LightSampleableByBSDF ? bsdfPdf : 0.0f);
}

// ---- firefly filter (PathTracerHelpers.hlsli:183-219) ---------------------------------------------------------------------
inline float ComputeRayConeSpreadAngleExpansionByScatterPDF(float pdf, float growthFactor = 0.3f)
{
    return growthFactor * 2.0f * FastACos(std::max(-1.0f, 1.0f - (1.0f / pdf) / (2.0f * K_PI)));
}
inline float ComputeNewScatterFireflyFilterK(float currentK, float bouncePDF, float lobeP)
{
    const float minK = 0.00001f;
    float angle = (bouncePDF == 0) ? 0 : ComputeRayConeSpreadAngleExpansionByScatterPDF(bouncePDF, 1.0f);
    const float k = 32;
    float p = k / (k + angle * angle);
    p *= FastSqrt(lobeP);
    return lp(std::max(minK, currentK * p));
}
inline float3 FireflyFilter(float3 signalIn, float threshold, float fireflyFilterK)
{
    // lpfloat arithmetic: every operation rounds to binary16 (Sample.cpp:1017 compiles the shaders with 16-bit types; pinned by tests/golden/helpers_golden.npz)
    signalIn = lp(signalIn);                                                          // the parameter is an lpfloat3
    float thr = lp(threshold * fireflyFilterK);
    float maxR = lp(lp(lp(signalIn.x + signalIn.y) + signalIn.z) / 3.0f);          // Average( lpfloat3 ), Utils.hlsli:63-66
    if (maxR > thr) signalIn = f3(lp(lp(signalIn.x / maxR) * thr), lp(lp(signalIn.y / maxR) * thr), lp(lp(signalIn.z / maxR) * thr));
    return signalIn;
}
inline float FireflyFilterShort(float signalAverage, float threshold, float fireflyFilterK)
{
    float thr = threshold * fireflyFilterK;
    return (signalAverage > thr) ? (1.0f / signalAverage * thr) : 1.0f;
}

// ---- camera (BridgeDonut:543-564, PathTracerHelpers.hlsli:126-153) ---------------------------------------------------------
inline void computeCameraRay(const PathTracerCtx& x, uint px, uint py, float3& origin, float3& dir)
{
    if (x.cameraRayOverride) { x.cameraRayOverride(px, py, origin, dir, x.cameraRayUser); return; }
    const RtxptCameraData& cam = x.c->camera;
    SampleSequenceGenerator sg = SampleSequenceGenerator::make(SampleGeneratorVertexBase::make((px << 16) | py, 0, x.sampleIndex));
    float r0 = sg.Next1D(), r1 = sg.Next1D();
    float2 subPixelOffset = f2(cam.Jitter[0] + (r0 - 0.5f) * x.c->perPixelJitterAAScale, cam.Jitter[1] + (r1 - 0.5f) * x.c->perPixelJitterAAScale);
    float d0 = sg.Next1D(), d1 = sg.Next1D();
    float2 p = f2((float(px) + 0.5f + (-subPixelOffset.x)) / float(cam.ViewportSize[0]), (float(py) + 0.5f + subPixelOffset.y) / float(cam.ViewportSize[1]));
    float2 ndc = f2(2 * p.x - 1, -2 * p.y + 1);
    float3 U = f3(cam.CameraU[0], cam.CameraU[1], cam.CameraU[2]), V = f3(cam.CameraV[0], cam.CameraV[1], cam.CameraV[2]), W = f3(cam.CameraW[0], cam.CameraW[1], cam.CameraW[2]);
    origin = f3(cam.PosW[0], cam.PosW[1], cam.PosW[2]);
    dir = ndc.x * U + ndc.y * V + W;
    float2 apertureSample = sample_disk(f2(d0, d1));
    float3 rayTarget = origin + dir;
    origin = origin + cam.ApertureRadius * (apertureSample.x * normalize(U) + apertureSample.y * normalize(V));
    dir = normalize(rayTarget - origin);
    float invCos = 1.f / dot(normalize(W), dir);
    float tMin = cam.NearZ * invCos;
    origin = origin + dir * tMin;
}

// ---- PathTracer.hlsli:382-404 ---------------------------------------------------------------------------------------------
inline void UpdatePathTravelled(PathState& path, float rayTCurrent)
{
    path.flagsAndVertexIndex += 1;
    path.rayCone = path.rayCone.propagateDistance(rayTCurrent);
    path.sceneLength = std::min(path.sceneLength + rayTCurrent, kMaxRayTravel);
}
// PathTracer.hlsli:139-162
inline void AccumulatePathRadiance(const PathTracerCtx& x, PathState& path, float3 radiance, float specularRadianceAvg, bool stablePlaneOnBranch)
{
    if (x.mode == MODE_REFERENCE) { float4 L = path.GetL(); path.SetL(f4(L.x + radiance.x, L.y + radiance.y, L.z + radiance.z, L.w)); }
    else if (x.mode == MODE_BUILD_STABLE_PLANES) x.sp->AccumulateStableRadiance(path.id >> 16, path.id & 0xFFFF, radiance);
    else if (!stablePlaneOnBranch)      // FILL: the stable part was captured by the BUILD pass
    {
        const float a = x.noisyRadianceAttenuation();
        float4 L = path.GetL();
        path.SetL(f4(L.x + radiance.x * a, L.y + radiance.y * a, L.z + radiance.z * a, L.w + specularRadianceAvg * a));
    }
}
inline void StablePlanesHandleMiss(const PathTracerCtx& x, PathState& path, float3 emission, float3 rayOrigin, float3 rayDir);
inline void ExportSpecHitTStop(const PathTracerCtx& x, const PathState& path)      // BridgeDonut:1160-1175
{
    float& t = x.sp->specularHitT[size_t(path.id & 0xFFFF) * x.sp->width + (path.id >> 16)];
    if (t < 0) t = std::max(0.0f, path.sceneLength + t);
}

// ---- miss (PathTracer.hlsli:407-503) ---------------------------------------------------------------------------------------
inline void HandleMiss(const PathTracerCtx& x, PathState& path, float3 rayDir, float rayTCurrent)
{
    const float3 rayOrigin = path.origin;
    UpdatePathTravelled(path, rayTCurrent);
    const bool build = x.mode == MODE_BUILD_STABLE_PLANES;
    if (x.mode == MODE_FILL_STABLE_PLANES && path.hasFlag(PF_exportSpecHitTQueued)) { ExportSpecHitTStop(x, path); path.setFlag(PF_exportSpecHitTQueued, false); }
    float3 environmentEmission = f3(0);
    NEEBSDFMISInfo misInfo = NEEBSDFMISInfo::Unpack16bit(build ? 0u : path.GetPackedMISInfo());
    if (x.lights->envEnabled)
    {
        float mipLevel = (path.getCounter(CTR_DiffuseBounces) > 1) ? x.c->EnvironmentMapDiffuseSampleMIPLevel : 0.0f;
        float3 localDir = envToLocal(x, rayDir);
        float3 Le = envEvalLocal(x, localDir, mipLevel);
        float misWeight = 1.0f;
        float bsdfScatterPdf = build ? 0.0f : path.GetBsdfScatterPdf();
        if (misInfo.LightSamplingEnabled && bsdfScatterPdf != 0)
        {
            float2 uv = ndir_to_oct_equal_area_unorm(localDir);
            uint cx = uint(uv.x * float(IMPORTANCE_MAP_DIM)), cy = uint(uv.y * float(IMPORTANCE_MAP_DIM));
            cx = std::min(cx, IMPORTANCE_MAP_DIM - 1); cy = std::min(cy, IMPORTANCE_MAP_DIM - 1);   // Texture2D.Load out of range returns 0 in D3D; uv==1 is the only way to get there
            uint envLightIndex = x.lights->envLookupMap[size_t(cy) * IMPORTANCE_MAP_DIM + cx];
            EnvironmentQuadLight eq = EnvironmentQuadLight::Create(x.lights->lights[envLightIndex]);
            misWeight = ComputeLightVsBSDF_MIS_ForBSDF(*x.lights, envLightIndex, bsdfScatterPdf, eq.SolidAnglePdf(), misInfo.FullSamples, x.neeat, path.id, misInfo.LightSamplingIsSSC, misInfo.CandidateSamples);
        }
        environmentEmission = lp(misWeight * Le);
    }
    float baseFFThreshold = lp(x.c->fireflyFilterThreshold);
    if (baseFFThreshold != 0 && !build) environmentEmission = FireflyFilter(environmentEmission, baseFFThreshold, path.GetFireflyFilterK());
    if (build) StablePlanesHandleMiss(x, path, environmentEmission, rayOrigin, rayDir);
    if (x.mode == MODE_REFERENCE && x.guide && x.worldToClip) { x.guide->depth = clipDepth(x.worldToClip, path.origin + rayDir * rayTCurrent); x.guide->throughput = 0; x.guide->motion[0] = x.guide->motion[1] = x.guide->motion[2] = 0; }     // ExportNonSurface (PathTracer.hlsli:487)
    if (any_gt0(environmentEmission))
    {
        const float3 radiance = path.GetThp() * environmentEmission;
        AccumulatePathRadiance(x, path, radiance, path.hasFlag(PF_stablePlaneBaseScatterDiff) ? 0.0f : Average(radiance), path.hasFlag(PF_stablePlaneOnBranch));
    }
    path.setFlag(PF_hit, false);
    path.terminate();
}

// ---- Russian roulette (PathTracer.hlsli:182-208) ------------------------------------------------------------------------------
inline bool HandleRussianRoulette(const PathTracerCtx& x, PathState& path, UniformSampleSequenceGenerator& sg)
{
    if (!x.c->enableRussianRoulette) return false;
    const float rrVal = sqrtf(Luminance(path.GetThp()));
    float prob = saturate(0.85f - rrVal); prob = prob * prob;
    prob = saturate(prob + std::max(0.0f, (float(path.getVertexIndex()) / float(x.c->bounceCount) - 0.4f)));
    if (sg.Next1D() < prob) return true;
    float thpRuRuCorrection = lp(1.0f / (1.0f - prob));
    path.SetPackedMISInfo_ThpRuRuCorrection(path.GetPackedMISInfo(), thpRuRuCorrection);
    return false;
}

// ---- scatter (PathTracer.hlsli:217-380) ---------------------------------------------------------------------------------------
inline void StablePlanesOnScatter(const PathTracerCtx& x, PathState& path, const BSDFSample& bs);
inline bool GenerateScatterRay(const PathTracerCtx& x, const ShadingData& sd, const StandardBSDF& bsdf, PathState& path, const SampleGeneratorVertexBase& sgBase)
{
    float u[4] = { 0, 0, 0, 0 };
    if (x.c->enableLDSamplerForBSDF && path.getCounter(CTR_DiffuseBounces) < 1) GenerateLD(3, sgBase, SeedScatterBSDF, u);
    else GenerateUniform(3, sgBase, SeedScatterBSDF, u);
    BSDFSample bs;
    if (!bsdf.sample(sd.frame(), u, bs)) return false;

    path.dir = bs.wo;
    const bool onDominantDenoisingLayer = path.hasFlag(PF_stablePlaneOnPlane) && path.hasFlag(PF_stablePlaneOnDominantBranch);
    path.SetThp(path.GetThp() * bs.weight);
    path.clearScatterEventFlags();
    path.origin = sd.computeNewRayOrigin(bs.isLobe(Lobe_Reflection));
    const float roughness = bsdf.data.roughness;
    bool isDiffuse = bs.isLobe(Lobe_DiffuseReflection) || bs.isLobe(Lobe_DiffuseTransmission) || roughness > 0.25f;
    if (isDiffuse)
    {
        if (!(bs.isLobe(Lobe_DiffuseTransmission) && ((path.getVertexIndex() % 2) == 1))) path.incrementCounter(CTR_DiffuseBounces);
    }
    else path.setFlag(PF_specular);
    if (bs.isLobe(Lobe_Transmission))
    {
        path.setFlag(PF_transmission);
        if (x.c->nestedDielectricsQuality > 0 && !sd.thinSurface)
        {
            path.interiorList.handleIntersection(sd.materialID, sd.nestedPriority, sd.frontFacing);
            path.setFlag(PF_insideDielectricVolume, !path.interiorList.isEmpty());
        }
    }
    if (bs.isLobe(Lobe_Delta)) path.setFlag(PF_delta);
    else
    {
        path.setFlag(PF_deltaOnlyPath, false);
        path.rayCone = RayCone::make(path.rayCone.getWidth(), std::min(path.rayCone.getSpreadAngle() + ComputeRayConeSpreadAngleExpansionByScatterPDF(bs.pdf), 2.0f * K_PI));
    }
    if (x.mode == MODE_FILL_STABLE_PLANES)
    {   // specular hit distance of the dominant plane (PathTracer.hlsli:295-324)
        const bool isDiffuseForSpecHitT = bs.isLobe(Lobe_DiffuseReflection) || bs.isLobe(Lobe_DiffuseTransmission) || roughness > 0.35f;
        if (onDominantDenoisingLayer && !isDiffuseForSpecHitT)
        {
            if (!sd.psdBlockMotionVectorsAtSurface)
            {
                path.setFlag(PF_exportSpecHitTQueued, true);
                x.sp->specularHitT[size_t(path.id & 0xFFFF) * x.sp->width + (path.id >> 16)] = -path.sceneLength;      // Bridge::ExportSpecHitTStart
            }
        }
        else if (path.hasFlag(PF_exportSpecHitTQueued))
        {
            const bool hasNonDeltaLobes = (bsdf.getLobes() & Lobe_NonDelta) != 0;
            if (hasNonDeltaLobes || path.getCounter(CTR_BouncesFromStablePlane) > 4) { ExportSpecHitTStop(x, path); path.setFlag(PF_exportSpecHitTQueued, false); }
        }
    }
    float fireflyFilterK = (x.c->fireflyFilterThreshold != 0) ? ComputeNewScatterFireflyFilterK(path.GetFireflyFilterK(), bs.pdf, bs.lobeP) : 0.0f;
    path.SetFireflyFilterK_BsdfScatterPdf(fireflyFilterK, bs.pdf);
    if (x.mode == MODE_FILL_STABLE_PLANES) StablePlanesOnScatter(x, path, bs);
    path.setFlag(PF_enableThreadReorder, true);
    return true;
}

// ---- NEE (PathTracerNEE.hlsli) --------------------------------------------------------------------------------------------------
struct NEEResult
{
    uint pkg[2]; NEEBSDFMISInfo BSDFMISInfo;
    NEEResult() { pkg[0] = Fp32ToFp16(f2(0, 0)); pkg[1] = Fp32ToFp16(f2(0, 0)); }
    float4 Get() const { float2 a = Fp16ToFp32(pkg[0]), b = Fp16ToFp32(pkg[1]); return f4(a.x, a.y, b.x, b.y); }
    void Accumulate(float3 radiance, float specAvg) { float4 v = Get(); pkg[0] = Fp32ToFp16(f2(v.x + radiance.x, v.y + radiance.y)); pkg[1] = Fp32ToFp16(f2(v.z + radiance.z, v.w + specAvg)); }
};

inline NEEResult HandleNEE(const PathTracerCtx& x, const PathState& pre, const ShadingData& sd, const StandardBSDF& bsdf, UniformSampleSequenceGenerator& sg)
{
    NEEResult result;
    const LightTable& lt = *x.lights;
    if (!x.c->NEEEnabled) return result;
    const uint fullSamples = std::min(63u, x.c->NEEFullSamples);
    const bool hasNonDeltaLobes = (bsdf.getLobes() & Lobe_NonDelta) != 0;
    if (!(hasNonDeltaLobes && !lt.IsEmpty() && fullSamples > 0)) return result;
    const uint candidateSampleCount = x.c->NEECandidateSamples;
    const BSDFFrame frame = sd.frame();
    result.BSDFMISInfo.LightSamplingEnabled = true;
    const bool isSSC = IsScreenSpaceCoherentHeuristic(x.neeat ? x.neeat->settings.screenSpaceVsWorldSpaceThreshold : 0.3f, pre.rayCone.getWidth(), pre.sceneLength);
    result.BSDFMISInfo.LightSamplingIsSSC = isSSC;
    result.BSDFMISInfo.CandidateSamples = candidateSampleCount;
    result.BSDFMISInfo.FullSamples = fullSamples;
    // GetCandidateSampleCounts: local candidates only for screen-space-coherent vertices and only once last frame's feedback has built the tile samplers
    const uint localCount = (x.neeat && isSSC) ? ComputeCandidateSampleLocalCount(x.neeat->localToGlobalSampleRatio, candidateSampleCount) : 0u, globalCount = candidateSampleCount - localCount;
    const uint pixelX = pre.id >> 16, pixelY = pre.id & 0xFFFFu;
    const uint tileAddress = x.neeat ? LocalSamplingTilePos(*x.neeat, pixelX, pixelY) : 0u;
    for (uint sampleIndex = 0; sampleIndex < fullSamples; sampleIndex++)
    {
        // GenerateLightSample: weighted reservoir sampling over the candidates
        LightSample picked; float weightSum = 0, candidateWeight = 0;
        for (uint i = 0; i < candidateSampleCount; i++)
        {
            const bool sampleIsLocal = i >= globalCount;
            float rnd = sg.Next1D();
            uint lightIndex; float selectionPdf;
            if (sampleIsLocal) lightIndex = SampleLocal(*x.neeat, tileAddress, rnd, selectionPdf);
            else lightIndex = SampleGlobal(lt, rnd, selectionPdf);
            const PolymorphicLightInfo& li = lt.lights[lightIndex];
            float2 interiorRnd; interiorRnd.x = sg.Next1D(); interiorRnd.y = sg.Next1D();
            PolymorphicLightSample ls = {};
            if (LightType(li) == kLightTypeTriangle) ls = TriangleLight::Create(li).CalcSample(interiorRnd, sd.posW);
            else if (LightType(li) == kLightTypeSphere)
            {   // PolymorphicLight::CalcSample, kSphere + the shaping factor applied to every sample with a positive pdf (PolymorphicLight.hlsli:643-676)
                const PolymorphicLightInfoEx ex = lt.exOf(lightIndex);
                ls = SphereLight::Create(li, ex).CalcSample(interiorRnd, sd.posW);
                if (ls.SolidAnglePdf > 0) ls.Radiance = ls.Radiance * evaluateLightShaping(unpackLightShaping(li, ex), sd.posW, ls.Position);
            }
            else if (LightType(li) == kLightTypeEnvironmentQuad)
            {
                EnvironmentQuadLight e = EnvironmentQuadLight::Create(li);
                float2 subTexelPos = f2((float(e.NodeX) + interiorRnd.x) / float(e.NodeDim), (float(e.NodeY) + interiorRnd.y) / float(e.NodeDim));
                float3 worldDir = envToWorld(x, oct_to_ndir_equal_area_unorm(subTexelPos));
                ls.Position = sd.posW + worldDir * DISTANT_LIGHT_DISTANCE;
                ls.Normal = -worldDir; ls.Radiance = e.Radiance; ls.SolidAnglePdf = e.SolidAnglePdf(); ls.LightSampleableByBSDF = true;
            }
            LightSample cs;
            const float pdf = ls.SolidAnglePdf * selectionPdf;
            cs.Li = pdf > 0.f ? (ls.Radiance / pdf) : f3(0);
            cs.SolidAnglePdf = ls.SolidAnglePdf;
            float3 surfToLight = ls.Position - sd.posW;
            cs.Distance = length(surfToLight);
            cs.Direction = surfToLight / std::max(cs.Distance, 1e-7f);
            cs.LightIndex = lightIndex; cs.SelectionPdf = selectionPdf; cs.LightSampleableByBSDF = ls.LightSampleableByBSDF; cs.FromLocalDistribution = sampleIsLocal;
            float wrsWeight = max3(cs.Li) * bsdf.evalPdf(frame, cs.Direction);
            float wrsRnd = sg.Next1D();
            weightSum += wrsWeight;
            float wrsThreshold = saturate(wrsWeight / weightSum);
            if (wrsRnd < wrsThreshold) { picked = cs; candidateWeight = wrsWeight; }
        }
        picked.Li = picked.Li * (1.0f / (candidateWeight / weightSum));

        // ProcessLightSample
        bool visible = false;
        if (picked.Valid())
        {
            float faceSide = dot(sd.N, picked.Direction) >= 0 ? 1.0f : -1.0f;
            float3 o = ComputeRayOrigin(sd.posW, sd.faceNCorrected * faceSide);
            if (x.stats) x.stats->shadowRays++;
            if (x.visibilityOverride) visible = x.visibilityOverride(o, picked.Direction, picked.Distance * 0.9985f, x.visibilityUser);
            else
            {
                Hit h = x.bvh->trace(*x.scene, o, picked.Direction, 0.0f, picked.Distance * 0.9985f, true, x.stats ? &x.stats->nodeVisits : nullptr, x.stats ? &x.stats->triTests : nullptr);
                visible = !h.valid();
            }
        }
        if (visible)
        {
            float fadeOut = (sd.shadowNoLFadeout > 0) ? saturate((dot(picked.Direction, sd.vertexN) - sd.shadowNoLFadeout) / (2.0f * sd.shadowNoLFadeout)) : 1.0f;
            // ComputeLightSelectionPdfs: the pdf the other sampler would have picked this light with
            float thisPdf, otherPdf, thisCount, otherCount;
            ComputeLightSelectionPdfs(lt, x.neeat, tileAddress, picked, localCount, globalCount, thisPdf, otherPdf, thisCount, otherCount);
            float wrsMIS = EvalMISBalance(1, thisPdf, 1, otherPdf) / thisCount;
            float scatterPdfForDir = bsdf.evalPdf(frame, picked.Direction);
            float pathMIS = ComputeLightVsBSDF_MIS_ForLight(picked, thisPdf, otherPdf, fullSamples, scatterPdfForDir);
            float3 Li = picked.Li * (fadeOut * wrsMIS * pathMIS / float(fullSamples));
            float4 bsdfThp = bsdf.eval(frame, picked.Direction);
            float3 radiance = xyz(bsdfThp) * Li;
            float radianceAvg = Average(radiance);
            float specAvg = bsdfThp.w * Average(Li);
            if (x.c->fireflyFilterThreshold != 0)
            {
                const float pdf = picked.SelectionPdf * picked.SolidAnglePdf;
                float neeFireflyFilterK = ComputeNewScatterFireflyFilterK(pre.GetFireflyFilterK(), pdf, 1.0f);
                radiance *= FireflyFilterShort(radianceAvg, x.c->fireflyFilterThreshold, neeFireflyFilterK);
            }
            float3 preScatterThp = pre.GetThp();
            radiance *= preScatterThp;
            specAvg *= Average(preScatterThp);
            radianceAvg *= Average(preScatterThp);
            result.Accumulate(radiance, specAvg);
            // temporal feedback for NEE-AT: how much this pixel wanted this light (path throughput x BSDF x light, un-filtered), PathTracerNEE.hlsli:276-283
            if (x.neeat && picked.LightIndex != RTXPT_INVALID_LIGHT_INDEX && x.neeat->temporalFeedbackRequired)
                InsertFeedbackFromNEE(*x.neeat, lt, pixelX, pixelY, isSSC, picked.LightIndex, radianceAvg, sg.Next1D());
        }
    }
    return result;
}

// ---- hit (PathTracer.hlsli:505-762) --------------------------------------------------------------------------------------------
inline void StablePlanesHandleHit(const PathTracerCtx& x, PathState& path, float3 rayOrigin, float3 rayDir, float rayTCurrent, const SurfaceData& surfaceData, bool pathStopping);
inline void HandleHitSurface(const PathTracerCtx& x, PathState& path, float3 rayOrigin, float3 rayDir, float rayTCurrent, SurfaceData surface);
inline void HandleHit(const PathTracerCtx& x, PathState& path, float3 rayOrigin, float3 rayDir, float rayTCurrent, const Tri& tri, float2 barycentrics)
{
    UpdatePathTravelled(path, rayTCurrent);
    HandleHitSurface(x, path, rayOrigin, rayDir, rayTCurrent, loadSurface(*x.scene, tri.instanceIndex, tri.geometryIndex, tri.primitiveIndex, barycentrics, rayDir, path.rayCone, x.c->texLODBias, path.getVertexIndex(),
                                                                         path.id >> 16, path.id & 0xFFFF, x.sampleIndex));
}
// PathTracer::HandleHit after Bridge::loadSurface (PathTracer.hlsli:517-763); the path has been advanced to this vertex (UpdatePathTravelled)
inline void HandleHitSurface(const PathTracerCtx& x, PathState& path, float3 rayOrigin, float3 rayDir, float rayTCurrent, SurfaceData surface)
{
    const bool build = x.mode == MODE_BUILD_STABLE_PLANES;
    const uint ndq = x.c->nestedDielectricsQuality;
    if (ndq > 0 && !path.interiorList.isEmpty())
    {   // homogeneous absorption (BridgeDonut:871-887, HomogeneousVolumeSampler::evalTransmittance)
        uint materialID = path.interiorList.getTopMaterialID();
        float3 sigmaA = f3(0);
        if (materialID < x.scene->desc->materialCount)
        {
            const RtxptMaterialData& m = x.scene->desc->materials[materialID];
            float dist = std::max(1e-30f, m.VolumeAttenuationDistance);
            sigmaA = f3(-logf(clampf(m.VolumeAttenuationColor[0], 1e-7f, 1)) / dist, -logf(clampf(m.VolumeAttenuationColor[1], 1e-7f, 1)) / dist, -logf(clampf(m.VolumeAttenuationColor[2], 1e-7f, 1)) / dist);
        }
        float3 transmittance = f3(expf(-rayTCurrent * sigmaA.x), expf(-rayTCurrent * sigmaA.y), expf(-rayTCurrent * sigmaA.z));
        path.SetThp(path.GetThp() * transmittance);
    }
    if (ndq > 0 && !surface.sd.thinSurface)
    {   // HandleNestedDielectrics, quality 1: kMaxRejectedDielectricHits = 4, NESTED_DIELECTRICS_AVOID_TERMINATION
        uint nestedPriority = surface.sd.nestedPriority;
        if (path.getCounter(CTR_RejectedHits) < 4 && !path.interiorList.isTrueIntersection(nestedPriority))
        {
            path.incrementCounter(CTR_RejectedHits);
            path.interiorList.handleIntersection(surface.sd.materialID, nestedPriority, surface.sd.frontFacing);
            path.origin = ComputeRayOrigin(surface.sd.posW, -surface.sd.faceNCorrected);
            path.flagsAndVertexIndex -= 1;
            return;     // rejected false hit: same direction continues from the far side
        }
        // ComputeOutsideIoR + Bridge::updateOutsideIoR
        uint outsideMaterialID = path.interiorList.getTopMaterialID();
        if (!surface.sd.frontFacing && outsideMaterialID == surface.sd.materialID) outsideMaterialID = path.interiorList.getNextMaterialID();
        float outsideIoR = 1.f;
        if (outsideMaterialID != InteriorList::kNoMaterial) outsideIoR = (outsideMaterialID >= x.scene->desc->materialCount) ? 1.0f : lp(x.scene->desc->materials[outsideMaterialID].IoR);
        surface.sd.IoR = outsideIoR;
        surface.bsdf.data.eta = lp(surface.sd.frontFacing ? (surface.sd.IoR / surface.interiorIoR) : (surface.interiorIoR / surface.sd.IoR));
    }
    const ShadingData& sd = surface.sd;
    const StandardBSDF& bsdf = surface.bsdf;

    // emissive triangle radiance with BSDF-side MIS
    float3 surfaceEmission = f3(0);
    NEEBSDFMISInfo misInfo = NEEBSDFMISInfo::Unpack16bit(build ? 0u : path.GetPackedMISInfo());
    const float pathBsdfScatterPdf = build ? 0.0f : path.GetBsdfScatterPdf();     // PathState.hlsli:157-159: no NEE, no MIS while building the planes
    if (any_gt0(sd.emission))
    {
        float misWeight = 1.0f;
        float bsdfScatterPdf = pathBsdfScatterPdf;
        if (misInfo.LightSamplingEnabled && bsdfScatterPdf != 0 && surface.neeTriangleLightIndex != RTXPT_INVALID_LIGHT_INDEX)
        {
            TriangleLight tl = TriangleLight::Create(x.lights->lights[surface.neeTriangleLightIndex]);
            float solidAnglePdf = tl.CalcSolidAnglePdfForMIS(rayOrigin, sd.posW);
            misWeight = ComputeLightVsBSDF_MIS_ForBSDF(*x.lights, surface.neeTriangleLightIndex, bsdfScatterPdf, solidAnglePdf, misInfo.FullSamples, x.neeat, path.id, misInfo.LightSamplingIsSSC, misInfo.CandidateSamples);
        }
        surfaceEmission = lp(sd.emission * misWeight);
    }
    if (surface.neeAnalyticLightIndex != RTXPT_INVALID_LIGHT_INDEX)
    {   // LightSampler::ComputeAnalyticLightProxyContributionWithMIS (LightSampler.hlsli:363-394, PathTracer.hlsli:636-648): sphere lights only
        const PolymorphicLightInfo& li = x.lights->lights[surface.neeAnalyticLightIndex];
        if (LightType(li) == kLightTypeSphere)
        {
            const SphereLight sl = SphereLight::Create(li, x.lights->exOf(surface.neeAnalyticLightIndex));
            float3 radiance, lightSamplePosition;
            if (sl.Eval(rayOrigin, rayDir, radiance, lightSamplePosition))
            {
                float mis = 1.0f;
                const float bsdfPdf = misInfo.LightSamplingEnabled ? pathBsdfScatterPdf : 0.0f;
                if (bsdfPdf != 0) mis = ComputeLightVsBSDF_MIS_ForBSDF(*x.lights, surface.neeAnalyticLightIndex, bsdfPdf, sl.CalcSolidAnglePdfForMIS(rayOrigin), misInfo.FullSamples, x.neeat, path.id, misInfo.LightSamplingIsSSC, misInfo.CandidateSamples);
                surfaceEmission = surfaceEmission + lp(radiance * mis);
            }
        }
    }
    if (any_gt0(surfaceEmission))
    {
        float baseFFThreshold = lp(x.c->fireflyFilterThreshold);
        if (baseFFThreshold != 0 && !build) surfaceEmission = FireflyFilter(surfaceEmission, baseFFThreshold, path.GetFireflyFilterK());
        if (any_gt0(surfaceEmission))
        {
            const float3 radiance = path.GetThp() * surfaceEmission;
            AccumulatePathRadiance(x, path, radiance, path.hasFlag(PF_stablePlaneBaseScatterDiff) ? 0.0f : Average(radiance), path.hasFlag(PF_stablePlaneOnBranch));
        }
    }
    const bool pathStopping = path.hasFlag(PF_terminateAtNextBounce);
    if (build) StablePlanesHandleHit(x, path, rayOrigin, rayDir, rayTCurrent, surface, pathStopping);      // before the throughput update (PathTracer.hlsli:679-682)
    if (x.mode == MODE_REFERENCE && x.guide && x.worldToClip)
    {   // ExportSurface (PathTracer.hlsli:684, BridgeDonut:1105-1129): virtual position along the pixel's camera ray at the path's scene length
        float3 co, cd; computeCameraRay(x, path.id >> 16, path.id & 0xFFFF, co, cd);
        x.guide->depth = clipDepth(x.worldToClip, co + cd * path.sceneLength);
        float3 t = path.GetThp(); x.guide->throughput = Pack_R11G11B10_FLOAT(f3(saturate(t.x), saturate(t.y), saturate(t.z)));
        x.guide->motion[0] = x.guide->motion[1] = x.guide->motion[2] = 0;
    }
    if (pathStopping) { path.terminate(); return; }

    path.SetThp(path.GetThp() * (build ? 1.0f : path.GetThpRuRuCorrection()));
    if (build) return;      // the BUILD pass consumed the emission and either re-aimed or terminated the path itself (PathTracer.hlsli:696-699)

    const SampleGeneratorVertexBase sgBase = SampleGeneratorVertexBase::make(path.id, path.getVertexIndex(), x.sampleIndex);
    UniformSampleSequenceGenerator uniformSG = UniformSampleSequenceGenerator::make(sgBase, SeedBase);
    const PathState preScatterPath = path;
    bool scatterValid = GenerateScatterRay(x, sd, bsdf, path, sgBase);
    NEEResult neeResult = HandleNEE(x, preScatterPath, sd, bsdf, uniformSG);
    path.SetPackedMISInfo_ThpRuRuCorrection(neeResult.BSDFMISInfo.Pack16bit(), path.GetThpRuRuCorrection());
    float4 nee = neeResult.Get();
    if (nee.x > 0 || nee.y > 0 || nee.z > 0 || nee.w > 0)
    {
        float specRadianceAvg = 0;
        if (!preScatterPath.hasFlag(PF_stablePlaneBaseScatterDiff))
        {   // PathTracer.hlsli:731-743
            const int bouncesFromStablePlane = int(preScatterPath.getCounter(CTR_BouncesFromStablePlane)) + 1;
            const bool specialCondition = (bouncesFromStablePlane == 1) || (preScatterPath.hasFlag(PF_deltaOnlyPath) && bouncesFromStablePlane <= 3);
            specRadianceAvg = specialCondition ? nee.w : Average(xyz(nee));
        }
        AccumulatePathRadiance(x, path, xyz(nee), specRadianceAvg, false);
    }
    if (!scatterValid) path.terminate();
    bool shouldTerminate = HasFinishedSurfaceBounces(x, path.getVertexIndex() + 1, path.getCounter(CTR_DiffuseBounces));
    shouldTerminate |= HandleRussianRoulette(x, path, uniformSG);
    if (shouldTerminate) path.setFlag(PF_terminateAtNextBounce);
}

// PathTracer::EmptyPathInitialize (PathTracer.hlsli:45-92) for the pass x.mode names; the primary ray is set by the caller (SetupPathPrimaryRay)
inline PathState EmptyPathInitialize(const PathTracerCtx& x, uint px, uint py, float pixelConeSpreadAngle)
{
    PathState path;
    path.id = (px << 16) | py;
    path.SetThp(f3(1));
    path.setFlag(PF_active); path.setFlag(PF_deltaOnlyPath, true);
    path.rayCone = RayCone::make(0, pixelConeSpreadAngle);
    if (x.mode == MODE_BUILD_STABLE_PLANES)
    {
        path.SetImageXform(identity3());
        path.setFlag(PF_stablePlaneOnDominantBranch, true);
        path.SetMotionVectorSceneLength(0);
    }
    else
    {
        path.SetL(f4(0, 0, 0, 0));
        path.SetFireflyFilterK_BsdfScatterPdf(1.0f, 0.0f);
        path.SetPackedMISInfo_ThpRuRuCorrection(NEEBSDFMISInfo().Pack16bit(), 1.0f);
    }
    path.setStablePlaneIndex(0);
    path.stableBranchID = 1;
    if (HasFinishedSurfaceBounces(x, path.getVertexIndex() + 1, path.getCounter(CTR_DiffuseBounces))) path.setFlag(PF_terminateAtNextBounce);
    return path;
}

// ---- one pixel, one sample (PathTracerSample.hlsl:201-256): returns L.rgb as stored to u_OutputColor (RGBA16F) -------------------
struct PixelResult { float rgb[3]; float primaryT; uint primaryTri; float primaryU, primaryV; };

inline PixelResult tracePixel(const PathTracerCtx& x, uint px, uint py)
{
    PathState path = EmptyPathInitialize(x, px, py, x.c->camera.PixelConeSpreadAngle);
    computeCameraRay(x, px, py, path.origin, path.dir);

    PixelResult out = {}; out.primaryT = -1.0f; out.primaryTri = 0xFFFFFFFFu;
    if (x.guide) { x.guide->depth = 0; x.guide->throughput = 0; x.guide->motion[0] = x.guide->motion[1] = x.guide->motion[2] = 0; }      // ExportSurfaceInit
    bool first = true;
    while (path.isActive())
    {
        float3 o = path.origin, d = path.dir;
        if (x.stats) x.stats->scatterRays++;
        Hit h = x.bvh->trace(*x.scene, o, d, 0.0f, kMaxRayTravel, false, x.stats ? &x.stats->nodeVisits : nullptr, x.stats ? &x.stats->triTests : nullptr);
        if (first) { first = false; if (h.valid()) { out.primaryT = h.t; out.primaryTri = h.triId; out.primaryU = h.u; out.primaryV = h.v; } }
        if (!h.valid()) HandleMiss(x, path, d, kMaxRayTravel);
        else HandleHit(x, path, o, d, h.t, x.bvh->tris[h.triId], f2(h.u, h.v));
    }
    float4 L = path.GetL();
    out.rgb[0] = L.x; out.rgb[1] = L.y; out.rgb[2] = L.z;      // already fp16 values: the RGBA16F store is lossless
    return out;
}


// =====================================================================================================================================================
// Realtime mode: PathTracerStablePlanes.hlsli (SplitDeltaPath :25-99, StablePlanesHandleHit :102-326, StablePlanesOnScatter :329-380,
// StablePlanesHandleMiss :382-412), StablePlanes.hlsli:232-252 (CommitDenoiserRadiance), PathTracerSample.hlsl:33-93 (FirstHitFromVBuffer),
// :96-113 (postProcessHit), :201-232 (raygen), PathTracer.hlsli:47-91 (EmptyPathInitialize), PathPayload.hlsli:29-110
// =====================================================================================================================================================
inline void packPayload(const PathState& path, uint out[20])
{
    out[0] = asuint(path.origin.x); out[1] = asuint(path.origin.y); out[2] = asuint(path.origin.z); out[3] = path.id;
    out[4] = asuint(path.dir.x); out[5] = asuint(path.dir.y); out[6] = asuint(path.dir.z); out[7] = asuint(path.sceneLength);
    out[8] = path.pack23[0]; out[9] = path.pack23[1]; out[10] = path.pack45[0]; out[11] = path.pack45[1];
    out[12] = path.interiorList.slots[0]; out[13] = path.interiorList.slots[1]; out[14] = path.packedCounters; out[15] = path.stableBranchID;
    out[16] = path.rayCone.widthSpreadAngleFP16; out[17] = path.pack0; out[18] = path.pack1; out[19] = path.flagsAndVertexIndex;
}
inline PathState unpackPayload(const uint in[20])
{
    PathState path;
    path.origin = f3(asfloat(in[0]), asfloat(in[1]), asfloat(in[2])); path.id = in[3];
    path.dir = f3(asfloat(in[4]), asfloat(in[5]), asfloat(in[6])); path.sceneLength = asfloat(in[7]);
    path.pack23[0] = in[8]; path.pack23[1] = in[9]; path.pack45[0] = in[10]; path.pack45[1] = in[11];
    path.interiorList.slots[0] = in[12]; path.interiorList.slots[1] = in[13]; path.packedCounters = in[14]; path.stableBranchID = in[15];
    path.rayCone.widthSpreadAngleFP16 = in[16]; path.pack0 = in[17]; path.pack1 = in[18]; path.flagsAndVertexIndex = in[19];
    return path;
}

// splits out one delta lobe: the new path leaves the surface along lobe.dir; the accumulated rotation (imageXform) gets the local mirror / refraction turn
inline PathState SplitDeltaPath(const PathTracerCtx& x, const PathState& oldPath, float3 rayDir, const SurfaceData& surfaceData, const DeltaLobe& lobe, uint deltaLobeIndex, bool verifyDominantFlag)
{
    const ShadingData& sd = surfaceData.sd;
    PathState newPath = oldPath;
    newPath.dir = lobe.dir;
    newPath.SetThp(newPath.GetThp() * lobe.thp);
    newPath.origin = sd.computeNewRayOrigin(lobe.transmission == 0);
    newPath.stableBranchID = StablePlanesAdvanceBranchID(oldPath.stableBranchID, deltaLobeIndex);
    newPath.setFlag(PF_delta);
    if (!lobe.transmission) newPath.setFlag(PF_specular);
    else
    {
        newPath.setFlag(PF_transmission);
        if (x.c->nestedDielectricsQuality > 0 && !sd.thinSurface)
        {
            newPath.interiorList.handleIntersection(sd.materialID, sd.nestedPriority, sd.frontFacing);
            newPath.setFlag(PF_insideDielectricVolume, !newPath.interiorList.isEmpty());
        }
    }
    if (newPath.GetMotionVectorSceneLength() == 0)      // transform updates stop behind a surface that blocks motion vectors
    {
        mat3 localT;        // lpfloat3x3: fp16 elements; products accumulated in fp32 here and rounded once per element
        if (lobe.transmission) localT = lp(MatrixRotateFromTo(lobe.dir, rayDir));
        else
        {
            mat3 toTangent; toTangent.r[0] = lp(sd.T); toTangent.r[1] = lp(sd.B); toTangent.r[2] = lp(sd.N);
            mat3 mirror; mirror.r[0] = f3(1, 0, 0); mirror.r[1] = f3(0, 1, 0); mirror.r[2] = f3(0, 0, -1);
            localT = lp(mul(mirror, toTangent));
            localT = lp(mul(transpose(toTangent), localT));
        }
        newPath.SetImageXform(mul(newPath.GetImageXform(), localT));
    }
    if (verifyDominantFlag && newPath.hasFlag(PF_stablePlaneOnDominantBranch))
    {
        const int psdDominantDeltaLobeIndex = int(sd.psdDominantDeltaLobeP1) - 1;
        if (int(deltaLobeIndex) != psdDominantDeltaLobeIndex) newPath.setFlag(PF_stablePlaneOnDominantBranch, false);
    }
    return newPath;
}

inline float3 pathThroughputGuide(float3 thp) { return f3(saturate(thp.x), saturate(thp.y), saturate(thp.z)); }

inline void StablePlanesHandleHit(const PathTracerCtx& x, PathState& path, float3 rayOrigin, float3 rayDir, float rayTCurrent, const SurfaceData& surfaceData, bool pathStopping)
{
    const RealtimeTargets& T = *x.sp;
    const uint vertexIndex = path.getVertexIndex(), currentSPIndex = path.getStablePlaneIndex(), px = path.id >> 16, py = path.id & 0xFFFF;
    const ShadingData& sd = surfaceData.sd;
    if (sd.psdBlockMotionVectorsAtSurface && path.GetMotionVectorSceneLength() == 0) path.SetMotionVectorSceneLength(path.sceneLength);
    if (vertexIndex == 1) T.StoreFirstHitRayLengthAndClearDominantToZero(px, py, path.sceneLength);

    bool setAsBase = true;
    if (vertexIndex < T.rt->maxStablePlaneVertexDepth && !pathStopping)
    {
        DeltaLobe deltaLobes[cMaxDeltaLobes]; int deltaLobeCount; float nonDeltaPart;
        surfaceData.bsdf.evalDeltaLobes(sd.frame(), deltaLobes, deltaLobeCount, nonDeltaPart);
        deltaLobeCount = std::max(int(cMaxDeltaLobes) - 1, deltaLobeCount);
        bool potentiallyVolumeTransmission = false;
        const float nonDeltaIgnoreThreshold = 1e-5f, deltaIgnoreThreshold = 0.001f;
        const bool hasNonDeltaLobes = nonDeltaPart > nonDeltaIgnoreThreshold;
        int nonZeroDeltaLobes[cMaxDeltaLobes] = { 0, 0, 0 }; int nonZeroDeltaLobeCount = 0;
        for (int k = 0; k < deltaLobeCount; k++)
            if (Average(deltaLobes[k].thp) > deltaIgnoreThreshold) { nonZeroDeltaLobes[nonZeroDeltaLobeCount++] = k; potentiallyVolumeTransmission |= deltaLobes[k].transmission != 0; }
        if (nonZeroDeltaLobeCount > 0)
        {
            bool allowPSR = T.rt->allowPrimarySurfaceReplacement && (nonZeroDeltaLobeCount == 1) && (currentSPIndex == 0) && !potentiallyVolumeTransmission;
            allowPSR &= !sd.psdBlockMotionVectorsAtSurface;
            bool canReuseExisting = (currentSPIndex != 0) && (nonZeroDeltaLobeCount > 0);
            canReuseExisting |= allowPSR;
            canReuseExisting &= !hasNonDeltaLobes;
            int availablePlaneCount = 0; int availablePlanes[3];
            T.GetAvailableEmptyPlanes(px, py, availablePlaneCount, availablePlanes);
            canReuseExisting &= (currentSPIndex == 0) || (sd.psdDominantDeltaLobeP1 > 0);
            nonZeroDeltaLobeCount = std::min(nonZeroDeltaLobeCount, availablePlaneCount + int(canReuseExisting));
            int lobeForReuse = -1;
            if (canReuseExisting) { lobeForReuse = nonZeroDeltaLobes[nonZeroDeltaLobeCount - 1]; nonZeroDeltaLobeCount--; }
            for (int i = 0; i < nonZeroDeltaLobeCount; i++)
            {
                const int lobeToExplore = nonZeroDeltaLobes[i];
                PathState splitPath = SplitDeltaPath(x, path, rayDir, surfaceData, deltaLobes[lobeToExplore], uint(lobeToExplore), true);
                splitPath.setStablePlaneIndex(uint(availablePlanes[i]));
                uint payload[20]; packPayload(splitPath, payload);
                memcpy(&T.planes[T.PixelToAddress(px, py, uint(availablePlanes[i]))], payload, 80);          // StoreExplorationStart
                T.SetBranchID(px, py, uint(availablePlanes[i]), cStablePlaneEnqueuedBranchID);
            }
            if (lobeForReuse != -1)
            {
                setAsBase = false;
                path = SplitDeltaPath(x, path, rayDir, surfaceData, deltaLobes[lobeForReuse], uint(lobeForReuse), nonZeroDeltaLobeCount > 0);
            }
        }
    }
    if (setAsBase)
    {
        float3 camO, camD; computeCameraRay(x, px, py, camO, camD);
        const mat3 imageXform = path.GetImageXform();
        const bool blockedAtSurface = path.GetMotionVectorSceneLength() != 0;
        const float sceneLengthForMVs = blockedAtSurface ? path.GetMotionVectorSceneLength() : path.sceneLength;
        const float3 virtualWorldPos = camO + camD * sceneLengthForMVs;
        const float3 worldMotion = surfaceData.prevPosW - sd.posW;     // actual world-space motion: instance.prevTransform / the previous-position stream (PathTracerStablePlanes.hlsli:286)
        const float3 virtualWorldMotion = mul(imageXform, worldMotion);
        const float3 motionVectors = T.computeMotionVector(virtualWorldPos, virtualWorldPos + virtualWorldMotion);
        float roughness = saturate(surfaceData.bsdf.data.roughness);
        const float3 worldNormal = normalize(mul(imageXform, sd.N));
        float3 diffBSDFEstimate, specBSDFEstimate;
        surfaceData.bsdf.estimateSpecDiffBSDF(diffBSDFEstimate, specBSDFEstimate, sd.N, sd.V);
        if (blockedAtSurface) roughness *= 0.25f * 0.95f;      // kSpecularRoughnessThreshold * 0.95
        const bool isDominant = path.hasFlag(PF_stablePlaneOnDominantBranch);
        T.StoreStablePlane(px, py, currentSPIndex, vertexIndex, rayOrigin, rayDir, path.stableBranchID, path.sceneLength, rayTCurrent, path.GetThp(), motionVectors, roughness, worldNormal,
                           diffBSDFEstimate, specBSDFEstimate, isDominant);
        if (isDominant) T.exportGuides(px, py, clipDepth(x.worldToClip, camO + camD * sceneLengthForMVs), motionVectors, Pack_R11G11B10_FLOAT(pathThroughputGuide(path.GetThp())));   // Bridge::ExportSurface
        path.terminate();
    }
}

inline void StablePlanesHandleMiss(const PathTracerCtx& x, PathState& path, float3 emission, float3 rayOrigin, float3 rayDir)
{
    const RealtimeTargets& T = *x.sp;
    const uint px = path.id >> 16, py = path.id & 0xFFFF, vertexIndex = path.getVertexIndex();
    if (vertexIndex == 1) T.StoreFirstHitRayLengthAndClearDominantToZero(px, py, kMaxRayTravel);
    float3 camO, camD; computeCameraRay(x, px, py, camO, camD);
    const bool blockedAtSurface = path.GetMotionVectorSceneLength() != 0;
    const float sceneLengthForMVs = blockedAtSurface ? path.GetMotionVectorSceneLength() : kEnvironmentMapSceneDistance;
    const float3 virtualWorldPos = camO + camD * sceneLengthForMVs;
    const float3 motionVectors = T.computeMotionVector(virtualWorldPos, virtualWorldPos);
    const bool isDominant = path.hasFlag(PF_stablePlaneOnDominantBranch);
    const float3 r = ReinhardMax(emission), skyAlbedo = f3(sqrtf(r.x), sqrtf(r.y), sqrtf(r.z));
    T.StoreStablePlane(px, py, path.getStablePlaneIndex(), vertexIndex, rayOrigin, rayDir, path.stableBranchID, blockedAtSurface ? sceneLengthForMVs : INFINITY, 0.0f, path.GetThp(), motionVectors,
                       blockedAtSurface ? 0.1f : 1.0f, -rayDir, skyAlbedo, blockedAtSurface ? f3(0.5f) : f3(0), isDominant);
    if (isDominant) T.exportGuides(px, py, clipDepth(x.worldToClip, virtualWorldPos), motionVectors, 0u);      // Bridge::ExportNonSurface
}

inline void CommitDenoiserRadiance(const PathTracerCtx& x, PathState& path)
{
    RtxptStablePlane& sp = x.sp->planes[x.sp->PixelToAddress(path.id >> 16, path.id & 0xFFFF, path.getStablePlaneIndex())];
    float4 accum = path.GetL();
    if (sp.PackedNoisyRadianceAndSpecAvg[0] != 0 && sp.PackedNoisyRadianceAndSpecAvg[1] != 0)
    {
        const float2 a = Fp16ToFp32(sp.PackedNoisyRadianceAndSpecAvg[0]), b = Fp16ToFp32(sp.PackedNoisyRadianceAndSpecAvg[1]);
        accum = accum + f4(a.x, a.y, b.x, b.y);
    }
    sp.PackedNoisyRadianceAndSpecAvg[0] = Fp32ToFp16(f2(accum.x, accum.y)); sp.PackedNoisyRadianceAndSpecAvg[1] = Fp32ToFp16(f2(accum.z, accum.w));
    path.SetL(f4(0, 0, 0, 0));
}

inline void StablePlanesOnScatter(const PathTracerCtx& x, PathState& path, const BSDFSample& bs)
{
    const RealtimeTargets& T = *x.sp;
    const uint px = path.id >> 16, py = path.id & 0xFFFF;
    if (path.hasFlag(PF_stablePlaneOnPlane)) path.setFlag(PF_stablePlaneBaseScatterDiff, (bs.lobe & Lobe_Diffuse) != 0);
    path.setFlag(PF_stablePlaneOnPlane, false);
    const uint nextVertexIndex = path.getVertexIndex() + 1;
    if (path.hasFlag(PF_stablePlaneOnBranch) && nextVertexIndex <= cStablePlaneMaxVertexIndex)
    {
        path.stableBranchID = StablePlanesAdvanceBranchID(path.stableBranchID, bs.getDeltaLobeIndex());
        bool onStablePath = false;
        for (uint spi = 0; spi < cStablePlaneCount; spi++)
        {
            const uint planeBranchID = T.GetBranchID(px, py, spi);
            if (planeBranchID == cStablePlaneInvalidBranchID) continue;
            if (StablePlaneIsOnPlane(planeBranchID, path.stableBranchID))
            {
                CommitDenoiserRadiance(x, path);
                path.setStablePlaneIndex(spi);
                path.setFlag(PF_stablePlaneOnDominantBranch, spi == T.LoadDominantIndex(px, py));
                path.setFlag(PF_stablePlaneOnPlane, true);
                path.setCounter(CTR_BouncesFromStablePlane, 0);
                onStablePath = true;
                break;
            }
            onStablePath |= StablePlaneIsOnStablePath(planeBranchID, StablePlanesVertexIndexFromBranchID(planeBranchID), path.stableBranchID, nextVertexIndex);
        }
        path.setFlag(PF_stablePlaneOnBranch, onStablePath);
    }
    else
    {
        path.stableBranchID = cStablePlaneInvalidBranchID;
        path.setFlag(PF_stablePlaneOnBranch, false);
        path.incrementCounter(CTR_BouncesFromStablePlane);
    }
    if (!path.hasFlag(PF_stablePlaneOnPlane)) path.incrementCounter(CTR_BouncesFromStablePlane);
}

struct RealtimeStats { uint64_t rays = 0; };

// postProcessHit of the BUILD pass (PathTracerSample.hlsl:96-113): a finished path continues with the pixel's next enqueued branch, if any
inline void postProcessHit(const PathTracerCtx& x, PathState& path)
{
    const RealtimeTargets& T = *x.sp; const uint px = path.id >> 16, py = path.id & 0xFFFF;
    int next;
    if (!path.isActive() && (next = T.FindNextToExplore(px, py, path.getStablePlaneIndex() + 1)) != -1)
    {
        uint payload[20]; memcpy(payload, &T.planes[T.PixelToAddress(px, py, uint(next))], 80);          // ExplorationStart
        T.SetBranchID(px, py, uint(next), cStablePlaneJustStartedID);
        path = unpackPayload(payload);
    }
}
// FirstHitFromVBuffer( path, 0 ) of the FILL pass (PathTracerSample.hlsl:33-93): the path restarts from stable plane 0; tMin / tMax bracket the surface the plane stands on
inline void FirstHitFromVBuffer(const PathTracerCtx& x, PathState& path, float& tMin, float& tMax)
{
    const RealtimeTargets& T = *x.sp; const uint px = path.id >> 16, py = path.id & 0xFFFF;
    {
        const RtxptStablePlane& sp = T.planes[T.PixelToAddress(px, py, 0)];
        const uint stableBranchID = T.GetBranchID(px, py, 0);
        float sceneLength = sp.SceneLength; const float lastRayTCurrent = sp.LastRayTCurrent;
        const uint vertexIndex = sp.VertexIndexAndRoughness >> 16;
        float3 thp, dummy; UnpackTwoFp32ToFp16(sp.PackedThpAndMVs, thp, dummy);
        bool isMiss = false;
        if (!std::isfinite(sceneLength)) { sceneLength = kMaxRayTravel; isMiss = true; }
        else { tMin = lastRayTCurrent * 0.99f; tMax = lastRayTCurrent * 1.01f; sceneLength -= lastRayTCurrent; }
        path.setVertexIndex(vertexIndex - 1);
        path.dir = f3(sp.RayDir[0], sp.RayDir[1], sp.RayDir[2]); path.origin = f3(sp.RayOrigin[0], sp.RayOrigin[1], sp.RayOrigin[2]);
        path.setFlag(PF_stablePlaneOnPlane, true); path.setFlag(PF_stablePlaneOnBranch, true);
        path.setStablePlaneIndex(0);
        path.stableBranchID = stableBranchID;
        path.SetThp(thp);
        path.SetL(f4(0, 0, 0, 0));
        path.setFlag(PF_stablePlaneOnDominantBranch, T.LoadDominantIndex(px, py) == 0);
        path.setCounter(CTR_BouncesFromStablePlane, 0);
        if (HasFinishedSurfaceBounces(x, path.getVertexIndex() + 1, path.getCounter(CTR_DiffuseBounces))) path.setFlag(PF_terminateAtNextBounce);
        path.rayCone = path.rayCone.propagateDistance(sceneLength);                      // UpdatePathTravelledLengthOnly
        path.sceneLength = std::min(path.sceneLength + sceneLength, kMaxRayTravel);
        if (isMiss) HandleMiss(x, path, path.dir, sceneLength);
    }
}

// RayGen of the BUILD pass for one pixel
inline void buildStablePlanesPixel(const PathTracerCtx& x, uint px, uint py)
{
    const RealtimeTargets& T = *x.sp;
    PathState path = EmptyPathInitialize(x, px, py, x.c->camera.PixelConeSpreadAngle);
    computeCameraRay(x, px, py, path.origin, path.dir);
    T.StartPixel(px, py); T.ExportSurfaceInit(px, py);
    while (path.isActive())
    {
        const float3 o = path.origin, d = path.dir;
        if (x.stats) x.stats->scatterRays++;
        Hit h = x.bvh->trace(*x.scene, o, d, 0.0f, kMaxRayTravel, false, x.stats ? &x.stats->nodeVisits : nullptr, x.stats ? &x.stats->triTests : nullptr);
        if (!h.valid()) HandleMiss(x, path, d, kMaxRayTravel);
        else HandleHit(x, path, o, d, h.t, x.bvh->tris[h.triId], f2(h.u, h.v));
        postProcessHit(x, path);
    }
}

// RayGen of the FILL pass for one pixel and one sub-sample (x.sampleIndex = sampleBaseIndex + subSampleIndex)
inline void fillStablePlanesPixel(const PathTracerCtx& x, uint px, uint py)
{
    const RealtimeTargets& T = *x.sp;
    PathState path = EmptyPathInitialize(x, px, py, x.c->camera.PixelConeSpreadAngle);
    computeCameraRay(x, px, py, path.origin, path.dir);
    float tMin = 0, tMax = kMaxRayTravel;
    FirstHitFromVBuffer(x, path, tMin, tMax);
    while (path.isActive())
    {
        const float3 o = path.origin, d = path.dir;
        if (x.stats) x.stats->scatterRays++;
        Hit h = x.bvh->trace(*x.scene, o, d, tMin, tMax, false, x.stats ? &x.stats->nodeVisits : nullptr, x.stats ? &x.stats->triTests : nullptr);
        if (!h.valid()) HandleMiss(x, path, d, kMaxRayTravel);
        else HandleHit(x, path, o, d, h.t, x.bvh->tris[h.triId], f2(h.u, h.v));
        tMin = 0; tMax = kMaxRayTravel;
    }
    CommitDenoiserRadiance(x, path);        // PathTracer::CommitPixel
}

} // namespace orc
