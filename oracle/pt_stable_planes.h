// ORACLE — test infrastructure only (see pt_math.h).
// pt_stable_planes.h: storage and helpers of the realtime-mode path-space decomposition ("stable planes"), restated from
//   Rtxpt/Shaders/PathTracer/StablePlanes.hlsli:29-45 (constants), :48-80 (StablePlane), :82-274 (StablePlanesContext), :277-315 (branch IDs)
//   Rtxpt/Shaders/PathTracer/Utils/Utils.hlsli:118-189 (octahedral packing, PackOrthoMatrix), :262-362 (Morton, GenericTS addressing)
//   Rtxpt/Shaders/PathTracer/Utils/Packing.hlsli:194-199 (PackTwoFp32ToFp16)
//   Rtxpt/Shaders/PathTracer/PathTracerHelpers.hlsli:227-262 (MatrixRotateFromTo)
//   Rtxpt/Shaders/PathTracerBridgeDonut.hlsli:890-909 (computeMotionVector), :1096-1175 (guide export)
// Branch-ID helpers and GenericTS addressing are pinned against the reference's own C++ halves of the same headers (oracle/_ref/ref_kat_host,
// tests/golden/host_golden.json); the plane packing, exploration payloads and the BUILD / FILL logic that uses them are pinned through tests/golden/hit_golden.npz
// (StablePlanes.hlsli and PathTracerStablePlanes.hlsli compiled in place).  computeMotionVector and the guide export restate the Donut bridge (unpinned).
#pragma once
#include <vector>
#include "pt_math.h"
#include "pt_lights.h"
#include "../include/rtxpt_b200.h"
#include <cmath>

namespace orc {

static const uint cStablePlaneCount = 3, cStablePlaneMaxVertexIndex = 15;
static const uint cStablePlaneInvalidBranchID = 0xFFFFFFFFu, cStablePlaneEnqueuedBranchID = 0xFFFFFFFEu, cStablePlaneJustStartedID = 0;
static const float kMaxSceneDistance = 50000.0f, kEnvironmentMapSceneDistance = 50000.0f * 100.0f;       // Config.h:84-85

inline uint StablePlanesAdvanceBranchID(uint prevStableBranchID, uint deltaLobeID) { return (prevStableBranchID << 2) | deltaLobeID; }
inline uint StablePlanesVertexIndexFromBranchID(uint stableBranchID) { uint v = stableBranchID, r = 0; while (v >>= 1) r++; return r / 2 + 1; }     // firstbithigh(id)/2+1
inline bool StablePlaneIsOnPlane(uint planeBranchID, uint vertexBranchID) { return planeBranchID == vertexBranchID; }
inline bool StablePlaneIsOnStablePath(uint planeBranchID, uint planeVertexIndex, uint vertexBranchID, uint vertexIndex)
{
    if (vertexIndex > planeVertexIndex) return false;
    return (planeBranchID >> ((planeVertexIndex - vertexIndex) * 2)) == vertexBranchID;
}

// ---- GenericTS addressing (8x8 tiles, Morton order inside a tile) ---------------------------------------------------------------------------
inline uint Morton16BitEncode(uint x, uint y)
{
    uint temp = (x & 0xff) | ((y & 0xff) << 16);
    temp = (temp ^ (temp << 4)) & 0x0f0f0f0f; temp = (temp ^ (temp << 2)) & 0x33333333; temp = (temp ^ (temp << 1)) & 0x55555555;
    return ((temp >> 15) | temp) & 0xffff;
}
inline uint GenericTSComputeLineStride(uint w, uint) { return ((w + 7) / 8) * 8; }
inline uint GenericTSComputePlaneStride(uint w, uint h) { return GenericTSComputeLineStride(w, h) * ((h + 7) / 8) * 8; }
inline uint GenericTSPixelToAddress(uint px, uint py, uint planeIndex, uint lineStride, uint planeStride)
{
    const uint xInTile = px % 8, yInTile = py % 8;
    return (px - xInTile) * 8 + (py - yInTile) * lineStride + Morton16BitEncode(xInTile, yInTile) + planeIndex * planeStride;
}

// ---- packing ---------------------------------------------------------------------------------------------------------------------------------
inline uint PackTwoFp32ToFp16(float a, float b) { return (f32tof16(clampf(a, -HLF_MAX, HLF_MAX)) << 16) | f32tof16(clampf(b, -HLF_MAX, HLF_MAX)); }
inline void PackTwoFp32ToFp16(float3 a, float3 b, uint out[3]) { out[0] = PackTwoFp32ToFp16(a.x, b.x); out[1] = PackTwoFp32ToFp16(a.y, b.y); out[2] = PackTwoFp32ToFp16(a.z, b.z); }
inline void UnpackTwoFp32ToFp16(const uint p[3], float3& a, float3& b)
{
    a = f3(f16tof32(p[0] >> 16), f16tof32(p[1] >> 16), f16tof32(p[2] >> 16)); b = f3(f16tof32(p[0] & 0xFFFF), f16tof32(p[1] & 0xFFFF), f16tof32(p[2] & 0xFFFF));
}
inline float2 OctWrap(float2 v) { return f2((1.0f - fabsf(v.y)) * (v.x >= 0.0f ? 1.0f : -1.0f), (1.0f - fabsf(v.x)) * (v.y >= 0.0f ? 1.0f : -1.0f)); }
inline float2 Encode_Oct(float3 n)
{
    n = n / (fabsf(n.x) + fabsf(n.y) + fabsf(n.z));
    float2 xy = n.z >= 0.0f ? f2(n.x, n.y) : OctWrap(f2(n.x, n.y));
    return f2(xy.x * 0.5f + 0.5f, xy.y * 0.5f + 0.5f);
}
inline float3 Decode_Oct(float2 f)
{
    f = f2(f.x * 2.0f - 1.0f, f.y * 2.0f - 1.0f);
    float3 n = f3(f.x, f.y, 1.0f - fabsf(f.x) - fabsf(f.y));
    const float t = saturate(-n.z);
    n.x += (n.x >= 0.0f) ? -t : t; n.y += (n.y >= 0.0f) ? -t : t;
    return normalize(n);
}
// NDirToOctUnorm32 / OctToNDirUnorm32: pt_lights.h
inline uint NDirToOctUnorm30(float3 n)
{
    float2 p = Encode_Oct(n); p = f2(saturate(p.x * 0.5f + 0.5f), saturate(p.y * 0.5f + 0.5f));
    return (uint(p.x * float(0x7fff) + 0.5f) & 0x7fff) | ((uint(p.y * float(0x7fff) + 0.5f) & 0x7fff) << 15);
}
inline float3 OctToNDirUnorm30(uint u)
{
    float2 p = f2(saturate(float(u & 0x7fff) / float(0x7fff)), saturate(float(u >> 15) / float(0x7fff)));
    return Decode_Oct(f2(p.x * 2.0f - 1.0f, p.y * 2.0f - 1.0f));
}

struct mat3 { float3 r[3]; };       // rows, like an HLSL float3x3
inline mat3 identity3() { mat3 m; m.r[0] = f3(1, 0, 0); m.r[1] = f3(0, 1, 0); m.r[2] = f3(0, 0, 1); return m; }
inline float3 col(const mat3& m, int c) { return c == 0 ? f3(m.r[0].x, m.r[1].x, m.r[2].x) : (c == 1 ? f3(m.r[0].y, m.r[1].y, m.r[2].y) : f3(m.r[0].z, m.r[1].z, m.r[2].z)); }
inline mat3 mul(const mat3& a, const mat3& b) { mat3 o; for (int i = 0; i < 3; i++) o.r[i] = f3(dot(a.r[i], col(b, 0)), dot(a.r[i], col(b, 1)), dot(a.r[i], col(b, 2))); return o; }
inline float3 mul(const mat3& m, float3 v) { return f3(dot(m.r[0], v), dot(m.r[1], v), dot(m.r[2], v)); }
inline mat3 transpose(const mat3& m) { mat3 o; o.r[0] = col(m, 0); o.r[1] = col(m, 1); o.r[2] = col(m, 2); return o; }
inline mat3 lp(const mat3& m) { mat3 o; for (int i = 0; i < 3; i++) o.r[i] = lp(m.r[i]); return o; }
inline void PackOrthoMatrix(const mat3& x, uint out[2])
{
    const uint handedness = dot(cross(x.r[0], x.r[1]), x.r[2]) > 0 ? 1u : 0u;
    out[0] = NDirToOctUnorm30(x.r[0]); out[1] = NDirToOctUnorm30(x.r[1]) | (handedness << 31);
}
inline mat3 UnpackOrthoMatrix(const uint packed[2])
{
    mat3 x; const uint handedness = packed[1] >> 31;
    x.r[0] = OctToNDirUnorm30(packed[0]); x.r[1] = OctToNDirUnorm30(packed[1] & 0x7FFFFFFFu);
    x.r[2] = handedness ? cross(x.r[0], x.r[1]) : cross(x.r[1], x.r[0]);
    return x;
}
inline mat3 MatrixRotateFromTo(float3 from, float3 to)       // columnMajor = true branch
{
    const float e = dot(from, to), f = fabsf(e);
    if (f > float(1.0f - 1e-10f)) return identity3();
    const float3 v = cross(from, to);
    const float h = 1.0f / (1.0f + e), hvx = h * v.x, hvz = h * v.z, hvxy = hvx * v.y, hvxz = hvx * v.z, hvyz = hvz * v.y;
    mat3 m;
    m.r[0] = f3(e + hvx * v.x, hvxy - v.z, hvxz + v.y);
    m.r[1] = f3(hvxy + v.z, e + h * v.y * v.y, hvyz - v.x);
    m.r[2] = f3(hvxz - v.y, hvyz + v.x, e + hvz * v.z);
    return m;
}
inline float3 ReinhardMax(float3 color)
{
    const float luminance = std::max(1e-7f, std::max(std::max(color.x, color.y), color.z));
    const float reinhard = luminance / (luminance + 1);
    return color * (reinhard / luminance);
}

// ---- the realtime render targets ---------------------------------------------------------------------------------------------------------------
struct RealtimeTargets
{
    uint width = 0, height = 0, lineStride = 0, planeStride = 0;
    RtxptStablePlane* planes = nullptr;     // [3 * planeStride]
    uint* header = nullptr;                 // [4][height][width]
    uint16_t* stableRadiance = nullptr;     // RGBA16F
    float* depth = nullptr; uint16_t* motionVectors = nullptr; uint* throughput = nullptr; float* specularHitT = nullptr;
    const RtxptRealtimeConstants* rt = nullptr;

    uint PixelToAddress(uint px, uint py, uint plane) const { return GenericTSPixelToAddress(px, py, plane, lineStride, planeStride); }
    uint& hdr(uint px, uint py, uint layer) const { return header[(size_t(layer) * height + py) * width + px]; }
    uint GetBranchID(uint px, uint py, uint plane) const { return hdr(px, py, plane); }
    void SetBranchID(uint px, uint py, uint plane, uint id) const { hdr(px, py, plane) = id; }
    void StoreFirstHitRayLengthAndClearDominantToZero(uint px, uint py, float length) const { hdr(px, py, 3) = asuint(std::min(kMaxRayTravel, length)) & 0xFFFFFFFCu; }
    void StoreDominantIndex(uint px, uint py, uint index) const { hdr(px, py, 3) = (hdr(px, py, 3) & 0xFFFFFFFCu) | (3u & index); }
    uint LoadDominantIndex(uint px, uint py) const { return hdr(px, py, 3) & 3u; }
    uint activePlaneCount() const { return rt->activeStablePlaneCount; }

    float3 LoadStableRadiance(uint px, uint py) const { const uint16_t* p = stableRadiance + (size_t(py) * width + px) * 4; return f3(f16tof32(p[0]), f16tof32(p[1]), f16tof32(p[2])); }
    void StoreStableRadiance(uint px, uint py, float3 r) const
    {   // RGBA16F UAV store of clamp(radiance, 0, HLF_MAX)
        uint16_t* p = stableRadiance + (size_t(py) * width + px) * 4; r = clamp3(r, 0, HLF_MAX);
        p[0] = uint16_t(f32tof16(r.x)); p[1] = uint16_t(f32tof16(r.y)); p[2] = uint16_t(f32tof16(r.z)); p[3] = 0;
    }
    void AccumulateStableRadiance(uint px, uint py, float3 r) const
    {   // StableRadianceUAV[pixelPos].xyz += radiance: read fp16, add in fp32, store fp16 (alpha untouched)
        uint16_t* p = stableRadiance + (size_t(py) * width + px) * 4;
        p[0] = uint16_t(f32tof16(f16tof32(p[0]) + r.x)); p[1] = uint16_t(f32tof16(f16tof32(p[1]) + r.y)); p[2] = uint16_t(f32tof16(f16tof32(p[2]) + r.z));
    }
    void StartPixel(uint px, uint py) const
    {
        StoreStableRadiance(px, py, f3(0));
        for (uint i = 0; i < 3; i++) hdr(px, py, i) = cStablePlaneInvalidBranchID;
    }
    void ExportSurfaceInit(uint px, uint py) const { depth[size_t(py) * width + px] = 0; specularHitT[size_t(py) * width + px] = 0; }

    void StoreStablePlane(uint px, uint py, uint planeIndex, uint vertexIndex, float3 rayOrigin, float3 rayDir, uint stableBranchID, float sceneLength, float rayTCurrent,
                          float3 thp, float3 motionVectors, float roughness, float3 worldNormal, float3 diffBSDFEstimate, float3 specBSDFEstimate, bool dominantSP) const
    {
        RtxptStablePlane sp;
        sp.RayOrigin[0] = rayOrigin.x; sp.RayOrigin[1] = rayOrigin.y; sp.RayOrigin[2] = rayOrigin.z;
        sp.RayDir[0] = rayDir.x; sp.RayDir[1] = rayDir.y; sp.RayDir[2] = rayDir.z;
        sp.SceneLength = sceneLength;
        sp.VertexIndexAndRoughness = (vertexIndex << 16) | f32tof16(roughness);
        PackTwoFp32ToFp16(thp, motionVectors, sp.PackedThpAndMVs);
        const float kNRDMinReflectance = 0.04f, kNRDMaxReflectance = 6.5504e+4f;
        PackTwoFp32ToFp16(clamp3(diffBSDFEstimate, kNRDMinReflectance, kNRDMaxReflectance), clamp3(specBSDFEstimate, kNRDMinReflectance, kNRDMaxReflectance), sp.DenoiserPackedBSDFEstimate);
        sp.PackedNormal = NDirToOctUnorm32(worldNormal);
        sp.PackedNoisyRadianceAndSpecAvg[0] = Fp32ToFp16(f2(0, 0)); sp.PackedNoisyRadianceAndSpecAvg[1] = Fp32ToFp16(f2(0, 0));
        sp.LastRayTCurrent = rayTCurrent;
        sp.FlagsAndVertexIndex = 0; sp.PackedCounters = 0;
        planes[PixelToAddress(px, py, planeIndex)] = sp;
        SetBranchID(px, py, planeIndex, stableBranchID);
        if (dominantSP && planeIndex != 0) StoreDominantIndex(px, py, planeIndex);
    }
    int FindNextToExplore(uint px, uint py, uint fromPlane) const
    {
        for (uint i = fromPlane; i < cStablePlaneCount; i++) if (GetBranchID(px, py, i) == cStablePlaneEnqueuedBranchID) return int(i);
        return -1;
    }
    void GetAvailableEmptyPlanes(uint px, uint py, int& availableCount, int availablePlanes[3]) const
    {
        availableCount = 0;
        for (uint i = 1; i < std::min(activePlaneCount(), cStablePlaneCount); i++) if (GetBranchID(px, py, i) == cStablePlaneInvalidBranchID) availablePlanes[availableCount++] = int(i);
    }
    // Bridge::computeMotionVector
    float3 (*motionVectorOverride)(float3 posW, float3 prevPosW) = nullptr;       // test hook (oracle.cpp's known-answer mirrors): the stub bridge's closed form
    float3 computeMotionVector(float3 posW, float3 prevPosW) const
    {
        if (motionVectorOverride) return motionVectorOverride(posW, prevPosW);
        auto xf = [](const float* M, float3 p, float out[4]) { for (int c = 0; c < 4; c++) out[c] = ((p.x * M[c] + p.y * M[4 + c]) + p.z * M[8 + c]) + M[12 + c]; };
        float clip[4], prev[4]; xf(rt->matWorldToClipNoOffset, posW, clip); xf(rt->prevMatWorldToClipNoOffset, prevPosW, prev);
        const float cx = clip[0] / clip[3], cy = clip[1] / clip[3], pxx = prev[0] / prev[3], pyy = prev[1] / prev[3];
        if (clip[3] <= 0 || prev[3] <= 0) return f3(0);
        return f3((pxx - cx) * rt->clipToWindowScale[0], (pyy - cy) * rt->clipToWindowScale[1], prev[3] - clip[3]);
    }
    void exportGuides(uint px, uint py, float depthValue, float3 motion, uint packedThroughput) const
    {
        const size_t pix = size_t(py) * width + px;
        uint16_t* mv = motionVectors + pix * 4; mv[0] = uint16_t(f32tof16(motion.x)); mv[1] = uint16_t(f32tof16(motion.y)); mv[2] = uint16_t(f32tof16(motion.z)); mv[3] = 0;
        depth[pix] = depthValue; throughput[pix] = packedThroughput;
    }
    float3 GetNoisyRadiance(const RtxptStablePlane& sp) const { float2 a = Fp16ToFp32(sp.PackedNoisyRadianceAndSpecAvg[0]), b = Fp16ToFp32(sp.PackedNoisyRadianceAndSpecAvg[1]); return f3(a.x, a.y, b.x); }
    // StablePlanesContext::GetAllRadiance = PostProcess NO_DENOISER_FINAL_MERGE
    float3 GetAllRadiance(uint px, uint py) const
    {
        float3 pathL = LoadStableRadiance(px, py);
        for (uint i = 0; i < cStablePlaneCount; i++)
        {
            if (GetBranchID(px, py, i) == cStablePlaneInvalidBranchID) continue;
            pathL = pathL + GetNoisyRadiance(planes[PixelToAddress(px, py, i)]);
        }
        return pathL;
    }
};

} // namespace orc

// ---- RTXPT's side of the denoiser interface: PostProcess.hlsl DENOISER_PREPARE_INPUTS (ReBLUR variant, :444-570) and
// DENOISER_FINAL_MERGE (:577-690), NRD front/back-end packing (External/Nrd/Shaders/Include/NRD.hlsli:327-345, :362-381, :526-529, :646-677, :728-750, :869-874),
// Rtxpt/NRD/DenoiserNRD.hlsli:34-52.  Test infrastructure like the rest of this directory.
namespace orc {

struct DenoiserTargets
{
    float* viewZ; uint16_t* motion; uint32_t* normalRoughness; uint16_t* diffRadianceHitDist; uint16_t* specRadianceHitDist; uint8_t* disocclusionMix; uint8_t* historyClampRelax;
    uint16_t* outputColor;       // RGBA16F
};
inline uint8_t unorm8(float v) { return uint8_t(saturate(v) * 255.0f + 0.5f); }
inline float3 NRD_LinearToYCoCg(float3 c) { return f3(dot(c, f3(0.25f, 0.5f, 0.25f)), dot(c, f3(0.5f, 0.0f, -0.5f)), dot(c, f3(-0.25f, 0.5f, -0.25f))); }
inline float3 NRD_YCoCgToLinear(float3 c) { const float t = c.x - c.z; return max3v(f3(t + c.y, c.x + c.z, t - c.y), f3(0)); }
inline float2 NRD_EncodeUnitVectorUnsigned(float3 v)
{
    v = v / (fabsf(v.x) + fabsf(v.y) + fabsf(v.z));
    const float2 wrap = f2((1.0f - fabsf(v.y)) * ((v.x >= 0.0f ? 1.0f : 0.0f) * 2.0f - 1.0f), (1.0f - fabsf(v.x)) * ((v.y >= 0.0f ? 1.0f : 0.0f) * 2.0f - 1.0f));
    const float2 xy = v.z >= 0.0f ? f2(v.x, v.y) : wrap;
    return f2(xy.x * 0.5f + 0.5f, xy.y * 0.5f + 0.5f);
}
inline uint32_t packR10G10B10A2(float x, float y, float z, float w)
{
    return uint32_t(saturate(x) * 1023.0f + 0.5f) | (uint32_t(saturate(y) * 1023.0f + 0.5f) << 10) | (uint32_t(saturate(z) * 1023.0f + 0.5f) << 20) | (uint32_t(saturate(w) * 3.0f + 0.5f) << 30);
}
inline float REBLUR_GetHitDistanceNormalization(float viewZ, const float* hp, float roughness) { return (hp[0] + fabsf(viewZ) * hp[1]) * lerp(1.0f, hp[2], saturate(exp2f(hp[3] * roughness * roughness))); }
inline void storeRGBA16F(uint16_t* dst, float4 v) { dst[0] = uint16_t(f32tof16(v.x)); dst[1] = uint16_t(f32tof16(v.y)); dst[2] = uint16_t(f32tof16(v.z)); dst[3] = uint16_t(f32tof16(v.w)); }
inline void NRDRadianceClamp(float3& radiance, float preExposedGrayLuminance, float rangeK)
{
    const float kClampMax = std::min(255.0f, preExposedGrayLuminance * rangeK);
    const float lum = Luminance(radiance);
    if (lum > kClampMax) radiance = radiance * (kClampMax / lum);
}

inline void denoiserPrepareInputsPixel(const RealtimeTargets& T, const DenoiserTargets& D, const RtxptDenoiserConstants& k, uint px, uint py, uint stablePlaneIndex, bool initWithStableRadiance,
                                       float3 camOrigin, float3 camDir)
{
    const size_t pix = size_t(py) * T.width + px;
    if (initWithStableRadiance) { const float3 s = T.LoadStableRadiance(px, py); storeRGBA16F(D.outputColor + pix * 4, f4(s, 1.0f)); D.historyClampRelax[pix] = 0; }
    bool hasSurface = false;
    const uint spBranchID = T.GetBranchID(px, py, stablePlaneIndex);
    if (spBranchID != cStablePlaneInvalidBranchID)
    {
        const RtxptStablePlane& sp = T.planes[T.PixelToAddress(px, py, stablePlaneIndex)];
        if (std::isfinite(sp.SceneLength))
        {
            hasSurface = true;
            float3 diffEstimate, specEstimate; UnpackTwoFp32ToFp16(sp.DenoiserPackedBSDFEstimate, diffEstimate, specEstimate);
            const float3 virtualWorldPos = camOrigin + camDir * sp.SceneLength;
            const float* M = k.matWorldToView;
            const float virtualViewspaceZ = ((virtualWorldPos.x * M[2] + virtualWorldPos.y * M[6]) + virtualWorldPos.z * M[10]) + M[14];
            float3 thp, motionVectors; UnpackTwoFp32ToFp16(sp.PackedThpAndMVs, thp, motionVectors);
            D.viewZ[pix] = virtualViewspaceZ;
            storeRGBA16F(D.motion + pix * 4, f4(motionVectors, 0));
            const float spRoughness = f16tof32(sp.VertexIndexAndRoughness & 0xFFFF);
            float finalRoughness = std::max(0.2f, spRoughness);
            float specularSuppressionMul = 1.0f;
            if (stablePlaneIndex == 0 && k.stablePlanesSuppressPrimaryIndirectSpecularK != 0.0f && T.activePlaneCount() > 1)
            {
                bool shouldSuppress = true;
                for (uint i = 1; i < T.activePlaneCount(); i++) shouldSuppress &= T.GetBranchID(px, py, i) != cStablePlaneInvalidBranchID;
                if (shouldSuppress) specularSuppressionMul = saturate(1 - k.stablePlanesSuppressPrimaryIndirectSpecularK);
            }
            float disocclusionRelax = 0.0f;
            if (StablePlanesVertexIndexFromBranchID(spBranchID) > 1)
            {   // ComputeDisocclusionRelaxation: how much the (virtual) normal turns towards the four neighbours
                const float3 rayDirC = OctToNDirUnorm32(sp.PackedNormal);
                const int off[4][2] = { { -1, 0 }, { 1, 0 }, { 0, -1 }, { 0, 1 } };
                for (int n = 0; n < 4; n++)
                {
                    const uint nx = uint(std::min(std::max(int(px) + off[n][0], 0), int(T.width) - 1)), ny = uint(std::min(std::max(int(py) + off[n][1], 0), int(T.height) - 1));
                    if (T.GetBranchID(nx, ny, stablePlaneIndex) == cStablePlaneInvalidBranchID) disocclusionRelax += 0.02f;
                    else disocclusionRelax += 1 - dot(rayDirC, OctToNDirUnorm32(T.planes[T.PixelToAddress(nx, ny, stablePlaneIndex)].PackedNormal));
                }
                disocclusionRelax = saturate((disocclusionRelax - 0.00002f) * 25);
            }
            D.disocclusionMix[pix] = unorm8(disocclusionRelax);
            D.historyClampRelax[pix] = unorm8(saturate(float(D.historyClampRelax[pix]) / 255.0f + disocclusionRelax * saturate(Luminance(thp))));
            finalRoughness = saturate(finalRoughness + disocclusionRelax);
            // StablePlane::GetNoisyDiffRadiance / GetNoisySpecRadiance (StablePlanes.hlsli:66-67)
            const float2 a = Fp16ToFp32(sp.PackedNoisyRadianceAndSpecAvg[0]), b = Fp16ToFp32(sp.PackedNoisyRadianceAndSpecAvg[1]);
            const float3 l = f3(a.x, a.y, b.x); const float specAvg = b.y, totalAvg = Average(l);
            float3 diff = l * saturate(1.0f - specAvg / (totalAvg + 1e-12f)), spec = l * saturate(specAvg / (totalAvg + 1e-12f));
            diff = diff / diffEstimate; spec = spec / specEstimate;
            spec = spec * specularSuppressionMul;
            D.normalRoughness[pix] = [&] { const float2 e = NRD_EncodeUnitVectorUnsigned(OctToNDirUnorm32(sp.PackedNormal)); return packR10G10B10A2(e.x, e.y, finalRoughness, 0.0f); }();
            NRDRadianceClamp(diff, k.preExposedGrayLuminance, k.denoiserRadianceClampK * 16); NRDRadianceClamp(spec, k.preExposedGrayLuminance, k.denoiserRadianceClampK * 16);
            float specHitT = 0;
            if (T.LoadDominantIndex(px, py) == stablePlaneIndex) specHitT = T.specularHitT[pix];
            auto pack = [](float3 radiance, float normHitDist) {     // REBLUR_FrontEnd_PackRadianceAndNormHitDist(sanitize = true)
                const bool invalid = !std::isfinite(radiance.x) || !std::isfinite(radiance.y) || !std::isfinite(radiance.z);
                radiance = invalid ? f3(0) : clamp3(radiance, 0, 65504.0f);
                normHitDist = std::isfinite(normHitDist) ? saturate(normHitDist) : 0.0f;
                return f4(NRD_LinearToYCoCg(radiance), normHitDist); };
            storeRGBA16F(D.diffRadianceHitDist + pix * 4, pack(diff, 0.0f));
            const float specNorm = saturate(specHitT / REBLUR_GetHitDistanceNormalization(virtualViewspaceZ, k.hitDistanceParameters, spRoughness));
            storeRGBA16F(D.specRadianceHitDist + pix * 4, pack(spec, specNorm));
        }
    }
    if (!hasSurface) D.viewZ[pix] = 3.402823466e+38f;       // VIEWZ_SKY_MARKER
}

inline void denoiserFinalMergePixel(const RealtimeTargets& T, const DenoiserTargets& D, uint px, uint py, uint stablePlaneIndex, const uint16_t* denoisedDiff, const uint16_t* denoisedSpec)
{
    const size_t pix = size_t(py) * T.width + px;
    if (D.viewZ[pix] == 3.402823466e+38f) return;
    float3 diffEstimate, specEstimate; UnpackTwoFp32ToFp16(T.planes[T.PixelToAddress(px, py, stablePlaneIndex)].DenoiserPackedBSDFEstimate, diffEstimate, specEstimate);
    auto load = [&](const uint16_t* p) { return f3(f16tof32(p[pix * 4]), f16tof32(p[pix * 4 + 1]), f16tof32(p[pix * 4 + 2])); };
    const float3 diff = NRD_YCoCgToLinear(load(denoisedDiff)) * diffEstimate, spec = NRD_YCoCgToLinear(load(denoisedSpec)) * specEstimate;
    uint16_t* o = D.outputColor + pix * 4;
    const float3 sum = max3v(diff + spec, f3(0));
    o[0] = uint16_t(f32tof16(f16tof32(o[0]) + sum.x)); o[1] = uint16_t(f32tof16(f16tof32(o[1]) + sum.y)); o[2] = uint16_t(f32tof16(f16tof32(o[2]) + sum.z));
}


// DenoisingGuidesBaker::DenoiseSpecHitT (Rtxpt/ProcessingPasses/DenoisingGuidesBaker.hlsl:53-97, .cpp:62-84; run by Sample::PathTrace after the FILL pass, Sample.cpp:2541-2543):
// the specular hit distance guide is spread over a 5x5 neighbourhood of similar depth - pixels without a value borrow the neighbours' mean, pixels with one are capped at
// 1.5 x + 0.5 of their own.  One ping (guide -> scratch) and one pong (scratch -> guide).
inline float SpecHitTNeighbourhood(const float* src, const float* depth, int W, int H, int px, int py)
{
    const float centerD = depth[size_t(py) * W + px];
    float prevHitT = std::max(0.0f, src[size_t(py) * W + px]);
    if (prevHitT < 5e-2f) prevHitT = 0;
    float vAvg = prevHitT, sumW = prevHitT > 0 ? 1.0f : 0.0f;
    for (int x = -2; x <= 2; x++) for (int y = -2; y <= 2; y++)
    {
        if (x == 0 && y == 0) continue;
        const int nx = px + x, ny = py + y;
        if (nx < 0 || ny < 0 || nx >= W || ny >= H) continue;
        const float v = std::min(src[size_t(ny) * W + nx], 65504.0f), d = std::max(0.0f, depth[size_t(ny) * W + nx]);
        float weight = v > 0 ? 1.0f : 0.0f;
        weight *= fabsf(d - centerD) <= (d + centerD + 1e-5f) * 0.025f ? 1.0f : 0.0f;
        if (weight > 0) { vAvg += v * weight; sumW += weight; }
    }
    if (sumW == 0) return prevHitT;
    vAvg /= sumW;
    return prevHitT <= 0 ? vAvg : std::min(prevHitT * 1.5f + 0.5f, vAvg);
}
inline void denoiseSpecHitT(float* specHitT, const float* depth, int W, int H)
{
    std::vector<float> scratch(size_t(W) * H);
    for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) scratch[size_t(y) * W + x] = SpecHitTNeighbourhood(specHitT, depth, W, H, x, y);
    for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) specHitT[size_t(y) * W + x] = SpecHitTNeighbourhood(scratch.data(), depth, W, H, x, y);
}

} // namespace orc
