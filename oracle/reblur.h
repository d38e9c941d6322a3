// ORACLE — test infrastructure only (see pt_math.h).
// reblur.h: CPU restatement of NRD's REBLUR_DIFFUSE_SPECULAR denoiser (NRD 4.15.2 as vendored under External/Nrd), SURVEY §8 row a18 / K9 — SPATIAL HALF:
//   ClassifyTiles            External/Nrd/Shaders/Source/REBLUR_ClassifyTiles.cs.hlsl:21-60
//   HitDistReconstruction    External/Nrd/Shaders/Include/REBLUR_HitDistReconstruction.hlsli:11-155 (5x5, RTXPT's setting: Rtxpt/NRD/NrdConfig.cpp:49-61)
//   PrePass / Blur / PostBlur External/Nrd/Shaders/Include/REBLUR_PrePass.hlsli, REBLUR_Blur.hlsli, REBLUR_PostBlur.hlsli with the shared
//                            REBLUR_Common_DiffuseSpatialFilter.hlsli / REBLUR_Common_SpecularSpatialFilter.hlsli (screen-space diffuse, world-space specular kernels)
//   helpers                  Common.hlsli:222-560, REBLUR_Common.hlsli:13-130, :262-290, REBLUR_Config.hlsli, NRD.hlsli:327-410, :526-529, :606-640
//   constants                External/Nrd/Source/Reblur.cpp:280-392 (AddSharedConstants_Reblur), InstanceImpl.cpp:331-451 (rotators, frustum, unproject)
// The temporal passes (TemporalAccumulation, HistoryFix, TemporalStabilization) are not restated yet; the spatial passes take the per-pixel accumulated
// frame counts ("data1") they would produce as an input.
// THIRD-PARTY CODE ABSENT FROM /root/reference: NRD's shaders call NVIDIA MathLib (ml.hlsli: Math::, Geometry::, ImportanceSampling::, Sequence::, Rng::, Filtering::,
// Packing::; an un-vendored submodule of NRD at the revision NRD 4.15.2 pins).  Its functions are restated below from their published definitions and from what the
// call sites require geometrically (view-space reconstruction, rotators, bases); the GGX dominant direction is the copy NRD.hlsli carries (:392-406).  Parity unpinned.
#pragma once
#include "pt_math.h"
#include <vector>
#include <cmath>

namespace orc { namespace reblur {

// ---- MathLib restatement -----------------------------------------------------------------------------------------------------------------------
static const float NRD_EPS = 1e-6f, NRD_INF = 1e6f;
// HLSL saturate maps NaN to 0 - the spatial filters rely on it when a tap lands on a sky texel (viewZ = FLT_MAX makes the plane distance inf * 0)
inline float saturate(float x) { return x > 0.0f ? (x < 1.0f ? x : 1.0f) : 0.0f; }
inline float LinearStep(float a, float b, float x) { return saturate((x - a) / (b - a)); }
inline float SmoothStep01(float x) { x = saturate(x); return x * x * (3.0f - 2.0f * x); }
inline float SmoothStep(float a, float b, float x) { return SmoothStep01(LinearStep(a, b, x)); }
inline float Pow01(float x, float y) { return powf(saturate(x), y); }
inline float Sqrt01(float x) { return sqrtf(saturate(x)); }
inline float PositiveRcp(float x) { return 1.0f / std::max(x, 1.175494351e-38f); }
inline float AcosApprox(float x) { return sqrtf(2.0f) * sqrtf(saturate(1.0f - x)); }
inline float Pow5(float x) { return powf(saturate(1.0f - x), 5.0f); }                        // BRDF::Pow5
struct Rotator { float x, y, z, w; };                                                          // ( cos, sin, -sin, cos )
inline Rotator GetRotator(float angle) { const float ca = cosf(angle), sa = sinf(angle); return { ca, sa, -sa, ca }; }
inline Rotator CombineRotators(Rotator a, Rotator b) { return { a.x * b.x + a.z * b.y, a.y * b.x + a.w * b.y, a.x * b.z + a.z * b.w, a.y * b.z + a.w * b.w }; }   // r1.xyxy * r2.xxzz + r1.zwzw * r2.yyww
inline float2 RotateVector(Rotator r, float2 v) { return f2(v.x * r.x + v.y * r.y, v.x * r.z + v.y * r.w); }                                                      // v.x * r.xz + v.y * r.yw
inline Rotator ScaleRotator(Rotator r, float2 s) { return { r.x * s.x, r.y * s.x, r.z * s.y, r.w * s.y }; }
inline void GetBasis(float3 N, float3& T, float3& B)
{   // branchless orthonormal basis (Duff et al. 2017)
    const float sz = N.z >= 0.0f ? 1.0f : -1.0f, a = 1.0f / (sz + N.z), ya = N.y * a, b = N.x * ya, c = N.x * sz;
    T = f3(c * N.x * a - 1.0f, sz * b, c); B = f3(b, N.y * ya - sz, N.y);
}
inline float Weyl1D(float p, uint n) { const float v = p + float(n) * 0.6180339887498948f; return v - floorf(v); }
inline float Bayer4x4_00(uint frameIndex) { return float(frameIndex & 15u) / 16.0f; }          // Sequence::Bayer4x4( uint2( 0, 0 ), frameIndex ): the matrix entry at (0,0) is 0
struct HashRng          // Rng::Hash: a per-pixel integer hash stream (the exact mixer is MathLib's; any well-mixed stream is statistically equivalent)
{
    uint s;
    void Initialize(uint px, uint py, uint frameIndex) { s = Hash32Combine(Hash32Combine(Hash32(px), py), frameIndex); }
    float GetFloat() { s = Hash32(s); return Hash32ToFloat(s); }
};
// ImportanceSampling::GetSpecularDominantDirection( N, V, roughness, ML_SPECULAR_DOMINANT_DIRECTION_G2 ): factor as in NRD.hlsli:392-398
inline float4 GetSpecularDominantDirection(float3 N, float3 V, float roughness)
{
    const float NoV = fabsf(dot(N, V));
    const float a = 0.298475f * logf(39.4115f - 39.0029f * roughness);
    const float f = saturate(powf(saturate(1.0f - NoV), 10.8649f) * (1.0f - a) + a);
    const float3 R = N * (2.0f * dot(N, V)) - V;                     // reflect( -V, N )
    return f4(normalize(lerp(N, R, f)), f);
}
// ImportanceSampling::GetSpecularLobeTanHalfAngle: tangent of the half angle that encloses `percentOfVolume` of the GGX lobe, tan = m * sqrt( p / ( 1 - p ) ), m = roughness^2
// ( Reblur.cpp:368 squares lobeAngleFraction "because GetSpecularLobeTanHalfAngle has been fixed" - the square root here is that fix )
inline float GetSpecularLobeTanHalfAngle(float roughness, float percentOfVolume)
{
    roughness = saturate(roughness); percentOfVolume = saturate(percentOfVolume);
    const float m = roughness * roughness;
    return m * sqrtf(percentOfVolume / (1.0f - percentOfVolume + NRD_EPS));
}

// ---- settings and per-frame constants ----------------------------------------------------------------------------------------------------------------
struct Settings         // nrd::ReblurSettings defaults (NRDSettings.h:230-300) with RTXPT's overrides (Rtxpt/NRD/NrdConfig.cpp:49-61) and nrd::CommonSettings as NrdIntegration.cpp:375-408 sets them
{
    float hitDistParams[4] = { 3.0f, 0.1f, 20.0f, -25.0f };
    uint maxAccumulatedFrameNum = 50, maxFastAccumulatedFrameNum = 6, historyFixFrameNum = 3;
    float diffusePrepassBlurRadius = 15.0f, specularPrepassBlurRadius = 40.0f;
    float minHitDistanceWeight = 0.1f, minBlurRadius = 1.0f, maxBlurRadius = 30.0f, lobeAngleFraction = 0.15f, roughnessFraction = 0.15f, planeDistanceSensitivity = 0.02f;
    float minMaterialForDiffuse = 4.0f, minMaterialForSpecular = 4.0f;
    bool usePrepassOnlyForSpecularMotionEstimation = false;
    float denoisingRange = 100000.0f;       // kMaxSceneDistance * 2
    float viewZScale = 1.0f;
};
struct Constants
{
    uint W = 0, H = 0, frameIndex = 0;
    float viewToWorld[9];                   // rotation rows: view x,y,z axes in world space (camera-relative, translation removed)
    float viewToClip[16];                   // row-major, row vector x matrix, D3D clip space
    float frustum[4]; float unproject, minRectDimMulUnproject, orthoMode = 0.0f;
    Rotator rotatorPre, rotator, rotatorPost;
    Settings s;
    float gLobeAngleFraction, gMaxBlurRadius, gDiffPrepassBlurRadius, gSpecPrepassBlurRadius;
};
// worldToView / viewToClip: row-major, row vector x matrix (the convention of include/rtxpt_b200.h); view space is left-handed, +z forward, as RTXPT's (Donut's) is
inline Constants makeConstants(const Settings& s, uint W, uint H, const float* worldToView, const float* viewToClip, uint frameIndex)
{
    Constants c; c.s = s; c.W = W; c.H = H; c.frameIndex = frameIndex;
    for (int r = 0; r < 3; r++) for (int k = 0; k < 3; k++) c.viewToWorld[r * 3 + k] = worldToView[k * 4 + r];        // inverse of the rotation = transpose: row r = view axis r in world space
    memcpy(c.viewToClip, viewToClip, 64);
    const float P00 = viewToClip[0], P11 = viewToClip[5], P20 = viewToClip[8], P21 = viewToClip[9];
    c.frustum[0] = (-1.0f - P20) / P00; c.frustum[2] = 2.0f / P00; c.frustum[1] = (1.0f - P21) / P11; c.frustum[3] = -2.0f / P11;      // Xv.xy = ( uv * frustum.zw + frustum.xy ) * viewZ
    c.unproject = 1.0f / (0.5f * float(H) * P11);
    c.minRectDimMulUnproject = float(std::min(W, H)) * c.unproject;
    const float rad90 = 1.5707963267948966f, rad360 = 6.283185307179586f;
    c.rotatorPre = GetRotator(Weyl1D(0.5f, frameIndex) * rad90);
    c.rotator = CombineRotators(GetRotator(Weyl1D(0.0f, frameIndex * 2) * rad90), GetRotator(Bayer4x4_00(frameIndex * 2) * rad360));
    c.rotatorPost = CombineRotators(GetRotator(Weyl1D(0.0f, frameIndex * 2 + 1) * rad90), GetRotator(Bayer4x4_00(frameIndex * 2 + 1) * rad360));
    c.gLobeAngleFraction = s.lobeAngleFraction * s.lobeAngleFraction;
    c.gMaxBlurRadius = std::max(s.maxBlurRadius, s.minBlurRadius); c.gDiffPrepassBlurRadius = s.diffusePrepassBlurRadius; c.gSpecPrepassBlurRadius = s.specularPrepassBlurRadius;
    return c;
}

// ---- images -------------------------------------------------------------------------------------------------------------------------------------------
struct Image4 { uint W = 0, H = 0; std::vector<float4> v; void init(uint w, uint h) { W = w; H = h; v.assign(size_t(w) * h, f4(0, 0, 0, 0)); }
                float4& at(int x, int y) { return v[size_t(y) * W + x]; } const float4& at(int x, int y) const { return v[size_t(y) * W + x]; }
                void store(int x, int y, float4 c) { at(x, y) = f4(lp(c.x), lp(c.y), lp(c.z), lp(c.w)); } };       // REBLUR_FORMAT = RGBA16_SFLOAT
struct Inputs           // NRD's inputs for one frame (what rtxpt_b200_denoiser_prepare_inputs writes)
{
    uint W, H; const float* viewZ; const uint32_t* normalRoughness;        // R32F, R10G10B10A2_UNORM
    float4 unpackNormalRoughness(int x, int y, float& materialID) const
    {   // NRD_FrontEnd_UnpackNormalAndRoughness, NRD_NORMAL_ENCODING_R10G10B10A2_UNORM / NRD_ROUGHNESS_ENCODING_LINEAR
        const uint32_t p = normalRoughness[size_t(y) * W + x];
        const float px = float(p & 1023u) / 1023.0f * 2.0f - 1.0f, py = float((p >> 10) & 1023u) / 1023.0f * 2.0f - 1.0f;
        float3 n = f3(px, py, 1.0f - fabsf(px) - fabsf(py));
        const float t = saturate(-n.z);
        n.x -= t * ((n.x >= 0.0f ? 1.0f : 0.0f) * 2.0f - 1.0f); n.y -= t * ((n.y >= 0.0f ? 1.0f : 0.0f) * 2.0f - 1.0f);
        n = n * (1.0f / sqrtf(dot(n, n) + 1e-9f));                   // _NRD_SafeNormalize
        materialID = float(p >> 30) / 3.0f * 3.0f;
        return f4(n, float((p >> 20) & 1023u) / 1023.0f);
    }
    float unpackViewZ(int x, int y, const Constants& c) const { return fabsf(viewZ[size_t(y) * W + x] * c.s.viewZScale); }
};

// ---- shared helpers (Common.hlsli, REBLUR_Common.hlsli) ---------------------------------------------------------------------------------------------------
inline float GetHitDistanceNormalization(float viewZ, const float* hp, float roughness) { return (hp[0] + fabsf(viewZ) * hp[1]) * lerp(1.0f, hp[2], saturate(exp2f(hp[3] * roughness * roughness))); }
inline float3 ReconstructViewPosition(float2 uv, const Constants& c, float viewZ) { return f3((uv.x * c.frustum[2] + c.frustum[0]) * viewZ, (uv.y * c.frustum[3] + c.frustum[1]) * viewZ, viewZ); }
inline float3 worldToViewRotate(const Constants& c, float3 n) { return f3(dot(f3(c.viewToWorld[0], c.viewToWorld[1], c.viewToWorld[2]), n), dot(f3(c.viewToWorld[3], c.viewToWorld[4], c.viewToWorld[5]), n), dot(f3(c.viewToWorld[6], c.viewToWorld[7], c.viewToWorld[8]), n)); }   // Geometry::RotateVectorInverse( gViewToWorld, N )
inline float GetFrustumSize(const Constants& c, float viewZ) { return c.minRectDimMulUnproject * viewZ; }
inline float PixelRadiusToWorld(const Constants& c, float pixelRadius, float viewZ) { return pixelRadius * c.unproject * viewZ; }
inline float GetSpecMagicCurve(float roughness, float power = 0.25f) { float f = 1.0f - exp2f(-200.0f * roughness * roughness); return f * Pow01(roughness, power); }
inline float ExpApprox(float x) { return 1.0f / (x * x - x + 1.0f); }
inline float ComputeExponentialWeight(float x, float px, float py) { return ExpApprox(-3.0f * fabsf(x * px + py)); }
inline float ComputeWeight(float x, float px, float py) { return SmoothStep(1.0f, 0.0f, fabsf(x * px + py)); }          // ComputeNonExponentialWeight
inline float GetGaussianWeight(float r) { return expf(-0.66f * r * r); }
static const float kNormalEncodingError = 0.75f / 255.0f;           // NRD_NORMAL_ENCODING_ERROR for R10G10B10A2
inline float GetNormalWeightParam(float nonLinearAccumSpeed, float lobeAngleFraction, float roughness = 1.0f)
{
    const float percentOfVolume = 0.75f * lerp(lobeAngleFraction, 1.0f, nonLinearAccumSpeed);
    const float angle = std::max(atanf(GetSpecularLobeTanHalfAngle(roughness, percentOfVolume)), kNormalEncodingError);
    return 1.0f / angle;
}
inline float2 GetGeometryWeightParams(float planeDistSensitivity, float frustumSize, float3 Xv, float3 Nv) { const float a = 1.0f / (planeDistSensitivity * frustumSize); return f2(a, -dot(Nv, Xv) * a); }
inline float2 GetHitDistanceWeightParams(float hitDist, float nonLinearAccumSpeed, float roughness = 1.0f)
{
    const float norm = lerp(0.0005f, 1.0f, std::min(nonLinearAccumSpeed, GetSpecMagicCurve(roughness)));
    const float a = 1.0f / norm; return f2(a, -hitDist * a);
}
inline float2 GetRoughnessWeightParams(float roughness, float fraction, float sensitivity = 0.01f) { const float a = 1.0f / lerp(sensitivity, 1.0f, saturate(roughness * fraction)); return f2(a, -roughness * a); }
inline float2 GetRelaxedRoughnessWeightParams(float m, float fraction = 1.0f, float sensitivity = 0.01f) { const float a = 1.0f / lerp(sensitivity, 1.0f, lerp(m * m, m, fraction)); return f2(a, -m * a); }
inline float GetFadeBasedOnAccumulatedFrames(const Constants& c, float accumSpeed)
{
    const float n = float(c.s.historyFixFrameNum);
    return LinearStep(n * 2.0f / 3.0f + 1e-6f, n * 4.0f / 3.0f + 2e-6f, accumSpeed);
}
inline bool CompareMaterials(float m0, float m, float minm) { return std::max(m0, minm) == std::max(m, minm); }
inline void GetKernelBasis(float3 D, float3 N, float3& T, float3& B)
{
    GetBasis(N, T, B);
    if (fabsf(dot(D, N)) < 0.999f) { const float3 R = N * (2.0f * dot(N, D)) - D; T = normalize(cross(N, R)); B = cross(R, T); }
}
inline float2 GetKernelSampleCoordinates(const Constants& c, float3 offset, float3 X, float3 T, float3 B, Rotator rotator)
{
    const float2 o = RotateVector(rotator, f2(offset.x, offset.y));
    const float3 p = X + T * o.x + B * o.y;
    const float* M = c.viewToClip;
    const float cx = p.x * M[0] + p.y * M[4] + p.z * M[8] + M[12], cy = p.x * M[1] + p.y * M[5] + p.z * M[9] + M[13], cw = p.x * M[3] + p.y * M[7] + p.z * M[11] + M[15];
    return f2(cx / cw * 0.5f + 0.5f, -(cy / cw) * 0.5f + 0.5f);
}
static const float kSpecial8[8][3] = { { -1, 0, 1 }, { 0, 1, 1 }, { 1, 0, 1 }, { 0, -1, 1 }, { -0.35355339f, 0.35355339f, 0.5f }, { 0.35355339f, 0.35355339f, 0.5f }, { 0.35355339f, -0.35355339f, 0.5f }, { -0.35355339f, -0.35355339f, 0.5f } };

// ---- ClassifyTiles: a 16x16 tile is "sky" when every pixel lies beyond the denoising range ------------------------------------------------------------------------
inline std::vector<uint8_t> classifyTiles(const Constants& c, const Inputs& in)
{
    const uint tw = (c.W + 15) / 16, th = (c.H + 15) / 16;
    std::vector<uint8_t> tiles(size_t(tw) * th, 0);
    for (uint ty = 0; ty < th; ty++) for (uint tx = 0; tx < tw; tx++)
    {
        int sum = 0;
        for (uint j = 0; j < 16; j++) for (uint i = 0; i < 16; i++)
        {   // out-of-bounds texels read 0 (viewZ 0 is inside the range), so partial border tiles are never sky
            const uint x = tx * 16 + i, y = ty * 16 + j;
            const float z = (x < c.W && y < c.H) ? in.unpackViewZ(int(x), int(y), c) : 0.0f;
            sum += z > c.s.denoisingRange ? 1 : 0;
        }
        tiles[size_t(ty) * tw + tx] = sum == 256 ? 1 : 0;
    }
    return tiles;
}
inline bool tileIsSky(const Constants& c, const std::vector<uint8_t>& tiles, int x, int y) { return tiles[size_t(y >> 4) * ((c.W + 15) / 16) + (x >> 4)] != 0; }

// ---- HitDistReconstruction 5x5: pixels whose ray missed (hit distance 0) borrow a hit distance from the same surface nearby --------------------------------------------
inline void hitDistReconstruction(const Constants& c, const Inputs& in, const std::vector<uint8_t>& tiles, const Image4& inDiff, const Image4& inSpec, Image4& outDiff, Image4& outSpec)
{
    const int BORDER = 2;
    for (int y = 0; y < int(c.H); y++) for (int x = 0; x < int(c.W); x++)
    {
        if (tileIsSky(c, tiles, x, y)) continue;
        const float viewZ = in.unpackViewZ(x, y, c);
        if (viewZ > c.s.denoisingRange) continue;
        float mid; const float4 nr = in.unpackNormalRoughness(x, y, mid);
        const float3 N = xyz(nr); const float roughness = nr.w;
        const float2 pixelUv = f2((float(x) + 0.5f) / float(c.W), (float(y) + 0.5f) / float(c.H));
        const float3 Xv = ReconstructViewPosition(pixelUv, c, viewZ), Nv = worldToViewRotate(c, N);
        const float frustumSize = GetFrustumSize(c, viewZ);
        const float2 gw = GetGeometryWeightParams(c.s.planeDistanceSensitivity, frustumSize, Xv, Nv), rw = GetRelaxedRoughnessWeightParams(roughness * roughness);
        const float diffNormalW = GetNormalWeightParam(1.0f, 1.0f), specNormalW = GetNormalWeightParam(1.0f, 1.0f, roughness);
        float2 center = f2(inDiff.at(x, y).w, inSpec.at(x, y).w);
        float2 sum = f2(center.x != 0.0f ? 1000.0f : 0.0f, center.y != 0.0f ? 1000.0f : 0.0f);
        center = center * sum;
        for (int j = -BORDER; j <= BORDER; j++) for (int i = -BORDER; i <= BORDER; i++)
        {
            if (i == 0 && j == 0) continue;
            const int sx = std::min(std::max(x + i, 0), int(c.W) - 1), sy = std::min(std::max(y + j, 0), int(c.H) - 1);        // Preload clamps to the rect
            const float2 uv = f2(pixelUv.x + float(i) / float(c.W), pixelUv.y + float(j) / float(c.H));
            float w = (uv.x > 0.0f && uv.y > 0.0f && uv.x < 1.0f && uv.y < 1.0f) ? 1.0f : 0.0f;
            w *= GetGaussianWeight(sqrtf(float(i * i + j * j)) * 0.5f);
            const float zs = in.unpackViewZ(sx, sy, c);
            w *= ComputeWeight(dot(Nv, ReconstructViewPosition(uv, c, zs)), gw.x, gw.y);
            float ms; const float4 ns = in.unpackNormalRoughness(sx, sy, ms);
            const float angle = AcosApprox(dot(N, xyz(ns)));
            float2 ww = f2(w * ComputeExponentialWeight(angle, diffNormalW, 0.0f), w * ComputeExponentialWeight(angle, specNormalW, 0.0f) * ComputeExponentialWeight(ns.w * ns.w, rw.x, rw.y));
            float2 data = f2(inDiff.at(sx, sy).w, inSpec.at(sx, sy).w);
            if (ww.x == 0.0f) data.x = 0.0f;        // Denanify
            if (ww.y == 0.0f) data.y = 0.0f;
            ww = f2(data.x != 0.0f ? ww.x : 0.0f, data.y != 0.0f ? ww.y : 0.0f);
            center = center + data * ww; sum = sum + ww;
        }
        center = f2(center.x / std::max(sum.x, NRD_EPS), center.y / std::max(sum.y, NRD_EPS));
        const float4 d = inDiff.at(x, y), s = inSpec.at(x, y);
        outDiff.store(x, y, f4(d.x, d.y, d.z, center.x)); outSpec.store(x, y, f4(s.x, s.y, s.z, center.y));
    }
}

// ---- the spatial filter shared by PrePass (mode 0), Blur (1) and PostBlur (2) ---------------------------------------------------------------------------------------
enum SpatialMode { PRE_BLUR = 0, BLUR = 1, POST_BLUR = 2 };
struct SpatialOutputs { Image4* diff; Image4* spec; std::vector<float>* specHitDistForTracking; };      // the last one: PrePass only (R16F)

inline void spatialPass(const Constants& c, const Inputs& in, const std::vector<uint8_t>& tiles, SpatialMode mode, const Image4& inDiff, const Image4& inSpec,
                        const std::vector<float2>* data1 /* accumulated frames (diff, spec) per pixel; Blur / PostBlur */, SpatialOutputs out)
{
    const float fractionScale = mode == PRE_BLUR ? 2.0f : (mode == BLUR ? 1.0f : 0.5f), radiusScale = mode == POST_BLUR ? 2.0f : 1.0f;
    const Rotator baseRotator = mode == PRE_BLUR ? c.rotatorPre : (mode == BLUR ? c.rotator : c.rotatorPost);       // rotator mode NRD_FRAME: the per-frame rotator as is
    const float2 rectSizeInv = f2(1.0f / float(c.W), 1.0f / float(c.H));
    for (int y = 0; y < int(c.H); y++) for (int x = 0; x < int(c.W); x++)
    {
        if (tileIsSky(c, tiles, x, y)) continue;
        const float viewZ = in.unpackViewZ(x, y, c);
        if (viewZ > c.s.denoisingRange) continue;
        float materialID; const float4 nr = in.unpackNormalRoughness(x, y, materialID);
        const float3 N = xyz(nr), Nv = worldToViewRotate(c, N); const float roughness = nr.w;
        const float2 pixelUv = f2((float(x) + 0.5f) * rectSizeInv.x, (float(y) + 0.5f) * rectSizeInv.y);
        const float3 Xv = ReconstructViewPosition(pixelUv, c, viewZ), Vv = normalize(-Xv);
        const float NoV = fabsf(dot(Nv, Vv));
        const float frustumSize = GetFrustumSize(c, viewZ);
        const float2 d1 = data1 ? (*data1)[size_t(y) * c.W + x] : f2(0, 0);
        auto sampleCoords = [&](float2 uv, int& sx, int& sy, float2& uvSnapped) {
            uvSnapped = f2((floorf(uv.x * float(c.W)) + 0.5f) * rectSizeInv.x, (floorf(uv.y * float(c.H)) + 0.5f) * rectSizeInv.y);          // snap to the pixel centre
            const float2 cl = f2(std::min(uvSnapped.x, 1.0f - 0.5f * rectSizeInv.x), std::min(uvSnapped.y, 1.0f - 0.5f * rectSizeInv.y));      // ClampUvToViewport; gNearestClamp
            sx = std::min(std::max(int(floorf(cl.x * float(c.W))), 0), int(c.W) - 1); sy = std::min(std::max(int(floorf(cl.y * float(c.H))), 0), int(c.H) - 1);
        };
        // ---- diffuse (REBLUR_Common_DiffuseSpatialFilter.hlsli; screen-space sampling) ----
        {
            float sum = 1.0f; float4 diff = inDiff.at(x, y);
            if (mode != PRE_BLUR || c.gDiffPrepassBlurRadius != 0.0f)
            {
                const float hitDist = diff.w * GetHitDistanceNormalization(viewZ, c.s.hitDistParams, 1.0f);
                const float hitDistFactor = saturate(hitDist / frustumSize);
                float nonLinearAccumSpeed, blurRadius, areaFactor;
                if (mode == PRE_BLUR) { nonLinearAccumSpeed = 1.0f / 11.0f; blurRadius = c.gDiffPrepassBlurRadius; areaFactor = hitDistFactor; }
                else
                {
                    float boost = 1.0f - GetFadeBasedOnAccumulatedFrames(c, d1.x); boost *= 1.0f - Pow5(NoV);
                    nonLinearAccumSpeed = 1.0f / (1.0f + (1.0f - boost) * d1.x);
                    blurRadius = c.gMaxBlurRadius; areaFactor = hitDistFactor * nonLinearAccumSpeed;
                }
                blurRadius *= Sqrt01(areaFactor); blurRadius *= radiusScale; blurRadius = std::max(blurRadius, c.s.minBlurRadius);
                const float2 gw = GetGeometryWeightParams(c.s.planeDistanceSensitivity, frustumSize, Xv, Nv);
                const float normalW = GetNormalWeightParam(nonLinearAccumSpeed, c.gLobeAngleFraction) / fractionScale;
                const float2 hw = GetHitDistanceWeightParams(diff.w, nonLinearAccumSpeed);
                float minHitDistWeight = c.s.minHitDistanceWeight * fractionScale;
                if (mode != PRE_BLUR) minHitDistWeight *= sqrtf(nonLinearAccumSpeed);
                float2 skew = f2(1, 1);
                if (mode != PRE_BLUR) { skew = f2(lerp(1.0f - fabsf(Nv.x), 1.0f, NoV), lerp(1.0f - fabsf(Nv.y), 1.0f, NoV)); const float m = std::max(skew.x, skew.y); skew = f2(skew.x / m, skew.y / m); }
                skew = f2(skew.x * rectSizeInv.x * blurRadius, skew.y * rectSizeInv.y * blurRadius);
                const Rotator scaledRotator = ScaleRotator(baseRotator, skew);
                for (int n = 0; n < 8; n++)
                {
                    const float2 o = RotateVector(scaledRotator, f2(kSpecial8[n][0], kSpecial8[n][1]));
                    int sx, sy; float2 uv; sampleCoords(f2(pixelUv.x + o.x, pixelUv.y + o.y), sx, sy, uv);
                    const float zs = in.unpackViewZ(sx, sy, c);
                    float ms; const float4 Ns = in.unpackNormalRoughness(sx, sy, ms);
                    float w = (uv.x > 0.0f && uv.y > 0.0f && uv.x < 1.0f && uv.y < 1.0f) ? 1.0f : 0.0f;
                    w *= ComputeWeight(dot(Nv, ReconstructViewPosition(uv, c, zs)), gw.x, gw.y);
                    w *= CompareMaterials(materialID, ms, c.s.minMaterialForDiffuse) ? 1.0f : 0.0f;
                    w *= ComputeWeight(AcosApprox(dot(N, xyz(Ns))), normalW, 0.0f);
                    float4 s = inDiff.at(sx, sy); if (w == 0.0f) s = f4(0, 0, 0, 0);
                    w *= lerp(minHitDistWeight, 1.0f, ComputeExponentialWeight(s.w, hw.x, hw.y));
                    w *= GetGaussianWeight(kSpecial8[n][2]);
                    sum += w; diff = diff + s * w;
                }
                diff = diff * PositiveRcp(sum);
            }
            out.diff->store(x, y, diff);
        }
        // ---- specular (REBLUR_Common_SpecularSpatialFilter.hlsli; world-space kernel bent towards the dominant direction, screen space in the pre-pass) ----
        {
            float sum = 1.0f; float4 spec = inSpec.at(x, y);
            const float smc = GetSpecMagicCurve(roughness);
            if (mode != PRE_BLUR || c.gSpecPrepassBlurRadius != 0.0f)
            {
                HashRng rng; rng.Initialize(uint(x), uint(y), c.frameIndex);
                const float4 Dv = GetSpecularDominantDirection(Nv, Vv, roughness);
                const float NoD = fabsf(dot(Nv, xyz(Dv)));
                const float hitDist = spec.w * GetHitDistanceNormalization(viewZ, c.s.hitDistParams, roughness);
                const float hitDistFactor = saturate(hitDist / frustumSize);
                float hitDistForTracking = hitDist == 0.0f ? NRD_INF : hitDist;
                float nonLinearAccumSpeed, blurRadius, areaFactor;
                if (mode == PRE_BLUR) { nonLinearAccumSpeed = 1.0f / 11.0f; blurRadius = c.gSpecPrepassBlurRadius; areaFactor = roughness * hitDistFactor; }
                else
                {
                    float boost = 1.0f - GetFadeBasedOnAccumulatedFrames(c, d1.y); boost *= 1.0f - Pow5(NoV); boost *= smc;
                    nonLinearAccumSpeed = 1.0f / (1.0f + (1.0f - boost) * d1.y);
                    blurRadius = c.gMaxBlurRadius; areaFactor = roughness * hitDistFactor * nonLinearAccumSpeed;
                }
                blurRadius *= Sqrt01(areaFactor);
                if (mode == PRE_BLUR)
                {
                    const float lobeRadius = hitDist * NoD * GetSpecularLobeTanHalfAngle(roughness, 0.3f);
                    blurRadius = std::min(blurRadius, lobeRadius / PixelRadiusToWorld(c, 1.0f, viewZ + hitDist * Dv.w));
                }
                blurRadius *= radiusScale; blurRadius = std::max(blurRadius, c.s.minBlurRadius * smc);
                const float roughnessFractionScaled = saturate(c.s.roughnessFraction * fractionScale);
                const float2 gw = GetGeometryWeightParams(c.s.planeDistanceSensitivity, frustumSize, Xv, Nv);
                const float normalW = GetNormalWeightParam(nonLinearAccumSpeed, c.gLobeAngleFraction, roughness) / fractionScale;
                const float2 rw = GetRoughnessWeightParams(roughness, roughnessFractionScaled);
                const float2 hw = GetHitDistanceWeightParams(spec.w, nonLinearAccumSpeed, roughness);
                float minHitDistWeight = c.s.minHitDistanceWeight * fractionScale * smc;
                if (mode != PRE_BLUR) minHitDistWeight *= sqrtf(nonLinearAccumSpeed);
                Rotator scaledRotator = baseRotator; float3 Tv = f3(0), Bv = f3(0);
                if (mode == PRE_BLUR) scaledRotator = ScaleRotator(baseRotator, f2(rectSizeInv.x * blurRadius, rectSizeInv.y * blurRadius));
                else
                {
                    const float bentFactor = sqrtf(hitDistFactor);
                    float skewFactor = lerp(0.25f + 0.75f * roughness, 1.0f, NoD); skewFactor = lerp(skewFactor, 1.0f, nonLinearAccumSpeed); skewFactor = lerp(1.0f, skewFactor, bentFactor);
                    GetKernelBasis(normalize(lerp(Nv, xyz(Dv), bentFactor)), Nv, Tv, Bv);
                    const float worldRadius = PixelRadiusToWorld(c, blurRadius, viewZ);
                    Tv = Tv * (worldRadius * skewFactor); Bv = Bv * (worldRadius / skewFactor);
                }
                for (int n = 0; n < 8; n++)
                {
                    float2 uvRaw;
                    if (mode == PRE_BLUR) { const float2 o = RotateVector(scaledRotator, f2(kSpecial8[n][0], kSpecial8[n][1])); uvRaw = f2(pixelUv.x + o.x, pixelUv.y + o.y); }
                    else uvRaw = GetKernelSampleCoordinates(c, f3(kSpecial8[n][0], kSpecial8[n][1], kSpecial8[n][2]), Xv, Tv, Bv, baseRotator);
                    int sx, sy; float2 uv; sampleCoords(uvRaw, sx, sy, uv);
                    const float zs = in.unpackViewZ(sx, sy, c);
                    float ms; const float4 Ns = in.unpackNormalRoughness(sx, sy, ms);
                    const float3 Xvs = ReconstructViewPosition(uv, c, zs);
                    float w = (uv.x > 0.0f && uv.y > 0.0f && uv.x < 1.0f && uv.y < 1.0f) ? 1.0f : 0.0f;
                    w *= ComputeWeight(dot(Nv, Xvs), gw.x, gw.y);
                    w *= CompareMaterials(materialID, ms, c.s.minMaterialForSpecular) ? 1.0f : 0.0f;
                    w *= ComputeWeight(AcosApprox(dot(N, xyz(Ns))), normalW, 0.0f);
                    w *= ComputeWeight(Ns.w, rw.x, rw.y);
                    float4 s = inSpec.at(sx, sy); if (w == 0.0f) s = f4(0, 0, 0, 0);
                    if (mode == PRE_BLUR)
                    {
                        const float hs = s.w * GetHitDistanceNormalization(zs, c.s.hitDistParams, Ns.w);
                        const float geometryWeight = w * NoV * (hs != 0.0f ? 1.0f : 0.0f);
                        if (rng.GetFloat() < geometryWeight) hitDistForTracking = std::min(hitDistForTracking, hs);
                        w *= c.s.usePrepassOnlyForSpecularMotionEstimation ? 0.0f : 1.0f;
                        const float d = length(Xvs - Xv) + NRD_EPS, t = hs / (d + hitDist);
                        w *= lerp(saturate(t), 1.0f, LinearStep(0.5f, 1.0f, roughness));
                    }
                    w *= lerp(minHitDistWeight, 1.0f, ComputeExponentialWeight(s.w, hw.x, hw.y));
                    w *= GetGaussianWeight(kSpecial8[n][2]);
                    sum += w; spec = spec + s * w;
                }
                spec = spec * PositiveRcp(sum);
                if (mode == PRE_BLUR && out.specHitDistForTracking) (*out.specHitDistForTracking)[size_t(y) * c.W + x] = lp(hitDistForTracking == NRD_INF ? 0.0f : hitDistForTracking);
            }
            out.spec->store(x, y, spec);
        }
    }
}

} } // namespace orc::reblur
