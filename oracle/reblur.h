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
    #pragma omp parallel for schedule(dynamic, 4)          // pixels are independent within a pass (each writes its own outputs only)
    for (int y = 0; y < int(c.H); y++) for (int x = 0; x < int(c.W); x++)
    {
        if (tileIsSky(c, tiles, x, y)) continue;
        const float viewZ = in.unpackViewZ(x, y, c);
        if (viewZ > c.s.denoisingRange) continue;
        float mid; const float4 nr = in.unpackNormalRoughness(x, y, mid);
        const float3 N = xyz(nr); const float roughness = nr.w;
        const float2 rectSizeInv = f2(1.0f / float(c.W), 1.0f / float(c.H));
        const float2 pixelUv = f2((float(x) + 0.5f) * rectSizeInv.x, (float(y) + 0.5f) * rectSizeInv.y);          // ( pixelPos + 0.5 ) * gRectSizeInv
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
            const float2 uv = f2(pixelUv.x + float(i) * rectSizeInv.x, pixelUv.y + float(j) * rectSizeInv.y);
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
    #pragma omp parallel for schedule(dynamic, 4)          // pixels are independent within a pass (each writes its own outputs only)
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


// =====================================================================================================================================================
// TemporalAccumulation (External/Nrd/Shaders/Include/REBLUR_TemporalAccumulation.hlsli:11-937): reprojection of the diffuse / specular history by surface motion
// ("smb") and, for specular, by the motion of the virtual image of the reflection ("vmb"); disocclusion tests against the previous frame's depth, normals,
// roughness and material; history length bookkeeping; firefly suppression; fast history.  Checkerboarding, history-confidence inputs, SH and occlusion-only
// variants, orthographic cameras and object motion (gWorldPrevToWorld = identity) are outside RTXPT's configuration and not restated.
// =====================================================================================================================================================
struct FrameMatrices        // camera-relative like NRD makes them (InstanceImpl.cpp:404-415): the current camera sits at the origin of "world" space
{
    float viewToWorld[9];           // rows = view axes in world space
    float worldToClip[16], worldToClipPrev[16];
    float worldToViewPrev[12];      // rows 0-2: rotation (view axis k of the previous camera), column 3: translation, so Xv = R * X + t
    float frustumPrev[4];
    float3 cameraDelta;             // previous camera position - current camera position
};
struct History              // the permanent pool of one NRD instance (Reblur_DiffuseSpecular.hpp:21-49); RTXPT keeps one instance per stable plane
{
    uint W = 0, H = 0; bool valid = false;
    std::vector<float> prevViewZ; std::vector<uint32_t> prevNormalRoughness; std::vector<uint16_t> prevInternalData;
    Image4 diff, spec; std::vector<float> diffFast, specFast, specHitDistForTracking, diffLumaStabilized, specLumaStabilized;
    void init(uint w, uint h)
    {
        W = w; H = h; valid = false; const size_t n = size_t(w) * h;
        prevViewZ.assign(n, 0.0f); prevNormalRoughness.assign(n, 0u); prevInternalData.assign(n, 0); diff.init(w, h); spec.init(w, h); diffFast.assign(n, 0.0f); specFast.assign(n, 0.0f); specHitDistForTracking.assign(n, 0.0f); diffLumaStabilized.assign(n, 0.0f); specLumaStabilized.assign(n, 0.0f);
    }
};
struct TemporalParams
{
    float disocclusionThreshold = 0.03f, disocclusionThresholdAlternate = 0.2f;      // m_ui.NRDDisocclusionThreshold / ...Alternate (SampleUI.h:294-296) + the jitter bonus (Reblur.cpp:293)
    float framerateScale = 2.0f;                                                      // max( 33.333 ms / frame time, 1 ): 1/60 s frames
    float fireflySuppressorMinRelativeScale = 2.0f, responsiveAccumulationRoughnessThreshold = 0.0f;
    bool resetHistory = false;
    const uint8_t* disocclusionThresholdMix = nullptr;                               // R8_UNORM, IN_DISOCCLUSION_THRESHOLD_MIX
    const uint16_t* motion = nullptr;                                                // IN_MV RGBA16F: xy pixels (gMvScale = 1 / size), z view-depth delta
};
static const float kAlmostZeroAngle = 0.01745240643728351f;                            // cos( 89 degrees )
static const float kRoughnessSensitivityInTA = 0.01f * 0.3f;

inline uint16_t PackInternalData(float diffAccumSpeed, float specAccumSpeed, float materialID)
{   // Packing::RgbaToUint( t.xyzz, 6, 6, 4, 0 )
    const uint a = uint(saturate(diffAccumSpeed / 63.0f) * 63.0f + 0.5f), b = uint(saturate(specAccumSpeed / 63.0f) * 63.0f + 0.5f), c = uint(saturate(materialID / 15.0f) * 15.0f + 0.5f);
    return uint16_t(a | (b << 6) | (c << 12));
}
inline float3 UnpackInternalData(uint p) { return f3(float(p & 63u), float((p >> 6) & 63u), float((p >> 12) & 15u)); }       // already scaled back by 63 / 63 / 15
inline float GetModifiedRoughnessFromNormalVariance(float roughness, float3 nonNormalizedAverageNormal)
{
    const float l = length(nonNormalizedAverageNormal);
    const float kappa = saturate(1.0f - l * l) * PositiveRcp(l * (3.0f - l * l));
    return Sqrt01(roughness * roughness + kappa);
}
inline float GetSpecularDominantFactor(float NoV, float roughness)
{
    const float a = 0.298475f * logf(39.4115f - 39.0029f * roughness);
    return saturate(powf(saturate(1.0f - NoV), 10.8649f) * (1.0f - a) + a);
}
inline float2 GetScreenUv(const float* M, float3 X)
{
    const float cx = X.x * M[0] + X.y * M[4] + X.z * M[8] + M[12], cy = X.x * M[1] + X.y * M[5] + X.z * M[9] + M[13], cw = X.x * M[3] + X.y * M[7] + X.z * M[11] + M[15];
    if (cw < 0.0f) return f2(99999.0f, 99999.0f);
    return f2(cx / cw * 0.5f + 0.5f, -(cy / cw) * 0.5f + 0.5f);
}
inline float3 viewToWorldRotate(const float* r, float3 v) { return f3(r[0], r[1], r[2]) * v.x + f3(r[3], r[4], r[5]) * v.y + f3(r[6], r[7], r[8]) * v.z; }
inline float ComputeNonExponentialWeightWithSigma(float x, float px, float py, float sigma) { return SmoothStep(1.0f, 0.0f, fabsf(x * px + py) - sigma * px); }
inline float GetEncodingAwareNormalWeight(float3 Ncurr, float3 Nprev, float maxAngle, float curvatureAngle, float thresholdAngle)
{
    const float angle = AcosApprox(dot(Ncurr, Nprev));
    return SmoothStep01(1.0f - (angle - curvatureAngle - thresholdAngle) / maxAngle);
}
inline float3 GetXvirtual(float hitDist, float curvature, float3 X, float3 Xprev, float3 N, float3 V, float roughness)
{   // Common.hlsli:405-453, NRD_USE_SPECULAR_MOTION_V2
    const float4 D = GetSpecularDominantDirection(N, V, roughness);
    const float3 reflectionRay = xyz(D) * hitDist;
    float3 T, B; GetBasis(N, T, B);
    float3 O = f3(dot(T, reflectionRay), dot(B, reflectionRay), dot(N, reflectionRay));      // Geometry::RotateVector( basis, v )
    O.z = -O.z;
    float mag = 1.0f / (2.0f * curvature * O.z - 1.0f);
    float f = length(X); f *= 1.0f - fabsf(dot(N, V)); f *= std::max(curvature, 0.0f);
    mag *= 1.0f / (1.0f + f);
    const float3 Iw = V * length(O * mag);
    const float closenessToSurface = saturate(length(Iw) / (hitDist + NRD_EPS));
    const float3 origin = lerp(Xprev, X, closenessToSurface * D.w);
    return origin - Iw * D.w;
}
struct Bilinear { float2 origin, weights; };
inline Bilinear GetBilinearFilter(float2 uv, float2 texSize) { const float2 t = f2(uv.x * texSize.x - 0.5f, uv.y * texSize.y - 0.5f); Bilinear b; b.origin = f2(floorf(t.x), floorf(t.y)); b.weights = f2(t.x - b.origin.x, t.y - b.origin.y); return b; }
inline float ApplyBilinearFilter(float s00, float s10, float s01, float s11, const Bilinear& f) { return lerp(lerp(s00, s10, f.weights.x), lerp(s01, s11, f.weights.x), f.weights.y); }
inline float4 GetBilinearCustomWeights(const Bilinear& f, float4 cw)
{
    const float ox = 1.0f - f.weights.x, oy = 1.0f - f.weights.y;
    return f4(cw.x * ox * oy, cw.y * f.weights.x * oy, cw.z * ox * f.weights.y, cw.w * f.weights.x * f.weights.y);
}
inline float ApplyBilinearCustomWeights(float s00, float s10, float s01, float s11, float4 w)
{
    const float sum = w.x + w.y + w.z + w.w;
    return sum < 0.0001f ? 0.0f : (s00 * w.x + s10 * w.y + s01 * w.z + s11 * w.w) / sum;
}
inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
// texture.SampleLevel( gLinearClamp, uv ) on an RGBA16F image / an R16F image; uv in texels here
inline float4 sampleBilinear(const Image4& img, float2 posTexels)
{
    const float tx = posTexels.x - 0.5f, ty = posTexels.y - 0.5f; const float fx = floorf(tx), fy = floorf(ty), wx = tx - fx, wy = ty - fy;
    const int x0 = clampi(int(fx), 0, int(img.W) - 1), x1 = clampi(int(fx) + 1, 0, int(img.W) - 1), y0 = clampi(int(fy), 0, int(img.H) - 1), y1 = clampi(int(fy) + 1, 0, int(img.H) - 1);
    const float4 a = img.at(x0, y0), b = img.at(x1, y0), c = img.at(x0, y1), d = img.at(x1, y1);
    return (a * (1 - wx) + b * wx) * (1 - wy) + (c * (1 - wx) + d * wx) * wy;
}
inline float sampleBilinear(const std::vector<float>& img, uint W, uint H, float2 posTexels)
{
    const float tx = posTexels.x - 0.5f, ty = posTexels.y - 0.5f; const float fx = floorf(tx), fy = floorf(ty), wx = tx - fx, wy = ty - fy;
    const int x0 = clampi(int(fx), 0, int(W) - 1), x1 = clampi(int(fx) + 1, 0, int(W) - 1), y0 = clampi(int(fy), 0, int(H) - 1), y1 = clampi(int(fy) + 1, 0, int(H) - 1);
    return lerp(lerp(img[size_t(y0) * W + x0], img[size_t(y0) * W + x1], wx), lerp(img[size_t(y1) * W + x0], img[size_t(y1) * W + x1], wx), wy);
}
// BicubicFilterNoCornersWithFallbackToBilinearFilterWithCustomWeights (Common.hlsli:610-665): 12-tap Catmull-Rom through 5 bilinear fetches, or the 2x2 footprint with custom weights
inline void sampleHistory(float2 samplePos, float4 bilinearCustomWeights, bool useBicubic, const Image4& tex0, float4& c0, const std::vector<float>& tex1, float& c1)
{
    const float S = 0.5f;       // NRD_CATROM_SHARPNESS
    const float2 centerPos = f2(floorf(samplePos.x - 0.5f) + 0.5f, floorf(samplePos.y - 0.5f) + 0.5f);
    const float2 f = f2(saturate(samplePos.x - centerPos.x), saturate(samplePos.y - centerPos.y));
    auto w0f = [&](float t) { return t * (t * (-S * t + 2.0f * S) - S); }; auto w1f = [&](float t) { return t * (t * ((2.0f - S) * t - (3.0f - S))) + 1.0f; };
    auto w2f = [&](float t) { return t * (t * (-(2.0f - S) * t + (3.0f - 2.0f * S)) + S); }; auto w3f = [&](float t) { return t * (t * (S * t - S)); };
    const float2 w0 = f2(w0f(f.x), w0f(f.y)), w1 = f2(w1f(f.x), w1f(f.y)), w2 = f2(w2f(f.x), w2f(f.y)), w3 = f2(w3f(f.x), w3f(f.y));
    const float2 w12 = f2(w1.x + w2.x, w1.y + w2.y), tc = f2(w2.x / w12.x, w2.y / w12.y);
    float4 w = f4(w12.x * w0.y, w0.x * w12.y, w12.x * w12.y, w3.x * w12.y); float w4 = w12.x * w3.y;
    if (!useBicubic) { w = bilinearCustomWeights; w4 = 0.0f; }
    const float sum = w.x + w.y + w.z + w.w + w4;
    float4 color;
    if (useBicubic)
        color = sampleBilinear(tex0, f2(centerPos.x + tc.x, centerPos.y - 1.0f)) * w.x + sampleBilinear(tex0, f2(centerPos.x - 1.0f, centerPos.y + tc.y)) * w.y + sampleBilinear(tex0, f2(centerPos.x + tc.x, centerPos.y + tc.y)) * w.z
              + sampleBilinear(tex0, f2(centerPos.x + 2.0f, centerPos.y + tc.y)) * w.w + sampleBilinear(tex0, f2(centerPos.x + tc.x, centerPos.y + 2.0f)) * w4;
    else
        color = sampleBilinear(tex0, centerPos) * w.x + sampleBilinear(tex0, f2(centerPos.x + 1, centerPos.y)) * w.y + sampleBilinear(tex0, f2(centerPos.x, centerPos.y + 1)) * w.z + sampleBilinear(tex0, f2(centerPos.x + 1, centerPos.y + 1)) * w.w
              + sampleBilinear(tex0, f2(centerPos.x + f.x, centerPos.y + f.y)) * w4;
    c0 = sum < 0.0001f ? f4(0, 0, 0, 0) : color * (1.0f / sum);
    // fast history: always the 2x2 footprint with the custom weights
    const int ox = int(centerPos.x), oy = int(centerPos.y), W = int(tex0.W), H = int(tex0.H);
    auto ld = [&](int x, int y) { return tex1[size_t(clampi(y, 0, H - 1)) * W + clampi(x, 0, W - 1)]; };
    const float bsum = bilinearCustomWeights.x + bilinearCustomWeights.y + bilinearCustomWeights.z + bilinearCustomWeights.w;
    const float v = ld(ox, oy) * bilinearCustomWeights.x + ld(ox + 1, oy) * bilinearCustomWeights.y + ld(ox, oy + 1) * bilinearCustomWeights.z + ld(ox + 1, oy + 1) * bilinearCustomWeights.w;
    c1 = bsum < 0.0001f ? 0.0f : v / bsum;
}
inline float4 ClampNegativeToZero(float4 v)
{   // YCoCg -> linear (clamped at 0) -> YCoCg, hit distance saturated
    const float t = v.x - v.z; const float3 rgb = f3(std::max(t + v.y, 0.0f), std::max(v.x + v.z, 0.0f), std::max(t - v.y, 0.0f));
    return f4(dot(rgb, f3(0.25f, 0.5f, 0.25f)), dot(rgb, f3(0.5f, 0.0f, -0.5f)), dot(rgb, f3(-0.25f, 0.5f, -0.25f)), saturate(v.w));
}
inline float GetMinAllowedLimitForHitDistNonLinearAccumSpeed(float roughness, float maxAccumulatedFrameNum) { return 1.0f / (1.0f + 0.5f * GetSpecMagicCurve(roughness) * maxAccumulatedFrameNum); }
inline float4 MixHistoryAndCurrent(float4 history, float4 current, float f, float roughness, float maxAccumulatedFrameNum)
{
    const float fw = std::max(f, GetMinAllowedLimitForHitDistNonLinearAccumSpeed(roughness, maxAccumulatedFrameNum));
    return f4(lerp(history.x, current.x, f), lerp(history.y, current.y, f), lerp(history.z, current.z, f), lerp(history.w, current.w, fw));
}
inline float4 ChangeLuma(float4 v, float newLuma) { const float s = (newLuma + NRD_EPS) / (v.x + NRD_EPS); return f4(v.x * s, v.y * s, v.z * s, v.w); }

struct TemporalOutputs
{
    Image4* diff; Image4* spec; std::vector<float>* diffFast; std::vector<float>* specFast; std::vector<float>* specHitDistForTracking;
    std::vector<uint8_t>* data1;        // RG8_UNORM: accumulated frames / 63 (diffuse, specular)
    std::vector<uint32_t>* data2;       // occlusion bits, virtual history amount, CatRom flag, curvature (REBLUR_Common.hlsli:63-95)
};

inline void temporalAccumulation(const Constants& c, const FrameMatrices& m, const TemporalParams& tp, const Inputs& in, const std::vector<uint8_t>& tiles, const Image4& inDiff, const Image4& inSpec,
                                 const std::vector<float>& prepassHitDistForTracking, const History& h, TemporalOutputs out)
{
    const float maxAccum = tp.resetHistory || !h.valid ? 0.0f : float(std::min(c.s.maxAccumulatedFrameNum, 63u)), maxFastAccum = tp.resetHistory || !h.valid ? 0.0f : float(c.s.maxFastAccumulatedFrameNum);
    const float historyFixFrameNum = float(c.s.historyFixFrameNum);
    const float2 rectSize = f2(float(c.W), float(c.H)), rectSizeInv = f2(1.0f / float(c.W), 1.0f / float(c.H));
    const int W = int(c.W), H = int(c.H);
    auto normalAt = [&](int x, int y) { float mid; return in.unpackNormalRoughness(clampi(x, 0, W - 1), clampi(y, 0, H - 1), mid); };
    auto prevNormalAt = [&](int x, int y) { Inputs p; p.W = c.W; p.H = c.H; p.viewZ = nullptr; p.normalRoughness = h.prevNormalRoughness.data(); float mid; return p.unpackNormalRoughness(clampi(x, 0, W - 1), clampi(y, 0, H - 1), mid); };
    auto prevZ = [&](int x, int y) { return fabsf(h.prevViewZ[size_t(clampi(y, 0, H - 1)) * W + clampi(x, 0, W - 1)] * c.s.viewZScale); };
    auto prevInternal = [&](int x, int y) { return uint(h.prevInternalData[size_t(clampi(y, 0, H - 1)) * W + clampi(x, 0, W - 1)]); };
    auto affinePrev = [&](float3 X) { const float* t = m.worldToViewPrev; return f3(t[0] * X.x + t[1] * X.y + t[2] * X.z + t[3], t[4] * X.x + t[5] * X.y + t[6] * X.z + t[7], t[8] * X.x + t[9] * X.y + t[10] * X.z + t[11]); };
    auto rotatePrevInverse = [&](float3 v) { const float* t = m.worldToViewPrev; return f3(t[0] * v.x + t[4] * v.y + t[8] * v.z, t[1] * v.x + t[5] * v.y + t[9] * v.z, t[2] * v.x + t[6] * v.y + t[10] * v.z); };
    auto inScreenBilinear = [&](float2 origin) {
        const float px[4] = { origin.x, origin.y, origin.x + 1, origin.y + 1 }; float r[4];
        for (int k = 0; k < 4; k++) r[k] = (px[k] >= 0.0f && px[k] < ((k & 1) ? rectSize.y : rectSize.x)) ? 1.0f : 0.0f;
        return f4(r[0] * r[1], r[2] * r[1], r[0] * r[3], r[2] * r[3]); };       // r.xzxz * r.yyww
    #pragma omp parallel for schedule(dynamic, 4)          // pixels are independent within a pass (each writes its own outputs only)
    for (int y = 0; y < H; y++) for (int x = 0; x < W; x++)
    {
        if (tileIsSky(c, tiles, x, y)) continue;
        const float viewZ = in.unpackViewZ(x, y, c);
        if (viewZ > c.s.denoisingRange) continue;
        const size_t pix = size_t(y) * W + x;
        const float2 pixelUv = f2((float(x) + 0.5f) * rectSizeInv.x, (float(y) + 0.5f) * rectSizeInv.y);
        const float3 Xv = ReconstructViewPosition(pixelUv, c, viewZ), X = viewToWorldRotate(m.viewToWorld, Xv);
        // 3x3 neighbourhood: tracking distance, averaged normal (2x2 towards +x,+y... the first 2x2 of the 3x3 block), roughness variance
        float3 Navg = f3(0); float hitDistForTracking = NRD_INF, roughnessM1 = 0, roughnessM2 = 0;
        for (int j = 0; j <= 2; j++) for (int i = 0; i <= 2; i++)
        {
            const int sx = clampi(x + i - 1, 0, W - 1), sy = clampi(y + j - 1, 0, H - 1);
            const float4 nr = normalAt(sx, sy);
            if (i < 2 && j < 2) Navg = Navg + xyz(nr);
            float hd = c.gSpecPrepassBlurRadius == 0.0f ? inSpec.at(sx, sy).w : prepassHitDistForTracking[size_t(sy) * W + sx];
            hitDistForTracking = std::min(hitDistForTracking, hd == 0.0f ? NRD_INF : hd);
            const float r2 = nr.w * nr.w; roughnessM1 += r2; roughnessM2 += r2 * r2;
        }
        Navg = Navg / 4.0f;
        float materialID; const float4 nrC = in.unpackNormalRoughness(x, y, materialID);
        const float3 N = xyz(nrC); const float roughness = nrC.w;
        const float roughnessModified = GetModifiedRoughnessFromNormalVariance(roughness, Navg);
        roughnessM1 /= 9.0f; roughnessM2 /= 9.0f;
        const float roughnessSigma = sqrtf(fabsf(roughnessM2 - roughnessM1 * roughnessM1));
        HashRng rng; rng.Initialize(uint(x), uint(y), c.frameIndex);
        hitDistForTracking = hitDistForTracking == NRD_INF ? 0.0f : hitDistForTracking;
        const float hitDistNormalization = GetHitDistanceNormalization(viewZ, c.s.hitDistParams, roughness);
        hitDistForTracking *= c.gSpecPrepassBlurRadius == 0.0f ? hitDistNormalization : 1.0f;
        (*out.specHitDistForTracking)[pix] = lp(hitDistForTracking);
        // previous position, surface motion
        float3 mv = f3(0);
        if (tp.motion) mv = f3(f16tof32(tp.motion[pix * 4]) * rectSizeInv.x, f16tof32(tp.motion[pix * 4 + 1]) * rectSizeInv.y, f16tof32(tp.motion[pix * 4 + 2]));
        const float2 smbPixelUv = f2(pixelUv.x + mv.x, pixelUv.y + mv.y);
        const float viewZprev = viewZ + mv.z;
        const float3 Xvprevlocal = f3((smbPixelUv.x * m.frustumPrev[2] + m.frustumPrev[0]) * viewZprev, (smbPixelUv.y * m.frustumPrev[3] + m.frustumPrev[1]) * viewZprev, viewZprev);
        const float3 Xprev = rotatePrevInverse(Xvprevlocal) + m.cameraDelta;
        // previous depth in the 4x4 block without corners, 2x2 bilinear footprint in its middle
        const Bilinear smbBil = GetBilinearFilter(smbPixelUv, rectSize);
        const int bx = int(smbBil.origin.x), by = int(smbBil.origin.y), ox = bx - 1, oy = by - 1;     // Catmull-Rom origin = bilinear origin - 1
        const float pz0[3] = { prevZ(ox + 1, oy), prevZ(ox, oy + 1), prevZ(ox + 1, oy + 1) }, pz1[3] = { prevZ(ox + 2, oy), prevZ(ox + 2, oy + 1), prevZ(ox + 3, oy + 1) };
        const float pz2[3] = { prevZ(ox, oy + 2), prevZ(ox + 1, oy + 2), prevZ(ox + 1, oy + 3) }, pz3[3] = { prevZ(ox + 2, oy + 2), prevZ(ox + 3, oy + 2), prevZ(ox + 2, oy + 3) };
        float3 smbNavg = f3(0);
        {
            float sum = 0; const float wz[4] = { pz0[2] < c.s.denoisingRange ? 1.0f : 0.0f, pz1[1] < c.s.denoisingRange ? 1.0f : 0.0f, pz2[1] < c.s.denoisingRange ? 1.0f : 0.0f, pz3[0] < c.s.denoisingRange ? 1.0f : 0.0f };
            const int qx[4] = { 0, 1, 0, 1 }, qy[4] = { 0, 0, 1, 1 };
            for (int k = 0; k < 4; k++) { smbNavg = smbNavg + xyz(prevNormalAt(bx + qx[k], by + qy[k])) * wz[k]; sum += wz[k]; }
            smbNavg = smbNavg / (sum == 0.0f ? 1.0f : sum);
        }
        auto parallaxInPixels = [&](float3 Xp, float2 uvZero, const float* M) { const float2 uv = GetScreenUv(M, Xp); return length(f2((uv.x - uvZero.x) * rectSize.x, (uv.y - uvZero.y) * rectSize.y)); };
        const float smbParallax1 = parallaxInPixels(Xprev + m.cameraDelta, smbPixelUv, m.worldToClipPrev), smbParallax2 = parallaxInPixels(Xprev - m.cameraDelta, pixelUv, m.worldToClip);
        const float smbParallaxMax = std::max(smbParallax1, smbParallax2), smbParallaxMin = std::min(smbParallax1, smbParallax2);
        const float pixelSize = PixelRadiusToWorld(c, 1.0f, viewZ), frustumSize = GetFrustumSize(c, viewZ);
        const float mix = tp.disocclusionThresholdMix ? float(tp.disocclusionThresholdMix[pix]) / 255.0f : 0.0f;
        const float disocclusionThreshold = lerp(tp.disocclusionThreshold, tp.disocclusionThresholdAlternate, mix);
        const float smallParallax = LinearStep(0.25f, 0.0f, smbParallaxMax);
        const float thresholdAngle = kAlmostZeroAngle - 0.25f * smallParallax;
        const float3 V = normalize(-X);
        const float NoV = fabsf(dot(N, V));
        const float NoVstrict = lerp(NoV, 1.0f, saturate(smbParallaxMax / 30.0f));
        float4 smbThr = f4(1, 1, 1, 1) * (frustumSize * saturate(disocclusionThreshold / std::max(0.05f, NoVstrict)));
        smbThr = smbThr * (dot(smbNavg, Navg) > thresholdAngle ? 1.0f : 0.0f);
        { const float4 s = inScreenBilinear(smbBil.origin); smbThr = f4(smbThr.x * s.x - NRD_EPS, smbThr.y * s.y - NRD_EPS, smbThr.z * s.z - NRD_EPS, smbThr.w * s.w - NRD_EPS); }
        const float XvprevZ = affinePrev(Xprev).z;
        float occ0[3], occ1[3], occ2[3], occ3[3];
        for (int k = 0; k < 3; k++) { occ0[k] = fabsf(pz0[k] - XvprevZ) <= smbThr.x ? 1.0f : 0.0f; occ1[k] = fabsf(pz1[k] - XvprevZ) <= smbThr.y ? 1.0f : 0.0f; occ2[k] = fabsf(pz2[k] - XvprevZ) <= smbThr.z ? 1.0f : 0.0f; occ3[k] = fabsf(pz3[k] - XvprevZ) <= smbThr.w ? 1.0f : 0.0f; }
        {   // material ids of the previous frame (R10G10B10A2 build)
            const float minMat = std::min(c.s.minMaterialForSpecular, c.s.minMaterialForDiffuse);
            const int t0[3][2] = { { 1, 0 }, { 0, 1 }, { 1, 1 } }, t1[3][2] = { { 2, 0 }, { 2, 1 }, { 3, 1 } }, t2[3][2] = { { 0, 2 }, { 1, 2 }, { 1, 3 } }, t3[3][2] = { { 2, 2 }, { 3, 2 }, { 2, 3 } };
            for (int k = 0; k < 3; k++)
            {
                occ0[k] *= CompareMaterials(materialID, UnpackInternalData(prevInternal(ox + t0[k][0], oy + t0[k][1])).z, minMat) ? 1.0f : 0.0f; occ1[k] *= CompareMaterials(materialID, UnpackInternalData(prevInternal(ox + t1[k][0], oy + t1[k][1])).z, minMat) ? 1.0f : 0.0f;
                occ2[k] *= CompareMaterials(materialID, UnpackInternalData(prevInternal(ox + t2[k][0], oy + t2[k][1])).z, minMat) ? 1.0f : 0.0f; occ3[k] *= CompareMaterials(materialID, UnpackInternalData(prevInternal(ox + t3[k][0], oy + t3[k][1])).z, minMat) ? 1.0f : 0.0f;
            }
        }
        const float4 smbOcc = f4(occ0[2], occ1[1], occ2[1], occ3[0]);
        const float4 smbOcclusionWeights = GetBilinearCustomWeights(smbBil, smbOcc);
        float occSum = 0; for (int k = 0; k < 3; k++) occSum += occ0[k] + occ1[k] + occ2[k] + occ3[k];
        const bool smbAllowCatRom = occSum > 11.5f;
        float fbits = smbOcc.x * 1.0f + smbOcc.y * 2.0f + smbOcc.z * 4.0f + smbOcc.w * 8.0f;
        const float3 id00 = UnpackInternalData(prevInternal(bx, by)), id10 = UnpackInternalData(prevInternal(bx + 1, by)), id01 = UnpackInternalData(prevInternal(bx, by + 1)), id11 = UnpackInternalData(prevInternal(bx + 1, by + 1));
        float diffAccumSpeed = ApplyBilinearCustomWeights(id00.x, id10.x, id01.x, id11.x, smbOcclusionWeights);
        float smbSpecAccumSpeed = ApplyBilinearCustomWeights(id00.y, id10.y, id01.y, id11.y, smbOcclusionWeights);
        // footprint quality
        const float3 smbVprev = normalize(m.cameraDelta - Xprev);
        const float NoVprev = fabsf(dot(N, smbVprev));
        float sizeQuality = (NoVprev + 1e-3f) / (NoV + 1e-3f); sizeQuality *= sizeQuality; sizeQuality = lerp(0.1f, 1.0f, saturate(sizeQuality));
        float smbFootprintQuality = Sqrt01(ApplyBilinearFilter(smbOcc.x, smbOcc.y, smbOcc.z, smbOcc.w, smbBil)) * sizeQuality;
        const float2 smbSamplePos = f2(saturate(smbPixelUv.x) * rectSize.x, saturate(smbPixelUv.y) * rectSize.y);

        // ---------------------------------------------------------------- specular ----------------------------------------------------------------
        float specAccumSpeed, curvature = 0.0f, virtualHistoryAmount;
        {
            smbSpecAccumSpeed *= lerp(smbFootprintQuality, 1.0f, 1.0f / (1.0f + smbSpecAccumSpeed));
            smbSpecAccumSpeed = std::min(smbSpecAccumSpeed, maxAccum);
            const float4 spec = inSpec.at(x, y);
            {   // curvature along the predicted motion
                float2 deltaUv = f2(smbPixelUv.x, smbPixelUv.y); { const float2 p = GetScreenUv(m.worldToClipPrev, Xprev + m.cameraDelta); deltaUv = f2((deltaUv.x - p.x) * rectSize.x, (deltaUv.y - p.y) * rectSize.y); }
                { const float d = std::max(smbParallax1, 1.0f / 256.0f); deltaUv = f2(deltaUv.x / d, deltaUv.y / d); }
                auto edgePoint = [&](float2 uvOffset) { const float3 xv = ReconstructViewPosition(f2(pixelUv.x + uvOffset.x, pixelUv.y + uvOffset.y), c, 1.0f); const float3 xw = viewToWorldRotate(m.viewToWorld, xv); const float3 v = normalize(-xw); return v * (dot(X, N) / dot(N, v)); };
                const float3 x10 = edgePoint(f2(rectSizeInv.x, 0)), x01 = edgePoint(f2(0, rectSizeInv.y));
                const float3 n10 = xyz(normalAt(x + 1, y)), n01 = xyz(normalAt(x, y + 1));
                float2 w = f2(fabsf(deltaUv.x) + 1.0f / 256.0f, fabsf(deltaUv.y) + 1.0f / 256.0f); { const float s = w.x + w.y; w = f2(w.x / s, w.y / s); }
                float3 xe = x10 * w.x + x01 * w.y, ne = normalize(n10 * w.x + n01 * w.y);
                float deltaUvLenFixed = smbParallaxMin;
                deltaUvLenFixed *= 1.0f + tp.framerateScale * (float(((uint(x) & 3u) + ((uint(y) & 3u) << 2) + c.frameIndex) & 15u) / 16.0f);     // Sequence::Bayer4x4 dither (ordering of the 4x4 matrix: MathLib's; any permutation dithers alike)
                float2 motionUvHigh = f2(pixelUv.x + deltaUvLenFixed * deltaUv.x * rectSizeInv.x, pixelUv.y + deltaUvLenFixed * deltaUv.y * rectSizeInv.y);
                motionUvHigh = f2((floorf(motionUvHigh.x * rectSize.x) + 0.5f) * rectSizeInv.x, (floorf(motionUvHigh.y * rectSize.y) + 0.5f) * rectSizeInv.y);
                if (deltaUvLenFixed > 1.0f && motionUvHigh.x > 0 && motionUvHigh.y > 0 && motionUvHigh.x < 1 && motionUvHigh.y < 1)
                {
                    const int hx = clampi(int(floorf(motionUvHigh.x * rectSize.x)), 0, W - 1), hy = clampi(int(floorf(motionUvHigh.y * rectSize.y)), 0, H - 1);
                    const float zHigh = in.unpackViewZ(hx, hy, c);
                    const float3 xHigh = viewToWorldRotate(m.viewToWorld, ReconstructViewPosition(motionUvHigh, c, zHigh));
                    const float zError = fabsf(zHigh - viewZ) / std::max(zHigh, viewZ);
                    if (zError < 0.1f) { ne = xyz(normalAt(hx, hy)); xe = xHigh; }
                }
                const float3 edge = xe - X;
                curvature = dot(ne - N, edge) * PositiveRcp(dot(edge, edge));
            }
            // virtual motion
            const float3 Xvirtual = GetXvirtual(hitDistForTracking, curvature, X, Xprev, N, V, roughness);
            const float XvirtualLength = length(Xvirtual);
            const float2 vmbPixelUv = GetScreenUv(m.worldToClipPrev, Xvirtual);
            float2 vmbDelta = f2(vmbPixelUv.x - smbPixelUv.x, vmbPixelUv.y - smbPixelUv.y);
            const float vmbPixelsTraveled = length(f2(vmbDelta.x * rectSize.x, vmbDelta.y * rectSize.y));
            const Bilinear vmbBil = GetBilinearFilter(vmbPixelUv, rectSize);
            const int vx = int(vmbBil.origin.x), vy = int(vmbBil.origin.y);
            float2 rrw = GetRelaxedRoughnessWeightParams(roughness * roughness, c.s.roughnessFraction, kRoughnessSensitivityInTA);
            float rwgt[4]; const int qx[4] = { 0, 1, 0, 1 }, qy[4] = { 0, 0, 1, 1 };
            for (int k = 0; k < 4; k++) { const float r = prevNormalAt(vx + qx[k], vy + qy[k]).w; rwgt[k] = lerp(SmoothStep(1.0f, 0.0f, smbParallaxMax), 1.0f, ComputeNonExponentialWeightWithSigma(r * r, rrw.x, rrw.y, roughnessSigma)); }
            float virtualHistoryRoughnessBasedConfidence = ApplyBilinearFilter(rwgt[0], rwgt[1], rwgt[2], rwgt[3], vmbBil);
            auto stochasticPrevNormal = [&](float2 uv) {      // StochasticBilinear + nearest fetch of the previous normal / roughness
                const Bilinear f = GetBilinearFilter(uv, rectSize); const float r0 = rng.GetFloat(), r1 = rng.GetFloat();
                return prevNormalAt(int(f.origin.x) + (r0 < f.weights.x ? 1 : 0), int(f.origin.y) + (r1 < f.weights.y ? 1 : 0)); };
            const float4 vmbNormalAndRoughness = stochasticPrevNormal(vmbPixelUv);
            const float3 vmbN = xyz(vmbNormalAndRoughness);
            const float Dfactor = GetSpecularDominantFactor(NoV, roughness);
            float virtualHistoryNormalBasedConfidence = 1.0f / (1.0f + 0.5f * Dfactor * saturate(length(N - vmbN) - kNormalEncodingError) * vmbPixelsTraveled);
            if (smbFootprintQuality == 0.0f) smbNavg = vmbN;
            float vmbOcc[4];
            {
                float thr = disocclusionThreshold * frustumSize * lerp(0.25f, 1.0f, NoV);
                thr *= dot(vmbN, N) > thresholdAngle ? 1.0f : 0.0f; thr *= dot(vmbN, smbNavg) > thresholdAngle ? 1.0f : 0.0f;
                const float4 s = inScreenBilinear(vmbBil.origin); const float sArr[4] = { s.x, s.y, s.z, s.w };
                const float3 vmbVv = f3(vmbPixelUv.x * m.frustumPrev[2] + m.frustumPrev[0], vmbPixelUv.y * m.frustumPrev[3] + m.frustumPrev[1], 1.0f), vmbV = rotatePrevInverse(vmbVv);
                const float NoXcurr = dot(N, Xprev - m.cameraDelta);
                for (int k = 0; k < 4; k++)
                {
                    const float z = prevZ(vx + qx[k], vy + qy[k]);
                    const float NoXprev = (N.x * vmbV.x + N.y * vmbV.y) * z + N.z * vmbV.z * z;
                    vmbOcc[k] = (fabsf(NoXprev - NoXcurr) <= thr * sArr[k] - NRD_EPS ? 1.0f : 0.0f) * (rwgt[k] >= 0.5f ? 1.0f : 0.0f);
                    vmbOcc[k] *= CompareMaterials(materialID, UnpackInternalData(prevInternal(vx + qx[k], vy + qy[k])).z, c.s.minMaterialForSpecular) ? 1.0f : 0.0f;
                }
            }
            fbits += vmbOcc[0] * 16.0f + vmbOcc[1] * 32.0f + vmbOcc[2] * 64.0f + vmbOcc[3] * 128.0f;
            const float4 vmbOcclusionWeights = GetBilinearCustomWeights(vmbBil, f4(vmbOcc[0], vmbOcc[1], vmbOcc[2], vmbOcc[3]));
            float vmbSpecAccumSpeed = ApplyBilinearCustomWeights(UnpackInternalData(prevInternal(vx, vy)).y, UnpackInternalData(prevInternal(vx + 1, vy)).y, UnpackInternalData(prevInternal(vx, vy + 1)).y, UnpackInternalData(prevInternal(vx + 1, vy + 1)).y, vmbOcclusionWeights);
            const float vmbFootprintQuality = Sqrt01(ApplyBilinearFilter(vmbOcc[0], vmbOcc[1], vmbOcc[2], vmbOcc[3], vmbBil));
            vmbSpecAccumSpeed *= lerp(vmbFootprintQuality, 1.0f, 1.0f / (1.0f + vmbSpecAccumSpeed));
            const bool vmbAllowCatRom = (vmbOcc[0] + vmbOcc[1] + vmbOcc[2] + vmbOcc[3]) > 3.5f && smbAllowCatRom;
            float curvatureAngleTan = pixelSize * fabsf(curvature); curvatureAngleTan *= std::max(vmbPixelsTraveled / std::max(NoV, 0.01f), 1.0f); curvatureAngleTan *= 2.0f;
            const float curvatureAngle = atanf(curvatureAngleTan);
            const float lobeTanHalfAngle = GetSpecularLobeTanHalfAngle(roughnessModified, 0.75f / (1.0f + vmbSpecAccumSpeed));
            const float lobeHalfAngle = std::max(atanf(lobeTanHalfAngle), kNormalEncodingError);
            float normalWeight = GetEncodingAwareNormalWeight(N, vmbN, lobeHalfAngle, curvatureAngle, kNormalEncodingError);
            normalWeight = lerp(SmoothStep(1.0f, 0.0f, vmbPixelsTraveled), 1.0f, normalWeight);
            virtualHistoryNormalBasedConfidence = std::min(virtualHistoryNormalBasedConfidence, normalWeight);
            virtualHistoryAmount = SmoothStep(0.05f, 0.95f, Dfactor) * virtualHistoryNormalBasedConfidence;
            float virtualHistoryParallaxBasedConfidence;
            {
                const float hitDistForTrackingPrev = sampleBilinear(h.specHitDistForTracking, c.W, c.H, f2(vmbPixelUv.x * rectSize.x, vmbPixelUv.y * rectSize.y));
                const float2 vmbPixelUvPrev = GetScreenUv(m.worldToClipPrev, GetXvirtual(hitDistForTrackingPrev, curvature, X, Xprev, N, V, roughness));
                const float pixelSizeAtXvirtual = PixelRadiusToWorld(c, 1.0f, XvirtualLength);
                const float r = std::max((lobeTanHalfAngle + curvatureAngleTan) * std::min(hitDistForTracking, hitDistForTrackingPrev) / pixelSizeAtXvirtual, 0.1f);
                const float d = length(f2((vmbPixelUvPrev.x - vmbPixelUv.x) * rectSize.x, (vmbPixelUvPrev.y - vmbPixelUv.y) * rectSize.y));
                virtualHistoryParallaxBasedConfidence = LinearStep(r, 0.0f, d);
            }
            {   // prev-prev test along the virtual motion (1 iteration)
                const float stepBetweenTaps = std::min(vmbPixelsTraveled * tp.framerateScale, 2.0f) + vmbPixelsTraveled;
                const float inv = 1.0f / sqrtf(dot(vmbDelta, vmbDelta)); vmbDelta = f2(vmbDelta.x * inv / rectSize.x, vmbDelta.y * inv / rectSize.y);
                rrw = GetRelaxedRoughnessWeightParams(vmbNormalAndRoughness.w * vmbNormalAndRoughness.w, c.s.roughnessFraction, kRoughnessSensitivityInTA);
                const float2 uvPrev = f2(vmbPixelUv.x + vmbDelta.x * stepBetweenTaps, vmbPixelUv.y + vmbDelta.y * stepBetweenTaps);
                if (std::isfinite(uvPrev.x) && std::isfinite(uvPrev.y) && uvPrev.x > 0 && uvPrev.y > 0 && uvPrev.x < 1 && uvPrev.y < 1)
                {
                    const float4 nrPrev = stochasticPrevNormal(uvPrev);
                    float wn = GetEncodingAwareNormalWeight(vmbN, xyz(nrPrev), lobeHalfAngle, curvatureAngle * (1.0f + stepBetweenTaps), kNormalEncodingError);
                    float wr = ComputeNonExponentialWeightWithSigma(nrPrev.w * nrPrev.w, rrw.x, rrw.y, roughnessSigma);
                    wn = lerp(1.0f, wn, saturate(stepBetweenTaps)); wr = lerp(1.0f, wr, saturate(stepBetweenTaps));
                    virtualHistoryNormalBasedConfidence = std::min(virtualHistoryNormalBasedConfidence, wn); virtualHistoryRoughnessBasedConfidence = std::min(virtualHistoryRoughnessBasedConfidence, wr);
                }
            }
            const float virtualHistoryConfidenceForSmbRelaxation = virtualHistoryNormalBasedConfidence * virtualHistoryRoughnessBasedConfidence;
            const float virtualHistoryConfidence = virtualHistoryConfidenceForSmbRelaxation * virtualHistoryParallaxBasedConfidence;
            virtualHistoryAmount *= virtualHistoryRoughnessBasedConfidence;
            float4 smbSpecHistory; float smbSpecFastHistory; sampleHistory(smbSamplePos, smbOcclusionWeights, smbAllowCatRom, h.spec, smbSpecHistory, h.specFast, smbSpecFastHistory);
            float surfaceHistoryConfidence;
            {
                const float a = atanf(smbParallaxMax * pixelSize / length(X));
                const float nl = 1.0f / (1.0f + smbSpecAccumSpeed);
                const float hd = lerp(smbSpecHistory.w, spec.w, nl) * hitDistNormalization;
                float tana0 = GetSpecularLobeTanHalfAngle(roughnessModified, 0.75f); tana0 *= lerp(NoV, 1.0f, roughnessModified); tana0 *= nl; tana0 /= saturate(hd / frustumSize) + NRD_EPS;
                const float a0 = std::max(atanf(tana0), kNormalEncodingError);
                surfaceHistoryConfidence = Pow01(LinearStep(a0, 0.0f, a), 4.0f);
            }
            float2 maxResponsiveFrameNum;
            {
                const float amount = (roughness + NRD_EPS) / (tp.responsiveAccumulationRoughnessThreshold + NRD_EPS), responsiveFactor = SmoothStep01(amount), smc = GetSpecMagicCurve(roughnessModified);
                const float fa = lerp(smc, 1.0f, responsiveFactor), pw = lerp(32.0f, 1.0f, smc) * (1.0f - responsiveFactor);
                maxResponsiveFrameNum = f2(std::max(maxAccum * fa * Pow01(dot(N, normalize(smbNavg)), pw), historyFixFrameNum), std::max(maxAccum * fa * Pow01(dot(N, vmbN), pw), historyFixFrameNum));
            }
            float smbMaxFrameNum = std::min(maxAccum * surfaceHistoryConfidence, maxResponsiveFrameNum.x);
            const float smbBoostedMaxFrameNum = std::max(smbMaxFrameNum, historyFixFrameNum * (1.0f - virtualHistoryConfidenceForSmbRelaxation));
            const float smbSpecAccumSpeedBoosted = std::min(smbSpecAccumSpeed, smbBoostedMaxFrameNum);
            const float vmbMaxFrameNum = std::min(maxAccum * virtualHistoryConfidence, maxResponsiveFrameNum.y);
            smbSpecAccumSpeed = std::min(smbSpecAccumSpeed, smbMaxFrameNum); vmbSpecAccumSpeed = std::min(vmbSpecAccumSpeed, vmbMaxFrameNum);
            const float virtualHistoryAmountUnbiased = virtualHistoryAmount;
            {
                const float magic = vmbSpecAccumSpeed > smbSpecAccumSpeed ? 8.0f : 0.5f;
                virtualHistoryAmount *= 1.0f + (vmbSpecAccumSpeed - smbSpecAccumSpeed) / (magic * std::max(vmbSpecAccumSpeed, smbSpecAccumSpeed) + 1.0f);
                virtualHistoryAmount = saturate(virtualHistoryAmount);
            }
            float4 vmbSpecHistory; float vmbSpecFastHistory; sampleHistory(f2(saturate(vmbPixelUv.x) * rectSize.x, saturate(vmbPixelUv.y) * rectSize.y), vmbOcclusionWeights, vmbAllowCatRom, h.spec, vmbSpecHistory, h.specFast, vmbSpecFastHistory);
            smbSpecHistory = ClampNegativeToZero(smbSpecHistory); vmbSpecHistory = ClampNegativeToZero(vmbSpecHistory);
            const float smbNL = 1.0f / (1.0f + smbSpecAccumSpeed), vmbNL = 1.0f / (1.0f + vmbSpecAccumSpeed);
            const float4 smbSpec = MixHistoryAndCurrent(smbSpecHistory, spec, smbNL, roughnessModified, maxAccum), vmbSpec = MixHistoryAndCurrent(vmbSpecHistory, spec, vmbNL, roughnessModified, maxAccum);
            float4 specResult = smbSpec * (1.0f - virtualHistoryAmount) + vmbSpec * virtualHistoryAmount;
            specAccumSpeed = lerp(smbSpecAccumSpeedBoosted, vmbSpecAccumSpeed, virtualHistoryAmount);
            const float4 specHistory = smbSpecHistory * (1.0f - virtualHistoryAmount) + vmbSpecHistory * virtualHistoryAmount;
            const float specMaxRelativeIntensity = tp.fireflySuppressorMinRelativeScale + 38.0f / (specAccumSpeed + 1.0f);
            float specAntifireflyFactor = specAccumSpeed * c.gMaxBlurRadius * 0.1f; specAntifireflyFactor /= 1.0f + specAntifireflyFactor;
            { const float l = specResult.x; float lc = std::min(l, specHistory.x * specMaxRelativeIntensity); lc = lerp(l, lc, specAntifireflyFactor); specResult = ChangeLuma(specResult, lc); }
            out.spec->store(x, y, specResult);
            auto nonLinearFast = [&](float accumSpeed, float confidence) { return std::max(1.0f - confidence, 1.0f / (1.0f + std::min(accumSpeed, maxFastAccum))); };      // GetNonLinearAccumSpeed, hasData = true
            const float smbSpecFast = lerp(smbSpecFastHistory, spec.x, nonLinearFast(smbSpecAccumSpeed, surfaceHistoryConfidence)), vmbSpecFast = lerp(vmbSpecFastHistory, spec.x, nonLinearFast(vmbSpecAccumSpeed, virtualHistoryConfidence));
            float specFastResult = lerp(smbSpecFast, vmbSpecFast, virtualHistoryAmountUnbiased);
            specFastResult = lerp(specFastResult, std::min(specFastResult, specHistory.x * specMaxRelativeIntensity * 4.0f), specAntifireflyFactor);
            (*out.specFast)[pix] = lp(specFastResult);
        }
        {   // PackData2
            uint p = uint(fbits + 0.5f); p |= uint(saturate(virtualHistoryAmount) * 127.0f + 0.5f) << 8; p |= smbAllowCatRom ? (1u << 15) : 0u; p |= f32tof16(curvature) << 16;
            (*out.data2)[pix] = p;
        }
        // ---------------------------------------------------------------- diffuse ----------------------------------------------------------------
        {
            diffAccumSpeed *= lerp(smbFootprintQuality, 1.0f, 1.0f / (1.0f + diffAccumSpeed));
            diffAccumSpeed = std::min(diffAccumSpeed, maxAccum);
            const float4 diff = inDiff.at(x, y);
            float4 smbDiffHistory; float smbDiffFastHistory; sampleHistory(smbSamplePos, smbOcclusionWeights, smbAllowCatRom, h.diff, smbDiffHistory, h.diffFast, smbDiffFastHistory);
            smbDiffHistory = ClampNegativeToZero(smbDiffHistory);
            const float nl = 1.0f / (1.0f + diffAccumSpeed);
            float4 diffResult = MixHistoryAndCurrent(smbDiffHistory, diff, nl, 1.0f, maxAccum);
            const float diffMaxRelativeIntensity = tp.fireflySuppressorMinRelativeScale + 38.0f / (diffAccumSpeed + 1.0f);
            float diffAntifireflyFactor = diffAccumSpeed * c.gMaxBlurRadius * 0.1f; diffAntifireflyFactor /= 1.0f + diffAntifireflyFactor;
            { const float l = diffResult.x; float lc = std::min(l, smbDiffHistory.x * diffMaxRelativeIntensity); lc = lerp(l, lc, diffAntifireflyFactor); diffResult = ChangeLuma(diffResult, lc); }
            out.diff->store(x, y, diffResult);
            float diffFastResult = lerp(smbDiffFastHistory, diff.x, 1.0f / (1.0f + std::min(diffAccumSpeed, maxFastAccum)));
            diffFastResult = lerp(diffFastResult, std::min(diffFastResult, smbDiffHistory.x * diffMaxRelativeIntensity * 4.0f), diffAntifireflyFactor);
            (*out.diffFast)[pix] = lp(diffFastResult);
        }
        (*out.data1)[pix * 2] = uint8_t(saturate(diffAccumSpeed / 63.0f) * 255.0f + 0.5f); (*out.data1)[pix * 2 + 1] = uint8_t(saturate(specAccumSpeed / 63.0f) * 255.0f + 0.5f);      // PackData1 into RG8_UNORM
    }
}


// =====================================================================================================================================================
// HistoryFix (External/Nrd/Shaders/Include/REBLUR_HistoryFix.hlsli:11-496): pixels whose history is shorter than historyFixFrameNum are rebuilt from a sparse 5x5
// (no corners) cross with a stride that shrinks as history grows; then the luminance is clamped to the fast history's local 5x5 statistics and, with the anti-firefly
// option (RTXPT: on), to the statistics of the 9x9 ring around the 3x3 centre.
// =====================================================================================================================================================
inline void historyFix(const Constants& c, const Inputs& in, const std::vector<uint8_t>& tiles, const std::vector<uint8_t>& data1, const Image4& inDiff, const Image4& inSpec,
                       const std::vector<float>& inDiffFast, const std::vector<float>& inSpecFast, bool antiFirefly, float historyFixBasePixelStride, bool historyValid,
                       Image4& outDiff, Image4& outSpec, std::vector<float>& outDiffFast, std::vector<float>& outSpecFast)
{
    const int W = int(c.W), H = int(c.H);
    const float historyFixFrameNum = float(c.s.historyFixFrameNum), maxAccum = historyValid ? float(std::min(c.s.maxAccumulatedFrameNum, 63u)) : 0.0f, maxFast = historyValid ? float(c.s.maxFastAccumulatedFrameNum) : 0.0f;
    const float2 rectSizeInv = f2(1.0f / float(W), 1.0f / float(H));
    auto frames = [&](int x, int y) { const size_t p = (size_t(clampi(y, 0, H - 1)) * W + clampi(x, 0, W - 1)) * 2; return f2(float(data1[p]) / 255.0f * 63.0f, float(data1[p + 1]) / 255.0f * 63.0f); };
    auto fastAt = [&](const std::vector<float>& img, int x, int y) { return img[size_t(clampi(y, 0, H - 1)) * W + clampi(x, 0, W - 1)]; };
    #pragma omp parallel for schedule(dynamic, 4)          // pixels are independent within a pass (each writes its own outputs only)
    for (int y = 0; y < H; y++) for (int x = 0; x < W; x++)
    {
        if (tileIsSky(c, tiles, x, y)) continue;
        const float viewZ = in.unpackViewZ(x, y, c);
        if (viewZ > c.s.denoisingRange) continue;
        float materialID; const float4 nr = in.unpackNormalRoughness(x, y, materialID);
        const float3 N = xyz(nr), Nv = worldToViewRotate(c, N); const float roughness = nr.w;
        const float frustumSize = GetFrustumSize(c, viewZ);
        const float2 pixelUv = f2((float(x) + 0.5f) * rectSizeInv.x, (float(y) + 0.5f) * rectSizeInv.y);
        const float3 Xv = ReconstructViewPosition(pixelUv, c, viewZ);
        const float2 frameNum = frames(x, y);
        float2 stride;
        {
            float2 avg = frameNum, sum = f2(1, 1); const float inv = 1.0f / (historyFixFrameNum + NRD_EPS);
            for (int i = -1; i <= 1; i++) for (int j = -1; j <= 1; j++)
            {
                if (i == 0 && j == 0) continue;
                const float2 f = frames(x + i, y + j); const float2 w = f2(f.x >= frameNum.x ? 1.0f : 0.0f, f.y >= frameNum.y ? 1.0f : 0.0f);
                avg = f2(avg.x + saturate(f.x * inv) * w.x, avg.y + saturate(f.y * inv) * w.y); sum = sum + w;
            }
            avg = f2(avg.x / sum.x, avg.y / sum.y);
            stride = f2(historyFixBasePixelStride / (2.0f + avg.x * historyFixFrameNum) * (frameNum.x < historyFixFrameNum ? 1.0f : 0.0f), historyFixBasePixelStride / (2.0f + avg.y * historyFixFrameNum) * (frameNum.y < historyFixFrameNum ? 1.0f : 0.0f));
        }
        for (int channel = 0; channel < 2; channel++)
        {
            const bool isSpec = channel == 1;
            const Image4& src = isSpec ? inSpec : inDiff; const std::vector<float>& fast = isSpec ? inSpecFast : inDiffFast;
            float4 v = src.at(x, y);
            const float smc = GetSpecMagicCurve(roughness), fn = isSpec ? frameNum.y : frameNum.x, r = isSpec ? roughness : 1.0f;
            float st = isSpec ? stride.y * lerp(0.5f, 1.0f, smc) : stride.x; st = floorf(st);
            if (st != 0.0f)
            {
                const int sti = int(st + 0.5f);
                const float nl = 1.0f / (1.0f + fn);
                const float normalW = GetNormalWeightParam(nl, c.gLobeAngleFraction, r);
                const float2 gw = GetGeometryWeightParams(c.s.planeDistanceSensitivity, frustumSize, Xv, Nv), rw = GetRelaxedRoughnessWeightParams(roughness * roughness, sqrtf(c.s.roughnessFraction));
                const float hitDistScale = GetHitDistanceNormalization(viewZ, c.s.hitDistParams, r), hitDist = v.w * hitDistScale;
                const float2 hw = GetHitDistanceWeightParams(saturate(hitDist / frustumSize), nl, r);
                float sum = 1.0f + fn; v = v * sum;
                for (int j = -2; j <= 2; j++) for (int i = -2; i <= 2; i++)
                {
                    if ((i == 0 && j == 0) || (std::abs(i) + std::abs(j) == 4)) continue;
                    const float2 uv = f2(pixelUv.x + float(i) * st * rectSizeInv.x, pixelUv.y + float(j) * st * rectSizeInv.y);
                    const int sx = clampi(x + i * sti, 0, W - 1), sy = clampi(y + j * sti, 0, H - 1);
                    const float zs = in.unpackViewZ(sx, sy, c); float ms; const float4 Ns = in.unpackNormalRoughness(sx, sy, ms);
                    float w = (uv.x > 0 && uv.y > 0 && uv.x < 1 && uv.y < 1) ? 1.0f : 0.0f;
                    w *= ComputeWeight(dot(Nv, ReconstructViewPosition(uv, c, zs)), gw.x, gw.y);
                    w *= CompareMaterials(materialID, ms, isSpec ? c.s.minMaterialForSpecular : c.s.minMaterialForDiffuse) ? 1.0f : 0.0f;
                    w *= ComputeExponentialWeight(AcosApprox(dot(xyz(Ns), N)), normalW, 0.0f);
                    if (isSpec) w *= ComputeExponentialWeight(Ns.w * Ns.w, rw.x, rw.y);
                    const float2 fs = frames(sx, sy); w *= 1.0f + (isSpec ? fs.y : fs.x);
                    float4 sv = src.at(sx, sy); if (w == 0.0f) sv = f4(0, 0, 0, 0);
                    const float hs = sv.w * hitDistScale;
                    w *= ComputeExponentialWeight(saturate(hs / frustumSize), hw.x, hw.y);
                    if (isSpec) { const float d = fabsf(hitDist - hs) / (std::max(hitDist, hs) + 0.001f), b = LinearStep(0.03f, 0.05f, roughness); w *= SmoothStep(0.2f + b, 0.05f + b, d); }
                    sum += w; v = v + sv * w;
                }
                v = v * PositiveRcp(sum);
            }
            float center = fastAt(fast, x, y), m1 = center, m2 = center * center;
            float f = saturate(fn / (historyFixFrameNum + NRD_EPS)); if (isSpec) f = lerp(1.0f, f, smc);
            center = lerp(v.x, center, f);
            (isSpec ? outSpecFast : outDiffFast)[size_t(y) * W + x] = lp(center);
            for (int j = -2; j <= 2; j++) for (int i = -2; i <= 2; i++) { if (i == 0 && j == 0) continue; const float d = fastAt(fast, x + i, y + j); m1 += d; m2 += d * d; }
            float luma = v.x;
            if (antiFirefly)
            {
                float a1 = 0, a2 = 0;
                for (int j = -4; j <= 4; j++) for (int i = -4; i <= 4; i++) { if (std::abs(i) <= 1 && std::abs(j) <= 1) continue; const float d = fastAt(fast, x + i, y + j); a1 += d; a2 += d * d; }
                a1 /= 72.0f; a2 /= 72.0f;
                const float sigma = sqrtf(fabsf(a2 - a1 * a1)) * 2.0f;
                luma = clampf(luma, a1 - sigma, a1 + sigma);
            }
            m1 /= 25.0f; m2 /= 25.0f;
            const float sigma = sqrtf(fabsf(m2 - m1 * m1)) * 2.0f;
            const float clamped = clampf(luma, m1 - sigma, m1 + sigma);
            luma = lerp(clamped, luma, 1.0f / (1.0f + (maxFast < maxAccum ? 1.0f : 0.0f) * fn * 2.0f));
            v = ChangeLuma(v, luma);
            (isSpec ? outSpec : outDiff).store(x, y, v);
        }
    }
}


// =====================================================================================================================================================
// TemporalStabilization (External/Nrd/Shaders/Include/REBLUR_TemporalStabilization.hlsli:11-369): the luminance of the post-blurred signal is blended with a
// reprojected, variance-clamped history of stabilised luminance (surface motion; for specular also virtual motion), gated by the antilag term; history lengths are
// incremented and written back for the next frame.  The motion-vector patching for specular-dominant pixels needs IN_BASECOLOR_METALNESS, which RTXPT does not supply.
// =====================================================================================================================================================
inline float sampleLumaHistory(float2 samplePos, float4 bilinearCustomWeights, bool useBicubic, const std::vector<float>& tex, uint W, uint H)
{   // BicubicFilterNoCornersWithFallbackToBilinearFilterWithCustomWeights1
    const float S = 0.5f;
    const float2 centerPos = f2(floorf(samplePos.x - 0.5f) + 0.5f, floorf(samplePos.y - 0.5f) + 0.5f), f = f2(saturate(samplePos.x - centerPos.x), saturate(samplePos.y - centerPos.y));
    auto w0f = [&](float t) { return t * (t * (-S * t + 2.0f * S) - S); }; auto w1f = [&](float t) { return t * (t * ((2.0f - S) * t - (3.0f - S))) + 1.0f; };
    auto w2f = [&](float t) { return t * (t * (-(2.0f - S) * t + (3.0f - 2.0f * S)) + S); }; auto w3f = [&](float t) { return t * (t * (S * t - S)); };
    const float2 w12 = f2(w1f(f.x) + w2f(f.x), w1f(f.y) + w2f(f.y)), tc = f2(w2f(f.x) / w12.x, w2f(f.y) / w12.y);
    float4 w = f4(w12.x * w0f(f.y), w0f(f.x) * w12.y, w12.x * w12.y, w3f(f.x) * w12.y); float w4 = w12.x * w3f(f.y);
    if (!useBicubic) { w = bilinearCustomWeights; w4 = 0.0f; }
    const float sum = w.x + w.y + w.z + w.w + w4;
    float color;
    if (useBicubic) color = sampleBilinear(tex, W, H, f2(centerPos.x + tc.x, centerPos.y - 1)) * w.x + sampleBilinear(tex, W, H, f2(centerPos.x - 1, centerPos.y + tc.y)) * w.y + sampleBilinear(tex, W, H, f2(centerPos.x + tc.x, centerPos.y + tc.y)) * w.z
                        + sampleBilinear(tex, W, H, f2(centerPos.x + 2, centerPos.y + tc.y)) * w.w + sampleBilinear(tex, W, H, f2(centerPos.x + tc.x, centerPos.y + 2)) * w4;
    else color = sampleBilinear(tex, W, H, centerPos) * w.x + sampleBilinear(tex, W, H, f2(centerPos.x + 1, centerPos.y)) * w.y + sampleBilinear(tex, W, H, f2(centerPos.x, centerPos.y + 1)) * w.z + sampleBilinear(tex, W, H, f2(centerPos.x + 1, centerPos.y + 1)) * w.w;
    return sum < 0.0001f ? 0.0f : color / sum;
}
struct StabilizationParams { float antilagSigmaScale = 4.0f, antilagSensitivity = 3.0f, stabilizationStrength = 63.0f / 64.0f; };      // nrd::ReblurAntilagSettings, maxStabilizedFrameNum = 63
inline void temporalStabilization(const Constants& c, const FrameMatrices& m, const TemporalParams& tp, const StabilizationParams& sp, const Inputs& in, const std::vector<uint8_t>& tiles,
                                  const std::vector<uint8_t>& data1, const std::vector<uint32_t>& data2, const Image4& inDiff, const Image4& inSpec, const std::vector<float>& specHitDistForTracking,
                                  const History& h, Image4& outDiff, Image4& outSpec, std::vector<float>& outDiffLuma, std::vector<float>& outSpecLuma, std::vector<uint16_t>& outInternalData)
{
    const int W = int(c.W), H = int(c.H);
    const float2 rectSize = f2(float(W), float(H)), rectSizeInv = f2(1.0f / float(W), 1.0f / float(H));
    const float historyFixFrameNum = float(c.s.historyFixFrameNum), strength = (tp.resetHistory || !h.valid) ? 0.0f : sp.stabilizationStrength;
    auto rotatePrevInverse = [&](float3 v) { const float* t = m.worldToViewPrev; return f3(t[0] * v.x + t[4] * v.y + t[8] * v.z, t[1] * v.x + t[5] * v.y + t[9] * v.z, t[2] * v.x + t[6] * v.y + t[10] * v.z); };
    auto antilag = [&](float history, float avg, float sigma, float accumSpeed) {      // ComputeAntilag, REBLUR_ANTILAG_MODE 2
        const float s = sigma * sp.antilagSigmaScale, magic = sp.antilagSensitivity * tp.framerateScale * tp.framerateScale;
        const float hc = clampf(history, avg - s, avg + s);
        const float d = fabsf(history - hc) / (std::max(history, hc) + NRD_EPS);
        return 1.0f / (1.0f + d * accumSpeed / magic); };
    #pragma omp parallel for schedule(dynamic, 4)          // pixels are independent within a pass (each writes its own outputs only)
    for (int y = 0; y < H; y++) for (int x = 0; x < W; x++)
    {
        if (tileIsSky(c, tiles, x, y)) continue;
        const float viewZ = in.unpackViewZ(x, y, c);
        if (viewZ > c.s.denoisingRange) continue;
        const size_t pix = size_t(y) * W + x;
        const float2 pixelUv = f2((float(x) + 0.5f) * rectSizeInv.x, (float(y) + 0.5f) * rectSizeInv.y);
        const float3 Xv = ReconstructViewPosition(pixelUv, c, viewZ), X = viewToWorldRotate(m.viewToWorld, Xv);
        float3 mv = f3(0);
        if (tp.motion) mv = f3(f16tof32(tp.motion[pix * 4]) * rectSizeInv.x, f16tof32(tp.motion[pix * 4 + 1]) * rectSizeInv.y, f16tof32(tp.motion[pix * 4 + 2]));
        const float2 smbPixelUv = f2(pixelUv.x + mv.x, pixelUv.y + mv.y);
        const float viewZprev = viewZ + mv.z;
        const float3 Xprev = rotatePrevInverse(f3((smbPixelUv.x * m.frustumPrev[2] + m.frustumPrev[0]) * viewZprev, (smbPixelUv.y * m.frustumPrev[3] + m.frustumPrev[1]) * viewZprev, viewZprev)) + m.cameraDelta;
        float materialID; const float4 nr = in.unpackNormalRoughness(x, y, materialID);
        const float3 N = xyz(nr); const float roughness = nr.w;
        float2 d1 = f2(float(data1[pix * 2]) / 255.0f * 63.0f, float(data1[pix * 2 + 1]) / 255.0f * 63.0f);
        const uint p2 = data2[pix], bits = p2 & 0xFFu; const bool smbAllowCatRom = (p2 & (1u << 15)) != 0;
        const float virtualHistoryAmount = float((p2 >> 8) & 127u) / 127.0f, curvature = f16tof32(p2 >> 16);
        const Bilinear smbBil = GetBilinearFilter(smbPixelUv, rectSize);
        const float4 smbOcc = f4((bits & 1u) ? 1.0f : 0.0f, (bits & 2u) ? 1.0f : 0.0f, (bits & 4u) ? 1.0f : 0.0f, (bits & 8u) ? 1.0f : 0.0f);
        const float4 smbW = GetBilinearCustomWeights(smbBil, smbOcc);
        const float smbFootprintQuality = Sqrt01(ApplyBilinearFilter(smbOcc.x, smbOcc.y, smbOcc.z, smbOcc.w, smbBil));
        const float2 smbSamplePos = f2(saturate(smbPixelUv.x) * rectSize.x, saturate(smbPixelUv.y) * rectSize.y);
        auto moments = [&](const Image4& img, float& luma, float& m1, float& sigma) {
            luma = img.at(x, y).x; float a = luma, b = luma * luma, mn = NRD_INF, mx = -NRD_INF;
            for (int j = -1; j <= 1; j++) for (int i = -1; i <= 1; i++) { if (i == 0 && j == 0) continue; const float d = img.at(clampi(x + i, 0, W - 1), clampi(y + j, 0, H - 1)).x; a += d; b += d * d; mn = std::min(mn, d); mx = std::max(mx, d); }
            m1 = a / 9.0f; sigma = sqrtf(fabsf(b / 9.0f - m1 * m1));
            if (c.gMaxBlurRadius != 0.0f) luma = clampf(luma, mn, mx); };       // RCRS
        // diffuse
        {
            float luma, m1, sigma; moments(inDiff, luma, m1, sigma);
            float hist = std::max(sampleLumaHistory(smbSamplePos, smbW, smbAllowCatRom, h.diffLumaStabilized, c.W, c.H), 0.0f);
            const float al = antilag(hist, m1, sigma, smbFootprintQuality * d1.x);
            float wgt = smbFootprintQuality * (d1.x / (1.0f + d1.x)); const float clampScale = 1.0f + 3.0f * tp.framerateScale * wgt;       // GetTemporalAccumulationParams
            wgt *= al;
            hist = clampf(hist, m1 - sigma * clampScale, m1 + sigma * clampScale);
            const float stabilized = lerp(luma, hist, std::min(wgt, strength));
            outDiff.store(x, y, ChangeLuma(inDiff.at(x, y), stabilized)); outDiffLuma[pix] = lp(stabilized);
            d1.x += 1.0f; d1.x = lerp(std::min(d1.x, historyFixFrameNum), d1.x, al);
        }
        // specular
        {
            float luma, m1, sigma; moments(inSpec, luma, m1, sigma);
            const float4 spec = inSpec.at(x, y);
            float hitDistForTracking = spec.w * GetHitDistanceNormalization(viewZ, c.s.hitDistParams, roughness);
            if (c.gSpecPrepassBlurRadius != 0.0f) hitDistForTracking = std::min(hitDistForTracking, specHitDistForTracking[pix]);
            const float3 V = normalize(-X);
            const float2 vmbPixelUv = GetScreenUv(m.worldToClipPrev, GetXvirtual(hitDistForTracking, curvature, X, Xprev, N, V, roughness));
            float smbHist = sampleLumaHistory(smbSamplePos, smbW, smbAllowCatRom, h.specLumaStabilized, c.W, c.H);
            const Bilinear vmbBil = GetBilinearFilter(vmbPixelUv, rectSize);
            const float4 vmbOcc = f4((bits & 16u) ? 1.0f : 0.0f, (bits & 32u) ? 1.0f : 0.0f, (bits & 64u) ? 1.0f : 0.0f, (bits & 128u) ? 1.0f : 0.0f);
            const bool vmbAllowCatRom = (vmbOcc.x + vmbOcc.y + vmbOcc.z + vmbOcc.w) > 3.5f && smbAllowCatRom;
            const float vmbFootprintQuality = Sqrt01(ApplyBilinearFilter(vmbOcc.x, vmbOcc.y, vmbOcc.z, vmbOcc.w, vmbBil));
            float vmbHist = sampleLumaHistory(f2(saturate(vmbPixelUv.x) * rectSize.x, saturate(vmbPixelUv.y) * rectSize.y), GetBilinearCustomWeights(vmbBil, vmbOcc), vmbAllowCatRom, h.specLumaStabilized, c.W, c.H);
            smbHist = std::max(smbHist, 0.0f); vmbHist = std::max(vmbHist, 0.0f);
            float hist = lerp(smbHist, vmbHist, virtualHistoryAmount);
            const float footprintQuality = lerp(smbFootprintQuality, vmbFootprintQuality, virtualHistoryAmount);
            const float al = antilag(hist, m1, sigma, footprintQuality * d1.y);
            float wgt = footprintQuality * (d1.y / (1.0f + d1.y)); const float clampScale = 1.0f + 3.0f * tp.framerateScale * wgt;
            wgt *= al;
            const float responsiveFactor = SmoothStep01((roughness + NRD_EPS) / (tp.responsiveAccumulationRoughnessThreshold + NRD_EPS)), smc = GetSpecMagicCurve(roughness);
            wgt *= lerp(smc, 1.0f, 0.5f + responsiveFactor * 0.5f);
            hist = clampf(hist, m1 - sigma * clampScale, m1 + sigma * clampScale);
            const float stabilized = lerp(luma, hist, std::min(wgt, strength));
            outSpec.store(x, y, ChangeLuma(spec, stabilized)); outSpecLuma[pix] = lp(stabilized);
            d1.y += 1.0f; d1.y = lerp(std::min(d1.y, historyFixFrameNum), d1.y, al);
        }
        outInternalData[pix] = PackInternalData(d1.x, d1.y, materialID);
    }
}

// =====================================================================================================================================================
// One frame of REBLUR_DIFFUSE_SPECULAR as RTXPT configures it (Reblur_DiffuseSpecular.hpp:71-270, Reblur.cpp:98-200): ClassifyTiles -> HitDistReconstruction 5x5 ->
// PrePass -> TemporalAccumulation -> HistoryFix -> Blur -> PostBlur -> TemporalStabilization, with the resource routing of the dispatch graph.
// =====================================================================================================================================================
struct FrameOutputs { Image4 diff, spec; std::vector<uint8_t> data1; std::vector<uint32_t> data2; bool keepStages = false; std::vector<Image4> stageDiff, stageSpec; };      // stage*: the images after each of the six middle passes (debugging aid for tests/test_reblur_port.py)
inline void denoiseFrame(const Settings& settings, uint W, uint H, const float* worldToView16, const float* viewToClip16, const float* worldToViewPrev16, const float* viewToClipPrev16, uint frameIndex,
                         const TemporalParams& tp, const Inputs& in, const Image4& inDiff, const Image4& inSpec, History& h, FrameOutputs& out)
{
    const Constants c = makeConstants(settings, W, H, worldToView16, viewToClip16, frameIndex);
    if (h.W != W || h.H != H) h.init(W, H);
    FrameMatrices m;
    memcpy(m.viewToWorld, c.viewToWorld, sizeof(m.viewToWorld));
    auto camPos = [](const float* wv) { return f3(-(wv[12] * wv[0] + wv[13] * wv[1] + wv[14] * wv[2]), -(wv[12] * wv[4] + wv[13] * wv[5] + wv[14] * wv[6]), -(wv[12] * wv[8] + wv[13] * wv[9] + wv[14] * wv[10])); };   // -t * R^T
    const float3 pos = camPos(worldToView16), posPrev = camPos(worldToViewPrev16);
    m.cameraDelta = posPrev - pos;
    auto relativeWorldToClip = [](const float* wv, float3 camRel, const float* vc, float* outM, float* outAffine) {
        // world (relative to the CURRENT camera) -> view of the camera sitting at camRel: Xv = R * ( X - camRel ); then * viewToClip
        float R[9]; for (int r = 0; r < 3; r++) for (int k = 0; k < 3; k++) R[r * 3 + k] = wv[k * 4 + r];
        const float t[3] = { -(R[0] * camRel.x + R[1] * camRel.y + R[2] * camRel.z), -(R[3] * camRel.x + R[4] * camRel.y + R[5] * camRel.z), -(R[6] * camRel.x + R[7] * camRel.y + R[8] * camRel.z) };
        if (outAffine) for (int r = 0; r < 3; r++) { outAffine[r * 4] = R[r * 3]; outAffine[r * 4 + 1] = R[r * 3 + 1]; outAffine[r * 4 + 2] = R[r * 3 + 2]; outAffine[r * 4 + 3] = t[r]; }
        float wvRel[16] = { R[0], R[3], R[6], 0, R[1], R[4], R[7], 0, R[2], R[5], R[8], 0, t[0], t[1], t[2], 1 };       // row-major, row vector x matrix
        for (int r = 0; r < 4; r++) for (int k = 0; k < 4; k++) { float a = 0; for (int j = 0; j < 4; j++) a += wvRel[r * 4 + j] * vc[j * 4 + k]; outM[r * 4 + k] = a; } };
    relativeWorldToClip(worldToView16, f3(0), viewToClip16, m.worldToClip, nullptr);
    relativeWorldToClip(worldToViewPrev16, m.cameraDelta, viewToClipPrev16, m.worldToClipPrev, m.worldToViewPrev);
    { const float P00 = viewToClipPrev16[0], P11 = viewToClipPrev16[5], P20 = viewToClipPrev16[8], P21 = viewToClipPrev16[9]; m.frustumPrev[0] = (-1.0f - P20) / P00; m.frustumPrev[2] = 2.0f / P00; m.frustumPrev[1] = (1.0f - P21) / P11; m.frustumPrev[3] = -2.0f / P11; }
    const size_t n = size_t(W) * H;
    const std::vector<uint8_t> tiles = classifyTiles(c, in);
    Image4 t1d = inDiff, t1s = inSpec, t2d = inDiff, t2s = inSpec;
    auto keep = [&](const Image4& d, const Image4& s) { if (out.keepStages) { out.stageDiff.push_back(d); out.stageSpec.push_back(s); } };
    out.stageDiff.clear(); out.stageSpec.clear();
    hitDistReconstruction(c, in, tiles, inDiff, inSpec, t2d, t2s); keep(t2d, t2s);
    std::vector<float> prepassTracking(n, 0.0f);
    { SpatialOutputs o{ &t1d, &t1s, &prepassTracking }; spatialPass(c, in, tiles, PRE_BLUR, t2d, t2s, nullptr, o); keep(t1d, t1s); }
    std::vector<float> diffFastT(n, 0.0f), specFastT(n, 0.0f), trackingPong(n, 0.0f);
    out.data1.assign(n * 2, 0); out.data2.assign(n, 0u);
    { TemporalOutputs o{ &t2d, &t2s, &diffFastT, &specFastT, &trackingPong, &out.data1, &out.data2 }; temporalAccumulation(c, m, tp, in, tiles, t1d, t1s, prepassTracking, h, o); keep(t2d, t2s); }
    historyFix(c, in, tiles, out.data1, t2d, t2s, diffFastT, specFastT, true, 14.0f, h.valid && !tp.resetHistory, t1d, t1s, h.diffFast, h.specFast); keep(t1d, t1s);
    std::vector<float2> frames(n);
    for (size_t i = 0; i < n; i++) frames[i] = f2(float(out.data1[2 * i]) / 255.0f * 63.0f, float(out.data1[2 * i + 1]) / 255.0f * 63.0f);
    { SpatialOutputs o{ &t2d, &t2s, nullptr }; spatialPass(c, in, tiles, BLUR, t1d, t1s, &frames, o); keep(t2d, t2s); }
    for (size_t i = 0; i < n; i++) h.prevViewZ[i] = in.viewZ[i];                     // Blur copies viewZ (sky included)
    { SpatialOutputs o{ &h.diff, &h.spec, nullptr }; Image4 keepD = h.diff, keepS = h.spec; spatialPass(c, in, tiles, POST_BLUR, t2d, t2s, &frames, o); (void)keepD; (void)keepS; }
    for (size_t i = 0; i < n; i++) h.prevNormalRoughness[i] = in.normalRoughness[i];  // PostBlur copies the packed normal / roughness (written for non-sky pixels; sky texels are rejected by viewZ)
    keep(h.diff, h.spec);
    out.diff = h.diff; out.spec = h.spec;
    std::vector<float> lumaD = h.diffLumaStabilized, lumaS = h.specLumaStabilized;
    temporalStabilization(c, m, tp, StabilizationParams(), in, tiles, out.data1, out.data2, h.diff, h.spec, trackingPong, h, out.diff, out.spec, lumaD, lumaS, h.prevInternalData);
    h.diffLumaStabilized.swap(lumaD); h.specLumaStabilized.swap(lumaS); h.specHitDistForTracking.swap(trackingPong);
    h.valid = true;
}

} } // namespace orc::reblur
