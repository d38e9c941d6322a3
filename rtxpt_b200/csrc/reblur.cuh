// reblur.cuh — parameter block and device helpers of the ReBLUR kernels (reblur_kernels.cu): NRD's REBLUR_DIFFUSE_SPECULAR denoiser (NRD 4.15.2 as vendored with RTXPT under
// External/Nrd) for one stable plane, SURVEY §8 row a18 / K9.  Citations: External/Nrd/Shaders/Include/{Common,REBLUR_Common,REBLUR_Config,NRD}.hlsli for the helpers,
// External/Nrd/Source/Reblur.cpp:280-392 and InstanceImpl.cpp:331-451 for the constants.  NRD's shaders also call NVIDIA MathLib (ml.hlsli), which is not vendored with the
// reference; those functions are written from their published definitions (Math::SmoothStep / LinearStep / AcosApprox, Geometry::GetRotator / RotateVector / GetBasis /
// ReconstructViewPosition, ImportanceSampling::GetSpecularLobeTanHalfAngle, Filtering::GetBilinearFilter ...).
#pragma once
#include "device_math.cuh"

namespace pt { namespace rb {

constexpr float kEps = 1e-6f, kInf = 1e6f;
constexpr float kNormalEncodingError = 0.75f / 255.0f;         // NRD_NORMAL_ENCODING_ERROR for R10G10B10A2 normals
constexpr float kAlmostZeroAngle = 0.01745240643728351f;       // cos( 89 deg )

struct Rotator { float x, y, z, w; };                          // ( cos, sin, -sin, cos )

struct Params
{
    // image
    uint W, H, tilesW, frameIndex;
    // camera (camera-relative world space: the current camera is the origin, as NRD makes its matrices, InstanceImpl.cpp:404-415)
    float viewToWorld[9];           // rows = view axes in world space
    float viewToClip[16], worldToClip[16], worldToClipPrev[16];
    float worldToViewPrev[12];      // 3x4: Xv = R * X + t
    float frustum[4], frustumPrev[4];
    float cameraDelta[3];
    float unproject, minRectDimMulUnproject;
    Rotator rotatorPre, rotator, rotatorPost;
    // settings (nrd::ReblurSettings + the common settings RTXPT passes)
    float hitDistParams[4];
    float maxAccumulatedFrameNum, maxFastAccumulatedFrameNum, historyFixFrameNum, historyFixBasePixelStride;
    float diffPrepassBlurRadius, specPrepassBlurRadius, minBlurRadius, maxBlurRadius, lobeAngleFraction /* squared */, roughnessFraction, planeDistSensitivity, minHitDistanceWeight;
    float minMaterialDiff, minMaterialSpec, denoisingRange, viewZScale;
    float disocclusionThreshold, disocclusionThresholdAlternate, framerateScale, fireflySuppressorMinRelativeScale, responsiveAccumulationRoughnessThreshold;
    float antilagSigmaScale, antilagSensitivity, stabilizationStrength;
    uint antiFirefly, usePrepassOnlyForSpecularMotionEstimation;
    // inputs (written by k_dn_prepare_inputs)
    const float* viewZ; const uint* normalRoughness; const uint2* motion; const unsigned char* disocclusionMix; const uint2* inDiff; const uint2* inSpec;
    // transient pool
    unsigned char* tiles; uint2* tmp1Diff; uint2* tmp1Spec; uint2* tmp2Diff; uint2* tmp2Spec; unsigned short* trackingTransient; unsigned short* diffFastTransient; unsigned short* specFastTransient;
    uchar2* data1; uint* data2;
    // permanent pool of this plane's instance
    float* prevViewZ; uint* prevNormalRoughness; unsigned short* prevInternalData; uint2* diffHistory; uint2* specHistory; unsigned short* diffFast; unsigned short* specFast;
    unsigned short* trackingPrev; unsigned short* trackingCurr; unsigned short* diffLumaPrev; unsigned short* diffLumaCurr; unsigned short* specLumaPrev; unsigned short* specLumaCurr;
    // outputs
    uint2* outDiff; uint2* outSpec;
};

// ---- MathLib --------------------------------------------------------------------------------------------------------------------------------------------
PT_HD float linearStep(float a, float b, float x) { return sat((x - a) / (b - a)); }
PT_HD float smoothStep01(float x) { x = sat(x); return x * x * (3.0f - 2.0f * x); }
PT_HD float smoothStep(float a, float b, float x) { return smoothStep01(linearStep(a, b, x)); }
PT_HD float pow01(float x, float y) { return powf(sat(x), y); }
PT_HD float sqrt01(float x) { return sqrtf(sat(x)); }
PT_HD float positiveRcp(float x) { return 1.0f / fmaxf(x, 1.175494351e-38f); }
PT_HD float acosApprox(float x) { return 1.41421356237f * sqrtf(sat(1.0f - x)); }
PT_HD float pow5(float x) { return powf(sat(1.0f - x), 5.0f); }
PT_HD float2 rotate(Rotator r, float2 v) { return mk2(v.x * r.x + v.y * r.y, v.x * r.z + v.y * r.w); }
PT_HD Rotator scaleRotator(Rotator r, float2 s) { Rotator o = { r.x * s.x, r.y * s.x, r.z * s.y, r.w * s.y }; return o; }
PT_HD void getBasis(float3 N, float3& T, float3& B)
{
    const float sz = N.z >= 0.0f ? 1.0f : -1.0f, a = 1.0f / (sz + N.z), ya = N.y * a, b = N.x * ya, c = N.x * sz;
    T = mk3(c * N.x * a - 1.0f, sz * b, c); B = mk3(b, N.y * ya - sz, N.y);
}
struct Rng { uint s; PT_HD void init(uint x, uint y, uint f) { s = hash32Combine(hash32Combine(hash32(x), y), f); } PT_HD float next() { s = hash32(s); return hashToFloat(s); } };
PT_HD float specularDominantFactor(float NoV, float roughness)          // NRD.hlsli:392-398
{
    const float a = 0.298475f * logf(39.4115f - 39.0029f * roughness);
    return sat(powf(sat(1.0f - NoV), 10.8649f) * (1.0f - a) + a);
}
PT_HD float4 specularDominantDirection(float3 N, float3 V, float roughness)
{
    const float f = specularDominantFactor(fabsf(dot3(N, V)), roughness);
    const float3 R = N * (2.0f * dot3(N, V)) - V;
    const float3 D = norm3(lerp3(N, R, f));
    return make_float4(D.x, D.y, D.z, f);
}
PT_HD float specularLobeTanHalfAngle(float roughness, float percentOfVolume)
{
    roughness = sat(roughness); percentOfVolume = sat(percentOfVolume);
    return roughness * roughness * sqrtf(percentOfVolume / (1.0f - percentOfVolume + kEps));
}

// ---- images --------------------------------------------------------------------------------------------------------------------------------------------------
PT_HD float4 unpackRGBA16F(uint2 v) { return make_float4(f16tof32(v.x), f16tof32(v.x >> 16), f16tof32(v.y), f16tof32(v.y >> 16)); }
PT_HD uint2 packRGBA16F(float4 v) { return make_uint2(f32tof16(v.x) | (f32tof16(v.y) << 16), f32tof16(v.z) | (f32tof16(v.w) << 16)); }
using pt::operator+; using pt::operator-; using pt::operator*; using pt::operator/;       // the float4 overloads below would otherwise hide pt's float3 operators
PT_HD float4 operator+(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
PT_HD float4 operator*(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
PT_HD int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
PT_HD float ldHalf(const unsigned short* p, size_t i) { return __half2float(__ushort_as_half(p[i])); }
PT_HD void stHalf(unsigned short* p, size_t i, float v) { p[i] = __half_as_ushort(__float2half_rn(v)); }
// NRD_FrontEnd_UnpackNormalAndRoughness for R10G10B10A2_UNORM + linear roughness; .w = roughness, materialID out
PT_HD float4 unpackNormalRoughness(uint p, float& materialID)
{
    const float px = float(p & 1023u) / 1023.0f * 2.0f - 1.0f, py = float((p >> 10) & 1023u) / 1023.0f * 2.0f - 1.0f;
    float3 n = mk3(px, py, 1.0f - fabsf(px) - fabsf(py));
    const float t = sat(-n.z);
    n.x -= t * (n.x >= 0.0f ? 1.0f : -1.0f); n.y -= t * (n.y >= 0.0f ? 1.0f : -1.0f);
    n = n * (1.0f / sqrtf(dot3(n, n) + 1e-9f));
    materialID = float(p >> 30);
    return make_float4(n.x, n.y, n.z, float((p >> 20) & 1023u) / 1023.0f);
}
PT_HD float3 xyz(float4 v) { return mk3(v.x, v.y, v.z); }

// ---- Common.hlsli / REBLUR_Common.hlsli ----------------------------------------------------------------------------------------------------------------------
PT_HD float hitDistanceNormalization(const Params& p, float viewZ, float roughness) { return (p.hitDistParams[0] + fabsf(viewZ) * p.hitDistParams[1]) * lerpf(1.0f, p.hitDistParams[2], sat(exp2f(p.hitDistParams[3] * roughness * roughness))); }
PT_HD float3 reconstructViewPosition(const float* frustum, float2 uv, float viewZ) { return mk3((uv.x * frustum[2] + frustum[0]) * viewZ, (uv.y * frustum[3] + frustum[1]) * viewZ, viewZ); }
PT_HD float3 worldToViewRotate(const Params& p, float3 n) { return mk3(dot3(mk3(p.viewToWorld[0], p.viewToWorld[1], p.viewToWorld[2]), n), dot3(mk3(p.viewToWorld[3], p.viewToWorld[4], p.viewToWorld[5]), n), dot3(mk3(p.viewToWorld[6], p.viewToWorld[7], p.viewToWorld[8]), n)); }
PT_HD float3 viewToWorldRotate(const Params& p, float3 v) { return mk3(p.viewToWorld[0], p.viewToWorld[1], p.viewToWorld[2]) * v.x + mk3(p.viewToWorld[3], p.viewToWorld[4], p.viewToWorld[5]) * v.y + mk3(p.viewToWorld[6], p.viewToWorld[7], p.viewToWorld[8]) * v.z; }
PT_HD float frustumSize(const Params& p, float viewZ) { return p.minRectDimMulUnproject * viewZ; }
PT_HD float pixelRadiusToWorld(const Params& p, float pixelRadius, float viewZ) { return pixelRadius * p.unproject * viewZ; }
PT_HD float specMagicCurve(float roughness) { return (1.0f - exp2f(-200.0f * roughness * roughness)) * pow01(roughness, 0.25f); }
PT_HD float expApprox(float x) { return 1.0f / (x * x - x + 1.0f); }
PT_HD float exponentialWeight(float x, float px, float py) { return expApprox(-3.0f * fabsf(x * px + py)); }
PT_HD float weight(float x, float px, float py) { return smoothStep(1.0f, 0.0f, fabsf(x * px + py)); }
PT_HD float weightWithSigma(float x, float px, float py, float sigma) { return smoothStep(1.0f, 0.0f, fabsf(x * px + py) - sigma * px); }
PT_HD float gaussianWeight(float r) { return expf(-0.66f * r * r); }
PT_HD float normalWeightParam(float nonLinearAccumSpeed, float lobeAngleFraction, float roughness) { return 1.0f / fmaxf(atanf(specularLobeTanHalfAngle(roughness, 0.75f * lerpf(lobeAngleFraction, 1.0f, nonLinearAccumSpeed))), kNormalEncodingError); }
PT_HD float2 geometryWeightParams(const Params& p, float fs, float3 Xv, float3 Nv) { const float a = 1.0f / (p.planeDistSensitivity * fs); return mk2(a, -dot3(Nv, Xv) * a); }
PT_HD float2 hitDistanceWeightParams(float hitDist, float nonLinearAccumSpeed, float roughness) { const float a = 1.0f / lerpf(0.0005f, 1.0f, fminf(nonLinearAccumSpeed, specMagicCurve(roughness))); return mk2(a, -hitDist * a); }
PT_HD float2 roughnessWeightParams(float roughness, float fraction) { const float a = 1.0f / lerpf(0.01f, 1.0f, sat(roughness * fraction)); return mk2(a, -roughness * a); }
PT_HD float2 relaxedRoughnessWeightParams(float m, float fraction, float sensitivity) { const float a = 1.0f / lerpf(sensitivity, 1.0f, lerpf(m * m, m, fraction)); return mk2(a, -m * a); }
PT_HD float fadeBasedOnAccumulatedFrames(const Params& p, float accumSpeed) { return linearStep(p.historyFixFrameNum * 2.0f / 3.0f + 1e-6f, p.historyFixFrameNum * 4.0f / 3.0f + 2e-6f, accumSpeed); }
PT_HD bool compareMaterials(float m0, float m, float minm) { return fmaxf(m0, minm) == fmaxf(m, minm); }
PT_HD float2 screenUv(const float* M, float3 X)
{
    const float cx = X.x * M[0] + X.y * M[4] + X.z * M[8] + M[12], cy = X.x * M[1] + X.y * M[5] + X.z * M[9] + M[13], cw = X.x * M[3] + X.y * M[7] + X.z * M[11] + M[15];
    if (cw < 0.0f) return mk2(99999.0f, 99999.0f);
    return mk2(cx / cw * 0.5f + 0.5f, -(cy / cw) * 0.5f + 0.5f);
}
PT_HD float4 changeLuma(float4 v, float newLuma) { const float s = (newLuma + kEps) / (v.x + kEps); return make_float4(v.x * s, v.y * s, v.z * s, v.w); }
PT_HD float4 clampNegativeToZero(float4 v)
{
    const float t = v.x - v.z; const float3 rgb = mk3(fmaxf(t + v.y, 0.0f), fmaxf(v.x + v.z, 0.0f), fmaxf(t - v.y, 0.0f));
    return make_float4(dot3(rgb, mk3(0.25f, 0.5f, 0.25f)), dot3(rgb, mk3(0.5f, 0.0f, -0.5f)), dot3(rgb, mk3(-0.25f, 0.5f, -0.25f)), sat(v.w));
}
PT_HD float4 mixHistoryAndCurrent(const Params& p, float4 history, float4 current, float f, float roughness)
{
    const float fw = fmaxf(f, 1.0f / (1.0f + 0.5f * specMagicCurve(roughness) * p.maxAccumulatedFrameNum));
    return make_float4(lerpf(history.x, current.x, f), lerpf(history.y, current.y, f), lerpf(history.z, current.z, f), lerpf(history.w, current.w, fw));
}
PT_HD uint packInternalData(float diffAccumSpeed, float specAccumSpeed, float materialID)
{
    return uint(sat(diffAccumSpeed / 63.0f) * 63.0f + 0.5f) | (uint(sat(specAccumSpeed / 63.0f) * 63.0f + 0.5f) << 6) | (uint(sat(materialID / 15.0f) * 15.0f + 0.5f) << 12);
}
PT_HD float3 unpackInternalData(uint v) { return mk3(float(v & 63u), float((v >> 6) & 63u), float((v >> 12) & 15u)); }
PT_HD float3 xVirtual(float hitDist, float curvature, float3 X, float3 Xprev, float3 N, float3 V, float roughness)      // Common.hlsli:405-453 (NRD_USE_SPECULAR_MOTION_V2)
{
    const float4 D = specularDominantDirection(N, V, roughness);
    const float3 ray = xyz(D) * hitDist;
    float3 T, B; getBasis(N, T, B);
    float3 O = mk3(dot3(T, ray), dot3(B, ray), -dot3(N, ray));
    float mag = 1.0f / (2.0f * curvature * O.z - 1.0f);
    const float f = len3(X) * (1.0f - fabsf(dot3(N, V))) * fmaxf(curvature, 0.0f);
    mag *= 1.0f / (1.0f + f);
    const float3 Iw = V * len3(O * mag);
    const float closeness = sat(len3(Iw) / (hitDist + kEps));
    return lerp3(Xprev, X, closeness * D.w) - Iw * D.w;
}
struct Bilinear { float2 origin, weights; };
PT_HD Bilinear bilinearFilter(float2 uv, float W, float H) { const float tx = uv.x * W - 0.5f, ty = uv.y * H - 0.5f; Bilinear b; b.origin = mk2(floorf(tx), floorf(ty)); b.weights = mk2(tx - b.origin.x, ty - b.origin.y); return b; }
PT_HD float applyBilinear(float s00, float s10, float s01, float s11, const Bilinear& f) { return lerpf(lerpf(s00, s10, f.weights.x), lerpf(s01, s11, f.weights.x), f.weights.y); }
PT_HD float4 bilinearCustomWeights(const Bilinear& f, float4 cw) { const float ox = 1.0f - f.weights.x, oy = 1.0f - f.weights.y; return make_float4(cw.x * ox * oy, cw.y * f.weights.x * oy, cw.z * ox * f.weights.y, cw.w * f.weights.x * f.weights.y); }
PT_HD float applyCustomWeights(float s00, float s10, float s01, float s11, float4 w) { const float sum = w.x + w.y + w.z + w.w; return sum < 0.0001f ? 0.0f : (s00 * w.x + s10 * w.y + s01 * w.z + s11 * w.w) / sum; }
// SampleLevel( gLinearClamp ) of an RGBA16F / R16F image, position in texels
PT_HD float4 sampleBilinear4(const uint2* img, int W, int H, float2 pos)
{
    const float tx = pos.x - 0.5f, ty = pos.y - 0.5f, fx = floorf(tx), fy = floorf(ty), wx = tx - fx, wy = ty - fy;
    const int x0 = clampi(int(fx), 0, W - 1), x1 = clampi(int(fx) + 1, 0, W - 1), y0 = clampi(int(fy), 0, H - 1), y1 = clampi(int(fy) + 1, 0, H - 1);
    const float4 a = unpackRGBA16F(img[size_t(y0) * W + x0]), b = unpackRGBA16F(img[size_t(y0) * W + x1]), c = unpackRGBA16F(img[size_t(y1) * W + x0]), d = unpackRGBA16F(img[size_t(y1) * W + x1]);
    return (a * (1 - wx) + b * wx) * (1 - wy) + (c * (1 - wx) + d * wx) * wy;
}
PT_HD float sampleBilinear1(const unsigned short* img, int W, int H, float2 pos)
{
    const float tx = pos.x - 0.5f, ty = pos.y - 0.5f, fx = floorf(tx), fy = floorf(ty), wx = tx - fx, wy = ty - fy;
    const int x0 = clampi(int(fx), 0, W - 1), x1 = clampi(int(fx) + 1, 0, W - 1), y0 = clampi(int(fy), 0, H - 1), y1 = clampi(int(fy) + 1, 0, H - 1);
    return lerpf(lerpf(ldHalf(img, size_t(y0) * W + x0), ldHalf(img, size_t(y0) * W + x1), wx), lerpf(ldHalf(img, size_t(y1) * W + x0), ldHalf(img, size_t(y1) * W + x1), wx), wy);
}
// Catmull-Rom (12 taps through 5 bilinear fetches) with fallback to the 2x2 footprint and custom weights (Common.hlsli:610-665)
struct CatRom { float2 centerPos, tc, f; float4 w; float w4, sum; bool bicubic; };
PT_HD CatRom catRomSetup(float2 samplePos, float4 customWeights, bool useBicubic)
{
    const float S = 0.5f; CatRom c; c.bicubic = useBicubic;
    c.centerPos = mk2(floorf(samplePos.x - 0.5f) + 0.5f, floorf(samplePos.y - 0.5f) + 0.5f); c.f = mk2(sat(samplePos.x - c.centerPos.x), sat(samplePos.y - c.centerPos.y));
    const float fx = c.f.x, fy = c.f.y;
    const float w0x = fx * (fx * (-S * fx + 2.0f * S) - S), w1x = fx * (fx * ((2.0f - S) * fx - (3.0f - S))) + 1.0f, w2x = fx * (fx * (-(2.0f - S) * fx + (3.0f - 2.0f * S)) + S), w3x = fx * (fx * (S * fx - S));
    const float w0y = fy * (fy * (-S * fy + 2.0f * S) - S), w1y = fy * (fy * ((2.0f - S) * fy - (3.0f - S))) + 1.0f, w2y = fy * (fy * (-(2.0f - S) * fy + (3.0f - 2.0f * S)) + S), w3y = fy * (fy * (S * fy - S));
    const float w12x = w1x + w2x, w12y = w1y + w2y; c.tc = mk2(w2x / w12x, w2y / w12y);
    c.w = make_float4(w12x * w0y, w0x * w12y, w12x * w12y, w3x * w12y); c.w4 = w12x * w3y;
    if (!useBicubic) { c.w = customWeights; c.w4 = 0.0f; }
    c.sum = c.w.x + c.w.y + c.w.z + c.w.w + c.w4;
    return c;
}
PT_HD float4 catRomSample4(const CatRom& c, const uint2* img, int W, int H)
{
    float4 color;
    if (c.bicubic)
        color = sampleBilinear4(img, W, H, mk2(c.centerPos.x + c.tc.x, c.centerPos.y - 1.0f)) * c.w.x + sampleBilinear4(img, W, H, mk2(c.centerPos.x - 1.0f, c.centerPos.y + c.tc.y)) * c.w.y + sampleBilinear4(img, W, H, mk2(c.centerPos.x + c.tc.x, c.centerPos.y + c.tc.y)) * c.w.z
              + sampleBilinear4(img, W, H, mk2(c.centerPos.x + 2.0f, c.centerPos.y + c.tc.y)) * c.w.w + sampleBilinear4(img, W, H, mk2(c.centerPos.x + c.tc.x, c.centerPos.y + 2.0f)) * c.w4;
    else
        color = sampleBilinear4(img, W, H, c.centerPos) * c.w.x + sampleBilinear4(img, W, H, mk2(c.centerPos.x + 1, c.centerPos.y)) * c.w.y + sampleBilinear4(img, W, H, mk2(c.centerPos.x, c.centerPos.y + 1)) * c.w.z + sampleBilinear4(img, W, H, mk2(c.centerPos.x + 1, c.centerPos.y + 1)) * c.w.w;
    return c.sum < 0.0001f ? make_float4(0, 0, 0, 0) : color * (1.0f / c.sum);
}
PT_HD float catRomSample1(const CatRom& c, const unsigned short* img, int W, int H)
{
    float color;
    if (c.bicubic)
        color = sampleBilinear1(img, W, H, mk2(c.centerPos.x + c.tc.x, c.centerPos.y - 1.0f)) * c.w.x + sampleBilinear1(img, W, H, mk2(c.centerPos.x - 1.0f, c.centerPos.y + c.tc.y)) * c.w.y + sampleBilinear1(img, W, H, mk2(c.centerPos.x + c.tc.x, c.centerPos.y + c.tc.y)) * c.w.z
              + sampleBilinear1(img, W, H, mk2(c.centerPos.x + 2.0f, c.centerPos.y + c.tc.y)) * c.w.w + sampleBilinear1(img, W, H, mk2(c.centerPos.x + c.tc.x, c.centerPos.y + 2.0f)) * c.w4;
    else
        color = sampleBilinear1(img, W, H, c.centerPos) * c.w.x + sampleBilinear1(img, W, H, mk2(c.centerPos.x + 1, c.centerPos.y)) * c.w.y + sampleBilinear1(img, W, H, mk2(c.centerPos.x, c.centerPos.y + 1)) * c.w.z + sampleBilinear1(img, W, H, mk2(c.centerPos.x + 1, c.centerPos.y + 1)) * c.w.w;
    return c.sum < 0.0001f ? 0.0f : color / c.sum;
}
// the 2x2 footprint of an R16F image with custom weights (fast history)
PT_HD float footprintSample1(float2 centerPos, float4 w, const unsigned short* img, int W, int H)
{
    const int ox = int(centerPos.x), oy = int(centerPos.y);
    const float sum = w.x + w.y + w.z + w.w;
    const float v = ldHalf(img, size_t(clampi(oy, 0, H - 1)) * W + clampi(ox, 0, W - 1)) * w.x + ldHalf(img, size_t(clampi(oy, 0, H - 1)) * W + clampi(ox + 1, 0, W - 1)) * w.y
                  + ldHalf(img, size_t(clampi(oy + 1, 0, H - 1)) * W + clampi(ox, 0, W - 1)) * w.z + ldHalf(img, size_t(clampi(oy + 1, 0, H - 1)) * W + clampi(ox + 1, 0, W - 1)) * w.w;
    return sum < 0.0001f ? 0.0f : v / sum;
}

} } // namespace pt::rb
