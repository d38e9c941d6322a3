// tonemap.cuh - tone mapping / auto exposure for presentable output (replaces ToneMappingPass::Render, Rtxpt/ToneMapper/ToneMappingPasses.cpp:230-360 with
// luminance_ps.hlsl and ToneMapping.ps.hlsli): host-side derivation of the constants (white balance by von Kries scaling in CAT02 space for a colour temperature, exposure
// compensation, manual exposure: ToneMappingPasses.cpp:393-441, ColorUtils.h:44-204 with its matrices read the way donut's row-major float3x3 reads them) and the per-pixel bodies
// as __host__ __device__ functions (kernels: tonemap_kernels.cu; host build: tests/emu).  The reference averages log-luminance through a MIP chain and maps with the value a CPU
// read-back captured a frame earlier; here the mean is over all pixels of the frame being mapped.  The SRGBA8 target's sRGB encode is written out.
#pragma once
#include "device_math.cuh"
#include "../../include/rtxpt_b200.h"
#include <math.h>

namespace pt { namespace tonemap {

struct Params       // ToneMappingConstants (ToneMapping_cb.h:28-43)
{
    uint op, clamped, autoExposure, enabled;
    float whiteScale, whiteMaxLuminance, autoExposureLumValueMin, autoExposureLumValueMax;
    float colorTransform[9];        // row-major, applied as M * c
};

// ---- host: RtxptToneMappingParams -> constants --------------------------------------------------------------------------------------------------------------------------
struct M3 { float m[3][3]; };
inline M3 mul3(const M3& a, const M3& b) { M3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { float s = 0; for (int k = 0; k < 3; k++) s += a.m[i][k] * b.m[k][j]; r.m[i][j] = s; } return r; }
inline void mul3v(const M3& a, const float* v, float* o) { for (int i = 0; i < 3; i++) { float s = 0; for (int j = 0; j < 3; j++) s += a.m[i][j] * v[j]; o[i] = s; } }
inline void colorTemperatureToXYZ(float T, float* xyz)
{   // Kang et al. 2002, piecewise rational polynomials, 1667 K .. 25000 K
    xyz[0] = xyz[1] = xyz[2] = 0;
    if (T < 1667.f || T > 25000.f) return;
    const double t = T, t2 = t * t, t3 = t * t * t;
    const double xc = T < 4000.f ? -0.2661239e9 / t3 - 0.2343580e6 / t2 + 0.8776956e3 / t + 0.179910 : -3.0258469e9 / t3 + 2.1070379e6 / t2 + 0.2226347e3 / t + 0.240390;
    const double x = xc, x2 = x * x, x3 = x * x * x;
    const double yc = T < 2222.f ? -1.1063814 * x3 - 1.34811020 * x2 + 2.18555832 * x - 0.20219683 : (T < 4000.f ? -0.9549476 * x3 - 1.37418593 * x2 + 2.09137015 * x - 0.16748867 : 3.0817580 * x3 - 5.87338670 * x2 + 3.75112997 * x - 0.37001483);
    const float fx = float(xc), fy = float(yc);
    xyz[0] = fx * 1.0f / fy; xyz[1] = 1.0f; xyz[2] = (1.f - fx - fy) * 1.0f / fy;
}
inline M3 whiteBalanceTransform(float T)
{
    static const M3 rgbToXyz = { { { 0.4123907992659595f, 0.2126390058715104f, 0.0193308187155918f }, { 0.3575843393838780f, 0.7151686787677559f, 0.1191947797946259f }, { 0.1804807884018343f, 0.0721923153607337f, 0.9505321522496608f } } };
    static const M3 xyzToRgb = { { { 3.2409699419045213f, -0.9692436362808798f, 0.0556300796969936f }, { -1.5373831775700935f, 1.8759675015077206f, -0.2039769588889765f }, { -0.4986107602930033f, 0.0415550574071756f, 1.0569715142428784f } } };
    static const M3 xyzToLms = { { { 0.7328f, -0.7036f, 0.0030f }, { 0.4296f, 1.6975f, 0.0136f }, { -0.1624f, 0.0061f, 0.9834f } } };
    static const M3 lmsToXyz = { { { 1.096123820835514f, 0.454369041975359f, -0.009627608738429f }, { -0.278869000218287f, 0.473533154307412f, -0.005698031216113f }, { 0.182745179382773f, 0.072097803717229f, 1.015325639954543f } } };
    const M3 MA = mul3(xyzToLms, rgbToXyz), invMA = mul3(xyzToRgb, lmsToXyz);
    float d65[3], src[3], wd[3], ws[3]; colorTemperatureToXYZ(6500.f, d65); colorTemperatureToXYZ(T, src); mul3v(xyzToLms, d65, wd); mul3v(xyzToLms, src, ws);
    const M3 D = { { { wd[0] / ws[0], 0, 0 }, { 0, wd[1] / ws[1], 0 }, { 0, 0, wd[2] / ws[2] } } };
    return mul3(mul3(invMA, D), MA);
}
inline M3 colorTransformOf(const RtxptToneMappingParams& u)
{
    M3 wb = { { { 1, 0, 0 }, { 0, 1, 0 }, { 0, 0, 1 } } };
    if (u.whiteBalance) wb = whiteBalanceTransform(u.whitePoint);
    const float exposureScale = powf(2.f, u.exposureCompensation);
    float manual = 1.f;
    if (!u.autoExposure) manual = ((1.f / 100.f) * u.filmSpeed) / (u.shutter * u.fNumber * u.fNumber);
    M3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = wb.m[i][j] * exposureScale * manual;
    return r;
}
inline Params makeParams(const RtxptToneMappingParams& u)
{
    Params p{};
    p.op = u.toneMapOperator; p.clamped = u.clamped; p.autoExposure = u.autoExposure; p.enabled = u.enabled; p.whiteScale = u.whiteScale; p.whiteMaxLuminance = u.whiteMaxLuminance;
    p.autoExposureLumValueMin = exp2f(u.autoExposure ? u.exposureValueMin : -16.0f); p.autoExposureLumValueMax = exp2f(u.autoExposure ? u.exposureValueMax : 16.0f);
    const M3 ct = colorTransformOf(u); for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) p.colorTransform[i * 3 + j] = ct.m[i][j];
    return p;
}
inline void preExposedGray(const RtxptToneMappingParams& u, float avgLuminance, float* out)
{   // ToneMappingPass::GetPreExposedGray: inverse( colour transform ) * 0.18, over the auto-exposure factor
    const M3 m = colorTransformOf(u);
    const double a = m.m[0][0], b = m.m[0][1], c = m.m[0][2], d = m.m[1][0], e = m.m[1][1], f = m.m[1][2], g = m.m[2][0], h = m.m[2][1], i = m.m[2][2];
    const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
    const double inv[3][3] = { { (e * i - f * h) / det, (c * h - b * i) / det, (b * f - c * e) / det }, { (f * g - d * i) / det, (a * i - c * g) / det, (c * d - a * f) / det }, { (d * h - e * g) / det, (b * g - a * h) / det, (a * e - b * d) / det } };
    for (int r = 0; r < 3; r++) { out[r] = float((inv[r][0] + inv[r][1] + inv[r][2]) * 0.18); if (u.autoExposure) out[r] = out[r] / (0.042f / avgLuminance); }
}

// ---- per-pixel bodies -----------------------------------------------------------------------------------------------------------------------------------------------------
PT_HD float tmLuminance(float3 c) { return dot3(c, mk3(0.299f, 0.587f, 0.114f)); }
PT_HD float logLuminance(float3 c) { return log2f(fmaxf(0.0001f, tmLuminance(c))); }                 // luminance_ps.hlsl
PT_HD float uc2f(float x) { const float A = 0.22f, B = 0.3f, C = 0.1f, D = 0.2f, E = 0.01f, F = 0.3f; return ((x * (A * x + C * B) + D * E) / (x * (A * x + B) + D * F)) - (E / F); }
PT_HD float hejlf(float x) { x = fmaxf(0.0f, x - 0.004f); x = (x * (6.2f * x + 0.5f)) / (x * (6.2f * x + 1.7f) + 0.06f); return powf(x, 2.2f); }
PT_HD float acesf(float x) { x *= 0.6f; return sat((x * (2.51f * x + 0.03f)) / (x * (2.43f * x + 0.59f) + 0.14f)); }
PT_HD float3 toneMapOp(const Params& p, float3 c)
{
    switch (p.op)
    {
    case 1: { const float l = tmLuminance(c), r = l / (l + 1); return c * (r / l); }
    case 2: { const float l = tmLuminance(c), r = l * (1 + l / (p.whiteMaxLuminance * p.whiteMaxLuminance)) * (1 + l); return c * (r / l); }
    case 3: return mk3(hejlf(c.x), hejlf(c.y), hejlf(c.z));
    case 4: { const float ws = 1 / uc2f(p.whiteScale); return mk3(uc2f(c.x * 2.0f), uc2f(c.y * 2.0f), uc2f(c.z * 2.0f)) * ws; }
    case 5: return mk3(acesf(c.x), acesf(c.y), acesf(c.z));
    default: return c;
    }
}
PT_HD float3 applyToneMapping(const Params& p, float avgLuminance, float3 c)
{
    if (p.autoExposure) c = c * clampf(0.042f / avgLuminance, p.autoExposureLumValueMin, p.autoExposureLumValueMax);         // TONEMAPPING_EXPOSURE_KEY
    if (p.enabled)
    {
        const float* m = p.colorTransform;
        c = mk3(m[0] * c.x + m[1] * c.y + m[2] * c.z, m[3] * c.x + m[4] * c.y + m[5] * c.z, m[6] * c.x + m[7] * c.y + m[8] * c.z);
        c = toneMapOp(p, c);
        if (p.clamped) c = mk3(sat(c.x), sat(c.y), sat(c.z));
    }
    return c;
}
PT_HD uint srgb8(float v) { v = sat(v); const float e = v <= 0.0031308f ? v * 12.92f : 1.055f * powf(v, 1.0f / 2.4f) - 0.055f; return uint(e * 255.0f + 0.5f); }
PT_HD uint packLdr(float3 c, float alpha) { return srgb8(c.x) | (srgb8(c.y) << 8) | (srgb8(c.z) << 16) | (uint(sat(alpha) * 255.0f + 0.5f) << 24); }

} } // namespace pt::tonemap
