// ORACLE — test infrastructure only (see pt_math.h).
// pt_tonemap.h: tone mapping / auto exposure for presentable output, restated from
//   Rtxpt/ToneMapper/luminance_ps.hlsl:14-29 (log2 luminance), ToneMapping.hlsl:24-32 (capture of the last MIP = the mean), ToneMapping.ps.hlsli:31-171 (operators, exposure,
//   colour transform, clamp), ToneMapping_cb.h:14-43, ToneMappingPasses.cpp:316-347 (constants), :393-441 (white balance, exposure value, colour transform), :443-459
//   (GetPreExposedGray), ColorUtils.h:44-204 (Rec.709 / CAT02 matrices as the file stores them, Kang et al. colour temperature, von Kries white balance)
// Differences kept on purpose: the reference averages log-luminance through a MIP chain and uses the value captured a frame earlier (CPU read-back); here the mean is over
// all pixels of the frame being mapped.  The LDR target is SRGBA8: the sRGB encode the ROP applies is written out.
#pragma once
#include "pt_math.h"
#include <cmath>

namespace orc { namespace tonemap {

struct Params       // ToneMappingParameters (ToneMappingPasses.h) + what RTXPT's UI sets
{
    uint op = 5;                      // 0 Linear, 1 Reinhard, 2 ReinhardModified, 3 HejiHableAlu, 4 HableUc2, 5 Aces
    uint clamped = 1, autoExposure = 0, enabled = 1, whiteBalance = 0;
    float exposureCompensation = 0, exposureValueMin = -16, exposureValueMax = 16, whiteScale = 11.2f, whiteMaxLuminance = 1.0f, whitePoint = 6500.0f;
    float filmSpeed = 100.0f, fNumber = 1.0f, shutter = 1.0f;
};
struct M3 { float m[3][3]; };
inline M3 mul(const M3& a, const M3& b) { M3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { float s = 0; for (int k = 0; k < 3; k++) s += a.m[i][k] * b.m[k][j]; r.m[i][j] = s; } return r; }
inline float3 mul(const M3& a, float3 v) { const float c[3] = { v.x, v.y, v.z }; float r[3]; for (int i = 0; i < 3; i++) { float s = 0; for (int j = 0; j < 3; j++) s += a.m[i][j] * c[j]; r[i] = s; } return f3(r[0], r[1], r[2]); }
// ColorUtils.h's arrays read the way donut's row-major float3x3 reads them
static const M3 kRGBtoXYZ = { { { 0.4123907992659595f, 0.2126390058715104f, 0.0193308187155918f }, { 0.3575843393838780f, 0.7151686787677559f, 0.1191947797946259f }, { 0.1804807884018343f, 0.0721923153607337f, 0.9505321522496608f } } };
static const M3 kXYZtoRGB = { { { 3.2409699419045213f, -0.9692436362808798f, 0.0556300796969936f }, { -1.5373831775700935f, 1.8759675015077206f, -0.2039769588889765f }, { -0.4986107602930033f, 0.0415550574071756f, 1.0569715142428784f } } };
static const M3 kXYZtoLMS = { { { 0.7328f, -0.7036f, 0.0030f }, { 0.4296f, 1.6975f, 0.0136f }, { -0.1624f, 0.0061f, 0.9834f } } };
static const M3 kLMStoXYZ = { { { 1.096123820835514f, 0.454369041975359f, -0.009627608738429f }, { -0.278869000218287f, 0.473533154307412f, -0.005698031216113f }, { 0.182745179382773f, 0.072097803717229f, 1.015325639954543f } } };
inline float3 colorTemperatureToXYZ(float T)
{
    if (T < 1667.f || T > 25000.f) return f3(0);
    const double t = T, t2 = t * t, t3 = t * t * t;
    const double xc = T < 4000.f ? -0.2661239e9 / t3 - 0.2343580e6 / t2 + 0.8776956e3 / t + 0.179910 : -3.0258469e9 / t3 + 2.1070379e6 / t2 + 0.2226347e3 / t + 0.240390;
    const double x = xc, x2 = x * x, x3 = x * x * x;
    const double yc = T < 2222.f ? -1.1063814 * x3 - 1.34811020 * x2 + 2.18555832 * x - 0.20219683 : (T < 4000.f ? -0.9549476 * x3 - 1.37418593 * x2 + 2.09137015 * x - 0.16748867 : 3.0817580 * x3 - 5.87338670 * x2 + 3.75112997 * x - 0.37001483);
    const float fx = float(xc), fy = float(yc);
    return f3(fx * 1.0f / fy, 1.0f, (1.f - fx - fy) * 1.0f / fy);
}
inline M3 whiteBalanceTransform(float T)
{
    const M3 MA = mul(kXYZtoLMS, kRGBtoXYZ), invMA = mul(kXYZtoRGB, kLMStoXYZ);
    const float3 wd = mul(kXYZtoLMS, colorTemperatureToXYZ(6500.f)), ws = mul(kXYZtoLMS, colorTemperatureToXYZ(T));
    M3 D = { { { wd.x / ws.x, 0, 0 }, { 0, wd.y / ws.y, 0 }, { 0, 0, wd.z / ws.z } } };
    return mul(mul(invMA, D), MA);
}
// m_ColorTransform = whiteBalance * 2^exposureCompensation * manual exposure (only without auto exposure)
inline M3 colorTransform(const Params& p)
{
    M3 wb = { { { 1, 0, 0 }, { 0, 1, 0 }, { 0, 0, 1 } } };
    if (p.whiteBalance) wb = whiteBalanceTransform(p.whitePoint);
    const float exposureScale = powf(2.f, p.exposureCompensation);
    float manual = 1.f;
    if (!p.autoExposure) manual = ((1.f / 100.f) * p.filmSpeed) / (p.shutter * p.fNumber * p.fNumber);
    M3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = wb.m[i][j] * exposureScale * manual;
    return r;
}
inline float luminance(float3 c) { return dot(c, f3(0.299f, 0.587f, 0.114f)); }
// exp2 of the mean of log2( max( 1e-4, luminance ) ): what the luminance pass + MIP chain + capture hand the CPU
inline float averageLuminance(const float* rgba, size_t pixelCount)
{
    double s = 0; for (size_t i = 0; i < pixelCount; i++) s += double(log2f(std::max(0.0001f, luminance(f3(rgba[4 * i], rgba[4 * i + 1], rgba[4 * i + 2])))));
    return exp2f(float(s / double(pixelCount)));
}
inline float3 uc2(float3 c) { const float A = 0.22f, B = 0.3f, C = 0.1f, D = 0.2f, E = 0.01f, F = 0.3f; auto f = [&](float x) { return ((x * (A * x + C * B) + D * E) / (x * (A * x + B) + D * F)) - (E / F); }; return f3(f(c.x), f(c.y), f(c.z)); }
inline float3 toneMapOp(const Params& p, float3 c)
{
    switch (p.op)
    {
    case 1: { const float l = luminance(c), r = l / (l + 1); return c * (r / l); }
    case 2: { const float l = luminance(c), r = l * (1 + l / (p.whiteMaxLuminance * p.whiteMaxLuminance)) * (1 + l); return c * (r / l); }
    case 3: { auto f = [](float x) { x = std::max(0.0f, x - 0.004f); x = (x * (6.2f * x + 0.5f)) / (x * (6.2f * x + 1.7f) + 0.06f); return powf(x, 2.2f); }; return f3(f(c.x), f(c.y), f(c.z)); }
    case 4: { const float3 v = uc2(c * 2.0f); const float ws = 1 / uc2(f3(p.whiteScale)).x; return v * ws; }
    case 5: { auto f = [](float x) { x *= 0.6f; return saturate((x * (2.51f * x + 0.03f)) / (x * (2.43f * x + 0.59f) + 0.14f)); }; return f3(f(c.x), f(c.y), f(c.z)); }
    default: return c;
    }
}
inline float3 apply(const Params& p, const M3& ct, float avgLuminance, float3 c)
{
    if (p.autoExposure) c = c * std::min(std::max(0.042f / avgLuminance, exp2f(p.exposureValueMin)), exp2f(p.exposureValueMax));
    if (p.enabled) { c = mul(ct, c); c = toneMapOp(p, c); if (p.clamped) c = f3(saturate(c.x), saturate(c.y), saturate(c.z)); }
    return c;
}
inline uint8_t srgb8(float v)
{
    v = saturate(v);
    const float e = v <= 0.0031308f ? v * 12.92f : 1.055f * powf(v, 1.0f / 2.4f) - 0.055f;
    return uint8_t(e * 255.0f + 0.5f);
}
inline float3 preExposedGray(const Params& p, float avgLuminance)
{   // inverse( m_ColorTransform ) * 0.18, divided by the auto-exposure factor
    const M3 m = colorTransform(p);
    const double a = m.m[0][0], b = m.m[0][1], c = m.m[0][2], d = m.m[1][0], e = m.m[1][1], f = m.m[1][2], g = m.m[2][0], h = m.m[2][1], i = m.m[2][2];
    const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
    const double inv[3][3] = { { (e * i - f * h) / det, (c * h - b * i) / det, (b * f - c * e) / det }, { (f * g - d * i) / det, (a * i - c * g) / det, (c * d - a * f) / det }, { (d * h - e * g) / det, (b * g - a * h) / det, (a * e - b * d) / det } };
    float3 r = f3(float((inv[0][0] + inv[0][1] + inv[0][2]) * 0.18), float((inv[1][0] + inv[1][1] + inv[1][2]) * 0.18), float((inv[2][0] + inv[2][1] + inv[2][2]) * 0.18));
    if (p.autoExposure) r = r / (0.042f / avgLuminance);
    return r;
}

} } // namespace orc::tonemap
