// ORACLE — test infrastructure only (see pt_math.h).
// pt_envbake.h: the environment-map baking path: an equirectangular or cube source plus up to 16 directional lights baked into the cube the path
// tracer samples, and its MIP chain with solid-angle weights.  Restated from
//   Rtxpt/Lighting/Distant/EnvMapBaker.hlsl:64-99 (CubemapGetDirectionFor, SampleSource), :101-151 (texel solid angles), :153-180 (ComputeLightContribution),
//     :197-238 (GenerateTexel), :262-311 (BaseLayerCS: four texels + the first MIP), :313-356 (MIPReduceCS)
//   Rtxpt/Lighting/Distant/EnvMapBaker.cpp:146-159 (samplers: equirect = bilinear, wrap in u, clamp in v), :425-600 (Update: constants, dispatch order)
//   Rtxpt/Shaders/PathTracer/Utils/Math/MathHelpers.hlsli:92-99 (world_to_latlong_map)
// Not restated: the procedural sky (SampleProceduralSky.*), the BC6U compression of the baked cube (BC6UCompress.hlsl; RTXPT's default stores the cube block-compressed,
// this path keeps the RGBA16F cube, EnvMapBaker's compression quality 0), and the seamless cross-face filtering of a cube SOURCE (taps are clamped to the face).
#pragma once
#include "pt_math.h"
#include <vector>
#include <cmath>

namespace orc { namespace envbake {

struct DirectionalLight { float colorIntensity[4]; float direction[3]; float angularSize; };          // EMB_DirectionalLight: W/sr, incoming direction, radians
struct Desc
{
    uint cubeDim = 0;
    uint sourceType = 0;                      // 0 none, 1 equirectangular, 2 cube (6 faces of sourceWidth^2, +x -x +y -y +z -z)
    uint sourceWidth = 0, sourceHeight = 0; const float* source = nullptr;      // RGBA32F
    float scaleColor[3] = { 1, 1, 1 };
    uint directionalLightCount = 0; DirectionalLight lights[16];
};

inline float3 CubemapGetDirectionFor(uint face, float u, float v)
{
    const float cx = (u * 2.0f) - 1, cy = 1 - (v * 2.0f);
    const float l = sqrtf(cx * cx + cy * cy + 1);
    float3 d = f3(0);
    switch (face)
    {
    case 0: d = f3(1, cy, -cx); break;  case 1: d = f3(-1, cy, cx); break;
    case 2: d = f3(cx, 1, -cy); break;  case 3: d = f3(cx, -1, cy); break;
    case 4: d = f3(cx, cy, 1); break;   case 5: d = f3(-cx, cy, -1); break;
    }
    return d * (1 / l);
}
inline float4 div4(float4 v, float s) { return f4(v.x / s, v.y / s, v.z / s, v.w / s); }
inline float SphereQuadrantArea(float x, float y) { return atan2f(x * y, sqrtf(x * x + y * y + 1)); }
// solid angles of the 2x2 texels whose top-left is coordTL (CubemapTexelSolidAngle4): .x 00, .y 01, .z 10, .w 11 (first digit x)
inline float4 CubemapTexelSolidAngle4(float cubeDim, uint tlx, uint tly)
{
    const float iDim = 1.0f / cubeDim;
    const float s = ((float(tlx) + 0.5f) * 2 * iDim) - 1, t = ((float(tly) + 0.5f) * 2 * iDim) - 1;
    const float x0 = s - iDim, y0 = t - iDim, x1 = s + iDim, y1 = t + iDim, x2 = s + iDim * 3, y2 = t + iDim * 3;
    const float a00 = SphereQuadrantArea(x0, y0), a01 = SphereQuadrantArea(x0, y1), a10 = SphereQuadrantArea(x1, y0), a11 = SphereQuadrantArea(x1, y1), a20 = SphereQuadrantArea(x2, y0),
                a21 = SphereQuadrantArea(x2, y1), a02 = SphereQuadrantArea(x0, y2), a12 = SphereQuadrantArea(x1, y2), a22 = SphereQuadrantArea(x2, y2);
    return f4(std::max(1e-6f, fabsf(a00 - a01 - a10 + a11)), std::max(1e-6f, fabsf(a01 - a02 - a11 + a12)), std::max(1e-6f, fabsf(a10 - a11 - a20 + a21)), std::max(1e-6f, fabsf(a11 - a12 - a21 + a22)));
}
// bilinear fetch of an RGBA32F image, texel centres at +0.5; wrapU: wrap in x, else clamp
inline float3 sampleBilinear(const float* img, int W, int H, float u, float v, bool wrapU)
{
    const float tx = u * float(W) - 0.5f, ty = v * float(H) - 0.5f, fx = floorf(tx), fy = floorf(ty), wx = tx - fx, wy = ty - fy;
    auto ax = [&](int x) { return wrapU ? ((x % W) + W) % W : std::min(std::max(x, 0), W - 1); };
    auto ay = [&](int y) { return std::min(std::max(y, 0), H - 1); };
    const int x0 = ax(int(fx)), x1 = ax(int(fx) + 1), y0 = ay(int(fy)), y1 = ay(int(fy) + 1);
    auto px = [&](int x, int y) { const float* p = img + (size_t(y) * W + x) * 4; return f3(p[0], p[1], p[2]); };
    return (px(x0, y0) * (1 - wx) + px(x1, y0) * wx) * (1 - wy) + (px(x0, y1) * (1 - wx) + px(x1, y1) * wx) * wy;
}
inline float3 SampleSource(const Desc& d, uint px, uint py, uint face)
{
    const float3 dir = CubemapGetDirectionFor(face, (float(px) + 0.5f) / float(d.cubeDim), (float(py) + 0.5f) / float(d.cubeDim));
    if (d.sourceType == 1)
    {   // world_to_latlong_map
        const float3 p = normalize(dir);
        const float u = atan2f(p.x, -p.z) * 0.15915494309189535f + 0.5f, v = acosf(p.y) * 0.3183098861837907f;
        return sampleBilinear(d.source, int(d.sourceWidth), int(d.sourceHeight), u, v, true);
    }
    if (d.sourceType == 2)
    {   // major axis -> face and in-face uv (the inverse of CubemapGetDirectionFor)
        const float ax = fabsf(dir.x), ay = fabsf(dir.y), az = fabsf(dir.z);
        uint f; float cx, cy;
        if (ax >= ay && ax >= az) { f = dir.x > 0 ? 0u : 1u; cx = dir.x > 0 ? -dir.z / ax : dir.z / ax; cy = dir.y / ax; }
        else if (ay >= az) { f = dir.y > 0 ? 2u : 3u; cx = dir.x / ay; cy = dir.y > 0 ? -dir.z / ay : dir.z / ay; }
        else { f = dir.z > 0 ? 4u : 5u; cx = dir.z > 0 ? dir.x / az : -dir.x / az; cy = dir.y / az; }
        const float u = (cx + 1) * 0.5f, v = (1 - cy) * 0.5f;
        return sampleBilinear(d.source + size_t(f) * d.sourceWidth * d.sourceWidth * 4, int(d.sourceWidth), int(d.sourceWidth), u, v, false);
    }
    return f3(0);
}
inline float3 ComputeLightContribution(const Desc& d, uint px, uint py, uint face, const DirectionalLight& light)
{
    const float fade = 1.1f, dim = float(d.cubeDim);
    const float3 toLight = f3(-light.direction[0], -light.direction[1], -light.direction[2]);
    float dotMin = 1e30f, dotMax = -1e30f;
    const float ox[4] = { -fade, fade, -fade, fade }, oy[4] = { -fade, -fade, fade, fade };
    for (int k = 0; k < 4; k++)
    {
        const float3 dir = CubemapGetDirectionFor(face, (float(px) + 0.5f + 0.5f * ox[k]) / dim, (float(py) + 0.5f + 0.5f * oy[k]) / dim);
        const float c = dot(toLight, dir); dotMin = std::min(dotMin, c); dotMax = std::max(dotMax, c);
    }
    const float angleMin = acosf(std::min(std::max(dotMax, -1.0f), 1.0f)), angleMax = acosf(std::min(std::max(dotMin, -1.0f), 1.0f));
    float coverage = saturate(((light.angularSize * 0.5f) - angleMin) / (angleMax - angleMin + 1e-24f));
    coverage = powf(coverage, 4.0f);
    const float lightSolidAngle = 2 * K_PI * (1 - cosf(light.angularSize * 0.5f));
    return f3(light.colorIntensity[0], light.colorIntensity[1], light.colorIntensity[2]) * (coverage * (light.colorIntensity[3] / lightSolidAngle));
}
inline float4 GenerateTexel(const Desc& d, uint px, uint py, uint face)
{
    float3 c = SampleSource(d, px, py, face);
    for (uint i = 0; i < d.directionalLightCount; i++) c = c + ComputeLightContribution(d, px, py, face, d.lights[i]);
    c = c * f3(d.scaleColor[0], d.scaleColor[1], d.scaleColor[2]);
    const float hmax = 65504.0f;
    c = f3(std::min(std::max(c.x, 0.0f), hmax), std::min(std::max(c.y, 0.0f), hmax), std::min(std::max(c.z, 0.0f), hmax));
    return f4(lp(c.x), lp(c.y), lp(c.z), 1.0f);                    // the cube is RGBA16F
}

// mips[m]: 6 faces of (cubeDim >> m)^2 RGBA32F texels holding fp16 values; mip 0 and 1 by BaseLayerCS, the rest by MIPReduceCS
inline void bake(const Desc& d, std::vector<std::vector<float>>& mips)
{
    uint levels = 1; while ((d.cubeDim >> levels) > 0) levels++;
    mips.assign(levels, std::vector<float>());
    for (uint m = 0; m < levels; m++) mips[m].assign(size_t(6) * (d.cubeDim >> m) * (d.cubeDim >> m) * 4, 0.0f);
    auto at = [&](uint m, uint face, uint x, uint y) { const uint n = d.cubeDim >> m; return mips[m].data() + ((size_t(face) * n + y) * n + x) * 4; };
    auto store = [&](float* p, float4 v) { p[0] = lp(v.x); p[1] = lp(v.y); p[2] = lp(v.z); p[3] = lp(v.w); };
    const uint half = d.cubeDim / 2;
    for (uint face = 0; face < 6; face++)
    {
        #pragma omp parallel for schedule(dynamic, 4)
        for (int y = 0; y < int(half); y++) for (uint x = 0; x < half; x++)
        {
            const float4 w = CubemapTexelSolidAngle4(float(d.cubeDim), x * 2, uint(y) * 2);
            const float4 e00 = GenerateTexel(d, x * 2, uint(y) * 2, face), e01 = GenerateTexel(d, x * 2, uint(y) * 2 + 1, face), e10 = GenerateTexel(d, x * 2 + 1, uint(y) * 2, face), e11 = GenerateTexel(d, x * 2 + 1, uint(y) * 2 + 1, face);
            store(at(0, face, x * 2, uint(y) * 2), e00); store(at(0, face, x * 2, uint(y) * 2 + 1), e01); store(at(0, face, x * 2 + 1, uint(y) * 2), e10); store(at(0, face, x * 2 + 1, uint(y) * 2 + 1), e11);
            const float wsum = w.x + w.y + w.z + w.w;
            if (levels > 1) store(at(1, face, x, uint(y)), div4(e00 * w.x + e01 * w.y + e10 * w.z + e11 * w.w, wsum));
        }
    }
    for (uint m = 2; m < levels; m++)
    {
        const uint n = d.cubeDim >> m;
        for (uint face = 0; face < 6; face++) for (uint y = 0; y < n; y++) for (uint x = 0; x < n; x++)
        {
            const float4 w = CubemapTexelSolidAngle4(float(n * 2), x * 2, y * 2);
            auto ld = [&](uint sx, uint sy) { const float* p = at(m - 1, face, sx, sy); return f4(p[0], p[1], p[2], p[3]); };
            const float wsum = w.x + w.y + w.z + w.w;
            store(at(m, face, x, y), div4(ld(x * 2, y * 2) * w.x + ld(x * 2, y * 2 + 1) * w.y + ld(x * 2 + 1, y * 2) * w.z + ld(x * 2 + 1, y * 2 + 1) * w.w, wsum));
        }
    }
}

} } // namespace orc::envbake
