// opacity_masks.cpp — host-side baker of the per-triangle opacity masks (opacity_masks.h): the stand-in for the reference's OMM bake (Rtxpt/OpacityMicroMap/OmmBaker.cpp,
// OmmBuildQueue.cpp:30-60: alpha texture + cutoff + texcoords + indices per alpha-tested geometry).  Runs inside rtxpt_b200_upload_scene; plain C++, OpenMP over triangles.
#include "opacity_masks.h"
#include "../../include/rtxpt_b200.h"
#include <algorithm>
#include <cmath>
#include <cstring>

namespace pt { namespace om {

// alpha of texel (x, y) of mip 0 with wrap addressing, as the byte the sampler normalises (RGBA8) or as 255 * value (RGBA32F)
static float alphaAt(const AlphaSource& a, int x, int y)
{
    x %= a.width; if (x < 0) x += a.width; y %= a.height; if (y < 0) y += a.height;
    return a.rgba8 ? float(a.rgba8[(size_t(y) * a.width + x) * 4 + 3]) : a.rgba32f[(size_t(y) * a.width + x) * 4 + 3] * 255.0f;
}

// The alpha test passes when bilinear( alpha / 255 ) >= cutoffByte / 255 (traverse.cuh alphaTestPasses; PathTracerBridgeDonut.hlsli:969 samples mip 0).  A bilinear tap at
// texel-space position p reads texels floor(p - 0.5) and floor(p - 0.5) + 1 per axis and returns a convex combination of them, so a micro-triangle whose every reachable
// texel is > cutoff (strictly: one byte of margin keeps the filter's rounding away from the comparison) is opaque, < cutoff transparent.
static uint32_t classify(const AlphaSource& a, float cutoffByte, const float uvCorner[3][2])
{
    float lo[2] = { 3.0e38f, 3.0e38f }, hi[2] = { -3.0e38f, -3.0e38f }, mag = 0.0f;
    for (int k = 0; k < 3; k++) for (int c = 0; c < 2; c++) { const float t = uvCorner[k][c] * float(c == 0 ? a.width : a.height); if (!std::isfinite(t)) return kUnknown; lo[c] = std::min(lo[c], t); hi[c] = std::max(hi[c], t); mag = std::max(mag, std::fabs(t)); }
    if (!(mag < 1.0e6f)) return kUnknown;                                       // absurd coordinates (NaN returned above): let the texture test decide
    const float pad = 0.05f + mag * 1.0e-6f;                                    // hit-point UVs are interpolated in binary32 on the device (and with FMA contraction in the fast build)
    const int x0 = int(std::floor(lo[0] - 0.5f - pad)), x1 = int(std::floor(hi[0] - 0.5f + pad)) + 1, y0 = int(std::floor(lo[1] - 0.5f - pad)), y1 = int(std::floor(hi[1] - 0.5f + pad)) + 1;
    if (int64_t(x1 - x0 + 1) * int64_t(y1 - y0 + 1) > 4096) return kUnknown;    // a footprint this large is not uniform in practice; bounding the scan bounds the bake
    float mn = 3.0e38f, mx = -3.0e38f;
    for (int y = y0; y <= y1; y++) for (int x = x0; x <= x1; x++) { const float v = alphaAt(a, x, y); mn = std::min(mn, v); mx = std::max(mx, v); }
    if (mn >= cutoffByte + 1.0f) return kOpaque;
    if (mx <= cutoffByte - 1.0f) return kTransparent;
    return kUnknown;
}

void bakeTriangle(const AlphaSource& a, uint32_t cutoffByte, const float uv[3][2], uint32_t out[4], uint32_t counts[3])
{
    out[0] = out[1] = out[2] = out[3] = 0;
    const float inv = 1.0f / float(kN);
    auto uvAt = [&](float u, float v, float* o) { const float w = 1.0f - u - v; o[0] = uv[0][0] * w + uv[1][0] * u + uv[2][0] * v; o[1] = uv[0][1] * w + uv[1][1] * u + uv[2][1] * v; };
    for (int iv = 0; iv < kN; iv++) for (int iu = 0; iu < kN - iv; iu++) for (int upper = 0; upper < ((iu + iv < kN - 1) ? 2 : 1); upper++)
    {
        float c[3][2];
        const float u0 = float(iu) * inv, u1 = float(iu + 1) * inv, v0 = float(iv) * inv, v1 = float(iv + 1) * inv;
        if (!upper) { uvAt(u0, v0, c[0]); uvAt(u1, v0, c[1]); uvAt(u0, v1, c[2]); } else { uvAt(u1, v0, c[0]); uvAt(u1, v1, c[1]); uvAt(u0, v1, c[2]); }
        const uint32_t state = classify(a, float(cutoffByte), c), micro = uint32_t(iv * (2 * kN - iv) + 2 * iu + upper);
        out[micro >> 4] |= state << ((micro & 15u) * 2u);
        if (counts) counts[state]++;
    }
}

} } // namespace pt::om

// ---- host-only inspection hooks (no CUDA device needed): what tests/test_opacity_masks.py drives ------------------------------------------------------------------------------
extern "C" RTXPT_API int rtxpt_b200_host_bake_opacity_mask(const void* mip0, uint32_t width, uint32_t height, uint32_t format, uint32_t alphaCutoffByte, const float uv[6], uint32_t outMask[4])
{
    if (!mip0 || !uv || !outMask || width == 0 || height == 0) return RTXPT_ERR_INVALID_ARGUMENT;
    pt::om::AlphaSource a; a.width = int(width); a.height = int(height); a.rgba8 = nullptr; a.rgba32f = nullptr;
    if (format == RTXPT_FORMAT_RGBA32_FLOAT) a.rgba32f = static_cast<const float*>(mip0); else a.rgba8 = static_cast<const uint8_t*>(mip0);
    const float t[3][2] = { { uv[0], uv[1] }, { uv[2], uv[3] }, { uv[4], uv[5] } };
    pt::om::bakeTriangle(a, alphaCutoffByte, t, outMask, nullptr);
    return RTXPT_OK;
}
extern "C" RTXPT_API uint32_t rtxpt_b200_host_opacity_micro_index(float u, float v) { return pt::om::microIndex(u, v); }
