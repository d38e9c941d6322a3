// lights_bake.cpp — host-side light list baking at scene upload / constants change; stands in for the reference's per-frame GPU passes
//   EnvMapImportanceSamplingBaker::BuildMIPDescentImportanceMapCS  (Rtxpt/Lighting/Distant/EnvMapImportanceSamplingBaker.hlsl:62-96)
//   EnvLightsSubdivideBase / EnvLightsSubdivideBoost / EnvLightsFillLookupMap (Rtxpt/Lighting/LightsBaker.hlsl:262-500)
//   BakeEmissiveTriangles (:544-717), ComputeWeights (:835-879), ComputeProxyCounts (:881-948), ExecuteProxyJobs (:1048-1066)
//   LightsBaker::ProcessEmissiveGeometry buffer order (Rtxpt/Lighting/LightsBaker.cpp:663-827, :1076-1087)
// Tier: power-based global table (no temporal feedback, LightsBaker.cpp:1050-1051), importance boosters off, no analytic lights.
// It runs on the host so that the quantised results (fp16 edges, RGB8+log radiance, proxy counts, quad-tree topology) are reproducible
// bit for bit; moving it to CUDA is later-round work.  The float "atomic" weight sum is taken in index order.
#include "lights_bake.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <omp.h>

namespace pt {
namespace {

// ---- scalar helpers (host) ---------------------------------------------------------------------------------------------------
inline uint32_t bitsOf(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline float floatOf(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
inline float satf(float x) { return std::min(std::max(x, 0.0f), 1.0f); }

uint32_t toHalf(float value)        // round to nearest even, like __float2half_rn
{
    const uint32_t x = bitsOf(value), sign = (x >> 16) & 0x8000u; uint32_t mant = x & 0x007FFFFFu; const int exp = int((x >> 23) & 0xFF);
    if (exp == 0xFF) return sign | 0x7C00u | (mant ? (0x200u | (mant >> 13)) : 0u);
    const int e = exp - 112;
    if (e >= 31) return sign | 0x7C00u;
    if (e <= 0)
    {
        if (e < -10) return sign;
        mant |= 0x00800000u;
        const uint32_t shift = uint32_t(14 - e); uint32_t h = mant >> shift; const uint32_t rem = mant & ((1u << shift) - 1u), half = 1u << (shift - 1);
        if (rem > half || (rem == half && (h & 1u))) h++;
        return sign | h;
    }
    uint32_t h = (uint32_t(e) << 10) | (mant >> 13); const uint32_t rem = mant & 0x1FFFu;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) h++;
    return sign | h;
}
float fromHalf(uint32_t h)
{
    const uint32_t sign = (h & 0x8000u) << 16, exp = (h >> 10) & 0x1Fu, mant = h & 0x3FFu;
    if (exp == 0) { if (mant == 0) return floatOf(sign); return floatOf(bitsOf(float(mant) * (1.0f / 16777216.0f)) | sign); }
    if (exp == 31) return floatOf(sign | 0x7F800000u | (mant << 13));
    return floatOf(sign | ((exp + 112u) << 23) | (mant << 13));
}
inline float roundHalf(float v) { return fromHalf(toHalf(v)); }
inline float fastSqrtHost(float x) { int i; memcpy(&i, &x, 4); i = 0x1fbd1df5 + (i >> 1); float r; memcpy(&r, &i, 4); return r; }
inline uint32_t highBit(uint32_t v) { uint32_t r = 0; while (v >>= 1) r++; return r; }
inline float q8(float a) { return std::floor(a * 256.0f + 0.5f) * (1.0f / 256.0f); }     // 8 fractional bits, like the texture units

struct V3 { float x, y, z; };
inline V3 operator+(V3 a, V3 b) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
inline V3 operator-(V3 a, V3 b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
inline V3 operator*(V3 a, float s) { return { a.x * s, a.y * s, a.z * s }; }
inline V3 operator/(V3 a, float s) { return { a.x / s, a.y / s, a.z / s }; }
inline V3 crossv(V3 a, V3 b) { return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; }
inline float dotv(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 xformPoint(const float* m, V3 v) { return { m[0] * v.x + m[1] * v.y + m[2] * v.z + m[3], m[4] * v.x + m[5] * v.y + m[6] * v.z + m[7], m[8] * v.x + m[9] * v.y + m[10] * v.z + m[11] }; }

V3 octToDir(float u, float v)      // oct_to_ndir_equal_area_unorm
{
    const float px = u * 2.f - 1.f, py = v * 2.f - 1.f;
    const float d = 1.f - (std::fabs(px) + std::fabs(py));
    const float r = 1.f - std::fabs(d);
    const float phi = (r > 0.f) ? ((std::fabs(py) - std::fabs(px)) / r + 1.f) * 0.78539816339744830961f : 0.f;
    const float f = r * std::sqrt(2.f - r * r);
    auto sg = [](float x) { return (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f); };
    return { f * sg(px) * std::cos(phi), f * sg(py) * std::sin(phi), sg(d) * (1.f - r * r) };
}

// ---- host samplers (env cube for the importance map, emissive textures for the triangle bake) -------------------------------
struct HostCube
{
    const RtxptEnvCubeDesc* d;
    V3 texel(uint32_t face, int x, int y) const
    {
        const int n = int(d->faceSize);
        x = std::min(std::max(x, 0), n - 1); y = std::min(std::max(y, 0), n - 1);
        const float* p = d->faces[face][0] + (size_t(y) * n + x) * 4;
        return { p[0], p[1], p[2] };
    }
    V3 sampleMip0(V3 v) const
    {
        const float ax = std::fabs(v.x), ay = std::fabs(v.y), az = std::fabs(v.z);
        float m, s, t; uint32_t face;
        if (ax >= ay && ax >= az) { m = ax; if (v.x >= 0) { face = 0; s = -v.z; t = -v.y; } else { face = 1; s = v.z; t = -v.y; } }
        else if (ay >= az)        { m = ay; if (v.y >= 0) { face = 2; s = v.x; t = v.z; }  else { face = 3; s = v.x; t = -v.z; } }
        else                      { m = az; if (v.z >= 0) { face = 4; s = v.x; t = -v.y; } else { face = 5; s = -v.x; t = -v.y; } }
        const float n = float(d->faceSize);
        const float x = (s / m + 1.0f) * 0.5f * n - 0.5f, y = (t / m + 1.0f) * 0.5f * n - 0.5f;
        const float fx = std::floor(x), fy = std::floor(y);
        const float wx = q8(x - fx), wy = q8(y - fy);
        const int ix = int(fx), iy = int(fy);
        const V3 a = texel(face, ix, iy) * (1 - wx) + texel(face, ix + 1, iy) * wx;
        const V3 b = texel(face, ix, iy + 1) * (1 - wx) + texel(face, ix + 1, iy + 1) * wx;
        return a * (1 - wy) + b * wy;
    }
};

struct HostTexture
{
    const RtxptTextureDesc* d;
    static const float* srgbTable()
    {
        static float t[256]; static bool init = false;
        if (!init) { for (int i = 0; i < 256; i++) { const float c = float(i) / 255.0f; t[i] = (c <= 0.04045f) ? c / 12.92f : std::pow((c + 0.055f) / 1.055f, 2.4f); } init = true; }
        return t;
    }
    void texel(uint32_t level, int x, int y, float out[4]) const
    {
        const int w = int(std::max(1u, d->width >> level)), h = int(std::max(1u, d->height >> level));
        x = ((x % w) + w) % w; y = ((y % h) + h) % h;
        const size_t i = size_t(y) * w + x;
        if (d->format == RTXPT_FORMAT_RGBA32_FLOAT) { const float* p = (const float*)d->mips[level] + i * 4; out[0] = p[0]; out[1] = p[1]; out[2] = p[2]; out[3] = p[3]; return; }
        const uint8_t* p = (const uint8_t*)d->mips[level] + i * 4;
        if (d->format == RTXPT_FORMAT_RGBA8_SRGB) { const float* l = srgbTable(); out[0] = l[p[0]]; out[1] = l[p[1]]; out[2] = l[p[2]]; out[3] = float(p[3]) / 255.0f; }
        else { out[0] = float(p[0]) / 255.0f; out[1] = float(p[1]) / 255.0f; out[2] = float(p[2]) / 255.0f; out[3] = float(p[3]) / 255.0f; }
    }
    void bilinear(uint32_t level, float u, float v, float out[4]) const
    {
        const float w = float(std::max(1u, d->width >> level)), h = float(std::max(1u, d->height >> level));
        const float x = u * w - 0.5f, y = v * h - 0.5f, fx = std::floor(x), fy = std::floor(y);
        const float ax = q8(x - fx), ay = q8(y - fy);
        float t00[4], t10[4], t01[4], t11[4];
        texel(level, int(fx), int(fy), t00); texel(level, int(fx) + 1, int(fy), t10); texel(level, int(fx), int(fy) + 1, t01); texel(level, int(fx) + 1, int(fy) + 1, t11);
        for (int k = 0; k < 4; k++) { const float a = t00[k] * (1 - ax) + t10[k] * ax, b = t01[k] * (1 - ax) + t11[k] * ax; out[k] = a * (1 - ay) + b * ay; }
    }
    void trilinear(float u, float v, float lod, float out[4]) const
    {
        lod = std::min(std::max(lod, 0.0f), float(d->mipLevels - 1));
        const float fl = std::floor(lod), a = q8(lod - fl);
        const uint32_t l0 = uint32_t(fl), l1 = std::min(l0 + 1, d->mipLevels - 1);
        bilinear(l0, u, v, out);
        if (a == 0.0f || l1 == l0) return;
        float c1[4]; bilinear(l1, u, v, c1);
        for (int k = 0; k < 4; k++) out[k] = out[k] * (1 - a) + c1[k] * a;
    }
};

// ---- packed light helpers ------------------------------------------------------------------------------------------------------
constexpr uint32_t kTypeSphere = 0, kTypeTriangle = 1, kTypePoint = 4, kTypeEnvQuad = 5;
constexpr uint32_t kShapingEnableBit = 1u << 28, kShapingUseMinFalloff = 1u << 30;
float unpackLogRadiance(uint32_t lr) { return (lr == 0) ? 0.f : std::exp2((float(lr - 1) / 65534.0f) * 48.0f + -8.0f); }
void packColor(V3 radiance, BakedLight& li)       // PolymorphicLight::PackColor (PolymorphicLight.hlsli:776-795)
{
    const float intensity = std::max(radiance.x, std::max(radiance.y, radiance.z));
    if (!(intensity > 0.0f)) return;
    const float logRadiance = satf((std::log2(intensity) - (-8.0f)) / 48.0f);
    const uint32_t packed = std::min(uint32_t(std::ceil(logRadiance * 65534.0f)) + 1, 0xffffu);
    const float unpacked = unpackLogRadiance(packed);
    auto pk = [](float c) { return uint32_t(std::floor(satf(c) * 255.0f + 0.5f)) & 0xFFu; };
    li.logRadiance |= packed;
    li.colorTypeAndFlags |= pk(radiance.x / unpacked) | (pk(radiance.y / unpacked) << 8) | (pk(radiance.z / unpacked) << 16);
}
V3 unpackColor(const BakedLight& li)
{
    const float r = unpackLogRadiance(li.logRadiance & 0xffff);
    return { float(li.colorTypeAndFlags & 0xFF) / 255.0f * r, float((li.colorTypeAndFlags >> 8) & 0xFF) / 255.0f * r, float((li.colorTypeAndFlags >> 16) & 0xFF) / 255.0f * r };
}

// ---- env quad tree --------------------------------------------------------------------------------------------------------------
constexpr uint32_t QT_BASE = 4, QT_SUBDIV = 24, QT_UNBOOSTED = QT_BASE * QT_BASE + 3 * QT_SUBDIV, QT_BOOST_DEPTH = 3, QT_BOOST_SUBDIV = 20;
constexpr uint32_t QT_BOOST_MULT = QT_BOOST_SUBDIV * 3 + 1;
constexpr uint32_t QT_TOTAL = QT_UNBOOSTED * QT_BOOST_MULT;
static_assert(QT_TOTAL == kEnvQuadLightCount, "env quad-tree node count");

struct QuadTreeBuilder
{
    const LightBakeState& st;
    const float* texel(uint32_t x, uint32_t y, uint32_t mip) const { const uint32_t n = kEnvImportanceMapDim >> mip; return &st.envRadianceMips[mip][(size_t(y) * n + x) * 4]; }
    static uint32_t pack(uint32_t dim, uint32_t x, uint32_t y) { return (highBit(dim) << 28) | (x << 14) | y; }
    static void unpack(uint32_t p, uint32_t& dim, uint32_t& x, uint32_t& y) { dim = 1u << (p >> 28); x = (p >> 14) & 0x3FFF; y = p & 0x3FFF; }
    uint32_t buildWeight(uint32_t dim, uint32_t x, uint32_t y, uint32_t index, uint32_t depthLimit) const     // EnvironmentComputeWeightForQTBuild
    {
        const uint32_t mip = st.envMipCount - highBit(dim) - 1;
        float w = float(1u << (mip * 2)) * texel(x, y, mip)[3];
        w = std::max((1.0f / 100.0f) * (1.0f / 100.0f) * float(mip), w);
        w *= (mip > depthLimit) ? 1.0f : 0.0f;
        return (std::min(uint32_t(fastSqrtHost(w) * 100 + 0.5f), 0x000FFFFFu) << 12) | index;
    }
    void refine(std::vector<uint32_t>& nodes, std::vector<uint32_t>& keys, uint32_t count, uint32_t steps, uint32_t depthLimit) const
    {
        for (uint32_t s = 0; s < steps; s++)
        {
            uint32_t best = 0;
            for (uint32_t i = 0; i < count; i++) best = std::max(best, keys[i]);
            const uint32_t at = best & 0xFFF;
            uint32_t dim, x, y; unpack(nodes[at], dim, x, y);
            for (uint32_t k = 0; k < 4; k++)
            {
                const uint32_t ni = (k == 0) ? at : (count + k - 1);
                nodes[ni] = pack(dim * 2, x * 2 + (k % 2), y * 2 + (k / 2));
                keys[ni] = buildWeight(dim * 2, x * 2 + (k % 2), y * 2 + (k / 2), ni, depthLimit);
            }
            count += 3;
        }
    }
};

} // namespace

void LightBaker::buildEnvRadianceMap(const RtxptEnvCubeDesc& cube, LightBakeState& st)
{
    const uint32_t N = kEnvImportanceMapDim, S = 4;
    st.envMipCount = highBit(N) + 1;
    st.envRadianceMips.assign(st.envMipCount, std::vector<float>());
    st.envRadianceMips[0].resize(size_t(N) * N * 4);
    const HostCube hc{ &cube };
    const float inv = 1.0f / float(S * S);
    #pragma omp parallel for schedule(dynamic, 8)
    for (int py = 0; py < int(N); py++)
        for (uint32_t px = 0; px < N; px++)
        {
            float L = 0.f; V3 R = { 0, 0, 0 };
            for (uint32_t y = 0; y < S; y++)
                for (uint32_t x = 0; x < S; x++)
                {
                    const V3 dir = octToDir((float(px * S + x) + 0.5f) / float(N * S), (float(uint32_t(py) * S + y) + 0.5f) / float(N * S));
                    const V3 rad = hc.sampleMip0(dir);
                    const float lum = rad.x * 0.2126f + rad.y * 0.7152f + rad.z * 0.0722f, avg = (rad.x + rad.y + rad.z) / 3.0f;
                    L += (lum + avg) * 0.5f;
                    R = R + rad;
                }
            float* o = &st.envRadianceMips[0][(size_t(py) * N + px) * 4];
            o[0] = roundHalf(R.x * inv); o[1] = roundHalf(R.y * inv); o[2] = roundHalf(R.z * inv); o[3] = roundHalf(L * inv);     // RGBA16F target
        }
    for (uint32_t m = 1; m < st.envMipCount; m++)
    {
        const uint32_t n = N >> m, pn = N >> (m - 1);
        st.envRadianceMips[m].resize(size_t(n) * n * 4);
        const std::vector<float>& src = st.envRadianceMips[m - 1];
        for (uint32_t y = 0; y < n; y++) for (uint32_t x = 0; x < n; x++) for (uint32_t c = 0; c < 4; c++)
        {
            const float s = src[(size_t(2 * y) * pn + 2 * x) * 4 + c] + src[(size_t(2 * y) * pn + 2 * x + 1) * 4 + c]
                          + src[(size_t(2 * y + 1) * pn + 2 * x) * 4 + c] + src[(size_t(2 * y + 1) * pn + 2 * x + 1) * 4 + c];
            st.envRadianceMips[m][(size_t(y) * n + x) * 4 + c] = roundHalf(s * 0.25f);
        }
    }
}


// ---- analytic scene lights -> packed records (the CPU half of the reference's light list: Rtxpt/Lighting/LightsBaker.cpp:414-556) ---------
// The reference packs these on the host with its own helpers: a truncating float->half (multiply by 2^-112, shift the mantissa) and an
// octahedral encoding that maps to [0,1] twice; the shader-side decoders (Utils.hlsli:128-153) undo exactly that.
static uint32_t halfTruncating(float v)
{
    const uint32_t u = bitsOf(v * floatOf(0x07800000u));
    return (((u & 0x80000000u) >> 16) | ((u & 0x0fffffffu) >> 13)) & 0xFFFFu;
}
static uint32_t octUnorm32(V3 n)
{
    const float l1 = std::fabs(n.x) + std::fabs(n.y) + std::fabs(n.z);
    float x = n.x / l1, y = n.y / l1; const float z = n.z / l1;
    if (!(z >= 0.0f)) { const float wx = (1.0f - std::fabs(y)) * (x >= 0.0f ? 1.0f : -1.0f), wy = (1.0f - std::fabs(x)) * (y >= 0.0f ? 1.0f : -1.0f); x = wx; y = wy; }
    x = x * 0.5f + 0.5f; y = y * 0.5f + 0.5f;
    const float px = satf(x * 0.5f + 0.5f), py = satf(y * 0.5f + 0.5f);
    return uint32_t(px * float(0xfffe)) | (uint32_t(py * float(0xfffe)) << 16);
}
[[maybe_unused]] static V3 octUnorm32ToDir(uint32_t p)
{
    float fx = satf(float(p & 0xffff) / float(0xfffe)) * 2.0f - 1.0f, fy = satf(float(p >> 16) / float(0xfffe)) * 2.0f - 1.0f;
    fx = fx * 2.0f - 1.0f; fy = fy * 2.0f - 1.0f;
    V3 n = { fx, fy, 1.0f - std::fabs(fx) - std::fabs(fy) };
    const float t = satf(-n.z);
    n.x += (n.x >= 0.0f) ? -t : t; n.y += (n.y >= 0.0f) ? -t : t;
    return n / std::sqrt(dotv(n, n));
}
static void convertAnalyticLight(const RtxptLightDesc& L, BakedLight& li, BakedLightEx& ex)
{
    memset(&li, 0, sizeof(li)); memset(&ex, 0, sizeof(ex));
    const float pi = 3.14159265358979323846f, toRad = pi / 180.0f;
    const V3 flux = V3{ L.color[0], L.color[1], L.color[2] } * L.intensity;
    const bool spot = (L.type == RTXPT_LIGHT_SPOT);
    const uint32_t falloffFlag = (spot && L.outerAngle < 0) ? kShapingUseMinFalloff : 0u;
    V3 axis = { 0, 0, 0 };
    if (spot) { axis = { L.direction[0], L.direction[1], L.direction[2] }; axis = axis / std::sqrt(dotv(axis, axis)); }
    if (L.radius == 0.f)
    {   // kPoint record: compiled out of the reference's sampling code, kept for the index layout
        li.colorTypeAndFlags = (kTypePoint << 24) | falloffFlag;
        packColor(flux, li);
        if (spot) { li.direction1 = octUnorm32(axis); li.direction2 = halfTruncating(std::fabs(L.outerAngle) * toRad) | (halfTruncating(L.innerAngle * toRad) << 16); }
        else li.direction2 = halfTruncating(pi) | (halfTruncating(0.0f) << 16);
    }
    else
    {   // sphere light; radiance = flux over the projected disc
        li.colorTypeAndFlags = (kTypeSphere << 24) | falloffFlag | (spot ? kShapingEnableBit : 0u);
        packColor(flux / (pi * (L.radius * L.radius)), li);
        li.scalars = halfTruncating(L.radius);
        if (spot && std::fabs(L.outerAngle) > 0)
        {
            const float softness = satf(1.f - L.innerAngle / std::fabs(L.outerAngle));
            ex.primaryAxis = octUnorm32(axis);
            ex.cosConeAngleAndSoftness = halfTruncating(std::cos(std::fabs(L.outerAngle) * toRad)) | (halfTruncating(softness) << 16);
        }
    }
    li.center[0] = L.position[0]; li.center[1] = L.position[1]; li.center[2] = L.position[2];
}
static float sphereLightPower(const BakedLight& l, const BakedLightEx& ex)      // SphereLight::GetPower x getShapingFluxFactor
{
    const float pi = 3.14159265358979323846f;
    const float radius = fromHalf(l.scalars & 0xffff);
    const V3 rad = unpackColor(l);
    float shaping = 1.0f;
    if (l.colorTypeAndFlags & kShapingEnableBit)
    {
        const float cosCone = fromHalf(ex.cosConeAngleAndSoftness & 0xffff), softness = fromHalf(ex.cosConeAngleAndSoftness >> 16);
        shaping = (1.0f - cosCone) * (1.0f + (0.5f - 1.0f) * softness) * 0.5f;
    }
    return (4 * pi * (radius * radius)) * pi * (rad.x * 0.2126f + rad.y * 0.7152f + rad.z * 0.0722f) * shaping;        // getSurfaceArea() = 4 pi sq(radius) first (PolymorphicLight.hlsli:224-232; pinned by tests/golden/lights_golden.npz)
}

void LightBaker::finalize(const RtxptPathTracerConstants& consts, LightBakeState& st)
{
    st.envEnabled = st.hasEnvCube && consts.envMap.Enabled != 0.0f;
    st.lights.assign(kEnvQuadLightCount, BakedLight());
    for (BakedLight& l : st.lights) { memset(&l, 0, sizeof(l)); l.colorTypeAndFlags = kTypeEnvQuad << 24; }
    st.envLookupMap.clear();
    if (st.envEnabled)
    {
        QuadTreeBuilder qt{ st };
        std::vector<uint32_t> nodes(QT_UNBOOSTED), keys(QT_UNBOOSTED);
        for (uint32_t i = 0; i < QT_BASE * QT_BASE; i++) { nodes[i] = QuadTreeBuilder::pack(QT_BASE, i / QT_BASE, i % QT_BASE); keys[i] = qt.buildWeight(QT_BASE, i / QT_BASE, i % QT_BASE, i, QT_BOOST_DEPTH); }
        qt.refine(nodes, keys, QT_BASE * QT_BASE, QT_SUBDIV, QT_BOOST_DEPTH);
        const float* cm = consts.envMap.ColorMultiplier;
        const float avgMul = (cm[0] + cm[1] + cm[2]) / 3.0f;
        const float relImportance = consts.distantVsLocalImportance * 0.0002f;        // LightsBaker.cpp:1029-1030
        std::vector<uint32_t> bn(QT_BOOST_MULT), bk(QT_BOOST_MULT);
        for (uint32_t g = 0; g < QT_UNBOOSTED; g++)
        {
            uint32_t dim, x, y; QuadTreeBuilder::unpack(nodes[g], dim, x, y);
            bn[0] = nodes[g]; bk[0] = qt.buildWeight(dim, x, y, 0, 0);
            qt.refine(bn, bk, 1, QT_BOOST_SUBDIV, 0);
            for (uint32_t i = 0; i < QT_BOOST_MULT; i++)
            {
                QuadTreeBuilder::unpack(bn[i], dim, x, y);
                const uint32_t mip = st.envMipCount - highBit(dim) - 1;
                const float* v = qt.texel(x, y, mip);
                const float weight = float(1u << (mip * 2)) * std::max(0.0f, v[3] * avgMul * relImportance);
                BakedLight li; memset(&li, 0, sizeof(li));
                packColor({ v[0] * cm[0], v[1] * cm[1], v[2] * cm[2] }, li);
                li.direction1 = (x << 16) | y; li.direction2 = dim << 16; li.scalars = bitsOf(weight);
                li.colorTypeAndFlags |= kTypeEnvQuad << 24;
                const V3 local = octToDir((float(x) + 0.5f) / float(dim), (float(y) + 0.5f) / float(dim));
                const float* m = consts.envMap.Transform;
                li.center[0] = (local.x * m[0] + local.y * m[4] + local.z * m[8]) * 100000.0f;
                li.center[1] = (local.x * m[1] + local.y * m[5] + local.z * m[9]) * 100000.0f;
                li.center[2] = (local.x * m[2] + local.y * m[6] + local.z * m[10]) * 100000.0f;
                st.lights[g * QT_BOOST_MULT + i] = li;
            }
        }
        st.envLookupMap.assign(size_t(kEnvImportanceMapDim) * kEnvImportanceMapDim, 0);
        for (uint32_t li = 0; li < kEnvQuadLightCount; li++)
        {
            const BakedLight& l = st.lights[li];
            const uint32_t nx = l.direction1 >> 16, ny = l.direction1 & 0xFFFF, nd = l.direction2 >> 16, ds = kEnvImportanceMapDim / nd;
            for (uint32_t yy = 0; yy < ds; yy++) for (uint32_t xx = 0; xx < ds; xx++) st.envLookupMap[size_t(ny * ds + yy) * kEnvImportanceMapDim + nx * ds + xx] = li;
        }
    }
    st.lights.insert(st.lights.end(), st.analyticLights.begin(), st.analyticLights.end());
    st.lights.insert(st.lights.end(), st.triangleLights.begin(), st.triangleLights.end());
    finalizeWeightsAndProxies(consts, st);
}

void LightBaker::setAnalyticLights(const RtxptLightDesc* lights, uint32_t count, LightBakeState& st)
{
    st.analyticLights.clear(); st.analyticLightsEx.clear();
    for (uint32_t i = 0; i < count && lights; i++)
    {
        BakedLight li; BakedLightEx ex; convertAnalyticLight(lights[i], li, ex);
        st.analyticLights.push_back(li); st.analyticLightsEx.push_back(ex);
    }
}
void LightBaker::snapshot(const LightBakeState& st, LightListSnapshot& out)
{
    out.valid = st.lights.size() >= kEnvQuadLightCount; out.envEnabled = st.envEnabled;
    out.analyticCount = uint32_t(st.analyticLights.size()); out.triangleCount = uint32_t(st.triangleLights.size());
    out.envNodes.resize(size_t(kEnvQuadLightCount) * 2);
    for (uint32_t i = 0; i < kEnvQuadLightCount && out.valid; i++) { out.envNodes[2 * i] = st.lights[i].direction1; out.envNodes[2 * i + 1] = st.lights[i].direction2; }
    out.envLookupMap = st.envLookupMap;
}
void LightBaker::buildRemap(const LightListSnapshot& past, const LightBakeState& st, std::vector<uint32_t>& pastToCurrent, std::vector<uint32_t>& currentToPast)
{
    const uint32_t E = kEnvQuadLightCount, nPast = past.analyticCount, nCur = uint32_t(st.analyticLights.size()), tPast = past.triangleCount, tCur = uint32_t(st.triangleLights.size());
    const uint32_t kInvalid = 0xFFFFFFFFu;
    pastToCurrent.assign(past.total(), kInvalid); currentToPast.assign(size_t(E) + nCur + tCur, kInvalid);
    const bool envBoth = past.envEnabled && st.envEnabled && past.envLookupMap.size() == size_t(kEnvImportanceMapDim) * kEnvImportanceMapDim && st.envLookupMap.size() == past.envLookupMap.size();
    if (envBoth)
        for (uint32_t i = 0; i < E; i++)
        {   // the node's corner texel looked up in the other frame's importance map (the mapping need not be one to one)
            uint32_t dim = past.envNodes[2 * i + 1] >> 16, x = past.envNodes[2 * i] >> 16, y = past.envNodes[2 * i] & 0xFFFFu;
            if (dim) { const uint32_t ds = kEnvImportanceMapDim / dim; pastToCurrent[i] = st.envLookupMap[size_t(y * ds) * kEnvImportanceMapDim + x * ds]; }
            dim = st.lights[i].direction2 >> 16; x = st.lights[i].direction1 >> 16; y = st.lights[i].direction1 & 0xFFFFu;
            if (dim) { const uint32_t ds = kEnvImportanceMapDim / dim; currentToPast[i] = past.envLookupMap[size_t(y * ds) * kEnvImportanceMapDim + x * ds]; }
        }
    for (uint32_t k = 0; k < nPast; k++) pastToCurrent[E + k] = k < nCur ? E + k : kInvalid;
    for (uint32_t k = 0; k < nCur; k++) currentToPast[E + k] = k < nPast ? E + k : kInvalid;
    for (uint32_t t = 0; t < tPast; t++) pastToCurrent[E + nPast + t] = t < tCur ? E + nCur + t : kInvalid;       // emissive geometry is baked at upload: same triangles, shifted block
    for (uint32_t t = 0; t < tCur; t++) currentToPast[E + nCur + t] = t < tPast ? E + nPast + t : kInvalid;
}

void LightBaker::prepareScene(const RtxptSceneDesc& scene, std::vector<RtxptSubInstanceData>& subInstances, LightBakeState& st)
{
    st.hasEnvCube = scene.envCube.faceSize != 0;
    st.envRadianceMips.clear();
    if (st.hasEnvCube) buildEnvRadianceMap(scene.envCube, st);
    // analytic lights sit between the env quad-tree slots and the emissive triangles (LightsBaker.cpp:596-640)
    setAnalyticLights(scene.lights, scene.lights ? scene.lightCount : 0u, st);
    const uint32_t firstTriangleLight = kEnvQuadLightCount + uint32_t(st.analyticLights.size());
    // emissive triangles: one light per triangle of every emissive geometry instance
    st.triangleLights.clear();
    st.triangleLightCount = 0;
    for (uint32_t ii = 0; ii < scene.instanceCount; ii++)
    {
        const RtxptInstanceData& inst = scene.instances[ii];
        for (uint32_t gi = 0; gi < inst.numGeometries; gi++)
        {
            RtxptSubInstanceData& sub = subInstances[inst.firstGeometryInstanceIndex + gi];
            const RtxptGeometryData& g = scene.geometries[inst.firstGeometryIndex + gi];
            const RtxptMaterialData& m = scene.materials[sub.GlobalGeometryIndex_PTMaterialDataIndex & 0xFFFF];
            const uint32_t triCount = g.numIndices / 3;
            const bool emissive = m.EmissiveColor[0] > 0 || m.EmissiveColor[1] > 0 || m.EmissiveColor[2] > 0;       // PTMaterial::IsEmissive
            if (!emissive || firstTriangleLight + st.triangleLights.size() + triCount >= kMaxLights) { sub.EmissiveLightMappingOffset = 0xFFFFFFFFu; continue; }
            sub.EmissiveLightMappingOffset = firstTriangleLight + uint32_t(st.triangleLights.size());
            const float* xf = inst.transform;
            const float det = xf[0] * (xf[5] * xf[10] - xf[6] * xf[9]) - xf[1] * (xf[4] * xf[10] - xf[6] * xf[8]) + xf[2] * (xf[4] * xf[9] - xf[5] * xf[8]);
            const uint8_t* ib = (const uint8_t*)scene.buffers[g.indexBufferIndex].data; const uint8_t* vb = (const uint8_t*)scene.buffers[g.vertexBufferIndex].data;
            for (uint32_t t = 0; t < triCount; t++)
            {
                uint32_t idx[3]; memcpy(idx, ib + g.indexOffset + t * 12, 12);
                V3 p[3];
                for (int k = 0; k < 3; k++) { float v[3]; memcpy(v, vb + g.positionOffset + idx[k] * 12, 12); p[k] = xformPoint(xf, { v[0], v[1], v[2] }); }
                V3 radiance = { m.EmissiveColor[0], m.EmissiveColor[1], m.EmissiveColor[2] };
                if (m.EmissiveTextureIndex != 0xFFFFFFFFu && g.texCoord1Offset != ~0u && (m.Flags & RTXPT_MATFLAG_UseEmissiveTexture))
                {   // the reference takes one anisotropic SampleGrad over the inscribed ellipse; here: one trilinear tap at the LOD of the longer axis
                    float uv[3][2]; for (int k = 0; k < 3; k++) memcpy(uv[k], vb + g.texCoord1Offset + idx[k] * 8, 8);
                    const float e[3][2] = { { uv[1][0] - uv[0][0], uv[1][1] - uv[0][1] }, { uv[2][0] - uv[1][0], uv[2][1] - uv[1][1] }, { uv[0][0] - uv[2][0], uv[0][1] - uv[2][1] } };
                    const float l[3] = { std::sqrt(e[0][0] * e[0][0] + e[0][1] * e[0][1]), std::sqrt(e[1][0] * e[1][0] + e[1][1] * e[1][1]), std::sqrt(e[2][0] * e[2][0] + e[2][1] * e[2][1]) };
                    int s, a, b;
                    if (l[0] < l[1] && l[0] < l[2]) { s = 0; a = 1; b = 2; } else if (l[1] < l[2]) { s = 1; a = 2; b = 0; } else { s = 2; a = 0; b = 1; }
                    const float sg[2] = { e[s][0] * (2.0f / 3.0f), e[s][1] * (2.0f / 3.0f) }, lg[2] = { (e[a][0] + e[b][0]) / 3.0f, (e[a][1] + e[b][1]) / 3.0f };
                    const HostTexture tex{ &scene.textures[m.EmissiveTextureIndex & 0xFFFF] };
                    const float dw = float(tex.d->width), dh = float(tex.d->height);
                    const float la = std::sqrt(sg[0] * dw * sg[0] * dw + sg[1] * dh * sg[1] * dh), lb = std::sqrt(lg[0] * dw * lg[0] * dw + lg[1] * dh * lg[1] * dh);
                    float c[4]; tex.trilinear((uv[0][0] + uv[1][0] + uv[2][0]) / 3.0f, (uv[0][1] + uv[1][1] + uv[2][1]) / 3.0f, std::log2(std::max(std::max(la, lb), 1e-8f)), c);
                    radiance = { radiance.x * c[0], radiance.y * c[1], radiance.z * c[2] };
                }
                radiance = { std::max(radiance.x, 0.0f), std::max(radiance.y, 0.0f), std::max(radiance.z, 0.0f) };
                const V3 e1 = (det < 0.f) ? (p[2] - p[0]) : (p[1] - p[0]), e2 = (det < 0.f) ? (p[1] - p[0]) : (p[2] - p[0]);
                if (std::max(radiance.x, std::max(radiance.y, radiance.z)) < 1e-7f) radiance = { 0, 0, 0 };
                BakedLight li; memset(&li, 0, sizeof(li));
                packColor(radiance, li);
                const V3 c = p[0] + ((e1 + e2) / 3.0f);
                li.center[0] = c.x; li.center[1] = c.y; li.center[2] = c.z;
                // TriangleLight::Store (PolymorphicLight.hlsli:510-513) passes the packed words through a `float3 edges`: uint -> float -> uint, each word rounded to 24
                // significant bits (the low bits of edge1's halves are lost).  Reproduced as is - these are the records the reference samples (tests/golden/lights_golden.npz)
                auto viaFloat = [](uint32_t packed) { const float f = float(packed); return f >= 4294967296.0f ? 0xFFFFFFFFu : uint32_t(f); };
                li.direction1 = viaFloat((toHalf(e1.x) & 0xffff) | (toHalf(e2.x) << 16));
                li.direction2 = viaFloat((toHalf(e1.y) & 0xffff) | (toHalf(e2.y) << 16));
                li.scalars = viaFloat((toHalf(e1.z) & 0xffff) | (toHalf(e2.z) << 16));
                li.colorTypeAndFlags |= kTypeTriangle << 24;
                st.triangleLights.push_back(li);
                st.triangleLightCount++;
            }
        }
    }
}

void LightBaker::finalizeWeightsAndProxies(const RtxptPathTracerConstants& consts, LightBakeState& st)
{
    // weights = flux^0.8 of the PACKED light (ComputeWeight, LightsBaker.hlsl:738-750)
    const uint32_t n = uint32_t(st.lights.size());
    std::vector<float> w(n);
    for (uint32_t i = 0; i < n; i++)
    {
        const BakedLight& l = st.lights[i];
        float flux = 0;
        const uint32_t type = (l.colorTypeAndFlags >> 24) & 0xf;
        if (type == kTypeTriangle)
        {
            const V3 e1 = { fromHalf(l.direction1 & 0xffff), fromHalf(l.direction2 & 0xffff), fromHalf(l.scalars & 0xffff) };
            const V3 e2 = { fromHalf(l.direction1 >> 16), fromHalf(l.direction2 >> 16), fromHalf(l.scalars >> 16) };
            const V3 nrm = crossv(e1, e2);
            const float len = std::sqrt(dotv(nrm, nrm));
            const float area = (len > 0.0f) ? 0.5f * len : 0.0f;
            const V3 rad = unpackColor(l);
            flux = area * 3.14159265358979323846f * (rad.x * 0.2126f + rad.y * 0.7152f + rad.z * 0.0722f);
        }
        else if (type == kTypeEnvQuad) flux = floatOf(l.scalars);
        else if (type == kTypeSphere) flux = sphereLightPower(l, st.analyticLightsEx[i - kEnvQuadLightCount]);
        float weight = std::pow(flux, 0.8f);
        if (weight < 1e-8f) weight = 0;
        w[i] = weight;
    }
    float total = 0;
    for (uint32_t g0 = 0; g0 < n; g0 += 32 * 128)
    {
        float groupSum = 0;
        for (uint32_t b0 = g0; b0 < std::min(n, g0 + 32 * 128); b0 += 32)
        {
            float blockSum = 0;
            for (uint32_t i = b0; i < std::min(n, b0 + 32); i++) blockSum += w[i];
            groupSum += blockSum;
        }
        total += groupSum;
    }
    st.weightsSum = total; st.weights = w;
    const uint32_t budget = 12 * std::max(n, kMaxLights / 10);
    st.proxyCounters.assign(n, 0); st.proxyIndices.clear();
    for (uint32_t i = 0; i < n; i++)
    {
        uint32_t proxies = 0;
        if (w[i] > 0) proxies = (consts.NEEType == 0) ? 1u : uint32_t(std::ceil((float(budget - n) * w[i]) / total));
        proxies = std::min(proxies, 256u * 1024u - 1u);
        st.proxyCounters[i] = proxies;
        st.proxyIndices.insert(st.proxyIndices.end(), proxies, i);
    }
}

} // namespace pt
