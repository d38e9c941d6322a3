// ORACLE — test infrastructure only (see pt_math.h).
// pt_lights.h: light list baking and the runtime light sampler, restated from
//   Rtxpt/Shaders/PathTracer/Lighting/PolymorphicLight.h:14-72, PolymorphicLight.hlsli:395-520 (TriangleLight), :560-640
//   (EnvironmentQuadLight), :643-676 (CalcSample dispatch), :760-795 (radiance/colour packing)
//   Rtxpt/Lighting/LightsBaker.hlsl:167-201 (env radiance/weight), :241-470 (env quad-tree subdivision base+boost),
//   :473-500 (lookup map), :544-717 (BakeEmissiveTriangles), :738-750 (ComputeWeight), :835-948 (weights, proxy counts),
//   :1010-1066 (proxy fill);  Rtxpt/Lighting/LightsBaker.cpp:561-583, :663-827, :1029-1051, :1076-1087 (buffer order)
//   Rtxpt/Lighting/Distant/EnvMapImportanceSamplingBaker.hlsl:62-96 (importance + radiance map)
//   Rtxpt/Shaders/PathTracer/Lighting/LightSampler.hlsli:109-117 (SampleGlobal), :282-328 (MIS), :412-432
// Tier: power-based global proxy table only (first frame / no feedback: LightsBaker.cpp:1050-1051 forces
// LocalToGlobalSampleRatio = GlobalFeedbackUseWeight = 0); importance boosters off (LightsBaker.h:245-248 UI toggles);
// analytic lights are not listed yet.  The float atomic weight sum (LightsBaker.hlsl:719-735, order undefined in the reference)
// is taken in index order: 32-light blocks, then 128-block groups, then groups.
#pragma once
#include "pt_scene.h"
#include "pt_rng.h"

namespace orc {

static const uint kLightTypeSphere = 0, kLightTypeTriangle = 1, kLightTypePoint = 4, kLightTypeEnvironmentQuad = 5;      // PolymorphicLight.h:28-40
static const uint kPolymorphicLightShapingEnableBit = 1u << 28, kPolymorphicLightShapingUseMinFalloff = 1u << 30;
static const float kMinSpotlightFalloff = 0.0001f;
static const uint kPolymorphicLightTypeShift = 24;
static const float kMinLog2Radiance = -8.f, kMaxLog2Radiance = 40.f;
static const float DISTANT_LIGHT_DISTANCE = 100000.0f;
static const uint ENVQT_BASE_RES = 4, ENVQT_SUBDIV = 24, ENVQT_UNBOOSTED = ENVQT_BASE_RES * ENVQT_BASE_RES + 3 * ENVQT_SUBDIV;   // 88
static const uint ENVQT_BOOST_DPT = 3, ENVQT_BOOST_SUBDIV = 20, ENVQT_BOOST_MULT = ENVQT_BOOST_SUBDIV * 3 + 1;                  // 61
static const uint ENVQT_TOTAL = ENVQT_UNBOOSTED * ENVQT_BOOST_MULT;                                                             // 5368
static const uint IMPORTANCE_MAP_DIM = 1024;
static const uint LIGHTING_MAX_LIGHTS = 512 * 1024, LIGHTING_PROXY_RATIO = 12, LIGHTING_MAX_PROXIES_PER_LIGHT = 256 * 1024;
static const float LIGHTING_MIN_WEIGHT = 1e-8f;
static const uint RTXPT_INVALID_LIGHT_INDEX = 0xFFFFFFFFu;

struct PolymorphicLightInfo { float Center[3]; uint ColorTypeAndFlags; uint Direction1, Direction2, Scalars, LogRadiance; };
struct PolymorphicLightInfoEx { uint IesProfileIndex, PrimaryAxis, CosConeAngleAndSoftness, UniqueID; };      // PolymorphicLight.h:62-71

inline float UnpackRadiance(uint logRadiance)
{
    return (logRadiance == 0) ? 0 : exp2f((float(logRadiance - 1) / 65534.0f) * (kMaxLog2Radiance - kMinLog2Radiance) + kMinLog2Radiance);
}
inline float3 UnpackLightColor(const PolymorphicLightInfo& li) { return Unpack_R8G8B8_UFLOAT(li.ColorTypeAndFlags) * UnpackRadiance(li.LogRadiance & 0xffff); }
inline void PackLightColor(float3 radiance, PolymorphicLightInfo& li)
{
    float intensity = max3(radiance);
    if (intensity > 0.0f)
    {
        float logRadiance = saturate((log2f(intensity) - kMinLog2Radiance) / (kMaxLog2Radiance - kMinLog2Radiance));
        uint packedRadiance = std::min(uint(ceilf(logRadiance * 65534.0f)) + 1, 0xffffu);
        float unpackedRadiance = UnpackRadiance(packedRadiance);
        float3 normalizedRadiance = saturate(radiance / unpackedRadiance);
        li.LogRadiance |= packedRadiance;
        li.ColorTypeAndFlags |= Pack_R8G8B8_UFLOAT(normalizedRadiance);
    }
}
inline uint LightType(const PolymorphicLightInfo& li) { return (li.ColorTypeAndFlags >> kPolymorphicLightTypeShift) & 0xf; }

struct PolymorphicLightSample { float3 Position, Normal, Radiance; float SolidAnglePdf; bool LightSampleableByBSDF; };

struct TriangleLight
{
    float3 base, edge1, edge2, radiance, normal; float surfaceArea;
    static TriangleLight Create(const PolymorphicLightInfo& li)
    {
        TriangleLight t;
        t.edge1 = f3(f16tof32(li.Direction1 & 0xffff), f16tof32(li.Direction2 & 0xffff), f16tof32(li.Scalars & 0xffff));
        t.edge2 = f3(f16tof32(li.Direction1 >> 16), f16tof32(li.Direction2 >> 16), f16tof32(li.Scalars >> 16));
        t.base = f3(li.Center[0], li.Center[1], li.Center[2]) - ((t.edge1 + t.edge2) / 3.0f);
        t.radiance = UnpackLightColor(li);
        float3 n = cross(t.edge1, t.edge2);
        float len = length(n);
        if (len > 0.0f) { t.surfaceArea = 0.5f * len; t.normal = n / len; }
        else { t.surfaceArea = 0.0f; t.normal = f3(0); }
        return t;
    }
    PolymorphicLightInfo Store() const
    {
        PolymorphicLightInfo li = {};
        PackLightColor(radiance, li);
        float3 c = base + ((edge1 + edge2) / 3.0f);
        li.Center[0] = c.x; li.Center[1] = c.y; li.Center[2] = c.z;
        // PolymorphicLight.hlsli:510-513 keeps the three packed words in a `float3 edges` before they are stored: uint -> float -> uint, i.e. each word is rounded to 24
        // significant bits and the low bits of edge1's halves are lost (RTXDI's original has `uint3 edges`).  Reproduced as is: these are the records the reference's NEE samples
        // (pinned by tests/golden/lights_golden.npz, generated from the unmodified header)
        auto viaFloat = [](uint packed) { const float f = float(packed); return f >= 4294967296.0f ? 0xFFFFFFFFu : uint(f); };
        li.Direction1 = viaFloat((f32tof16(edge1.x) & 0xffff) | (f32tof16(edge2.x) << 16));
        li.Direction2 = viaFloat((f32tof16(edge1.y) & 0xffff) | (f32tof16(edge2.y) << 16));
        li.Scalars    = viaFloat((f32tof16(edge1.z) & 0xffff) | (f32tof16(edge2.z) << 16));
        li.ColorTypeAndFlags |= kLightTypeTriangle << kPolymorphicLightTypeShift;
        return li;
    }
    PolymorphicLightSample CalcSample(float2 random, float3 viewerPosition) const
    {
        PolymorphicLightSample r = {};
        float3 bary = SampleTriangleUniform(random);
        r.Position = base + edge1 * bary.y + edge2 * bary.z;
        r.Position = ComputeRayOrigin(r.Position, normal);
        r.Normal = normal;
        float3 toLight = r.Position - viewerPosition;
        float distSqr = std::max(2e-9f, dot(toLight, toLight));
        float distance = sqrtf(distSqr);
        float3 dir = toLight / distance;
        float cosTheta = dot(normal, -dir);
        r.SolidAnglePdf = 0.f; r.Radiance = f3(0);
        if (cosTheta <= 0.f) return r;
        float areaPdf = std::max(2e-9f, 1.0f / surfaceArea);
        r.SolidAnglePdf = std::min(1e10f, pdfAtoW(areaPdf, distance, cosTheta));
        r.Radiance = radiance;
        r.LightSampleableByBSDF = true;
        return r;
    }
    float CalcSolidAnglePdfForMIS(float3 viewerPosition, float3 lightSamplePosition) const
    {
        float3 toLight = lightSamplePosition - viewerPosition;
        float distSqr = std::max(2e-9f, dot(toLight, toLight));
        float distance = sqrtf(distSqr);
        float3 dir = toLight / distance;
        float cosTheta = dot(normal, -dir);
        float areaPdf = std::max(2e-9f, 1.0f / surfaceArea);
        return std::min(1e10f, pdfAtoW(areaPdf, distance, cosTheta));
    }
    float GetPower() const { return surfaceArea * K_PI * Luminance(radiance); }
};

struct EnvironmentQuadLight
{
    uint NodeX, NodeY, NodeDim; float Weight; float3 Radiance;
    static EnvironmentQuadLight Create(const PolymorphicLightInfo& li)
    {
        EnvironmentQuadLight e;
        e.NodeX = li.Direction1 >> 16; e.NodeY = li.Direction1 & 0xFFFF; e.NodeDim = li.Direction2 >> 16;
        e.Weight = asfloat(li.Scalars); e.Radiance = UnpackLightColor(li);
        return e;
    }
    PolymorphicLightInfo Store() const
    {
        PolymorphicLightInfo li = {};
        PackLightColor(Radiance, li);
        li.Direction1 = (NodeX << 16) | NodeY; li.Direction2 = NodeDim << 16; li.Scalars = asuint(Weight);
        li.ColorTypeAndFlags |= kLightTypeEnvironmentQuad << kPolymorphicLightTypeShift;
        return li;
    }
    float SolidAnglePdf() const { return float(NodeDim * NodeDim) / (4.0f * K_PI); }
};


// ---- analytic lights (sphere lights with optional spot shaping) -------------------------------------------------------------------------
// host-side packing helpers of Rtxpt/Lighting/LightsBaker.cpp:414-454 (note the truncating fp16 conversion and the doubled [0,1] mapping
// of the octahedral encoding, both reproduced on purpose)
inline uint fp32ToFp16Truncating(float v)
{
    const float multiple = asfloat(0x07800000u);        // 2^-112
    const uint u = asuint(v * multiple);
    const uint sign = u & 0x80000000u, body = u & 0x0fffffffu;
    return ((sign >> 16) | (body >> 13)) & 0xFFFFu;
}
inline float2 OctWrapHost(float2 v) { return f2((1.0f - fabsf(v.y)) * ((v.x >= 0.0f) ? 1.0f : -1.0f), (1.0f - fabsf(v.x)) * ((v.y >= 0.0f) ? 1.0f : -1.0f)); }
inline uint NDirToOctUnorm32(float3 n3)
{
    n3 = n3 / (fabsf(n3.x) + fabsf(n3.y) + fabsf(n3.z));
    float2 n = f2(n3.x, n3.y);
    n = n3.z >= 0.0f ? n : OctWrapHost(n);
    n = n * 0.5f + f2(0.5f, 0.5f);
    float2 p = n * 0.5f + f2(0.5f, 0.5f);
    p.x = saturate(p.x); p.y = saturate(p.y);
    return uint(p.x * float(0xfffe)) | (uint(p.y * float(0xfffe)) << 16);
}
inline float3 OctToNDirUnorm32(uint pUnorm)      // Utils.hlsli:128-153
{
    float2 p = f2(saturate(float(pUnorm & 0xffff) / float(0xfffe)), saturate(float(pUnorm >> 16) / float(0xfffe)));
    p = p * 2.0f - f2(1.0f, 1.0f);
    float2 f = p * 2.0f - f2(1.0f, 1.0f);            // Decode_Oct
    float3 n = f3(f.x, f.y, 1.0f - fabsf(f.x) - fabsf(f.y));
    float t = saturate(-n.z);
    n.x += (n.x >= 0.0f) ? -t : t; n.y += (n.y >= 0.0f) ? -t : t;
    return normalize(n);
}

struct LightShaping { float cosConeAngle = 0; float3 primaryAxis = f3(0); float cosConeSoftness = 0; bool isSpot = false; float minFalloff = 0; };
inline LightShaping unpackLightShaping(const PolymorphicLightInfo& base, const PolymorphicLightInfoEx& ex)     // LightShaping.hlsli:26-42
{
    LightShaping s;
    if (base.ColorTypeAndFlags & kPolymorphicLightShapingEnableBit)
    {
        s.isSpot = true; s.primaryAxis = OctToNDirUnorm32(ex.PrimaryAxis);
        s.cosConeAngle = f16tof32(ex.CosConeAngleAndSoftness & 0xffff); s.cosConeSoftness = f16tof32(ex.CosConeAngleAndSoftness >> 16);
        s.minFalloff = (base.ColorTypeAndFlags & kPolymorphicLightShapingUseMinFalloff) ? kMinSpotlightFalloff : 0.0f;
    }
    return s;
}
inline float smoothstepf(float a, float b, float x) { float t = saturate((x - a) / (b - a)); return t * t * (3.0f - 2.0f * t); }
inline float evaluateLightShaping(const LightShaping& s, float3 surfacePosition, float3 lightSamplePosition)   // LightShaping.hlsli:76-95
{
    if (!s.isSpot) return 1.0f;
    const float3 lightToSurface = normalize(surfacePosition - lightSamplePosition);
    const float cosTheta = dot(s.primaryAxis, lightToSurface);
    const float smoothFalloff = smoothstepf(s.cosConeAngle, s.cosConeAngle + s.cosConeSoftness, cosTheta);
    const float softSpotlight = std::max(s.minFalloff, smoothFalloff);
    return softSpotlight <= 0 ? 0.0f : softSpotlight;
}
inline float getShapingFluxFactor(const LightShaping& s)        // LightShaping.hlsli:161-174
{
    if (!s.isSpot) return 1.0f;
    float solidAngleOverTwoPi = (1.0f - s.cosConeAngle);
    solidAngleOverTwoPi *= (1.0f + (0.5f - 1.0f) * s.cosConeSoftness);      // lerp(1, 0.5, softness)
    return solidAngleOverTwoPi * 0.5f;
}
inline void BranchlessONB(float3 normal, float3& tangent, float3& bitangent)    // Utils/Geometry.hlsli:17-24
{
    float sign = (normal.z >= 0) ? 1.0f : -1.0f;
    float a = -1.0f / (sign + normal.z);
    float b = normal.x * normal.y * a;
    tangent = f3(1.0f + sign * normal.x * normal.x * a, sign * b, -sign * normal.x);
    bitangent = f3(b, sign + normal.y * normal.y * a, -normal.y);
}

struct SphereLight      // PolymorphicLight.hlsli:93-259
{
    float3 position, radiance; float radius; LightShaping shaping;
    static SphereLight Create(const PolymorphicLightInfo& li, const PolymorphicLightInfoEx& ex)
    {
        SphereLight s; s.position = f3(li.Center[0], li.Center[1], li.Center[2]); s.radius = f16tof32(li.Scalars & 0xffff);
        s.radiance = UnpackLightColor(li); s.shaping = unpackLightShaping(li, ex);
        return s;
    }
    PolymorphicLightSample CalcSample(float2 random, float3 viewerPosition) const
    {
        PolymorphicLightSample r = {};
        const float3 lightVector = position - viewerPosition;
        const float lightDistance2 = dot(lightVector, lightVector), radius2 = radius * radius;
        if (lightDistance2 < radius2)
        {   // inside the sphere: no emission (single-sided, like emissive triangles)
            r.Position = position; r.Normal = f3(0); r.Radiance = f3(0); r.SolidAnglePdf = 1.0f; r.LightSampleableByBSDF = false;
            return r;
        }
        const float lightDistance = sqrtf(lightDistance2);
        const float sinThetaMax2 = radius2 / lightDistance2;
        const float cosThetaMax = sqrtf(std::max(0.0f, 1.0f - sinThetaMax2));
        const float phi = 2.0f * K_PI * random.x;
        const float cosTheta = cosThetaMax + (1.0f - cosThetaMax) * random.y;         // lerp(cosThetaMax, 1, u.y)
        const float sinTheta = sqrtf(std::max(0.0f, 1.0f - cosTheta * cosTheta));
        const float sinTheta2 = sinTheta * sinTheta;
        const float dc = lightDistance, dc2 = lightDistance2;
        const float ds = dc * cosTheta - sqrtf(std::max(1e-10f, radius2 - dc2 * sinTheta2));
        const float cosAlpha = (dc2 + radius2 - ds * ds) / (2.0f * dc * radius);
        const float sinAlpha = sqrtf(std::max(0.0f, 1.0f - cosAlpha * cosAlpha));
        const float3 sampleSpaceNormal = normalize(lightVector);
        float3 T, B; BranchlessONB(sampleSpaceNormal, T, B);
        const float sinPhi = sinf(phi), cosPhi = cosf(phi);
        const float3 radiusVector = (-T) * (sinAlpha * cosPhi) + (-B) * (sinAlpha * sinPhi) + (-sampleSpaceNormal) * cosAlpha;
        r.Position = position + radiusVector * radius;
        r.Normal = normalize(radiusVector);
        r.Radiance = radiance;
        r.SolidAnglePdf = 1.0f / (2.0f * K_PI * (1.0f - cosThetaMax));
        r.LightSampleableByBSDF = false;
        return r;
    }
    // Eval (PolymorphicLight.hlsli:190-205) with IntersectRaySphere (Utils/Geometry.hlsli:85-118): what a BSDF ray sees when it reaches the light's proxy geometry
    bool Eval(float3 rayPos, float3 rayDir, float3& outRadiance, float3& outLightSamplePosition) const
    {
        const float3 lightVector = position - rayPos;
        if (dot(lightVector, lightVector) < radius * radius) return false;
        const float3 oc = rayPos - position;
        const float b = 2.0f * dot(oc, rayDir), c = dot(oc, oc) - radius * radius;
        const float discriminant = b * b - 4.0f * c;
        if (discriminant < 0.0f) return false;
        const float sqrtDisc = sqrtf(discriminant);
        const float t1 = (-b - sqrtDisc) / 2.0f, t2 = (-b + sqrtDisc) / 2.0f;
        const float t = (t1 >= 0.0f) ? t1 : ((t2 >= 0.0f) ? t2 : -1.0f);
        if (t < 0.0f) return false;
        outLightSamplePosition = rayPos + rayDir * t;
        outRadiance = radiance * evaluateLightShaping(shaping, rayPos, position);
        return true;
    }
    float CalcSolidAnglePdfForMIS(float3 viewerPosition) const
    {
        const float3 lightVector = position - viewerPosition;
        const float sinThetaMax2 = (radius * radius) / dot(lightVector, lightVector);
        const float cosThetaMax = sqrtf(std::max(0.0f, 1.0f - sinThetaMax2));
        return 1.0f / (2.0f * K_PI * (1.0f - cosThetaMax));
    }
    float GetPower() const { return (4 * K_PI * (radius * radius)) * K_PI * Luminance(radiance) * getShapingFluxFactor(shaping); }        // getSurfaceArea() = 4 pi sq(radius) first (PolymorphicLight.hlsli:224-232)
};

// Rtxpt/Lighting/LightsBaker.cpp:456-556 (ConvertLight)
inline void ConvertLight(const RtxptLightDesc& L, PolymorphicLightInfo& base, PolymorphicLightInfoEx& ex)
{
    base = PolymorphicLightInfo(); ex = PolymorphicLightInfoEx();
    const float3 color = f3(L.color[0], L.color[1], L.color[2]);
    const float kPi = 3.14159265358979323846f;
    auto radians = [&](float deg) { return deg * (kPi / 180.0f); };
    if (L.type == RTXPT_LIGHT_SPOT)
    {
        const float3 dir = normalize(f3(L.direction[0], L.direction[1], L.direction[2]));
        const uint minFalloff = (L.outerAngle < 0) ? kPolymorphicLightShapingUseMinFalloff : 0u;
        if (L.radius == 0.f)
        {
            base.ColorTypeAndFlags = (kLightTypePoint << kPolymorphicLightTypeShift) | minFalloff;
            PackLightColor(color * L.intensity, base);
            base.Direction1 = NDirToOctUnorm32(dir);
            base.Direction2 = fp32ToFp16Truncating(radians(fabsf(L.outerAngle))) | (fp32ToFp16Truncating(radians(L.innerAngle)) << 16);
        }
        else
        {
            const float projectedArea = kPi * (L.radius * L.radius);
            const float3 radiance = color * L.intensity / projectedArea;
            const float softness = saturate(1.f - L.innerAngle / fabsf(L.outerAngle));
            base.ColorTypeAndFlags = (kLightTypeSphere << kPolymorphicLightTypeShift) | minFalloff | kPolymorphicLightShapingEnableBit;
            PackLightColor(radiance, base);
            base.Scalars = fp32ToFp16Truncating(L.radius);
            if (fabsf(L.outerAngle) > 0)
            {
                ex.PrimaryAxis = NDirToOctUnorm32(dir);
                ex.CosConeAngleAndSoftness = fp32ToFp16Truncating(cosf(radians(fabsf(L.outerAngle)))) | (fp32ToFp16Truncating(softness) << 16);
            }
        }
    }
    else
    {
        if (L.radius == 0.f)
        {
            base.ColorTypeAndFlags = kLightTypePoint << kPolymorphicLightTypeShift;
            PackLightColor(color * L.intensity, base);
            base.Direction2 = fp32ToFp16Truncating(kPi) | (fp32ToFp16Truncating(0.0f) << 16);
        }
        else
        {
            const float projectedArea = kPi * (L.radius * L.radius);
            base.ColorTypeAndFlags = kLightTypeSphere << kPolymorphicLightTypeShift;
            PackLightColor(color * L.intensity / projectedArea, base);
            base.Scalars = fp32ToFp16Truncating(L.radius);
        }
    }
    base.Center[0] = L.position[0]; base.Center[1] = L.position[1]; base.Center[2] = L.position[2];
}

// ---------------------------------------------------------------------------------------------------------------------
// Bake
// ---------------------------------------------------------------------------------------------------------------------
struct LightTable
{
    std::vector<PolymorphicLightInfo> lights;
    std::vector<uint> proxyCounters, proxyIndices;
    std::vector<uint> envLookupMap;                         // IMPORTANCE_MAP_DIM^2 light indices, empty without an env map
    std::vector<std::vector<float>> radianceMips;           // RGBA (rgb radiance, a importance), fp16-rounded like the RGBA16F texture
    uint envQuadNodeCount = 0, triangleLightCount = 0, samplingProxyCount = 0;
    float weightsSum = 0;
    std::vector<float> weights;                             // per light, power-based (ComputeWeight), what ComputeProxyCounts blends the usage feedback into (pt_neeat.h)
    bool envEnabled = false;
    std::vector<PolymorphicLightInfoEx> lightsEx;      // analytic lights only: light index - ENVQT_TOTAL
    uint analyticLightCount = 0;
    uint exBase = ENVQT_TOTAL;                              // first light with an Extended record (the known-answer mirrors of oracle.cpp index a 16-light table from 0)
    PolymorphicLightInfoEx exOf(uint lightIndex) const { uint k = lightIndex - exBase; return k < analyticLightCount ? lightsEx[k] : PolymorphicLightInfoEx(); }
    uint importanceMipCount = 0;
    bool IsEmpty() const { return samplingProxyCount == 0; }
};

inline uint firstbithigh(uint v) { uint r = 0; while (v >>= 1) r++; return r; }

inline void buildEnvRadianceMap(const Scene& sc, LightTable& lt)
{
    const uint N = IMPORTANCE_MAP_DIM, S = 4;                                       // 16 samples per texel (EMISB_IMPORTANCE_SAMPLES_PER_PIXEL)
    lt.importanceMipCount = firstbithigh(N) + 1;
    lt.radianceMips.resize(lt.importanceMipCount);
    lt.radianceMips[0].resize(size_t(N) * N * 4);
    const float invSamples = 1.0f / float(S * S);
    #pragma omp parallel for schedule(dynamic, 8)
    for (int py = 0; py < int(N); py++)
        for (uint px = 0; px < N; px++)
        {
            float L = 0.f; float3 R = f3(0);
            for (uint y = 0; y < S; y++)
                for (uint x = 0; x < S; x++)
                {
                    float2 p = f2((float(px * S + x) + 0.5f) / float(N * S), (float(uint(py) * S + y) + 0.5f) / float(N * S));
                    float3 dir = oct_to_ndir_equal_area_unorm(p);
                    float3 radiance = sc.env.sampleLevel(dir, 0);
                    L += (Luminance(radiance) + Average(radiance)) * 0.5f;
                    R += radiance;
                }
            float* o = &lt.radianceMips[0][(size_t(py) * N + px) * 4];
            o[0] = lp(R.x * invSamples); o[1] = lp(R.y * invSamples); o[2] = lp(R.z * invSamples); o[3] = lp(L * invSamples);
        }
    for (uint m = 1; m < lt.importanceMipCount; m++)                                // donut MipMapGenPass MODE_COLOR: 2x2 box
    {
        uint n = N >> m, pn = N >> (m - 1);
        lt.radianceMips[m].resize(size_t(n) * n * 4);
        const std::vector<float>& src = lt.radianceMips[m - 1];
        for (uint y = 0; y < n; y++)
            for (uint x = 0; x < n; x++)
                for (uint c = 0; c < 4; c++)
                {
                    float s = src[(size_t(2 * y) * pn + 2 * x) * 4 + c] + src[(size_t(2 * y) * pn + 2 * x + 1) * 4 + c]
                            + src[(size_t(2 * y + 1) * pn + 2 * x) * 4 + c] + src[(size_t(2 * y + 1) * pn + 2 * x + 1) * 4 + c];
                    lt.radianceMips[m][(size_t(y) * n + x) * 4 + c] = lp(s * 0.25f);
                }
    }
}

struct EnvBakeConsts { float colorMultiplier[3]; float distantVsLocalRelativeImportance; const float* transform; };

inline const float* envTexel(const LightTable& lt, uint x, uint y, uint mip) { uint n = IMPORTANCE_MAP_DIM >> mip; return &lt.radianceMips[mip][(size_t(y) * n + x) * 4]; }

inline uint envWeightForQTBuild(const LightTable& lt, uint dim, uint x, uint y, uint lightIndex, uint depthLimit)
{
    uint mipLevel = lt.importanceMipCount - firstbithigh(dim) - 1;
    float areaMul = float(1u << (mipLevel * 2));
    float ret = areaMul * envTexel(lt, x, y, mipLevel)[3];
    ret = std::max(sq(1.0f / 100.0f) * float(mipLevel), ret);
    ret *= (mipLevel > depthLimit) ? 1.0f : 0.0f;
    return (std::min(uint(FastSqrt(ret) * 100 + 0.5f), 0x000FFFFFu) << 12) | lightIndex;
}
inline uint EQTNodePack(uint dim, uint x, uint y) { return (firstbithigh(dim) << 28) | (x << 14) | y; }
inline void EQTNodeUnpack(uint p, uint& dim, uint& x, uint& y) { dim = 1u << (p >> 28); x = (p >> 14) & 0x3FFF; y = p & 0x3FFF; }

inline void subdivide(const LightTable& lt, std::vector<uint>& nodes, std::vector<uint>& weights, uint nodeCount, uint subdivisions, uint depthLimit)
{
    for (uint si = 0; si < subdivisions; si++)
    {
        uint packed = 0;
        for (uint i = 0; i < nodeCount; i++) packed = std::max(packed, weights[i]);
        uint idx = packed & 0xFFF;
        uint dim, x, y; EQTNodeUnpack(nodes[idx], dim, x, y);
        for (uint k = 0; k < 4; k++)
        {
            uint nd = dim * 2, nx = x * 2 + (k % 2), ny = y * 2 + (k / 2);
            uint ni = (k == 0) ? idx : (nodeCount + k - 1);
            nodes[ni] = EQTNodePack(nd, nx, ny);
            weights[ni] = envWeightForQTBuild(lt, nd, nx, ny, ni, depthLimit);
        }
        nodeCount += 3;
    }
}

inline void bakeEnvLights(LightTable& lt, const EnvBakeConsts& c)
{
    std::vector<uint> nodes(ENVQT_UNBOOSTED), weights(ENVQT_UNBOOSTED);
    for (uint i = 0; i < ENVQT_BASE_RES * ENVQT_BASE_RES; i++)
    {
        uint x = i / ENVQT_BASE_RES, y = i % ENVQT_BASE_RES;
        nodes[i] = EQTNodePack(ENVQT_BASE_RES, x, y);
        weights[i] = envWeightForQTBuild(lt, ENVQT_BASE_RES, x, y, i, ENVQT_BOOST_DPT);
    }
    subdivide(lt, nodes, weights, ENVQT_BASE_RES * ENVQT_BASE_RES, ENVQT_SUBDIV, ENVQT_BOOST_DPT);
    const float avgMul = (c.colorMultiplier[0] + c.colorMultiplier[1] + c.colorMultiplier[2]) / 3.0f;
    for (uint g = 0; g < ENVQT_UNBOOSTED; g++)
    {
        std::vector<uint> bn(ENVQT_BOOST_MULT), bw(ENVQT_BOOST_MULT);
        uint dim, x, y; EQTNodeUnpack(nodes[g], dim, x, y);
        bn[0] = nodes[g]; bw[0] = envWeightForQTBuild(lt, dim, x, y, 0, 0);
        subdivide(lt, bn, bw, 1, ENVQT_BOOST_SUBDIV, 0);
        for (uint i = 0; i < ENVQT_BOOST_MULT; i++)
        {
            EnvironmentQuadLight e; EQTNodeUnpack(bn[i], e.NodeDim, e.NodeX, e.NodeY);
            uint mipLevel = lt.importanceMipCount - firstbithigh(e.NodeDim) - 1;
            float areaMul = float(1u << (mipLevel * 2));
            const float* v = envTexel(lt, e.NodeX, e.NodeY, mipLevel);
            e.Weight = areaMul * std::max(0.0f, v[3] * avgMul * c.distantVsLocalRelativeImportance);
            e.Radiance = f3(v[0] * c.colorMultiplier[0], v[1] * c.colorMultiplier[1], v[2] * c.colorMultiplier[2]);
            PolymorphicLightInfo li = e.Store();
            float3 localDir = oct_to_ndir_equal_area_unorm(f2((float(e.NodeX) + 0.5f) / float(e.NodeDim), (float(e.NodeY) + 0.5f) / float(e.NodeDim)));
            float3 worldDir = mul_vec_33of34(localDir, c.transform) * DISTANT_LIGHT_DISTANCE;
            li.Center[0] = worldDir.x; li.Center[1] = worldDir.y; li.Center[2] = worldDir.z;
            lt.lights[g * ENVQT_BOOST_MULT + i] = li;
        }
    }
    lt.envLookupMap.assign(size_t(IMPORTANCE_MAP_DIM) * IMPORTANCE_MAP_DIM, 0);
    for (uint li = 0; li < ENVQT_TOTAL; li++)
    {
        EnvironmentQuadLight e = EnvironmentQuadLight::Create(lt.lights[li]);
        uint ds = IMPORTANCE_MAP_DIM / e.NodeDim;
        for (uint yy = 0; yy < ds; yy++) for (uint xx = 0; xx < ds; xx++)
            lt.envLookupMap[size_t(e.NodeY * ds + yy) * IMPORTANCE_MAP_DIM + (e.NodeX * ds + xx)] = li;
    }
}

// emissive texture tap for the triangle bake: the reference uses one anisotropic SampleGrad (LightsBaker.hlsl:585-650); this
// restatement (and the product) takes one trilinear tap at the LOD of the longer gradient.
inline float3 sampleEmissiveForBake(const Scene& sc, uint textureIndexAndInfo, float2 uv, float2 gradA, float2 gradB)
{
    const Texture2D& t = sc.textures[textureIndexAndInfo & 0xFFFF];
    float2 dims = f2(float(t.d->width), float(t.d->height));
    float la = length(gradA * dims), lb = length(gradB * dims);
    float lod = log2f(std::max(std::max(la, lb), 1e-8f));
    return xyz(t.sampleLevel(uv, lod));
}

inline void bakeLights(Scene& sc, const RtxptPathTracerConstants& consts, LightTable& lt, bool rebuildEnvMaps = true)
{
    const RtxptSceneDesc& d = *sc.desc;
    lt.envEnabled = (d.envCube.faceSize != 0) && (consts.envMap.Enabled != 0.0f);
    lt.lights.assign(ENVQT_TOTAL, PolymorphicLightInfo());
    for (uint i = 0; i < ENVQT_TOTAL; i++) { lt.lights[i] = PolymorphicLightInfo(); lt.lights[i].ColorTypeAndFlags = kLightTypeEnvironmentQuad << kPolymorphicLightTypeShift; }
    lt.envQuadNodeCount = ENVQT_TOTAL;
    lt.envLookupMap.clear();
    if (lt.envEnabled)
    {
        if (rebuildEnvMaps || lt.radianceMips.empty()) buildEnvRadianceMap(sc, lt);
        EnvBakeConsts c; c.colorMultiplier[0] = consts.envMap.ColorMultiplier[0]; c.colorMultiplier[1] = consts.envMap.ColorMultiplier[1]; c.colorMultiplier[2] = consts.envMap.ColorMultiplier[2];
        c.distantVsLocalRelativeImportance = consts.distantVsLocalImportance * 0.0002f; c.transform = consts.envMap.Transform;
        bakeEnvLights(lt, c);
    }
    // analytic lights (LightsBaker.cpp:596-640): after the environment nodes, before the emissive triangles
    lt.lightsEx.clear(); lt.analyticLightCount = 0;
    for (uint i = 0; i < d.lightCount; i++)
    {
        PolymorphicLightInfo base; PolymorphicLightInfoEx ex; ConvertLight(d.lights[i], base, ex);
        lt.lights.push_back(base); lt.lightsEx.push_back(ex); lt.analyticLightCount++;
    }
    // emissive triangles, one light per triangle of every emissive geometry instance, in instance/geometry order
    lt.triangleLightCount = 0;
    for (uint ii = 0; ii < d.instanceCount; ii++)
    {
        const RtxptInstanceData& inst = d.instances[ii];
        for (uint gi = 0; gi < inst.numGeometries; gi++)
        {
            RtxptSubInstanceData& sub = sc.subInstances[inst.firstGeometryInstanceIndex + gi];
            const RtxptGeometryData& g = d.geometries[inst.firstGeometryIndex + gi];
            const RtxptMaterialData& m = d.materials[sub.GlobalGeometryIndex_PTMaterialDataIndex & 0xFFFF];
            uint triCount = g.numIndices / 3;
            bool emissive = (m.EmissiveColor[0] > 0 || m.EmissiveColor[1] > 0 || m.EmissiveColor[2] > 0);
            bool overflow = lt.lights.size() + triCount >= LIGHTING_MAX_LIGHTS;
            if (!emissive || overflow) { sub.EmissiveLightMappingOffset = 0xFFFFFFFFu; continue; }
            sub.EmissiveLightMappingOffset = uint(lt.lights.size());
            const float* xf = inst.transform;
            float det = xf[0] * (xf[5] * xf[10] - xf[6] * xf[9]) - xf[1] * (xf[4] * xf[10] - xf[6] * xf[8]) + xf[2] * (xf[4] * xf[9] - xf[5] * xf[8]);
            bool isFlipped = det < 0.f;
            for (uint t = 0; t < triCount; t++)
            {
                uint idx[3]; float3 p[3];
                for (int k = 0; k < 3; k++) { idx[k] = sc.load32(g.indexBufferIndex, g.indexOffset + t * 12 + k * 4); p[k] = mul34_point(xf, sc.loadFloat3(g.vertexBufferIndex, g.positionOffset + idx[k] * 12)); }
                float3 radiance = f3(m.EmissiveColor[0], m.EmissiveColor[1], m.EmissiveColor[2]);
                if (m.EmissiveTextureIndex != 0xFFFFFFFFu && g.texCoord1Offset != ~0u && (m.Flags & RTXPT_MATFLAG_UseEmissiveTexture))
                {
                    float2 uvs[3]; for (int k = 0; k < 3; k++) uvs[k] = sc.loadFloat2(g.vertexBufferIndex, g.texCoord1Offset + idx[k] * 8);
                    float2 e[3] = { uvs[1] - uvs[0], uvs[2] - uvs[1], uvs[0] - uvs[2] };
                    float l[3] = { length(e[0]), length(e[1]), length(e[2]) };
                    float2 shortEdge, long1, long2;
                    if (l[0] < l[1] && l[0] < l[2]) { shortEdge = e[0]; long1 = e[1]; long2 = e[2]; }
                    else if (l[1] < l[2]) { shortEdge = e[1]; long1 = e[2]; long2 = e[0]; }
                    else { shortEdge = e[2]; long1 = e[0]; long2 = e[1]; }
                    float2 shortGradient = shortEdge * (2.0f / 3.0f), longGradient = (long1 + long2) / 3.0f;
                    float2 centerUV = (uvs[0] + uvs[1] + uvs[2]) / 3.0f;
                    radiance *= sampleEmissiveForBake(sc, m.EmissiveTextureIndex, centerUV, shortGradient, longGradient);
                }
                radiance = max3v(radiance, f3(0));
                TriangleLight tl; tl.base = p[0];
                if (!isFlipped) { tl.edge1 = p[1] - p[0]; tl.edge2 = p[2] - p[0]; } else { tl.edge1 = p[2] - p[0]; tl.edge2 = p[1] - p[0]; }
                if (max3(radiance) < 1e-7f) radiance = f3(0);
                tl.radiance = radiance;
                lt.lights.push_back(tl.Store());
                lt.triangleLightCount++;
            }
        }
    }
    // weights (ComputeWeight: pow(flux, 0.8), threshold), deterministic sum order
    const uint n = uint(lt.lights.size());
    std::vector<float> w(n);
    for (uint i = 0; i < n; i++)
    {
        const PolymorphicLightInfo& li = lt.lights[i];
        float flux = 0;
        if (LightType(li) == kLightTypeTriangle) flux = TriangleLight::Create(li).GetPower();
        else if (LightType(li) == kLightTypeEnvironmentQuad) flux = asfloat(li.Scalars);
        else if (LightType(li) == kLightTypeSphere) flux = SphereLight::Create(li, lt.exOf(i)).GetPower();
        float weight = powf(flux, 0.8f);
        if (weight < LIGHTING_MIN_WEIGHT) weight = 0;
        w[i] = weight;
    }
    float total = 0;
    for (uint g0 = 0; g0 < n; g0 += 32 * 128)
    {
        float groupSum = 0;
        for (uint b0 = g0; b0 < std::min(n, g0 + 32 * 128); b0 += 32)
        {
            float blockSum = 0;
            for (uint i = b0; i < std::min(n, b0 + 32); i++) blockSum += w[i];
            groupSum += blockSum;
        }
        total += groupSum;
    }
    lt.weightsSum = total; lt.weights = w;
    // proxy counts + table (ComputeProxyCounts / ExecuteProxyJobs)
    const uint budget = LIGHTING_PROXY_RATIO * std::max(n, LIGHTING_MAX_LIGHTS / 10);
    lt.proxyCounters.assign(n, 0);
    lt.proxyIndices.clear();
    for (uint i = 0; i < n; i++)
    {
        uint proxies = 0;
        if (w[i] > 0) proxies = (consts.NEEType == 0) ? 1u : uint(ceilf((float(budget - n) * w[i]) / total));
        proxies = std::min(proxies, LIGHTING_MAX_PROXIES_PER_LIGHT - 1);
        lt.proxyCounters[i] = proxies;
        lt.proxyIndices.insert(lt.proxyIndices.end(), proxies, i);
    }
    lt.samplingProxyCount = uint(lt.proxyIndices.size());
}

} // namespace orc
