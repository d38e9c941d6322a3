// ORACLE — test infrastructure only (see pt_math.h).
// pt_scene.h: scene access layer restated from
//   Rtxpt/Shaders/PathTracerBridgeDonut.hlsli:152-256 (getGeometryFromHit), :269-278 (sampleTexture), :280-309 (normal map),
//   :311-392 (EvaluateSceneMaterialRTXPT), :394-428 (sampleGeometryMaterialRTXPT), :612-853 (Bridge::loadSurface),
//   :871-887 (loadHomogeneousVolumeData), :929-989 (alpha test)
//   Rtxpt/Shaders/PathTracer/Scene/Material/ShadingUtils.hlsli:110-166, Scene/Material/TextureSampler.hlsli (ray-cone LOD sampler),
//   Rtxpt/Shaders/PathTracer/Rendering/Materials/TexLODHelpers.hlsli:41-161, Rtxpt/Shaders/PathTracer/PathTracerHelpers.hlsli:29-42
// Texture filtering is done in software here (the reference uses the TMU): trilinear, wrap addressing, filter weights quantised to
// 8 fractional bits like the hardware (CUDA programming guide, "Texture Fetching").
#pragma once
#include "pt_math.h"
#include "pt_bsdf.h"
#include "../include/rtxpt_b200.h"
#include <vector>

namespace orc {

// ---------------------------------------------------------------------------------------------------------------------
// Textures
// ---------------------------------------------------------------------------------------------------------------------
inline float srgbToLinear8(uint8_t v)
{
    float c = float(v) / 255.0f;
    return (c <= 0.04045f) ? c / 12.92f : powf((c + 0.055f) / 1.055f, 2.4f);
}
struct SrgbLut { float t[256]; SrgbLut() { for (int i = 0; i < 256; i++) t[i] = srgbToLinear8(uint8_t(i)); } };
inline const float* srgbLut() { static SrgbLut l; return l.t; }

inline float quant8(float a) { return floorf(a * 256.0f + 0.5f) * (1.0f / 256.0f); }

struct Texture2D
{
    const RtxptTextureDesc* d = nullptr;
    float4 texel(uint level, int x, int y) const
    {
        uint w = std::max(1u, d->width >> level), h = std::max(1u, d->height >> level);
        x = ((x % int(w)) + int(w)) % int(w); y = ((y % int(h)) + int(h)) % int(h);     // wrap
        size_t idx = size_t(y) * w + size_t(x);
        if (d->format == RTXPT_FORMAT_RGBA32_FLOAT) { const float* p = (const float*)d->mips[level] + idx * 4; return f4(p[0], p[1], p[2], p[3]); }
        const uint8_t* p = (const uint8_t*)d->mips[level] + idx * 4;
        if (d->format == RTXPT_FORMAT_RGBA8_SRGB) { const float* l = srgbLut(); return f4(l[p[0]], l[p[1]], l[p[2]], float(p[3]) / 255.0f); }
        return f4(float(p[0]) / 255.0f, float(p[1]) / 255.0f, float(p[2]) / 255.0f, float(p[3]) / 255.0f);
    }
    float4 bilinear(uint level, float2 uv) const
    {
        uint w = std::max(1u, d->width >> level), h = std::max(1u, d->height >> level);
        float x = uv.x * float(w) - 0.5f, y = uv.y * float(h) - 0.5f;
        float fx = floorf(x), fy = floorf(y);
        float ax = quant8(x - fx), ay = quant8(y - fy);
        int ix = int(fx), iy = int(fy);
        float4 t00 = texel(level, ix, iy), t10 = texel(level, ix + 1, iy), t01 = texel(level, ix, iy + 1), t11 = texel(level, ix + 1, iy + 1);
        float4 a = t00 * (1 - ax) + t10 * ax;
        float4 b = t01 * (1 - ax) + t11 * ax;
        return a * (1 - ay) + b * ay;
    }
    float4 sampleLevel(float2 uv, float lod) const      // Texture2D.SampleLevel with a trilinear wrap sampler
    {
        float maxLod = float(d->mipLevels - 1);
        lod = clampf(lod, 0.0f, maxLod);
        float fl = floorf(lod);
        float a = quant8(lod - fl);
        uint l0 = uint(fl), l1 = std::min(l0 + 1, d->mipLevels - 1);
        float4 c0 = bilinear(l0, uv);
        if (a == 0.0f || l1 == l0) return c0;
        float4 c1 = bilinear(l1, uv);
        return c0 * (1 - a) + c1 * a;
    }
};

// Environment cube: D3D face order/orientation; bilinear inside the face with clamp (non-seamless), integer mip.
struct EnvCube
{
    const RtxptEnvCubeDesc* d = nullptr;
    static void dirToFace(float3 v, uint& face, float2& uv)
    {
        float ax = fabsf(v.x), ay = fabsf(v.y), az = fabsf(v.z);
        float m, s, t;
        if (ax >= ay && ax >= az) { m = ax; if (v.x >= 0) { face = 0; s = -v.z; t = -v.y; } else { face = 1; s = v.z; t = -v.y; } }
        else if (ay >= az)        { m = ay; if (v.y >= 0) { face = 2; s = v.x; t = v.z; }  else { face = 3; s = v.x; t = -v.z; } }
        else                      { m = az; if (v.z >= 0) { face = 4; s = v.x; t = -v.y; } else { face = 5; s = -v.x; t = -v.y; } }
        uv = f2((s / m + 1.0f) * 0.5f, (t / m + 1.0f) * 0.5f);
    }
    float3 texel(uint face, uint level, int x, int y) const
    {
        int n = int(std::max(1u, d->faceSize >> level));
        x = std::min(std::max(x, 0), n - 1); y = std::min(std::max(y, 0), n - 1);
        const float* p = d->faces[face][level] + (size_t(y) * n + x) * 4;
        return f3(p[0], p[1], p[2]);
    }
    float3 sampleLevel(float3 dir, float lod) const
    {
        if (!d || d->faceSize == 0) return f3(0);
        uint level = uint(clampf(floorf(lod + 0.5f), 0.0f, float(d->mipLevels - 1)));
        uint face; float2 uv; dirToFace(dir, face, uv);
        float n = float(std::max(1u, d->faceSize >> level));
        float x = uv.x * n - 0.5f, y = uv.y * n - 0.5f;
        float fx = floorf(x), fy = floorf(y);
        float ax = quant8(x - fx), ay = quant8(y - fy);
        int ix = int(fx), iy = int(fy);
        float3 a = texel(face, level, ix, iy) * (1 - ax) + texel(face, level, ix + 1, iy) * ax;
        float3 b = texel(face, level, ix, iy + 1) * (1 - ax) + texel(face, level, ix + 1, iy + 1) * ax;
        return a * (1 - ay) + b * ay;
    }
};

// ---------------------------------------------------------------------------------------------------------------------
// Ray cone (TexLODHelpers.hlsli:41-120; USE_RAYCONES_WITH_FP16_IN_RAYPAYLOAD)
// ---------------------------------------------------------------------------------------------------------------------
inline float SafeLog2(float x) { return log2f(clampf(x, FLT_MIN_, FLT_MAX_)); }
struct RayCone
{
    uint widthSpreadAngleFP16 = 0;
    float getWidth() const { return f16tof32(widthSpreadAngleFP16 >> 16); }
    float getSpreadAngle() const { return f16tof32(widthSpreadAngleFP16 & 0xFFFF); }
    static RayCone make(float width, float angle) { RayCone r; r.widthSpreadAngleFP16 = (f32tof16(width) << 16) | f32tof16(angle); return r; }
    RayCone propagateDistance(float hitT) const { float a = getSpreadAngle(), w = getWidth(); return make(a * hitT + w, a); }
    float computeLOD(float triLODConstant, float3 rayDir, float3 normal, bool moreDetailOnSlopes) const
    {
        float lambda = triLODConstant;
        float distTerm = fabsf(getWidth());
        float normalTerm = fabsf(dot(rayDir, normal));
        if (moreDetailOnSlopes) normalTerm = sqrtf(normalTerm);
        lambda += SafeLog2(distTerm / normalTerm);
        return lambda;
    }
};

// PathTracerHelpers.hlsli:29-42 (Wächter & Binder integer-offset origin)
inline float3 ComputeRayOrigin(float3 worldPosition, float3 faceNormal)
{
    const float origin = 1.f / 16.f, fScale = 3.f / 65536.f, iScale = 3 * 256.f;
    int iOff[3] = { int(faceNormal.x * iScale), int(faceNormal.y * iScale), int(faceNormal.z * iScale) };
    float p[3] = { worldPosition.x, worldPosition.y, worldPosition.z };
    float n[3] = { faceNormal.x, faceNormal.y, faceNormal.z };
    float r[3];
    for (int i = 0; i < 3; i++)
    {
        float iPos = asfloat_i(asint(p[i]) + ((p[i] < 0.f) ? -iOff[i] : iOff[i]));
        float fOff = n[i] * fScale;
        r[i] = (fabsf(p[i]) < origin) ? p[i] + fOff : iPos;
    }
    return f3(r[0], r[1], r[2]);
}

// ---------------------------------------------------------------------------------------------------------------------
// Shading data
// ---------------------------------------------------------------------------------------------------------------------
struct ShadingData      // Scene/ShadingData.hlsli:38-61
{
    float3 posW, faceNCorrected, V, N, T, B, vertexN;
    bool frontFacing;
    uint nestedPriority; bool thinSurface; uint activeLobes; bool psdExclude;   // MaterialHeader (Scene/Material/MaterialData.hlsli)
    uint psdDominantDeltaLobeP1 = 0; bool psdBlockMotionVectorsAtSurface = false;
    uint materialID;
    float IoR, shadowNoLFadeout;
    float3 emission;
    float3 computeNewRayOrigin(bool viewside) const { return ComputeRayOrigin(posW, viewside ? faceNCorrected : -faceNCorrected); }
    BSDFFrame frame() const { BSDFFrame f; f.T = T; f.B = B; f.N = N; f.V = V; f.thinSurface = thinSurface; f.psdExclude = psdExclude; f.activeLobes = activeLobes; return f; }
};

struct SurfaceData      // PathTracerTypes.hlsli:52-94
{
    ShadingData sd;
    StandardBSDF bsdf;
    float interiorIoR;
    uint neeTriangleLightIndex;
    uint neeAnalyticLightIndex;     // light the hit geometry stands in for (PTMaterialFlags_EnableAsAnalyticLightProxy), else 0xFFFFFFFF
    float3 prevPosW;                // instance.prevTransform x prevObjectSpacePosition (BridgeDonut:631); read by the BUILD pass's motion vectors only
};

struct Scene
{
    const RtxptSceneDesc* desc = nullptr;
    std::vector<Texture2D> textures;
    EnvCube env;
    std::vector<RtxptSubInstanceData> subInstances;    // copy: EmissiveLightMappingOffset is filled by the light bake

    void init(const RtxptSceneDesc* d)
    {
        desc = d;
        textures.resize(d->textureCount);
        for (uint i = 0; i < d->textureCount; i++) textures[i].d = &d->textures[i];
        env.d = &d->envCube;
        subInstances.assign(d->subInstances, d->subInstances + d->subInstanceCount);
        // AnalyticProxyLightIndex arrives as an index into desc->lights; the light list puts analytic lights after the 5368 environment nodes
        for (RtxptSubInstanceData& si : subInstances) si.AnalyticProxyLightIndex = (si.AnalyticProxyLightIndex < d->lightCount) ? si.AnalyticProxyLightIndex + 5368u : 0xFFFFFFFFu;
    }
    uint load32(int buffer, uint byteOffset) const { uint v; memcpy(&v, (const uint8_t*)desc->buffers[buffer].data + byteOffset, 4); return v; }
    float3 loadFloat3(int buffer, uint byteOffset) const { float v[3]; memcpy(v, (const uint8_t*)desc->buffers[buffer].data + byteOffset, 12); return f3(v[0], v[1], v[2]); }
    float2 loadFloat2(int buffer, uint byteOffset) const { float v[2]; memcpy(v, (const uint8_t*)desc->buffers[buffer].data + byteOffset, 8); return f2(v[0], v[1]); }
};

inline float3 SafeNormalize(float3 v) { float l2 = dot(v, v); return v * (1.0f / sqrtf(std::max(1.175494351e-38f, l2))); }
inline float3 FlipIfOpposite(float3 n, float3 ref) { return (dot(n, ref) >= 0) ? n : -n; }

struct GeometrySample   // DonutGeometrySample, PathTracerBridgeDonut.hlsli:57-80 (only what the reference-mode path reads)
{
    const RtxptInstanceData* instance; const RtxptGeometryData* geometry;
    float3 vertexPositions[3]; float2 vertexTexcoords[3];
    float3 objectSpacePosition; float2 texcoord;
    float3 flatNormal, geometryNormal; float4 tangent; bool frontFacing;
    float3 prevObjectSpacePosition;     // GeomAttr_PrevPosition: last frame's position where the geometry carries a previous-position stream (skinned meshes), else the current one
    float curvatureWS;                  // TriangleCurvatureApprox_GradN
};

// PathTracerBridgeDonut.hlsli:104-150: RMS magnitude of the gradient of the (object-space, unit) vertex normals over the world-space triangle, 1 / position units
inline float TriangleCurvatureApprox_GradN(const float3 vertexPositions[3], const float3 vertexNormals[3], const float* xf)
{
    const float eps = 1e-8f;
    float3 e10 = mul34_vec(xf, vertexPositions[1] - vertexPositions[0]);
    float e10Len = length(e10);
    if (e10Len < eps) return 0.0f;
    float3 e1 = e10 / e10Len;
    float3 e20 = mul34_vec(xf, vertexPositions[2] - vertexPositions[0]);
    float u2 = dot(e20, e1);
    float3 t = e20 - e1 * u2;
    float tLen = length(t);
    if (tLen < eps) return 0.0f;
    float3 e2 = t / tLen;
    float u1 = e10Len, v2 = dot(e20, e2);
    float3 dn1 = vertexNormals[1] - vertexNormals[0], dn2 = vertexNormals[2] - vertexNormals[0];
    float3 a = dn1 / std::max(u1, eps);
    float denomV = fabsf(v2) < eps ? (v2 >= 0.0f ? eps : -eps) : v2;
    float3 b = (dn2 - a * u2) / denomV;
    return sqrtf(dot(a, a) + dot(b, b));
}
// Libraries/MicroRng.hlsli:12-58
struct MicroRng
{
    uint N;
    static MicroRng make(uint x, uint y, uint seedValueA, uint seedValueB)
    {
        MicroRng r; r.N = ((x << 16) | y) ^ 0x9e3779b9u;
        r.N = r.N ^ (seedValueA + (r.N << 6) + (r.N >> 2));
        r.N = r.N ^ (seedValueB + (r.N << 6) + (r.N >> 2));
        return r;
    }
    uint Next() { N ^= N >> 16; N *= 0x21f0aaadu; N ^= N >> 15; N *= 0xf35a2d97u; N ^= N >> 15; return N; }
    float NextFloat() { return float(Next() >> 8) / 16777216.0f; }
};

inline GeometrySample getGeometryFromHit(const Scene& sc, uint instanceIndex, uint geometryIndex, uint triangleIndex, float2 rayBary, float3 rayDirection)
{
    GeometrySample gs = {};
    gs.instance = &sc.desc->instances[instanceIndex];
    gs.geometry = &sc.desc->geometries[gs.instance->firstGeometryIndex + geometryIndex];
    const RtxptGeometryData& g = *gs.geometry;
    float3 bary = f3(1.0f - (rayBary.x + rayBary.y), rayBary.x, rayBary.y);
    uint idx[3];
    for (int k = 0; k < 3; k++) idx[k] = sc.load32(g.indexBufferIndex, g.indexOffset + triangleIndex * 12 + k * 4);
    for (int k = 0; k < 3; k++) gs.vertexPositions[k] = sc.loadFloat3(g.vertexBufferIndex, g.positionOffset + idx[k] * 12);
    gs.objectSpacePosition = gs.vertexPositions[0] * bary.x + gs.vertexPositions[1] * bary.y + gs.vertexPositions[2] * bary.z;
    gs.prevObjectSpacePosition = gs.objectSpacePosition;
    if (g.prevPositionOffset != ~0u)        // only present for skinned objects (BridgeDonut:187-199)
    {
        float3 pv[3]; for (int k = 0; k < 3; k++) pv[k] = sc.loadFloat3(g.vertexBufferIndex, g.prevPositionOffset + idx[k] * 12);
        gs.prevObjectSpacePosition = pv[0] * bary.x + pv[1] * bary.y + pv[2] * bary.z;
    }
    if (g.texCoord1Offset != ~0u)
    {
        for (int k = 0; k < 3; k++) gs.vertexTexcoords[k] = sc.loadFloat2(g.vertexBufferIndex, g.texCoord1Offset + idx[k] * 8);
        gs.texcoord = gs.vertexTexcoords[0] * bary.x + gs.vertexTexcoords[1] * bary.y + gs.vertexTexcoords[2] * bary.z;
    }
    float3 vBA = gs.vertexPositions[1] - gs.vertexPositions[0], vCA = gs.vertexPositions[2] - gs.vertexPositions[0];
    float3 objectSpaceFlatNormal = SafeNormalize(cross(vBA, vCA));
    const float* xf = gs.instance->transform;
    if (g.normalOffset != ~0u)
    {
        float3 n[3];
        for (int k = 0; k < 3; k++)
        {
            n[k] = normalize(Unpack_RGB8_SNORM(sc.load32(g.vertexBufferIndex, g.normalOffset + idx[k] * 4)));
            n[k] = FlipIfOpposite(n[k], objectSpaceFlatNormal);
        }
        gs.curvatureWS = TriangleCurvatureApprox_GradN(gs.vertexPositions, n, xf);
        gs.geometryNormal = n[0] * bary.x + n[1] * bary.y + n[2] * bary.z;
        gs.geometryNormal = SafeNormalize(mul34_vec(xf, gs.geometryNormal));
    }
    if (g.tangentOffset != ~0u)
    {
        float4 t[3];
        for (int k = 0; k < 3; k++) t[k] = Unpack_RGBA8_SNORM(sc.load32(g.vertexBufferIndex, g.tangentOffset + idx[k] * 4));
        float3 txyz = xyz(t[0]) * bary.x + xyz(t[1]) * bary.y + xyz(t[2]) * bary.z;
        txyz = SafeNormalize(mul34_vec(xf, txyz));
        gs.tangent = f4(txyz, t[0].w);
    }
    gs.flatNormal = SafeNormalize(mul34_vec(xf, objectSpaceFlatNormal));
    gs.frontFacing = dot(-rayDirection, gs.flatNormal) >= 0.0f;
    return gs;
}

// TexLODHelpers.hlsli:129-142 ; worldMat passed as transpose((float3x3)transform) and used as mul(v, worldMat) == mul((float3x3)transform, v)
inline float computeRayConeTriangleLODValue(const float3 v[3], const float2 tx[3], const float* xf)
{
    float2 tx10 = tx[1] - tx[0], tx20 = tx[2] - tx[0];
    float Ta = fabsf(tx10.x * tx20.y - tx20.x * tx10.y);
    float3 edge01 = mul34_vec(xf, v[1] - v[0]), edge02 = mul34_vec(xf, v[2] - v[0]);
    float Pa = length(cross(edge01, edge02));
    return 0.5f * SafeLog2(Ta / Pa);
}

// ExplicitRayConesLodTextureSampler::sampleTexture + Bridge sampleTexture (TextureSampler.hlsli, BridgeDonut:269-278)
inline float4 sampleMaterialTexture(const Scene& sc, uint textureIndexAndInfo, float rayconesLODWithoutTexDims, float2 uv)
{
    uint textureIndex = textureIndexAndInfo & 0xFFFF, baseLOD = textureIndexAndInfo >> 24, mipLevels = (textureIndexAndInfo >> 16) & 0xFF;
    float lambda = 0.5f * float(baseLOD) + rayconesLODWithoutTexDims;
    lambda = std::min(lambda, std::max(float(mipLevels) - 5.0f, 0.0f));
    return sc.textures[textureIndex].sampleLevel(uv, lambda);
}

// ShadingUtils.hlsli:110-137
inline void computeTangentSpace(ShadingData& sd, float4 tangentW, bool ignoreTangent)
{
    float3 t = xyz(tangentW);
    float NdotT = dot(t, sd.N);
    bool nonParallel = fabsf(NdotT) < 0.9999f;
    bool nonZero = dot(t, t) > 0.f;
    bool valid = tangentW.w != 0.f && nonZero && nonParallel;
    if (!ignoreTangent && valid) { sd.T = normalize(t - sd.N * NdotT); sd.B = cross(sd.N, sd.T) * tangentW.w; }
    else { sd.T = perp_stark(sd.N); sd.B = cross(sd.N, sd.T); }
}
// ShadingUtils.hlsli:144-166
inline void adjustShadingNormal(ShadingData& sd, float4 tangentW, bool recomputeTangentSpace, bool ignoreTangent)
{
    float3 Ng = sd.faceNCorrected;
    float signN = dot(sd.N, Ng) >= 0.f ? 1.f : -1.f;
    float3 Ns = signN * sd.N;
    const float kCosThetaThreshold = 0.1f;
    float cosTheta = dot(sd.V, Ns);
    if (cosTheta <= kCosThetaThreshold)
    {
        float t = saturate(cosTheta * (1.f / kCosThetaThreshold));
        sd.N = signN * normalize(lerp(Ng, Ns, t));
    }
    if (cosTheta <= kCosThetaThreshold || recomputeTangentSpace) computeTangentSpace(sd, tangentW, ignoreTangent);
}

struct MaterialProperties   // PathTracer/Materials/MaterialTypes.hlsli:17-58 ; lpfloat fields carry fp16-rounded values
{
    float3 shadingNormal, geometryNormal, emissiveColor, baseColor;
    float opacity, roughness, metalness, transmission, diffuseTransmission, ior, shadowNoLFadeout;
    uint flags;
};

inline float GetPerceivedBrightness(float3 c) { return sqrtf(0.299f * c.x * c.x + 0.587f * c.y * c.y + 0.114f * c.z * c.z); }
inline float SolveMetalness(float diffuse, float specular, float oneMinusSpecularStrength)
{
    const float ds = 0.04f;
    if (specular < ds) return 0;
    float a = ds, b = diffuse * oneMinusSpecularStrength / (1 - ds) + specular - 2 * ds, c = ds - specular;
    float D = std::max(b * b - 4 * a * c, 0.f);
    return clampf((-b + sqrtf(D)) / (2 * a), 0, 1);
}

inline MaterialProperties sampleGeometryMaterial(const Scene& sc, const GeometrySample& gs, uint materialIndex, float lodNoDims)
{
    const RtxptMaterialData& m = sc.desc->materials[materialIndex];
    float4 texBase = f4(1, 1, 1, 1), texMR = f4(1, 1, 1, 1), texEmissive = f4(1, 1, 1, 1), texNormal = f4(0.5f, 0.5f, 1.0f, 0.0f), texTrans = f4(1, 1, 1, 1);
    if (m.Flags & RTXPT_MATFLAG_UseBaseOrDiffuseTexture) texBase = sampleMaterialTexture(sc, m.BaseOrDiffuseTextureIndex, lodNoDims, gs.texcoord);
    if (m.Flags & RTXPT_MATFLAG_UseEmissiveTexture) texEmissive = sampleMaterialTexture(sc, m.EmissiveTextureIndex, lodNoDims, gs.texcoord);
    if (m.Flags & RTXPT_MATFLAG_UseNormalTexture) texNormal = sampleMaterialTexture(sc, m.NormalTextureIndex, lodNoDims, gs.texcoord);
    if (m.Flags & RTXPT_MATFLAG_UseMetalRoughOrSpecularTexture) texMR = sampleMaterialTexture(sc, m.MetalRoughOrSpecularTextureIndex, lodNoDims, gs.texcoord);
    if (m.Flags & RTXPT_MATFLAG_UseTransmissionTexture) texTrans = sampleMaterialTexture(sc, m.TransmissionTextureIndex, lodNoDims, gs.texcoord);

    MaterialProperties r = {};
    r.opacity = 1; r.ior = 1.5f;
    r.geometryNormal = normalize(gs.geometryNormal);
    r.shadingNormal = r.geometryNormal;
    r.flags = m.Flags;
    float3 base = f3(m.BaseOrDiffuseColor[0], m.BaseOrDiffuseColor[1], m.BaseOrDiffuseColor[2]);
    if (m.Flags & RTXPT_MATFLAG_UseSpecularGlossModel)
    {
        float3 diffuseColor = base * xyz(texBase);
        float3 specularColor = f3(m.SpecularColor[0], m.SpecularColor[1], m.SpecularColor[2]) * xyz(texMR);
        r.roughness = lp(1.0f - texMR.w * (1.0f - m.Roughness));
        const float epsilon = 1e-6f;
        float oneMinusSpecularStrength = 1.0f - max3(specularColor);
        float metalness = SolveMetalness(GetPerceivedBrightness(diffuseColor), GetPerceivedBrightness(specularColor), oneMinusSpecularStrength);
        float3 fromDiffuse = diffuseColor * (oneMinusSpecularStrength / (1 - 0.04f) / std::max(1 - metalness, epsilon));
        float3 fromSpecular = specularColor - f3(0.04f * (1 - metalness) / std::max(metalness, epsilon));
        r.baseColor = lp(saturate(lerp(fromDiffuse, fromSpecular, metalness * metalness)));
        r.metalness = lp(metalness);
    }
    else
    {
        r.baseColor = lp(base * xyz(texBase));
        r.roughness = lp(m.Roughness * texMR.y);
        r.metalness = lp(m.Metalness * ((m.Flags & RTXPT_MATFLAG_MetalnessInRedChannel) ? texMR.x : texMR.z));
    }
    r.opacity = lp(m.Opacity);
    if (m.Flags & RTXPT_MATFLAG_UseBaseOrDiffuseTexture) r.opacity = lp(r.opacity * lp(texBase.w));
    r.opacity = saturate(r.opacity);
    r.transmission = lp(m.TransmissionFactor);
    r.diffuseTransmission = lp(m.DiffuseTransmissionFactor);
    if (m.Flags & RTXPT_MATFLAG_UseTransmissionTexture) { r.transmission = lp(r.transmission * lp(texTrans.x)); r.diffuseTransmission = lp(r.diffuseTransmission * lp(texTrans.x)); }
    r.emissiveColor = lp(f3(m.EmissiveColor[0], m.EmissiveColor[1], m.EmissiveColor[2]));
    if (m.Flags & RTXPT_MATFLAG_UseEmissiveTexture) r.emissiveColor = lp(r.emissiveColor * lp(xyz(texEmissive)));
    r.ior = lp(m.IoR);
    r.shadowNoLFadeout = lp(m.ShadowNoLFadeout);
    if (m.Flags & RTXPT_MATFLAG_UseNormalTexture)
    {   // ApplyNormalMapRTXPT, BridgeDonut:280-309
        float4 tangent = gs.tangent;
        float squareTangentLength = dot(xyz(tangent), xyz(tangent));
        if (squareTangentLength != 0 && tangent.w != 0)
        {
            float nx = (texNormal.x * 2.0f - 1.0f) * m.NormalTextureScale, ny = (texNormal.y * 2.0f - 1.0f) * m.NormalTextureScale, nz;
            if (texNormal.z <= 0) nz = sqrtf(saturate(1.0f - nx * nx - ny * ny));
            else nz = fabsf(texNormal.z * 2.0f - 1.0f);
            float sqLen = nx * nx + ny * ny + nz * nz;
            if (sqLen != 0)
            {
                float len = sqrtf(sqLen);
                float3 localNormal = f3(nx / len, ny / len, nz / len);
                float3 t = xyz(tangent) * (1.0f / sqrtf(squareTangentLength));
                float3 bitangent = cross(r.geometryNormal, t) * tangent.w;
                r.shadingNormal = normalize(t * localNormal.x + bitangent * localNormal.y + r.geometryNormal * localNormal.z);
            }
        }
    }
    return r;
}

// Bridge::loadSurface, PathTracerBridgeDonut.hlsli:612-853
inline SurfaceData loadSurface(const Scene& sc, uint instanceIndex, uint geometryIndex, uint triangleIndex, float2 barycentrics,
                               float3 rayDir, RayCone rayCone, float texLODBias, uint pathVertexIndex = 1, uint pixelX = 0, uint pixelY = 0, uint sampleIndex = 0)
{
    GeometrySample gs = getGeometryFromHit(sc, instanceIndex, geometryIndex, triangleIndex, barycentrics, rayDir);
    float3 posW = mul34_point(gs.instance->transform, gs.objectSpacePosition);
    float coneTexLODValue = computeRayConeTriangleLODValue(gs.vertexPositions, gs.vertexTexcoords, gs.instance->transform);
    float lambda = rayCone.computeLOD(coneTexLODValue, rayDir, gs.flatNormal, true) + texLODBias;

    SurfaceData out = {};
    out.prevPosW = mul34_point(gs.instance->prevTransform, gs.prevObjectSpacePosition);
    ShadingData& sd = out.sd;
    sd.posW = posW;
    sd.V = -rayDir;
    sd.N = gs.geometryNormal;
    uint subInstanceDataIndex = gs.instance->firstGeometryInstanceIndex + geometryIndex;
    uint materialIndex = sc.subInstances[subInstanceDataIndex].GlobalGeometryIndex_PTMaterialDataIndex & 0xFFFF;
    MaterialProperties mat = sampleGeometryMaterial(sc, gs, materialIndex, lambda);
    bool ignoreTangent = (mat.flags & RTXPT_MATFLAG_IgnoreMeshTangentSpace) != 0;
    computeTangentSpace(sd, gs.tangent, ignoreTangent);
    sd.faceNCorrected = gs.frontFacing ? gs.flatNormal : -gs.flatNormal;
    sd.vertexN = gs.frontFacing ? gs.geometryNormal : -gs.geometryNormal;
    sd.frontFacing = gs.frontFacing;
    sd.N = gs.frontFacing ? mat.shadingNormal : -mat.shadingNormal;
    sd.materialID = materialIndex;
    sd.nestedPriority = std::min(15u, 1u + (mat.flags >> RTXPT_MATFLAG_NestedPriorityShift));
    sd.thinSurface = (mat.flags & RTXPT_MATFLAG_ThinSurface) != 0;
    sd.psdExclude = (mat.flags & RTXPT_MATFLAG_PSDExclude) != 0;
    sd.psdDominantDeltaLobeP1 = (mat.flags & 0x0F000000u) >> 24;                                    // PTMaterialFlags_PSDDominantDeltaLobeP1Mask/Shift (BridgeDonut:700)
    {   // stopping motion vectors from being calculated behind this surface (BridgeDonut:702-718): 0 Off, 1 AutoLow, 2 AutoHigh (curvature seen through the ray cone), 3 Full
        const uint blockType = (mat.flags >> 13) & 3u;
        bool blockMVs = blockType == 3u;
        if (blockType == 1u || blockType == 2u)
        {
            const float projectionTerm = fabsf(dot(rayDir, -sd.N));
            const float pixelCurvature = (gs.curvatureWS * rayCone.getWidth()) / std::max(projectionTerm, 1e-6f);
            const float threshold = blockType == 1u ? 0.03f : 0.0005f;
            MicroRng rng = MicroRng::make(pixelX, pixelY, pathVertexIndex, sampleIndex);
            blockMVs |= pixelCurvature > ((rng.NextFloat() * 0.9f + 0.3f) * threshold);
        }
        sd.psdBlockMotionVectorsAtSurface = blockMVs;
    }
    adjustShadingNormal(sd, gs.tangent, true, ignoreTangent);
    sd.shadowNoLFadeout = mat.shadowNoLFadeout;

    float matIoR = mat.ior;
    StandardBSDFData& b = out.bsdf.data;
    b.specularTransmission = lp(mat.transmission * (1 - mat.metalness));
    b.diffuseTransmission = lp(mat.diffuseTransmission * (1 - mat.metalness));
    b.transmission = mat.baseColor;
    sd.activeLobes = Lobe_All;
    float f = (matIoR - 1.f) / (matIoR + 1.f);
    float F0 = f * f;
    b.diffuse = lp(lerp(mat.baseColor, f3(0), mat.metalness));
    b.specular = lp(lerp(lp(f3(F0)), mat.baseColor, mat.metalness));
    b.roughness = mat.roughness;
    b.metallic = mat.metalness;
    sd.IoR = 1.f;
    b.eta = lp(sd.IoR / matIoR);
    if (!sd.thinSurface && !sd.frontFacing) b.eta = lp(matIoR / sd.IoR);
    out.neeTriangleLightIndex = 0xFFFFFFFFu;
    out.neeAnalyticLightIndex = (mat.flags & RTXPT_MATFLAG_EnableAsAnalyticLightProxy) ? sc.subInstances[subInstanceDataIndex].AnalyticProxyLightIndex : 0xFFFFFFFFu;   // BridgeDonut:828-829
    sd.emission = f3(0);
    if (sd.frontFacing && any_gt0(mat.emissiveColor))
    {
        sd.emission = mat.emissiveColor;
        uint baseIndex = sc.subInstances[subInstanceDataIndex].EmissiveLightMappingOffset;
        if (baseIndex != 0xFFFFFFFFu) out.neeTriangleLightIndex = baseIndex + triangleIndex;
    }
    out.interiorIoR = matIoR;
    return out;
}

// AlphaTestImpl, PathTracerBridgeDonut.hlsli:929-971 (SUBINSTANCEDATA_EXTENDED path); true = opaque at this point
inline bool alphaTest(const Scene& sc, const RtxptSubInstanceData& s, uint triangleIndex, float2 rayBary)
{
    if ((s.FlagsAndAlphaInfo & RTXPT_SUBINST_FLAG_ALPHA_TESTED) == 0) return true;
    int ib = int(s.IndexBufferIndex_VertexBufferIndex >> 16), vb = int(s.IndexBufferIndex_VertexBufferIndex & 0xFFFF);
    float2 uv[3];
    for (int k = 0; k < 3; k++) uv[k] = sc.loadFloat2(vb, s.TexCoord1Offset + sc.load32(ib, s.IndexOffset + triangleIndex * 12 + k * 4) * 8);
    float3 bary = f3(1.0f - (rayBary.x + rayBary.y), rayBary.x, rayBary.y);
    float2 texcoord = uv[0] * bary.x + uv[1] * bary.y + uv[2] * bary.z;
    float opacity = sc.textures[s.FlagsAndAlphaInfo & 0xFFFF].sampleLevel(texcoord, 0).w;
    return opacity >= float(s.FlagsAndAlphaInfo >> 24) / 255.0f;
}

} // namespace orc
