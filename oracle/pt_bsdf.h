// ORACLE — test infrastructure only (see pt_math.h).
// pt_bsdf.h: StandardBSDF (Falcor BSDF) restated from
//   Rtxpt/Shaders/PathTracer/Rendering/Materials/Fresnel.hlsli:27-75
//   Rtxpt/Shaders/PathTracer/Rendering/Materials/Microfacet.hlsli:33-240
//   Rtxpt/Shaders/PathTracer/Rendering/Materials/BxDF.hlsli:31-51 (constants), :157-247 (diffuse lobes), :249-381 (specular
//   reflection), :385-607 (specular reflection+transmission), :709-1000 (FalcorBSDF mixture)
//   Rtxpt/Shaders/PathTracer/Rendering/Materials/StandardBSDF.hlsli:53-91
#pragma once
#include "pt_math.h"

namespace orc {

enum LobeType : uint {
    Lobe_None = 0x00, Lobe_DiffuseReflection = 0x01, Lobe_SpecularReflection = 0x02, Lobe_DeltaReflection = 0x04,
    Lobe_DiffuseTransmission = 0x10, Lobe_SpecularTransmission = 0x20, Lobe_DeltaTransmission = 0x40,
    Lobe_Diffuse = 0x11, Lobe_Specular = 0x22, Lobe_Delta = 0x44, Lobe_NonDelta = 0x33,
    Lobe_Reflection = 0x0f, Lobe_Transmission = 0xf0, Lobe_All = 0xff
};

static const float kMinCosTheta = 1e-6f;
static const float kMinGGXAlpha = 0.0064f;
static const float cOneMinusEpsilon = 0.99999994f;      // MathConstants.hlsli cFloatOneMinusEpsilon (0x1.fffffep-1)

// ---- Fresnel.hlsli ---------------------------------------------------------------------------------------------------
inline float pow5(float x) { return powf(x, 5.0f); }    // HLSL pow(x,5); keep the library call so both sides use "pow"
inline float3 evalFresnelSchlick(float3 f0, float f90, float cosTheta)
{
    float p = pow5(std::max(1 - cosTheta, 0.f));
    return f0 + (f3(f90) - f0) * p;
}
inline float evalFresnelSchlick(float f0, float f90, float cosTheta) { return f0 + (f90 - f0) * pow5(std::max(1 - cosTheta, 0.f)); }
inline float evalFresnelDielectric(float eta, float cosThetaI, float& cosThetaT)
{
    if (cosThetaI < 0) { eta = 1 / eta; cosThetaI = -cosThetaI; }
    float sinThetaTSq = eta * eta * (1 - cosThetaI * cosThetaI);
    if (sinThetaTSq > 1) { cosThetaT = 0; return 1; }
    cosThetaT = sqrtf(1 - sinThetaTSq);
    float Rs = (eta * cosThetaI - cosThetaT) / (eta * cosThetaI + cosThetaT);
    float Rp = (eta * cosThetaT - cosThetaI) / (eta * cosThetaT + cosThetaI);
    return 0.5f * (Rs * Rs + Rp * Rp);
}
inline float evalFresnelDielectric(float eta, float cosThetaI) { float t; return evalFresnelDielectric(eta, cosThetaI, t); }

// ---- Microfacet.hlsli --------------------------------------------------------------------------------------------------
inline float evalNdfGGX(float alpha, float cosTheta)
{
    float a2 = alpha * alpha;
    float d = ((cosTheta * a2 - cosTheta) * cosTheta + 1);
    return a2 / (d * d * K_PI);
}
inline float evalLambdaGGX(float alphaSqr, float cosTheta)
{
    if (cosTheta <= 0) return 0;
    float cosThetaSqr = cosTheta * cosTheta;
    float tanThetaSqr = std::max(1 - cosThetaSqr, 0.f) / cosThetaSqr;
    return 0.5f * (-1 + sqrtf(1 + alphaSqr * tanThetaSqr));
}
inline float evalMaskingSmithGGXCorrelated(float alpha, float cosThetaI, float cosThetaO)
{
    float alphaSqr = alpha * alpha;
    return 1 / (1 + evalLambdaGGX(alphaSqr, cosThetaI) + evalLambdaGGX(alphaSqr, cosThetaO));
}
// Bounded VNDF (Microfacet.hlsli:108-130 pdf, :187-207 sample)
// Microfacet.hlsli:282-355, SpecularMaskingFunctionSmithGGXCorrelated coefficients (BxDFConfig.hlsli:33)
inline float3 approxSpecularIntegralGGX(float3 specularReflectance, float alpha, float cosTheta)
{
    cosTheta = fabsf(cosTheta);
    const float X[4] = { 1.f, cosTheta, cosTheta * cosTheta, cosTheta * (cosTheta * cosTheta) };
    const float Y[4] = { 1.f, alpha, alpha * alpha, alpha * (alpha * alpha) };
    const float M1[2][2] = { { 0.995367f, -1.38839f }, { -0.24751f, 1.97442f } };
    const float M2[3][3] = { { 1.0f, 2.68132f, 52.366f }, { 16.0932f, -3.98452f, 59.3013f }, { -5.18731f, 255.259f, 2544.07f } };
    const float M3[2][2] = { { -0.0564526f, 3.82901f }, { 16.91f, -11.0303f } };
    const float M4[3][3] = { { 1.0f, 4.11118f, -1.37886f }, { 19.3254f, -28.9947f, 16.9514f }, { 0.545386f, 96.0994f, -79.4492f } };
    auto bil2 = [](const float M[2][2], float x0, float x1, float y0, float y1) { return (M[0][0] * x0 + M[0][1] * x1) * y0 + (M[1][0] * x0 + M[1][1] * x1) * y1; };
    auto bil3 = [](const float M[3][3], float x0, float x1, float x2, float y0, float y1, float y2) {
        return ((M[0][0] * x0 + M[0][1] * x1) + M[0][2] * x2) * y0 + ((M[1][0] * x0 + M[1][1] * x1) + M[1][2] * x2) * y1 + ((M[2][0] * x0 + M[2][1] * x1) + M[2][2] * x2) * y2; };
    float bias = bil2(M1, X[0], X[1], Y[0], Y[1]) * (1.0f / bil3(M2, X[0], X[1], X[3], Y[0], Y[1], Y[3]));
    const float scale = bil2(M3, X[0], X[1], Y[0], Y[1]) * (1.0f / bil3(M4, X[0], X[2], X[3], Y[0], Y[1], Y[3]));
    const float luma = dot(specularReflectance, f3(1.f / 3.f));
    bias *= saturate(luma * 50.0f);
    return specularReflectance * std::max(0.0f, scale) + f3(std::max(0.0f, bias));
}
inline float evalPdfGGX_BVNDF(float alpha, float3 i, float3 m)
{
    float ndf = evalNdfGGX(alpha, m.z);
    float2 ai = f2(alpha * i.x, alpha * i.y);
    float len2 = dot(ai, ai);
    float t = sqrtf(len2 + i.z * i.z);
    float a = saturate(alpha);
    float s = 1.0f + length(f2(i.x, i.y));
    float a2 = a * a, s2 = s * s;
    float k = (1.0f - a2) * s2 / (s2 + a2 * i.z * i.z);
    return ndf / (2.0f * (k * i.z + t));
}
inline float3 sampleGGX_BVNDF(float alpha, float3 i, float2 rand)
{
    float3 i_std = normalize(f3(i.x * alpha, i.y * alpha, i.z));
    float phi = 2.0f * K_PI * rand.x;
    float a = saturate(alpha);
    float s = 1.0f + length(f2(i.x, i.y));
    float a2 = a * a, s2 = s * s;
    float k = (1.0f - a2) * s2 / (s2 + a2 * i.z * i.z);
    float b = i.z > 0 ? k * i_std.z : i_std.z;
    float z = (1.0f - rand.y) * (1.0f + b) + (-b);      // mad(1-rand.y, 1+b, -b)
    float sinTheta = sqrtf(saturate(1.0f - z * z));
    float3 o_std = f3(sinTheta * cosf(phi), sinTheta * sinf(phi), z);
    float3 m_std = i_std + o_std;
    return normalize(f3(m_std.x * alpha, m_std.y * alpha, m_std.z));
}

// ---- BSDF inputs -------------------------------------------------------------------------------------------------------
struct StandardBSDFData     // BxDF.hlsli:612-705; every field is an lpfloat (fp16-rounded by the producer)
{
    float3 diffuse; float roughness;
    float3 specular; float metallic;
    float3 transmission; float diffuseTransmission, specularTransmission;
    float eta;
};

struct BSDFFrame            // the part of ShadingData the BSDF reads (Scene/ShadingData.hlsli:38-61)
{
    float3 T, B, N, V;
    bool thinSurface; bool psdExclude; uint activeLobes;
    float3 toLocal(float3 v) const { return f3(dot(v, T), dot(v, B), dot(v, N)); }
    float3 fromLocal(float3 v) const { return T * v.x + B * v.y + N * v.z; }
};

struct BSDFSample { float3 wo; float pdf; float3 weight; uint lobe; float lobeP; bool isLobe(uint t) const { return (lobe & t) != 0; }
                    // IBSDF.hlsli:53-60: 0 = delta transmission, 1 = delta reflection, 0xFFFFFFFF = not a delta lobe
                    uint getDeltaLobeIndex() const { if ((lobe & Lobe_Delta) == 0u) return 0xFFFFFFFFu; return (lobe & Lobe_Transmission) == 0u ? 1u : 0u; } };
struct DeltaLobe { float3 thp = f3(0); float probability = 0; float3 dir = f3(0); int transmission = 0; };    // IBSDF.hlsli:24-34
static const uint cMaxDeltaLobes = 3;       // IBSDF.hlsli:22

// ---- lobes ---------------------------------------------------------------------------------------------------------------
struct DiffuseReflectionFrostbite   // BxDF.hlsli:157-208
{
    float3 albedo; float roughness;
    float3 evalWeight(float3 wi, float3 wo) const
    {
        float3 h = normalize(wi + wo);
        float woDotH = dot(wo, h);
        float energyBias = lerp(0.f, 0.5f, roughness);
        float energyFactor = lerp(1.f, 1.f / 1.51f, roughness);
        float fd90 = energyBias + 2.f * woDotH * woDotH * roughness;
        float wiScatter = evalFresnelSchlick(1.f, fd90, wi.z);
        float woScatter = evalFresnelSchlick(1.f, fd90, wo.z);
        return albedo * wiScatter * woScatter * energyFactor;
    }
    float3 eval(float3 wi, float3 wo) const
    {
        if (std::min(wi.z, wo.z) < kMinCosTheta) return f3(0);
        return evalWeight(wi, wo) * K_1_PI * wo.z;
    }
    bool sample(float3 wi, float3& wo, float& pdf, float3& weight, uint& lobe, float& lobeP, float3 u) const
    {
        wo = sample_cosine_hemisphere_concentric(f2(u.x, u.y), pdf);
        lobe = Lobe_DiffuseReflection;
        if (std::min(wi.z, wo.z) < kMinCosTheta) { weight = f3(0); lobeP = 0; return false; }
        weight = evalWeight(wi, wo); lobeP = 1.0f;
        return true;
    }
    float evalPdf(float3 wi, float3 wo) const { return (std::min(wi.z, wo.z) < kMinCosTheta) ? 0.f : K_1_PI * wo.z; }
};

struct DiffuseTransmissionLambert   // BxDF.hlsli:212-247
{
    float3 albedo;
    float3 eval(float3 wi, float3 wo) const { return (std::min(wi.z, -wo.z) < kMinCosTheta) ? f3(0) : K_1_PI * albedo * -wo.z; }
    bool sample(float3 wi, float3& wo, float& pdf, float3& weight, uint& lobe, float& lobeP, float3 u) const
    {
        wo = sample_cosine_hemisphere_concentric(f2(u.x, u.y), pdf);
        wo.z = -wo.z;
        lobe = Lobe_DiffuseTransmission;
        if (std::min(wi.z, -wo.z) < kMinCosTheta) { weight = f3(0); lobeP = 0; return false; }
        weight = albedo; lobeP = 1.0f;
        return true;
    }
    float evalPdf(float3 wi, float3 wo) const { return (std::min(wi.z, -wo.z) < kMinCosTheta) ? 0.f : K_1_PI * -wo.z; }
};

// BxDF.hlsli:249-269 (Turquin-style multiple-scattering approximation)
inline float EmsApprox(float r2, float NdV) { float r4 = r2 * r2; return lerp(0.2f * r2, 0.32f * r2 + 1.94f * r4, NdV); }
inline float3 MultiScatterSpecularApprox(float alpha, float NdV, float3 F0) { return f3(1) + F0 * EmsApprox(alpha, NdV); }

struct SpecularReflectionMicrofacet  // BxDF.hlsli:273-381
{
    float3 albedo; float alpha; uint activeLobes;
    bool hasLobe(uint l) const { return (activeLobes & l) != 0; }
    float3 eval(float3 wi, float3 wo) const
    {
        if (std::min(wi.z, wo.z) < kMinCosTheta) return f3(0);
        if (alpha == 0.f) return f3(0);
        if (!hasLobe(Lobe_SpecularReflection)) return f3(0);
        float3 h = normalize(wi + wo);
        float wiDotH = dot(wi, h);
        float D = evalNdfGGX(alpha, h.z);
        float G = evalMaskingSmithGGXCorrelated(alpha, wi.z, wo.z);
        float3 F = evalFresnelSchlick(albedo, 1.f, wiDotH);
        float3 ms = MultiScatterSpecularApprox(alpha, wi.z, albedo);
        return ms * F * (D * G * 0.25f / wi.z);
    }
    float evalPdf(float3 wi, float3 wo) const
    {
        if (std::min(wi.z, wo.z) < kMinCosTheta) return 0.f;
        if (alpha == 0.f) return 0.f;
        if (!hasLobe(Lobe_SpecularReflection)) return 0.f;
        float3 h = normalize(wi + wo);
        return evalPdfGGX_BVNDF(alpha, wi, h);
    }
    bool sample(float3 wi, float3& wo, float& pdf, float3& weight, uint& lobe, float& lobeP, float3 u) const
    {
        wo = f3(0); weight = f3(0); pdf = 0.f; lobe = Lobe_SpecularReflection; lobeP = 1.0f;
        if (wi.z < kMinCosTheta) return false;
        if (alpha == 0.f)
        {
            if (!hasLobe(Lobe_DeltaReflection)) return false;
            wo = f3(-wi.x, -wi.y, wi.z);
            pdf = 0.f;
            weight = evalFresnelSchlick(albedo, 1.f, wi.z);
            lobe = Lobe_DeltaReflection;
            return true;
        }
        if (!hasLobe(Lobe_SpecularReflection)) return false;
        float3 h = sampleGGX_BVNDF(alpha, wi, f2(u.x, u.y));
        float wiDotH = dot(wi, h);
        wo = 2.f * wiDotH * h - wi;
        if (wo.z < kMinCosTheta) return false;
        pdf = evalPdf(wi, wo);
        weight = eval(wi, wo) / pdf;
        lobe = Lobe_SpecularReflection;
        return true;
    }
};

struct SpecularReflectionTransmissionMicrofacet     // BxDF.hlsli:385-607
{
    float3 transmissionAlbedo; float alpha; float eta; uint activeLobes; bool isThinSurface;
    bool hasLobe(uint l) const { return (activeLobes & l) != 0; }
    float3 eval(float3 wi, float3 wo) const
    {
        if (std::min(wi.z, fabsf(wo.z)) < kMinCosTheta) return f3(0);
        if (alpha == 0.f) return f3(0);
        const bool hasReflection = hasLobe(Lobe_SpecularReflection), hasTransmission = hasLobe(Lobe_SpecularTransmission);
        const bool isReflection = wo.z > 0.f;
        if ((isReflection && !hasReflection) || (!isReflection && !hasTransmission)) return f3(0);
        float actualEta = (isThinSurface && !isReflection) ? 1.0f : eta;
        float3 h = normalize(wo + wi * (isReflection ? 1.f : actualEta));
        h = h * signf(h.z);
        float wiDotH = dot(wi, h), woDotH = dot(wo, h);
        float D = evalNdfGGX(alpha, h.z);
        float G = evalMaskingSmithGGXCorrelated(alpha, wi.z, fabsf(wo.z));
        float F = evalFresnelDielectric(actualEta, wiDotH);
        if (isReflection) return f3(F * D * G * 0.25f / wi.z);
        float sqrtDenom = woDotH + actualEta * wiDotH;
        float t = actualEta * actualEta * wiDotH * woDotH / (wi.z * sqrtDenom * sqrtDenom);
        return transmissionAlbedo * (1.f - F) * D * G * fabsf(t);      // left to right as written in BxDF.hlsli:436 (pinned by tests/golden/bsdf_golden.npz)
    }
    float evalPdf(float3 wi, float3 wo) const
    {
        if (std::min(wi.z, fabsf(wo.z)) < kMinCosTheta) return 0.f;
        if (alpha == 0.f) return 0.f;
        bool isReflection = wo.z > 0.f;
        const bool hasReflection = hasLobe(Lobe_SpecularReflection), hasTransmission = hasLobe(Lobe_SpecularTransmission);
        if ((isReflection && !hasReflection) || (!isReflection && !hasTransmission)) return 0.f;
        float actualEta = (isThinSurface && !isReflection) ? 1.0f : eta;
        float3 h = normalize(wo + wi * (isReflection ? 1.f : actualEta));
        h = h * signf(h.z);
        float wiDotH = dot(wi, h), woDotH = dot(wo, h);
        float F = evalFresnelDielectric(actualEta, wiDotH);
        float pdf = evalPdfGGX_BVNDF(alpha, wi, h);
        if (isReflection)
        {
            if (woDotH <= 0.f) return 0.f;
            pdf *= wiDotH / woDotH;
        }
        else
        {
            if (woDotH > 0.f) return 0.f;
            pdf *= wiDotH * 4.0f;
            float sqrtDenom = woDotH + actualEta * wiDotH;
            float denom = sqrtDenom * sqrtDenom;
            pdf *= fabsf(woDotH) / denom;
        }
        if (hasReflection && hasTransmission) pdf *= isReflection ? F : 1.f - F;
        return clampf(pdf, 0, FLT_MAX_);
    }
    bool sample(float3 wi, float3& wo, float& pdf, float3& weight, uint& lobe, float& lobeP, float3 u) const
    {
        wo = f3(0); weight = f3(0); pdf = 0.f; lobe = Lobe_SpecularReflection; lobeP = 1;
        if (wi.z < kMinCosTheta) return false;
        float lobeSample = u.z;
        if (alpha == 0.f)
        {
            const bool hasReflection = hasLobe(Lobe_DeltaReflection), hasTransmission = hasLobe(Lobe_DeltaTransmission);
            if (!(hasReflection || hasTransmission)) return false;
            float cosThetaT;
            float F = evalFresnelDielectric(eta, wi.z, cosThetaT);
            bool isReflection = hasReflection;
            if (hasReflection && hasTransmission) { isReflection = lobeSample < F; lobeP = isReflection ? F : (1 - F); }
            else if (hasTransmission && F == 1.f) return false;
            float actualEta = eta;
            if (isThinSurface && !isReflection) { actualEta = 1.0f; F = evalFresnelDielectric(actualEta, wi.z, cosThetaT); }
            pdf = 0.f;
            weight = isReflection ? f3(1) : transmissionAlbedo;
            if (!(hasReflection && hasTransmission)) weight *= (isReflection ? F : 1.f - F);
            wo = isReflection ? f3(-wi.x, -wi.y, wi.z) : f3(-wi.x * actualEta, -wi.y * actualEta, -cosThetaT);
            lobe = isReflection ? Lobe_DeltaReflection : Lobe_DeltaTransmission;
            if (fabsf(wo.z) < kMinCosTheta || ((wo.z > 0.f) != isReflection)) return false;
            return true;
        }
        const bool hasReflection = hasLobe(Lobe_SpecularReflection), hasTransmission = hasLobe(Lobe_SpecularTransmission);
        if (!(hasReflection || hasTransmission)) return false;
        float3 h = sampleGGX_BVNDF(alpha, wi, f2(u.x, u.y));
        float wiDotH = dot(wi, h);
        float cosThetaT;
        float F = evalFresnelDielectric(eta, wiDotH, cosThetaT);
        bool isReflection = hasReflection;
        if (hasReflection && hasTransmission) isReflection = lobeSample < F;
        else if (hasTransmission && F == 1.f) return false;
        float actualEta = eta;
        if (isThinSurface && !isReflection) { actualEta = 1.0f; F = evalFresnelDielectric(actualEta, wi.z, cosThetaT); }  // sic: wi.z (BxDF.hlsli:532)
        wo = isReflection ? (2.f * wiDotH * h - wi) : ((actualEta * wiDotH - cosThetaT) * h - actualEta * wi);
        if (fabsf(wo.z) < kMinCosTheta || ((wo.z > 0.f) != isReflection)) return false;
        lobe = isReflection ? Lobe_SpecularReflection : Lobe_SpecularTransmission;
        pdf = evalPdf(wi, wo);
        weight = pdf > 0.f ? eval(wi, wo) / pdf : f3(0);
        return true;
    }
};

// ---- the mixture (BxDF.hlsli:709-1000) -----------------------------------------------------------------------------------
struct FalcorBSDF
{
    DiffuseReflectionFrostbite diffuseReflection;
    DiffuseTransmissionLambert diffuseTransmission;
    SpecularReflectionMicrofacet specularReflection;
    SpecularReflectionTransmissionMicrofacet specularReflectionTransmission;
    float diffTrans, specTrans;
    float pDiffuseReflection, pDiffuseTransmission, pSpecularReflection, pSpecularReflectionTransmission;

    static FalcorBSDF make(const BSDFFrame& sd, const StandardBSDFData& data)
    {
        FalcorBSDF b;
        bool isThinSurface = sd.thinSurface;
        float3 dataTransmission = data.transmission;
        float3 transmissionAlbedo = isThinSurface ? dataTransmission : f3(sqrtf(dataTransmission.x), sqrtf(dataTransmission.y), sqrtf(dataTransmission.z));
        float dataRoughness = data.roughness;
        b.diffuseReflection.albedo = data.diffuse;
        b.diffuseReflection.roughness = dataRoughness;
        b.diffuseTransmission.albedo = transmissionAlbedo;
        float alpha = dataRoughness * dataRoughness;
        if (alpha < kMinGGXAlpha) alpha = 0.f;
        const uint activeLobes = sd.activeLobes;
        b.specularReflection.albedo = data.specular;
        b.specularReflection.alpha = alpha;
        b.specularReflection.activeLobes = activeLobes;
        b.specularReflectionTransmission.transmissionAlbedo = transmissionAlbedo;
        b.specularReflectionTransmission.alpha = (data.eta == 1.f) ? 0.f : alpha;
        b.specularReflectionTransmission.eta = data.eta;
        b.specularReflectionTransmission.activeLobes = activeLobes;
        b.specularReflectionTransmission.isThinSurface = isThinSurface;
        b.diffTrans = data.diffuseTransmission;
        b.specTrans = data.specularTransmission;
        float metallicBRDF = data.metallic * (1.f - b.specTrans);
        float dielectricBSDF = (1.f - data.metallic) * (1.f - b.specTrans);
        float specularBSDF = b.specTrans;
        float diffuseWeight = Luminance(data.diffuse);
        float specularWeight = Luminance(evalFresnelSchlick(data.specular, 1.f, dot(sd.V, sd.N)));
        b.pDiffuseReflection = (activeLobes & Lobe_DiffuseReflection) ? diffuseWeight * dielectricBSDF * (1.f - b.diffTrans) : 0.f;
        b.pDiffuseTransmission = (activeLobes & Lobe_DiffuseTransmission) ? diffuseWeight * dielectricBSDF * b.diffTrans : 0.f;
        b.pSpecularReflection = (activeLobes & (Lobe_SpecularReflection | Lobe_DeltaReflection)) ? specularWeight * (metallicBRDF + dielectricBSDF) : 0.f;
        b.pSpecularReflectionTransmission = (activeLobes & (Lobe_SpecularReflection | Lobe_DeltaReflection | Lobe_SpecularTransmission | Lobe_DeltaTransmission)) ? specularBSDF : 0.f;
        float normFactor = b.pDiffuseReflection + b.pDiffuseTransmission + b.pSpecularReflection + b.pSpecularReflectionTransmission;
        if (normFactor > 0.f)
        {
            normFactor = 1.f / normFactor;
            b.pDiffuseReflection *= normFactor; b.pDiffuseTransmission *= normFactor;
            b.pSpecularReflection *= normFactor; b.pSpecularReflectionTransmission *= normFactor;
        }
        return b;
    }

    // BxDF.hlsli:972-1053; wi in the local frame, directions returned in the local frame
    void evalDeltaLobes(float3 wi, bool psdExclude, DeltaLobe deltaLobes[cMaxDeltaLobes], int& deltaLobeCount, float& nonDeltaPart) const
    {
        deltaLobeCount = 2;
        for (uint i = 0; i < cMaxDeltaLobes; i++) deltaLobes[i] = DeltaLobe();
        nonDeltaPart = pDiffuseReflection + pDiffuseTransmission;
        if (specularReflection.alpha > 0) nonDeltaPart += pSpecularReflection;
        if (specularReflectionTransmission.alpha > 0) nonDeltaPart += pSpecularReflectionTransmission;
        if ((pSpecularReflection + pSpecularReflectionTransmission) == 0 || psdExclude) return;
        DeltaLobe deltaReflection, deltaTransmission;
        deltaReflection.transmission = 0; deltaTransmission.transmission = 1;
        deltaReflection.dir = f3(-wi.x, -wi.y, wi.z);
        if (specularReflection.alpha == 0 && specularReflection.hasLobe(Lobe_DeltaReflection))
        {
            deltaReflection.probability = pSpecularReflection;
            deltaReflection.thp = (1 - pSpecularReflectionTransmission) * evalFresnelSchlick(specularReflection.albedo, 1.f, wi.z);
        }
        if (specularReflectionTransmission.alpha == 0.f)
        {
            const bool hasReflection = specularReflectionTransmission.hasLobe(Lobe_DeltaReflection), hasTransmission = specularReflectionTransmission.hasLobe(Lobe_DeltaTransmission);
            if (hasReflection || hasTransmission)
            {
                float cosThetaT;
                float F = evalFresnelDielectric(specularReflectionTransmission.eta, wi.z, cosThetaT);
                if (hasReflection)
                {
                    float localProbability = pSpecularReflectionTransmission * F;
                    deltaReflection.thp = deltaReflection.thp + f3(1) * localProbability;
                    deltaReflection.probability += localProbability;
                }
                if (hasTransmission)
                {
                    float actualEta = specularReflectionTransmission.eta;
                    if (specularReflectionTransmission.isThinSurface) { actualEta = 1.0f; F = evalFresnelDielectric(actualEta, wi.z, cosThetaT); }
                    float localProbability = pSpecularReflectionTransmission * (1.0f - F);
                    deltaTransmission.dir = f3(-wi.x * actualEta, -wi.y * actualEta, -cosThetaT);
                    deltaTransmission.thp = specularReflectionTransmission.transmissionAlbedo * localProbability;
                    deltaTransmission.probability = localProbability;
                }
            }
        }
        deltaLobes[0] = deltaTransmission; deltaLobes[1] = deltaReflection;
    }

    static uint getLobes(const StandardBSDFData& data)
    {
        float alpha = data.roughness * data.roughness;
        bool isDelta = alpha < kMinGGXAlpha;
        float diffTrans = data.diffuseTransmission, specTrans = data.specularTransmission;
        uint lobes = isDelta ? Lobe_DeltaReflection : Lobe_SpecularReflection;
        if (any_gt0(data.diffuse) && specTrans < 1.f)
        {
            if (diffTrans < 1.f) lobes |= Lobe_DiffuseReflection;
            if (diffTrans > 0.f) lobes |= Lobe_DiffuseTransmission;
        }
        if (specTrans > 0.f) lobes |= (isDelta ? Lobe_DeltaTransmission : Lobe_SpecularTransmission);
        return lobes;
    }

    float4 eval(float3 wi, float3 wo) const
    {
        float3 diffuse = f3(0), specular = f3(0);
        if (pDiffuseReflection > 0.f) diffuse += (1.f - specTrans) * (1.f - diffTrans) * diffuseReflection.eval(wi, wo);
        if (pDiffuseTransmission > 0.f) diffuse += (1.f - specTrans) * diffTrans * diffuseTransmission.eval(wi, wo);
        if (pSpecularReflection > 0.f) specular += (1.f - specTrans) * specularReflection.eval(wi, wo);
        if (pSpecularReflectionTransmission > 0.f) specular += specTrans * specularReflectionTransmission.eval(wi, wo);
        return f4(diffuse + specular, Average(specular));
    }

    float evalPdf(float3 wi, float3 wo) const
    {
        float pdf = 0.f;
        if (pDiffuseReflection > 0.f) pdf += pDiffuseReflection * diffuseReflection.evalPdf(wi, wo);
        if (pDiffuseTransmission > 0.f) pdf += pDiffuseTransmission * diffuseTransmission.evalPdf(wi, wo);
        if (pSpecularReflection > 0.f) pdf += pSpecularReflection * specularReflection.evalPdf(wi, wo);
        if (pSpecularReflectionTransmission > 0.f) pdf += pSpecularReflectionTransmission * specularReflectionTransmission.evalPdf(wi, wo);
        return pdf;
    }

    // RecycleSelectSamples = 1 (BxDF.hlsli:45): three random numbers, .z selects the lobe and is re-stretched for the chosen lobe
    bool sample(float3 wi, float3& wo, float& pdf, float3& weight, uint& lobe, float& lobeP, float3 u) const
    {
        wo = f3(0); weight = f3(0); pdf = 0.f; lobe = Lobe_DiffuseReflection; lobeP = 0.0f;
        bool valid = false;
        float uSelect = u.z;
        if (uSelect < pDiffuseReflection)
        {
            u.z = clampf(uSelect / pDiffuseReflection, 0, cOneMinusEpsilon);
            valid = diffuseReflection.sample(wi, wo, pdf, weight, lobe, lobeP, u);
            weight /= pDiffuseReflection;
            weight *= (1.f - specTrans) * (1.f - diffTrans);
            pdf *= pDiffuseReflection;
            lobeP *= pDiffuseReflection;
            if (pSpecularReflection > 0.f) pdf += pSpecularReflection * specularReflection.evalPdf(wi, wo);
            if (pSpecularReflectionTransmission > 0.f) pdf += pSpecularReflectionTransmission * specularReflectionTransmission.evalPdf(wi, wo);
        }
        else if (uSelect < pDiffuseReflection + pDiffuseTransmission)
        {
            valid = diffuseTransmission.sample(wi, wo, pdf, weight, lobe, lobeP, u);
            weight /= pDiffuseTransmission;
            weight *= (1.f - specTrans) * diffTrans;
            pdf *= pDiffuseTransmission;
            lobeP *= pDiffuseTransmission;
            if (pSpecularReflectionTransmission > 0.f) pdf += pSpecularReflectionTransmission * specularReflectionTransmission.evalPdf(wi, wo);
        }
        else if (uSelect < pDiffuseReflection + pDiffuseTransmission + pSpecularReflection)
        {
            u.z = clampf((uSelect - (pDiffuseReflection + pDiffuseTransmission)) / pSpecularReflection, 0, cOneMinusEpsilon);
            valid = specularReflection.sample(wi, wo, pdf, weight, lobe, lobeP, u);
            weight /= pSpecularReflection;
            weight *= (1.f - specTrans);
            pdf *= pSpecularReflection;
            lobeP *= pSpecularReflection;
            if (pDiffuseReflection > 0.f) pdf += pDiffuseReflection * diffuseReflection.evalPdf(wi, wo);
            if (pSpecularReflectionTransmission > 0.f) pdf += pSpecularReflectionTransmission * specularReflectionTransmission.evalPdf(wi, wo);
        }
        else if (pSpecularReflectionTransmission > 0.f)
        {
            u.z = clampf((uSelect - (pDiffuseReflection + pDiffuseTransmission + pSpecularReflection)) / pSpecularReflectionTransmission, 0, cOneMinusEpsilon);
            valid = specularReflectionTransmission.sample(wi, wo, pdf, weight, lobe, lobeP, u);
            weight /= pSpecularReflectionTransmission;
            weight *= specTrans;
            pdf *= pSpecularReflectionTransmission;
            lobeP *= pSpecularReflectionTransmission;
            if (pDiffuseReflection > 0.f) pdf += pDiffuseReflection * diffuseReflection.evalPdf(wi, wo);
            if (pDiffuseTransmission > 0.f) pdf += pDiffuseTransmission * diffuseTransmission.evalPdf(wi, wo);
            if (pSpecularReflection > 0.f) pdf += pSpecularReflection * specularReflection.evalPdf(wi, wo);
        }
        if (!valid || (lobe & Lobe_Delta) != 0) pdf = 0.0f;
        return valid;
    }
};

// StandardBSDF.hlsli:53-91 (world-space wrappers; kUseBSDFSampling = true, PathTracer.hlsli:21)
struct StandardBSDF
{
    StandardBSDFData data;
    float4 eval(const BSDFFrame& sd, float3 wo) const { return FalcorBSDF::make(sd, data).eval(sd.toLocal(sd.V), sd.toLocal(wo)); }
    float evalPdf(const BSDFFrame& sd, float3 wo) const { return FalcorBSDF::make(sd, data).evalPdf(sd.toLocal(sd.V), sd.toLocal(wo)); }
    bool sample(const BSDFFrame& sd, const float u[4], BSDFSample& r) const
    {
        float3 woLocal = f3(0);
        bool valid = FalcorBSDF::make(sd, data).sample(sd.toLocal(sd.V), woLocal, r.pdf, r.weight, r.lobe, r.lobeP, f3(u[0], u[1], u[2]));
        r.wo = sd.fromLocal(woLocal);
        return valid;
    }
    uint getLobes() const { return FalcorBSDF::getLobes(data); }
    // StandardBSDF.hlsli:227-237
    void evalDeltaLobes(const BSDFFrame& sd, DeltaLobe deltaLobes[cMaxDeltaLobes], int& deltaLobeCount, float& nonDeltaPart) const
    {
        FalcorBSDF::make(sd, data).evalDeltaLobes(sd.toLocal(sd.V), sd.psdExclude, deltaLobes, deltaLobeCount, nonDeltaPart);
        for (int i = 0; i < deltaLobeCount; i++) deltaLobes[i].dir = sd.fromLocal(deltaLobes[i].dir);
    }
    // StandardBSDF.hlsli:93-121; albedo products are lpfloat (fp16) arithmetic in the reference
    void estimateSpecDiffBSDF(float3& outDiffEstimate, float3& outSpecEstimate, float3 normal, float3 viewVector) const
    {
        const float dataRoughness = data.roughness;
        const float alpha = dataRoughness * dataRoughness;
        const float roughness = alpha < kMinGGXAlpha ? 0.f : dataRoughness;
        const float dT = data.diffuseTransmission, sT = data.specularTransmission;
        const float3 diffuseReflectionAlbedo = lp(lp(lp(1.f - dT) * lp(1.f - sT)) * data.diffuse);
        const float3 diffuseTransmissionAlbedo = lp(lp(dT * data.transmission) * lp(1.f - sT));
        const float3 specularReflectionAlbedo = lp(lp(1.f - sT) * data.specular);
        const float3 specularTransmissionAlbedo = lp(sT * data.transmission);
        outDiffEstimate = lp(diffuseReflectionAlbedo + diffuseTransmissionAlbedo);
        const float NdotV = saturate(dot(normal, viewVector));
        const float ggxAlpha = roughness * roughness;
        outSpecEstimate = approxSpecularIntegralGGX(specularReflectionAlbedo, ggxAlpha, NdotV) + specularTransmissionAlbedo;
    }
};

} // namespace orc
