// bsdf.cuh — StandardBSDF (Falcor BSDF) for the shade kernel: Frostbite diffuse, Lambert diffuse transmission, GGX specular
// reflection (bounded-VNDF sampling, height-correlated Smith, Turquin multiple-scattering approximation), GGX dielectric
// reflection+transmission, delta lobes below kMinGGXAlpha, luminance-weighted lobe selection.
// Reference: Rtxpt/Shaders/PathTracer/Rendering/Materials/StandardBSDF.hlsli:53-91, BxDF.hlsli:157-247, :249-381, :385-607,
// :709-1000, Microfacet.hlsli:33-207, Fresnel.hlsli:27-75.
// Structure differs from the reference on purpose: the lobe mixture is set up once per vertex (BsdfSetup) and reused for the one
// sample, the NEE candidate pdfs and the final eval, instead of being rebuilt inside every eval/evalPdf/sample call.
#pragma once
#include "device_math.cuh"

namespace pt {

enum : uint {
    kLobeDiffuseReflection = 0x01, kLobeSpecularReflection = 0x02, kLobeDeltaReflection = 0x04,
    kLobeDiffuseTransmission = 0x10, kLobeSpecularTransmission = 0x20, kLobeDeltaTransmission = 0x40,
    kLobeDelta = 0x44, kLobeNonDelta = 0x33, kLobeReflection = 0x0f, kLobeTransmission = 0xf0, kLobeAll = 0xff
};
constexpr float kMinCosTheta = 1e-6f;
constexpr float kMinGGXAlpha = 0.0064f;
constexpr float kOneMinusEpsilon = 0.99999994f;

struct BsdfParams           // StandardBSDFData (BxDF.hlsli:612-705); lpfloat fields hold fp16-rounded values
{
    float3 diffuse; float roughness;
    float3 specular; float metallic;
    float3 transmission; float diffuseTransmission, specularTransmission, eta;
};

#if defined(PT_FAST_MATH)
PT_DEVICE float schlickPow5(float cosTheta) { const float x = fmaxf(1.0f - cosTheta, 0.0f), x2 = x * x; return x2 * x2 * x; }
#else
PT_DEVICE float schlickPow5(float cosTheta) { return powf(fmaxf(1.0f - cosTheta, 0.0f), 5.0f); }     // pow() as in the reference (and the oracle)
#endif
PT_DEVICE float3 fresnelSchlick3(float3 f0, float f90, float cosTheta) { return f0 + (mk3(f90) - f0) * schlickPow5(cosTheta); }
PT_DEVICE float fresnelSchlick1(float f0, float f90, float cosTheta) { return f0 + (f90 - f0) * schlickPow5(cosTheta); }
PT_DEVICE float fresnelDielectric(float eta, float cosThetaI, float& cosThetaT)
{
    if (cosThetaI < 0) { eta = 1 / eta; cosThetaI = -cosThetaI; }
    const float sinThetaTSq = eta * eta * (1 - cosThetaI * cosThetaI);
    if (sinThetaTSq > 1) { cosThetaT = 0; return 1; }
    cosThetaT = sqrtf(1 - sinThetaTSq);
    const float Rs = (eta * cosThetaI - cosThetaT) / (eta * cosThetaI + cosThetaT);
    const float Rp = (eta * cosThetaT - cosThetaI) / (eta * cosThetaT + cosThetaI);
    return 0.5f * (Rs * Rs + Rp * Rp);
}
PT_DEVICE float ggxNdf(float alpha, float cosTheta)
{
    const float a2 = alpha * alpha;
    const float d = ((cosTheta * a2 - cosTheta) * cosTheta + 1);
    return a2 / (d * d * kPi);
}
PT_DEVICE float ggxLambda(float alphaSqr, float cosTheta)
{
    if (cosTheta <= 0) return 0;
    const float c2 = cosTheta * cosTheta;
    const float tan2 = fmaxf(1 - c2, 0.f) / c2;
    return 0.5f * (-1 + sqrtf(1 + alphaSqr * tan2));
}
PT_DEVICE float smithGGXCorrelated(float alpha, float cosI, float cosO) { const float a2 = alpha * alpha; return 1 / (1 + ggxLambda(a2, cosI) + ggxLambda(a2, cosO)); }
PT_DEVICE float boundedVndfPdf(float alpha, float3 i, float3 m)        // Microfacet.hlsli:108-130
{
    const float ndf = ggxNdf(alpha, m.z);
    const float aix = alpha * i.x, aiy = alpha * i.y;
    const float t = sqrtf((aix * aix + aiy * aiy) + i.z * i.z);
    const float a = sat(alpha);
    const float s = 1.0f + sqrtf(i.x * i.x + i.y * i.y);
    const float a2 = a * a, s2 = s * s;
    const float k = (1.0f - a2) * s2 / (s2 + a2 * i.z * i.z);
    return ndf / (2.0f * (k * i.z + t));
}
PT_DEVICE float3 boundedVndfSample(float alpha, float3 i, float u0, float u1)    // Microfacet.hlsli:187-207
{
    const float3 iStd = norm3(mk3(i.x * alpha, i.y * alpha, i.z));
    const float phi = 2.0f * kPi * u0;
    const float a = sat(alpha);
    const float s = 1.0f + sqrtf(i.x * i.x + i.y * i.y);
    const float a2 = a * a, s2 = s * s;
    const float k = (1.0f - a2) * s2 / (s2 + a2 * i.z * i.z);
    const float b = i.z > 0 ? k * iStd.z : iStd.z;
    const float z = (1.0f - u1) * (1.0f + b) + (-b);
    const float sinTheta = sqrtf(sat(1.0f - z * z));
    const float3 mStd = iStd + mk3(sinTheta * cosf(phi), sinTheta * sinf(phi), z);
    return norm3(mk3(mStd.x * alpha, mStd.y * alpha, mStd.z));
}

struct BsdfSample { float3 wo; float pdf; float3 weight; uint lobe; float lobeP; };

struct BsdfSetup
{
    // frame
    float3 T, B, N;
    float3 wi;                  // view direction in the local frame
    // lobes
    float3 diffuseAlbedo; float roughness;
    float3 transAlbedo;
    float3 specAlbedo; float alphaRefl;
    float alphaTrans, eta;
    bool thin;
    float diffTrans, specTrans;
    float pDR, pDT, pSR, pSRT;  // selection probabilities

    PT_DEVICE float3 toLocal(float3 v) const { return mk3(dot3(v, T), dot3(v, B), dot3(v, N)); }
    PT_DEVICE float3 fromLocal(float3 v) const { return T * v.x + B * v.y + N * v.z; }

    // FalcorBSDF::__init (BxDF.hlsli:740-813) with all lobes active (BridgeDonut:723 sets LobeType::All)
    PT_DEVICE void init(float3 t, float3 b, float3 n, float3 v, bool thinSurface, const BsdfParams& d)
    {
        T = t; B = b; N = n; wi = toLocal(v); thin = thinSurface;
        transAlbedo = thinSurface ? d.transmission : mk3(sqrtf(d.transmission.x), sqrtf(d.transmission.y), sqrtf(d.transmission.z));
        diffuseAlbedo = d.diffuse; roughness = d.roughness;
        float alpha = d.roughness * d.roughness;
        if (alpha < kMinGGXAlpha) alpha = 0.f;
        specAlbedo = d.specular; alphaRefl = alpha;
        alphaTrans = (d.eta == 1.f) ? 0.f : alpha; eta = d.eta;
        diffTrans = d.diffuseTransmission; specTrans = d.specularTransmission;
        const float metallicBRDF = d.metallic * (1.f - specTrans);
        const float dielectricBSDF = (1.f - d.metallic) * (1.f - specTrans);
        const float diffuseWeight = luminance(d.diffuse);
        const float specularWeight = luminance(fresnelSchlick3(d.specular, 1.f, dot3(v, n)));
        pDR = diffuseWeight * dielectricBSDF * (1.f - diffTrans);
        pDT = diffuseWeight * dielectricBSDF * diffTrans;
        pSR = specularWeight * (metallicBRDF + dielectricBSDF);
        pSRT = specTrans;
        float norm = pDR + pDT + pSR + pSRT;
        if (norm > 0.f) { norm = 1.f / norm; pDR *= norm; pDT *= norm; pSR *= norm; pSRT *= norm; }
    }

    // ---- diffuse reflection, Frostbite (BxDF.hlsli:157-208) ----
    PT_DEVICE float3 drWeight(float3 wo) const
    {
        const float3 h = norm3(wi + wo);
        const float woDotH = dot3(wo, h);
        const float energyBias = lerpf(0.f, 0.5f, roughness);
        const float energyFactor = lerpf(1.f, 1.f / 1.51f, roughness);
        const float fd90 = energyBias + 2.f * woDotH * woDotH * roughness;
        return diffuseAlbedo * fresnelSchlick1(1.f, fd90, wi.z) * fresnelSchlick1(1.f, fd90, wo.z) * energyFactor;
    }
    PT_DEVICE float3 drEval(float3 wo) const { return (fminf(wi.z, wo.z) < kMinCosTheta) ? mk3(0.f) : drWeight(wo) * k1OverPi * wo.z; }
    PT_DEVICE float drPdf(float3 wo) const { return (fminf(wi.z, wo.z) < kMinCosTheta) ? 0.f : k1OverPi * wo.z; }
    // ---- diffuse transmission, Lambert (BxDF.hlsli:212-247) ----
    PT_DEVICE float3 dtEval(float3 wo) const { return (fminf(wi.z, -wo.z) < kMinCosTheta) ? mk3(0.f) : k1OverPi * transAlbedo * -wo.z; }
    PT_DEVICE float dtPdf(float3 wo) const { return (fminf(wi.z, -wo.z) < kMinCosTheta) ? 0.f : k1OverPi * -wo.z; }
    // ---- specular reflection (BxDF.hlsli:273-381) ----
    PT_DEVICE float3 srEval(float3 wo) const
    {
        if (fminf(wi.z, wo.z) < kMinCosTheta || alphaRefl == 0.f) return mk3(0.f);
        const float3 h = norm3(wi + wo);
        const float D = ggxNdf(alphaRefl, h.z);
        const float G = smithGGXCorrelated(alphaRefl, wi.z, wo.z);
        const float3 F = fresnelSchlick3(specAlbedo, 1.f, dot3(wi, h));
        const float r4 = alphaRefl * alphaRefl;
        const float ems = lerpf(0.2f * alphaRefl, 0.32f * alphaRefl + 1.94f * r4, wi.z);     // EmsApprox (BxDF.hlsli:251-259)
        const float3 ms = mk3(1.f) + specAlbedo * ems;
        return ms * F * (D * G * 0.25f / wi.z);
    }
    PT_DEVICE float srPdf(float3 wo) const
    {
        if (fminf(wi.z, wo.z) < kMinCosTheta || alphaRefl == 0.f) return 0.f;
        return boundedVndfPdf(alphaRefl, wi, norm3(wi + wo));
    }
    // ---- specular reflection + transmission (BxDF.hlsli:385-607) ----
    PT_DEVICE float3 srtEval(float3 wo) const
    {
        if (fminf(wi.z, fabsf(wo.z)) < kMinCosTheta || alphaTrans == 0.f) return mk3(0.f);
        const bool isReflection = wo.z > 0.f;
        const float actualEta = (thin && !isReflection) ? 1.0f : eta;
        float3 h = norm3(wo + wi * (isReflection ? 1.f : actualEta));
        h = h * sgn(h.z);
        const float wiDotH = dot3(wi, h), woDotH = dot3(wo, h);
        const float D = ggxNdf(alphaTrans, h.z);
        const float G = smithGGXCorrelated(alphaTrans, wi.z, fabsf(wo.z));
        float ct; const float F = fresnelDielectric(actualEta, wiDotH, ct);
        if (isReflection) return mk3(F * D * G * 0.25f / wi.z);
        const float sqrtDenom = woDotH + actualEta * wiDotH;
        const float t = actualEta * actualEta * wiDotH * woDotH / (wi.z * sqrtDenom * sqrtDenom);
        return transAlbedo * (1.f - F) * D * G * fabsf(t);        // left to right, BxDF.hlsli:436
    }
    PT_DEVICE float srtPdf(float3 wo) const
    {
        if (fminf(wi.z, fabsf(wo.z)) < kMinCosTheta || alphaTrans == 0.f) return 0.f;
        const bool isReflection = wo.z > 0.f;
        const float actualEta = (thin && !isReflection) ? 1.0f : eta;
        float3 h = norm3(wo + wi * (isReflection ? 1.f : actualEta));
        h = h * sgn(h.z);
        const float wiDotH = dot3(wi, h), woDotH = dot3(wo, h);
        float ct; const float F = fresnelDielectric(actualEta, wiDotH, ct);
        float pdf = boundedVndfPdf(alphaTrans, wi, h);
        if (isReflection) { if (woDotH <= 0.f) return 0.f; pdf *= wiDotH / woDotH; }
        else
        {
            if (woDotH > 0.f) return 0.f;
            pdf *= wiDotH * 4.0f;
            const float sqrtDenom = woDotH + actualEta * wiDotH;
            pdf *= fabsf(woDotH) / (sqrtDenom * sqrtDenom);
        }
        pdf *= isReflection ? F : 1.f - F;
        return clampf(pdf, 0.f, kFltMax);
    }

    // FalcorBSDF::eval (BxDF.hlsli:853-862): rgb and the average of the specular part
    PT_DEVICE float4 eval(float3 woWorld) const
    {
        const float3 wo = toLocal(woWorld);
        float3 diffuse = mk3(0.f), specular = mk3(0.f);
        if (pDR > 0.f) diffuse = diffuse + (1.f - specTrans) * (1.f - diffTrans) * drEval(wo);
        if (pDT > 0.f) diffuse = diffuse + (1.f - specTrans) * diffTrans * dtEval(wo);
        if (pSR > 0.f) specular = specular + (1.f - specTrans) * srEval(wo);
        if (pSRT > 0.f) specular = specular + specTrans * srtEval(wo);
        const float3 sum = diffuse + specular;
        return make_float4(sum.x, sum.y, sum.z, average(specular));
    }
    // FalcorBSDF::evalPdf (BxDF.hlsli:948-956)
    PT_DEVICE float pdfLocal(float3 wo) const
    {
        float pdf = 0.f;
        if (pDR > 0.f) pdf += pDR * drPdf(wo);
        if (pDT > 0.f) pdf += pDT * dtPdf(wo);
        if (pSR > 0.f) pdf += pSR * srPdf(wo);
        if (pSRT > 0.f) pdf += pSRT * srtPdf(wo);
        return pdf;
    }
    PT_DEVICE float pdf(float3 woWorld) const { return pdfLocal(toLocal(woWorld)); }

    // FalcorBSDF::sample (BxDF.hlsli:864-946), RecycleSelectSamples: u2 picks the lobe and is re-stretched for it
    PT_DEVICE bool sample(float u0, float u1, float u2, BsdfSample& r) const
    {
        float3 wo = mk3(0.f); float3 weight = mk3(0.f); float pdf = 0.f; uint lobe = kLobeDiffuseReflection; float lobeP = 0.f;
        bool valid = false;
        const float uSelect = u2;
        if (uSelect < pDR)
        {
            wo = sampleCosineHemisphereConcentric(u0, u1, pdf);
            lobe = kLobeDiffuseReflection;
            if (fminf(wi.z, wo.z) < kMinCosTheta) { weight = mk3(0.f); lobeP = 0.f; valid = false; }
            else { weight = drWeight(wo); lobeP = 1.f; valid = true; }
            weight = weight / pDR; weight = weight * ((1.f - specTrans) * (1.f - diffTrans));
            pdf *= pDR; lobeP *= pDR;
            if (pSR > 0.f) pdf += pSR * srPdf(wo);
            if (pSRT > 0.f) pdf += pSRT * srtPdf(wo);
        }
        else if (uSelect < pDR + pDT)
        {
            wo = sampleCosineHemisphereConcentric(u0, u1, pdf);
            wo.z = -wo.z;
            lobe = kLobeDiffuseTransmission;
            if (fminf(wi.z, -wo.z) < kMinCosTheta) { weight = mk3(0.f); lobeP = 0.f; valid = false; }
            else { weight = transAlbedo; lobeP = 1.f; valid = true; }
            weight = weight / pDT; weight = weight * ((1.f - specTrans) * diffTrans);
            pdf *= pDT; lobeP *= pDT;
            if (pSRT > 0.f) pdf += pSRT * srtPdf(wo);
        }
        else if (uSelect < pDR + pDT + pSR)
        {
            lobe = kLobeSpecularReflection; lobeP = 1.f;
            if (wi.z >= kMinCosTheta)
            {
                if (alphaRefl == 0.f)
                {
                    wo = mk3(-wi.x, -wi.y, wi.z); pdf = 0.f;
                    weight = fresnelSchlick3(specAlbedo, 1.f, wi.z);
                    lobe = kLobeDeltaReflection; valid = true;
                }
                else
                {
                    const float3 h = boundedVndfSample(alphaRefl, wi, u0, u1);
                    const float wiDotH = dot3(wi, h);
                    const float3 w = 2.f * wiDotH * h - wi;
                    wo = w;
                    if (w.z >= kMinCosTheta)
                    {
                        pdf = srPdf(w);
                        weight = srEval(w) / pdf;
                        valid = true;
                    }
                }
            }
            weight = weight / pSR; weight = weight * (1.f - specTrans);
            pdf *= pSR; lobeP *= pSR;
            if (pDR > 0.f) pdf += pDR * drPdf(wo);
            if (pSRT > 0.f) pdf += pSRT * srtPdf(wo);
        }
        else if (pSRT > 0.f)
        {
            const float lobeSample = clampf((uSelect - (pDR + pDT + pSR)) / pSRT, 0.f, kOneMinusEpsilon);
            lobe = kLobeSpecularReflection; lobeP = 1.f;
            if (wi.z >= kMinCosTheta)
            {
                if (alphaTrans == 0.f)
                {
                    float cosThetaT;
                    float F = fresnelDielectric(eta, wi.z, cosThetaT);
                    const bool isReflection = lobeSample < F;
                    lobeP = isReflection ? F : (1.f - F);
                    float actualEta = eta;
                    if (thin && !isReflection) { actualEta = 1.0f; F = fresnelDielectric(actualEta, wi.z, cosThetaT); }
                    pdf = 0.f;
                    weight = isReflection ? mk3(1.f) : transAlbedo;
                    wo = isReflection ? mk3(-wi.x, -wi.y, wi.z) : mk3(-wi.x * actualEta, -wi.y * actualEta, -cosThetaT);
                    lobe = isReflection ? kLobeDeltaReflection : kLobeDeltaTransmission;
                    valid = !(fabsf(wo.z) < kMinCosTheta || ((wo.z > 0.f) != isReflection));
                }
                else
                {
                    const float3 h = boundedVndfSample(alphaTrans, wi, u0, u1);
                    const float wiDotH = dot3(wi, h);
                    float cosThetaT;
                    float F = fresnelDielectric(eta, wiDotH, cosThetaT);
                    const bool isReflection = lobeSample < F;
                    float actualEta = eta;
                    if (thin && !isReflection) { actualEta = 1.0f; F = fresnelDielectric(actualEta, wi.z, cosThetaT); }    // sic: wi.z (BxDF.hlsli:532)
                    const float3 w = isReflection ? (2.f * wiDotH * h - wi) : ((actualEta * wiDotH - cosThetaT) * h - actualEta * wi);
                    wo = w;
                    if (!(fabsf(w.z) < kMinCosTheta || ((w.z > 0.f) != isReflection)))
                    {
                        lobe = isReflection ? kLobeSpecularReflection : kLobeSpecularTransmission;
                        pdf = srtPdf(w);
                        weight = pdf > 0.f ? srtEval(w) / pdf : mk3(0.f);
                        valid = true;
                    }
                }
            }
            weight = weight / pSRT; weight = weight * specTrans;
            pdf *= pSRT; lobeP *= pSRT;
            if (pDR > 0.f) pdf += pDR * drPdf(wo);
            if (pDT > 0.f) pdf += pDT * dtPdf(wo);
            if (pSR > 0.f) pdf += pSR * srPdf(wo);
        }
        if (!valid || (lobe & kLobeDelta) != 0) pdf = 0.0f;
        r.wo = fromLocal(wo); r.pdf = pdf; r.weight = weight; r.lobe = lobe; r.lobeP = lobeP;
        return valid;
    }
};

// FalcorBSDF::getLobes (BxDF.hlsli:831-851)
PT_DEVICE uint bsdfLobes(const BsdfParams& d)
{
    const bool isDelta = (d.roughness * d.roughness) < kMinGGXAlpha;
    uint lobes = isDelta ? kLobeDeltaReflection : kLobeSpecularReflection;
    if (anyPositive(d.diffuse) && d.specularTransmission < 1.f)
    {
        if (d.diffuseTransmission < 1.f) lobes |= kLobeDiffuseReflection;
        if (d.diffuseTransmission > 0.f) lobes |= kLobeDiffuseTransmission;
    }
    if (d.specularTransmission > 0.f) lobes |= (isDelta ? kLobeDeltaTransmission : kLobeSpecularTransmission);
    return lobes;
}

} // namespace pt
