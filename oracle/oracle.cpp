// ORACLE — test infrastructure only.  C entry points (ctypes) over the CPU restatement of RTXPT's PathTrace hot path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may build, load or call this library.
// The product (rtxpt_b200/) never links it.
//
// Parity status: RNG / packing are pinned against the reference's own C++ halves, and the path tracer's HLSL (material model, lights, NEE-AT sampler and frame-end passes,
// PathTracer::HandleHit / HandleMiss in all three passes, the per-pixel driver steps) against golden vectors made by compiling those headers in place as C++ (oracle/_ref,
// tests/golden/): the oracle_*_funcs / oracle_hit_funcs / oracle_baker_* mirrors below are what tests/test_oracle_golden.py compares.  Everything that depends on
// the DXR driver (BVH, traversal, intersection) or on TMU filtering, the Donut bridge's scene access, the environment bake and ReBLUR stay "parity unpinned" — the reference
// ships no runnable golden data for them.
#include "pt_path.h"
#include "reblur.h"
#include "pt_envbake.h"
#include "pt_tonemap.h"
#include "pt_skinning.h"
#include <cstdio>
#include <chrono>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace orc;

struct OracleCtx
{
    RtxptSceneDesc descCopy{}; std::vector<RtxptLightDesc> ownLights;
    Scene scene;
    Bvh2 bvh;
    LightTable lights;
    RtxptPathTracerConstants consts;
    bool haveConsts = false, haveView = false;
    float worldToClip[16] = {};
    double bvhBuildSeconds = 0;
    NeeatState neeat;              // NEE-AT temporal feedback (pt_neeat.h); used when consts.NEEATFeedback != 0
    NeeatState* neeatPtr() { return (haveConsts && consts.NEEATFeedback != 0 && consts.NEEType == 2 && neeat.W == consts.imageWidth && neeat.H == consts.imageHeight) ? &neeat : nullptr; }
};

extern "C" {

#define ORC_API __attribute__((visibility("default")))

ORC_API uint32_t oracle_hash32(uint32_t x) { return Hash32(x); }
ORC_API uint32_t oracle_hash32_combine(uint32_t s, uint32_t v) { return Hash32Combine(s, v); }
ORC_API float    oracle_hash32_to_float(uint32_t h) { return Hash32ToFloat(h); }
ORC_API uint32_t oracle_sobol(uint32_t index, uint32_t dim) { return bhos_sobol(index, dim); }
ORC_API uint32_t oracle_owen_scramble(uint32_t x, uint32_t seed) { return bhos_owen_scramble(x, seed); }
ORC_API uint32_t oracle_f32tof16(float f) { return f32tof16(f); }
ORC_API float    oracle_f16tof32(uint32_t h) { return f16tof32(h); }
ORC_API uint32_t oracle_pack_snorm8(float v) { return uint32_t(int(clampf(v, -1.0f, 1.0f) * 127.0f) & 0xff); }
ORC_API float    oracle_unpack_snorm8(uint32_t v) { return Unpack_R8_SNORM(v); }

// in: (pixelX, pixelY, vertexIndex, sampleIndex) ; out: 4 uniform draws (effect seed Base) then 4 LD draws (seed ScatterBSDF, raw u32 via float bits)
ORC_API void oracle_rng(const uint32_t* in, uint32_t count, uint32_t* out)
{
    for (uint32_t i = 0; i < count; i++)
    {
        SampleGeneratorVertexBase base = SampleGeneratorVertexBase::make((in[i * 4 + 0] << 16) | in[i * 4 + 1], in[i * 4 + 2], in[i * 4 + 3]);
        UniformSampleSequenceGenerator u = UniformSampleSequenceGenerator::make(base, SeedBase);
        for (int k = 0; k < 4; k++) out[i * 8 + k] = u.Next();
        float ld[4]; GenerateLD(4, base, SeedScatterBSDF, ld);
        for (int k = 0; k < 4; k++) out[i * 8 + 4 + k] = asuint(ld[k]);
    }
}

// Record layout shared with rtxpt_b200_debug_bsdf: 36 floats in, 16 floats out (see include/rtxpt_b200.h)
ORC_API void oracle_bsdf(const float* in, uint32_t count, float* out)
{
    for (uint32_t i = 0; i < count; i++)
    {
        const float* r = in + size_t(i) * 36; float* o = out + size_t(i) * 16;
        BSDFFrame f; f.V = f3(r[0], r[1], r[2]); f.N = f3(r[3], r[4], r[5]); f.T = f3(r[6], r[7], r[8]); f.B = f3(r[9], r[10], r[11]);
        float3 wo = f3(r[12], r[13], r[14]);
        float u[4] = { r[15], r[16], r[17], 0 };
        StandardBSDF b;
        b.data.diffuse = f3(r[18], r[19], r[20]); b.data.roughness = r[21]; b.data.specular = f3(r[22], r[23], r[24]); b.data.metallic = r[25];
        b.data.transmission = f3(r[26], r[27], r[28]); b.data.diffuseTransmission = r[29]; b.data.specularTransmission = r[30]; b.data.eta = r[31];
        f.thinSurface = r[32] != 0.0f; f.activeLobes = uint(r[33]); f.psdExclude = false;
        float4 e = b.eval(f, wo);
        o[0] = e.x; o[1] = e.y; o[2] = e.z; o[3] = e.w;
        o[4] = b.evalPdf(f, wo);
        BSDFSample s; bool valid = b.sample(f, u, s);
        o[5] = valid ? 1.0f : 0.0f; o[6] = s.wo.x; o[7] = s.wo.y; o[8] = s.wo.z; o[9] = s.pdf; o[10] = s.weight.x; o[11] = s.weight.y; o[12] = s.weight.z;
        o[13] = float(s.lobe); o[14] = s.lobeP; o[15] = float(b.getLobes());
    }
}

// oracle_bsdf's 16 floats followed by evalDeltaLobes and estimateSpecDiffBSDF: the 40-float layout of oracle/ref_kat_bsdf_main.cpp's "bsdf" mode
ORC_API void oracle_bsdf_ex(const float* in, uint32_t count, float* out)
{
    for (uint32_t i = 0; i < count; i++)
    {
        const float* r = in + size_t(i) * 36; float* o = out + size_t(i) * 40;
        oracle_bsdf(r, 1, o);
        BSDFFrame f; f.V = f3(r[0], r[1], r[2]); f.N = f3(r[3], r[4], r[5]); f.T = f3(r[6], r[7], r[8]); f.B = f3(r[9], r[10], r[11]);
        StandardBSDF b;
        b.data.diffuse = f3(r[18], r[19], r[20]); b.data.roughness = r[21]; b.data.specular = f3(r[22], r[23], r[24]); b.data.metallic = r[25];
        b.data.transmission = f3(r[26], r[27], r[28]); b.data.diffuseTransmission = r[29]; b.data.specularTransmission = r[30]; b.data.eta = r[31];
        f.thinSurface = r[32] != 0.0f; f.activeLobes = uint(r[33]); f.psdExclude = false;
        DeltaLobe lobes[cMaxDeltaLobes]; int n = 0; float nonDelta = 0;
        b.evalDeltaLobes(f, lobes, n, nonDelta);
        for (int k = 0; k < 2; k++) { float* d = o + 16 + k * 8; d[0] = lobes[k].thp.x; d[1] = lobes[k].thp.y; d[2] = lobes[k].thp.z; d[3] = lobes[k].probability; d[4] = lobes[k].dir.x; d[5] = lobes[k].dir.y; d[6] = lobes[k].dir.z; d[7] = float(lobes[k].transmission); }
        o[32] = nonDelta; o[33] = float(n);
        float3 de, se; b.estimateSpecDiffBSDF(de, se, f.N, f.V);
        o[34] = de.x; o[35] = de.y; o[36] = de.z; o[37] = se.x; o[38] = se.y; o[39] = se.z;
    }
}

// the scalar building blocks of the material model on the inputs oracle/ref_kat_bsdf_main.cpp derives from 8 uniforms per record; same 40-float output layout.  Slots of functions
// the live path does not use (and the oracle therefore does not restate: G1, separable masking, the unbounded VNDF) are left as NaN.
ORC_API void oracle_bsdf_funcs(const float* in, uint32_t count, float* out)
{
    const float nan = std::nanf("");
    for (uint32_t i = 0; i < count; i++)
    {
        const float* u = in + size_t(i) * 8; float* o = out + size_t(i) * 40;
        for (int k = 0; k < 40; k++) o[k] = nan;
        const float alpha = 0.0064f + u[0] * u[0] * 0.99f, cosI = 0.001f + 0.999f * u[1], cosO = 0.001f + 0.999f * u[2], eta = u[3] < 0.5f ? 1.0f / (1.0f + 1.2f * u[4]) : 1.0f + 1.2f * u[4];
        const float phi = 6.2831853f * u[5], sI = sqrtf(std::max(0.0f, 1.0f - cosI * cosI));
        const float3 wi = f3(sI * cosf(phi), sI * sinf(phi), cosI);
        o[0] = evalFresnelSchlick(u[6], 1.0f, cosI);
        const float3 fs = evalFresnelSchlick(f3(u[6], u[7], u[0]), 1.0f, cosO); o[1] = fs.x; o[2] = fs.y; o[3] = fs.z;
        float cosT = -1.0f; o[4] = evalFresnelDielectric(eta, cosI, cosT); o[5] = cosT;
        float cosT2 = -1.0f; o[6] = evalFresnelDielectric(eta, -cosI, cosT2); o[7] = cosT2;
        o[8] = evalNdfGGX(alpha, cosO); o[9] = evalLambdaGGX(alpha * alpha, cosI);
        o[11] = evalMaskingSmithGGXCorrelated(alpha, cosI, cosO);
        const float3 h = sampleGGX_BVNDF(alpha, wi, f2(u[6], u[7])); o[13] = h.x; o[14] = h.y; o[15] = h.z;
        o[16] = evalPdfGGX_BVNDF(alpha, wi, h);
        const float3 asi = approxSpecularIntegralGGX(f3(u[6], u[7], u[0]), alpha, cosI); o[21] = asi.x; o[22] = asi.y; o[23] = asi.z;
        float pdf = 0; const float3 ch = sample_cosine_hemisphere_concentric(f2(u[6], u[7]), pdf); o[24] = ch.x; o[25] = ch.y; o[26] = ch.z; o[27] = pdf;
        const float2 dk = sample_disk_concentric(f2(u[0], u[1])); o[28] = dk.x; o[29] = dk.y;
        const float3 ps = perp_stark(wi); o[30] = ps.x; o[31] = ps.y; o[32] = ps.z;
        const float2 oc = ndir_to_oct_equal_area_unorm(wi); o[33] = oc.x; o[34] = oc.y; const float3 od = oct_to_ndir_equal_area_unorm(f2(u[2], u[3])); o[35] = od.x; o[36] = od.y; o[37] = od.z;
        o[38] = Luminance(f3(u[0], u[1], u[2])); o[39] = Average(f3(u[0], u[1], u[2]));
    }
}

// Utils/Utils.hlsli as the oracle restates it, on the inputs and in the 24-float layout of oracle/ref_kat_bsdf_main.cpp's "utils" mode.  Slots of functions the oracle does
// not restate (LuminanceClamp, power heuristic, three-way MIS, WeightedAverage, RelativelyEqual, Reinhard: unused by the live path) are NaN.
ORC_API void oracle_utils_funcs(const float* in, uint32_t count, float* out)
{
    const float nan = std::nanf("");
    for (uint32_t i = 0; i < count; i++)
    {
        const float* u = in + size_t(i) * 8; float* o = out + size_t(i) * 24;
        for (int k = 0; k < 24; k++) o[k] = nan;
        const float3 dir = normalize(f3(2.0f * u[0] - 1.0f, 2.0f * u[1] - 1.0f, 2.0f * u[2] - 1.0f) + f3(1e-3f, 0.0f, 0.0f));
        const float n0 = 1.0f + floorf(u[4] * 4.0f), n1 = 1.0f + floorf(u[6] * 4.0f), p0 = u[5] * 3.0f, p1 = u[7] * 3.0f;
        o[3] = EvalMISBalance(n0, p0, n1, p1);
        const float2 eo = Encode_Oct(dir); o[6] = eo.x; o[7] = eo.y;
        const float3 dn = Decode_Oct(f2(u[0], u[1])); o[8] = dn.x; o[9] = dn.y; o[10] = dn.z;
        const uint p32 = NDirToOctUnorm32(dir); memcpy(&o[11], &p32, 4);
        const uint q32 = (uint(u[4] * 65534.0f) & 0xffffu) | (uint(u[5] * 65534.0f) << 16); const float3 d32 = OctToNDirUnorm32(q32); o[12] = d32.x; o[13] = d32.y; o[14] = d32.z;
        const uint p30 = NDirToOctUnorm30(dir); memcpy(&o[15], &p30, 4);
        const uint q30 = (uint(u[6] * 32767.0f) & 0x7fffu) | ((uint(u[7] * 32767.0f) & 0x7fffu) << 15); const float3 d30 = OctToNDirUnorm30(q30); o[16] = d30.x; o[17] = d30.y; o[18] = d30.z;
        o[19] = FastSqrt(u[0] * 10.0f); o[20] = FastACos(2.0f * u[1] - 1.0f);
    }
}

// PathTracerHelpers.hlsli as the oracle restates it, on the inputs and in the 16-float layout of oracle/ref_kat_bsdf_main.cpp's "helpers" mode (NaN: not restated as a function -
// the grazing-angle falloff is written inline in HandleNEE, the roughness-based cone growth is unused by the live path)
ORC_API void oracle_helper_funcs(const float* in, uint32_t count, float* out)
{
    const float nan = std::nanf("");
    for (uint32_t i = 0; i < count; i++)
    {
        const float* u = in + size_t(i) * 8; float* o = out + size_t(i) * 16;
        for (int k = 0; k < 16; k++) o[k] = nan;
        const float scale = u[6] < 0.25f ? 0.05f : (u[6] < 0.5f ? 1.0f : (u[6] < 0.75f ? 40.0f : 3000.0f));
        const float3 pos = f3(2.0f * u[0] - 1.0f, 2.0f * u[1] - 1.0f, 2.0f * u[2] - 1.0f) * scale;
        const float3 nrm = normalize(f3(2.0f * u[3] - 1.0f, 2.0f * u[4] - 1.0f, 2.0f * u[5] - 1.0f) + f3(0.0f, 1e-3f, 0.0f));
        const float3 ro = ComputeRayOrigin(pos, nrm); o[0] = ro.x; o[1] = ro.y; o[2] = ro.z;
        const float3 l = normalize(f3(2.0f * u[5] - 1.0f, 2.0f * u[6] - 1.0f, 2.0f * u[7] - 1.0f) + f3(1e-3f, 0.0f, 0.0f));
        const float from = 0.01f + 0.24f * u[7]; o[3] = saturate((dot(l, nrm) - from) / (2.0f * from));        // HandleNEE's fade-out term (pt_path.h), same expression
        const float pdf = u[3] < 0.05f ? 0.0f : u[3] * u[3] * 40.0f;
        o[7] = pdf == 0.0f ? 0.0f : ComputeRayConeSpreadAngleExpansionByScatterPDF(pdf);
        o[8] = ComputeNewScatterFireflyFilterK(lp(u[4]), pdf, u[5]);
        const float3 ff = FireflyFilter(f3(lp(u[0] * 8.0f), lp(u[1] * 8.0f), lp(u[2] * 8.0f)), lp(2.0f + u[3]), lp(u[4])); o[9] = ff.x; o[10] = ff.y; o[11] = ff.z;
        o[12] = FireflyFilterShort(u[0] * 8.0f, 2.0f + u[3], u[4]);
        o[13] = EvalMISBalance(1.0f, u[5] * 3.0f, 1.0f, u[7] * 3.0f);
        { const mat3 m = MatrixRotateFromTo(nrm, l); o[14] = m.r[0].x + m.r[0].y * 0.5f + m.r[0].z * 0.25f + m.r[1].x * 0.125f + m.r[1].y * 3.0f; o[15] = m.r[1].z + m.r[2].x * 0.5f + m.r[2].y * 0.25f + m.r[2].z * 3.0f; }
    }
}

// Lighting/PolymorphicLight.hlsli's emissive-triangle light as the oracle restates it (pt_lights.h), on the inputs and in the layout of ref_kat_bsdf_main.cpp's "lights" mode
ORC_API void oracle_light_funcs(const float* in, uint32_t count, float* out)
{
    for (uint32_t i = 0; i < count; i++)
    {
        const float* u = in + size_t(i) * 24; float* o = out + size_t(i) * 24;
        TriangleLight t; t.base = f3(u[0], u[1], u[2]); t.edge1 = f3(u[3], u[4], u[5]); t.edge2 = f3(u[6], u[7], u[8]); t.radiance = f3(u[9], u[10], u[11]);
        const PolymorphicLightInfo li = t.Store();
        const uint words[8] = { asuint(li.Center[0]), asuint(li.Center[1]), asuint(li.Center[2]), li.ColorTypeAndFlags, li.Direction1, li.Direction2, li.Scalars, li.LogRadiance };
        memcpy(o, words, 32);
        const TriangleLight r = TriangleLight::Create(li);
        const float3 viewer = f3(u[14], u[15], u[16]);
        const PolymorphicLightSample s = r.CalcSample(f2(u[12], u[13]), viewer);
        o[8] = s.Position.x; o[9] = s.Position.y; o[10] = s.Position.z; o[11] = s.Normal.x; o[12] = s.Normal.y; o[13] = s.Normal.z; o[14] = s.Radiance.x; o[15] = s.Radiance.y; o[16] = s.Radiance.z;
        o[17] = s.SolidAnglePdf; o[18] = r.CalcSolidAnglePdfForMIS(viewer, s.Position); o[19] = r.GetPower();
        o[20] = r.edge1.x; o[21] = r.edge1.y; o[22] = r.edge1.z; o[23] = r.surfaceArea;
    }
}

// the analytic sphere / spot light, layout of ref_kat_bsdf_main.cpp's "spheres" mode: the record is assembled with the oracle's packers, then Create / CalcSample / pdf / power
ORC_API void oracle_sphere_light_funcs(const float* in, uint32_t count, float* out)
{
    for (uint32_t i = 0; i < count; i++)
    {
        const float* u = in + size_t(i) * 24; float* o = out + size_t(i) * 24;
        PolymorphicLightInfo li = {}; PolymorphicLightInfoEx ex = {};
        li.Center[0] = u[0]; li.Center[1] = u[1]; li.Center[2] = u[2]; li.Scalars = f32tof16(u[3]);
        PackLightColor(f3(u[4], u[5], u[6]), li);
        li.ColorTypeAndFlags |= kLightTypeSphere << kPolymorphicLightTypeShift;
        if (u[7] > 0.5f)
        {
            li.ColorTypeAndFlags |= kPolymorphicLightShapingEnableBit | (u[8] > 0.5f ? kPolymorphicLightShapingUseMinFalloff : 0u);
            ex.PrimaryAxis = NDirToOctUnorm32(normalize(f3(u[9], u[10], u[11])));
            ex.CosConeAngleAndSoftness = f32tof16(u[12]) | (f32tof16(u[13]) << 16);
        }
        const uint words[12] = { asuint(li.Center[0]), asuint(li.Center[1]), asuint(li.Center[2]), li.ColorTypeAndFlags, li.Direction1, li.Direction2, li.Scalars, li.LogRadiance, ex.IesProfileIndex, ex.PrimaryAxis, ex.CosConeAngleAndSoftness, ex.UniqueID };
        memcpy(o, words, 48);
        const SphereLight s = SphereLight::Create(li, ex);
        const float3 viewer = f3(u[16], u[17], u[18]);
        PolymorphicLightSample r = s.CalcSample(f2(u[14], u[15]), viewer);
        if (r.SolidAnglePdf > 0) r.Radiance = r.Radiance * evaluateLightShaping(s.shaping, viewer, r.Position);       // PolymorphicLight::CalcSample, PolymorphicLight.hlsli:670-673
        o[12] = r.Position.x; o[13] = r.Position.y; o[14] = r.Position.z; o[15] = r.Normal.x; o[16] = r.Normal.y; o[17] = r.Normal.z; o[18] = r.Radiance.x; o[19] = r.Radiance.y; o[20] = r.Radiance.z;
        o[21] = r.SolidAnglePdf; o[22] = s.CalcSolidAnglePdfForMIS(viewer); o[23] = s.GetPower();
    }
}

// the six tone-mapping operators as the oracle restates them (pt_tonemap.h), in the layout of ref_kat_bsdf_main.cpp's "tonemap" mode: 8 floats in, rgb + luminance out
ORC_API void oracle_tonemap_ops(const float* in, uint32_t count, float* out)
{
    for (uint32_t i = 0; i < count; i++)
    {
        const float* u = in + size_t(i) * 8; float* o = out + size_t(i) * 4;
        tonemap::Params p; p.op = uint(u[3]); p.whiteMaxLuminance = u[4]; p.whiteScale = u[5];
        const float3 c = tonemap::toneMapOp(p, f3(u[0], u[1], u[2])); o[0] = c.x; o[1] = c.y; o[2] = c.z; o[3] = tonemap::luminance(f3(u[0], u[1], u[2]));
    }
}

// TexLODHelpers.hlsli as the oracle restates it (pt_scene.h), layout of ref_kat_bsdf_main.cpp's "texlod" mode: 40 floats in, 8 out (slot 6, addToSpreadAngle, is not restated as
// a function: the path adds to the angle through RayCone::make directly)
ORC_API void oracle_texlod_funcs(const float* in, uint32_t count, float* out)
{
    for (uint32_t i = 0; i < count; i++)
    {
        const float* u = in + size_t(i) * 40; float* o = out + size_t(i) * 8;
        const float3 v[3] = { f3(u[0], u[1], u[2]), f3(u[3], u[4], u[5]), f3(u[6], u[7], u[8]) }; const float2 t[3] = { f2(u[9], u[10]), f2(u[11], u[12]), f2(u[13], u[14]) };
        const float xf[12] = { u[15], u[16], u[17], 0.0f, u[18], u[19], u[20], 0.0f, u[21], u[22], u[23], 0.0f };
        o[0] = computeRayConeTriangleLODValue(v, t, xf);
        RayCone rc = RayCone::make(u[24], u[25]); o[1] = rc.getWidth(); o[2] = rc.getSpreadAngle();
        rc = rc.propagateDistance(u[26]); o[3] = rc.getWidth();
        const float3 dir = normalize(f3(u[27], u[28], u[29])), nrm = normalize(f3(u[30], u[31], u[32]));
        o[4] = rc.computeLOD(o[0], dir, nrm, true); o[5] = rc.computeLOD(o[0], dir, nrm, false);
        o[6] = RayCone::make(rc.getWidth(), rc.getSpreadAngle() + u[33]).getSpreadAngle(); o[7] = SafeLog2(u[34]);
    }
}

// InteriorList.hlsli as the oracle restates it (pt_path.h), layout of ref_kat_bsdf_main.cpp's "interior" mode: 12 crossings per record, 6 values after each
ORC_API void oracle_interior_list(const float* in, uint32_t count, float* out)
{
    for (uint32_t i = 0; i < count; i++)
    {
        const float* u = in + size_t(i) * 48; float* o = out + size_t(i) * 72;
        InteriorList il;
        for (int k = 0; k < 12; k++)
        {
            const uint mat = uint(u[4 * k]), prio = uint(u[4 * k + 1]); const bool entering = u[4 * k + 2] != 0.0f; const uint probe = uint(u[4 * k + 3]);
            const bool isTrue = il.isTrueIntersection(probe);
            il.handleIntersection(mat, prio, entering);
            const uint w[2] = { il.slots[0], il.slots[1] }; memcpy(o + 6 * k, w, 8);
            o[6 * k + 2] = float(il.getTopNestedPriority()); o[6 * k + 3] = float(int(il.getTopMaterialID())); o[6 * k + 4] = float(int(il.getNextMaterialID())); o[6 * k + 5] = isTrue ? 1.0f : 0.0f;
        }
    }
}

// LightSampler.hlsli's sampler side as the oracle restates it (pt_neeat.h, pt_path.h), layout of ref_kat_bsdf_main.cpp's "sampler" mode: one 16-light, 2 x 2-tile scenario and
// 8 queries per record (680 floats in, 8 x 16 out)
ORC_API void oracle_sampler_funcs(const float* in, uint32_t count, float* out)
{
    for (uint32_t i = 0; i < count; i++)
    {
        const float* r = in + size_t(i) * 680; float* o = out + size_t(i) * 128;
        LightTable lt; lt.samplingProxyCount = uint(r[0]); lt.proxyCounters.resize(16); lt.proxyIndices.resize(64); lt.lights.resize(16);
        for (int k = 0; k < 16; k++) lt.proxyCounters[k] = uint(r[8 + k]);
        for (int k = 0; k < 64; k++) lt.proxyIndices[k] = uint(r[24 + k]);
        NeeatState ns; ns.init(8, 8); ns.jitter[0] = uint(r[1]); ns.jitter[1] = uint(r[2]); ns.localToGlobalSampleRatio = r[5]; ns.settings.screenSpaceVsWorldSpaceThreshold = r[6];
        memcpy(ns.localSamplingBuffer.data(), r + 88, 512 * sizeof(uint));
        const uint candidateSampleCount = uint(r[3]), fullSamples = uint(r[4]);
        for (int q = 0; q < 8; q++)
        {
            const float* u = r + 600 + q * 10; float* d = o + q * 16;
            const uint px = uint(u[0]), py = uint(u[1]), lightIndex = uint(u[3]), flags = uint(u[4]); const bool isSSC = (flags & 1u) != 0;
            const uint tileAddress = LocalSamplingTilePos(ns, px, py);
            float pdf = 0; uint idx = SampleGlobal(lt, u[2], pdf); d[0] = float(idx); d[1] = pdf;
            idx = SampleLocal(ns, tileAddress, u[2], pdf); d[2] = float(idx); d[3] = pdf;
            d[4] = SampleGlobalPDF(lt, lightIndex); d[5] = SampleLocalPDF(ns, tileAddress, lightIndex);
            const uint localCount = isSSC ? ComputeCandidateSampleLocalCount(ns.localToGlobalSampleRatio, candidateSampleCount) : 0u, globalCount = candidateSampleCount - localCount;
            d[6] = float(localCount); d[7] = float(globalCount);
            d[8] = ComputeLightVsBSDF_MIS_ForBSDF(lt, lightIndex, lp(u[7]), u[8], fullSamples, &ns, (px << 16) | py, isSSC, candidateSampleCount);
            LightSample s; s.LightIndex = lightIndex; s.SelectionPdf = u[9]; s.SolidAnglePdf = u[8]; s.FromLocalDistribution = (flags & 2u) != 0; s.LightSampleableByBSDF = (flags & 4u) != 0;
            float thisPdf, otherPdf, thisCount, otherCount; ComputeLightSelectionPdfs(lt, &ns, tileAddress, s, localCount, globalCount, thisPdf, otherPdf, thisCount, otherCount);
            d[9] = otherPdf; d[10] = thisCount; d[11] = otherCount;
            d[12] = ComputeLightVsBSDF_MIS_ForLight(s, thisPdf, otherPdf, fullSamples, u[7]);
            InsertFeedbackFromNEE(ns, lt, px, py, isSSC, lightIndex, u[5], u[6]);
            d[13] = ns.feedback.weight[py * 8 + px]; memcpy(d + 14, &ns.feedback.candidate[py * 8 + px], 4);
            d[15] = IsScreenSpaceCoherentHeuristic(ns.settings.screenSpaceVsWorldSpaceThreshold, u[8], u[5]) ? 1.0f : 0.0f;
        }
    }
}

// NEE-AT's feedback passes as the oracle restates them (pt_neeat.h), layout of oracle/ref_kat_baker_main.cpp (3056 floats in, 4241 out): P0, P1a, P1b, P2 (FillTile),
// ClearFeedbackHistory and P3 (the sorted lists with their run lengths) on a 16 x 16 image with 3 x 3 tiles, in UpdateEnd's order
ORC_API void oracle_baker_feedback(const float* in, uint32_t count, float* out)
{
    for (uint32_t i = 0; i < count; i++)
    {
        const float* r = in + size_t(i) * 3056; float* o = out + size_t(i) * 4241;
        const uint W = 16, H = 16, P = W * H;
        NeeatState s; s.init(W, H);
        const uint total = uint(r[0]); s.historicTotalLightCount = uint(r[1]); s.updateCounter = uint(r[2]); s.jitter[0] = uint(r[3]); s.jitter[1] = uint(r[4]); s.jitterPrev[0] = uint(r[5]); s.jitterPrev[1] = uint(r[6]);
        s.lastFrameTemporalFeedbackAvailable = r[7] != 0.0f; s.lastFrameLocalSamplesAvailable = r[8] != 0.0f;
        s.settings.depthDisocclusionThreshold = r[10]; s.settings.enableMotionReprojection = r[11] != 0.0f; s.settings.reservoirHistoryDropoff = r[12];
        LightTable lt; lt.lights.resize(total); lt.samplingProxyCount = uint(r[9]); lt.proxyIndices.resize(64); for (int k = 0; k < 64; k++) lt.proxyIndices[k] = uint(r[48 + k]);
        s.pastToCurrent.resize(32); memcpy(s.pastToCurrent.data(), r + 32, 64); for (int k = 16; k < 32; k++) s.pastToCurrent[k] = RTXPT_INVALID_LIGHT_INDEX;
        memcpy(s.feedback.weight.data(), r + 112, P * 4); memcpy(s.feedback.candidate.data(), r + 368, P * 4); memcpy(s.historyDepth.data(), r + 624, P * 4);
        std::vector<float> depth(r + 880, r + 880 + P); std::vector<uint16_t> motion(size_t(P) * 4, 0);
        for (uint k = 0; k < P; k++) { motion[4 * k] = uint16_t(f32tof16(r[1136 + 3 * k])); motion[4 * k + 1] = uint16_t(f32tof16(r[1137 + 3 * k])); motion[4 * k + 2] = uint16_t(f32tof16(r[1138 + 3 * k])); }
        memcpy(s.localSamplingBuffer.data(), r + 1904, 1152 * 4);
        ProcessFeedbackHistoryP0(s, total);
        memcpy(o, s.feedback.weight.data(), P * 4); memcpy(o + 256, s.feedback.candidate.data(), P * 4); for (uint k = 0; k < 17; k++) o[512 + k] = k <= total ? float(s.feedbackCounters[k]) : 0.0f;
        ProcessFeedbackHistoryP1a(s, lt, depth.data(), motion.data());
        memcpy(o + 529, s.blended.weight.data(), 64 * 4); memcpy(o + 593, s.blended.candidate.data(), 64 * 4);
        ProcessFeedbackHistoryP1b(s, lt, depth.data(), motion.data());
        memcpy(o + 657, s.scratch.weight.data(), P * 4); memcpy(o + 913, s.scratch.candidate.data(), P * 4);
        for (uint ty = 0; ty < 3; ty++) for (uint tx = 0; tx < 3; tx++)
        {
            uint list[NEEAT_LOCAL_PROXY_COUNT]; FillTile(s, tx, ty, list);
            for (uint k = 0; k < NEEAT_LOCAL_PROXY_COUNT; k++) { const uint w = PackMiniListLightAndCount(list[k], 1); memcpy(o + 1169 + s.tileBaseAddress(tx, ty) + k, &w, 4); }
        }
        ClearFeedbackHistory(s, depth.data());
        memcpy(o + 2321, s.feedback.weight.data(), P * 4); memcpy(o + 2577, s.feedback.candidate.data(), P * 4); memcpy(o + 2833, s.historyDepth.data(), P * 4);
        ProcessFeedbackHistoryP2P3(s);      // (reads the scratch and blended images, which ClearFeedbackHistory left alone)
        memcpy(o + 3089, s.localSamplingBuffer.data(), 1152 * 4);
    }
}

// ComputeProxyCounts as the oracle restates it (pt_neeat.h: RebuildGlobalProxies), layout of ref_kat_baker_main.cpp's "counts" mode (64 floats in, 40 out)
ORC_API void oracle_baker_counts(const float* in, uint32_t count, float* out)
{
    for (uint32_t i = 0; i < count; i++)
    {
        const float* r = in + size_t(i) * 64; float* o = out + size_t(i) * 40;
        for (int k = 0; k < 40; k++) o[k] = 0.0f;
        const uint n = uint(r[0]);
        NeeatState s; s.lastFrameTemporalFeedbackAvailable = r[1] != 0.0f; s.globalFeedbackUseWeight = r[3]; s.currentWeightsSum = r[5];
        s.currentWeights.assign(r + 8, r + 8 + n); s.feedbackCounters.resize(n + 1); for (uint k = 0; k <= n; k++) s.feedbackCounters[k] = uint(r[24 + k]);
        s.validFeedbackCount = uint(r[2]) - s.feedbackCounters[n];
        LightTable lt; lt.lights.resize(n); lt.proxyCounters.resize(n);
        RebuildGlobalProxies(s, lt, uint(r[4]) == 0 ? 0u : 2u);
        uint offset = 0;
        for (uint k = 0; k < n; k++) { o[k] = float(lt.proxyCounters[k]); o[17 + k] = float(offset); offset += lt.proxyCounters[k]; }
        for (uint k = n; k < 16; k++) o[k] = r[24 + k];      // the slots behind the lights keep what they held (slot n: reservoirs without a light)
        o[16] = float(lt.samplingProxyCount); o[33] = float(offset);
    }
}

// the environment-quad light, layout of ref_kat_bsdf_main.cpp's "envquads" mode: Store, Create, the sample HandleNEE draws from it (pt_path.h), pdf, power
ORC_API void oracle_envquad_light_funcs(const float* in, uint32_t count, float* out)
{
    for (uint32_t i = 0; i < count; i++)
    {
        const float* u = in + size_t(i) * 24; float* o = out + size_t(i) * 24;
        EnvironmentQuadLight e; e.NodeX = uint(u[0]); e.NodeY = uint(u[1]); e.NodeDim = uint(u[2]); e.Weight = u[3]; e.Radiance = f3(u[4], u[5], u[6]);
        const PolymorphicLightInfo li = e.Store();
        uint words[12] = {}; memcpy(words, &li, 32); words[11] = uint(u[7]);
        memcpy(o, words, 48);
        RtxptPathTracerConstants c; memset(&c, 0, sizeof(c));
        const bool rotated = u[8] != 0.0f || u[9] != 0.0f || u[10] != 0.0f;
        for (int a = 0; a < 3; a++) for (int k = 0; k < 3; k++) { const float v = rotated ? u[8 + 3 * a + k] : (a == k ? 1.0f : 0.0f); c.envMap.Transform[a * 4 + k] = v; c.envMap.InvTransform[k * 4 + a] = v; }
        PathTracerCtx x; x.c = &c;
        const EnvironmentQuadLight q = EnvironmentQuadLight::Create(li);
        const float3 viewer = f3(u[19], u[20], u[21]);
        const float2 subTexelPos = f2((float(q.NodeX) + u[17]) / float(q.NodeDim), (float(q.NodeY) + u[18]) / float(q.NodeDim));
        const float3 worldDir = envToWorld(x, oct_to_ndir_equal_area_unorm(subTexelPos));
        const float3 position = viewer + worldDir * DISTANT_LIGHT_DISTANCE, normal = -worldDir;
        o[12] = position.x; o[13] = position.y; o[14] = position.z; o[15] = normal.x; o[16] = normal.y; o[17] = normal.z; o[18] = q.Radiance.x; o[19] = q.Radiance.y; o[20] = q.Radiance.z;
        o[21] = q.SolidAnglePdf(); o[22] = q.SolidAnglePdf(); o[23] = q.Weight;
    }
}

// PathTracer::HandleHit on one path vertex as the oracle restates it (pt_path.h: HandleHitSurface, HandleNEE, GenerateScatterRay, HandleRussianRoulette, nested dielectrics), layout
// of ref_kat_bsdf_main.cpp's "hit" mode (1024 floats in, 128 out).  The scene side is data, as behind the stub bridge there: the surface comes from the record, materials are the
// IoR / absorption table, a shadow ray is answered by the same function of its bits (ShimVisibilityRule in oracle/ref_bridge_stub.h, restated here).  mode: 0 reference, 2 FILL
struct HitMirrorVisibility { uint queries = 0; float3 o = f3(0), d = f3(0); float tMax = 0; bool last = false; };
static float3 hitMirrorEnvCube(float3 d, float lod) { const float k = exp2f(-lod); return f3((0.5f + 0.5f * d.x) * k, (0.5f + 0.25f * d.y) * k, (0.75f + 0.25f * d.z) * k); }
static const std::vector<uint>& hitMirrorEnvLookup()
{
    static std::vector<uint> m;
    if (m.empty()) { m.resize(size_t(IMPORTANCE_MAP_DIM) * IMPORTANCE_MAP_DIM); for (uint y = 0; y < IMPORTANCE_MAP_DIM; y++) for (uint x = 0; x < IMPORTANCE_MAP_DIM; x++) m[size_t(y) * IMPORTANCE_MAP_DIM + x] = ((x >> 6) + (y >> 6) * 3u) & 3u; }
    return m;
}
struct HitMirrorCamera { float3 pos, base, dx, dy; };
static void hitMirrorCameraRay(uint px, uint py, float3& origin, float3& dir, void* user)
{
    const HitMirrorCamera& c = *static_cast<const HitMirrorCamera*>(user);
    origin = c.pos; dir = normalize(c.base + c.dx * float(px) + c.dy * float(py));
}
static float3 hitMirrorMotionVector(float3 posW, float3 prevPosW) { return (prevPosW - posW) * 0.5f; }
static bool hitMirrorVisibility(float3 o, float3 d, float tMax, void* user)
{
    HitMirrorVisibility& v = *static_cast<HitMirrorVisibility*>(user);
    const uint h = asuint(o.x) ^ (asuint(o.y) >> 1) ^ (asuint(o.z) >> 2) ^ asuint(d.x) ^ (asuint(d.y) >> 1) ^ (asuint(d.z) >> 2) ^ asuint(tMax);
    v.queries++; v.o = o; v.d = d; v.tMax = tMax; v.last = (h & 3u) != 0u;
    return v.last;
}
ORC_API void oracle_hit_funcs(const float* in, uint32_t count, float* out, uint32_t mode)
{
    for (uint32_t i = 0; i < count; i++)
    {
        const float* r = in + size_t(i) * 1024; float* o = out + size_t(i) * 128;
        for (int k = 0; k < 128; k++) o[k] = 0.0f;
        // surface
        SurfaceData sf; ShadingData& sd = sf.sd;
        sd.posW = f3(r[28], r[29], r[30]); sd.faceNCorrected = f3(r[31], r[32], r[33]); sd.V = -f3(r[23], r[24], r[25]); sd.N = f3(r[34], r[35], r[36]); sd.T = f3(r[37], r[38], r[39]);
        sd.B = f3(r[40], r[41], r[42]); sd.vertexN = f3(r[43], r[44], r[45]); sd.frontFacing = r[46] != 0.0f; sd.nestedPriority = uint(r[47]); sd.activeLobes = uint(r[48]); sd.thinSurface = r[49] != 0.0f;
        sd.psdExclude = r[50] != 0.0f; sd.psdBlockMotionVectorsAtSurface = r[57] != 0.0f; sd.psdDominantDeltaLobeP1 = uint(r[58]); sd.materialID = uint(r[51]); sd.IoR = r[52]; sd.shadowNoLFadeout = r[53];
        sd.emission = f3(r[54], r[55], r[56]);
        const float* b = r + 42;
        sf.bsdf.data.diffuse = f3(b[18], b[19], b[20]); sf.bsdf.data.roughness = b[21]; sf.bsdf.data.specular = f3(b[22], b[23], b[24]); sf.bsdf.data.metallic = b[25];
        sf.bsdf.data.transmission = f3(b[26], b[27], b[28]); sf.bsdf.data.diffuseTransmission = b[29]; sf.bsdf.data.specularTransmission = b[30]; sf.bsdf.data.eta = b[31];
        sf.interiorIoR = r[74]; sf.neeTriangleLightIndex = r[75] < 0 ? 0xFFFFFFFFu : uint(r[75]); sf.neeAnalyticLightIndex = r[76] < 0 ? 0xFFFFFFFFu : uint(r[76]); sf.prevPosW = f3(r[77], r[78], r[79]);
        // constants, materials
        RtxptPathTracerConstants c; memset(&c, 0, sizeof(c));
        c.imageWidth = c.imageHeight = 8; c.bounceCount = uint(r[80]); c.diffuseBounceCount = uint(r[81]); c.NEEEnabled = 1; c.NEEType = 2; c.NEECandidateSamples = uint(r[83]); c.NEEFullSamples = uint(r[84]);
        c.fireflyFilterThreshold = r[85]; c.enableRussianRoulette = 1; c.enableLDSamplerForBSDF = 1; c.nestedDielectricsQuality = 1; c.EnvironmentMapDiffuseSampleMIPLevel = r[93]; c.NEEATFeedback = 1;
        for (int a = 0; a < 3; a++) for (int k = 0; k < 3; k++) { c.envMap.Transform[a * 4 + k] = r[950 + 3 * a + k]; c.envMap.InvTransform[k * 4 + a] = r[950 + 3 * a + k]; }
        for (int k = 0; k < 3; k++) c.envMap.ColorMultiplier[k] = r[959];
        RtxptMaterialData mats[8]; memset(mats, 0, sizeof(mats));
        for (int m = 0; m < 8; m++) { mats[m].IoR = r[96 + m]; for (int k = 0; k < 3; k++) mats[m].VolumeAttenuationColor[k] = r[104 + 3 * m + k]; mats[m].VolumeAttenuationDistance = r[128 + m]; }
        RtxptSceneDesc desc; memset(&desc, 0, sizeof(desc)); desc.materials = mats; desc.materialCount = 8;
        Scene sc; sc.desc = &desc;
        // lights
        static thread_local LightTable lt;      // (kept across records: its 1024 x 1024 environment lookup map is filled once)
        lt.samplingProxyCount = uint(r[90]); lt.proxyCounters.resize(16); lt.proxyIndices.resize(64); lt.lights.resize(16); lt.lightsEx.resize(16); lt.exBase = 0; lt.analyticLightCount = 16;
        for (int k = 0; k < 16; k++) lt.proxyCounters[k] = uint(r[136 + k]);
        for (int k = 0; k < 64; k++) lt.proxyIndices[k] = uint(r[152 + k]);
        for (int k = 0; k < 16; k++) { memcpy(&lt.lights[k], r + 728 + 12 * k, 32); memcpy(&lt.lightsEx[k], r + 728 + 12 * k + 8, 16); }
        lt.envEnabled = true; if (lt.envLookupMap.empty()) lt.envLookupMap = hitMirrorEnvLookup();
        NeeatState ns; ns.init(8, 8); ns.jitter[0] = uint(r[88]); ns.jitter[1] = uint(r[89]); ns.localToGlobalSampleRatio = r[86]; ns.settings.screenSpaceVsWorldSpaceThreshold = r[91];
        ns.temporalFeedbackRequired = r[92] != 0.0f; memcpy(ns.localSamplingBuffer.data(), r + 216, 512 * sizeof(uint));
        // the vertex
        HitMirrorVisibility vis;
        PathTracerCtx x; x.scene = &sc; x.bvh = nullptr; x.lights = &lt; x.c = &c; x.sampleIndex = uint(r[82]); x.stats = nullptr; x.mode = mode; x.neeat = &ns;
        x.envEvalOverride = hitMirrorEnvCube; x.visibilityOverride = hitMirrorVisibility; x.visibilityUser = &vis; x.noisyRadianceAttenuationOverride = r[87];
        uint payload[20]; memcpy(payload, r, 80);
        PathState path = unpackPayload(payload); const uint payloadIn14 = payload[14];
        // FILL: the stable planes of an 8 x 8 image, the pixel's entries from the record
        const uint pxi = (path.id >> 16) & 7u, pyi = path.id & 7u;
        std::vector<RtxptStablePlane> planes; std::vector<uint> header, throughput; std::vector<float> specHitT, depth; std::vector<uint16_t> stableRadiance, motion; RtxptRealtimeConstants rtc; memset(&rtc, 0, sizeof(rtc)); RealtimeTargets rtt;
        HitMirrorCamera cam; const float identity[16] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1 };
        if (mode != MODE_REFERENCE)
        {
            rtt.width = rtt.height = 8; rtt.lineStride = GenericTSComputeLineStride(8, 8); rtt.planeStride = GenericTSComputePlaneStride(8, 8);
            planes.assign(size_t(3) * rtt.planeStride, RtxptStablePlane{}); header.assign(4 * 64, 0xFFFFFFFFu); specHitT.assign(64, 0.0f);
            rtt.planes = planes.data(); rtt.header = header.data(); rtt.specularHitT = specHitT.data(); rtc.subSampleCount = 1; rtc.activeStablePlaneCount = 3; rtt.rt = &rtc;
            for (uint k = 0; k < 4; k++) memcpy(&rtt.hdr(pxi, pyi, k), r + 920 + k, 4);
            for (uint k = 0; k < 3; k++) memcpy(planes[rtt.PixelToAddress(pxi, pyi, k)].PackedNoisyRadianceAndSpecAvg, r + 924 + 2 * k, 8);
            specHitT[pyi * 8 + pxi] = r[930];
            if (r[27] >= 3.0f) for (uint k = 0; k < 3; k++) memcpy(&planes[rtt.PixelToAddress(pxi, pyi, k)], r + 960 + 20 * k, 80);
            stableRadiance.assign(64 * 4, 0); motion.assign(64 * 4, 0); depth.assign(64, -1.0f); throughput.assign(64, 0u);
            rtt.stableRadiance = stableRadiance.data(); rtt.motionVectors = motion.data(); rtt.depth = depth.data(); rtt.throughput = throughput.data();
            for (int k = 0; k < 4; k++) stableRadiance[(pyi * 8 + pxi) * 4 + k] = uint16_t(f32tof16(r[946 + k]));
            rtc.maxStablePlaneVertexDepth = uint(r[943]); rtc.allowPrimarySurfaceReplacement = uint(r[944]);
            cam.pos = f3(r[931], r[932], r[933]); cam.base = f3(r[934], r[935], r[936]); cam.dx = f3(r[937], r[938], r[939]); cam.dy = f3(r[940], r[941], r[942]);
            x.cameraRayOverride = hitMirrorCameraRay; x.cameraRayUser = &cam; rtt.motionVectorOverride = hitMirrorMotionVector; x.worldToClip = identity;
            x.sp = &rtt;
        }
        const float3 rayOrigin = f3(r[20], r[21], r[22]), rayDir = f3(r[23], r[24], r[25]);
        float tMin = 0.0f, tMax = 0.0f;
        if (r[27] == 1.0f) HandleMiss(x, path, rayDir, r[26]);          // the payload's origin is the ray's
        else if (r[27] == 0.0f) { UpdatePathTravelled(path, r[26]); HandleHitSurface(x, path, rayOrigin, rayDir, r[26], sf); }
        else if (r[27] == 2.0f) path = EmptyPathInitialize(x, path.id >> 16, path.id & 0xFFFF, r[1020]);
        else if (r[27] == 3.0f) { path = EmptyPathInitialize(x, path.id >> 16, path.id & 0xFFFF, r[1020]); tMax = kMaxRayTravel; FirstHitFromVBuffer(x, path, tMin, tMax); }
        else if (r[27] == 4.0f) postProcessHit(x, path);
        o[120] = tMin; o[121] = tMax;
        packPayload(path, payload); memcpy(o, payload, 80);
        o[20] = float(vis.queries); o[21] = vis.o.x; o[22] = vis.o.y; o[23] = vis.o.z; o[24] = vis.d.x; o[25] = vis.d.y; o[26] = vis.d.z; o[27] = vis.tMax; o[28] = vis.last ? 1.0f : 0.0f;
        const bool rejectedFalseHit = path.getCounter(CTR_RejectedHits) != ((payloadIn14 >> 8) & 0xFFu);
        const bool miss = r[27] == 1.0f, vertexOp = r[27] <= 1.0f;
        if (mode == MODE_REFERENCE && miss) { const float3 v = rayOrigin + rayDir * r[26]; o[31] = 1.0f; o[32] = v.x; o[33] = v.y; o[34] = v.z; }      // Bridge::ExportNonSurface( path, rayOrigin + rayDir * rayTCurrent )
        if (mode == MODE_REFERENCE && vertexOp && !miss && !rejectedFalseHit) { o[29] = 1.0f; o[30] = path.sceneLength; }      // Bridge::ExportSurface( path, surface, path.GetSceneLength() ): once per accepted hit
        if (mode != MODE_REFERENCE)
        {
            o[37] = specHitT[pyi * 8 + pxi];
            for (uint k = 0; k < 3; k++) memcpy(o + 41 + 2 * k, planes[rtt.PixelToAddress(pxi, pyi, k)].PackedNoisyRadianceAndSpecAvg, 8);
            for (uint k = 0; k < 4; k++) memcpy(o + 47 + k, &rtt.hdr(pxi, pyi, k), 4);
            for (int k = 0; k < 4; k++) o[52 + k] = f16tof32(stableRadiance[(pyi * 8 + pxi) * 4 + k]);
            for (uint k = 0; k < 3; k++) memcpy(o + 56 + 20 * k, &planes[rtt.PixelToAddress(pxi, pyi, k)], 80);
            if (mode == MODE_BUILD_STABLE_PLANES) o[miss ? 31 : 29] = depth[pyi * 8 + pxi] != -1.0f ? 1.0f : 0.0f;        // Bridge::ExportSurface / ExportNonSurface from the dominant base plane
        }
        const uint px = (path.id >> 16) & 7u, py = path.id & 7u;
        o[39] = ns.feedback.weight[py * 8 + px]; memcpy(o + 40, &ns.feedback.candidate[py * 8 + px], 4);
    }
}

ORC_API void* oracle_create(const RtxptSceneDesc* desc)
{
    OracleCtx* c = new OracleCtx();
    c->descCopy = *desc;                 // a private copy of the table of pointers: oracle_set_lights swaps the light array
    c->scene.init(&c->descCopy);
    auto t0 = std::chrono::steady_clock::now();
    c->bvh.build(c->scene);
    c->bvhBuildSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return c;
}
ORC_API void oracle_destroy(void* p) { delete (OracleCtx*)p; }
ORC_API double oracle_bvh_build_seconds(void* p) { return ((OracleCtx*)p)->bvhBuildSeconds; }
ORC_API uint32_t oracle_triangle_count(void* p) { return uint32_t(((OracleCtx*)p)->bvh.tris.size()); }

ORC_API int oracle_set_constants(void* p, const RtxptPathTracerConstants* consts)
{
    OracleCtx* c = (OracleCtx*)p;
    bool rebuildEnv = !c->haveConsts;
    c->consts = *consts; c->haveConsts = true;
    bakeLights(c->scene, c->consts, c->lights, rebuildEnv);
    return 0;
}

// the scene's analytic lights changed (rtxpt_b200_update_lights): re-bake the light list; NEE-AT's next update_begin maps last frame's feedback onto it
ORC_API int oracle_set_lights(void* p, const RtxptLightDesc* lights, uint32_t count)
{
    OracleCtx* c = (OracleCtx*)p; if (count && !lights) return -1;
    c->ownLights.assign(lights, lights + count); c->descCopy.lights = c->ownLights.data(); c->descCopy.lightCount = count;
    if (c->haveConsts) bakeLights(c->scene, c->consts, c->lights, false);
    return 0;
}

// ---- NEE-AT temporal feedback (pt_neeat.h).  Frame order as in Sample.cpp: set_constants; neeat_update_begin; [BUILD pass -> depth, motion]; neeat_update_end( depth, motion ); render.
ORC_API int oracle_neeat_reset(void* p)
{
    OracleCtx* c = (OracleCtx*)p; if (!c->haveConsts) return -1;
    c->neeat.init(c->consts.imageWidth, c->consts.imageHeight);
    return 0;
}
ORC_API int oracle_neeat_update_begin(void* p)
{
    OracleCtx* c = (OracleCtx*)p; if (!c->neeatPtr() || c->lights.IsEmpty()) return -1;
    NeeatUpdateBegin(c->neeat, c->lights, c->consts.NEEType, c->consts.NEEATImportanceBoost, c->haveView ? c->worldToClip : nullptr);
    return 0;
}
// depth: R32F guide; motion: RGBA16F guide (pixels) or NULL
ORC_API int oracle_neeat_update_end(void* p, const float* depth, const uint16_t* motion)
{
    OracleCtx* c = (OracleCtx*)p; if (!c->neeatPtr() || !depth) return -1;
    NeeatUpdateEnd(c->neeat, c->lights, depth, motion);
    return 0;
}
// what: 0 feedback weight (f32 WxH), 1 feedback candidate (u32 WxH), 2 scratch weight, 3 scratch candidate, 4 blended weight, 5 blended candidate (ceil(W/2) x ceil(H/2)),
//       6 local sampling buffer (u32 tilesX*tilesY*128), 7 proxy counters (u32 lightCount), 8 control { tilesX, tilesY, jitterX, jitterY, samplingProxyCount, updateCounter, available, validFeedbackCount }
ORC_API int oracle_neeat_get(void* p, int what, void* out, size_t bytes)
{
    OracleCtx* c = (OracleCtx*)p; const NeeatState& s = c->neeat;
    const void* src = nullptr; size_t n = 0; uint32_t ctl[8];
    switch (what)
    {
    case 0: src = s.feedback.weight.data(); n = s.feedback.weight.size() * 4; break;
    case 1: src = s.feedback.candidate.data(); n = s.feedback.candidate.size() * 4; break;
    case 2: src = s.scratch.weight.data(); n = s.scratch.weight.size() * 4; break;
    case 3: src = s.scratch.candidate.data(); n = s.scratch.candidate.size() * 4; break;
    case 4: src = s.blended.weight.data(); n = s.blended.weight.size() * 4; break;
    case 5: src = s.blended.candidate.data(); n = s.blended.candidate.size() * 4; break;
    case 6: src = s.localSamplingBuffer.data(); n = s.localSamplingBuffer.size() * 4; break;
    case 7: src = c->lights.proxyCounters.data(); n = c->lights.proxyCounters.size() * 4; break;
    case 8: ctl[0] = s.tilesX; ctl[1] = s.tilesY; ctl[2] = s.jitter[0]; ctl[3] = s.jitter[1]; ctl[4] = c->lights.samplingProxyCount; ctl[5] = s.updateCounter; ctl[6] = s.lastFrameTemporalFeedbackAvailable; ctl[7] = s.validFeedbackCount; src = ctl; n = sizeof(ctl); break;
    case 13: src = s.currentWeights.data(); n = s.currentWeights.size() * 4; break;          // boosted weights of the frame (f32 lightCount)
    case 14: src = &s.currentWeightsSum; n = 4; break;
    case 15: src = c->lights.lights.data(); n = c->lights.lights.size() * sizeof(PolymorphicLightInfo); break;      // the light records (32 B each)
    case 9: src = c->lights.weights.data(); n = c->lights.weights.size() * 4; break;            // power-based light weights (f32 lightCount)
    case 10: src = &c->lights.weightsSum; n = 4; break;
    case 11: src = c->lights.proxyIndices.data(); n = c->lights.proxyIndices.size() * 4; break;
    case 12: { static uint32_t cnt; cnt = uint32_t(c->lights.lights.size()); src = &cnt; n = 4; break; }
    default: return -1;
    }
    if (bytes < n) return -2;
    memcpy(out, src, n);
    return int(n);
}
// sampler-side functions of the local tile sampler for pixel (px, py): out = { SampleLocal's light, its pdf, SampleLocalPDF( lightForPdf ) }
ORC_API int oracle_neeat_sample_local(void* p, uint32_t px, uint32_t py, float rnd, uint32_t lightForPdf, float* out)
{
    OracleCtx* c = (OracleCtx*)p; const NeeatState& s = c->neeat; if (s.W == 0) return -1;
    const uint tile = LocalSamplingTilePos(s, px, py); float pdf = 0;
    const uint light = SampleLocal(s, tile, rnd, pdf);
    out[0] = float(light); out[1] = pdf; out[2] = SampleLocalPDF(s, tile, lightForPdf);
    return 0;
}
// test hook: overwrite the feedback reservoirs (same layouts as `what` 0 / 1)
ORC_API int oracle_neeat_set_feedback(void* p, const float* weight, const uint32_t* candidate)
{
    OracleCtx* c = (OracleCtx*)p; NeeatState& s = c->neeat; if (s.W == 0) return -1;
    memcpy(s.feedback.weight.data(), weight, s.feedback.weight.size() * 4); memcpy(s.feedback.candidate.data(), candidate, s.feedback.candidate.size() * 4);
    s.feedbackBufferFilled = true;
    return 0;
}

ORC_API int oracle_trace_rays(void* p, const RtxptRay* rays, uint32_t count, int anyHit, RtxptHit* out)
{
    OracleCtx* c = (OracleCtx*)p;
    #pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < int(count); i++)
    {
        const RtxptRay& r = rays[i];
        Hit h = c->bvh.trace(c->scene, f3(r.origin[0], r.origin[1], r.origin[2]), f3(r.dir[0], r.dir[1], r.dir[2]), r.tMin, r.tMax, anyHit != 0);
        RtxptHit& o = out[i];
        if (h.valid()) { const Tri& t = c->bvh.tris[h.triId]; o.t = h.t; o.u = h.u; o.v = h.v; o.instanceIndex = t.instanceIndex; o.geometryIndex = t.geometryIndex; o.primitiveIndex = t.primitiveIndex; }
        else { o.t = -1.0f; o.u = o.v = 0; o.instanceIndex = o.geometryIndex = o.primitiveIndex = 0xFFFFFFFFu; }
    }
    return 0;
}

ORC_API int oracle_set_view(void* p, const float* worldToClip16)
{
    OracleCtx* c = (OracleCtx*)p;
    memcpy(c->worldToClip, worldToClip16, 64); c->haveView = true;
    return 0;
}
// guide buffers of sub-sample `subSample` (the reference overwrites them every sub-sample: the last one stays)
ORC_API int oracle_render_guides(void* p, uint32_t subSample, uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1, float* depth, uint32_t* throughput, int threads)
{
    OracleCtx* c = (OracleCtx*)p;
    if (!c->haveConsts || !c->haveView) return -1;
    const uint32_t W = c->consts.imageWidth;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
    #pragma omp parallel
    {
        PathTracerCtx x; x.scene = &c->scene; x.bvh = &c->bvh; x.lights = &c->lights; x.c = &c->consts; x.stats = nullptr; x.neeat = c->neeatPtr();
        x.sampleIndex = c->consts.sampleBaseIndex + subSample; x.worldToClip = c->worldToClip;
        #pragma omp for schedule(dynamic, 1)
        for (int y = int(y0); y < int(y1); y++)
            for (uint32_t px = x0; px < x1; px++)
            {
                GuideOut g; x.guide = &g;
                tracePixel(x, px, uint32_t(y));
                depth[size_t(y) * W + px] = g.depth; throughput[size_t(y) * W + px] = g.throughput;
            }
    }
    return 0;
}
ORC_API int oracle_get_lights(void* p, void* outLightInfos, uint32_t* ioLightCount, uint32_t* outProxyCounters, uint32_t* outProxyIndices, uint32_t* ioProxyCount)
{
    OracleCtx* c = (OracleCtx*)p;
    uint32_t n = uint32_t(c->lights.lights.size()), m = c->lights.samplingProxyCount;
    if (outLightInfos && *ioLightCount >= n) memcpy(outLightInfos, c->lights.lights.data(), size_t(n) * 32);
    if (outProxyCounters && *ioLightCount >= n) memcpy(outProxyCounters, c->lights.proxyCounters.data(), size_t(n) * 4);
    if (outProxyIndices && *ioProxyCount >= m) memcpy(outProxyIndices, c->lights.proxyIndices.data(), size_t(m) * 4);
    *ioLightCount = n; *ioProxyCount = m;
    return 0;
}
ORC_API int oracle_get_lights_ex(void* p, void* outEx, uint32_t* ioCount)
{
    OracleCtx* c = (OracleCtx*)p;
    uint32_t n = c->lights.analyticLightCount;
    if (outEx && *ioCount >= n) memcpy(outEx, c->lights.lightsEx.data(), size_t(n) * 16);
    *ioCount = n;
    return 0;
}
ORC_API int oracle_get_sub_instances(void* p, RtxptSubInstanceData* out, uint32_t count)
{
    OracleCtx* c = (OracleCtx*)p;
    if (count < c->scene.subInstances.size()) return -1;
    memcpy(out, c->scene.subInstances.data(), c->scene.subInstances.size() * sizeof(RtxptSubInstanceData));
    return 0;
}

struct OracleRenderStats { uint64_t scatterRays, shadowRays, nodeVisits, triTests; double seconds; int threads; };

// Renders sub-samples [first, first+count) of the pixel rectangle [x0,x1)x[y0,y1) and folds each into `accum` (RGBA32F, full image
// pitch) with the reference accumulation (AccumulationPass.hlsl:57-65: blend = 1/(n+1), lerp; Sample.cpp:2775).
// `lastOutput` (optional, RGB32F full image pitch) receives the per-sample u_OutputColor of the last sub-sample;
// `primary` (optional, 4 floats per pixel: t,u,v,triId-as-bits) the primary hit of the last sub-sample.
ORC_API int oracle_render(void* p, uint32_t firstSubSample, uint32_t subSampleCount, uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1,
                          float* accum, uint32_t* ioAccumCount, float* lastOutput, float* primary, int threads, OracleRenderStats* outStats)
{
    OracleCtx* c = (OracleCtx*)p;
    if (!c->haveConsts) return -1;
    const uint32_t W = c->consts.imageWidth;
    auto t0 = std::chrono::steady_clock::now();
    RenderStats total;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
    int usedThreads = (threads > 0) ? threads : omp_get_max_threads();
#else
    int usedThreads = 1;
#endif
    for (uint32_t s = 0; s < subSampleCount; s++)
    {
        const uint32_t n = *ioAccumCount;
        const float blend = 1.0f / (float(n) + 1.0f);
        #pragma omp parallel
        {
            RenderStats local;
            PathTracerCtx x; x.scene = &c->scene; x.bvh = &c->bvh; x.lights = &c->lights; x.c = &c->consts; x.stats = &local; x.neeat = c->neeatPtr();
            x.sampleIndex = c->consts.sampleBaseIndex + firstSubSample + s;
            #pragma omp for schedule(dynamic, 1)
            for (int y = int(y0); y < int(y1); y++)
                for (uint32_t px = x0; px < x1; px++)
                {
                    PixelResult r = tracePixel(x, px, uint32_t(y));
                    size_t pix = size_t(y) * W + px;
                    float sample[4] = { r.rgb[0], r.rgb[1], r.rgb[2], 1.0f };
                    for (int k = 0; k < 4; k++)
                    {
                        float prev = accum[pix * 4 + k];
                        accum[pix * 4 + k] = (blend < 1.0f) ? (prev + (sample[k] - prev) * blend) : sample[k];
                    }
                    if (lastOutput) { lastOutput[pix * 3 + 0] = r.rgb[0]; lastOutput[pix * 3 + 1] = r.rgb[1]; lastOutput[pix * 3 + 2] = r.rgb[2]; }
                    if (primary) { primary[pix * 4 + 0] = r.primaryT; primary[pix * 4 + 1] = r.primaryU; primary[pix * 4 + 2] = r.primaryV; primary[pix * 4 + 3] = asfloat(r.primaryTri); }
                }
            #pragma omp critical
            { total.scatterRays += local.scatterRays; total.shadowRays += local.shadowRays; total.nodeVisits += local.nodeVisits; total.triTests += local.triTests; }
        }
        *ioAccumCount = n + 1;
    }
    if (outStats)
    {
        outStats->scatterRays = total.scatterRays; outStats->shadowRays = total.shadowRays; outStats->nodeVisits = total.nodeVisits; outStats->triTests = total.triTests;
        outStats->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); outStats->threads = usedThreads;
    }
    return 0;
}

// Realtime mode over the pixel rectangle: BUILD pass, then rt->subSampleCount FILL passes (sample index = sampleBaseIndex + sub-sample), then
// the no-denoiser merge (stable radiance + every valid plane's noisy radiance) into `merged` (RGB32F, full image pitch, optional).
// All targets are caller-allocated at full image size: planes[3 * planeStride] (GenericTS addressing), header[4][H][W], stableRadiance RGBA16F,
// depth R32F, motionVectors RGBA16F, throughput R32_UINT, specularHitT R32F.
ORC_API int oracle_render_realtime(void* p, const RtxptRealtimeConstants* rt, uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1,
                                   RtxptStablePlane* planes, uint32_t* header, uint16_t* stableRadiance, float* depth, uint16_t* motionVectors, uint32_t* throughput, float* specularHitT,
                                   float* merged, int threads)
{
    OracleCtx* c = (OracleCtx*)p;
    if (!c->haveConsts || !c->haveView || !rt || rt->subSampleCount == 0 || rt->activeStablePlaneCount == 0 || rt->activeStablePlaneCount > 3) return -1;
    RealtimeTargets T;
    T.width = c->consts.imageWidth; T.height = c->consts.imageHeight;
    T.lineStride = GenericTSComputeLineStride(T.width, T.height); T.planeStride = GenericTSComputePlaneStride(T.width, T.height);
    T.planes = planes; T.header = header; T.stableRadiance = stableRadiance; T.depth = depth; T.motionVectors = motionVectors; T.throughput = throughput; T.specularHitT = specularHitT; T.rt = rt;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
    for (uint32_t pass = 0; pass <= rt->subSampleCount; pass++)
    {
        #pragma omp parallel
        {
            PathTracerCtx x; x.scene = &c->scene; x.bvh = &c->bvh; x.lights = &c->lights; x.c = &c->consts; x.stats = nullptr; x.worldToClip = c->worldToClip; x.sp = &T; x.neeat = c->neeatPtr();
            x.mode = pass == 0 ? MODE_BUILD_STABLE_PLANES : MODE_FILL_STABLE_PLANES;
            x.sampleIndex = c->consts.sampleBaseIndex + (pass == 0 ? 0u : pass - 1u);
            #pragma omp for schedule(dynamic, 1)
            for (int y = int(y0); y < int(y1); y++)
                for (uint32_t px = x0; px < x1; px++)
                    if (pass == 0) buildStablePlanesPixel(x, px, uint32_t(y)); else fillStablePlanesPixel(x, px, uint32_t(y));
        }
        // LightsBaker::UpdateEnd sits between the BUILD pass and the radiance passes (Sample.cpp:2495): it needs this frame's depth and motion vectors
        if (pass == 0 && c->neeatPtr() && x0 == 0 && y0 == 0 && x1 == T.width && y1 == T.height) NeeatUpdateEnd(c->neeat, c->lights, depth, motionVectors);
    }
    if (merged)
        for (uint32_t y = y0; y < y1; y++) for (uint32_t px = x0; px < x1; px++)
        {   // u_OutputColor is RGBA16F
            const float3 r = T.GetAllRadiance(px, y); float* o = merged + (size_t(y) * T.width + px) * 3;
            o[0] = lp(r.x); o[1] = lp(r.y); o[2] = lp(r.z);
        }
    return 0;
}
ORC_API uint32_t oracle_branch_advance(uint32_t prev, uint32_t lobe) { return StablePlanesAdvanceBranchID(prev, lobe); }
ORC_API uint32_t oracle_branch_vertex_index(uint32_t id) { return StablePlanesVertexIndexFromBranchID(id); }
ORC_API uint32_t oracle_branch_on_stable_path(uint32_t planeId, uint32_t planeVertex, uint32_t vertexId, uint32_t vertexIndex) { return StablePlaneIsOnStablePath(planeId, planeVertex, vertexId, vertexIndex) ? 1u : 0u; }
ORC_API uint32_t oracle_generic_ts_address(uint32_t x, uint32_t y, uint32_t plane, uint32_t lineStride, uint32_t planeStride) { return GenericTSPixelToAddress(x, y, plane, lineStride, planeStride); }
ORC_API uint32_t oracle_generic_ts_line_stride(uint32_t w, uint32_t h) { return GenericTSComputeLineStride(w, h); }
ORC_API uint32_t oracle_generic_ts_plane_stride(uint32_t w, uint32_t h) { return GenericTSComputePlaneStride(w, h); }
ORC_API void oracle_pack_ortho(const float* m9, uint32_t* out2) { mat3 m; for (int i = 0; i < 3; i++) m.r[i] = f3(m9[3 * i], m9[3 * i + 1], m9[3 * i + 2]); PackOrthoMatrix(m, out2); }
ORC_API void oracle_unpack_ortho(const uint32_t* in2, float* m9) { mat3 m = UnpackOrthoMatrix(in2); for (int i = 0; i < 3; i++) { m9[3 * i] = m.r[i].x; m9[3 * i + 1] = m.r[i].y; m9[3 * i + 2] = m.r[i].z; } }

// RTXPT's side of the denoiser interface for one stable plane.  realtimeTargets = { planes, header, stableRadiance, depth, motionVectors, throughput, specularHitT } as filled
// by oracle_render_realtime; denoiserTargets = { viewZ f32, motion RGBA16F, normalRoughness R10G10B10A2, diffRadianceHitDist RGBA16F, specRadianceHitDist RGBA16F,
// disocclusionMix R8, historyClampRelax R8, outputColor RGBA16F }, all caller-allocated at full image size.
static RealtimeTargets makeTargets(OracleCtx* c, const RtxptRealtimeConstants* rt, void* const* r)
{
    RealtimeTargets T;
    T.width = c->consts.imageWidth; T.height = c->consts.imageHeight;
    T.lineStride = GenericTSComputeLineStride(T.width, T.height); T.planeStride = GenericTSComputePlaneStride(T.width, T.height);
    T.planes = (RtxptStablePlane*)r[0]; T.header = (uint32_t*)r[1]; T.stableRadiance = (uint16_t*)r[2]; T.depth = (float*)r[3]; T.motionVectors = (uint16_t*)r[4]; T.throughput = (uint32_t*)r[5];
    T.specularHitT = (float*)r[6]; T.rt = rt;
    return T;
}
static DenoiserTargets makeDenoiserTargets(void* const* d)
{
    DenoiserTargets D; D.viewZ = (float*)d[0]; D.motion = (uint16_t*)d[1]; D.normalRoughness = (uint32_t*)d[2]; D.diffRadianceHitDist = (uint16_t*)d[3]; D.specRadianceHitDist = (uint16_t*)d[4];
    D.disocclusionMix = (uint8_t*)d[5]; D.historyClampRelax = (uint8_t*)d[6]; D.outputColor = (uint16_t*)d[7];
    return D;
}
ORC_API int oracle_denoiser_prepare_inputs(void* p, const RtxptRealtimeConstants* rt, const RtxptDenoiserConstants* k, uint32_t stablePlaneIndex, int initWithStableRadiance,
                                           void* const* realtimeTargets, void* const* denoiserTargets)
{
    OracleCtx* c = (OracleCtx*)p;
    if (!c->haveConsts || !rt || !k || stablePlaneIndex >= 3) return -1;
    const RealtimeTargets T = makeTargets(c, rt, realtimeTargets); const DenoiserTargets D = makeDenoiserTargets(denoiserTargets);
    PathTracerCtx x; x.scene = &c->scene; x.bvh = &c->bvh; x.lights = &c->lights; x.c = &c->consts; x.stats = nullptr; x.sampleIndex = c->consts.sampleBaseIndex; x.neeat = c->neeatPtr();
    for (uint32_t y = 0; y < T.height; y++) for (uint32_t px = 0; px < T.width; px++)
    {
        float3 co, cd; computeCameraRay(x, px, y, co, cd);
        denoiserPrepareInputsPixel(T, D, *k, px, y, stablePlaneIndex, initWithStableRadiance != 0, co, cd);
    }
    return 0;
}
ORC_API int oracle_denoiser_final_merge(void* p, const RtxptRealtimeConstants* rt, uint32_t stablePlaneIndex, void* const* realtimeTargets, void* const* denoiserTargets,
                                        const uint16_t* denoisedDiff, const uint16_t* denoisedSpec)
{
    OracleCtx* c = (OracleCtx*)p;
    if (!c->haveConsts || !rt || stablePlaneIndex >= 3) return -1;
    const RealtimeTargets T = makeTargets(c, rt, realtimeTargets); const DenoiserTargets D = makeDenoiserTargets(denoiserTargets);
    for (uint32_t y = 0; y < T.height; y++) for (uint32_t px = 0; px < T.width; px++) denoiserFinalMergePixel(T, D, px, y, stablePlaneIndex, denoisedDiff, denoisedSpec);
    return 0;
}

// Donut's skinning pass + the shade-record rewrite (pt_skinning.h); normals / tangents may be NULL; triShade: 24 words per source triangle, updated in place
ORC_API int oracle_skin(uint32_t numVertices, uint32_t numTriangles, uint32_t firstGid, const float* positions, const uint32_t* normals, const uint32_t* tangents, const uint16_t* jointIndices,
                        const float* jointWeights, const float* jointMatrices, const uint32_t* indices, float* outPositions, uint32_t* outNormals, uint32_t* outTangents, uint32_t* triShade)
{
    skinning::skinVertices(numVertices, positions, normals, tangents, jointIndices, jointWeights, jointMatrices, outPositions, outNormals, outTangents);
    if (triShade) skinning::gatherShadeRecords(numTriangles, firstGid, indices, outPositions, normals ? outNormals : nullptr, tangents ? outTangents : nullptr, triShade);
    return 0;
}

// pins against tests/golden/host_golden.json (reference headers compiled in place): MicroRng, candidate counts, white balance
ORC_API void oracle_micro_rng(uint32_t x, uint32_t y, uint32_t a, uint32_t b, uint32_t n, uint32_t* outNext, float* outFloats)
{ MicroRng r = MicroRng::make(x, y, a, b); for (uint32_t i = 0; i < n; i++) outNext[i] = r.Next(); for (uint32_t i = 0; i < n; i++) outFloats[i] = r.NextFloat(); }
ORC_API uint32_t oracle_candidate_local_count(float ratio, uint32_t total) { return ComputeCandidateSampleLocalCount(ratio, total); }
ORC_API void oracle_white_balance(float T, float* outM9, float* outXyz3)
{ const tonemap::M3 m = tonemap::whiteBalanceTransform(T); for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) outM9[i * 3 + j] = m.m[i][j]; const float3 x = tonemap::colorTemperatureToXYZ(T); outXyz3[0] = x.x; outXyz3[1] = x.y; outXyz3[2] = x.z; }

// ToneMappingPass (pt_tonemap.h): RGBA32F frame -> SRGBA8; params as RtxptToneMappingParams; outAux = { average luminance, pre-exposed gray r, g, b }
ORC_API int oracle_tone_map(const RtxptToneMappingParams* u, const float* rgba, uint32_t pixelCount, uint8_t* outRGBA8, float* outAux)
{
    if (!u || !rgba || !outRGBA8) return -1;
    tonemap::Params p; p.op = u->toneMapOperator; p.clamped = u->clamped; p.autoExposure = u->autoExposure; p.enabled = u->enabled; p.whiteBalance = u->whiteBalance;
    p.exposureCompensation = u->exposureCompensation; p.exposureValueMin = u->exposureValueMin; p.exposureValueMax = u->exposureValueMax; p.whiteScale = u->whiteScale; p.whiteMaxLuminance = u->whiteMaxLuminance;
    p.whitePoint = u->whitePoint; p.filmSpeed = u->filmSpeed; p.fNumber = u->fNumber; p.shutter = u->shutter;
    const tonemap::M3 ct = tonemap::colorTransform(p);
    const float avg = tonemap::averageLuminance(rgba, pixelCount);
    for (uint32_t i = 0; i < pixelCount; i++)
    {
        const float3 c = tonemap::apply(p, ct, p.autoExposure ? avg : 1.0f, f3(rgba[4 * i], rgba[4 * i + 1], rgba[4 * i + 2]));
        outRGBA8[4 * i] = tonemap::srgb8(c.x); outRGBA8[4 * i + 1] = tonemap::srgb8(c.y); outRGBA8[4 * i + 2] = tonemap::srgb8(c.z); outRGBA8[4 * i + 3] = uint8_t(saturate(rgba[4 * i + 3]) * 255.0f + 0.5f);
    }
    if (outAux) { outAux[0] = avg; const float3 g = tonemap::preExposedGray(p, avg); outAux[1] = g.x; outAux[2] = g.y; outAux[3] = g.z; }
    return 0;
}

// EnvMapBaker (pt_envbake.h): source + directional lights -> cube MIP chain.  lights: 8 floats each (colour rgb, intensity W/sr, incoming direction xyz, angular size rad);
// out: all MIPs back to back, MIP m = 6 faces of (cubeDim >> m)^2 RGBA32F texels (fp16 values)
ORC_API int oracle_bake_env_map(uint32_t cubeDim, uint32_t sourceType, uint32_t sourceWidth, uint32_t sourceHeight, const float* source, const float* scaleColor, uint32_t lightCount, const float* lights, float* out)
{
    if (cubeDim < 2 || (cubeDim & (cubeDim - 1)) || lightCount > 16) return -1;
    envbake::Desc d; d.cubeDim = cubeDim; d.sourceType = sourceType; d.sourceWidth = sourceWidth; d.sourceHeight = sourceHeight; d.source = source;
    for (int k = 0; k < 3; k++) d.scaleColor[k] = scaleColor[k];
    d.directionalLightCount = lightCount;
    for (uint32_t i = 0; i < lightCount; i++) { memcpy(d.lights[i].colorIntensity, lights + 8 * i, 16); memcpy(d.lights[i].direction, lights + 8 * i + 4, 12); d.lights[i].angularSize = lights[8 * i + 7]; }
    std::vector<std::vector<float>> mips; envbake::bake(d, mips);
    for (auto& m : mips) { memcpy(out, m.data(), m.size() * 4); out += m.size(); }
    return 0;
}

// DenoisingGuidesBaker::DenoiseSpecHitT: in place on the R32F specular hit distance guide, with the R32F depth guide
ORC_API int oracle_denoise_spec_hit_t(uint32_t W, uint32_t H, const float* depth, float* specHitT) { if (!depth || !specHitT) return -1; denoiseSpecHitT(specHitT, depth, int(W), int(H)); return 0; }

// ReBLUR, spatial half (oracle/reblur.h): ClassifyTiles -> [HitDistReconstruction 5x5] -> PrePass -> Blur -> PostBlur on NRD's inputs as rtxpt_b200_denoiser_prepare_inputs /
// oracle_denoiser_prepare_inputs write them.  accumulatedFrames (2 floats per pixel: diffuse, specular; may be NULL = 0) stands in for the history lengths the temporal passes
// would hand to Blur / PostBlur.  stages: bit 0 hit-distance reconstruction, bit 1 pre-pass, bit 2 blur, bit 3 post-blur.  Images are RGBA16F; matrices row-major, row vector x matrix.
ORC_API int oracle_reblur_spatial(uint32_t W, uint32_t H, const float* worldToView16, const float* viewToClip16, uint32_t frameIndex, const float* viewZ, const uint32_t* normalRoughness,
                                  const uint16_t* inDiff, const uint16_t* inSpec, const float* accumulatedFrames, uint32_t stages, uint16_t* outDiff, uint16_t* outSpec, float* outSpecHitDistForTracking,
                                  uint8_t* outTiles)
{
    using namespace orc::reblur;
    Settings settings; const Constants c = makeConstants(settings, W, H, worldToView16, viewToClip16, frameIndex);
    Inputs in; in.W = W; in.H = H; in.viewZ = viewZ; in.normalRoughness = normalRoughness;
    auto load = [&](const uint16_t* src, Image4& img) { img.init(W, H); for (size_t i = 0; i < size_t(W) * H; i++) img.v[i] = f4(f16tof32(src[4 * i]), f16tof32(src[4 * i + 1]), f16tof32(src[4 * i + 2]), f16tof32(src[4 * i + 3])); };
    auto save = [&](const Image4& img, uint16_t* dst) { for (size_t i = 0; i < size_t(W) * H; i++) { dst[4 * i] = uint16_t(f32tof16(img.v[i].x)); dst[4 * i + 1] = uint16_t(f32tof16(img.v[i].y)); dst[4 * i + 2] = uint16_t(f32tof16(img.v[i].z)); dst[4 * i + 3] = uint16_t(f32tof16(img.v[i].w)); } };
    Image4 diffA, specA, diffB, specB; load(inDiff, diffA); load(inSpec, specA); diffB = diffA; specB = specA;
    const std::vector<uint8_t> tiles = classifyTiles(c, in);
    if (outTiles) memcpy(outTiles, tiles.data(), tiles.size());
    std::vector<float2> data1(size_t(W) * H, f2(0, 0));
    if (accumulatedFrames) for (size_t i = 0; i < size_t(W) * H; i++) data1[i] = f2(accumulatedFrames[2 * i], accumulatedFrames[2 * i + 1]);
    std::vector<float> tracking(size_t(W) * H, 0.0f);
    if (stages & 1u) { hitDistReconstruction(c, in, tiles, diffA, specA, diffB, specB); diffA = diffB; specA = specB; }
    if (stages & 2u) { SpatialOutputs o{ &diffB, &specB, &tracking }; spatialPass(c, in, tiles, PRE_BLUR, diffA, specA, nullptr, o); diffA = diffB; specA = specB; }
    if (stages & 4u) { SpatialOutputs o{ &diffB, &specB, nullptr }; spatialPass(c, in, tiles, BLUR, diffA, specA, &data1, o); diffA = diffB; specA = specB; }
    if (stages & 8u) { SpatialOutputs o{ &diffB, &specB, nullptr }; spatialPass(c, in, tiles, POST_BLUR, diffA, specA, &data1, o); diffA = diffB; specA = specB; }
    save(diffA, outDiff); save(specA, outSpec);
    if (outSpecHitDistForTracking) memcpy(outSpecHitDistForTracking, tracking.data(), tracking.size() * 4);
    return 0;
}

// Full ReBLUR frame (oracle/reblur.h: denoiseFrame) with a persistent history per instance - RTXPT keeps one NRD instance per stable plane (Sample.h:327).
struct ReblurInstance { orc::reblur::History history; orc::reblur::FrameOutputs last; };
ORC_API void* oracle_reblur_create() { return new ReblurInstance(); }
ORC_API void oracle_reblur_destroy(void* p) { delete (ReblurInstance*)p; }
// debugging aid: keep / fetch the images after pass `stage` of the last frame (0 HitDistReconstruction, 1 PrePass, 2 TemporalAccumulation, 3 HistoryFix, 4 Blur, 5 PostBlur)
ORC_API void oracle_reblur_keep_stages(void* p, int on) { ((ReblurInstance*)p)->last.keepStages = on != 0; }
ORC_API int oracle_reblur_stage(void* p, uint32_t stage, uint16_t* outDiff, uint16_t* outSpec)
{
    ReblurInstance* inst = (ReblurInstance*)p;
    if (stage >= inst->last.stageDiff.size()) return -1;
    const orc::reblur::Image4& d = inst->last.stageDiff[stage]; const orc::reblur::Image4& s = inst->last.stageSpec[stage];
    for (size_t i = 0; i < d.v.size(); i++) for (int k = 0; k < 4; k++) { outDiff[4 * i + k] = uint16_t(f32tof16((&d.v[i].x)[k])); outSpec[4 * i + k] = uint16_t(f32tof16((&s.v[i].x)[k])); }
    return 0;
}
// matrices: row-major, row vector x matrix; motion: IN_MV RGBA16F (pixels, view-depth delta) or NULL; disocclusionMix: R8 or NULL; outputs RGBA16F; outAccumFrames: 2 floats per pixel (optional)
ORC_API int oracle_reblur_denoise(void* p, uint32_t W, uint32_t H, const float* worldToView16, const float* viewToClip16, const float* worldToViewPrev16, const float* viewToClipPrev16, uint32_t frameIndex,
                                  int resetHistory, const float* viewZ, const uint32_t* normalRoughness, const uint16_t* motion, const uint8_t* disocclusionMix, const uint16_t* inDiff, const uint16_t* inSpec,
                                  uint16_t* outDiff, uint16_t* outSpec, float* outAccumFrames)
{
    using namespace orc::reblur;
    ReblurInstance* inst = (ReblurInstance*)p;
    Inputs in; in.W = W; in.H = H; in.viewZ = viewZ; in.normalRoughness = normalRoughness;
    TemporalParams tp; tp.motion = motion; tp.disocclusionThresholdMix = disocclusionMix; tp.resetHistory = resetHistory != 0;
    tp.disocclusionThreshold = 0.03f + 1.0f / float(H); tp.disocclusionThresholdAlternate = 0.2f + 1.0f / float(H);      // + ( 1 + jitterDelta ) / rectH, no jitter
    const size_t n = size_t(W) * H;
    Image4 d, s; d.init(W, H); s.init(W, H);
    for (size_t i = 0; i < n; i++) { d.v[i] = f4(f16tof32(inDiff[4 * i]), f16tof32(inDiff[4 * i + 1]), f16tof32(inDiff[4 * i + 2]), f16tof32(inDiff[4 * i + 3])); s.v[i] = f4(f16tof32(inSpec[4 * i]), f16tof32(inSpec[4 * i + 1]), f16tof32(inSpec[4 * i + 2]), f16tof32(inSpec[4 * i + 3])); }
    denoiseFrame(Settings(), W, H, worldToView16, viewToClip16, worldToViewPrev16, viewToClipPrev16, frameIndex, tp, in, d, s, inst->history, inst->last);
    for (size_t i = 0; i < n; i++)
    {
        const float4 a = inst->last.diff.v[i], b = inst->last.spec.v[i];
        outDiff[4 * i] = uint16_t(f32tof16(a.x)); outDiff[4 * i + 1] = uint16_t(f32tof16(a.y)); outDiff[4 * i + 2] = uint16_t(f32tof16(a.z)); outDiff[4 * i + 3] = uint16_t(f32tof16(a.w));
        outSpec[4 * i] = uint16_t(f32tof16(b.x)); outSpec[4 * i + 1] = uint16_t(f32tof16(b.y)); outSpec[4 * i + 2] = uint16_t(f32tof16(b.z)); outSpec[4 * i + 3] = uint16_t(f32tof16(b.w));
        if (outAccumFrames) { outAccumFrames[2 * i] = float(inst->last.data1[2 * i]) / 255.0f * 63.0f; outAccumFrames[2 * i + 1] = float(inst->last.data1[2 * i + 1]) / 255.0f * 63.0f; }
    }
    return 0;
}

// primary-hit triangle id -> (instance, geometry, primitive), for comparing with the product's hit records
ORC_API int oracle_tri_info(void* p, uint32_t triId, uint32_t* out3)
{
    OracleCtx* c = (OracleCtx*)p;
    if (triId >= c->bvh.tris.size()) return -1;
    const Tri& t = c->bvh.tris[triId];
    out3[0] = t.instanceIndex; out3[1] = t.geometryIndex; out3[2] = t.primitiveIndex;
    return 0;
}

} // extern "C"
