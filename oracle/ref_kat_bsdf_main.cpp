// ORACLE/_ref — TEST INFRASTRUCTURE: known-answer generator for the floating-point material model, built from the UNMODIFIED reference headers
//   Rtxpt/Shaders/PathTracer/Rendering/Materials/{Fresnel,Microfacet,BxDF,StandardBSDF,IBSDF}.hlsli, Utils/Math/MathHelpers.hlsli, Scene/{ShadingData,Material/MaterialData}.hlsli
// compiled in place through oracle/ref_hlsl_shim.h (oracle/ref_hlsl_tu.sh appends this file to the stream it feeds g++; it is not compiled on its own).
//   ref_kat_bsdf bsdf  in.f32 out.f32     in: N x 36 floats (tests/bsdf_records.py layout, the one oracle_bsdf and rtxpt_b200_debug_bsdf take)
//                                         out: N x 40 floats: [0..15] eval.xyzw, evalPdf, sample{valid, wo, pdf, weight, lobe, lobeP}, getLobes (= oracle_bsdf's layout),
//                                              [16..31] evalDeltaLobes: 2 x {thp.xyz, probability, dir.xyz, transmission}, [32] nonDeltaPart, [33] deltaLobeCount,
//                                              [34..39] estimateSpecDiffBSDF( N, V ): diffuse estimate, specular estimate
//   ref_kat_bsdf funcs in.f32 out.f32     in: N x 8 uniform floats; out: N x 40 floats, the scalar building blocks (see below)
// tests/golden/make_bsdf_golden.py runs it on seeded inputs and commits the vectors; tests/test_oracle_golden.py holds the oracle's restatement to them.
#include <cstdio>
#include <vector>

static std::vector<float> readAll(const char* path)
{
    FILE* f = fopen(path, "rb"); if (!f) { perror(path); exit(2); }
    fseek(f, 0, SEEK_END); const long n = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<float> v(size_t(n) / 4); if (fread(v.data(), 4, v.size(), f) != v.size()) { perror("read"); exit(2); }
    fclose(f); return v;
}
static void writeAll(const char* path, const std::vector<float>& v)
{
    FILE* f = fopen(path, "wb"); if (!f) { perror(path); exit(2); }
    fwrite(v.data(), 4, v.size(), f); fclose(f);
}

// EnvironmentQuadLight::ToWorld / ToLocal are PathTracerNEE.hlsli's, over Bridge::CreateEnvMap(): the stub bridge's environment transform starts as the identity
float4 (*g_shimCubeSample)(float3 dir, float lod) = nullptr;
static void shimIdentityEnv() { memset(&g_bridge.env, 0, sizeof(g_bridge.env)); for (int k = 0; k < 3; k++) g_bridge.env.Transform.m[k][k] = g_bridge.env.InvTransform.m[k][k] = 1.0f; g_bridge.env.ColorMultiplier = float3(1, 1, 1); }


// ---- "hit" mode: PathTracer::HandleHit / HandleMiss behind the stub bridge (ref_bridge_stub.h) ------------------------------------------------------------------------------------------
// record layout (floats; words marked * are bit patterns):  0-19* path payload | 20-22 ray origin, 23-25 ray dir, 26 rayTCurrent, 27 miss? | 28-58 surface (posW, faceNCorrected, N, T, B,
// vertexN, frontFacing, nestedPriority, activeLobes, thinSurface, psdExclude, materialID, IoR, shadowNoLFadeout, emission, psdBlockMotionVectors, psdDominantDeltaLobeP1) | 60-73 BSDF data
// (order of the "bsdf" mode's words 18-31) | 74 interior IoR, 75 / 76 emissive-triangle / analytic-proxy light (-1: none), 77-79 prevPosW | 80-93 constants | 96-135 materials (IoR,
// attenuation colour, attenuation distance x 8) | 136-151 proxy counters, 152-215 proxy indices, 216-727* 2 x 2 local tiles, 728-919* 16 light records (Base + Extended) | FILL mode: 920-923* the pixel's stable-plane header (3 branch ids, first-hit length | dominant index), 924-929* the three planes'
// packed noisy radiance, 930 the pixel's specular hit distance | BUILD mode: 931-942 the stub camera (position, direction at pixel 0, per-pixel steps), 943 maxStablePlaneVertexDepth,
// 944 allowPrimarySurfaceReplacement, 945 stablePlanesSplitStopThreshold, 946-949 the pixel's stable radiance (RGBA16F values) | 950-958 the environment map's rotation (rows), 959 its intensity | 960-1019* the pixel's three stable planes (ops 3, 4), 1020 the pixel cone spread angle (op 2)
// word 27 selects the call: 0 HandleHit, 1 HandleMiss, 2 EmptyPathInitialize, 3 FirstHitFromVBuffer (FILL binary), 4 postProcessHit (BUILD binary)
static const int kHitIn = 1024, kHitOut = 128;
struct ShimHitScenario
{
    LightingControlData cd; uint counters[16], indices[64], local[512]; PolymorphicLightInfo lights[16]; PolymorphicLightInfoEx lightsEx[16]; float fbWeight[64]; uint fbCand[64]; uint envLookup[1];
    PathTracerConstants pt;
    uint spHeader[4 * 64]; StablePlane spPlanes[3 * 256]; float4 spRadiance[64];
};
static float4 shimCube(float3 d, float lod) { const float k = exp2(-lod); return float4((0.5f + 0.5f * d.x) * k, (0.5f + 0.25f * d.y) * k, (0.75f + 0.25f * d.z) * k, 1.0f); }
static void shimLoadHitScenario(const float* r, ShimHitScenario& S)
{
    g_bridge = ShimBridgeScenario(); shimIdentityEnv(); g_shimCubeSample = shimCube;
    memset(&S.cd, 0, sizeof(S.cd)); memset(&S.pt, 0, sizeof(S.pt)); memset(S.lights, 0, sizeof(S.lights)); memset(S.lightsEx, 0, sizeof(S.lightsEx));
    // surface
    ShadingData sd = ShadingData::make();
    sd.posW = float3(r[28], r[29], r[30]); sd.faceNCorrected = float3(r[31], r[32], r[33]); sd.V = -float3(r[23], r[24], r[25]); sd.N = float3(r[34], r[35], r[36]); sd.T = float3(r[37], r[38], r[39]);
    sd.B = float3(r[40], r[41], r[42]); sd.vertexN = float3(r[43], r[44], r[45]); sd.frontFacing = r[46] != 0.0f;
    sd.mtl = MaterialHeader::make(); sd.mtl.setNestedPriority(uint(r[47])); sd.mtl.setActiveLobes(uint(r[48])); sd.mtl.setThinSurface(r[49] != 0.0f); sd.mtl.setPSDExclude(r[50] != 0.0f);
    sd.mtl.setPSDBlockMotionVectorsAtSurface(r[57] != 0.0f); sd.mtl.setPSDDominantDeltaLobeP1(uint(r[58]));
    sd.materialID = uint(r[51]); sd.IoR = lpfloat(r[52]); sd.shadowNoLFadeout = lpfloat(r[53]); sd.emission = lpfloat3(lpfloat(r[54]), lpfloat(r[55]), lpfloat(r[56]));
    const float* b = r + 42;        // the "bsdf" mode's word k sits at k + 42
    StandardBSDF bsdf = StandardBSDF::make(StandardBSDFData::make(lpfloat3(lpfloat(b[18]), lpfloat(b[19]), lpfloat(b[20])), lpfloat3(lpfloat(b[22]), lpfloat(b[23]), lpfloat(b[24])), lpfloat(b[21]), lpfloat(b[25]),
                                                                     lpfloat(b[31]), lpfloat3(lpfloat(b[26]), lpfloat(b[27]), lpfloat(b[28])), lpfloat(b[29]), lpfloat(b[30])));
    const uint neeTri = r[75] < 0 ? 0xFFFFFFFFu : uint(r[75]), neeAna = r[76] < 0 ? 0xFFFFFFFFu : uint(r[76]);
#if PATH_TRACER_MODE == PATH_TRACER_MODE_BUILD_STABLE_PLANES
    g_bridge.surface = PathTracer::SurfaceData::make(sd, bsdf, float3(r[77], r[78], r[79]), lpfloat(r[74]), neeTri, neeAna);
#else
    g_bridge.surface = PathTracer::SurfaceData::make(sd, bsdf, lpfloat(r[74]), neeTri, neeAna);
#endif
    // constants
    g_bridge.maxBounces = uint(r[80]); g_bridge.maxDiffuseBounces = uint(r[81]); g_bridge.sampleIndex = uint(r[82]); g_bridge.noisyRadianceAttenuation = r[87]; g_bridge.envMipOffset = r[93];
    S.pt.bounceCount = uint(r[80]); S.pt.diffuseBounceCount = uint(r[81]); S.pt.NEEEnabled = 1; S.pt.NEECandidateSamples = uint(r[83]); S.pt.NEEFullSamples = uint(r[84]); S.pt.fireflyFilterThreshold = r[85];
    S.pt.invSubSampleCount = r[87]; S.pt.EnvironmentMapDiffuseSampleMIPLevel = r[93];
    // materials: the bridge's absorption rule (PathTracerBridgeDonut.hlsli:881-887), restated
    g_bridge.materialCount = 8;
    for (int m = 0; m < 8; m++)
    {
        g_bridge.ior[m] = r[96 + m]; const float dist = max(1e-30f, r[128 + m]);
        g_bridge.sigmaA[m] = float3(-log(clamp(r[104 + 3 * m], 1e-7f, 1.0f)) / dist, -log(clamp(r[105 + 3 * m], 1e-7f, 1.0f)) / dist, -log(clamp(r[106 + 3 * m], 1e-7f, 1.0f)) / dist);
    }
    // lights
    S.cd.TotalLightCount = 16; S.cd.SamplingProxyCount = uint(r[90]); S.cd.LocalSamplingTileJitter = uint2((uint)r[88], (uint)r[89]); S.cd.LocalToGlobalSampleRatio = r[86];
    S.cd.LocalSamplingResolution = uint2(2u, 2u); S.cd.TemporalFeedbackRequired = uint(r[92]); S.cd.ScreenSpaceVsWorldSpaceThreshold = r[91];
    for (int k = 0; k < 16; k++) S.counters[k] = uint(r[136 + k]);
    for (int k = 0; k < 64; k++) S.indices[k] = uint(r[152 + k]);
    memcpy(S.local, r + 216, sizeof(S.local));
    for (int k = 0; k < 16; k++)
    {
        uint w[12]; memcpy(w, r + 728 + 12 * k, 48);
        S.lights[k].Center = float3(asfloat(w[0]), asfloat(w[1]), asfloat(w[2])); S.lights[k].ColorTypeAndFlags = w[3]; S.lights[k].Direction1 = w[4]; S.lights[k].Direction2 = w[5]; S.lights[k].Scalars = w[6];
        S.lights[k].LogRadiance = w[7]; S.lightsEx[k].IesProfileIndex = w[8]; S.lightsEx[k].PrimaryAxis = w[9]; S.lightsEx[k].CosConeAngleAndSoftness = w[10]; S.lightsEx[k].UniqueID = w[11];
    }
    for (int k = 0; k < 64; k++) { S.fbWeight[k] = 0.0f; S.fbCand[k] = 0xFFFFFFFFu; }
    // the environment lookup map (direction -> environment-quad light): 1024 x 1024 texels of the equal-area octahedral map, in 64-texel blocks over lights 0-3 (what the record's first
    // four lights are when it has environment quads)
    static std::vector<uint> envLookup1024;
    if (envLookup1024.empty()) { envLookup1024.resize(1024 * 1024); for (uint y = 0; y < 1024; y++) for (uint x = 0; x < 1024; x++) envLookup1024[y * 1024 + x] = ((x >> 6) + (y >> 6) * 3u) & 3u; }
    memset(&g_bridge.env, 0, sizeof(g_bridge.env));
    for (int a = 0; a < 3; a++) for (int c = 0; c < 3; c++) { g_bridge.env.Transform.m[a][c] = r[950 + 3 * a + c]; g_bridge.env.InvTransform.m[c][a] = r[950 + 3 * a + c]; }
    g_bridge.env.ColorMultiplier = float3(r[959], r[959], r[959]); g_bridge.env.Enabled = 1.0f;
    g_bridge.control.p = &S.cd; g_bridge.control.n = 1; g_bridge.lights.p = S.lights; g_bridge.lights.n = 16; g_bridge.lightsEx.p = S.lightsEx; g_bridge.lightsEx.n = 16;
    g_bridge.proxyCounters.p = S.counters; g_bridge.proxyCounters.n = 16; g_bridge.proxyIndices.p = S.indices; g_bridge.proxyIndices.n = S.cd.SamplingProxyCount; g_bridge.localSampling.p = S.local; g_bridge.localSampling.n = 512;
    g_bridge.envLookup.p = envLookup1024.data(); g_bridge.envLookup.w = g_bridge.envLookup.h = 1024; g_bridge.feedbackWeight.p = S.fbWeight; g_bridge.feedbackWeight.w = g_bridge.feedbackWeight.h = 8;
    g_bridge.feedbackCandidates.p = S.fbCand; g_bridge.feedbackCandidates.w = g_bridge.feedbackCandidates.h = 8;
    g_bridge.hasEnvMap = true;
}

int main(int argc, char** argv)
{
    if (argc != 4) { fprintf(stderr, "usage: %s bsdf|funcs|utils|helpers|lights|spheres|tonemap|texlod|interior|sampler|hit|envquads in.f32 out.f32\n", argv[0]); return 2; }
    const std::vector<float> in = readAll(argv[2]); std::vector<float> out;
    shimIdentityEnv();
    if (std::string(argv[1]) == "bsdf")
    {
        const size_t n = in.size() / 36; out.assign(n * 40, 0.0f);
        for (size_t i = 0; i < n; i++)
        {
            const float* r = &in[i * 36]; float* o = &out[i * 40];
            ShadingData sd = ShadingData::make();
            sd.V = float3(r[0], r[1], r[2]); sd.N = float3(r[3], r[4], r[5]); sd.T = float3(r[6], r[7], r[8]); sd.B = float3(r[9], r[10], r[11]);
            const float3 wo(r[12], r[13], r[14]);
            sd.mtl = MaterialHeader::make(); sd.mtl.setActiveLobes(uint(r[33])); sd.mtl.setThinSurface(r[32] != 0.0f); sd.mtl.setPSDExclude(false);
            StandardBSDF b = StandardBSDF::make(StandardBSDFData::make(lpfloat3(lpfloat(r[18]), lpfloat(r[19]), lpfloat(r[20])), lpfloat3(lpfloat(r[22]), lpfloat(r[23]), lpfloat(r[24])), lpfloat(r[21]), lpfloat(r[25]),
                                                                         lpfloat(r[31]), lpfloat3(lpfloat(r[26]), lpfloat(r[27]), lpfloat(r[28])), lpfloat(r[29]), lpfloat(r[30])));
            const float4 e = b.eval(sd, wo);
            o[0] = e.x; o[1] = e.y; o[2] = e.z; o[3] = e.w;
            o[4] = b.evalPdf(sd, wo, true);
            BSDFSample s; s.wo = float3(0, 0, 0); s.pdf = 0; s.weight = float3(0, 0, 0); s.lobe = 0; s.lobeP = 0;
            const bool valid = b.sample(sd, float4(r[15], r[16], r[17], 0.0f), s, true);
            o[5] = valid ? 1.0f : 0.0f; o[6] = s.wo.x; o[7] = s.wo.y; o[8] = s.wo.z; o[9] = s.pdf; o[10] = s.weight.x; o[11] = s.weight.y; o[12] = s.weight.z;
            o[13] = float(s.lobe); o[14] = s.lobeP; o[15] = float(b.getLobes(sd));
            DeltaLobe lobes[cMaxDeltaLobes]; int count = 0; float nonDelta = 0;
            b.evalDeltaLobes(sd, lobes, count, nonDelta);
            for (int k = 0; k < 2; k++) { float* d = o + 16 + k * 8; d[0] = lobes[k].thp.x; d[1] = lobes[k].thp.y; d[2] = lobes[k].thp.z; d[3] = lobes[k].probability; d[4] = lobes[k].dir.x; d[5] = lobes[k].dir.y; d[6] = lobes[k].dir.z; d[7] = float(lobes[k].transmission); }
            o[32] = nonDelta; o[33] = float(count);
            float3 de, se; b.estimateSpecDiffBSDF(de, se, sd.N, sd.V);
            o[34] = de.x; o[35] = de.y; o[36] = de.z; o[37] = se.x; o[38] = se.y; o[39] = se.z;
        }
    }
    else if (std::string(argv[1]) == "interior")
    {   // Rendering/Materials/InteriorList.hlsli: the two-slot stack of nested dielectrics.  A record is a sequence of 12 surface crossings (material, nested priority, entering,
        // probe priority); after each the packed slots and every query are written out (6 values per crossing; slots as bit patterns)
        const size_t n = in.size() / 48; out.assign(n * 72, 0.0f);
        for (size_t i = 0; i < n; i++)
        {
            const float* u = &in[i * 48]; float* o = &out[i * 72];
            InteriorList il; il.slots = uint2(0u, 0u);
            for (int k = 0; k < 12; k++)
            {
                const uint mat = uint(u[4 * k]), prio = uint(u[4 * k + 1]); const bool entering = u[4 * k + 2] != 0.0f; const uint probe = uint(u[4 * k + 3]);
                const bool isTrue = il.isTrueIntersection(probe);                       // asked BEFORE the crossing is committed, as HandleHit does
                il.handleIntersection(mat, prio, entering);
                const uint w[2] = { il.slots.x, il.slots.y }; memcpy(o + 6 * k, w, 8);
                o[6 * k + 2] = float(il.getTopNestedPriority()); o[6 * k + 3] = float(int(il.getTopMaterialID())); o[6 * k + 4] = float(int(il.getNextMaterialID())); o[6 * k + 5] = isTrue ? 1.0f : 0.0f;
            }
        }
    }
    else if (std::string(argv[1]) == "sampler")
    {   // Lighting/LightSampler.hlsli (+ LightingTypes.hlsli's LightFeedbackReservoir, LightingAlgorithms.hlsli's LocalLightBinarySearch): the NEE-AT sampler side a path
        // vertex calls.  A record is one small scenario - 16 lights' proxy counters, <= 64 global proxies, 2 x 2 tiles of 128 packed (light, count) tuples, an 8 x 8 cleared
        // feedback image - and 8 queries run in order against it (the feedback inserts accumulate).  680 floats in (packed words as bit patterns), 8 x 16 out
        const size_t n = in.size() / 680; out.assign(n * 128, 0.0f);
        for (size_t i = 0; i < n; i++)
        {
            const float* r = &in[i * 680]; float* o = &out[i * 128];
            LightingControlData cd; memset(&cd, 0, sizeof(cd));
            cd.TotalLightCount = 16; cd.SamplingProxyCount = uint(r[0]); cd.LocalSamplingTileJitter = uint2(uint(r[1]), uint(r[2])); cd.LocalToGlobalSampleRatio = r[5];
            cd.LocalSamplingResolution = uint2(2u, 2u); cd.TemporalFeedbackRequired = 1; cd.ScreenSpaceVsWorldSpaceThreshold = r[6];
            const uint candidateSampleCount = uint(r[3]), fullSamples = uint(r[4]);
            uint counters[16], indices[64], local[512]; PolymorphicLightInfo lights[16]; PolymorphicLightInfoEx lightsEx[16]; memset(lights, 0, sizeof(lights)); memset(lightsEx, 0, sizeof(lightsEx));
            for (int k = 0; k < 16; k++) counters[k] = uint(r[8 + k]);
            for (int k = 0; k < 64; k++) indices[k] = uint(r[24 + k]);
            memcpy(local, r + 88, sizeof(local));
            float fbWeight[64]; uint fbCand[64]; uint envLookup[1] = { 0u };
            for (int k = 0; k < 64; k++) { fbWeight[k] = 0.0f; fbCand[k] = 0xFFFFFFFFu; }
            StructuredBuffer<LightingControlData> bControl; bControl.p = &cd; bControl.n = 1;
            StructuredBuffer<PolymorphicLightInfo> bLights; bLights.p = lights; bLights.n = 16;
            StructuredBuffer<PolymorphicLightInfoEx> bLightsEx; bLightsEx.p = lightsEx; bLightsEx.n = 16;
            Buffer<uint> bCounters; bCounters.p = counters; bCounters.n = 16;
            Buffer<uint> bIndices; bIndices.p = indices; bIndices.n = cd.SamplingProxyCount;
            Buffer<uint> bLocal; bLocal.p = local; bLocal.n = 512;
            Texture2D<uint> tEnv; tEnv.p = envLookup; tEnv.w = tEnv.h = 1;
            RWTexture2D<float> tWeight; tWeight.p = fbWeight; tWeight.w = tWeight.h = 8;
            RWTexture2D<uint> tCand; tCand.p = fbCand; tCand.w = tCand.h = 8;
            for (int q = 0; q < 8; q++)
            {
                const float* u = r + 600 + q * 10; float* d = o + q * 16;
                const uint2 pixel = uint2((uint)u[0], (uint)u[1]); const uint lightIndex = uint(u[3]); const uint flags = uint(u[4]); const bool isSSC = (flags & 1u) != 0;
                LightSampler ls = LightSampler::make(bControl, bLights, bLightsEx, bCounters, bIndices, bLocal, tEnv, tWeight, tCand, pixel, isSSC);
                float pdf = 0; uint idx = ls.SampleGlobal(u[2], pdf); d[0] = float(idx); d[1] = pdf;
                idx = ls.SampleLocal(u[2], pdf); d[2] = float(idx); d[3] = pdf;
                d[4] = ls.SampleGlobalPDF(lightIndex); d[5] = ls.SampleLocalPDF(lightIndex);
                uint localCount = 0, globalCount = 0; ls.GetCandidateSampleCounts(candidateSampleCount, localCount, globalCount); d[6] = float(localCount); d[7] = float(globalCount);
                d[8] = ls.ComputeLightVsBSDF_MIS_ForBSDF(lightIndex, lpfloat(u[7]), u[8], candidateSampleCount, fullSamples);
                LightSample s = LightSample::make(); s.LightIndex = lightIndex; s.SelectionPdf = u[9]; s.SolidAnglePdf = u[8]; s.FromLocalDistribution = (flags & 2u) != 0; s.LightSampleableByBSDF = (flags & 4u) != 0;
                float thisPdf, otherPdf, thisCount, otherCount; ls.ComputeLightSelectionPdfs(s, localCount, globalCount, thisPdf, otherPdf, thisCount, otherCount);
                d[9] = otherPdf; d[10] = thisCount; d[11] = otherCount;
                d[12] = ls.ComputeLightVsBSDF_MIS_ForLight(float3(0, 0, 0), s, thisPdf, otherPdf, thisCount, otherCount, candidateSampleCount, fullSamples, u[7]);
                ls.InsertFeedbackFromNEE(lightIndex, u[5], u[6]);
                d[13] = fbWeight[pixel.y * 8 + pixel.x]; memcpy(d + 14, &fbCand[pixel.y * 8 + pixel.x], 4);
                d[15] = LightSampler::IsScreenSpaceCoherentHeuristic(bControl, u[8], u[5]) ? 1.0f : 0.0f;
            }
        }
    }
    else if (std::string(argv[1]) == "hit")
    {   // PathTracer::HandleHit / HandleMiss (PathTracer.hlsli:391-763) on one path vertex: the incoming path state, the ray that found the surface, the surface itself (what
        // Bridge::loadSurface would return) and a small light scenario in; the outgoing path state and everything the call exported out
        const size_t n = in.size() / kHitIn; out.assign(n * kHitOut, 0.0f);
        ShimHitScenario* S = new ShimHitScenario();
        for (size_t i = 0; i < n; i++)
        {
            const float* r = &in[i * kHitIn]; float* o = &out[i * kHitOut];
            shimLoadHitScenario(r, *S);
            PathPayload payload; memcpy(payload.packed, r, 80);
            PathState path = PathPayload::unpack(payload);
            PathTracer::WorkingContext wc; wc.PtConsts = S->pt;
#if PATH_TRACER_MODE != PATH_TRACER_MODE_REFERENCE
            // the stable planes of an 8 x 8 image; the record carries the pixel's own entries
            const uint2 pixelIn = PathPayload::unpack(payload).GetPixelPos(); const uint pxi = pixelIn.x & 7u, pyi = pixelIn.y & 7u;
            wc.PtConsts.imageWidth = wc.PtConsts.imageHeight = 8; wc.PtConsts.genericTSLineStride = GenericTSComputeLineStride(8, 8); wc.PtConsts.genericTSPlaneStride = GenericTSComputePlaneStride(8, 8);
            memset(S->spHeader, 0xFF, sizeof(S->spHeader)); memset(S->spPlanes, 0, sizeof(S->spPlanes)); memset(S->spRadiance, 0, sizeof(S->spRadiance));
            RWTexture2DArray<uint> hdr; hdr.p = S->spHeader; hdr.w = hdr.h = 8; hdr.d = 4; RWStructuredBuffer<StablePlane> planes; planes.p = S->spPlanes; planes.n = 3 * 256; RWTexture2D<float4> rad; rad.p = S->spRadiance; rad.w = rad.h = 8;
            wc.StablePlanes = StablePlanesContext::make(hdr, planes, rad, wc.PtConsts);
            for (uint k = 0; k < 4; k++) memcpy(&S->spHeader[(k * 8 + pyi) * 8 + pxi], r + 920 + k, 4);
            for (uint k = 0; k < 3; k++) { uint w[2]; memcpy(w, r + 924 + 2 * k, 8); S->spPlanes[wc.StablePlanes.PixelToAddress(uint2(pxi, pyi), k)].PackedNoisyRadianceAndSpecAvg = uint2(w[0], w[1]); }
            g_bridge.specularHitT = r[930];
            if (r[27] >= 3.0f) for (uint k = 0; k < 3; k++)
            {   // ops 3, 4 read whole planes: a stored base plane (FirstHitFromVBuffer) or an enqueued path (postProcessHit: StablePlane::UnpackCustomPayload is member by member too)
                uint w[20]; memcpy(w, r + 960 + 20 * k, 80); StablePlane& sp = S->spPlanes[wc.StablePlanes.PixelToAddress(uint2(pxi, pyi), k)];
                sp.RayOrigin = float3(asfloat(w[0]), asfloat(w[1]), asfloat(w[2])); sp.LastRayTCurrent = asfloat(w[3]); sp.RayDir = float3(asfloat(w[4]), asfloat(w[5]), asfloat(w[6])); sp.SceneLength = asfloat(w[7]);
                sp.PackedThpAndMVs = uint3(w[8], w[9], w[10]); sp.VertexIndexAndRoughness = w[11]; sp.DenoiserPackedBSDFEstimate = uint3(w[12], w[13], w[14]); sp.PackedNormal = w[15];
                sp.PackedNoisyRadianceAndSpecAvg = uint2(w[16], w[17]); sp.FlagsAndVertexIndex = w[18]; sp.PackedCounters = w[19];
            }
            g_bridge.cameraPos = float3(r[931], r[932], r[933]); g_bridge.cameraDirBase = float3(r[934], r[935], r[936]); g_bridge.cameraDirDx = float3(r[937], r[938], r[939]); g_bridge.cameraDirDy = float3(r[940], r[941], r[942]);
            wc.PtConsts.maxStablePlaneVertexDepth = uint(r[943]); wc.PtConsts.allowPrimarySurfaceReplacement = uint(r[944]); wc.PtConsts.stablePlanesSplitStopThreshold = r[945];
            wc.StablePlanes.PTConstants = wc.PtConsts;
            S->spRadiance[pyi * 8 + pxi] = float4(r[946], r[947], r[948], r[949]);
#endif
            const float3 rayOrigin(r[20], r[21], r[22]), rayDir(r[23], r[24], r[25]);
            float2 tMinMax = float2(0.0f, 0.0f);
            if (r[27] == 1.0f) PathTracer::HandleMiss(path, rayOrigin, rayDir, r[26], wc);
            else if (r[27] == 0.0f) PathTracer::HandleHit(path, rayOrigin, rayDir, r[26], float2(0.25f, 0.25f), wc);
            else if (r[27] == 2.0f) path = PathTracer::EmptyPathInitialize(path.GetPixelPos(), r[1020]);
#if PATH_TRACER_MODE == PATH_TRACER_MODE_FILL_STABLE_PLANES
            else if (r[27] == 3.0f) { path = PathTracer::EmptyPathInitialize(path.GetPixelPos(), r[1020]); tMinMax = FirstHitFromVBuffer(path, 0, wc); }
#endif
#if PATH_TRACER_MODE == PATH_TRACER_MODE_BUILD_STABLE_PLANES
            else if (r[27] == 4.0f) postProcessHit(path, wc);
#endif
            o[120] = tMinMax.x; o[121] = tMinMax.y;
            payload = PathPayload::pack(path); memcpy(o, payload.packed, 80);
            o[20] = float(g_bridge.visibilityQueries); o[21] = g_bridge.lastVisibilityRay.Origin.x; o[22] = g_bridge.lastVisibilityRay.Origin.y; o[23] = g_bridge.lastVisibilityRay.Origin.z;
            o[24] = g_bridge.lastVisibilityRay.Direction.x; o[25] = g_bridge.lastVisibilityRay.Direction.y; o[26] = g_bridge.lastVisibilityRay.Direction.z; o[27] = g_bridge.lastVisibilityRay.TMax;
            o[28] = g_bridge.lastVisibility ? 1.0f : 0.0f; o[29] = float(g_bridge.exportSurfaceCalls); o[30] = g_bridge.exportSceneLength; o[31] = float(g_bridge.exportNonSurfaceCalls);
            o[32] = g_bridge.exportVirtualPos.x; o[33] = g_bridge.exportVirtualPos.y; o[34] = g_bridge.exportVirtualPos.z; o[35] = float(g_bridge.specHitTStarts); o[36] = float(g_bridge.specHitTStops);
            o[37] = g_bridge.specularHitT;
#if PATH_TRACER_MODE != PATH_TRACER_MODE_REFERENCE
            for (uint k = 0; k < 3; k++) { const uint2 w = S->spPlanes[wc.StablePlanes.PixelToAddress(uint2(pxi, pyi), k)].PackedNoisyRadianceAndSpecAvg; memcpy(o + 41 + 2 * k, &w.x, 4); memcpy(o + 42 + 2 * k, &w.y, 4); }
            for (uint k = 0; k < 4; k++) memcpy(o + 47 + k, &S->spHeader[(k * 8 + pyi) * 8 + pxi], 4);
            // the stable radiance target is RGBA16F: the format conversion of the UAV store, which the stand-in texture does not have, is applied here
            { const float4 sr = S->spRadiance[pyi * 8 + pxi]; o[52] = f16tof32(f32tof16(sr.x)); o[53] = f16tof32(f32tof16(sr.y)); o[54] = f16tof32(f32tof16(sr.z)); o[55] = f16tof32(f32tof16(sr.w)); }
            for (uint k = 0; k < 3; k++)
            {   // the 80-byte record, member by member (the shim's vector types are wider than their HLSL counterparts)
                const StablePlane& sp = S->spPlanes[wc.StablePlanes.PixelToAddress(uint2(pxi, pyi), k)];
                const uint w[20] = { asuint(sp.RayOrigin.x), asuint(sp.RayOrigin.y), asuint(sp.RayOrigin.z), asuint(sp.LastRayTCurrent), asuint(sp.RayDir.x), asuint(sp.RayDir.y), asuint(sp.RayDir.z), asuint(sp.SceneLength),
                                     sp.PackedThpAndMVs.x, sp.PackedThpAndMVs.y, sp.PackedThpAndMVs.z, sp.VertexIndexAndRoughness, sp.DenoiserPackedBSDFEstimate.x, sp.DenoiserPackedBSDFEstimate.y, sp.DenoiserPackedBSDFEstimate.z,
                                     sp.PackedNormal, sp.PackedNoisyRadianceAndSpecAvg.x, sp.PackedNoisyRadianceAndSpecAvg.y, sp.FlagsAndVertexIndex, sp.PackedCounters };
                memcpy(o + 56 + 20 * k, w, 80);
            }
            o[117] = g_bridge.exportMotion.x; o[118] = g_bridge.exportMotion.y; o[119] = g_bridge.exportMotion.z;
#endif
            const uint2 px = path.GetPixelPos(); const uint at = (px.y & 7u) * 8u + (px.x & 7u);
            o[39] = S->fbWeight[at]; memcpy(o + 40, &S->fbCand[at], 4);
        }
        delete S;
    }
    else if (std::string(argv[1]) == "envquads")
    {   // Lighting/PolymorphicLight.hlsli: an environment-quad light (a node of the quad tree over the equal-area octahedral environment map) through EnvironmentQuadLight::Store, Create,
        // CalcSample (through the environment rotation words 8-16 give; identity when all zero), CalcSolidAnglePdfForMIS, GetPower.  24 floats in, 24 out (words 0-11: the record)
        const size_t n = in.size() / 24; out.assign(n * 24, 0.0f);
        for (size_t i = 0; i < n; i++)
        {
            const float* u = &in[i * 24]; float* o = &out[i * 24];
            EnvironmentQuadLight e; e.NodeX = uint(u[0]); e.NodeY = uint(u[1]); e.NodeDim = uint(u[2]); e.Weight = u[3]; e.Radiance = float3(u[4], u[5], u[6]);
            const PolymorphicLightInfoFull full = e.Store(uint(u[7]));
            const uint words[12] = { asuint(full.Base.Center.x), asuint(full.Base.Center.y), asuint(full.Base.Center.z), full.Base.ColorTypeAndFlags, full.Base.Direction1, full.Base.Direction2, full.Base.Scalars,
                                     full.Base.LogRadiance, full.Extended.IesProfileIndex, full.Extended.PrimaryAxis, full.Extended.CosConeAngleAndSoftness, full.Extended.UniqueID };
            memcpy(o, words, 48);
            shimIdentityEnv();
            if (u[8] != 0.0f || u[9] != 0.0f || u[10] != 0.0f) for (int a = 0; a < 3; a++) for (int c = 0; c < 3; c++) { g_bridge.env.Transform.m[a][c] = u[8 + 3 * a + c]; g_bridge.env.InvTransform.m[c][a] = u[8 + 3 * a + c]; }
            const float3 viewer(u[19], u[20], u[21]);
            const PolymorphicLightSample r = PolymorphicLight::CalcSample(full, float2(u[17], u[18]), viewer);
            o[12] = r.Position.x; o[13] = r.Position.y; o[14] = r.Position.z; o[15] = r.Normal.x; o[16] = r.Normal.y; o[17] = r.Normal.z; o[18] = r.Radiance.x; o[19] = r.Radiance.y; o[20] = r.Radiance.z;
            const EnvironmentQuadLight c = EnvironmentQuadLight::Create(full);
            o[21] = r.SolidAnglePdf; o[22] = c.CalcSolidAnglePdfForMIS(viewer, r.Position); o[23] = c.GetPower();
        }
    }
    else if (std::string(argv[1]) == "texlod")
    {   // Rendering/Materials/TexLODHelpers.hlsli:40-161: the ray cone (fp16-packed width / spread angle), its propagation, the per-triangle LOD constant and computeLOD - what
        // every material texture fetch of the path derives its MIP level from.  40 floats in, 8 out
        const size_t n = in.size() / 40; out.assign(n * 8, 0.0f);
        for (size_t i = 0; i < n; i++)
        {
            const float* u = &in[i * 40]; float* o = &out[i * 8];
            float3 v[3] = { float3(u[0], u[1], u[2]), float3(u[3], u[4], u[5]), float3(u[6], u[7], u[8]) }; float2 t[3] = { float2(u[9], u[10]), float2(u[11], u[12]), float2(u[13], u[14]) };
            const float3x3 M(u[15], u[16], u[17], u[18], u[19], u[20], u[21], u[22], u[23]);       // (float3x3)transform; the bridge passes its transpose and multiplies from the left
            float3x3 Mt; for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) Mt[r][c] = M[c][r];
            o[0] = computeRayConeTriangleLODValue(v, t, Mt);
            RayCone rc = RayCone::make(u[24], u[25]); o[1] = rc.getWidth(); o[2] = rc.getSpreadAngle();
            rc = rc.propagateDistance(u[26]); o[3] = rc.getWidth();
            const float3 dir = normalize(float3(u[27], u[28], u[29])), nrm = normalize(float3(u[30], u[31], u[32]));
            o[4] = rc.computeLOD(o[0], dir, nrm, true); o[5] = rc.computeLOD(o[0], dir, nrm, false);
            rc = rc.addToSpreadAngle(u[33]); o[6] = rc.getSpreadAngle(); o[7] = SafeLog2(u[34]);
        }
    }
    else if (std::string(argv[1]) == "tonemap")
    {   // Rtxpt/ToneMapper/ToneMapping.ps.hlsli:31-129: the six operators through toneMap().  8 floats in (rgb, operator, whiteMaxLuminance, whiteScale), 4 out (rgb, luminance)
        const size_t n = in.size() / 8; out.assign(n * 4, 0.0f);
        for (size_t i = 0; i < n; i++)
        {
            const float* u = &in[i * 8]; float* o = &out[i * 4];
            gParams.toneMapOperator = uint(u[3]); gParams.whiteMaxLuminance = u[4]; gParams.whiteScale = u[5];
            const float3 c = toneMap(float3(u[0], u[1], u[2])); o[0] = c.x; o[1] = c.y; o[2] = c.z; o[3] = calcLuminance(float3(u[0], u[1], u[2]));
        }
    }
    else if (std::string(argv[1]) == "spheres")
    {   // Lighting/PolymorphicLight.hlsli: the analytic sphere / spot light.  The record is assembled here from the reference's own packers (PackColor, NDirToOctUnorm32, f32tof16:
        // the host does this in LightsBaker.cpp), then SphereLight::Create, CalcSample (incl. evaluateLightShaping), CalcSolidAnglePdfForMIS, GetPower.  24 floats in, 24 out
        // (words 0-11: Base + Extended as bit patterns)
        const size_t n = in.size() / 24; out.assign(n * 24, 0.0f);
        for (size_t i = 0; i < n; i++)
        {
            const float* u = &in[i * 24]; float* o = &out[i * 24];
            PolymorphicLightInfoFull full = PolymorphicLightInfoFull{}; full.Extended = PolymorphicLightInfoEx::empty();
            full.Base.Center = float3(u[0], u[1], u[2]); full.Base.ColorTypeAndFlags = 0; full.Base.Direction1 = 0; full.Base.Direction2 = 0; full.Base.LogRadiance = 0;
            full.Base.Scalars = f32tof16(u[3]);
            PolymorphicLight::PackColor(float3(u[4], u[5], u[6]), full.Base);
            full.Base.ColorTypeAndFlags |= uint(PolymorphicLightType::kSphere) << kPolymorphicLightTypeShift;
            if (u[7] > 0.5f)
            {   // spot: shaping enabled, axis / cone as the host packs them
                full.Base.ColorTypeAndFlags |= kPolymorphicLightShapingEnableBit | (u[8] > 0.5f ? kPolymorphicLightShapingUseMinFalloff : 0u);
                full.Extended.PrimaryAxis = NDirToOctUnorm32(normalize(float3(u[9], u[10], u[11])));
                full.Extended.CosConeAngleAndSoftness = f32tof16(u[12]) | (f32tof16(u[13]) << 16);
            }
            const uint words[12] = { asuint(full.Base.Center.x), asuint(full.Base.Center.y), asuint(full.Base.Center.z), full.Base.ColorTypeAndFlags, full.Base.Direction1, full.Base.Direction2, full.Base.Scalars,
                                     full.Base.LogRadiance, full.Extended.IesProfileIndex, full.Extended.PrimaryAxis, full.Extended.CosConeAngleAndSoftness, full.Extended.UniqueID };
            memcpy(o, words, 48);
            const SphereLight s = SphereLight::Create(full);
            const float3 viewer(u[16], u[17], u[18]);
            const PolymorphicLightSample r = PolymorphicLight::CalcSample(full, float2(u[14], u[15]), viewer);       // the dispatcher: SphereLight::CalcSample x evaluateLightShaping
            o[12] = r.Position.x; o[13] = r.Position.y; o[14] = r.Position.z; o[15] = r.Normal.x; o[16] = r.Normal.y; o[17] = r.Normal.z; o[18] = r.Radiance.x; o[19] = r.Radiance.y; o[20] = r.Radiance.z;
            o[21] = r.SolidAnglePdf; o[22] = s.CalcSolidAnglePdfForMIS(viewer, r.Position); o[23] = s.GetPower();
        }
    }
    else if (std::string(argv[1]) == "lights")
    {   // Lighting/PolymorphicLight.hlsli: an emissive triangle through TriangleLight::Store (what LightsBaker writes into the light buffer: PackColor, half-packed edges),
        // TriangleLight::Create (what NEE reads back), CalcSample, CalcSolidAnglePdfForMIS, GetPower.  24 floats in, 24 out (words 0-7: the PolymorphicLightInfo as bit patterns)
        const size_t n = in.size() / 24; out.assign(n * 24, 0.0f);
        for (size_t i = 0; i < n; i++)
        {
            const float* u = &in[i * 24]; float* o = &out[i * 24];
            TriangleLight t; t.base = float3(u[0], u[1], u[2]); t.edge1 = float3(u[3], u[4], u[5]); t.edge2 = float3(u[6], u[7], u[8]); t.radiance = float3(u[9], u[10], u[11]);
            const PolymorphicLightInfoFull full = t.Store(0u);
            const uint words[8] = { asuint(full.Base.Center.x), asuint(full.Base.Center.y), asuint(full.Base.Center.z), full.Base.ColorTypeAndFlags, full.Base.Direction1, full.Base.Direction2, full.Base.Scalars, full.Base.LogRadiance };
            memcpy(o, words, 32);
            const TriangleLight r = TriangleLight::Create(full);
            const float3 viewer(u[14], u[15], u[16]);
            const PolymorphicLightSample s = r.CalcSample(float2(u[12], u[13]), viewer);
            o[8] = s.Position.x; o[9] = s.Position.y; o[10] = s.Position.z; o[11] = s.Normal.x; o[12] = s.Normal.y; o[13] = s.Normal.z; o[14] = s.Radiance.x; o[15] = s.Radiance.y; o[16] = s.Radiance.z;
            o[17] = s.SolidAnglePdf; o[18] = r.CalcSolidAnglePdfForMIS(viewer, s.Position); o[19] = r.GetPower();
            o[20] = r.edge1.x; o[21] = r.edge1.y; o[22] = r.edge1.z; o[23] = r.surfaceArea;
        }
    }
    else if (std::string(argv[1]) == "helpers")
    {   // PathTracerHelpers.hlsli:26-66, :155-219: self-intersection offset, grazing-angle falloff, ray-cone growth, firefly filter (layout shared with oracle_helper_funcs)
        const size_t n = in.size() / 8; out.assign(n * 16, 0.0f);
        for (size_t i = 0; i < n; i++)
        {
            const float* u = &in[i * 8]; float* o = &out[i * 16];
            const float scale = u[6] < 0.25f ? 0.05f : (u[6] < 0.5f ? 1.0f : (u[6] < 0.75f ? 40.0f : 3000.0f));            // positions inside and outside the |p| < 1/16 switch, up to kilometres
            const float3 pos = float3(2.0f * u[0] - 1.0f, 2.0f * u[1] - 1.0f, 2.0f * u[2] - 1.0f) * scale;
            const float3 nrm = normalize(float3(2.0f * u[3] - 1.0f, 2.0f * u[4] - 1.0f, 2.0f * u[5] - 1.0f) + float3(0.0f, 1e-3f, 0.0f));
            const float3 ro = ComputeRayOrigin(pos, nrm); o[0] = ro.x; o[1] = ro.y; o[2] = ro.z;
            const float3 l = normalize(float3(2.0f * u[5] - 1.0f, 2.0f * u[6] - 1.0f, 2.0f * u[7] - 1.0f) + float3(1e-3f, 0.0f, 0.0f));
            const float from = 0.01f + 0.24f * u[7]; o[3] = ComputeLowGrazingAngleFalloff(l, nrm, from, 2.0f * from);
            o[4] = RoughnessToVariance(u[0]); o[5] = GetAngleFromGGXRoughness(u[1]); o[6] = ComputeRayConeSpreadAngleExpansionByRoughness(u[2]);
            const float pdf = u[3] < 0.05f ? 0.0f : u[3] * u[3] * 40.0f;
            o[7] = pdf == 0.0f ? 0.0f : ComputeRayConeSpreadAngleExpansionByScatterPDF(pdf);
            o[8] = ComputeNewScatterFireflyFilterK(lpfloat(u[4]), pdf, u[5]);
            const lpfloat3 ff = FireflyFilter(lpfloat3(lpfloat(u[0] * 8.0f), lpfloat(u[1] * 8.0f), lpfloat(u[2] * 8.0f)), lpfloat(2.0f + u[3]), lpfloat(u[4])); o[9] = ff.x; o[10] = ff.y; o[11] = ff.z;
            o[12] = FireflyFilterShort(u[0] * 8.0f, 2.0f + u[3], u[4]);
            o[13] = BalanceHeuristic(1.0f, u[5] * 3.0f, 1.0f, u[7] * 3.0f);
            // MatrixRotateFromTo (:227-270), the rotation a delta transmission applies to the stable plane's image transform: first two rows (the third follows from orthogonality... it is
            // compared through the oracle's own packed form elsewhere); 14, 15 = m[0][0], m[1][2]
            { const float3x3 m = MatrixRotateFromTo(nrm, l); o[14] = m[0][0] + m[0][1] * 0.5f + m[0][2] * 0.25f + m[1][0] * 0.125f + m[1][1] * 3.0f; o[15] = m[1][2] + m[2][0] * 0.5f + m[2][1] * 0.25f + m[2][2] * 3.0f; }
        }
    }
    else if (std::string(argv[1]) == "utils")
    {   // Utils/Utils.hlsli: MIS heuristics, octahedral encodings, luminance clamp, the fast approximations (inputs: 8 uniforms per record; layout shared with oracle_utils_funcs)
        const size_t n = in.size() / 8; out.assign(n * 24, 0.0f);
        for (size_t i = 0; i < n; i++)
        {
            const float* u = &in[i * 8]; float* o = &out[i * 24];
            const float3 dir = normalize(float3(2.0f * u[0] - 1.0f, 2.0f * u[1] - 1.0f, 2.0f * u[2] - 1.0f) + float3(1e-3f, 0.0f, 0.0f));
            const float n0 = 1.0f + floor(u[4] * 4.0f), n1 = 1.0f + floor(u[6] * 4.0f), p0 = u[5] * 3.0f, p1 = u[7] * 3.0f;
            const float3 lc = LuminanceClamp(float3(u[0], u[1], u[2]) * 4.0f, 0.2f + u[3]); o[0] = lc.x; o[1] = lc.y; o[2] = lc.z;
            o[3] = EvalMIS(MISHeuristic::Balance, n0, p0, n1, p1); o[4] = EvalMIS(MISHeuristic::PowerTwo, n0, p0, n1, p1); o[5] = EvalMIS(MISHeuristic::Balance, n0, p0, n1, p1, 2.0f, u[3]);
            const float2 eo = Encode_Oct(dir); o[6] = eo.x; o[7] = eo.y;
            const float3 dn = Decode_Oct(float2(u[0], u[1])); o[8] = dn.x; o[9] = dn.y; o[10] = dn.z;
            const uint p32 = NDirToOctUnorm32(dir); memcpy(&o[11], &p32, 4);
            const uint q32 = (uint(u[4] * 65534.0f) & 0xffffu) | (uint(u[5] * 65534.0f) << 16); const float3 d32 = OctToNDirUnorm32(q32); o[12] = d32.x; o[13] = d32.y; o[14] = d32.z;
            const uint p30 = NDirToOctUnorm30(dir); memcpy(&o[15], &p30, 4);
            const uint q30 = (uint(u[6] * 32767.0f) & 0x7fffu) | ((uint(u[7] * 32767.0f) & 0x7fffu) << 15); const float3 d30 = OctToNDirUnorm30(q30); o[16] = d30.x; o[17] = d30.y; o[18] = d30.z;
            o[19] = FastSqrt(u[0] * 10.0f); o[20] = FastACos(2.0f * u[1] - 1.0f); o[21] = WeightedAverage(u[0], u[1], u[2], u[3]);
            o[22] = RelativelyEqual(u[0], u[0] * (1.0f + 2e-4f * u[1])) ? 1.0f : 0.0f; o[23] = Reinhard(float3(u[0], u[1], u[2]) * 3.0f).y;
        }
    }
    else
    {
        const size_t n = in.size() / 8; out.assign(n * 40, 0.0f);
        for (size_t i = 0; i < n; i++)
        {
            const float* u = &in[i * 8]; float* o = &out[i * 40];
            // inputs derived from the uniforms the same way on both sides (tests/golden/make_bsdf_golden.py documents them)
            const float alpha = 0.0064f + u[0] * u[0] * 0.99f, cosI = 0.001f + 0.999f * u[1], cosO = 0.001f + 0.999f * u[2], eta = u[3] < 0.5f ? 1.0f / (1.0f + 1.2f * u[4]) : 1.0f + 1.2f * u[4];
            const float phi = 6.2831853f * u[5], sI = sqrt(max(0.0f, 1.0f - cosI * cosI));
            const float3 wi(sI * cos(phi), sI * sin(phi), cosI);
            o[0] = evalFresnelSchlick(u[6], 1.0f, cosI);
            const float3 fs = evalFresnelSchlick(float3(u[6], u[7], u[0]), float3(1.0f, 1.0f, 1.0f), cosO); o[1] = fs.x; o[2] = fs.y; o[3] = fs.z;
            float cosT = -1.0f; o[4] = evalFresnelDielectric(eta, cosI, cosT); o[5] = cosT;
            float cosT2 = -1.0f; o[6] = evalFresnelDielectric(eta, -cosI, cosT2); o[7] = cosT2;
            o[8] = evalNdfGGX(alpha, cosO); o[9] = evalLambdaGGX(alpha * alpha, cosI); o[10] = evalG1GGX(alpha * alpha, cosI);
            o[11] = evalMaskingSmithGGXCorrelated(alpha, cosI, cosO); o[12] = evalMaskingSmithGGXSeparable(alpha, cosI, cosO);
            const float3 h = sampleGGX_BVNDF(alpha, wi, float2(u[6], u[7])); o[13] = h.x; o[14] = h.y; o[15] = h.z;
            o[16] = evalPdfGGX_BVNDF(alpha, wi, h);
            const float3 hv = sampleGGX_VNDF(alpha, wi, float2(u[6], u[7])); o[17] = hv.x; o[18] = hv.y; o[19] = hv.z;
            o[20] = evalPdfGGX_VNDF(alpha, wi, hv);
            const float3 asi = approxSpecularIntegralGGX(float3(u[6], u[7], u[0]), alpha, cosI); o[21] = asi.x; o[22] = asi.y; o[23] = asi.z;
            float pdf = 0; const float3 ch = sample_cosine_hemisphere_concentric(float2(u[6], u[7]), pdf); o[24] = ch.x; o[25] = ch.y; o[26] = ch.z; o[27] = pdf;
            const float2 dk = sample_disk_concentric(float2(u[0], u[1])); o[28] = dk.x; o[29] = dk.y;
            const float3 ps = perp_stark(wi); o[30] = ps.x; o[31] = ps.y; o[32] = ps.z;
            const float2 oc = ndir_to_oct_equal_area_unorm(wi); o[33] = oc.x; o[34] = oc.y; const float3 od = oct_to_ndir_equal_area_unorm(float2(u[2], u[3])); o[35] = od.x; o[36] = od.y; o[37] = od.z;
            o[38] = Luminance(float3(u[0], u[1], u[2])); o[39] = Average(float3(u[0], u[1], u[2]));
        }
    }
    writeAll(argv[3], out);
    return 0;
}
