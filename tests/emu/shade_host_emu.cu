// TEST INFRASTRUCTURE ONLY - never linked into librtxpt_b200*.so.  Host build of the product's per-vertex shading functions (rtxpt_b200/csrc/shade.cuh: shadeHit / shadeMiss with
// the NEE-AT sampler of neeat.cuh, bsdf.cuh, the sample generators of device_math.cuh), so that tests/test_shade_port.py can hold the CUDA source - not only the oracle - to the golden
// vectors generated from the reference's own PathTracer.hlsli (tests/golden/hit_golden.npz) on the CPU.  The device-only spellings of those headers are mapped onto host equivalents
// below (bit casts, read-only loads, IEEE single operations); texture fetches and the triangle gather are what the golden's stub bridge replaces (PT_HOST_EMU hooks in shade.cuh).
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstring>
#include <cstdint>
#include <cmath>
#include <vector>
namespace emu {
inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; } inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline float i2f(int i) { float f; memcpy(&f, &i, 4); return f; } inline int f2i(float f) { int i; memcpy(&i, &f, 4); return i; }
inline uint32_t brev(uint32_t v) { uint32_t r = 0; for (int i = 0; i < 32; i++) r |= ((v >> i) & 1u) << (31 - i); return r; }
template <typename T> inline T ldg(const T* p) { return *p; }
template <typename T, typename U> inline T atomicAddHost(T* p, U v) { const T old = *p; *p = old + T(v); return old; }
template <typename T> T tex2DLodStub(cudaTextureObject_t, float, float, float) { return T{}; }
template <typename T> T tex2DLayeredLodStub(cudaTextureObject_t, float, float, int, float) { return T{}; }
}
#define __uint_as_float(x) emu::u2f(x)
#define __float_as_uint(x) emu::f2u(x)
#define __int_as_float(x) emu::i2f(x)
#define __float_as_int(x) emu::f2i(x)
#define __ldg(p) emu::ldg(p)
#define __ldcs(p) (*(p))
#define __stcs(p, v) (*(p) = (v))
#define __popc(x) __builtin_popcount(x)
#define __ffs(x) __builtin_ffs(int(x))
#define __clz(x) ((x) == 0 ? 32 : __builtin_clz(x))
#define __brev(x) emu::brev(x)
#define atomicAdd(p, v) emu::atomicAddHost(p, v)
#define __ballot_sync(m, p) ((p) ? 1u : 0u)
#define __shfl_sync(m, v, l) (v)
#define __fdiv_rn(a, b) ((a) / (b))
#define __fmul_rn(a, b) ((a) * (b))
#define __fadd_rn(a, b) ((a) + (b))
#define __fsub_rn(a, b) ((a) - (b))
#define tex2DLod emu::tex2DLodStub
#define tex2DLayeredLod emu::tex2DLayeredLodStub
// constant-memory tables (Sobol directions) are plain initialised arrays in the host build - a __constant__ variable's host shadow is not
#ifndef __CUDA_ARCH__
#undef __constant__
#define __constant__
#endif
#define PT_DEVICE __host__ __device__ __forceinline__
#define PT_HOST_EMU 1
#include "../../rtxpt_b200/csrc/shade.cuh"

using namespace pt;

// ---- what the golden's stub bridge supplies -------------------------------------------------------------------------------------------------------------------------------
static thread_local const Surface* gSurface = nullptr;
__host__ __device__ void pt::emuLoadSurface(Surface& s) {
#ifndef __CUDA_ARCH__
    s = *gSurface;
#endif
}
__host__ __device__ float3 pt::emuEnvCube(float3 d, float lod) { const float k = exp2f(-lod); return mk3((0.5f + 0.5f * d.x) * k, (0.5f + 0.25f * d.y) * k, (0.75f + 0.25f * d.z) * k); }
static thread_local float3 gCamPos, gCamBase, gCamDx, gCamDy;
__host__ __device__ void pt::emuCameraRay(uint id, float3& origin, float3& dir)
{
#ifndef __CUDA_ARCH__
    const float3 d = gCamBase + gCamDx * float(id >> 16) + gCamDy * float(id & 0xFFFFu); origin = gCamPos; dir = d / sqrtf(dot3(d, d));       // normalize( ) as the stub bridge spells it: x / length( x )
#endif
}
__host__ __device__ float3 pt::emuMotionVector(float3 posW, float3 prevPosW) { return (prevPosW - posW) * 0.5f; }
static bool visibilityRule(float4 o, float4 d)
{   // ShimVisibilityRule of oracle/ref_bridge_stub.h
    const uint32_t h = emu::f2u(o.x) ^ (emu::f2u(o.y) >> 1) ^ (emu::f2u(o.z) >> 2) ^ emu::f2u(d.x) ^ (emu::f2u(d.y) >> 1) ^ (emu::f2u(d.z) >> 2) ^ emu::f2u(o.w);
    return (h & 3u) != 0u;
}
static const std::vector<uint32_t>& envLookup()
{
    static std::vector<uint32_t> m;
    if (m.empty()) { m.resize(1024 * 1024); for (uint32_t y = 0; y < 1024; y++) for (uint32_t x = 0; x < 1024; x++) m[y * 1024 + x] = ((x >> 6) + (y >> 6) * 3u) & 3u; }
    return m;
}

// One path vertex of the reference-mode shade kernel with NEE-AT (k_shade<.., true, true, true>: shadeHit / shadeMiss) followed by what k_trace_shadow does with the vertex's
// shadow record (kernels.cu:190-222: radiance into L, feedback reservoir, the Russian-roulette outcome of a visible sample, which the path's next shadeHit applies - applied here),
// on records of tests/golden/hit_golden.npz (layout: oracle/ref_kat_bsdf_main.cpp, "hit" mode; ops 0 and 1).  Returns 0, or -1 for a record this path does not cover.
template <int MODE> static int shadeVertex(const float* r, float* o)
{
    constexpr bool kRealtime = MODE != kModeReference, kNeeat = MODE != kModeBuildStablePlanes;     // k_shade< .., NEEAT >, k_rt_shade< BUILD, .. >, k_rt_shade< FILL, .., NEEAT >
    for (int k = 0; k < 128; k++) o[k] = 0.0f;
    if (r[27] > 1.0f) return -1;
    LaunchParams p; memset(&p, 0, sizeof(p));
    // surface
    Surface s; memset(&s, 0, sizeof(s));
    s.posW = mk3(r[28], r[29], r[30]); s.faceN = mk3(r[31], r[32], r[33]); s.V = mk3(-r[23], -r[24], -r[25]); s.N = mk3(r[34], r[35], r[36]); s.T = mk3(r[37], r[38], r[39]); s.B = mk3(r[40], r[41], r[42]);
    s.vertexN = mk3(r[43], r[44], r[45]); s.frontFacing = r[46] != 0.0f; s.nestedPriority = uint(r[47]); s.thin = r[49] != 0.0f; s.psdExclude = r[50] != 0.0f; s.materialID = uint(r[51]); s.IoR = r[52];
    s.shadowNoLFadeout = r[53]; s.emission = mk3(r[54], r[55], r[56]); s.psdBlockMVs = r[57] != 0.0f; s.psdDominantDeltaLobeP1 = uint(r[58]);
    const float* b = r + 42;
    s.bsdf.diffuse = mk3(b[18], b[19], b[20]); s.bsdf.roughness = b[21]; s.bsdf.specular = mk3(b[22], b[23], b[24]); s.bsdf.metallic = b[25]; s.bsdf.transmission = mk3(b[26], b[27], b[28]);
    s.bsdf.diffuseTransmission = b[29]; s.bsdf.specularTransmission = b[30]; s.bsdf.eta = b[31];
    s.interiorIoR = r[74]; s.neeTriangleLightIndex = r[75] < 0 ? 0xFFFFFFFFu : uint(r[75]); s.neeAnalyticLightIndex = r[76] < 0 ? 0xFFFFFFFFu : uint(r[76]); s.prevPosW = mk3(r[77], r[78], r[79]);
    gSurface = &s;
    // constants
    RtxptPathTracerConstants& c = p.c;
    c.imageWidth = c.imageHeight = 8; c.bounceCount = uint(r[80]); c.diffuseBounceCount = uint(r[81]); c.NEEEnabled = 1; c.NEEType = 2; c.NEECandidateSamples = uint(r[83]); c.NEEFullSamples = uint(r[84]);
    c.fireflyFilterThreshold = r[85]; c.enableRussianRoulette = 1; c.enableLDSamplerForBSDF = 1; c.nestedDielectricsQuality = 1; c.EnvironmentMapDiffuseSampleMIPLevel = r[93]; c.NEEATFeedback = 1;
    for (int a = 0; a < 3; a++) for (int k = 0; k < 3; k++) { c.envMap.Transform[a * 4 + k] = r[950 + 3 * a + k]; c.envMap.InvTransform[k * 4 + a] = r[950 + 3 * a + k]; }
    for (int k = 0; k < 3; k++) c.envMap.ColorMultiplier[k] = r[959];
    // scene side: materials, lights
    RtxptMaterialData mats[8]; memset(mats, 0, sizeof(mats));
    for (int m = 0; m < 8; m++) { mats[m].IoR = r[96 + m]; for (int k = 0; k < 3; k++) mats[m].VolumeAttenuationColor[k] = r[104 + 3 * m + k]; mats[m].VolumeAttenuationDistance = r[128 + m]; }
    p.scene.materials = mats; p.scene.materialCount = 8;
    LightInfo lights[16]; uint4 lightsEx[16]; uint32_t counters[16], indices[64], local[512], proxyCount = uint(r[90]);
    for (int k = 0; k < 16; k++) { memcpy(&lights[k], r + 728 + 12 * k, 32); memcpy(&lightsEx[k], r + 728 + 12 * k + 8, 16); counters[k] = uint(r[136 + k]); }
    for (int k = 0; k < 64; k++) indices[k] = uint(r[152 + k]);
    memcpy(local, r + 216, sizeof(local));
    p.scene.lights = lights; p.scene.lightCount = 16; p.scene.analyticLightCount = 16; p.scene.envEnabled = 1; p.scene.envLookupMap = envLookup().data();
    // shade.cuh reads the Extended record of light i at lightsEx[ uint( i - 5368 ) ] (the analytic lights follow the 5368 environment nodes): make that land on entry i of this 16-light table
    p.scene.lightsEx = reinterpret_cast<const uint4*>(reinterpret_cast<uintptr_t>(lightsEx) - (uintptr_t(0x100000000ull) - 5368ull) * sizeof(uint4));
    p.scene.proxyCounters = counters; p.scene.proxyIndices = indices; p.scene.samplingProxyCount = proxyCount;
    float fbWeight[64]; uint32_t fbCand[64]; for (int k = 0; k < 64; k++) { fbWeight[k] = 0.0f; fbCand[k] = 0xFFFFFFFFu; }
    p.na.W = p.na.H = 8; p.na.tilesX = p.na.tilesY = 2; p.na.lightCount = 16; p.na.neeType = 2; p.na.jitterX = uint(r[88]); p.na.jitterY = uint(r[89]); p.na.localToGlobalSampleRatio = r[86];
    p.na.screenSpaceVsWorldSpaceThreshold = r[91]; p.na.temporalFeedbackRequired = uint(r[92]); p.na.fbWeight = fbWeight; p.na.fbCandidate = fbCand; p.na.localSamplingBuffer = local;
    p.na.proxyCounters = counters; p.na.proxyIndices = indices; p.na.samplingProxyCount = &proxyCount;
    uint32_t rrFix[1] = { 0u }; p.naRrFix = rrFix;
    // realtime passes: the stable planes of an 8 x 8 image, the pixel's entries from the record; the stub bridge's camera
    const uint32_t pid = emu::f2u(r[3]), px = (pid >> 16) & 7u, py = pid & 7u;
    RtxptStablePlane planes[3 * 64]; uint32_t header[4 * 64]; uint2 stableRadiance[64]; float specHitT[64], depth[64]; uint2 motion[64]; uint32_t throughput[64];
    if (kRealtime)
    {
        memset(planes, 0, sizeof(planes)); memset(header, 0xFF, sizeof(header)); memset(stableRadiance, 0, sizeof(stableRadiance)); memset(motion, 0, sizeof(motion)); memset(throughput, 0, sizeof(throughput));
        for (int k = 0; k < 64; k++) { specHitT[k] = 0.0f; depth[k] = -1.0f; }
        p.rt.planes = planes; p.rt.header = header; p.rt.stableRadiance = stableRadiance; p.rt.specularHitT = specHitT; p.rt.lineStride = 8; p.rt.planeStride = 64; p.rt.activePlaneCount = 3;
        p.rt.maxVertexDepth = uint(r[943]); p.rt.allowPSR = uint(r[944]); p.rt.attenuation = r[87]; p.depth = depth; p.motionVectors = motion; p.throughput = throughput;
        for (int k = 0; k < 16; k += 5) p.worldToClip[k] = 1.0f;
        for (uint32_t k = 0; k < 4; k++) memcpy(&header[(k * 8 + py) * 8 + px], r + 920 + k, 4);
        for (uint32_t k = 0; k < 3; k++) memcpy(planes[planeAddress(p.rt, (px << 16) | py, k)].PackedNoisyRadianceAndSpecAvg, r + 924 + 2 * k, 8);
        specHitT[py * 8 + px] = r[930];
        stableRadiance[py * 8 + px] = make_uint2(f32tof16(r[946]) | (f32tof16(r[947]) << 16), f32tof16(r[948]) | (f32tof16(r[949]) << 16));
        gCamPos = mk3(r[931], r[932], r[933]); gCamBase = mk3(r[934], r[935], r[936]); gCamDx = mk3(r[937], r[938], r[939]); gCamDy = mk3(r[940], r[941], r[942]);
        p.firstSampleIndex = uint(r[82]);
    }
    // the path: the 80-byte payload in the order of wavefront.cuh's five state words (the stableBranchID word carries the sample index in reference mode)
    uint32_t w[20]; memcpy(w, r, 80);
    PathRegs path;
    path.origin = mk3(emu::u2f(w[0]), emu::u2f(w[1]), emu::u2f(w[2])); path.id = w[3]; path.dir = mk3(emu::u2f(w[4]), emu::u2f(w[5]), emu::u2f(w[6])); path.sceneLength = emu::u2f(w[7]);
    path.thpXY = w[8]; path.thpZ = w[9]; path.lXY = w[10]; path.lZW = w[11]; path.interior0 = w[12]; path.interior1 = w[13]; path.packedCounters = w[14]; path.rayCone = w[16];
    path.pack0 = w[17]; path.pack1 = w[18]; path.flagsAndVertexIndex = w[19]; path.sampleIndex = kRealtime ? w[15] : uint(r[82]);
    HitOutputs out; out.continuePath = false; out.emitShadow = false; out.naRecord = make_uint4(0xFFFFFFFFu, 0u, 0u, 0u);
    if (r[27] == 1.0f) shadeMiss<false, MODE, kNeeat>(p, path);
    else shadeHit<false, true, MODE, kNeeat>(p, path, 0u, make_float4(r[26], 0.25f, 0.25f, 0.0f), out);
    // k_trace_shadow's half of ProcessLightSample
    if (out.emitShadow)
    {
        const bool visible = visibilityRule(out.shadow.originTMax, out.shadow.dirPath);
        o[20] = 1.0f; o[21] = out.shadow.originTMax.x; o[22] = out.shadow.originTMax.y; o[23] = out.shadow.originTMax.z; o[24] = out.shadow.dirPath.x; o[25] = out.shadow.dirPath.y; o[26] = out.shadow.dirPath.z;
        o[27] = out.shadow.originTMax.w; o[28] = visible ? 1.0f : 0.0f;
        if (visible)
        {
            const uint2 rad = out.shadow.radiance;
            const float rx = f16tof32(rad.x & 0x7FFFu), ry = f16tof32((rad.x >> 16) & 0x7FFFu), rz = f16tof32(rad.y), rw = f16tof32(rad.y >> 16);
            if (rx > 0 || ry > 0 || rz > 0 || rw > 0)
            {
                float lx, ly, lz, lw;
                if (kRealtime)
                {
                    const float a = p.rt.attenuation, spec = (rad.x & 0x00008000u) ? rw : ((rad.x & 0x80000000u) ? (rx + ry + rz) / 3.0f : 0.0f);
                    lx = f16tof32(path.lXY) + rx * a; ly = f16tof32(path.lXY >> 16) + ry * a; lz = f16tof32(path.lZW) + rz * a; lw = f16tof32(path.lZW >> 16) + spec * a;
                }
                else { lx = f16tof32(path.lXY) + rx; ly = f16tof32(path.lXY >> 16) + ry; lz = f16tof32(path.lZW) + rz; lw = f16tof32(path.lZW >> 16); }
                path.lXY = packHalf2NoClamp(clampf(lx, 0.f, kHalfMax), clampf(ly, 0.f, kHalfMax)); path.lZW = packHalf2NoClamp(clampf(lz, 0.f, kHalfMax), clampf(lw, 0.f, kHalfMax));
            }
            const uint4 fb = out.naRecord;
            if (fb.w & 0x80000000u)
            {
                neeat::Reservoir::at(p.na.fbWeight, p.na.fbCandidate, size_t(path.id & 7u) * 8 + ((path.id >> 16) & 7u)).add(emu::u2f(fb.z), fb.x & 0x7FFFFFFFu, emu::u2f(fb.y), (fb.x & 0x80000000u) != 0);
                path.setFlag(kPFTerminateAtNextBounce, (fb.w & 0x40000000u) != 0); path.setMisInfo_RuRu(path.misInfo(), f16tof32(fb.w & 0xFFFFu));     // shadeHit's prologue at the path's next vertex
            }
        }
    }
    uint32_t q[20] = { emu::f2u(path.origin.x), emu::f2u(path.origin.y), emu::f2u(path.origin.z), path.id, emu::f2u(path.dir.x), emu::f2u(path.dir.y), emu::f2u(path.dir.z), emu::f2u(path.sceneLength),
                       path.thpXY, path.thpZ, path.lXY, path.lZW, path.interior0, path.interior1, path.packedCounters, kRealtime ? path.sampleIndex : w[15], path.rayCone, path.pack0, path.pack1, path.flagsAndVertexIndex };
    memcpy(o, q, 80);
    if (kRealtime)
    {
        o[37] = specHitT[py * 8 + px];
        for (uint32_t k = 0; k < 3; k++) { const RtxptStablePlane& sp = planes[planeAddress(p.rt, (px << 16) | py, k)]; memcpy(o + 41 + 2 * k, sp.PackedNoisyRadianceAndSpecAvg, 8); memcpy(o + 56 + 20 * k, &sp, 80); }
        for (uint32_t k = 0; k < 4; k++) memcpy(o + 47 + k, &header[(k * 8 + py) * 8 + px], 4);
        const uint2 sr = stableRadiance[py * 8 + px]; o[52] = f16tof32(sr.x); o[53] = f16tof32(sr.x >> 16); o[54] = f16tof32(sr.y); o[55] = f16tof32(sr.y >> 16);
        if (MODE == kModeBuildStablePlanes) o[r[27] == 1.0f ? 31 : 29] = depth[py * 8 + px] != -1.0f ? 1.0f : 0.0f;
    }
    const uint32_t at = (path.id & 7u) * 8 + ((path.id >> 16) & 7u);
    o[39] = fbWeight[at]; memcpy(o + 40, &fbCand[at], 4);
    return 0;
}
extern "C" int shade_emu_reference_vertex(const float* r, float* o) { return shadeVertex<kModeReference>(r, o); }
extern "C" int shade_emu_build_vertex(const float* r, float* o) { return shadeVertex<kModeBuildStablePlanes>(r, o); }
extern "C" int shade_emu_fill_vertex(const float* r, float* o) { return shadeVertex<kModeFillStablePlanes>(r, o); }

// k_debug_bsdf's body (shade_kernels.cu) on the host: StandardBSDF eval / evalPdf / sample / getLobes of bsdf.cuh on the records of tests/golden/bsdf_golden.npz (36 floats in, 16 out)
extern "C" void shade_emu_bsdf(const float* in, uint32_t count, float* out)
{
    for (uint32_t i = 0; i < count; i++)
    {
        const float* r = in + size_t(i) * 36; float* o = out + size_t(i) * 16;
        BsdfParams d;
        d.diffuse = mk3(r[18], r[19], r[20]); d.roughness = r[21]; d.specular = mk3(r[22], r[23], r[24]); d.metallic = r[25];
        d.transmission = mk3(r[26], r[27], r[28]); d.diffuseTransmission = r[29]; d.specularTransmission = r[30]; d.eta = r[31];
        BsdfSetup b; b.init(mk3(r[6], r[7], r[8]), mk3(r[9], r[10], r[11]), mk3(r[3], r[4], r[5]), mk3(r[0], r[1], r[2]), r[32] != 0.0f, d);
        const float3 wo = mk3(r[12], r[13], r[14]);
        const float4 e = b.eval(wo);
        o[0] = e.x; o[1] = e.y; o[2] = e.z; o[3] = e.w; o[4] = b.pdf(wo);
        BsdfSample s; const bool valid = b.sample(r[15], r[16], r[17], s);
        o[5] = valid ? 1.0f : 0.0f; o[6] = s.wo.x; o[7] = s.wo.y; o[8] = s.wo.z; o[9] = s.pdf; o[10] = s.weight.x; o[11] = s.weight.y; o[12] = s.weight.z;
        o[13] = float(s.lobe); o[14] = s.lobeP; o[15] = float(bsdfLobes(d));
    }
}
// batch form: `count` records of 1024 floats, 128 floats out each; status[i] = 0 where the record was run
extern "C" void shade_emu_vertices(const float* in, uint32_t count, float* out, int32_t* status, uint32_t mode)
{
    for (uint32_t i = 0; i < count; i++)
    {
        const float* r = in + size_t(i) * 1024; float* o = out + size_t(i) * 128;
        status[i] = mode == 0 ? shadeVertex<kModeReference>(r, o) : (mode == 1 ? shadeVertex<kModeBuildStablePlanes>(r, o) : shadeVertex<kModeFillStablePlanes>(r, o));
    }
}
