// TEST INFRASTRUCTURE ONLY - never linked into librtxpt_b200*.so, never loaded by rtxpt_b200/.
// Host build of the product's ReBLUR pass bodies (rtxpt_b200/csrc/reblur_passes.cuh: __host__ __device__ functions, the same source the CUDA kernels wrap) with the product's host
// constants (reblur_host.h), run pixel by pixel on the CPU with the dispatch order of launchReblurFrame (reblur_kernels.cu).  It lets tests/test_reblur_port.py hold the CUDA
// source to the oracle (oracle/reblur.h) without a GPU: what it proves is the logic and the resource routing of the port; what it cannot prove is anything only the GPU build has
// (libdevice maths, fast-math flags, launch configuration) - that stays with tests/test_gpu_reblur.py.
#include "../../rtxpt_b200/csrc/reblur_passes.cuh"
#include "../../rtxpt_b200/csrc/reblur_host.h"
#include "../../rtxpt_b200/csrc/guides_filter.cuh"
#include "../../rtxpt_b200/csrc/envbake.cuh"
#include "../../rtxpt_b200/csrc/refit.cuh"
#include "../../rtxpt_b200/csrc/tonemap.cuh"
#include "../../rtxpt_b200/csrc/denoiser_iface.cuh"
#include "../../rtxpt_b200/csrc/skinning.cuh"
#include <vector>
#include <cstdint>

using namespace pt;

namespace {
struct Instance
{
    uint32_t W = 0, H = 0; bool valid = false; uint32_t pingPong = 0;
    std::vector<float> prevViewZ; std::vector<uint32_t> prevNormalRoughness; std::vector<unsigned short> prevInternalData, diffFast, specFast, tracking[2], diffLuma[2], specLuma[2];
    std::vector<uint2> diffHistory, specHistory;
    bool keepStages = false; std::vector<std::vector<uint2>> stageDiff, stageSpec;       // images after each of the six middle passes of the last frame (debugging aid)
    void init(uint32_t w, uint32_t h)
    {
        W = w; H = h; valid = false; pingPong = 0; const size_t n = size_t(w) * h;
        prevViewZ.assign(n, 0.0f); prevNormalRoughness.assign(n, 0u); prevInternalData.assign(n, 0); diffFast.assign(n, 0); specFast.assign(n, 0); diffHistory.assign(n, make_uint2(0, 0)); specHistory.assign(n, make_uint2(0, 0));
        for (int i = 0; i < 2; i++) { tracking[i].assign(n, 0); diffLuma[i].assign(n, 0); specLuma[i].assign(n, 0); }
    }
};
template <typename F> void forEachPixel(const rb::Params& p, F f)
{
    #pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < int(p.H); y++) for (int x = 0; x < int(p.W); x++) f(x, y);
}
}

extern "C" void* rb_emu_create() { return new Instance(); }
extern "C" void rb_emu_destroy(void* p) { delete static_cast<Instance*>(p); }
// debugging aid: keep / fetch the images after pass `stage` of the last frame (0 HitDistReconstruction, 1 PrePass, 2 TemporalAccumulation, 3 HistoryFix, 4 Blur, 5 PostBlur)
extern "C" void rb_emu_keep_stages(void* p, int on) { static_cast<Instance*>(p)->keepStages = on != 0; }
extern "C" int rb_emu_stage(void* p, uint32_t stage, uint16_t* outDiff, uint16_t* outSpec)
{
    Instance& h = *static_cast<Instance*>(p);
    if (stage >= h.stageDiff.size()) return -1;
    memcpy(outDiff, h.stageDiff[stage].data(), h.stageDiff[stage].size() * 8); memcpy(outSpec, h.stageSpec[stage].data(), h.stageSpec[stage].size() * 8);
    return 0;
}
// images: RGBA16F as 4 x uint16 per pixel; motion / disocclusionMix may be NULL; outFrames: 2 floats per pixel (accumulated frames after TemporalAccumulation)
extern "C" int rb_emu_denoise(void* instance, uint32_t W, uint32_t H, const RtxptReblurFrame* frame, const float* viewZ, const uint32_t* normalRoughness, const uint16_t* motion, const uint8_t* disocclusionMix,
                              const uint16_t* inDiff, const uint16_t* inSpec, uint16_t* outDiff, uint16_t* outSpec, float* outFrames)
{
    Instance& h = *static_cast<Instance*>(instance);
    if (h.W != W || h.H != H) h.init(W, H);
    const size_t n = size_t(W) * H, tiles = size_t((W + 15) / 16) * ((H + 15) / 16);
    rb::Params p{}; rb::fillFrameParams(p, W, H, frame, h.valid);
    std::vector<unsigned char> tileSky(tiles, 0); std::vector<uint2> tmp1Diff(n, make_uint2(0, 0)), tmp1Spec(n, make_uint2(0, 0)), tmp2Diff(n, make_uint2(0, 0)), tmp2Spec(n, make_uint2(0, 0)), oDiff(n, make_uint2(0, 0)), oSpec(n, make_uint2(0, 0));
    std::vector<unsigned short> trackingT(n, 0), diffFastT(n, 0), specFastT(n, 0); std::vector<uchar2> data1(n, make_uchar2(0, 0)); std::vector<uint32_t> data2(n, 0);
    p.viewZ = viewZ; p.normalRoughness = normalRoughness; p.motion = frame->ignoreMotionVectors ? nullptr : reinterpret_cast<const uint2*>(motion); p.disocclusionMix = disocclusionMix;
    p.inDiff = reinterpret_cast<const uint2*>(inDiff); p.inSpec = reinterpret_cast<const uint2*>(inSpec);
    p.tiles = tileSky.data(); p.tmp1Diff = tmp1Diff.data(); p.tmp1Spec = tmp1Spec.data(); p.tmp2Diff = tmp2Diff.data(); p.tmp2Spec = tmp2Spec.data();
    p.trackingTransient = trackingT.data(); p.diffFastTransient = diffFastT.data(); p.specFastTransient = specFastT.data(); p.data1 = data1.data(); p.data2 = data2.data();
    p.prevViewZ = h.prevViewZ.data(); p.prevNormalRoughness = h.prevNormalRoughness.data(); p.prevInternalData = h.prevInternalData.data(); p.diffHistory = h.diffHistory.data(); p.specHistory = h.specHistory.data();
    p.diffFast = h.diffFast.data(); p.specFast = h.specFast.data();
    const uint32_t prev = h.pingPong, curr = prev ^ 1u;
    p.trackingPrev = h.tracking[prev].data(); p.trackingCurr = h.tracking[curr].data(); p.diffLumaPrev = h.diffLuma[prev].data(); p.diffLumaCurr = h.diffLuma[curr].data(); p.specLumaPrev = h.specLuma[prev].data(); p.specLumaCurr = h.specLuma[curr].data();
    p.outDiff = oDiff.data(); p.outSpec = oSpec.data();
    // launchReblurFrame's order
    for (uint32_t ty = 0; ty < (H + 15) / 16; ty++) for (uint32_t tx = 0; tx < (W + 15) / 16; tx++)
    {
        int sky = 0;
        for (int j = 0; j < 16; j++) for (int i = 0; i < 16; i++) sky += rb::beyondDenoisingRange(p, int(tx * 16 + i), int(ty * 16 + j)) ? 1 : 0;
        tileSky[size_t(ty) * p.tilesW + tx] = sky == 256 ? 1 : 0;
    }
    h.stageDiff.clear(); h.stageSpec.clear();
    auto keep = [&](const std::vector<uint2>& d, const std::vector<uint2>& s) { if (h.keepStages) { h.stageDiff.push_back(d); h.stageSpec.push_back(s); } };
    forEachPixel(p, [&](int x, int y) { rb::hitDistReconstructionPixel(p, x, y); }); keep(tmp2Diff, tmp2Spec);
    forEachPixel(p, [&](int x, int y) { rb::spatialPixel<0>(p, x, y); }); keep(tmp1Diff, tmp1Spec);
    forEachPixel(p, [&](int x, int y) { rb::temporalAccumulationPixel(p, x, y); }); keep(tmp2Diff, tmp2Spec);
    if (outFrames) for (size_t i = 0; i < n; i++) { outFrames[2 * i] = float(data1[i].x) / 255.0f * 63.0f; outFrames[2 * i + 1] = float(data1[i].y) / 255.0f * 63.0f; }
    forEachPixel(p, [&](int x, int y) { rb::historyFixPixel(p, x, y); }); keep(tmp1Diff, tmp1Spec);
    forEachPixel(p, [&](int x, int y) { rb::spatialPixel<1>(p, x, y); }); keep(tmp2Diff, tmp2Spec);
    forEachPixel(p, [&](int x, int y) { rb::spatialPixel<2>(p, x, y); }); keep(h.diffHistory, h.specHistory);
    forEachPixel(p, [&](int x, int y) { rb::temporalStabilizationPixel(p, x, y); });
    h.pingPong = curr; h.valid = true;
    memcpy(outDiff, oDiff.data(), n * 8); memcpy(outSpec, oSpec.data(), n * 8);
    return 0;
}

// DenoiseSpecHitT: the product's pixel function (guides_filter.cuh), ping then pong as rtxpt_b200_denoise_spec_hit_t launches it
extern "C" int emu_denoise_spec_hit_t(uint32_t W, uint32_t H, const float* depth, float* specHitT)
{
    std::vector<float> scratch(size_t(W) * H);
    for (int y = 0; y < int(H); y++) for (int x = 0; x < int(W); x++) scratch[size_t(y) * W + x] = pt::specHitTNeighbourhood(specHitT, depth, int(W), int(H), x, y);
    for (int y = 0; y < int(H); y++) for (int x = 0; x < int(W); x++) specHitT[size_t(y) * W + x] = pt::specHitTNeighbourhood(scratch.data(), depth, int(W), int(H), x, y);
    return 0;
}

// EnvMapBaker: the product's BaseLayer / MIPReduce bodies (envbake.cuh) in launchEnvBake's order; lights: 8 floats each; out: all MIPs back to back
extern "C" int emu_bake_env_map(uint32_t cubeDim, uint32_t sourceType, uint32_t sourceWidth, uint32_t sourceHeight, const float* source, const float* scaleColor, uint32_t lightCount, const float* lights, float* out)
{
    pt::envbake::Params p{};
    p.cubeDim = cubeDim; p.sourceType = sourceType; p.sourceWidth = sourceWidth; p.sourceHeight = sourceHeight; p.source = source; memcpy(p.scaleColor, scaleColor, 12); p.lightCount = lightCount;
    for (uint32_t i = 0; i < lightCount; i++) { memcpy(p.lights[i].colorIntensity, lights + 8 * i, 16); memcpy(p.lights[i].direction, lights + 8 * i + 4, 12); p.lights[i].angularSize = lights[8 * i + 7]; }
    uint32_t levels = 0; while ((cubeDim >> levels) > 0) levels++;
    size_t off = 0; for (uint32_t m = 0; m < levels; m++) { p.mips[m] = out + off; off += size_t(6) * (cubeDim >> m) * (cubeDim >> m) * 4; }
    const uint32_t half = cubeDim / 2;
    for (uint32_t f = 0; f < 6; f++)
    {
        #pragma omp parallel for schedule(dynamic, 4)
        for (int y = 0; y < int(half); y++) for (uint32_t x = 0; x < half; x++) pt::envbake::baseLayerTexel(p, x, uint32_t(y), f, levels > 1);
    }
    for (uint32_t m = 2; m < levels; m++) { const uint32_t n = cubeDim >> m; for (uint32_t f = 0; f < 6; f++) for (uint32_t y = 0; y < n; y++) for (uint32_t x = 0; x < n; x++) pt::envbake::mipReduceTexel(p, m, x, y, f); }
    return 0;
}

// BVH refit: the product's bodies (refit.cuh) in launchRefit's order - every leaf triangle, then the levels deepest first; nodes / tris are updated in place, nodeBox is scratch
extern "C" int emu_refit(uint32_t* nodes, float* tris, const uint32_t* triShade, const RtxptInstanceData* instances, float* nodeBox, uint32_t nodeCount, uint32_t triCount, const uint32_t* levelStart, uint32_t levelCount)
{
    pt::refit::Params p{};
    p.nodes = reinterpret_cast<uint4*>(nodes); p.tris = reinterpret_cast<float4*>(tris); p.triShade = reinterpret_cast<const uint4*>(triShade); p.instances = instances; p.nodeBox = nodeBox; p.nodeCount = nodeCount; p.triCount = triCount;
    for (uint32_t i = 0; i < triCount; i++) pt::refit::refitTriangle(p, i);
    for (uint32_t d = levelCount; d-- > 0;) for (uint32_t ni = levelStart[d]; ni < levelStart[d + 1]; ni++) pt::refit::refitNode(p, ni);
    return 0;
}

// tone mapping: the product's host constants (makeParams / preExposedGray) and pixel bodies (tonemap.cuh); the log-luminance mean in double like the kernels' reduction
extern "C" int emu_tone_map(const RtxptToneMappingParams* u, const float* rgba, uint32_t pixelCount, uint8_t* outRGBA8, float* outAux)
{
    const pt::tonemap::Params p = pt::tonemap::makeParams(*u);
    double s = 0; for (uint32_t i = 0; i < pixelCount; i++) s += double(pt::tonemap::logLuminance(pt::mk3(rgba[4 * i], rgba[4 * i + 1], rgba[4 * i + 2])));
    const float avg = exp2f(float(s / double(pixelCount)));
    for (uint32_t i = 0; i < pixelCount; i++)
    {
        const uint32_t v = pt::tonemap::packLdr(pt::tonemap::applyToneMapping(p, p.autoExposure ? avg : 1.0f, pt::mk3(rgba[4 * i], rgba[4 * i + 1], rgba[4 * i + 2])), rgba[4 * i + 3]);
        memcpy(outRGBA8 + 4 * size_t(i), &v, 4);
    }
    if (outAux) { outAux[0] = avg; pt::tonemap::preExposedGray(*u, avg, outAux + 1); }
    return 0;
}

// RTXPT's side of the denoiser interface: the product's pixel bodies (denoiser_iface.cuh) over host copies of the realtime targets.  realtimeTargets = { planes, header, stableRadiance,
// depth, motion, throughput, specularHitT }, denoiserTargets = { viewZ, motion, normalRoughness, diff, spec, disocclusionMix, historyClampRelax, outputColor } - the layouts of
// oracle_denoiser_prepare_inputs.  mode 0: prepare inputs; 1: final merge with the given denoised images.
extern "C" int emu_denoiser_interface(const RtxptPathTracerConstants* consts, const RtxptRealtimeConstants* rt, const RtxptDenoiserConstants* k, uint32_t plane, int initWithStableRadiance, int mode,
                                      void* const* realtimeTargets, void* const* denoiserTargets, const void* denoisedDiff, const void* denoisedSpec)
{
    pt::LaunchParams p; memset(&p, 0, sizeof(p));
    p.c = *consts;
    const uint32_t W = consts->imageWidth, H = consts->imageHeight;
    p.rt.planes = static_cast<RtxptStablePlane*>(realtimeTargets[0]); p.rt.header = static_cast<uint32_t*>(realtimeTargets[1]); p.rt.stableRadiance = static_cast<uint2*>(realtimeTargets[2]);
    p.rt.specularHitT = static_cast<float*>(realtimeTargets[6]);
    p.rt.lineStride = rtxpt_b200_generic_ts_line_stride(W, H); p.rt.planeStride = rtxpt_b200_generic_ts_plane_stride(W, H); p.rt.activePlaneCount = rt->activeStablePlaneCount;
    p.rt.dnViewZ = static_cast<float*>(denoiserTargets[0]); p.rt.dnMotion = static_cast<uint2*>(denoiserTargets[1]); p.rt.dnNormalRoughness = static_cast<uint32_t*>(denoiserTargets[2]);
    p.rt.dnDiff = static_cast<uint2*>(denoiserTargets[3]); p.rt.dnSpec = static_cast<uint2*>(denoiserTargets[4]); p.rt.dnDisocclusionMix = static_cast<unsigned char*>(denoiserTargets[5]);
    p.rt.dnHistoryClampRelax = static_cast<unsigned char*>(denoiserTargets[6]); p.outputColor = static_cast<uint2*>(denoiserTargets[7]);
    p.rt.dnDenoisedDiff = static_cast<const uint2*>(denoisedDiff); p.rt.dnDenoisedSpec = static_cast<const uint2*>(denoisedSpec);
    if (k) p.rt.dn = *k;
    p.rt.dnPlane = plane; p.rt.dnInitWithStableRadiance = initWithStableRadiance ? 1u : 0u;
    for (uint32_t y = 0; y < H; y++) for (uint32_t x = 0; x < W; x++) { const uint32_t id = (x << 16) | y; if (mode == 0) pt::dnPrepareInputsPixel(p, id); else pt::dnFinalMergePixel(p, id); }
    return 0;
}

// pin against tests/golden/host_golden.json: the product's white balance transform (tonemap.cuh host half)
extern "C" void emu_white_balance(float T, float* outM9, float* outXyz3)
{ const pt::tonemap::M3 m = pt::tonemap::whiteBalanceTransform(T); for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) outM9[i * 3 + j] = m.m[i][j]; pt::tonemap::colorTemperatureToXYZ(T, outXyz3); }

// skinning: the product's bodies (skinning.cuh) in launchSkin's order; same argument list as oracle_skin
extern "C" int emu_skin(uint32_t numVertices, uint32_t numTriangles, uint32_t firstGid, const float* positions, const uint32_t* normals, const uint32_t* tangents, const uint16_t* jointIndices,
                        const float* jointWeights, const float* jointMatrices, const uint32_t* indices, float* outPositions, uint32_t* outNormals, uint32_t* outTangents, uint32_t* triShade)
{
    pt::skin::Params p{};
    p.numVertices = numVertices; p.numTriangles = triShade ? numTriangles : 0; p.firstGid = firstGid; p.flags = (normals ? 2u : 0u) | (tangents ? 4u : 0u);
    p.positions = positions; p.normals = normals; p.tangents = tangents; p.jointIndices = jointIndices; p.jointWeights = jointWeights; p.jointMatrices = jointMatrices;
    p.outPositions = outPositions; p.outNormals = outNormals; p.outTangents = outTangents; p.indices = indices; p.triShade = reinterpret_cast<uint4*>(triShade);
    for (uint32_t i = 0; i < p.numVertices; i++) pt::skin::skinVertex(p, i);
    for (uint32_t t = 0; t < p.numTriangles; t++) pt::skin::gatherTriangle(p, t);
    return 0;
}
