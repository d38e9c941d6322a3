// TEST INFRASTRUCTURE ONLY - never linked into librtxpt_b200*.so.  Host build of the product's NEE-AT feedback passes (rtxpt_b200/csrc/neeat.cuh: the __host__ __device__ bodies
// the kernels of neeat_kernels.cu wrap) with the product's host bookkeeping (neeat_host.h), run in the order rtxpt_b200_neeat_update_begin / _end launch them.  Lets
// tests/test_neeat_port.py hold the CUDA source to the oracle (oracle/pt_neeat.h) on the CPU; the device-only parts (atomics of P0, the scan, launch shapes) stay with the GPU tests.
#include "../../rtxpt_b200/csrc/neeat_host.h"
#include <vector>
#include <cstring>
#include <cstdint>

using namespace pt;

namespace {
struct Instance
{
    neeat::HostState host; neeat::Params p{};
    uint32_t neeType = 2, lightCount = 0; float weightsSum = 0;
    std::vector<float> fbW, scW, blW, historyDepth, weights, bw[2], groupSums; float weightsSumDev = 0; uint32_t pingPong = 0, boostFlags = 0; std::vector<uint32_t> lightRecords; float worldToClip[16]; bool haveView = false; std::vector<uint32_t> fbC, scC, blC, local, counters, proxyCounters, proxyOffsets, proxyIndices; uint32_t samplingProxyCount = 0;
    void bind()
    {
        p.fbWeight = fbW.data(); p.fbCandidate = fbC.data(); p.scratchWeight = scW.data(); p.scratchCandidate = scC.data(); p.blendedWeight = blW.data(); p.blendedCandidate = blC.data();
        p.historyDepth = historyDepth.data(); p.localSamplingBuffer = local.data(); p.feedbackCounters = counters.data(); p.lightWeights = weights.data(); p.proxyCounters = proxyCounters.data();
        p.proxyOffsets = proxyOffsets.data(); p.proxyIndices = proxyIndices.data(); p.samplingProxyCount = &samplingProxyCount;
        p.lightRecords = reinterpret_cast<const uint4*>(lightRecords.data()); p.curWeights = bw[pingPong].data(); p.histWeights = bw[pingPong ^ 1u].data(); p.weightGroupSums = groupSums.data(); p.weightsSumDev = &weightsSumDev;
    }
};
}

extern "C" void* neeat_emu_create(uint32_t W, uint32_t H, uint32_t lightCount, const float* weights, float weightsSum, uint32_t neeType)
{
    Instance* i = new Instance(); i->host.reset(W, H); i->neeType = neeType; i->lightCount = lightCount; i->weightsSum = weightsSum;
    const size_t P = size_t(W) * H, B = size_t((W + 1) / 2) * ((H + 1) / 2), T = size_t(neeat::HostState::tilesX(W)) * neeat::HostState::tilesY(H) * neeat::kLocalProxyCount;
    i->fbW.assign(P, 0.f); i->scW.assign(P, 0.f); i->blW.assign(B, 0.f); i->historyDepth.assign(P, 0.f); i->fbC.assign(P, neeat::kInvalidLight); i->scC.assign(P, neeat::kInvalidLight); i->blC.assign(B, neeat::kInvalidLight);
    i->local.assign(T, 0u); i->counters.assign(size_t(lightCount) + 1, 0u); i->weights.assign(weights, weights + lightCount); i->proxyCounters.assign(lightCount, 0u); i->proxyOffsets.assign(size_t(lightCount) + 1, 0u);
    const uint32_t budget = neeat::kProxyRatio * std::max(lightCount, neeat::kMaxLights / 10);
    i->proxyIndices.assign(size_t(budget) + lightCount, 0u);
    i->bw[0].assign(lightCount, 0.f); i->bw[1].assign(lightCount, 0.f); i->groupSums.assign((lightCount + 4095) / 4096 + 1, 0.f); i->lightRecords.assign(size_t(lightCount) * 8, 0u);
    return i;
}
// importance boosters: the light records (32 B each, for the centres and types), the flags of RtxptPathTracerConstants::NEEATImportanceBoost and view.matWorldToClip
extern "C" int neeat_emu_set_boost(void* h, uint32_t flags, const uint32_t* lightRecords, const float* worldToClip)
{
    Instance& i = *static_cast<Instance*>(h); i.boostFlags = flags;
    if (lightRecords) memcpy(i.lightRecords.data(), lightRecords, i.lightRecords.size() * 4);
    i.haveView = worldToClip != nullptr; if (worldToClip) memcpy(i.worldToClip, worldToClip, 64);
    return 0;
}
extern "C" void neeat_emu_destroy(void* h) { delete static_cast<Instance*>(h); }
extern "C" int neeat_emu_set_feedback(void* h, const float* weight, const uint32_t* candidate)
{
    Instance& i = *static_cast<Instance*>(h);
    memcpy(i.fbW.data(), weight, i.fbW.size() * 4); memcpy(i.fbC.data(), candidate, i.fbC.size() * 4); i.host.feedbackBufferFilled = true;
    return 0;
}
// rtxpt_b200_neeat_update_begin's order: [snapshot -> PreFilter -> P0] -> proxy counts -> exclusive scan -> proxy fill
extern "C" int neeat_emu_update_begin(void* h)
{
    Instance& i = *static_cast<Instance*>(h); neeat::Params& p = i.p;
    neeat::beginFrame(i.host, p, i.neeType, i.lightCount, i.weightsSum, i.boostFlags, i.haveView ? i.worldToClip : nullptr); i.pingPong ^= 1u; i.bind();
    std::fill(i.counters.begin(), i.counters.end(), 0u);
    if (p.lastFrameFeedbackAvailable)
    {
        if (i.host.settings.preFilter)
        {
            i.scW = i.fbW; i.scC = i.fbC;
            for (int y = 0; y < int(p.H); y++) for (int x = 0; x < int(p.W); x++) neeat::preFilterPixel(p, x, y);
        }
        for (int y = 0; y < int(p.H); y++) for (int x = 0; x < int(p.W); x++) i.counters[neeat::p0Pixel(p, x, y)]++;
    }
    {   // k_na_weights / k_na_weight_total: 32-light blocks, 128-block groups, groups in index order
        const uint32_t blocks = (p.lightCount + 31) / 32, groups = (p.lightCount + 4095) / 4096; float total = 0.0f;
        for (uint32_t g = 0; g < groups; g++) { float gs = 0.0f; for (uint32_t b = g * 128; b < std::min(blocks, g * 128 + 128); b++) gs = gs + neeat::weightBlock(p, b); i.groupSums[g] = gs; total = total + gs; }
        i.weightsSumDev = total;
    }
    for (uint32_t l = 0; l < p.lightCount; l++) i.proxyCounters[l] = neeat::proxyCountOfLight(p, l);
    uint32_t total = 0; for (uint32_t l = 0; l < p.lightCount; l++) { i.proxyOffsets[l] = total; total += i.proxyCounters[l]; } i.proxyOffsets[p.lightCount] = total;
    i.samplingProxyCount = total;
    for (uint32_t slot = 0; slot < total; slot++) i.proxyIndices[slot] = neeat::lightOfProxySlot(p, slot);
    return 0;
}
// rtxpt_b200_neeat_update_end's order: P1a -> P1b -> P2 + P3 per tile -> ClearFeedbackHistory
extern "C" int neeat_emu_update_end(void* h, const float* depth, const uint16_t* motion)
{
    Instance& i = *static_cast<Instance*>(h); neeat::Params& p = i.p; i.bind();
    p.depth = depth; p.motion = reinterpret_cast<const uint2*>(motion);
    for (uint32_t y = 0; y < p.blendedH; y++) for (uint32_t x = 0; x < p.blendedW; x++) neeat::p1aPixel(p, x, y);
    for (uint32_t y = 0; y < p.H; y++) for (uint32_t x = 0; x < p.W; x++) neeat::p1bPixel(p, x, y);
    for (uint32_t ty = 0; ty < p.tilesY; ty++) for (uint32_t tx = 0; tx < p.tilesX; tx++)
    {
        uint32_t data[neeat::kLocalProxyCount];
        for (uint32_t slot = 0; slot < neeat::kLocalProxyCount; slot++) data[slot] = neeat::fillTileEntry(p, tx, ty, slot);
        for (uint32_t k = 2; k <= neeat::kLocalProxyCount; k <<= 1) for (uint32_t j = k / 2; j > 0; j /= 2) for (uint32_t t = 0; t < neeat::kLocalProxyCount / 2; t++) neeat::bitonicStep(data, t, k, j);
        const uint32_t base = neeat::tileBaseAddress(p, tx, ty);
        for (uint32_t loc = 0; loc < neeat::kLocalProxyCount; loc++) i.local[base + loc] = neeat::packMiniList(data[loc], neeat::runLength(data, loc));
    }
    if (p.temporalFeedbackRequired) for (uint32_t y = 0; y < p.H; y++) for (uint32_t x = 0; x < p.W; x++) neeat::clearFeedbackPixel(p, x, y);
    neeat::endFrame(i.host, p);
    return 0;
}
// same `what` codes as oracle_neeat_get
extern "C" int neeat_emu_get(void* h, int what, void* out, size_t bytes)
{
    Instance& i = *static_cast<Instance*>(h); const void* src = nullptr; size_t n = 0; uint32_t ctl[8];
    switch (what)
    {
    case 0: src = i.fbW.data(); n = i.fbW.size() * 4; break;   case 1: src = i.fbC.data(); n = i.fbC.size() * 4; break;
    case 2: src = i.scW.data(); n = i.scW.size() * 4; break;   case 3: src = i.scC.data(); n = i.scC.size() * 4; break;
    case 4: src = i.blW.data(); n = i.blW.size() * 4; break;   case 5: src = i.blC.data(); n = i.blC.size() * 4; break;
    case 6: src = i.local.data(); n = i.local.size() * 4; break; case 7: src = i.proxyCounters.data(); n = i.proxyCounters.size() * 4; break;
    case 8: ctl[0] = i.p.tilesX; ctl[1] = i.p.tilesY; ctl[2] = i.p.jitterX; ctl[3] = i.p.jitterY; ctl[4] = i.samplingProxyCount; ctl[5] = i.p.updateCounter; ctl[6] = i.p.lastFrameFeedbackAvailable;
            ctl[7] = i.p.lastFrameFeedbackAvailable ? i.p.W * i.p.H - i.counters[i.p.lightCount] : 0u; src = ctl; n = sizeof(ctl); break;
    case 11: src = i.proxyIndices.data(); n = size_t(i.samplingProxyCount) * 4; break;
    case 13: src = i.bw[i.pingPong].data(); n = i.bw[i.pingPong].size() * 4; break;
    case 14: src = &i.weightsSumDev; n = 4; break;
    default: return -1;
    }
    if (bytes < n) return -2;
    memcpy(out, src, n);
    return int(n);
}

// sampler-side functions (what the NEEAT shade kernel calls) for pixel (px, py): out = { sampleLocal's light, its pdf, sampleLocalPdf( lightForPdf ) }
extern "C" int neeat_emu_sample_local(void* h, uint32_t px, uint32_t py, float rnd, uint32_t lightForPdf, float* out)
{
    Instance& i = *static_cast<Instance*>(h); i.bind();
    const uint32_t tile = neeat::localSamplingTilePos(i.p, px, py); float pdf = 0;
    const uint32_t light = neeat::sampleLocal(i.p, tile, rnd, pdf);
    out[0] = float(light); out[1] = pdf; out[2] = neeat::sampleLocalPdf(i.p, tile, lightForPdf);
    return 0;
}

// pins against tests/golden/host_golden.json: the product's MicroRng and candidate counts
extern "C" void emu_micro_rng(uint32_t x, uint32_t y, uint32_t a, uint32_t b, uint32_t n, uint32_t* outNext, float* outFloats)
{ neeat::MicroRng r = neeat::MicroRng::make(x, y, a, b); for (uint32_t i = 0; i < n; i++) outNext[i] = r.next(); for (uint32_t i = 0; i < n; i++) outFloats[i] = r.nextFloat(); }
extern "C" uint32_t emu_candidate_local_count(float ratio, uint32_t total) { return neeat::candidateLocalCount(ratio, total); }
