// ORACLE/_ref — TEST INFRASTRUCTURE: known-answer generator for NEE-AT's feedback passes, built from the UNMODIFIED Rtxpt/Lighting/LightsBaker.hlsl (oracle/ref_baker_tu.sh pipes
// the shim, the lighting headers, that file's passes and this file into g++; nothing of the reference is written to disk).  The passes run here are the ones whose threads do not
// talk to each other - ProcessFeedbackHistoryP0, P1a, P1b, P2 (FillTile), ClearFeedbackHistory - one thread after the other, in LightsBaker::UpdateEnd's order
// (LightsBaker.cpp:1331-1418); barriers are no-ops, a wave is one lane (P0's WaveMatch counting degenerates to one atomic add per pixel: same counters).  Not run: PreFilter and
// P3 (group-shared tiles, bitonic sort across a thread group), the proxy-table passes.
//   usage: ref_kat_baker feedback|counts in.f32 out.f32
// record (3056 floats; words marked * are bit patterns): 0-31 header | 32-47* past-to-current table (16) | 48-111 global proxies (64) | 112-367 feedback total weight (16 x 16),
// 368-623* feedback candidates | 624-879 history depth | 880-1135 depth | 1136-1903 motion vectors (xyz per pixel) | 1904-3055* last frame's local sampling buffer (3 x 3 tiles x 128)
// out (3089 floats): feedback weight / candidates after P0 (512), counters (17) | blended after P1a (64 + 64) | scratch after P1b (512) | tile lists after P2, unsorted (1152) |
// feedback after ClearFeedbackHistory (512), history depth (256) | tile lists after P3 (1152): ProcessFeedbackHistoryP3 synchronises its 64 threads (bitonic sort in group-shared
// memory), so every tile runs on 64 real threads with a barrier behind GroupMemoryBarrierWithGroupSync
#include <cstdio>
#include <vector>
#include <thread>
static std::vector<float> readAll(const char* path) { std::vector<float> v; FILE* f = fopen(path, "rb"); if (!f) { perror(path); exit(1); } fseek(f, 0, SEEK_END); v.resize(size_t(ftell(f)) / 4); fseek(f, 0, SEEK_SET); if (fread(v.data(), 4, v.size(), f) != v.size()) exit(1); fclose(f); return v; }
float4 (*g_shimCubeSample)(float3 dir, float lod) = nullptr;


int main(int argc, char** argv)
{
    if (argc != 4 || (std::string(argv[1]) != "feedback" && std::string(argv[1]) != "counts")) { fprintf(stderr, "usage: %s feedback|counts in.f32 out.f32\n", argv[0]); return 2; }
    if (std::string(argv[1]) == "counts")
    {   // ComputeProxyCounts (LightsBaker.hlsl:880-948, UpdateBegin): how many global sampling proxies every light gets from its power-based weight blended with last frame's usage
        // counters.  One thread per light in a group of 128 that synchronises once (thread 0 then sums the group): real threads, as for P3.
        // record (64 floats): 0 light count (<= 16), 1 LastFrameTemporalFeedbackAvailable, 2 TotalMaxFeedbackCount, 3 GlobalFeedbackUseWeight, 4 ImportanceSamplingType, 5 weights sum,
        // 8-23 weights, 24-40 usage counters (the last: reservoirs without a light)   out (40 floats): 0-15 proxy counters, 16 SamplingProxyCount, 17-32 offsets inside the group, 33 group total
        const std::vector<float> in = readAll(argv[2]); const size_t n = in.size() / 64; std::vector<float> out(n * 40, 0.0f);
        for (size_t i = 0; i < n; i++)
        {
            const float* r = &in[i * 64]; float* o = &out[i * 40];
            LightingControlData cd; memset(&cd, 0, sizeof(cd));
            const uint count = uint(r[0]); cd.TotalLightCount = count; cd.LastFrameTemporalFeedbackAvailable = uint(r[1]); cd.TotalMaxFeedbackCount = uint(r[2]); cd.GlobalFeedbackUseWeight = r[3];
            cd.ImportanceSamplingType = uint(r[4]); memcpy(&cd.WeightsSumUINT, r + 5, 4);
            std::vector<float> weights(r + 8, r + 24); std::vector<uint> counters(32, 0u), scratchList(32, 0u), proxies(64, 0u);
            for (int k = 0; k < 17; k++) counters[k] = uint(r[24 + k]);
            u_controlBuffer.p = &cd; u_controlBuffer.n = 1; u_lightWeights.p = weights.data(); u_lightWeights.n = 16; u_perLightProxyCounters.p = counters.data(); u_perLightProxyCounters.n = 32;
            u_scratchList.p = scratchList.data(); u_scratchList.n = 32; u_lightSamplingProxies.p = proxies.data(); u_lightSamplingProxies.n = 64;
            pthread_barrier_t barrier; pthread_barrier_init(&barrier, nullptr, count); g_shimGroupBarrier = &barrier;       // threads past the light count leave before the barrier
            std::vector<std::thread> threads;
            for (uint t = 0; t < 128; t++) threads.emplace_back([=]() { ComputeProxyCounts(t, t); });
            for (auto& th : threads) th.join();
            g_shimGroupBarrier = nullptr; pthread_barrier_destroy(&barrier);
            for (int k = 0; k < 16; k++) { o[k] = float(counters[k]); o[17 + k] = float(scratchList[k]); }
            o[16] = float(cd.SamplingProxyCount); o[33] = float(proxies[1]);
        }
        FILE* f = fopen(argv[3], "wb"); if (!f) { perror(argv[3]); return 1; } fwrite(out.data(), 4, out.size(), f); fclose(f);
        return 0;
    }
    const std::vector<float> in = readAll(argv[2]); const int kIn = 3056, kOut = 4241; const size_t n = in.size() / kIn; std::vector<float> out(n * kOut, 0.0f);
    const uint W = 16, H = 16, P = W * H, LW = 8, LH = 8, TX = 3, TY = 3;
    for (size_t i = 0; i < n; i++)
    {
        const float* r = &in[i * kIn]; float* o = &out[i * kOut];
        LightingControlData cd; memset(&cd, 0, sizeof(cd));
        cd.TotalLightCount = uint(r[0]); cd.HistoricTotalLightCount = uint(r[1]); cd.BakerConstants.UpdateCounter = uint(r[2]); cd.LocalSamplingTileJitter = uint2((uint)r[3], (uint)r[4]);
        cd.LocalSamplingTileJitterPrev = uint2((uint)r[5], (uint)r[6]); cd.LastFrameTemporalFeedbackAvailable = uint(r[7]); cd.LastFrameLocalSamplesAvailable = uint(r[8]); cd.SamplingProxyCount = uint(r[9]);
        cd.BakerConstants.DepthDisocclusionThreshold = r[10]; cd.BakerConstants.EnableMotionReprojection = uint(r[11]); cd.BakerConstants.ReservoirHistoryDropoff = r[12];
        cd.BakerConstants.PrevOverCurrentViewportSize = float2(1.0f, 1.0f); cd.BakerConstants.FeedbackResolution = uint2(W, H); cd.BakerConstants.BlendedFeedbackResolution = uint2(LW, LH);
        cd.LocalSamplingResolution = uint2(TX, TY); cd.BakerConstants.MouseCursorPos = uint2(0xFFFFu, 0xFFFFu); cd.BakerConstants.DebugDrawType = 0;
        std::vector<uint> remap(32, 0xFFFFFFFFu), proxies(64), counters(32, 0u), fbC(P), scC(P, 0xFFFFFFFFu), blC(LW * LH, 0xFFFFFFFFu), local(TX * TY * 128);
        std::vector<float> fbW(P), scW(P, 0.0f), blW(LW * LH, 0.0f), histDepth(P), depth(P); std::vector<float3> motion(P);
        memcpy(remap.data(), r + 32, 64); for (int k = 0; k < 64; k++) proxies[k] = uint(r[48 + k]);
        memcpy(fbW.data(), r + 112, P * 4); memcpy(fbC.data(), r + 368, P * 4); memcpy(histDepth.data(), r + 624, P * 4); memcpy(depth.data(), r + 880, P * 4);
        for (uint k = 0; k < P; k++) motion[k] = float3(r[1136 + 3 * k], r[1137 + 3 * k], r[1138 + 3 * k]);
        memcpy(local.data(), r + 1904, local.size() * 4);
        u_controlBuffer.p = &cd; u_controlBuffer.n = 1; u_historyRemapPastToCurrent.p = remap.data(); u_historyRemapPastToCurrent.n = 32; u_lightSamplingProxies.p = proxies.data(); u_lightSamplingProxies.n = 64;
        u_perLightProxyCounters.p = counters.data(); u_perLightProxyCounters.n = 32;
        u_feedbackTotalWeight.p = fbW.data(); u_feedbackTotalWeight.w = W; u_feedbackTotalWeight.h = H; u_feedbackCandidates.p = fbC.data(); u_feedbackCandidates.w = W; u_feedbackCandidates.h = H;
        u_feedbackTotalWeightScratch.p = scW.data(); u_feedbackTotalWeightScratch.w = W; u_feedbackTotalWeightScratch.h = H; u_feedbackCandidatesScratch.p = scC.data(); u_feedbackCandidatesScratch.w = W; u_feedbackCandidatesScratch.h = H;
        u_feedbackTotalWeightBlended.p = blW.data(); u_feedbackTotalWeightBlended.w = LW; u_feedbackTotalWeightBlended.h = LH; u_feedbackCandidatesBlended.p = blC.data(); u_feedbackCandidatesBlended.w = LW; u_feedbackCandidatesBlended.h = LH;
        u_historyDepth.p = histDepth.data(); u_historyDepth.w = W; u_historyDepth.h = H; t_depthBuffer.p = depth.data(); t_depthBuffer.w = W; t_depthBuffer.h = H; t_motionVectors.p = motion.data(); t_motionVectors.w = W; t_motionVectors.h = H;
        u_localSamplingBuffer.p = local.data(); u_localSamplingBuffer.n = uint(local.size());
        // P0 over the padded dispatch of 16 x 16 threads per group: the image is exactly one group here
        for (uint y = 0; y < H; y++) for (uint x = 0; x < W; x++) ProcessFeedbackHistoryP0(uint2(x, y));
        memcpy(o, fbW.data(), P * 4); memcpy(o + 256, fbC.data(), P * 4); for (int k = 0; k < 17; k++) o[512 + k] = float(counters[k]);
        for (uint y = 0; y < LH; y++) for (uint x = 0; x < LW; x++) ProcessFeedbackHistoryP1a(uint2(x, y));
        memcpy(o + 529, blW.data(), 64 * 4); memcpy(o + 593, blC.data(), 64 * 4);
        for (uint y = 0; y < H; y++) for (uint x = 0; x < W; x++) ProcessFeedbackHistoryP1b(uint2(x, y));
        memcpy(o + 657, scW.data(), P * 4); memcpy(o + 913, scC.data(), P * 4);
        for (uint y = 0; y < TY; y++) for (uint x = 0; x < TX; x++) ProcessFeedbackHistoryP2(uint2(x, y));
        memcpy(o + 1169, local.data(), 1152 * 4);
        for (uint y = 0; y < H; y++) for (uint x = 0; x < W; x++) ClearFeedbackHistory(uint2(x, y));
        memcpy(o + 2321, fbW.data(), P * 4); memcpy(o + 2577, fbC.data(), P * 4); memcpy(o + 2833, histDepth.data(), P * 4);
        for (uint ty = 0; ty < TY; ty++) for (uint tx = 0; tx < TX; tx++)
        {
            pthread_barrier_t barrier; pthread_barrier_init(&barrier, nullptr, 64); g_shimGroupBarrier = &barrier;
            std::vector<std::thread> threads;
            for (uint t = 0; t < 64; t++) threads.emplace_back([=]() { ProcessFeedbackHistoryP3(uint3(tx, ty, 0u), uint3(t, 0u, 0u)); });
            for (auto& th : threads) th.join();
            g_shimGroupBarrier = nullptr; pthread_barrier_destroy(&barrier);
        }
        memcpy(o + 3089, local.data(), 1152 * 4);
    }
    FILE* f = fopen(argv[3], "wb"); if (!f) { perror(argv[3]); return 1; } fwrite(out.data(), 4, out.size(), f); fclose(f);
    return 0;
}
