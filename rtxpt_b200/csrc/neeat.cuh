// neeat.cuh - NEE-AT temporal feedback: parameter block, the sampler-side functions the path tracer calls (local tile sampler, its pdf, feedback insertion)
// and the bodies of LightsBaker's feedback passes as __host__ __device__ functions (neeat_kernels.cu wraps them in kernels; the test-only host build tests/emu runs the same source
// on the CPU against the oracle).  Restates
//   Rtxpt/Shaders/PathTracer/Lighting/LightingTypes.hlsli:146-177 (candidate counts, mini-list tuples, tile address), :180-291 (LightFeedbackReservoir)
//   Rtxpt/Shaders/PathTracer/Lighting/LightSampler.hlsli:45-96, :120-199 (tile position, SampleLocal, SampleLocalPDF, InsertFeedbackFromNEE)
//   Rtxpt/Shaders/PathTracer/Lighting/LightingAlgorithms.hlsli:654-682 (binary search), Rtxpt/Shaders/Libraries/MicroRng.hlsli:12-60
//   Rtxpt/Lighting/LightsBaker.hlsl:774-823 (ClearFeedbackHistory), :880-948 (ComputeProxyCounts), :1068-1093, :1095-1181 (PreFilter), :1186-1318 (P0), :1321-1377, :1380-1452 (P1a),
//     :1456-1528 (P1b), :1531-1610 (FillTile), :1745-1850 (P3: bitonic sort, duplicate ranges)
// Light lists may change between frames (rtxpt_b200_update_lights): Params::pastToCurrent / currentToPast carry the reference's index tables for the one frame that needs them;
// NULL = unchanged list, the remap is the identity with its bounds checks.  PreFilter reads a snapshot of the reservoirs (the reference's in-place
// version is racy across thread groups).
#pragma once
#include "device_math.cuh"

namespace pt { namespace neeat {

constexpr uint kTileSize = 8, kWindowSize = 8, kLocalProxyCount = 128, kBinarySearchSteps = 8, kEarlyFeedbackTileSize = 2, kTopUpSamples = kLocalProxyCount - kWindowSize * kWindowSize;
constexpr uint kInvalidLight = 0xFFFFFFFFu, kSSCFlag = 0x80000000u;
constexpr float kMaxWeight = 1e12f;
constexpr uint kProxyRatio = 12, kMaxLights = 512 * 1024, kMaxProxiesPerLight = 256 * 1024;

struct Params
{
    uint W, H, blendedW, blendedH, tilesX, tilesY;
    uint lightCount, historicLightCount, updateCounter, neeType;
    uint jitterX, jitterY, jitterPrevX, jitterPrevY;
    uint lastFrameFeedbackAvailable, lastFrameLocalSamplesAvailable, temporalFeedbackRequired, enableMotionReprojection;
    float reservoirHistoryDropoff, depthDisocclusionThreshold, globalFeedbackUseWeight, localToGlobalSampleRatio, screenSpaceVsWorldSpaceThreshold, weightsSum;
    // feedback reservoirs: the ones NEE fills, the processed copy (also PreFilter's snapshot), the half-resolution blend; depth of the frame they belong to
    float* fbWeight; uint* fbCandidate; float* scratchWeight; uint* scratchCandidate; float* blendedWeight; uint* blendedCandidate; float* historyDepth;
    uint* localSamplingBuffer;          // tilesX * tilesY * 128 (light << 9 | count - 1), sorted by light inside a tile
    uint* feedbackCounters;             // [lightCount + 1]: reservoirs per light; last = reservoirs without a light
    // importance boosters (LightsBaker.hlsl:107-165): bit 0 of boostFlags frustum, bit 1 intensity delta
    uint boostFlags; float boostFrustumMul, boostFrustumFadeDistance, boostIntensityDeltaMul; float frustumPlanes[5][4];
    const uint4* lightRecords;          // PolymorphicLightInfo, 2 x uint4 each: [2 i] = centre.xyz, colorTypeAndFlags
    float* curWeights; const float* histWeights; float* weightGroupSums; float* weightsSumDev;      // boosted weights of this / the last frame, partial sums, their total
    // global proxy table
    const float* lightWeights; uint* proxyCounters; uint* proxyOffsets; uint* proxyIndices; uint* samplingProxyCount;
    // guides of the frame: depth R32F, screen motion RGBA16F (pixels)
    const float* depth; const uint2* motion;
    // dynamic light lists (LightsBaker.hlsl: u_historyRemapPastToCurrent / u_historyRemapCurrentToPast): index of last frame's light in this frame's list and back,
    // kInvalidLight where a light has no counterpart; NULL = the list did not change (identity)
    const uint* pastToCurrent; const uint* currentToPast;
};

// single IEEE operations the build flags cannot fuse or approximate (-fmad / -prec-div / -use_fast_math): the baker passes are bit-exact against the oracle
#ifdef __CUDA_ARCH__
PT_HD float fadd_rn(float a, float b) { return __fadd_rn(a, b); }
PT_HD float fsub_rn(float a, float b) { return __fsub_rn(a, b); }
PT_HD float fmul_rn(float a, float b) { return __fmul_rn(a, b); }
PT_HD float fdiv_rn(float a, float b) { return __fdiv_rn(a, b); }
#else
PT_HD float fadd_rn(float a, float b) { return a + b; }
PT_HD float fsub_rn(float a, float b) { return a - b; }
PT_HD float fmul_rn(float a, float b) { return a * b; }
PT_HD float fdiv_rn(float a, float b) { return a / b; }
#endif

struct MicroRng
{
    uint N;
    PT_HD static MicroRng make(uint x, uint y, uint a, uint b) { MicroRng r; r.N = ((x << 16) | y) ^ 0x9e3779b9u; r.N = r.N ^ (a + (r.N << 6) + (r.N >> 2)); r.N = r.N ^ (b + (r.N << 6) + (r.N >> 2)); return r; }
    PT_HD uint next() { N ^= N >> 16; N *= 0x21f0aaadu; N ^= N >> 15; N *= 0xf35a2d97u; N ^= N >> 15; return N; }
    PT_HD float nextFloat() { return float(next() >> 8) * 5.9604644775390625e-8f; }       // / 2^24, exact
};

PT_HD uint packMiniList(uint lightIndex, uint counter) { return ((lightIndex & 0x007FFFFFu) << 9) | ((counter - 1) & 0x1FFu); }
PT_HD uint miniListLight(uint v) { return v >> 9; }
PT_HD uint miniListCount(uint v) { return (v & 0x1FFu) + 1; }
PT_HD uint candidateLocalCount(float localToGlobalRatio, uint totalCandidateSamples) { return uint(float(totalCandidateSamples - 1) * localToGlobalRatio + 0.75f); }
PT_HD uint tileBaseAddress(const Params& p, uint tx, uint ty) { return (tx + ty * p.tilesX) * kLocalProxyCount; }
PT_HD int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
PT_HD uint minu(uint a, uint b) { return a < b ? a : b; }

// ---- reservoirs (LightFeedbackReservoir) over a weight / candidate image pair -----------------------------------------------------------------------------------------
struct Reservoir
{
    float* w; uint* c;
    PT_HD static Reservoir at(float* weight, uint* candidate, size_t i) { Reservoir r; r.w = weight + i; r.c = candidate + i; return r; }
    PT_HD float total() const { return *w; }
    PT_HD void setTotal(float v) { *w = fminf(kMaxWeight, v); }
    PT_HD bool empty() const { return *w == 0.0f; }
    PT_HD void clear() { *w = 0.0f; *c = kInvalidLight; }
    PT_HD void cloneFrom(const Reservoir& o, float scale) { if (o.total() > 0.0f) { setTotal(fmul_rn(o.total(), scale)); *c = *o.c; } else clear(); }
    PT_HD void add(float rnd, uint candidateIndex, float candidateWeight, bool ssc)
    {
        candidateWeight = fminf(kMaxWeight, candidateWeight);
        const float t = fadd_rn(total(), candidateWeight); setTotal(t);
        const float threshold = sat(fdiv_rn(candidateWeight, t));
        if (ssc) candidateIndex |= kSSCFlag;
        if (rnd < threshold) *c = candidateIndex;
    }
    PT_HD void merge(float rnd, const Reservoir& o, float otherScale)
    {
        const float otherTotal = fminf(kMaxWeight, fmul_rn(o.total(), otherScale));
        if (otherTotal > 0.0f) { const uint l = *o.c; if (l != kInvalidLight) add(rnd, l & ~kSSCFlag, otherTotal, (l & kSSCFlag) != 0); }
    }
};

// ---- sampler side ----------------------------------------------------------------------------------------------------------------------------------------------------
PT_HD uint localSamplingTilePos(const Params& p, uint px, uint py) { return tileBaseAddress(p, (px + p.jitterX) / kTileSize, (py + p.jitterY) / kTileSize); }
PT_HD uint sampleLocal(const Params& p, uint tileAddress, float rnd, float& pdf)
{
    const uint v = p.localSamplingBuffer[tileAddress + minu(uint(rnd * float(kLocalProxyCount)), kLocalProxyCount - 1)];
    pdf = float(miniListCount(v)) / float(kLocalProxyCount);
    return miniListLight(v);
}
// LocalLightBinarySearch as the reference runs it: exactly 8 steps, no empty-range test.  A light below every key of the tile makes step 8 read the word just before the tile -
// the previous tile's last entry (whose count comes back if it holds the light) or, for tile 0, address 0x7FFFFFFF, which a D3D typed buffer reads as 0 ("light 0, count 1");
// the bounds test below is that D3D read.  Pinned against the reference header by tests/golden/sampler_golden.npz (oracle side)
PT_HD float sampleLocalPdf(const Params& p, uint tileAddress, uint lightIndex)
{
    const uint total = p.tilesX * p.tilesY * kLocalProxyCount;
    uint left = tileAddress, right = tileAddress + kLocalProxyCount - 1;
    #pragma unroll
    for (uint i = 0; i < kBinarySearchSteps; i++)
    {
        const uint mid = (left + right) >> 1; const uint v = mid < total ? p.localSamplingBuffer[mid] : 0u, key = miniListLight(v);
        if (key < lightIndex) left = mid + 1;
        else if (key > lightIndex) right = mid - 1;
        else return float(miniListCount(v)) / float(kLocalProxyCount);
    }
    return 0.0f;
}
// feedbackWeight: how much this pixel wanted this light (throughput x BSDF x light); globalPdf = SampleGlobalPDF( lightIndex ); the caller supplies pow( globalPdf, 0.65 )'s input
PT_HD void insertFeedbackFromNEE(const Params& p, uint px, uint py, bool ssc, uint lightIndex, float pixelRadianceContributionAvg, float globalPdf, float rnd)
{
    const float w = fdiv_rn(pixelRadianceContributionAvg, powf(globalPdf, 0.65f));
    Reservoir::at(p.fbWeight, p.fbCandidate, size_t(py) * p.W + px).add(rnd, lightIndex, w, ssc);
}

// ---- baker passes ----------------------------------------------------------------------------------------------------------------------------------------------------------
PT_HD uint remapPastToCurrent(const Params& p, uint historic)
{
    if (historic == kInvalidLight) return kInvalidLight;
    uint idx = historic < p.historicLightCount ? (p.pastToCurrent ? p.pastToCurrent[historic] : historic) : kInvalidLight;
    if (idx != kInvalidLight && idx >= p.lightCount) idx = kInvalidLight;
    return idx;
}

// PreFilter: every reservoir is replaced by a weighted pick from its 3x3 neighbourhood (centre x48, same coherence class x128); reads the snapshot in scratch*, writes fb*
PT_HD void preFilterPixel(const Params& p, int x, int y)
{
    const int W = int(p.W), H = int(p.H);
    uint kIndex[9]; float kWeight[9], cdf[9];
    const size_t centre = size_t(y) * W + x;
    const uint cIndex = p.scratchCandidate[centre];
    const bool centerIsSSC = cIndex != kInvalidLight && (cIndex & kSSCFlag) != 0, centerIsNotEmpty = cIndex != kInvalidLight;
    float total = 0.0f; int n = 0;
    for (int dy = -1; dy <= 1; dy++) for (int dx = -1; dx <= 1; dx++)
    {
        const size_t at = size_t(clampi(y + dy, 0, H - 1)) * W + clampi(x + dx, 0, W - 1);
        kIndex[n] = p.scratchCandidate[at]; kWeight[n] = kIndex[n] == kInvalidLight ? 0.0f : p.scratchWeight[at];
        const bool ssc = kIndex[n] != kInvalidLight && (kIndex[n] & kSSCFlag) != 0;
        float mul = (dx == 0 && dy == 0) ? 48.0f : 1.0f;
        mul *= (centerIsSSC == ssc && centerIsNotEmpty) ? 128.0f : 1.0f;
        total = fadd_rn(total, fmul_rn(kWeight[n], mul)); cdf[n] = total; n++;
    }
    MicroRng rng = MicroRng::make(uint(x), uint(y), p.updateCounter, 7);
    const float rnd = rng.nextFloat();
    int pick = 8;
    for (int i = 0; i < 8; i++) if (rnd < fdiv_rn(cdf[i], total)) { pick = i; break; }
    p.fbCandidate[centre] = kIndex[pick]; p.fbWeight[centre] = fminf(kMaxWeight, kWeight[pick]);
}

// P0: remap, strip world-space-coherent candidates; returns the slot of feedbackCounters this reservoir counts towards (the caller adds 1 there, atomically on the device)
PT_HD uint p0Pixel(const Params& p, int x, int y)
{
    Reservoir r = Reservoir::at(p.fbWeight, p.fbCandidate, size_t(y) * p.W + x);
    uint lightIndexAll = kInvalidLight;
    if (!r.empty())
    {
        uint candidate = *r.c; bool ssc = false;
        if (candidate != kInvalidLight) { ssc = (candidate & kSSCFlag) != 0; candidate &= ~kSSCFlag; }
        candidate = remapPastToCurrent(p, candidate);
        lightIndexAll = candidate;
        if (!ssc) candidate = kInvalidLight;
        *r.c = candidate | (ssc ? kSSCFlag : 0u);
        if (candidate == kInvalidLight) r.clear();
    }
    return minu(lightIndexAll, p.lightCount);
}

// ComputeWeights + ImportanceBooster for the 32 lights of block `block` (the reference gives each thread LLB_LOCAL_BLOCK_SIZE = 32 lights, summed in index order): stores the boosted
// weights, returns the block's sum.  Single IEEE operations throughout: the sum feeds integer proxy counts that must equal the oracle's.
PT_HD float weightBlock(const Params& p, uint block)
{
    float blockSum = 0.0f;
    const uint end = minu(block * 32u + 32u, p.lightCount);
    for (uint i = block * 32u; i < end; i++)
    {
        float w = p.lightWeights[i];
        if (p.boostFlags & 1u)
        {
            const uint4 rec = p.lightRecords[size_t(i) * 2];
            float boostK;
            if (((rec.w >> 24) & 0xFu) == 5u) boostK = 0.5f;                 // kEnvironmentQuad: no position, half boost
            else
            {
                const float cx = bitsToFloat(rec.x), cy = bitsToFloat(rec.y), cz = bitsToFloat(rec.z);
                float distMin = 0.0f;
                for (int k = 0; k < 5; k++) distMin = fminf(distMin, fsub_rn(fadd_rn(fadd_rn(fmul_rn(cx, p.frustumPlanes[k][0]), fmul_rn(cy, p.frustumPlanes[k][1])), fmul_rn(cz, p.frustumPlanes[k][2])), p.frustumPlanes[k][3]));
                boostK = sat(fsub_rn(1.0f, fdiv_rn(fmaxf(0.0f, -distMin), fmaxf(1e-5f, p.boostFrustumFadeDistance))));
            }
            w = fmul_rn(w, fadd_rn(1.0f, fmul_rn(p.boostFrustumMul, boostK)));
        }
        if ((p.boostFlags & 2u) && p.lastFrameFeedbackAvailable)
        {
            float hist = 0.0f;                                                  // the light's boosted weight of last frame, 0 for a light that was not there (ImportanceBooster, LightsBaker.hlsl:139-141)
            if (!p.currentToPast) hist = p.histWeights[i]; else { const uint h = p.currentToPast[i]; if (h != kInvalidLight) hist = p.histWeights[h]; }
            const float d = fsub_rn(w, fmul_rn(hist, 1.1f));
            if (d > 0.0f) w = fadd_rn(w, fmul_rn(p.boostIntensityDeltaMul, d));
        }
        p.curWeights[i] = w;
        blockSum = fadd_rn(blockSum, w);
    }
    return blockSum;
}

// ComputeProxyCounts: proxies of one light from its power-based weight blended with last frame's usage; validFeedbackCount = W * H - feedbackCounters[ lightCount ]
PT_HD uint proxyCountOfLight(const Params& p, uint lightIndex)
{
    const float weightsSum = *p.weightsSumDev;
    float lightWeight = p.curWeights[lightIndex];
    if (p.lastFrameFeedbackAvailable)
    {
        const uint valid = p.W * p.H - p.feedbackCounters[p.lightCount];
        const float feedbackWeight = fdiv_rn(fmul_rn(float(p.feedbackCounters[lightIndex]), weightsSum), fmaxf(1.0f, float(valid)));
        lightWeight = fadd_rn(lightWeight, fmul_rn(fsub_rn(feedbackWeight, lightWeight), p.globalFeedbackUseWeight));
    }
    const uint budget = kProxyRatio * (p.lightCount > kMaxLights / 10 ? p.lightCount : kMaxLights / 10);
    uint proxies = 0;
    if (lightWeight > 0.0f) proxies = p.neeType == 0 ? 1u : uint(ceilf(fdiv_rn(fmul_rn(float(budget - p.lightCount), lightWeight), weightsSum)));
    return minu(proxies, kMaxProxiesPerLight - 1);
}
// proxy fill: the light that owns slot `slot` = the last light whose exclusive offset is <= slot (lights without proxies share their successor's offset and are skipped)
PT_HD uint lightOfProxySlot(const Params& p, uint slot)
{
    uint lo = 0, hi = p.lightCount;          // invariant: offsets[lo] <= slot, answer in [lo, hi)
    while (hi - lo > 1) { const uint mid = (lo + hi) >> 1; if (p.proxyOffsets[mid] <= slot) lo = mid; else hi = mid; }
    return lo;
}

PT_HD uint sampleLightGlobal(const Params& p, MicroRng& rng)
{
    const float rnd = rng.nextFloat(); const uint M = *p.samplingProxyCount;
    return p.proxyIndices[minu(uint(rnd * float(M)), M - 1)];
}
PT_HD int mirrorCoord(int v, int maxRes) { int r = v >= 0 ? v : -v; r = r < maxRes ? r : 2 * maxRes - 2 - r; return clampi(r, 0, maxRes - 1); }
// false when disoccluded (then the historic position is the pixel itself)
PT_HD bool reproject(const Params& p, int px, int py, int& hx, int& hy)
{
    if (!p.enableMotionReprojection) { hx = px; hy = py; return true; }
    float mx = 0.0f, my = 0.0f;
    if (p.motion) { const uint2 m = p.motion[size_t(py) * p.W + px]; mx = f16tof32(m.x); my = f16tof32(m.x >> 16); }
    const float cx = float(px) + 0.5f, cy = float(py) + 0.5f;
    mx = fsub_rn(fmul_rn(fadd_rn(cx, mx), 1.0f), cx); my = fsub_rn(fmul_rn(fadd_rn(cy, my), 1.0f), cy);          // ConvertMotionVectorToPixelSpace, PrevOverCurrentViewportSize = 1
    hx = int(fadd_rn(fadd_rn(float(px), mx), 0.5f)); hy = int(fadd_rn(fadd_rn(float(py), my), 0.5f));
    bool disocclusion = false;
    if (!(hx >= 0 && hy >= 0 && hx < int(p.W) && hy < int(p.H))) disocclusion = true;
    else
    {
        const float historic = p.historyDepth[size_t(hy) * p.W + hx], current = p.depth[size_t(py) * p.W + px];
        const float a = fdiv_rn(historic, current), b = fdiv_rn(current, historic);
        disocclusion = (a < b ? b : a) > p.depthDisocclusionThreshold;
    }
    if (disocclusion) { hx = px; hy = py; }
    return !disocclusion;
}

PT_HD void p1aPixel(const Params& p, uint lx, uint ly)
{
    const int T = int(kEarlyFeedbackTileSize);
    MicroRng rng = MicroRng::make(lx, ly, p.updateCounter, 3);
    Reservoir out = Reservoir::at(p.blendedWeight, p.blendedCandidate, size_t(ly) * p.blendedW + lx);
    out.clear();
    if (p.lastFrameFeedbackAvailable)
        for (int x = -1; x < T + 1; x++) for (int y = -1; y < T + 1; y++)
        {
            const int px = clampi(int(lx) * T + x, 0, int(p.W) - 1), py = clampi(int(ly) * T + y, 0, int(p.H) - 1);
            const float baseWeight = (x < 0 || y < 0 || x >= T || y >= T) ? p.reservoirHistoryDropoff : 1.0f;
            int hx, hy;
            if (reproject(p, px, py, hx, hy))
            {
                const Reservoir src = Reservoir::at(p.fbWeight, p.fbCandidate, size_t(hy) * p.W + hx);
                if (!src.empty()) out.merge(rng.nextFloat(), src, baseWeight);
            }
        }
    if (*out.c == kInvalidLight) *out.c = sampleLightGlobal(p, rng);
}

PT_HD void p1bPixel(const Params& p, uint x, uint y)
{
    MicroRng rng = MicroRng::make(x, y, p.updateCounter, 4);
    int hx, hy; const bool reprojectionValid = reproject(p, int(x), int(y), hx, hy);
    Reservoir target = Reservoir::at(p.scratchWeight, p.scratchCandidate, size_t(y) * p.W + x);
    if (!p.lastFrameFeedbackAvailable) { target.clear(); *target.c = sampleLightGlobal(p, rng); return; }
    target.cloneFrom(Reservoir::at(p.fbWeight, p.fbCandidate, size_t(hy) * p.W + hx), reprojectionValid ? 1.0f : 0.0f);
    const Reservoir src = Reservoir::at(p.blendedWeight, p.blendedCandidate, size_t(y / kEarlyFeedbackTileSize) * p.blendedW + x / kEarlyFeedbackTileSize);
    if (!src.empty()) target.merge(rng.nextFloat(), src, p.reservoirHistoryDropoff);
    uint res = *target.c;
    if (res == kInvalidLight)
    {
        if (reprojectionValid && p.lastFrameLocalSamplesAvailable)
        {
            const uint tx = (uint(hx) + p.jitterPrevX) / kTileSize, ty = (uint(hy) + p.jitterPrevY) / kTileSize;
            const uint indexInIndex = rng.next() % kLocalProxyCount;
            res = remapPastToCurrent(p, miniListLight(p.localSamplingBuffer[tileBaseAddress(p, tx, ty) + indexInIndex]));
        }
        if (res == kInvalidLight) res = sampleLightGlobal(p, rng);
        *target.c = res;
    }
}

// P2 (FillTile): key `slot` (0..127) of tile (tx, ty): slots 0..63 are the 8x8 window of processed reservoirs (x-major), slots 64..127 top-up picks from the blended image
// around the tile.  The top-up picks share one generator seeded per tile (two draws each, in order), so the thread that produces slot 64 + i first advances its own copy 2 i
// draws - at most 126 iterations of a five-instruction hash, which buys a tile filled by 128 independent lanes instead of one.
PT_HD uint fillTileEntry(const Params& p, uint tx, uint ty, uint slot)
{
    const int W = int(p.W), H = int(p.H), margin = int(kWindowSize - kTileSize) / 2;
    const int cellX = int(tx * kTileSize) - int(p.jitterX), cellY = int(ty * kTileSize) - int(p.jitterY);
    uint key;
    if (slot < kWindowSize * kWindowSize)
    {
        const int x = int(slot / kWindowSize), y = int(slot % kWindowSize);
        key = p.scratchCandidate[size_t(mirrorCoord(cellY - margin + y, H)) * W + mirrorCoord(cellX - margin + x, W)];
    }
    else
    {
        MicroRng rng = MicroRng::make(tx, ty, p.updateCounter, 5);
        for (uint i = 0; i < 2 * (slot - kWindowSize * kWindowSize); i++) rng.next();
        const float centerX = float(cellX) + 4.0f, centerY = float(cellY) + 4.0f, radius = float(kWindowSize) * 4.0f;      // + kTileSize * 0.5: exact
        const float ox = fmul_rn(fsub_rn(rng.nextFloat(), 0.5f), radius), oy = fmul_rn(fsub_rn(rng.nextFloat(), 0.5f), radius);
        const int px = mirrorCoord(int(fadd_rn(fadd_rn(centerX, ox), 0.5f)), W), py = mirrorCoord(int(fadd_rn(fadd_rn(centerY, oy), 0.5f)), H);
        key = p.blendedCandidate[size_t(py / int(kEarlyFeedbackTileSize)) * p.blendedW + px / int(kEarlyFeedbackTileSize)];
    }
    return key & 0x007FFFFFu;           // the key a tuple can carry
}
// P3 for one tile whose 128 keys sit in `data` (shared memory on the device): one compare-exchange of the bitonic network for `thread` in [0, 64)
PT_HD void bitonicStep(uint* data, uint thread, uint k, uint j)
{
    const uint mask = j - 1, index2 = ((thread & ~mask) << 1) | (thread & mask) | j, index1 = index2 ^ (k == 2 * j ? k - 1 : j);
    const uint a = data[index1], b = data[index2];
    if (a > b) { data[index1] = b; data[index2] = a; }
}
// run length of the key at `loc` in the sorted list (the reference finds it with a two-pass range scan; the result is the count of equal keys)
PT_HD uint runLength(const uint* data, uint loc)
{
    const uint key = data[loc]; uint l = loc, r = loc;
    while (l > 0 && data[l - 1] == key) l--;
    while (r + 1 < kLocalProxyCount && data[r + 1] == key) r++;
    return r - l + 1;
}

// ClearFeedbackHistory: next frame's reservoirs start from a faded copy of the processed history (own pixel + 4 neighbours); the depth is snapshotted for reprojection
PT_HD void clearFeedbackPixel(const Params& p, uint x, uint y)
{
    const size_t at = size_t(y) * p.W + x;
    p.historyDepth[at] = p.depth[at];
    Reservoir r = Reservoir::at(p.fbWeight, p.fbCandidate, at);
    if (p.lastFrameFeedbackAvailable)
    {
        const float dropOff = p.reservoirHistoryDropoff;
        r.cloneFrom(Reservoir::at(p.scratchWeight, p.scratchCandidate, at), dropOff);
        MicroRng rng = MicroRng::make(x, y, p.updateCounter, 6);
        const int ox[4] = { -1, 1, 0, 0 }, oy[4] = { 0, 0, -1, 1 };
        for (int i = 0; i < 4; i++)
        {
            const size_t s = size_t(clampi(int(y) + oy[i], 0, int(p.H) - 1)) * p.W + clampi(int(x) + ox[i], 0, int(p.W) - 1);
            const Reservoir src = Reservoir::at(p.scratchWeight, p.scratchCandidate, s);
            if (!src.empty()) r.merge(rng.nextFloat(), src, fmul_rn(dropOff, dropOff));
        }
        if (r.total() < 1e-12f) r.clear();
    }
    else r.clear();
}

} } // namespace pt::neeat
