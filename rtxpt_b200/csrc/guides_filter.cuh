// guides_filter.cuh - DenoisingGuidesBaker::DenoiseSpecHitT (Rtxpt/ProcessingPasses/DenoisingGuidesBaker.hlsl:53-115, .cpp:62-84): the 5x5 depth-aware spread of the specular hit
// distance guide that Sample::PathTrace runs after the FILL pass (Sample.cpp:2541-2543), as a __host__ __device__ pixel function (kernel: realtime_kernels.cu; host build for
// the CPU parity test: tests/emu).
#pragma once
#include "device_math.cuh"

namespace pt {

PT_HD float specHitTNeighbourhood(const float* src, const float* depth, int W, int H, int px, int py)
{
    const float centerD = depth[size_t(py) * W + px];
    float prevHitT = fmaxf(0.0f, src[size_t(py) * W + px]);
    if (prevHitT < 5e-2f) prevHitT = 0.0f;             // below the storage / computation precision of "normal" scene scales
    float vAvg = prevHitT, sumW = prevHitT > 0.0f ? 1.0f : 0.0f;
    for (int x = -2; x <= 2; x++) for (int y = -2; y <= 2; y++)
    {
        if (x == 0 && y == 0) continue;
        const int nx = px + x, ny = py + y;
        if (nx < 0 || ny < 0 || nx >= W || ny >= H) continue;
        const float v = fminf(src[size_t(ny) * W + nx], 65504.0f), d = fmaxf(0.0f, depth[size_t(ny) * W + nx]);
        if (v > 0.0f && fabsf(d - centerD) <= (d + centerD + 1e-5f) * 0.025f) { vAvg += v; sumW += 1.0f; }
    }
    if (sumW == 0.0f) return prevHitT;
    vAvg /= sumW;
    return prevHitT <= 0.0f ? vAvg : fminf(prevHitT * 1.5f + 0.5f, vAvg);
}

} // namespace pt
