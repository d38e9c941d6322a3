// opacity_masks.h — per-triangle opacity masks for alpha-tested geometry: this implementation's equivalent of the reference's Opacity Micro-Maps
// (Rtxpt/OpacityMicroMap/OmmBuildQueue.h:42-44: OC1_4_State, baked per alpha-tested mesh by the OMM SDK; attached to the BLAS, Rtxpt/SampleCommon/AccelerationStructureUtil.h:60-84).
// With OMMs the RT core resolves most candidate hits on alpha-tested triangles without running the any-hit shader (AlphaTestImpl, PathTracerBridgeDonut.hlsli:929-971); here the
// software traversal does the same: every alpha-tested triangle carries 64 two-bit states, one per micro-triangle of a 3-level uniform barycentric subdivision:
//     0 transparent   1 opaque   2 unknown (run the texture test)
// A state is only ever "known" when EVERY point of the micro-triangle passes / fails the alpha test, whatever the bilinear filter weights: all texels of the micro-triangle's
// (padded) mip-0 footprint are strictly above / strictly below the cutoff.  The masks therefore never change a hit: results are bit-identical with and without them
// (tests/test_gpu_parity.py::test_opacity_masks_do_not_change_hits), they only remove texture fetches from the traversal's inner loop.
#pragma once
#include <stdint.h>
#if defined(__CUDACC__)
#define OM_HD __host__ __device__ inline
#else
#define OM_HD inline
#endif

namespace pt { namespace om {

constexpr int kLevel = 3, kN = 1 << kLevel, kMicroTriangles = kN * kN;      // 8 x 8 barycentric grid, 64 micro-triangles, 128 bits per triangle
constexpr uint32_t kTransparent = 0, kOpaque = 1, kUnknown = 2;
constexpr uint32_t kNoMask = 0xFFFFFFFFu;

// micro-triangle of the barycentric point (u, v) (hit convention: P = (1-u-v) V0 + u V1 + v V2): rows of constant floor(8 v), in a row the cells by floor(8 u), two
// micro-triangles per cell (the one touching the cell's origin first), the last cell of a row has one.  Points on a shared edge may land on either side: both states are valid
// there, the baker pads every footprint.
OM_HD uint32_t microIndex(float u, float v)
{
    int iv = int(v * float(kN)); iv = iv < 0 ? 0 : (iv > kN - 1 ? kN - 1 : iv);
    int iu = int(u * float(kN)); iu = iu < 0 ? 0 : iu; if (iu > kN - 1 - iv) iu = kN - 1 - iv;
    const float fu = u * float(kN) - float(iu), fv = v * float(kN) - float(iv);
    const bool upper = (fu + fv > 1.0f) && (iu + iv < kN - 1);
    return uint32_t(iv * (2 * kN - iv) + 2 * iu + (upper ? 1 : 0));
}
OM_HD uint32_t stateOf(const uint32_t mask[4], uint32_t micro) { return (mask[micro >> 4] >> ((micro & 15u) * 2u)) & 3u; }

// ---- host-side baker (opacity_masks.cpp) ----
// mip 0 of the alpha texture: RGBA8 (rgba8 set) or RGBA32F (rgba32f set), tightly packed rows, sampled with wrap addressing
struct AlphaSource { const uint8_t* rgba8; const float* rgba32f; int width, height; };
// out: 64 x 2-bit states; counts[state] (optional) is incremented per micro-triangle
void bakeTriangle(const AlphaSource& a, uint32_t cutoffByte, const float uv[3][2], uint32_t out[4], uint32_t counts[3]);

} } // namespace pt::om
