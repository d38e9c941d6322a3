// bvh8.h — compressed 8-wide BVH (CWBVH8) layout shared by the host builder and the device traversal.
//
// The reference has no acceleration-structure format to follow: BLAS/TLAS are opaque driver objects
// (Rtxpt/Sample.cpp:1061-1240, Rtxpt/SampleCommon/AccelerationStructureUtil.h:34-100).  This is the layout of
//   H. Ylitie, T. Karras, S. Laine, "Efficient Incoherent Ray Traversal on GPUs Through Compressed Wide BVHs", HPG 2017
// 80-byte nodes (5 x 16 B, one 128-bit load each) and 48-byte triangles (3 x 16 B).
//
// Node (80 B):
//   q0: p.x p.y p.z | ex ey ez imask          origin of the quantisation grid, per-axis exponent (biased, 2^(e-127)), internal-child mask
//   q1: childBase | triBase | meta[0..3] | meta[4..7]
//        meta[i] = 0                        empty slot
//        meta[i] = 0b001_11sss (0x38 | s)   internal child in slot s (s == i)
//        meta[i] = uuu_ooooo                leaf: ooooo = first triangle (offset from triBase, 0..23), uuu = unary count 001/011/111
//   q2: qlo.x[0..7] | qlo.y[0..7]
//   q3: qlo.z[0..7] | qhi.x[0..7]
//   q4: qhi.y[0..7] | qhi.z[0..7]
//   Slot s holds the child lying towards direction (s&4 ? +x : -x, s&2 ? +y : -y, s&1 ? +z : -z) so that visiting slots in the
//   order (s XOR octant) gives front-to-back traversal for every ray octant.
// Triangle (48 B): v0.xyz gid | v1.xyz subInstanceAndFlags | v2.xyz opacity-mask slot (0xFFFFFFFF: none; field `primitiveIndex` of BuildTriangle / Bvh8Tri)
//   gid  = global triangle id (instance order, geometry order, primitive order) — the tie-break key for equal-t hits
//   subInstanceAndFlags = subInstanceIndex | (alphaTested << 30) | (excludeFromNEE << 31)
#pragma once
#include <stdint.h>
#include <math.h>
#include <vector>

namespace pt {

struct Bvh8Node { uint32_t w[20]; };        // 80 bytes
struct Bvh8Tri  { float v0[3]; uint32_t gid; float v1[3]; uint32_t subInstanceAndFlags; float v2[3]; uint32_t primitiveIndex; };   // 48 bytes

constexpr uint32_t kTriFlagAlphaTested = 1u << 30;
constexpr uint32_t kTriFlagExcludeFromNEE = 1u << 31;
constexpr uint32_t kTriSubInstanceMask = (1u << 30) - 1u;

struct BuildTriangle { float v0[3], v1[3], v2[3]; uint32_t gid, subInstanceAndFlags, primitiveIndex; };

struct Bvh8
{
    std::vector<Bvh8Node> nodes;    // breadth-first: the top of the tree is a prefix of the array (staged into shared memory by the kernels)
    std::vector<uint32_t> levelStart;   // nodes of depth d (root = 0) are [levelStart[d], levelStart[d + 1]): what a bottom-up refit walks, deepest level first (refit.cuh)
    std::vector<Bvh8Tri> tris;      // leaf order
    float sceneLo[3], sceneHi[3];
    double buildSeconds = 0;
    uint32_t maxDepth = 0;
};

// ---- quantisation frame, shared by the host builder and the device refit (refit.cuh) so that a refit of unmoved geometry reproduces the built nodes bit for bit ----
#if defined(__CUDACC__)
#define BVH8_HD __host__ __device__ inline
#else
#define BVH8_HD inline
#endif
// exponent of a node axis: the smallest e in [-126, 100] with ext / 2^e <= 255, found with exact operations only (traverse.cuh scales by a further 2^15 and by 1/|d| <= 1e20)
BVH8_HD int bvh8FrameExponent(double ext)
{
    if (!(ext > 0.0)) return -126;
    int k = 0; (void)frexp(ext, &k);                     // ext = m * 2^k, m in [0.5, 1)
    int e = k - 8;                                       // 255 < 2^8: ext / 2^(k-8) = m * 256 in [128, 256)
    if (e < -126) e = -126;
    if (e > 100) e = 100;
    while (e < 100 && ext / ldexp(1.0, e) > 255.0) e++;
    while (e > -126 && ext / ldexp(1.0, e - 1) <= 255.0) e--;
    return e;
}
// conservative 8-bit box of a child inside its parent's frame: floor / ceil in double, clamped to the grid
BVH8_HD void bvh8QuantizeChild(const float* nodeLo, const uint32_t* ebias, const float* childLo, const float* childHi, uint8_t* qlo, uint8_t* qhi)
{
    for (int a = 0; a < 3; a++)
    {
        const double scale = ldexp(1.0, int(ebias[a]) - 127);
        const double lo = floor((double(childLo[a]) - double(nodeLo[a])) / scale), hi = ceil((double(childHi[a]) - double(nodeLo[a])) / scale);
        qlo[a] = uint8_t(lo < 0.0 ? 0.0 : (lo > 255.0 ? 255.0 : lo)); qhi[a] = uint8_t(hi < 0.0 ? 0.0 : (hi > 255.0 ? 255.0 : hi));
    }
}

// Binned-SAH BVH2 -> greedy 8-wide collapse -> octant slot assignment -> quantisation.  Host only.
void buildBvh8(const std::vector<BuildTriangle>& tris, Bvh8& out);

} // namespace pt
