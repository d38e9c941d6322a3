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
// Triangle (48 B): v0.xyz gid | v1.xyz subInstanceAndFlags | v2.xyz primitiveIndex
//   gid  = global triangle id (instance order, geometry order, primitive order) — the tie-break key for equal-t hits
//   subInstanceAndFlags = subInstanceIndex | (alphaTested << 30) | (excludeFromNEE << 31)
#pragma once
#include <stdint.h>
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
    std::vector<Bvh8Tri> tris;      // leaf order
    float sceneLo[3], sceneHi[3];
    double buildSeconds = 0;
    uint32_t maxDepth = 0;
};

// Binned-SAH BVH2 -> greedy 8-wide collapse -> octant slot assignment -> quantisation.  Host only.
void buildBvh8(const std::vector<BuildTriangle>& tris, Bvh8& out);

} // namespace pt
