// refit.cuh - per-frame update of the acceleration structure for rigidly animated instances (what Sample::UpdateAccelStructs / BuildTLAS do through the
// driver every frame, Rtxpt/Sample.cpp:1170-1240): the leaf triangles are re-transformed with the instances' new matrices and the compressed 8-wide BVH is refitted bottom-up,
// topology unchanged.  Bodies are __host__ __device__ (kernels: refit_kernels.cu; host build: tests/emu); the quantisation frame is the builder's own (bvh8.h), so refitting
// unmoved geometry reproduces the built nodes bit for bit.
#pragma once
#include "device_math.cuh"
#include "bvh8.h"
#include "../../include/rtxpt_b200.h"
#include <string.h>

namespace pt { namespace refit {

#ifdef __CUDA_ARCH__
PT_HD uint __float_as_uint_hd(float f) { return __float_as_uint(f); }
PT_HD float __uint_as_float_hd(uint u) { return __uint_as_float(u); }
PT_HD uint popc_hd(uint v) { return uint(__popc(v)); }
PT_HD float fmul(float a, float b) { return __fmul_rn(a, b); }
PT_HD float fadd3(float a, float b, float c, float d) { return __fadd_rn(__fadd_rn(__fadd_rn(a, b), c), d); }       // ((a + b) + c) + d, never contracted
#else
PT_HD uint __float_as_uint_hd(float f) { uint u; memcpy(&u, &f, 4); return u; }
PT_HD float __uint_as_float_hd(uint u) { float f; memcpy(&f, &u, 4); return f; }
PT_HD uint popc_hd(uint v) { return uint(__builtin_popcount(v)); }
PT_HD float fmul(float a, float b) { return a * b; }
PT_HD float fadd3(float a, float b, float c, float d) { return ((a + b) + c) + d; }
#endif

struct Params
{
    uint4* nodes; float4* tris; const uint4* triShade; const RtxptInstanceData* instances;
    float* nodeBox;                 // 6 floats per node: the exact (unquantised) bounds, what a parent needs of an internal child
    uint nodeCount, triCount;
};

// leaf triangle i: object-space vertices of its source triangle (shade record, indexed by the global triangle id) x the instance's current transform - the arithmetic of the
// scene upload's flattening (api.cu: hostXformPoint), one rounding per operation
PT_HD void refitTriangle(const Params& p, uint i)
{
    const uint gid = __float_as_uint_hd(p.tris[size_t(i) * 3].w);
    const uint4* rec = p.triShade + size_t(gid) * 6;
    const float* m = p.instances[rec[5].y].transform;
    #pragma unroll
    for (int k = 0; k < 3; k++)
    {
        const float vx = __uint_as_float_hd(rec[k].x), vy = __uint_as_float_hd(rec[k].y), vz = __uint_as_float_hd(rec[k].z);
        float4 t = p.tris[size_t(i) * 3 + k];
        t.x = fadd3(fmul(m[0], vx), fmul(m[1], vy), fmul(m[2], vz), m[3]); t.y = fadd3(fmul(m[4], vx), fmul(m[5], vy), fmul(m[6], vz), m[7]); t.z = fadd3(fmul(m[8], vx), fmul(m[9], vy), fmul(m[10], vz), m[11]);
        p.tris[size_t(i) * 3 + k] = t;
    }
}

// node ni: bounds of every child (leaf: its triangles; internal: the child's stored exact box - children sit on the next level, refitted before), union, new frame, new
// quantised child boxes.  Child metadata, childBase and triBase stay.
PT_HD void refitNode(const Params& p, uint ni)
{
    uint4* n = p.nodes + size_t(ni) * 5;
    const uint4 n0 = n[0], n1 = n[1];
    const uint imask = n0.w >> 24, childBase = n1.x, triBase = n1.y;
    float clo[8][3], chi[8][3]; bool used[8];
    float lo[3] = { 3.0e38f, 3.0e38f, 3.0e38f }, hi[3] = { -3.0e38f, -3.0e38f, -3.0e38f };
    #pragma unroll
    for (int s = 0; s < 8; s++)
    {
        const uint meta = ((s < 4 ? n1.z : n1.w) >> ((s & 3) * 8)) & 0xFFu;
        used[s] = meta != 0;
        if (!used[s]) continue;
        for (int a = 0; a < 3; a++) { clo[s][a] = 3.0e38f; chi[s][a] = -3.0e38f; }
        if (imask & (1u << s))
        {
            const float* b = p.nodeBox + size_t(childBase + popc_hd(imask & ((1u << s) - 1u))) * 6;
            for (int a = 0; a < 3; a++) { clo[s][a] = b[a]; chi[s][a] = b[3 + a]; }
        }
        else
        {
            const uint first = triBase + (meta & 31u), count = popc_hd(meta >> 5);
            for (uint t = first; t < first + count; t++) for (int k = 0; k < 3; k++)
            {
                const float4 v = p.tris[size_t(t) * 3 + k]; const float c[3] = { v.x, v.y, v.z };
                for (int a = 0; a < 3; a++) { clo[s][a] = fminf(clo[s][a], c[a]); chi[s][a] = fmaxf(chi[s][a], c[a]); }
            }
        }
        for (int a = 0; a < 3; a++) { lo[a] = fminf(lo[a], clo[s][a]); hi[a] = fmaxf(hi[a], chi[s][a]); }
    }
    float* box = p.nodeBox + size_t(ni) * 6;
    for (int a = 0; a < 3; a++) { box[a] = lo[a]; box[3 + a] = hi[a]; }
    uint32_t ebias[3];
    for (int a = 0; a < 3; a++) ebias[a] = uint32_t(bvh8FrameExponent(double(hi[a]) - double(lo[a])) + 127);
    uint q[6][2] = { { 0, 0 }, { 0, 0 }, { 0, 0 }, { 0, 0 }, { 0, 0 }, { 0, 0 } };          // [qlo.x qlo.y qlo.z qhi.x qhi.y qhi.z][slots 0-3 | 4-7]
    #pragma unroll
    for (int s = 0; s < 8; s++)
    {
        if (!used[s]) continue;
        uint8_t ql[3], qh[3]; bvh8QuantizeChild(lo, ebias, clo[s], chi[s], ql, qh);
        for (int a = 0; a < 3; a++) { q[a][s >> 2] |= uint(ql[a]) << ((s & 3) * 8); q[3 + a][s >> 2] |= uint(qh[a]) << ((s & 3) * 8); }
    }
    n[0] = make_uint4(__float_as_uint_hd(lo[0]), __float_as_uint_hd(lo[1]), __float_as_uint_hd(lo[2]), ebias[0] | (ebias[1] << 8) | (ebias[2] << 16) | (imask << 24));
    n[2] = make_uint4(q[0][0], q[0][1], q[1][0], q[1][1]); n[3] = make_uint4(q[2][0], q[2][1], q[3][0], q[3][1]); n[4] = make_uint4(q[4][0], q[4][1], q[5][0], q[5][1]);
}

} } // namespace pt::refit
