// skinning.cuh - skinned-mesh animation (Donut's skinning pass that RTXPT runs before its BLAS updates, External/Donut/shaders/skinning_cs.hlsl:42-105,
// Rtxpt/Sample.cpp:1170-1198): stage 1 = the shader's per-vertex blend of up to four joint matrices (position, normal, tangent; snorm8 packing of donut/shaders/packing.hlsli:159-204);
// stage 2 = what the reference gets for free from shared vertex buffers: the per-triangle shade records of the path tracer (object-space corner positions, packed normals and
// tangents) are rewritten from the skinned vertices through the index buffer.  The BVH refit (refit.cuh) then picks the new positions up.  Bodies are __host__ __device__.
#pragma once
#include "device_math.cuh"

namespace pt { namespace skin {

struct Params
{
    uint numVertices, numTriangles, firstGid, flags;            // flags: bit 1 normals, bit 2 tangents (SkinningFlag_*)
    const float* positions; const uint* normals; const uint* tangents; const unsigned short* jointIndices; const float* jointWeights;     // bind pose, per vertex
    const float* jointMatrices;                                 // 16 floats per joint, row-major, row vector x matrix
    float* outPositions; uint* outNormals; uint* outTangents;   // skinned vertices
    const uint* indices;                                        // 3 per triangle (the geometry's index range)
    uint4* triShade;                                            // 6 x uint4 per source triangle
    float* triPrevPos;                                          // 9 floats per triangle of this geometry: the corners the records held before this update (BUILD pass motion vectors; Donut's prevPosition stream)
};

PT_HD uint packSnorm8x(float v) { return uint(int(clampf(v, -1.0f, 1.0f) * 127.0f)) & 0xffu; }          // Pack_R8_SNORM: truncation towards zero
PT_HD uint packSnorm8x4(float x, float y, float z, float w) { return packSnorm8x(x) | (packSnorm8x(y) << 8) | (packSnorm8x(z) << 16) | (packSnorm8x(w) << 24); }

PT_HD void skinVertex(const Params& p, uint i)
{
    float m[16]; for (int k = 0; k < 16; k++) m[k] = 0.0f;
    for (int j = 0; j < 4; j++)
    {
        const float w = p.jointWeights[size_t(i) * 4 + j];
        if (w > 0.0f) { const float* jm = p.jointMatrices + size_t(p.jointIndices[size_t(i) * 4 + j]) * 16; for (int k = 0; k < 16; k++) m[k] = m[k] + jm[k] * w; }
    }
    const float px = p.positions[size_t(i) * 3], py = p.positions[size_t(i) * 3 + 1], pz = p.positions[size_t(i) * 3 + 2];
    // mul( float4( position, 1 ), jointMatrix ).xyz
    p.outPositions[size_t(i) * 3] = ((px * m[0] + py * m[4]) + pz * m[8]) + m[12];
    p.outPositions[size_t(i) * 3 + 1] = ((px * m[1] + py * m[5]) + pz * m[9]) + m[13];
    p.outPositions[size_t(i) * 3 + 2] = ((px * m[2] + py * m[6]) + pz * m[10]) + m[14];
    for (int which = 0; which < 2; which++)
    {
        if (!(p.flags & (which == 0 ? 2u : 4u))) continue;
        const uint packed = which == 0 ? p.normals[i] : p.tangents[i];
        const float vx = unpackSnorm8(packed), vy = unpackSnorm8(packed >> 8), vz = unpackSnorm8(packed >> 16), vw = unpackSnorm8(packed >> 24);
        float3 t = mk3((vx * m[0] + vy * m[4]) + vz * m[8], (vx * m[1] + vy * m[5]) + vz * m[9], (vx * m[2] + vy * m[6]) + vz * m[10]);
        t = norm3(t);
        (which == 0 ? p.outNormals : p.outTangents)[i] = packSnorm8x4(t.x, t.y, t.z, vw);
    }
}
// shade record of source triangle firstGid + t: corner k = ( position bits, packed normal ), tangents in [4].z, [4].w, [5].x (api.cu: rtxpt_b200_upload_scene)
PT_HD void gatherTriangle(const Params& p, uint t)
{
    uint4* rec = p.triShade + size_t(p.firstGid + t) * 6;
    for (int k = 0; k < 3; k++)
    {
        const uint v = p.indices[size_t(t) * 3 + k];
        uint4 r = rec[k];
        if (p.triPrevPos) { float* q = p.triPrevPos + size_t(t) * 9 + k * 3; q[0] = bitsToFloat(r.x); q[1] = bitsToFloat(r.y); q[2] = bitsToFloat(r.z); }
        r.x = floatBits(p.outPositions[size_t(v) * 3]); r.y = floatBits(p.outPositions[size_t(v) * 3 + 1]); r.z = floatBits(p.outPositions[size_t(v) * 3 + 2]);
        if (p.flags & 2u) r.w = p.outNormals[v];
        rec[k] = r;
        if (p.flags & 4u) { const uint tan = p.outTangents[v]; if (k == 0) rec[4].z = tan; else if (k == 1) rec[4].w = tan; else rec[5].x = tan; }
    }
}

// previous-position range of a newly registered skin: the corners the shade records hold now (no motion until the first update)
PT_HD void initPrevTriangle(const Params& p, uint t)
{
    const uint4* rec = p.triShade + size_t(p.firstGid + t) * 6;
    for (int k = 0; k < 3; k++) { float* q = p.triPrevPos + size_t(t) * 9 + k * 3; q[0] = bitsToFloat(rec[k].x); q[1] = bitsToFloat(rec[k].y); q[2] = bitsToFloat(rec[k].z); }
}

} } // namespace pt::skin
