// ORACLE — test infrastructure only (see pt_math.h).
// pt_skinning.h: Donut's skinning pass (External/Donut/shaders/skinning_cs.hlsl:42-105; snorm8 packing donut/shaders/packing.hlsli:159-204) restated per vertex - the blend of up to
// four joint matrices applied to position, normal and tangent - and the rewrite of the per-triangle shade records the path tracer reads (what shared vertex buffers give the
// reference for free).
#pragma once
#include "pt_math.h"
#include <vector>

namespace orc { namespace skinning {

inline uint Pack_R8_SNORM(float v) { return uint(int(std::min(std::max(v, -1.0f), 1.0f) * 127.0f)) & 0xffu; }
inline float Unpack_R8_SNORM(uint v) { const int s = int(v << 24) >> 24; return std::min(std::max(float(s) / 127.0f, -1.0f), 1.0f); }

// in: bind pose; out: skinned positions (3 floats) and packed normals / tangents per vertex
inline void skinVertices(uint numVertices, const float* positions, const uint* normals, const uint* tangents, const uint16_t* jointIndices, const float* jointWeights, const float* jointMatrices,
                         float* outPositions, uint* outNormals, uint* outTangents)
{
    for (uint i = 0; i < numVertices; i++)
    {
        float m[16] = {};
        for (int j = 0; j < 4; j++)
        {
            const float w = jointWeights[size_t(i) * 4 + j];
            if (w > 0) { const float* jm = jointMatrices + size_t(jointIndices[size_t(i) * 4 + j]) * 16; for (int k = 0; k < 16; k++) m[k] += jm[k] * w; }
        }
        const float3 p = f3(positions[3 * size_t(i)], positions[3 * size_t(i) + 1], positions[3 * size_t(i) + 2]);
        outPositions[3 * size_t(i)] = ((p.x * m[0] + p.y * m[4]) + p.z * m[8]) + m[12]; outPositions[3 * size_t(i) + 1] = ((p.x * m[1] + p.y * m[5]) + p.z * m[9]) + m[13];
        outPositions[3 * size_t(i) + 2] = ((p.x * m[2] + p.y * m[6]) + p.z * m[10]) + m[14];
        for (int which = 0; which < 2; which++)
        {
            const uint* src = which == 0 ? normals : tangents; uint* dst = which == 0 ? outNormals : outTangents;
            if (!src) continue;
            const float vx = Unpack_R8_SNORM(src[i]), vy = Unpack_R8_SNORM(src[i] >> 8), vz = Unpack_R8_SNORM(src[i] >> 16), vw = Unpack_R8_SNORM(src[i] >> 24);
            const float3 t = normalize(f3((vx * m[0] + vy * m[4]) + vz * m[8], (vx * m[1] + vy * m[5]) + vz * m[9], (vx * m[2] + vy * m[6]) + vz * m[10]));
            dst[i] = Pack_R8_SNORM(t.x) | (Pack_R8_SNORM(t.y) << 8) | (Pack_R8_SNORM(t.z) << 16) | (Pack_R8_SNORM(vw) << 24);
        }
    }
}
// shade records: 6 x 4 words per source triangle; corner k = ( position bits, packed normal ), tangents in word 4.z, 4.w, 5.x
inline void gatherShadeRecords(uint numTriangles, uint firstGid, const uint* indices, const float* positions, const uint* normals, const uint* tangents, uint* triShade)
{
    for (uint t = 0; t < numTriangles; t++)
    {
        uint* rec = triShade + size_t(firstGid + t) * 24;
        for (int k = 0; k < 3; k++)
        {
            const uint v = indices[size_t(t) * 3 + k];
            memcpy(rec + 4 * k, positions + 3 * size_t(v), 12);
            if (normals) rec[4 * k + 3] = normals[v];
            if (tangents) rec[k == 0 ? 18 : (k == 1 ? 19 : 20)] = tangents[v];
        }
    }
}

} } // namespace orc::skinning
