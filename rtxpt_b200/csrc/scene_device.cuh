// scene_device.cuh — device-side view of the scene tables the reference binds for the dispatch (Rtxpt/Sample.cpp:2315-2427):
// t1 SubInstanceData, t2 InstanceData, t3 GeometryData, t5 PTMaterialData, bindless buffers/textures, t10 env cube,
// t12-t18 light buffers (Rtxpt/Shaders/Bindings/LightingBindings.hlsli:20-33), plus this implementation's BVH.
#pragma once
#include "device_math.cuh"
#include "../../include/rtxpt_b200.h"
#include "bvh8.h"

namespace pt {

struct LightInfo { float cx, cy, cz; uint colorTypeAndFlags; uint direction1, direction2, scalars, logRadiance; };   // PolymorphicLightInfo, 32 B

struct SceneView
{
    const RtxptInstanceData*    instances;
    const RtxptGeometryData*    geometries;
    const RtxptSubInstanceData* subInstances;       // EmissiveLightMappingOffset filled by the light bake
    const RtxptMaterialData*    materials;
    const uint8_t*              subInstanceClass;   // shade-queue class of each sub-instance (0..3), the SER sort key analogue
    uint                        materialCount;
    const uint8_t* const*       buffers;            // bindless ByteAddressBuffers
    const cudaTextureObject_t*  textures;           // bindless Texture2D (trilinear, wrap)
    cudaTextureObject_t         envCube;            // layered 2D (6 faces), bilinear + clamp inside a face, point mip
    uint                        envFaceSize, envMipLevels;
    // acceleration structure
    const uint4*                bvhNodes;           // 5 x uint4 per node
    const float4*               bvhTris;            // 3 x float4 per triangle
    const uint4*                triInfo;            // per global triangle id: instanceIndex, geometryIndex, primitiveIndex, subInstanceIndex
    const uint4*                triShade;           // per global triangle id, 6 x uint4 (kTriShadeWords): everything loadSurface gathers per vertex, see below
    const uint4*                opacityMasks;       // per alpha-tested triangle (slot = third w word of its BVH triangle): 64 x 2-bit states, opacity_masks.h
    uint                        bvhNodeCount, bvhTriCount;
    // last frame's object-space corner positions of the triangles of geometries that carry a previous-position stream (skinned meshes; Donut's prevPositionOffset):
    // prevPosBase[sub-instance] = first triangle slot in triPrevPos (9 floats per triangle) or 0xFFFFFFFF; read by the BUILD pass's motion vectors only
    const uint*                 prevPosBase;
    const float*                triPrevPos;
    // lights
    const LightInfo*            lights;
    const uint*                 proxyCounters;
    const uint*                 proxyIndices;
    const uint*                 envLookupMap;       // 1024 x 1024 light indices
    uint                        lightCount, samplingProxyCount, envEnabled;
    const uint4*                lightsEx;           // PolymorphicLightInfoEx of the analytic lights: index = light index - 5368
    uint                        analyticLightCount;
};

// Per-triangle shading record (96 B), built at upload from the caller's index / vertex buffers so that a hit costs one contiguous read instead of
// the chain triInfo -> instance / geometry -> 3 indices -> 3 x (position, texcoord, normal, tangent):
//   q0..q2: object-space position of vertex k (xyz), its packed snorm8 normal (w)
//   q3: uv0.xy uv1.xy        q4: uv2.xy, packed tangent 0, packed tangent 1        q5: packed tangent 2, instance index, sub-instance index, primitive index | presence bits
constexpr uint kTriShadeWords = 6;
constexpr uint kTriShadeHasUV = 1u << 29, kTriShadeHasNormal = 1u << 30, kTriShadeHasTangent = 1u << 31, kTriShadePrimMask = (1u << 29) - 1u;

PT_DEVICE uint load32(const SceneView& sc, uint buffer, uint byteOffset) { return __ldg(reinterpret_cast<const uint*>(sc.buffers[buffer] + byteOffset)); }
PT_DEVICE uint3 loadIndex3(const SceneView& sc, uint buffer, uint byteOffset)
{
    const uint* p = reinterpret_cast<const uint*>(sc.buffers[buffer] + byteOffset);
    return make_uint3(__ldg(p), __ldg(p + 1), __ldg(p + 2));
}
PT_DEVICE float3 loadFloat3(const SceneView& sc, uint buffer, uint byteOffset)
{
    const float* p = reinterpret_cast<const float*>(sc.buffers[buffer] + byteOffset);
    return mk3(__ldg(p), __ldg(p + 1), __ldg(p + 2));
}
PT_DEVICE float2 loadFloat2(const SceneView& sc, uint buffer, uint byteOffset)
{
    const float* p = reinterpret_cast<const float*>(sc.buffers[buffer] + byteOffset);    // only 4-byte aligned in Donut's SoA vertex buffer
    return mk2(__ldg(p), __ldg(p + 1));
}

} // namespace pt
