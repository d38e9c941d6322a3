// ORACLE/_ref — second known-answer binary: compiles, from where they lie under /root/reference (never copied), the host-visible C++ halves of
//   Rtxpt/Shaders/PathTracer/PathTracerShared.h   (PathTracerCameraData, PathTracerConstants, BridgeCamera :109-141)
//   Rtxpt/Shaders/PathTracer/Materials/MaterialPT.h (PTMaterialData, PTMaterialFlags_*), Rtxpt/Shaders/SubInstanceData.h,
//   Rtxpt/Shaders/PathTracer/Lighting/PolymorphicLight.h (PolymorphicLightInfo / Ex, type enum, flag bits),
//   External/Donut/include/donut/shaders/bindless.h (GeometryData, InstanceData), External/Donut/src/core/math/vector.cpp (vectorToSnorm8 / snorm8ToVector)
// and prints struct layouts, constants and function outputs as JSON.  tests/golden/make_host_golden.py commits the result as
// tests/golden/host_golden.json, which pins the C ABI struct mirrors, the camera bridge (C++ and Python) and the vertex packing.
#include <donut/core/math/math.h>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstring>
using namespace donut::math;
typedef uint32_t uint;
#include <donut/shaders/bindless.h>
#include "PathTracer/PathTracerShared.h"
#include "PathTracer/Materials/MaterialPT.h"
#include "SubInstanceData.h"
#include "PathTracer/Lighting/PolymorphicLight.h"
#include "PathTracer/Lighting/LightingTypes.hlsli"    // C++ half: LightingControlData, candidate sample counts; pulls LightingConfig.h (tile size, local proxy count ...)
#include "Libraries/MicroRng.hlsli"                    // plain struct code: compiles as C++ with donut's uint2 / float2
#include "../ToneMapper/ColorUtils.h"                  // colour temperature, white balance transform
#include "../ToneMapper/ToneMapping_cb.h"
#include "PathTracer/StablePlanes.hlsli"       // C++ half: StablePlane layout, branch-ID helpers; pulls Utils/Utils.hlsli (Morton / GenericTS addressing)

static uint32_t bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
#define OFF(T, f) printf("%s\"%s\": %zu", first ? "" : ", ", #f, offsetof(T, f)), first = false
#define LAYOUT_BEGIN(T) printf(" \"%s\": {\"size\": %zu, \"offsets\": {", #T, sizeof(T)); first = true
#define LAYOUT_END(last) printf("}}%s\n", last ? "" : ",")

int main()
{
    bool first;
    printf("{\n");
    LAYOUT_BEGIN(GeometryData);
    OFF(GeometryData, numIndices); OFF(GeometryData, numVertices); OFF(GeometryData, indexBufferIndex); OFF(GeometryData, indexOffset); OFF(GeometryData, vertexBufferIndex);
    OFF(GeometryData, positionOffset); OFF(GeometryData, prevPositionOffset); OFF(GeometryData, texCoord1Offset); OFF(GeometryData, texCoord2Offset); OFF(GeometryData, normalOffset);
    OFF(GeometryData, tangentOffset); OFF(GeometryData, curveRadiusOffset); OFF(GeometryData, materialIndex);
    LAYOUT_END(false);
    LAYOUT_BEGIN(InstanceData);
    OFF(InstanceData, flags); OFF(InstanceData, firstGeometryInstanceIndex); OFF(InstanceData, firstGeometryIndex); OFF(InstanceData, numGeometries); OFF(InstanceData, transform); OFF(InstanceData, prevTransform);
    LAYOUT_END(false);
    LAYOUT_BEGIN(SubInstanceData);
    OFF(SubInstanceData, FlagsAndAlphaInfo); OFF(SubInstanceData, GlobalGeometryIndex_PTMaterialDataIndex); OFF(SubInstanceData, EmissiveLightMappingOffset); OFF(SubInstanceData, AnalyticProxyLightIndex);
    OFF(SubInstanceData, IndexBufferIndex_VertexBufferIndex); OFF(SubInstanceData, IndexOffset); OFF(SubInstanceData, TexCoord1Offset);
    LAYOUT_END(false);
    LAYOUT_BEGIN(PTMaterialData);
    OFF(PTMaterialData, BaseOrDiffuseColor); OFF(PTMaterialData, Flags); OFF(PTMaterialData, SpecularColor); OFF(PTMaterialData, EmissiveColor); OFF(PTMaterialData, Opacity); OFF(PTMaterialData, Roughness);
    OFF(PTMaterialData, Metalness); OFF(PTMaterialData, NormalTextureScale); OFF(PTMaterialData, AlphaCutoff); OFF(PTMaterialData, TransmissionFactor); OFF(PTMaterialData, DiffuseTransmissionFactor);
    OFF(PTMaterialData, BaseOrDiffuseTextureIndex); OFF(PTMaterialData, MetalRoughOrSpecularTextureIndex); OFF(PTMaterialData, EmissiveTextureIndex); OFF(PTMaterialData, NormalTextureIndex);
    OFF(PTMaterialData, OcclusionTextureIndex); OFF(PTMaterialData, TransmissionTextureIndex); OFF(PTMaterialData, IoR); OFF(PTMaterialData, ThicknessFactor); OFF(PTMaterialData, Volume);
    OFF(PTMaterialData, ShadowNoLFadeout);
    LAYOUT_END(false);
    LAYOUT_BEGIN(PolymorphicLightInfo);
    OFF(PolymorphicLightInfo, Center); OFF(PolymorphicLightInfo, ColorTypeAndFlags); OFF(PolymorphicLightInfo, Direction1); OFF(PolymorphicLightInfo, Direction2); OFF(PolymorphicLightInfo, Scalars); OFF(PolymorphicLightInfo, LogRadiance);
    LAYOUT_END(false);
    LAYOUT_BEGIN(PolymorphicLightInfoEx);
    OFF(PolymorphicLightInfoEx, IesProfileIndex); OFF(PolymorphicLightInfoEx, PrimaryAxis); OFF(PolymorphicLightInfoEx, CosConeAngleAndSoftness); OFF(PolymorphicLightInfoEx, UniqueID);
    LAYOUT_END(false);
    LAYOUT_BEGIN(PathTracerCameraData);
    OFF(PathTracerCameraData, PosW); OFF(PathTracerCameraData, NearZ); OFF(PathTracerCameraData, DirectionW); OFF(PathTracerCameraData, PixelConeSpreadAngle); OFF(PathTracerCameraData, CameraU);
    OFF(PathTracerCameraData, FarZ); OFF(PathTracerCameraData, CameraV); OFF(PathTracerCameraData, FocalDistance); OFF(PathTracerCameraData, CameraW); OFF(PathTracerCameraData, AspectRatio);
    OFF(PathTracerCameraData, ViewportSize); OFF(PathTracerCameraData, ApertureRadius); OFF(PathTracerCameraData, Jitter);
    LAYOUT_END(false);
    printf(" \"constants\": {\"PTMaterialFlags_UseSpecularGlossModel\": %u, \"PTMaterialFlags_UseMetalRoughOrSpecularTexture\": %u, \"PTMaterialFlags_UseBaseOrDiffuseTexture\": %u, "
           "\"PTMaterialFlags_UseEmissiveTexture\": %u, \"PTMaterialFlags_UseNormalTexture\": %u, \"PTMaterialFlags_UseTransmissionTexture\": %u, \"PTMaterialFlags_MetalnessInRedChannel\": %u, "
           "\"PTMaterialFlags_ThinSurface\": %u, \"PTMaterialFlags_PSDExclude\": %u, \"PTMaterialFlags_EnableAsAnalyticLightProxy\": %u, \"PTMaterialFlags_NestedPriorityShift\": %u, "
           "\"kPolymorphicLightTypeShift\": %u, \"kPolymorphicLightShapingEnableBit\": %u, \"kPolymorphicLightShapingUseMinFalloff\": %u, "
           "\"kSphere\": %u, \"kTriangle\": %u, \"kPoint\": %u, \"kEnvironmentQuad\": %u, \"kPolymorphicLightMinLog2Radiance\": %g, \"kPolymorphicLightMaxLog2Radiance\": %g},\n",
           (uint)PTMaterialFlags_UseSpecularGlossModel, (uint)PTMaterialFlags_UseMetalRoughOrSpecularTexture, (uint)PTMaterialFlags_UseBaseOrDiffuseTexture, (uint)PTMaterialFlags_UseEmissiveTexture,
           (uint)PTMaterialFlags_UseNormalTexture, (uint)PTMaterialFlags_UseTransmissionTexture, (uint)PTMaterialFlags_MetalnessInRedChannel, (uint)PTMaterialFlags_ThinSurface, (uint)PTMaterialFlags_PSDExclude,
           (uint)PTMaterialFlags_EnableAsAnalyticLightProxy, (uint)PTMaterialFlags_NestedPriorityShift, kPolymorphicLightTypeShift, kPolymorphicLightShapingEnableBit, kPolymorphicLightShapingUseMinFalloff,
           (uint)PolymorphicLightType::kSphere, (uint)PolymorphicLightType::kTriangle, (uint)PolymorphicLightType::kPoint, (uint)PolymorphicLightType::kEnvironmentQuad,
           (double)kPolymorphicLightMinLog2Radiance, (double)kPolymorphicLightMaxLog2Radiance);
    // BridgeCamera: inputs and the 28 words of PathTracerCameraData it returns
    struct CamIn { uint w, h; float pos[3], dir[3], up[3], fov, nearZ, farZ, focal, aperture, jitter[2]; };
    const CamIn cams[] = {
        { 256, 256, { 2.78f, 2.73f, -8.0f }, { 0, 0, 1 }, { 0, 1, 0 }, 0.66f, 0.1f, 1e7f, 10000.0f, 0.0f, { 0, 0 } },
        { 1920, 1080, { -20, 1.8f, 12 }, { 0.7f, -0.1f, 0.7f }, { 0, 1, 0 }, 1.04f, 0.1f, 1e7f, 10000.0f, 0.0f, { 0.25f, -0.5f } },
        { 3840, 2160, { 3.5f, 1.2f, -7.25f }, { -0.31f, 0.22f, 0.92f }, { 0.05f, 1, 0.02f }, 0.785398f, 0.05f, 5000.0f, 4.5f, 0.035f, { -0.125f, 0.375f } },
        { 641, 359, { 100.5f, -3.25f, 0.001f }, { -1, -1, -1 }, { 0, 0, 1 }, 1.5f, 1.0f, 100.0f, 10.0f, 0.5f, { 0.5f, 0.5f } } };
    printf(" \"bridge_camera\": [");
    for (size_t i = 0; i < sizeof(cams) / sizeof(cams[0]); i++)
    {
        const CamIn& c = cams[i];
        PathTracerCameraData d = BridgeCamera(c.w, c.h, float(c.w) / float(c.h), float3(c.pos[0], c.pos[1], c.pos[2]), float3(c.dir[0], c.dir[1], c.dir[2]), float3(c.up[0], c.up[1], c.up[2]),
                                              c.fov, c.nearZ, c.farZ, c.focal, c.aperture, float2(c.jitter[0], c.jitter[1]));
        printf("%s{\"in\": [%u, %u, %.9g, %.9g, %.9g, %.9g, %.9g, %.9g, %.9g, %.9g, %.9g, %.9g, %.9g, %.9g, %.9g, %.9g, %.9g, %.9g], \"out\": [", i ? ", " : "", c.w, c.h, c.pos[0], c.pos[1], c.pos[2],
               c.dir[0], c.dir[1], c.dir[2], c.up[0], c.up[1], c.up[2], c.fov, c.nearZ, c.farZ, c.focal, c.aperture, c.jitter[0], c.jitter[1]);
        uint32_t w[28]; memcpy(w, &d, sizeof(w));
        for (int k = 0; k < 28; k++) printf("%s%u", k ? ", " : "", w[k]);
        printf("]}");
    }
    printf("],\n \"snorm8\": [");
    // vectorToSnorm8<float3/float4> and the decode, on a deterministic vector set
    uint32_t s = 12345u; first = true;
    for (int i = 0; i < 64; i++)
    {
        float v[4];
        for (int k = 0; k < 4; k++) { s = s * 1664525u + 1013904223u; v[k] = (float((s >> 8) & 0xFFFF) / 32767.5f - 1.0f) * ((i % 7 == 0) ? 3.0f : 1.0f); }
        if (i < 3) { v[0] = (i == 0); v[1] = (i == 1); v[2] = (i == 2); v[3] = (i == 1) ? -1.0f : 1.0f; }
        const uint p3 = vectorToSnorm8(float3(v[0], v[1], v[2])), p4 = vectorToSnorm8(float4(v[0], v[1], v[2], (v[3] >= 0 ? 1.0f : -1.0f)));
        const float3 d3 = snorm8ToVector<3>(p3); const float4 d4 = snorm8ToVector<4>(p4);
        printf("%s[%u, %u, %u, %u, %u, %u, %u, %u, %u, %u, %u, %u, %u]", first ? "" : ", ", bits(v[0]), bits(v[1]), bits(v[2]), bits(v[3] >= 0 ? 1.0f : -1.0f), p3, p4,
               bits(d3.x), bits(d3.y), bits(d3.z), bits(d4.x), bits(d4.y), bits(d4.z), bits(d4.w));
        first = false;
    }
    printf("],\n");
    // realtime mode: StablePlane layout, constants, branch-ID arithmetic and GenericTS addressing (StablePlanes.hlsli, Utils/Utils.hlsli)
    LAYOUT_BEGIN(StablePlane);
    OFF(StablePlane, RayOrigin); OFF(StablePlane, LastRayTCurrent); OFF(StablePlane, RayDir); OFF(StablePlane, SceneLength); OFF(StablePlane, PackedThpAndMVs); OFF(StablePlane, VertexIndexAndRoughness);
    OFF(StablePlane, DenoiserPackedBSDFEstimate); OFF(StablePlane, PackedNormal); OFF(StablePlane, PackedNoisyRadianceAndSpecAvg); OFF(StablePlane, FlagsAndVertexIndex); OFF(StablePlane, PackedCounters);
    LAYOUT_END(false);
    printf(" \"stable_plane_constants\": {\"cStablePlaneCount\": %u, \"cStablePlaneMaxVertexIndex\": %u, \"cStablePlaneInvalidBranchID\": %u, \"cStablePlaneEnqueuedBranchID\": %u, \"cStablePlaneJustStartedID\": %u, \"cMaxDeltaLobes\": %u},\n",
           (uint)cStablePlaneCount, cStablePlaneMaxVertexIndex, cStablePlaneInvalidBranchID, cStablePlaneEnqueuedBranchID, cStablePlaneJustStartedID, cMaxDeltaLobes);
    printf(" \"branch_ids\": [");
    {   // random delta-lobe walks from the camera vertex: [lobes..] -> id, vertex index, on-stable-path answers against a second walk
        uint32_t r = 777u; first = true;
        for (int i = 0; i < 96; i++)
        {
            uint ids[2], depth[2];
            for (int w = 0; w < 2; w++)
            {
                r = r * 1664525u + 1013904223u; depth[w] = 1 + (r >> 28) % 15; ids[w] = 1;
                for (uint v = 1; v < depth[w]; v++) { r = r * 1664525u + 1013904223u; ids[w] = StablePlanesAdvanceBranchID(ids[w], (i % 3 == 0 && w == 1 && v < depth[0]) ? ((ids[0] >> ((depth[0] - 1 - v) * 2)) & 3) : ((r >> 30) & 1)); }
            }
            printf("%s[%u, %u, %u, %u, %u, %u, %u]", first ? "" : ", ", ids[0], ids[1], StablePlanesVertexIndexFromBranchID(ids[0]), StablePlanesVertexIndexFromBranchID(ids[1]),
                   (uint)StablePlaneIsOnStablePath(ids[0], ids[1]), (uint)StablePlaneIsOnPlane(ids[0], ids[1]), StablePlanesGetParentLobeID(ids[0]));
            first = false;
        }
    }
    printf("],\n \"generic_ts\": [");
    {
        const uint sizes[][2] = { { 64, 64 }, { 96, 96 }, { 1920, 1080 }, { 641, 359 }, { 3840, 2160 }, { 7, 5 } };
        uint32_t r = 4242u; first = true;
        for (auto& sz : sizes)
        {
            const uint line = GenericTSComputeLineStride(sz[0], sz[1]), plane = GenericTSComputePlaneStride(sz[0], sz[1]);
            printf("%s{\"size\": [%u, %u], \"line\": %u, \"plane\": %u, \"count3\": %u, \"samples\": [", first ? "" : ", ", sz[0], sz[1], line, plane, GenericTSComputeStorageElementCount(sz[0], sz[1], 3));
            for (int i = 0; i < 24; i++)
            {
                r = r * 1664525u + 1013904223u; const uint x = (i == 0) ? sz[0] - 1 : (r >> 8) % sz[0]; r = r * 1664525u + 1013904223u; const uint y = (i == 0) ? sz[1] - 1 : (r >> 8) % sz[1]; const uint p = i % 3;
                printf("%s[%u, %u, %u, %u]", i ? ", " : "", x, y, p, GenericTSPixelToAddress(uint2(x, y), p, line, plane));
            }
            printf("]}"); first = false;
        }
    }
    printf("],\n \"neeat\": {\"tile\": %d, \"window\": %d, \"local_proxies\": %d, \"search_steps\": %d, \"top_up\": %d, \"early_tile\": %d, \"proxy_ratio\": %d, \"max_lights\": %d, \"max_proxies_per_light\": %d, \"control_size\": %zu, \"local_counts\": [",
           RTXPT_LIGHTING_SAMPLING_BUFFER_TILE_SIZE, RTXPT_LIGHTING_SAMPLING_BUFFER_WINDOW_SIZE, RTXPT_LIGHTING_LOCAL_PROXY_COUNT, RTXPT_LIGHTING_LOCAL_PROXY_BINARY_SEARCH_STEPS, RTXPT_LIGHTING_TOP_UP_SAMPLES,
           RTXPT_NEEAT_EARLY_FEEDBACK_TILE_SIZE, RTXPT_LIGHTING_SAMPLING_PROXY_RATIO, RTXPT_LIGHTING_MAX_LIGHTS, RTXPT_LIGHTING_MAX_SAMPLING_PROXIES_PER_LIGHT, sizeof(LightingControlData));
    {
        const float ratios[] = { 0.0f, 0.25f, 0.5f, 0.65f, 0.9f, 1.0f }; first = true;
        for (float r : ratios) for (uint n = 1; n <= 12; n++) { printf("%s[%u, %u, %u]", first ? "" : ", ", bits(r), n, ComputeCandidateSampleLocalCount(r, n)); first = false; }
    }
    printf("]},\n \"micro_rng\": [");
    {
        const uint seeds[][4] = { { 0, 0, 1, 1 }, { 17, 5, 2, 3 }, { 1919, 1079, 977, 4 }, { 240, 135, 65535, 5 }, { 3, 700, 123456, 7 } }; first = true;
        for (auto& sd : seeds)
        {
            MicroRng r = MicroRng::make(uint2(sd[0], sd[1]), sd[2], sd[3]);
            printf("%s{\"seed\": [%u, %u, %u, %u], \"next\": [", first ? "" : ", ", sd[0], sd[1], sd[2], sd[3]);
            for (int i = 0; i < 6; i++) printf("%s%u", i ? ", " : "", r.Next());
            printf("], \"floats\": ["); for (int i = 0; i < 6; i++) printf("%s%u", i ? ", " : "", bits(r.NextFloat()));
            printf("]}"); first = false;
        }
    }
    printf("],\n \"tone_mapping\": {\"exposure_key\": %u, \"cb_size\": %zu, \"white_balance\": [", bits((float)TONEMAPPING_EXPOSURE_KEY), sizeof(ToneMappingConstants));
    {
        const float temps[] = { 1000.f, 1667.f, 2000.f, 3200.f, 4000.f, 5000.f, 6500.f, 9000.f, 20000.f }; first = true;
        for (float T : temps)
        {
            const float3 xyz = colorTemperatureToXYZ(T); const float3x3 m = calculateWhiteBalanceTransformRGB_Rec709(T);
            printf("%s{\"T\": %u, \"xyz\": [%u, %u, %u], \"m\": [", first ? "" : ", ", bits(T), bits(xyz.x), bits(xyz.y), bits(xyz.z));
            for (int i = 0; i < 9; i++) printf("%s%u", i ? ", " : "", bits(m.m_data[i]));
            const float3 g = m * float3(0.25f, 0.5f, 0.75f);
            printf("], \"m_times_v\": [%u, %u, %u]}", bits(g.x), bits(g.y), bits(g.z)); first = false;
        }
    }
    printf("]}\n}\n");
    return 0;
}
