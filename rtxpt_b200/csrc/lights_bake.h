// lights_bake.h — host-side light baking (see lights_bake.cpp).
#pragma once
#include <stdint.h>
#include <vector>
#include "../../include/rtxpt_b200.h"

namespace pt {

constexpr uint32_t kEnvQuadLightCount = 5368;       // RTXPT_NEEAT_ENVMAP_QT_TOTAL_NODE_COUNT (Lighting/LightingConfig.h:58-66)
constexpr uint32_t kEnvImportanceMapDim = 1024;     // EMISB_IMPORTANCE_MAP_DIM
constexpr uint32_t kMaxLights = 512 * 1024;         // RTXPT_LIGHTING_MAX_LIGHTS

struct BakedLight { float center[3]; uint32_t colorTypeAndFlags, direction1, direction2, scalars, logRadiance; };   // PolymorphicLightInfo (32 B)
struct BakedLightEx { uint32_t iesProfileIndex, primaryAxis, cosConeAngleAndSoftness, uniqueID; };                  // PolymorphicLightInfoEx (16 B)

struct LightBakeState
{
    std::vector<BakedLight> lights;             // [0, kEnvQuadLightCount) env quad-tree nodes, analytic lights, then one light per emissive triangle
    std::vector<BakedLight> triangleLights;     // scene-only part, baked once at upload
    std::vector<BakedLight> analyticLights;     // converted scene lights (sphere / point records), baked once at upload
    std::vector<BakedLightEx> analyticLightsEx; // their shaping records: light index - kEnvQuadLightCount
    bool hasEnvCube = false;
    std::vector<uint32_t> proxyCounters, proxyIndices, envLookupMap;
    std::vector<std::vector<float>> envRadianceMips;    // RGBA, fp16-rounded (EnvRadianceMap RGBA16F, .a = importance)
    uint32_t envMipCount = 0, triangleLightCount = 0;
    float weightsSum = 0;
    std::vector<float> weights;                 // per light, power-based (ComputeWeight): what NEE-AT's usage feedback is blended into (neeat.cuh: proxyCountOfLight)
    bool envEnabled = false;
};

struct LightBaker
{
    static void buildEnvRadianceMap(const RtxptEnvCubeDesc& cube, LightBakeState& st);
    // scene upload: env radiance/importance map, emissive triangle lights; writes EmissiveLightMappingOffset into `subInstances`
    static void prepareScene(const RtxptSceneDesc& scene, std::vector<RtxptSubInstanceData>& subInstances, LightBakeState& st);
    // constants change: env quad-tree lights (tint / rotation / importance), weights, proxy table
    static void finalize(const RtxptPathTracerConstants& consts, LightBakeState& st);
    static void finalizeWeightsAndProxies(const RtxptPathTracerConstants& consts, LightBakeState& st);
};

} // namespace pt
