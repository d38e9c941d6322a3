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

// what NEE-AT's feedback of the last frame was indexed by: enough of that frame's light list to map its indices onto the current list (LightsBaker.cpp:1086-1225)
struct LightListSnapshot
{
    bool valid = false, envEnabled = false; uint32_t analyticCount = 0, triangleCount = 0;
    std::vector<uint32_t> envNodes;         // per env quad-tree light: direction1 (x << 16 | y), direction2 (dim << 16), interleaved
    std::vector<uint32_t> envLookupMap;
    uint32_t total() const { return kEnvQuadLightCount + analyticCount + triangleCount; }
};

struct LightBaker
{
    // the analytic lights of the scene changed (moved, dimmed, added, removed): reconvert them; finalize() rebuilds the list
    static void setAnalyticLights(const RtxptLightDesc* lights, uint32_t count, LightBakeState& st);
    static void snapshot(const LightBakeState& st, LightListSnapshot& out);
    // past -> current and current -> past light indices between the snapshot and the present list: environment nodes through the importance-map lookups (EnvLightsMapPastToCurrent,
    // LightsBaker.hlsl:440-465, :515-538), analytic lights by their position in the caller's array, emissive triangles by their block offset (:700-711)
    static void buildRemap(const LightListSnapshot& past, const LightBakeState& st, std::vector<uint32_t>& pastToCurrent, std::vector<uint32_t>& currentToPast);
    static void buildEnvRadianceMap(const RtxptEnvCubeDesc& cube, LightBakeState& st);
    // scene upload: env radiance/importance map, emissive triangle lights; writes EmissiveLightMappingOffset into `subInstances`
    static void prepareScene(const RtxptSceneDesc& scene, std::vector<RtxptSubInstanceData>& subInstances, LightBakeState& st);
    // constants change: env quad-tree lights (tint / rotation / importance), weights, proxy table
    static void finalize(const RtxptPathTracerConstants& consts, LightBakeState& st);
    static void finalizeWeightsAndProxies(const RtxptPathTracerConstants& consts, LightBakeState& st);
};

} // namespace pt
