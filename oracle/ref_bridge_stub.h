// ORACLE/_ref - TEST INFRASTRUCTURE.  A CPU stand-in for RTXPT's `Bridge` namespace (the interface Rtxpt/Shaders/PathTracerBridge.hlsli declares and PathTracerBridgeDonut.hlsli
// implements over Donut's bindless scene): what lets the UNMODIFIED PathTracer.hlsli / PathTracerNEE.hlsli / PathTracerStablePlanes.hlsli / PathTracerNestedDielectrics.hlsli be
// compiled as C++ and driven one call at a time (oracle/ref_hlsl_tu.sh emits this file between the reference's type headers and those four).  The scene side of the bridge is DATA
// here: the surface a hit loads, the medium / IoR table, the light tables and the answer of a visibility ray are whatever the known-answer generator put into g_bridge before the
// call.  So a golden made through this stub pins the path tracer's own logic (everything in namespace PathTracer), not the Donut bridge's scene access - that stays a restatement
// (oracle/pt_scene.h) held to CUDA-vs-oracle parity.  Exports (guide buffers, specular hit distance) are recorded for the generator to write out.
#pragma once

struct ShimBridgeScenario
{
    // Bridge::get*
    uint sampleIndex = 0, maxBounces = 0, maxDiffuseBounces = 0; float noisyRadianceAttenuation = 1.0f, envMipOffset = 0.0f;
    // Bridge::loadSurface: the surface of the next hit
    PathTracer::SurfaceData surface;
    // per material: IoR and absorption (Bridge::loadIoR / loadHomogeneousVolumeData); ids past the table answer like ids past g_Const.MaterialCount
    uint materialCount = 0; float ior[8] = {}; float3 sigmaA[8];
    // Bridge::CreateLightSampler
    StructuredBuffer<LightingControlData> control; StructuredBuffer<PolymorphicLightInfo> lights; StructuredBuffer<PolymorphicLightInfoEx> lightsEx;
    Buffer<uint> proxyCounters, proxyIndices, localSampling; Texture2D<uint> envLookup; RWTexture2D<float> feedbackWeight; RWTexture2D<uint> feedbackCandidates;
    // Bridge::CreateEnvMap
    EnvMapSceneParams env; bool hasEnvMap = false;
    // Bridge::computeCameraRay / computeMotionVector: closed forms the generator and the oracle-side mirror share
    float3 cameraPos; float3 cameraDirBase, cameraDirDx, cameraDirDy;
    // what the calls exported
    uint visibilityQueries = 0; RayDesc lastVisibilityRay; bool lastVisibility = false;
    uint exportSurfaceCalls = 0, exportNonSurfaceCalls = 0, specHitTStarts = 0, specHitTStops = 0; float exportSceneLength = 0; float3 exportMotion, exportVirtualPos; float specularHitT = 0;     // u_SpecularHitT of the pixel
};
static ShimBridgeScenario g_bridge;

// the answer of a shadow ray: a function of the ray's own bits (3 of 4 rays see their light), so that a ray off by one ulp anywhere shows up in the golden
inline bool ShimVisibilityRule(const RayDesc& ray)
{
    const uint h = asuint(ray.Origin.x) ^ (asuint(ray.Origin.y) >> 1) ^ (asuint(ray.Origin.z) >> 2) ^ asuint(ray.Direction.x) ^ (asuint(ray.Direction.y) >> 1) ^ (asuint(ray.Direction.z) >> 2) ^ asuint(ray.TMax);
    return (h & 3u) != 0u;
}

namespace Bridge
{
    static uint getSampleIndex() { return g_bridge.sampleIndex; }
    static float getNoisyRadianceAttenuation() { return g_bridge.noisyRadianceAttenuation; }
    static uint getMaxBounceLimit() { return g_bridge.maxBounces; }
    static uint getMaxDiffuseBounceLimit() { return g_bridge.maxDiffuseBounces; }
    static Ray computeCameraRay(const uint2 pixelPos)
    {
        Ray ray; ray.origin = g_bridge.cameraPos; ray.dir = normalize(g_bridge.cameraDirBase + g_bridge.cameraDirDx * float(pixelPos.x) + g_bridge.cameraDirDy * float(pixelPos.y)); ray.tMin = 0.0f; ray.tMax = kMaxRayTravel;
        return ray;
    }
    static PathTracer::SurfaceData loadSurface(const uint instanceIndex, const uint geometryIndex, const uint triangleIndex, const float2 barycentrics, const float3 rayDir, const RayCone rayCone,
                                               const int pathVertexIndex, const uint2 pixelPosition, DebugContext debug) { return g_bridge.surface; }
    static void updateOutsideIoR(PathTracer::SurfaceData& surfaceData, lpfloat outsideIoR)
    {   // the bridge's rule (PathTracerBridgeDonut.hlsli:855-861), restated: eta = incident IoR / transmissive IoR for the side the ray arrives on
        surfaceData.shadingData.IoR = outsideIoR;
        surfaceData.bsdf.data.SetEta(surfaceData.shadingData.frontFacing ? (surfaceData.shadingData.IoR / surfaceData.interiorIoR) : (surfaceData.interiorIoR / surfaceData.shadingData.IoR));
    }
    static lpfloat loadIoR(const uint materialID) { return materialID >= g_bridge.materialCount ? lpfloat(1.0f) : lpfloat(g_bridge.ior[materialID]); }
    static HomogeneousVolumeData loadHomogeneousVolumeData(const uint materialID)
    {
        HomogeneousVolumeData v; v.sigmaS = float3(0, 0, 0); v.sigmaA = float3(0, 0, 0); v.g = 0.0f;
        if (materialID < g_bridge.materialCount) v.sigmaA = g_bridge.sigmaA[materialID];
        return v;
    }
    static float3 computeMotionVector(float3 posW, float3 prevPosW) { return (prevPosW - posW) * 0.5f; }
    static float3 computeSkyMotionVector(const uint2 pixelPos) { return float3(0, 0, 0); }
    static bool traceVisibilityRay(RayDesc ray, const RayCone rayCone, const int pathVertexIndex, DebugContext debug)
    {
        g_bridge.visibilityQueries++; g_bridge.lastVisibilityRay = ray; g_bridge.lastVisibility = ShimVisibilityRule(ray);
        return g_bridge.lastVisibility;
    }
    static bool HasEnvMap() { return g_bridge.hasEnvMap; }
    static EnvMap CreateEnvMap() { return EnvMap::make(TextureCube<float4>(), SamplerState(), g_bridge.env); }
    static LightSampler CreateLightSampler(const uint2 pixelPos, bool isScreenSpaceCoherent)
    {
        return LightSampler::make(g_bridge.control, g_bridge.lights, g_bridge.lightsEx, g_bridge.proxyCounters, g_bridge.proxyIndices, g_bridge.localSampling, g_bridge.envLookup, g_bridge.feedbackWeight,
                                  g_bridge.feedbackCandidates, pixelPos, isScreenSpaceCoherent);
    }
    static LightSampler CreateLightSampler(const uint2 pixelPos, float rayConeWidth, float totalPathLength)
    {   // as the bridge does (PathTracerBridgeDonut.hlsli:1080-1089): the heuristic decides which local sampler a vertex uses
        return CreateLightSampler(pixelPos, LightSampler::IsScreenSpaceCoherentHeuristic(g_bridge.control, rayConeWidth, totalPathLength));
    }
    static float DiffuseEnvironmentMapMIPOffset() { return g_bridge.envMipOffset; }
    static void ExportSurfaceInit(uint2 pixelPos) {}
    static void ExportSurface(const PathState path, PathTracer::SurfaceData surfaceData, float sceneLength, float3 motionVectors)
    { g_bridge.exportSurfaceCalls++; g_bridge.exportSceneLength = sceneLength; g_bridge.exportMotion = motionVectors; }
    static void ExportNonSurface(const PathState path, float3 virtualWorldPos, float3 motionVectors) { g_bridge.exportNonSurfaceCalls++; g_bridge.exportVirtualPos = virtualWorldPos; g_bridge.exportMotion = motionVectors; }
    // the bridge's bookkeeping (PathTracerBridgeDonut.hlsli:1155-1176), restated: the start stores minus the scene length, the stop turns a negative entry into the distance travelled since
    static void ExportSpecHitTStart(const PathState path) { g_bridge.specHitTStarts++; g_bridge.specularHitT = -path.GetSceneLength(); }
    static void ExportSpecHitTStop(const PathState path)
    {
        g_bridge.specHitTStops++;
        const float denoisingSceneLength = g_bridge.specularHitT;
        if (denoisingSceneLength < 0) g_bridge.specularHitT = max(0.0f, path.GetSceneLength() + denoisingSceneLength);
    }
}
