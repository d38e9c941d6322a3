#!/usr/bin/env bash
# ORACLE/_ref — TEST INFRASTRUCTURE.  Writes ONE C++ translation unit to stdout: the HLSL shim, then the UNMODIFIED reference material headers read from where they lie under
# $REF (default /root/reference) through a stream filter, then the known-answer generator given as $1.  The Makefile pipes this into `g++ -x c++ -`; nothing of the reference is
# written to disk or into this repository.  The filter only rewrites spellings C++ has no equivalent for (it does not touch arithmetic):
#   out / inout T name        -> T& name                      (HLSL output parameters)
#   scalar.xx / scalar.x      -> float2( s, s ) / s            (swizzles of scalars)
#   0.xxx                     -> float3( 0, 0, 0 )
#   (e).xxx as the argument of float3( )                      -> (e)   (the constructor splats);  v.xyzw -> v
#   1.5 (unsuffixed literal)  -> 1.5f                          (an HLSL floating literal takes the type of the expression it meets - binary32 here - where C++ would make it a double
#                                                               and carry the whole expression in binary64)
#   c ? 0.f : dataRoughness  -> c ? 0.f : (float)dataRoughness (StandardBSDF.hlsli:98; HLSL promotes the float16_t arm, C++ finds the two arms ambiguous)
#   ( in T name / , in T name  -> T name                       (HLSL input qualifier);  (StructName)0 -> StructName{}  (zero initialisation)
#   radiance.xxx (scalar)     -> float3( radiance, .. )        (PolymorphicLight.hlsli:773, :787)
#   [unroll] [loop] [branch] [flatten] [mutating]              (attributes: dropped)
#   #include "local header"                                    (dropped: the files are emitted here in dependency order)
#   #if !defined(__cplusplus)                                  -> #if 1   (the shader half is what is being compiled)
set -euo pipefail
REF=${REF:-/root/reference}
PT=$REF/Rtxpt/Shaders/PathTracer
MAIN=$1
filter() {
  sed -E \
    -e 's/^[[:space:]]*#include[[:space:]]+".*$//' \
    -e 's/#if[[:space:]]+!defined\(__cplusplus\)/#if 1/' \
    -e 's/#ifndef[[:space:]]+__cplusplus/#if 1/' \
    -e 's/\b(in)?out[[:space:]]+([A-Za-z_][A-Za-z0-9_]*)[[:space:]]+([A-Za-z_][A-Za-z0-9_]*)\[/\2 \3[/g' \
    -e 's/\b(in)?out[[:space:]]+([A-Za-z_][A-Za-z0-9_]*)[[:space:]]+([A-Za-z_][A-Za-z0-9_]*)/\2\& \3/g' \
    -e 's/([(,][[:space:]]*)in[[:space:]]+/\1/g' \
    -e 's/\(([A-Z][A-Za-z0-9_]*)\)[[:space:]]*0([^.0-9a-zA-Z_]|$)/\1{}\2/g' \
    -e 's/\b(radiance|unpackedRadiance)\.xxx\b/float3(\1, \1, \1)/g' \
    -e 's/\buniform[[:space:]]+//g' \
    -e 's/#ifdef[[:space:]]+__cplusplus/#if 0/' \
    -e 's/#if[[:space:]]+defined\(__cplusplus\)/#if 0/' \
    -e 's/\b(RTXPT_LIGHTING_SAMPLING_BUFFER_TILE_SIZE|width)\.xx\b/uint2(\1, \1)/g' \
    -e 's/\bthis\./this->/g' \
    -e 's/float\(0\)\.rrr/float3(0,0,0)/g' \
    -e 's/1\.#INF/asfloat(0x7F800000u)/g' \
    -e 's/DeltaLobe deltaLobes\[cMaxDeltaLobes\]; uint deltaLobeCount;/DeltaLobe deltaLobes[cMaxDeltaLobes]; int deltaLobeCount;/' \
    -e 's/\b([0-9]+\.[0-9]+)\.xxx\b/float3(\1, \1, \1)/g' \
    -e 's/\b(kNRDMinReflectance|kNRDMaxReflectance)\.xxx\b/float3(\1, \1, \1)/g' \
    -e 's/\bxform\[([0-2])\]\.xyz\b/shimRow(xform, \1)/g' \
    -e 's/^([[:space:]]*)xform\[([0-2])\] = (.*);/\1shimSetRow(xform, \2, \3);/' \
    -e 's/\bxform\[([0-2])\]([,)[:space:]])/shimRow(xform, \1)\2/g' \
    -e 's/\b0\.xxx\b/float3(0,0,0)/g' \
    -e 's/\b0\.xxxx\b/float4(0,0,0,0)/g' \
    -e 's/\bHLF_MAX\.xxxx\b/float4(HLF_MAX,HLF_MAX,HLF_MAX,HLF_MAX)/g' \
    -e 's/\bHLF_MAX\.xxx\b/float3(HLF_MAX,HLF_MAX,HLF_MAX)/g' \
    -e 's/\b1\.xxx\b/float3(1,1,1)/g' \
    -e 's/\b_alpha\.xx\b/float2(_alpha, _alpha)/g' \
    -e 's/\bpackedData\.x\b/packedData/g' \
    -e 's/\?\(path\.GetBsdfScatterPdf\(\)\):\(0\.0\)/?((float)path.GetBsdfScatterPdf()):(0.0)/' \
    -e 's/([A-Za-z_.]+\(\))\.xxx\b/float3((float)\1, (float)\1, (float)\1)/g' \
    -e 's/\)\.xxx\b/)/g' \
    -e 's/\.rgba\b//g' \
    -e 's/CommitPixel\( const PathState path/CommitPixel( PathState path/' \
    -e 's/\.xyzw\b//g' \
    -e 's/\? 0\.f : dataRoughness/? 0.f : (float)dataRoughness/' \
    -e 's/(^|[^A-Za-z0-9_.])([0-9]+\.[0-9]*([eE][-+]?[0-9]+)?|\.[0-9]+([eE][-+]?[0-9]+)?)([^0-9A-Za-z_.]|$)/\1\2f\5/g' \
    -e 's/(^|[^A-Za-z0-9_.])([0-9]+\.[0-9]*([eE][-+]?[0-9]+)?|\.[0-9]+([eE][-+]?[0-9]+)?)([^0-9A-Za-z_.]|$)/\1\2f\5/g' \
    -e 's/\[(unroll|loop|branch|flatten|mutating)\]//g' \
    "$1"
}
echo '#include "ref_hlsl_shim.h"'
echo '#define RTXPT_LP_TYPES_USE_16BIT_PRECISION 1      /* Sample.cpp:1017, the default */'
# the switches Sample::FillPTPipelineGlobalMacros (Sample.cpp:988-1037) passes to every path-tracer shader, at the UI's defaults (SampleUI.h:181-220); the NEE sample counts stay
# undefined so that PathTracerNEE.hlsli reads them from the constant buffer (its #ifdef arms are otherwise identical).  PATH_TRACER_MODE comes from the compiler command line
cat <<'MACROS'
#define PT_ENABLE_RUSSIAN_ROULETTE 1
#define PT_NEE_ENABLED 1
#define PT_USE_RESTIR_DI 0
#define PT_USE_RESTIR_GI 0
#define RTXPT_USE_APPROXIMATE_MIS 0
#define RTXPT_DISCARD_NON_NEE_LIGHTING 0
#define RTXPT_DISCARD_NEE_LIGHTING 0
#define RTXPT_FIREFLY_FILTER 1
#define RTXPT_ACTIVE_STABLE_PLANE_COUNT 3
#define RTXPT_NESTED_DIELECTRICS_QUALITY 1
#define RTXPT_ENABLE_LOW_DISCREPANCY_SAMPLER_FOR_BSDF 1
#define NON_PATH_TRACING_PASS 0
#ifndef PATH_TRACER_MODE
#define PATH_TRACER_MODE 0
#endif
MACROS
echo 'float3 ComputeRayOrigin(float3 pos, float3 normal);      /* PathTracerHelpers.hlsli:29-42: ShadingData.hlsli names it before the helper ranges below define it */'
for f in Config.h Utils/Math/MathConstants.hlsli Utils/Utils.hlsli:28-67 Utils/Utils.hlsli:68-92 Utils/Utils.hlsli:115-169 Utils/Utils.hlsli:170-192 Utils/Utils.hlsli:193-198 Utils/Utils.hlsli:272-370 Utils/Utils.hlsli:392-499 Utils/Utils.hlsli:510-517 Rendering/Materials/BxDFConfig.hlsli Rendering/Materials/LobeType.hlsli Scene/Material/MaterialData.hlsli \
         Utils/ColorHelpers.hlsli Utils/Math/MathHelpers.hlsli Rendering/Materials/Fresnel.hlsli Rendering/Materials/Microfacet.hlsli Rendering/Materials/IBSDF.hlsli \
         Scene/ShadingData.hlsli Rendering/Materials/BxDF.hlsli Rendering/Materials/StandardBSDF.hlsli PathTracerHelpers.hlsli:26-66 PathTracerHelpers.hlsli:155-219 PathTracerHelpers.hlsli:221-270 Rendering/Materials/TexLODHelpers.hlsli:40-161 Rendering/Materials/InteriorList.hlsli Utils/Packing.hlsli:16-51 Utils/Packing.hlsli:194-265 Utils/Geometry.hlsli Lighting/PolymorphicLightPTConfig.h Lighting/PolymorphicLight.h Lighting/LightShaping.hlsli Lighting/PolymorphicLight.hlsli Lighting/LightingConfig.h Lighting/LightingTypes.hlsli Lighting/LightingAlgorithms.hlsli Lighting/LightSampler.hlsli PathTracerShared.h Utils/Math/Ray.hlsli Utils/NoiseAndSequences.hlsli:17-18 Utils/NoiseAndSequences.hlsli:58-96 Utils/NoiseAndSequences.hlsli:121-300 Utils/SampleGenerators.hlsli:16-41 Utils/StatelessSampleGenerators.hlsli Utils/SampleGenerators.hlsli:43-112 PathTracerHelpers.hlsli:318-319 Scene/HitInfoType.hlsli Scene/SceneTypes.hlsli Scene/HitInfo.hlsli PathState.hlsli PathPayload.hlsli StablePlanes.hlsli:1-318 StablePlanes.hlsli:336-371 PathTracerDebug.hlsli PathTracerTypes.hlsli:29-220 Lighting/EnvMap.hlsli:23-48 Lighting/EnvMap.hlsli:52-93 Scene/Material/HomogeneousVolumeData.hlsli Rendering/Volumes/HomogeneousVolumeSampler.hlsli local=ref_bridge_stub.h PathTracer.hlsli:18-25 PathTracerNestedDielectrics.hlsli PathTracerStablePlanes.hlsli PathTracerNEE.hlsli PathTracer.hlsli:33-764; do
  case "$f" in
    local=*) echo; echo "#line 1 \"${f#local=}\""; cat "${f#local=}" ;;
    *:*) range=${f#*:}; f=${f%%:*}; echo; echo "#line ${range%-*} \"$PT/$f\""; filter "$PT/$f" | sed -n "${range%-*},${range#*-}p" ;;       # a line range of a header whose other parts resist (Utils.hlsli: the lpfloat typedefs, Luminance / Average, LuminanceClamp, the octahedral encodings, EvalMIS, FastSqrt / FastACos, WeightedAverage; not: PackOrthoMatrix (matrix row swizzles; pinned through ref_kat_host instead), the debug text drawing, FastACosLp)
    *)   echo; echo "#line 1 \"$PT/$f\""; filter "$PT/$f"
         if [ "$f" = Config.h ]; then echo; echo '#undef ENABLE_DEBUG_VIZUALISATIONS'; echo '#define ENABLE_DEBUG_VIZUALISATIONS 0   /* the debug overlays (a build switch of Config.h:63) write to UAVs the path does not read */'; fi ;;
  esac
done
# the tone-mapping operators (Rtxpt/ToneMapper/ToneMapping.ps.hlsli:31-129: calcLuminance ... toneMap) read a constant buffer: the shared header defines it, a global stands in
TM=$REF/Rtxpt/ToneMapper
echo; echo "#line 1 \"$TM/ToneMapping_cb.h\""; filter "$TM/ToneMapping_cb.h"
echo; echo 'ToneMappingConstants gParams;'
echo; echo "#line 31 \"$TM/ToneMapping.ps.hlsli\""; filter "$TM/ToneMapping.ps.hlsli" | sed -n '31,129p'
# the per-pixel driver's own logic (Rtxpt/Shaders/PathTracerSample.hlsl): FirstHitFromVBuffer (FILL: restart from stable plane 0) and postProcessHit (BUILD: next enqueued branch);
# the ray-generation loop around them only alternates nextHit and postProcessHit
SH=$REF/Rtxpt/Shaders
echo; echo "#line 33 \"$SH/PathTracerSample.hlsl\""; filter "$SH/PathTracerSample.hlsl" | sed -n '33,113p'
echo; echo "#line 1 \"$MAIN\""
cat "$MAIN"
