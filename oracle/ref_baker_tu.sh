#!/usr/bin/env bash
# ORACLE/_ref — TEST INFRASTRUCTURE.  The baker variant of ref_hlsl_tu.sh (same filter, NEEAT_BAKER_ONLY = 1): the lighting headers, then Rtxpt/Lighting/LightsBaker.hlsl's feedback
# passes (lines 98-257 and 719-1855: everything but the environment quad-tree passes and BakeEmissiveTriangles, which read the scene's bindless buffers), then $1.
# Original header of ref_hlsl_tu.sh follows.  Writes ONE C++ translation unit to stdout: the HLSL shim, then the UNMODIFIED reference material headers read from where they lie under
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
#   return float2( NextFloat(), NextFloat() )                  -> two statements, x first   (MicroRng.hlsli:59: DXC evaluates arguments left to right, g++ does not promise to)
#   static const float neighbourWeight = <per-frame constant>  -> const float ...   (LightsBaker.hlsl:1490: a function-local static is initialised once per PROCESS in C++, per invocation in HLSL)
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
    -e 's/\b([0-9]+\.[0-9]+)\.xx\b/float2(\1, \1)/g' \
    -e 's/\b(RTXPT_NEEAT_EARLY_FEEDBACK_TILE_SIZE)\.xx\b/int2(\1, \1)/g' \
    -e 's/dispatchThreadID\.x \* LLB_LOCAL_BLOCK_SIZE/dispatchThreadID * LLB_LOCAL_BLOCK_SIZE/' \
    -e 's/([^0-9A-Za-z_.])([0-9]+)\.xx\b/\1int2(\2, \2)/g' \
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
    -e 's/[[:space:]]*:[[:space:]]*register\([^)]*\)//g' \
    -e 's/[[:space:]]*:[[:space:]]*SV_[A-Za-z]+//g' \
    -e 's/^\[numthreads\([^]]*\]//' \
    -e 's/\bgroupshared\b/static/g' \
    -e 's/VK_BINDING\([^)]*\)//g' \
    -e 's/static const float neighbourWeight/const float neighbourWeight/' \
    -e 's/return float2\(NextFloat\(\), NextFloat\(\)\);/float2 r2; r2.x = NextFloat(); r2.y = NextFloat(); return r2;/' \
    -e 's/\[allow_uav_condition\]//g' \
    -e 's/\bAllMemoryBarrierWithGroupSync\b/GroupMemoryBarrierWithGroupSync/g' \
    "$1"
}
echo '#include "ref_hlsl_shim.h"'
echo '#define RTXPT_LP_TYPES_USE_16BIT_PRECISION 1      /* Sample.cpp:1017, the default */'
# the switches Sample::FillPTPipelineGlobalMacros (Sample.cpp:988-1037) passes to every path-tracer shader, at the UI's defaults (SampleUI.h:181-220); the NEE sample counts stay
# undefined so that PathTracerNEE.hlsli reads them from the constant buffer (its #ifdef arms are otherwise identical).  PATH_TRACER_MODE comes from the compiler command line
cat <<'MACROS'
#define NEEAT_BAKER_ONLY 1
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
         Scene/ShadingData.hlsli Rendering/Materials/BxDF.hlsli Rendering/Materials/StandardBSDF.hlsli PathTracerHelpers.hlsli:26-66 PathTracerHelpers.hlsli:155-219 PathTracerHelpers.hlsli:221-270 Rendering/Materials/TexLODHelpers.hlsli:40-161 Rendering/Materials/InteriorList.hlsli Utils/Packing.hlsli:16-51 Utils/Packing.hlsli:194-265 Utils/Geometry.hlsli Lighting/PolymorphicLightPTConfig.h Lighting/PolymorphicLight.h Lighting/LightShaping.hlsli Lighting/PolymorphicLight.hlsli Lighting/LightingConfig.h Lighting/LightingTypes.hlsli Lighting/LightingAlgorithms.hlsli; do
  case "$f" in
    local=*) echo; echo "#line 1 \"${f#local=}\""; cat "${f#local=}" ;;
    *:*) range=${f#*:}; f=${f%%:*}; echo; echo "#line ${range%-*} \"$PT/$f\""; filter "$PT/$f" | sed -n "${range%-*},${range#*-}p" ;;       # a line range of a header whose other parts resist (Utils.hlsli: the lpfloat typedefs, Luminance / Average, LuminanceClamp, the octahedral encodings, EvalMIS, FastSqrt / FastACos, WeightedAverage; not: PackOrthoMatrix (matrix row swizzles; pinned through ref_kat_host instead), the debug text drawing, FastACosLp)
    *)   echo; echo "#line 1 \"$PT/$f\""; filter "$PT/$f"
         if [ "$f" = Config.h ]; then echo; echo '#undef ENABLE_DEBUG_VIZUALISATIONS'; echo '#define ENABLE_DEBUG_VIZUALISATIONS 0   /* the debug overlays (a build switch of Config.h:63) write to UAVs the path does not read */'; fi ;;
  esac
done
LB=$REF/Rtxpt/Lighting
echo; echo "#line 1 \"$REF/Rtxpt/Shaders/Libraries/MicroRng.hlsli\""; filter "$REF/Rtxpt/Shaders/Libraries/MicroRng.hlsli"
echo; echo "#line 1 \"$REF/Rtxpt/Shaders/Libraries/NEE-AT/NEEATBaker.hlsli\""; filter "$REF/Rtxpt/Shaders/Libraries/NEE-AT/NEEATBaker.hlsli"
echo; echo 'struct SubInstanceData { uint d[4]; }; struct InstanceData { uint d[4]; }; struct GeometryData { uint d[4]; }; struct PTMaterialData { uint d[4]; };   /* scene tables: only BakeEmissiveTriangles (not compiled here) reads them */'
echo; echo "#line 40 \"$LB/LightsBaker.hlsl\""; filter "$LB/LightsBaker.hlsl" | sed -n '40,75p'
echo 'SamplerState s_point, s_linear, s_materialSampler;'
echo; echo "#line 98 \"$LB/LightsBaker.hlsl\""; filter "$LB/LightsBaker.hlsl" | sed -n '98,257p'
echo; echo "#line 719 \"$LB/LightsBaker.hlsl\""; filter "$LB/LightsBaker.hlsl" | sed -n '719,1855p'
echo; echo "#line 1 \"$MAIN\""
cat "$MAIN"
