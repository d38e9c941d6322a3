// material_json.cpp — RTXPT's `.material.json` files (Assets/Materials/<model>.<name>.material.json, written by PTMaterial::Write and read back
// by PTMaterial::Read, Rtxpt/Materials/MaterialsBaker.cpp:84-245) -> the 128-byte PTMaterialData record the kernels read, following
// PTMaterial::FillData (MaterialsBaker.cpp:516-591) with the PTMaterial member defaults of MaterialsBaker.h:134-201 for absent keys.
// In the reference these files override what the glTF says (MaterialsBaker.cpp:707-747, :868-917); gltf_loader.cpp applies the same rule
// when it is given the materials directory.  Host only.
#include <algorithm>
#include <cfloat>
#include <cstring>
#include <string>
#include "../../include/rtxpt_b200.h"
#include "json_min.h"

using namespace rtxpt_host;

namespace rtxpt_host {

static void readFloat3(const JValue& j, const char* key, float out[3])
{
    const JValue* v = j.find(key);
    if (v && v->type == JValue::Array && v->size() >= 3) for (int k = 0; k < 3; k++) out[k] = float(v->arr[size_t(k)].num);
}
static bool readBool(const JValue& j, const char* key, bool def) { const JValue* v = j.find(key); return (v && v->type == JValue::Bool) ? v->b : def; }

void materialFromJson(const JValue& j, RtxptMaterialJsonInfo& out)
{
    memset(&out, 0, sizeof(out));
    if (j.type != JValue::Object) failf("material JSON: top level is not an object");
    // PTMaterial member defaults (MaterialsBaker.h:134-201)
    float baseColor[3] = { 1, 1, 1 }, specular[3] = { 0, 0, 0 }, emissive[3] = { 0, 0, 0 }, volumeColor[3] = { 1, 1, 1 };
    readFloat3(j, "BaseOrDiffuseColor", baseColor); readFloat3(j, "SpecularColor", specular); readFloat3(j, "EmissiveColor", emissive); readFloat3(j, "VolumeAttenuationColor", volumeColor);
    const float emissiveIntensity = float(j.number("EmissiveIntensity", 1.0)), metalness = float(j.number("Metalness", 0.0)), roughness = float(j.number("Roughness", 0.0));
    const float opacity = float(j.number("Opacity", 1.0)), transmission = float(j.number("TransmissionFactor", 0.0)), diffuseTransmission = float(j.number("DiffuseTransmissionFactor", 0.0));
    const float normalScale = float(j.number("NormalTextureScale", 1.0)), ior = float(j.number("IoR", 1.5)), alphaCutoff = float(j.number("AlphaCutoff", 0.5));
    const float volumeDistance = float(std::min(j.number("VolumeAttenuationDistance", double(FLT_MAX)), double(FLT_MAX))), shadowFade = float(j.number("ShadowNoLFadeout", 0.0));
    const bool specGloss = readBool(j, "UseSpecularGlossModel", false), alphaTest = readBool(j, "EnableAlphaTesting", false), enableTransmission = readBool(j, "EnableTransmission", false);
    const bool metalInRed = readBool(j, "MetalnessInRedChannel", false), thin = readBool(j, "ThinSurface", false), excludeNEE = readBool(j, "ExcludeFromNEE", false);
    const bool psdExclude = readBool(j, "PSDExclude", true), proxy = readBool(j, "EnableAsAnalyticLightProxy", false), ignoreTangents = readBool(j, "IgnoreMeshTangentSpace", false);
    const bool skipRender = readBool(j, "SkipRender", false);
    const int psdDominant = j.integer("PSDDominantDeltaLobe", -1), psdBlock = j.integer("PSDBlockMotionVectorsAtSurfaceType", 0), nestedPriority = j.integer("NestedPriority", 14);

    static const char* texKeys[5] = { "BaseTexture", "OcclusionRoughnessMetallicTexture", "NormalTexture", "EmissiveTexture", "TransmissionTexture" };
    static const char* enableKeys[5] = { "EnableBaseTexture", "EnableOcclusionRoughnessMetallicTexture", "EnableNormalTexture", "EnableEmissiveTexture", "EnableTransmissionTexture" };
    for (int t = 0; t < 5; t++)
    {
        const JValue* tj = j.find(texKeys[t]);
        std::string path = (tj && tj->type == JValue::Object) ? tj->string("path") : std::string();
        for (char& ch : path) if (ch == '\\') ch = '/';
        const bool enabled = readBool(j, enableKeys[t], true) && !path.empty() && (t != 4 || enableTransmission);
        out.textureEnabled[t] = enabled ? 1u : 0u;
        out.textureSRGB[t] = (tj && tj->type == JValue::Object && readBool(*tj, "sRGB", false)) ? 1u : 0u;
        strncpy(out.texturePath[t], path.c_str(), sizeof(out.texturePath[t]) - 1);
    }

    RtxptMaterialData& d = out.data;
    uint32_t flags = 0;
    if (specGloss) flags |= RTXPT_MATFLAG_UseSpecularGlossModel;
    if (metalInRed) flags |= RTXPT_MATFLAG_MetalnessInRedChannel;
    if (thin || !enableTransmission) flags |= RTXPT_MATFLAG_ThinSurface;        // materials with no transmission are thin surfaces
    if (psdExclude) flags |= RTXPT_MATFLAG_PSDExclude;
    if (psdBlock % 2) flags |= 1u << 13;                                         // PTMaterialFlags_PSDBlockMVsAtSurfaceTypeB0
    if (psdBlock / 2) flags |= 1u << 14;                                         // PTMaterialFlags_PSDBlockMVsAtSurfaceTypeB1
    if (proxy) flags |= RTXPT_MATFLAG_EnableAsAnalyticLightProxy;
    if (ignoreTangents) flags |= RTXPT_MATFLAG_IgnoreMeshTangentSpace;
    flags |= uint32_t(std::min(nestedPriority, 14)) << RTXPT_MATFLAG_NestedPriorityShift;
    flags |= uint32_t(std::min(std::max(psdDominant + 1, 0), 7)) << 24;          // PTMaterialFlags_PSDDominantDeltaLobeP1Shift
    d.Flags = flags;                                                             // Use*Texture bits are set by whoever binds the textures
    for (int k = 0; k < 3; k++) { d.BaseOrDiffuseColor[k] = baseColor[k]; d.SpecularColor[k] = specular[k]; d.EmissiveColor[k] = emissive[k] * emissiveIntensity; d.VolumeAttenuationColor[k] = volumeColor[k]; }
    d.Roughness = roughness; d.Metalness = metalness; d.NormalTextureScale = normalScale;
    d.TransmissionFactor = enableTransmission ? transmission : 0.0f; d.DiffuseTransmissionFactor = enableTransmission ? diffuseTransmission : 0.0f;
    d.Opacity = opacity; d.AlphaCutoff = alphaCutoff; d.IoR = ior; d.VolumeAttenuationDistance = volumeDistance;
    d.ShadowNoLFadeout = std::min(std::max(shadowFade, 0.0f), 0.25f);
    d.BaseOrDiffuseTextureIndex = d.MetalRoughOrSpecularTextureIndex = d.EmissiveTextureIndex = d.NormalTextureIndex = d.OcclusionTextureIndex = d.TransmissionTextureIndex = 0xFFFFFFFFu;
    d._padding0 = 42; d._padding1 = 42.0f;
    out.enableAlphaTesting = alphaTest; out.excludeFromNEE = excludeNEE; out.skipRender = skipRender; out.enableTransmission = enableTransmission;
}

} // namespace rtxpt_host

static thread_local std::string g_materialError;

extern "C" RTXPT_API int rtxpt_b200_parse_material_json(const char* jsonText, RtxptMaterialJsonInfo* out)
{
    if (!jsonText || !out) { g_materialError = "null argument"; return RTXPT_ERR_INVALID_ARGUMENT; }
    try
    {
        JParser jp{ jsonText, jsonText + strlen(jsonText) };
        const JValue root = jp.parse();
        materialFromJson(root, *out);
    }
    catch (const LoadError& e) { g_materialError = e.msg; return RTXPT_ERR_INVALID_ARGUMENT; }
    return RTXPT_OK;
}
extern "C" RTXPT_API const char* rtxpt_b200_parse_material_json_error(void) { return g_materialError.c_str(); }
