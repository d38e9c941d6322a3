// gltf_loader.cpp — host-side glTF 2.0 scene loader: .gltf (+ external .bin, data: URIs, PNG images) or .glb  ->  the GPU-table view of
// a scene that rtxpt_b200_upload_scene consumes (RtxptSceneDesc).
//
// In the reference this is the job of Donut's GltfImporter + Scene::CreateMeshBuffers and RTXPT's MaterialsBaker
// (External/Donut/src/engine/GltfImporter.cpp:641-1430, Scene.cpp:821-1000, Rtxpt/Materials/MaterialsBaker.cpp:516-591, :660-705, :960-1017),
// all of which sit on the host side of the PathTrace boundary.  This file follows their conventions so that a maintainer can swap either
// side: SoA vertex buffers (float3 position | float2 texcoord | snorm8x4 normal | snorm8x4 tangent), truncating snorm8 packing
// (core/math/vector.cpp:84-103), tangents computed when the asset has none (GltfImporter.cpp:1331-1421), emissive factor split into a
// normalised colour and an intensity (GltfImporter.cpp:957-963), domain from alphaMode + KHR_materials_transmission (:988-994), non-transmissive
// materials forced thin (MaterialsBaker.cpp:543-544), alpha cutoff quantised to 8 bits in SubInstanceData (:990-991).
// Buffer layout: two bindless buffers per glTF mesh (indices, vertices), one GeometryData per primitive, one InstanceData per node with a mesh
// in depth-first scene order — the same layout rtxpt_b200/scene_builder.py produces, which the tests compare against byte for byte.
// Not handled (reported as errors, never silently skipped): sparse accessors, KTX images and progressive JPEGs (PNG, baseline JPEG and DDS are read), Draco/meshopt compression, skins, morph targets.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>
#include <zlib.h>
#include "../../include/rtxpt_b200.h"
#include "json_min.h"

using namespace rtxpt_host;
namespace rtxpt_host {
void materialFromJson(const JValue& j, RtxptMaterialJsonInfo& out);       // material_json.cpp
struct DdsImage { uint32_t width = 0, height = 0; bool srgb = false; std::vector<std::vector<uint8_t>> mips; };
DdsImage decodeDds(const uint8_t* data, size_t size, const char* name);   // dds.cpp
struct DdsBlocks { uint32_t width = 0, height = 0, format = 0; bool srgb = false; std::vector<std::vector<uint8_t>> mips; };
bool extractDdsBlocks(const uint8_t* data, size_t size, const char* name, DdsBlocks& out);   // dds.cpp: BC1 / BC2 / BC3 / BC7 kept compressed
struct JpegImage { uint32_t w = 0, h = 0; std::vector<uint8_t> rgba; };
JpegImage decodeJpeg(const uint8_t* data, size_t size, const char* name);   // jpeg.cpp: baseline / sequential JPEG -> RGBA8
bool g_keepBlockCompression = false;                                        // rtxpt_b200_loader_keep_block_compression
}

namespace {

// ---- files, base64 ------------------------------------------------------------------------------------------------------------------------------
std::vector<uint8_t> readFile(const std::string& path)
{
    FILE* f = fopen(path.c_str(), "rb"); if (!f) failf("cannot open '%s'", path.c_str());
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<uint8_t> d(size_t(std::max(0L, n)));
    if (n > 0 && fread(d.data(), 1, size_t(n), f) != size_t(n)) { fclose(f); failf("short read on '%s'", path.c_str()); }
    fclose(f); return d;
}
std::vector<uint8_t> base64(const char* s, size_t n)
{
    std::vector<uint8_t> out; uint32_t acc = 0; int bits = 0;
    for (size_t i = 0; i < n; i++)
    {
        const char c = s[i]; int v;
        if (c >= 'A' && c <= 'Z') v = c - 'A'; else if (c >= 'a' && c <= 'z') v = c - 'a' + 26; else if (c >= '0' && c <= '9') v = c - '0' + 52;
        else if (c == '+' || c == '-') v = 62; else if (c == '/' || c == '_') v = 63; else continue;
        acc = (acc << 6) | uint32_t(v); bits += 6;
        if (bits >= 8) { bits -= 8; out.push_back(uint8_t(acc >> bits)); }
    }
    return out;
}
std::string uriToPath(const std::string& uri, const std::string& baseDir)
{
    std::string decoded;        // percent-decoding of file URIs
    for (size_t i = 0; i < uri.size(); i++)
        if (uri[i] == '%' && i + 2 < uri.size()) { decoded += char(strtoul(uri.substr(i + 1, 2).c_str(), nullptr, 16)); i += 2; } else decoded += uri[i];
    return baseDir + decoded;
}
std::vector<uint8_t> resolveUri(const std::string& uri, const std::string& baseDir)
{
    if (uri.rfind("data:", 0) == 0)
    {
        const size_t comma = uri.find(','); if (comma == std::string::npos) failf("glTF: malformed data URI");
        return base64(uri.c_str() + comma + 1, uri.size() - comma - 1);
    }
    return readFile(uriToPath(uri, baseDir));
}
bool fileExists(const std::string& path) { FILE* f = fopen(path.c_str(), "rb"); if (f) fclose(f); return f != nullptr; }
// "<name>.png" -> "<name>.dds" when that file exists: both the glTF importer (GltfImporter.cpp:652-654, 786-796) and the material reader
// (MaterialsBaker.cpp:179-192) prefer a block-compressed sibling of an uncompressed image
std::string preferDdsSibling(const std::string& path, bool pngOnly)
{
    const size_t dot = path.find_last_of('.'), slash = path.find_last_of("/\\");
    if (dot == std::string::npos || (slash != std::string::npos && dot < slash)) return path;
    std::string ext = path.substr(dot); for (char& c : ext) c = char(tolower(c));
    if (ext == ".dds" || (pngOnly && ext != ".png")) return path;
    const std::string dds = path.substr(0, dot) + ".dds";
    return fileExists(dds) ? dds : path;
}

// ---- PNG (8/16-bit, colour types 0/2/3/4/6, non-interlaced) -> RGBA8 --------------------------------------------------------------------------------
struct Image { uint32_t w = 0, h = 0; std::vector<uint8_t> rgba; };
uint32_t be32(const uint8_t* p) { return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | p[3]; }
Image decodePng(const std::vector<uint8_t>& d, const char* name)
{
    static const uint8_t sig[8] = { 0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A };
    if (d.size() < 8 || memcmp(d.data(), sig, 8) != 0) failf("image '%s' is not a PNG (PNG, JPEG and DDS images are decoded by this loader)", name);
    Image img; uint32_t depth = 0, ctype = 0, interlace = 0; std::vector<uint8_t> idat, plte, trns;
    for (size_t off = 8; off + 12 <= d.size();)
    {
        const uint32_t len = be32(&d[off]); const char* type = reinterpret_cast<const char*>(&d[off + 4]); const uint8_t* body = &d[off + 8];
        if (off + 12 + len > d.size()) failf("PNG '%s': truncated chunk", name);
        if (!memcmp(type, "IHDR", 4))
        {
            if (len != 13) failf("PNG '%s': IHDR chunk of %u bytes", name, len);
            img.w = be32(body); img.h = be32(body + 4); depth = body[8]; ctype = body[9]; interlace = body[12];
            if (img.w > 32768 || img.h > 32768) failf("PNG '%s': %u x %u is larger than 32768 x 32768", name, img.w, img.h);
        }
        else if (!memcmp(type, "PLTE", 4)) plte.assign(body, body + len);
        else if (!memcmp(type, "tRNS", 4)) trns.assign(body, body + len);
        else if (!memcmp(type, "IDAT", 4)) idat.insert(idat.end(), body, body + len);
        else if (!memcmp(type, "IEND", 4)) break;
        off += 12 + size_t(len);
    }
    if (!img.w || !img.h) failf("PNG '%s': missing IHDR", name);
    if (interlace) failf("PNG '%s': interlaced images are not supported", name);
    if (depth != 8 && depth != 16) failf("PNG '%s': bit depth %u is not supported", name, depth);
    const uint32_t channels = (ctype == 0) ? 1 : (ctype == 2) ? 3 : (ctype == 3) ? 1 : (ctype == 4) ? 2 : (ctype == 6) ? 4 : 0;
    if (!channels || (ctype == 3 && depth != 8)) failf("PNG '%s': colour type %u is not supported", name, ctype);
    const size_t bpp = channels * (depth / 8), stride = size_t(img.w) * bpp;
    std::vector<uint8_t> raw((stride + 1) * img.h);
    uLongf rawLen = uLongf(raw.size());
    if (uncompress(raw.data(), &rawLen, idat.data(), uLong(idat.size())) != Z_OK || rawLen != raw.size()) failf("PNG '%s': zlib stream is corrupt", name);
    std::vector<uint8_t> pix(stride * img.h);
    for (uint32_t y = 0; y < img.h; y++)
    {
        const uint8_t filter = raw[(stride + 1) * y]; const uint8_t* src = &raw[(stride + 1) * y + 1];
        uint8_t* dst = &pix[stride * y]; const uint8_t* up = y ? &pix[stride * (y - 1)] : nullptr;
        for (size_t x = 0; x < stride; x++)
        {
            const int a = (x >= bpp) ? dst[x - bpp] : 0, b = up ? up[x] : 0, c = (up && x >= bpp) ? up[x - bpp] : 0;
            int pred = 0;
            switch (filter)
            {
            case 0: pred = 0; break; case 1: pred = a; break; case 2: pred = b; break; case 3: pred = (a + b) >> 1; break;
            case 4: { const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c); pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); break; }
            default: failf("PNG '%s': bad filter type", name);
            }
            dst[x] = uint8_t(src[x] + pred);
        }
    }
    img.rgba.resize(size_t(img.w) * img.h * 4);
    const size_t step = depth / 8;      // 16-bit samples: keep the high byte
    for (size_t i = 0; i < size_t(img.w) * img.h; i++)
    {
        const uint8_t* s = &pix[i * bpp]; uint8_t* o = &img.rgba[i * 4];
        switch (ctype)
        {
        case 0: o[0] = o[1] = o[2] = s[0]; o[3] = 255; break;
        case 2: o[0] = s[0]; o[1] = s[step]; o[2] = s[2 * step]; o[3] = 255; break;
        case 3: { const uint32_t k = s[0]; if (k * 3 + 2 >= plte.size()) failf("PNG '%s': palette index out of range", name); o[0] = plte[k * 3]; o[1] = plte[k * 3 + 1]; o[2] = plte[k * 3 + 2]; o[3] = k < trns.size() ? trns[k] : 255; break; }
        case 4: o[0] = o[1] = o[2] = s[0]; o[3] = s[step]; break;
        case 6: o[0] = s[0]; o[1] = s[step]; o[2] = s[2 * step]; o[3] = s[3 * step]; break;
        }
    }
    return img;
}

// box-filtered mip chain down to 1x1; the float chain is carried unrounded from level to level, every level rounds half to even
std::vector<std::vector<uint8_t>> makeMips(const Image& img, std::vector<std::pair<uint32_t, uint32_t>>& dims)
{
    std::vector<std::vector<uint8_t>> mips; mips.push_back(img.rgba); dims.push_back({ img.w, img.h });
    std::vector<float> cur(img.rgba.begin(), img.rgba.end());
    uint32_t w = img.w, h = img.h;
    while (w > 1 || h > 1)
    {
        const uint32_t nw = std::max(1u, w / 2), nh = std::max(1u, h / 2);
        std::vector<float> nxt(size_t(nw) * nh * 4);
        for (uint32_t y = 0; y < nh; y++) for (uint32_t x = 0; x < nw; x++) for (int c = 0; c < 4; c++)
        {
            auto at = [&](uint32_t xx, uint32_t yy) { return cur[(size_t(yy) * w + xx) * 4 + c]; };
            float v;
            if (h > 1 && w > 1) v = (((at(2 * x, 2 * y) + at(2 * x, 2 * y + 1)) + at(2 * x + 1, 2 * y)) + at(2 * x + 1, 2 * y + 1)) * 0.25f;
            else if (h > 1) v = (at(x, 2 * y) + at(x, 2 * y + 1)) * 0.5f;
            else v = (at(2 * x, y) + at(2 * x + 1, y)) * 0.5f;
            nxt[(size_t(y) * nw + x) * 4 + c] = v;
        }
        std::vector<uint8_t> q(nxt.size());
        for (size_t i = 0; i < nxt.size(); i++) q[i] = uint8_t(std::min(255.0f, std::max(0.0f, std::nearbyintf(nxt[i]))));
        mips.push_back(std::move(q)); dims.push_back({ nw, nh });
        cur.swap(nxt); w = nw; h = nh;
    }
    return mips;
}

// ---- small vector helpers -----------------------------------------------------------------------------------------------------------------------------
struct F3 { float x, y, z; };
inline F3 operator-(F3 a, F3 b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
inline F3 operator+(F3 a, F3 b) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
inline F3 operator*(F3 a, float s) { return { a.x * s, a.y * s, a.z * s }; }
inline F3 cross(F3 a, F3 b) { return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; }
inline float norm(F3 a) { return std::sqrt((a.x * a.x + a.y * a.y) + a.z * a.z); }
inline bool finitef(float v) { return std::isfinite(v); }

uint32_t packSnorm8x3(F3 v)     // vectorToSnorm8<3>: truncating int(v * 127 / |v|)
{
    const float scale = 127.0f / std::max(norm(v), 1e-30f);
    return (uint32_t(int(v.x * scale)) & 0xff) | ((uint32_t(int(v.y * scale)) & 0xff) << 8) | ((uint32_t(int(v.z * scale)) & 0xff) << 16);
}
uint32_t packSnorm8x4(F3 v, float w)
{
    const float scale = 127.0f / std::max(norm(v), 1e-30f);
    return (uint32_t(int(v.x * scale)) & 0xff) | ((uint32_t(int(v.y * scale)) & 0xff) << 8) | ((uint32_t(int(v.z * scale)) & 0xff) << 16) | ((uint32_t(int(w * scale)) & 0xff) << 24);
}

struct Mat4 { double m[16]; };     // column-major like glTF
Mat4 mul(const Mat4& a, const Mat4& b) { Mat4 r; for (int c = 0; c < 4; c++) for (int rr = 0; rr < 4; rr++) { double s = 0; for (int k = 0; k < 4; k++) s += a.m[k * 4 + rr] * b.m[c * 4 + k]; r.m[c * 4 + rr] = s; } return r; }
Mat4 identity() { Mat4 r = {}; r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1; return r; }

// ---- the loaded scene: owns every byte RtxptSceneDesc points to ------------------------------------------------------------------------------------------
struct Primitive
{
    std::vector<F3> positions, normals; std::vector<float> uvs, tangents;     // uvs: 2 per vertex (empty: none), tangents: 4 per vertex (empty: compute)
    std::vector<uint32_t> indices; uint32_t material = 0;
};

} // namespace

struct rtxpt_host_scene
{
    RtxptSceneDesc desc = {};
    std::vector<RtxptInstanceData> instances; std::vector<RtxptGeometryData> geometries; std::vector<RtxptSubInstanceData> subInstances;
    std::vector<RtxptMaterialData> materials; std::vector<RtxptBufferDesc> buffers; std::vector<RtxptTextureDesc> textures; std::vector<RtxptLightDesc> lights;
    std::vector<std::vector<uint8_t>> blobs;            // index / vertex buffers and texture mips
    std::vector<RtxptGltfCamera> cameras;
    uint32_t triangleCount = 0;
    std::vector<bool> alphaTested, excludeFromNEE, skipRender;       // per material, beside `materials`
    RtxptSceneFileInfo info = {};                                    // .scene.json extras (environment light, settings); zero for plain glTF
};

namespace {

struct Accessor { const uint8_t* data; size_t stride; uint32_t count, componentType, components; bool normalized; };

struct Loader
{
    std::string baseDir; JValue root; std::vector<std::vector<uint8_t>> bufferData; std::vector<uint8_t> glbBin;
    rtxpt_host_scene* out = nullptr;
    std::map<std::pair<std::string, int>, uint32_t> textureSlot;        // (image source, sRGB) -> RtxptTextureDesc index
    std::string materialsDir, sceneMaterialsDir, modelName;     // RTXPT material overrides (Assets/Materials[/<scene>]); empty: none
    std::string mediaDir;                                       // what the texture paths of material files are relative to (Assets/)
    uint32_t overriddenMaterials = 0, materialBase = 0, bufferBase = 0;      // where this model's materials / buffers start in the shared scene

    // MaterialsBaker::Load search order (MaterialsBaker.cpp:707-747): scene-specialised folder first, then the shared one; <model>.<name> before <name>
    bool findMaterialFile(const std::string& name, std::string& text)
    {
        if (name.empty()) return false;
        const std::string cands[4] = { sceneMaterialsDir.empty() ? std::string() : sceneMaterialsDir + modelName + "." + name + ".material.json",
                                       sceneMaterialsDir.empty() ? std::string() : sceneMaterialsDir + name + ".material.json",
                                       materialsDir.empty() ? std::string() : materialsDir + modelName + "." + name + ".material.json",
                                       materialsDir.empty() ? std::string() : materialsDir + name + ".material.json" };
        for (const std::string& c : cands)
        {
            if (c.empty()) continue;
            FILE* f = fopen(c.c_str(), "rb"); if (!f) continue;
            fclose(f); const std::vector<uint8_t> d = readFile(c); text.assign(d.begin(), d.end()); return true;
        }
        return false;
    }

    const JValue& arrayItem(const char* name, int index)
    {
        const JValue* a = root.find(name);
        if (!a || index < 0 || size_t(index) >= a->size()) failf("glTF: %s[%d] does not exist", name, index);
        return a->arr[size_t(index)];
    }
    Accessor accessor(int index)
    {
        const JValue& a = arrayItem("accessors", index);
        if (a.find("sparse")) failf("glTF: sparse accessors are not supported");
        Accessor r; r.count = uint32_t(a.integer("count", 0)); r.componentType = uint32_t(a.integer("componentType", 0)); r.normalized = a.find("normalized") && a.at("normalized").b;
        const std::string type = a.string("type");
        r.components = type == "SCALAR" ? 1 : type == "VEC2" ? 2 : type == "VEC3" ? 3 : type == "VEC4" ? 4 : type == "MAT4" ? 16 : 0;
        if (!r.components) failf("glTF: accessor type '%s' is not supported", type.c_str());
        const size_t compSize = (r.componentType == 5120 || r.componentType == 5121) ? 1 : (r.componentType == 5122 || r.componentType == 5123) ? 2 : (r.componentType == 5125 || r.componentType == 5126) ? 4 : 0;
        if (!compSize) failf("glTF: component type %u is not supported", r.componentType);
        const int bvIndex = a.integer("bufferView", -1); if (bvIndex < 0) failf("glTF: accessor %d has no bufferView", index);
        const JValue& bv = arrayItem("bufferViews", bvIndex);
        const int bufIndex = bv.integer("buffer", -1); if (bufIndex < 0 || size_t(bufIndex) >= bufferData.size()) failf("glTF: bufferView %d references a missing buffer", bvIndex);
        // offsets, stride and lengths come from an untrusted file: finite, non-negative, inside the bufferView, and checked without wrapping
        const std::vector<uint8_t>& buf = bufferData[size_t(bufIndex)];
        auto field = [&](const JValue& o, const char* key, double def) { const double v = o.number(key, def); if (!(v >= 0.0) || v > double(buf.size()) || v != std::floor(v)) failf("glTF: accessor %d: '%s' is not a valid byte count", index, key); return size_t(v); };
        const size_t viewOffset = field(bv, "byteOffset", 0), viewLength = field(bv, "byteLength", double(buf.size() - std::min(buf.size(), viewOffset))), accOffset = field(a, "byteOffset", 0);
        r.stride = field(bv, "byteStride", 0); if (!r.stride) r.stride = compSize * r.components;
        if (r.stride < compSize * r.components || r.stride > 252 + compSize * r.components) failf("glTF: accessor %d: byteStride %zu is out of range", index, r.stride);
        if (viewOffset > buf.size() || viewLength > buf.size() - viewOffset) failf("glTF: bufferView %d lies outside buffer %d", bvIndex, bufIndex);
        const size_t elem = compSize * r.components;
        if (r.count)
        {   // last byte read = accOffset + stride * (count - 1) + elem, within the view
            if (accOffset > viewLength || elem > viewLength - accOffset || (r.count - 1) > (viewLength - accOffset - elem) / r.stride) failf("glTF: accessor %d reads past the end of bufferView %d", index, bvIndex);
        }
        const size_t offset = viewOffset + accOffset;
        r.data = buf.data() + offset;
        return r;
    }
    static float component(const Accessor& a, uint32_t element, uint32_t c)
    {
        const uint8_t* p = a.data + a.stride * element;
        switch (a.componentType)
        {
        case 5126: { float v; memcpy(&v, p + 4 * c, 4); return v; }
        case 5121: { const uint8_t v = p[c]; return a.normalized ? float(v) / 255.0f : float(v); }
        case 5123: { uint16_t v; memcpy(&v, p + 2 * c, 2); return a.normalized ? float(v) / 65535.0f : float(v); }
        case 5120: { const int8_t v = int8_t(p[c]); return a.normalized ? std::max(float(v) / 127.0f, -1.0f) : float(v); }
        case 5122: { int16_t v; memcpy(&v, p + 2 * c, 2); return a.normalized ? std::max(float(v) / 32767.0f, -1.0f) : float(v); }
        case 5125: { uint32_t v; memcpy(&v, p + 4 * c, 4); return float(v); }
        }
        return 0;
    }
    static uint32_t indexAt(const Accessor& a, uint32_t element)
    {
        const uint8_t* p = a.data + a.stride * element;
        switch (a.componentType) { case 5121: return p[0]; case 5123: { uint16_t v; memcpy(&v, p, 2); return v; } case 5125: { uint32_t v; memcpy(&v, p, 4); return v; } }
        failf("glTF: index component type %u is not supported", a.componentType);
    }

    // ---- textures: one RtxptTextureDesc per (source, colour space) in order of first use by the materials ----------------------------------------------
    // PNG files get a generated mip chain; DDS files (BC1-5, BC7, RGBA8) are decoded to RGBA8 with the mips they carry (dds.cpp)
    uint32_t registerTexture(const std::string& key, bool srgb, const std::vector<uint8_t>& bytes, const std::string& name)
    {
        std::vector<std::vector<uint8_t>> mips; uint32_t w, h; uint32_t bcFormat = 0;
        DdsBlocks raw;
        if (g_keepBlockCompression && extractDdsBlocks(bytes.data(), bytes.size(), name.c_str(), raw))
        {   // stays block-compressed on the device: the texture units decode it on fetch (include/rtxpt_b200.h RTXPT_FORMAT_BC*)
            w = raw.width; h = raw.height; mips = std::move(raw.mips); bcFormat = raw.format + (srgb ? 1u : 0u);
        }
        else if (bytes.size() >= 4 && memcmp(bytes.data(), "DDS ", 4) == 0)
        {
            DdsImage dds = decodeDds(bytes.data(), bytes.size(), name.c_str());
            w = dds.width; h = dds.height; mips = std::move(dds.mips);
        }
        else if (bytes.size() >= 2 && bytes[0] == 0xFF && bytes[1] == 0xD8)
        {   // image/jpeg, the other image format glTF 2.0 allows (jpeg.cpp)
            JpegImage j = decodeJpeg(bytes.data(), bytes.size(), name.c_str());
            Image img; img.w = j.w; img.h = j.h; img.rgba = std::move(j.rgba);
            std::vector<std::pair<uint32_t, uint32_t>> dims;
            mips = makeMips(img, dims); w = img.w; h = img.h;
        }
        else
        {
            Image img = decodePng(bytes, name.c_str());
            std::vector<std::pair<uint32_t, uint32_t>> dims;
            mips = makeMips(img, dims); w = img.w; h = img.h;
        }
        if (mips.size() > 16) failf("texture '%s' is larger than 32768 texels per side", name.c_str());
        RtxptTextureDesc d = {};
        d.width = w; d.height = h; d.mipLevels = uint32_t(mips.size()); d.format = bcFormat ? bcFormat : (srgb ? RTXPT_FORMAT_RGBA8_SRGB : RTXPT_FORMAT_RGBA8_UNORM);
        for (size_t m = 0; m < mips.size(); m++) { out->blobs.push_back(std::move(mips[m])); d.mips[m] = out->blobs.back().data(); }
        const uint32_t slot = uint32_t(out->textures.size()); out->textures.push_back(d); textureSlot[std::make_pair(key, srgb ? 1 : 0)] = slot;
        return slot;
    }
    uint32_t packedTextureIndex(uint32_t slot) const
    {   // the packed index of MaterialsBaker (baseLOD << 24 | mipLevels << 16 | bindless index)
        const RtxptTextureDesc& d = out->textures[slot];
        const uint32_t baseLod = uint32_t(std::log2(float(d.width) * float(d.height)) + 0.5f);       // MaterialsBaker.cpp:499-501
        return (baseLod << 24) | (d.mipLevels << 16) | slot;
    }
    uint32_t imageTexture(int imgIndex, bool srgb, bool searchForDds)
    {
        const JValue& im = arrayItem("images", imgIndex);
        std::vector<uint8_t> bytes; std::string name = im.string("uri", "<bufferView image>"), key = "image:" + std::to_string(imgIndex);
        const bool isFile = im.find("uri") && name.rfind("data:", 0) != 0;
        if (isFile)
        {
            std::string path = uriToPath(name, baseDir);
            if (searchForDds) path = preferDdsSibling(path, false);
            key = "file:" + path; name = path;
        }
        auto it = textureSlot.find(std::make_pair(key, srgb ? 1 : 0));
        if (it != textureSlot.end()) return it->second;
        if (isFile) bytes = readFile(name);
        else if (im.find("uri")) { bytes = resolveUri(im.at("uri").str, baseDir); name = "<data URI>"; }
        else
        {
            const JValue& bv = arrayItem("bufferViews", im.integer("bufferView", -1));
            const std::vector<uint8_t>& buf = bufferData.at(size_t(bv.integer("buffer", 0)));
            const size_t o = size_t(bv.number("byteOffset", 0)), n = size_t(bv.number("byteLength", 0));
            if (o + n > buf.size()) failf("glTF: image bufferView out of range");
            bytes.assign(buf.begin() + o, buf.begin() + o + n);
        }
        return registerTexture(key, srgb, bytes, name);
    }
    uint32_t textureInfo(const JValue* texRef, bool srgb)
    {   // 0xFFFFFFFF when absent
        if (!texRef) return 0xFFFFFFFFu;
        const int texIndex = texRef->integer("index", -1); if (texIndex < 0) return 0xFFFFFFFFu;
        if (texRef->integer("texCoord", 0) != 0) failf("glTF: only TEXCOORD_0 is supported for material textures");
        const JValue& tex = arrayItem("textures", texIndex);
        // an MSFT_texture_dds image wins over the plain source (GltfImporter.cpp:853-860); a plain file source is swapped for its .dds sibling when one exists
        int ddsIndex = -1;
        if (const JValue* e = tex.find("extensions")) if (const JValue* d = e->find("MSFT_texture_dds")) ddsIndex = d->integer("source", -1);
        const int imgIndex = tex.integer("source", -1);
        if (ddsIndex < 0 && imgIndex < 0) failf("glTF: texture %d has neither a PNG nor a DDS source", texIndex);
        return packedTextureIndex(ddsIndex >= 0 ? imageTexture(ddsIndex, srgb, false) : imageTexture(imgIndex, srgb, true));
    }
    // a texture named by an RTXPT material file: path relative to the media folder, .png swapped for a .dds sibling (MaterialsBaker.cpp:159-195)
    uint32_t materialFileTexture(const char* localPath, bool srgb)
    {
        if (mediaDir.empty() || !localPath[0]) return 0xFFFFFFFFu;
        const std::string path = preferDdsSibling(mediaDir + localPath, true);
        const std::string key = "file:" + path;
        auto it = textureSlot.find(std::make_pair(key, srgb ? 1 : 0));
        if (it != textureSlot.end()) return packedTextureIndex(it->second);
        if (!fileExists(path)) return 0xFFFFFFFFu;
        return packedTextureIndex(registerTexture(key, srgb, readFile(path), path));
    }

    void loadMaterials()
    {
        const JValue* mats = root.find("materials");
        const size_t n = mats ? mats->size() : 0;
        materialBase = uint32_t(out->materials.size());
        for (size_t i = 0; i <= n; i++)
        {   // one extra record at the end: the glTF default material for primitives that name none
            static const JValue empty;
            const JValue& m = (i < n) ? mats->arr[i] : empty;
            RtxptMaterialData d = {};
            uint32_t flags = 0;
            const JValue* pbr = m.find("pbrMetallicRoughness");
            float base[4] = { 1, 1, 1, 1 }; float metal = 1.0f, rough = 1.0f;
            if (i == n) { metal = 0.0f; }        // untextured grey for geometry without a material (Donut creates an equivalent default)
            const JValue* baseTex = nullptr; const JValue* ormTex = nullptr;
            if (pbr)
            {
                if (const JValue* f = pbr->find("baseColorFactor")) for (size_t k = 0; k < 4 && k < f->size(); k++) base[k] = float(f->arr[k].num);
                metal = float(pbr->number("metallicFactor", 1.0)); rough = float(pbr->number("roughnessFactor", 1.0));
                baseTex = pbr->find("baseColorTexture"); ormTex = pbr->find("metallicRoughnessTexture");
            }
            const JValue* ext = m.find("extensions");
            if (ext && ext->find("KHR_materials_pbrSpecularGlossiness")) failf("glTF: material %zu uses KHR_materials_pbrSpecularGlossiness, which this loader does not convert", i);
            auto use = [&](uint32_t info, uint32_t bit) { if (info != 0xFFFFFFFFu) flags |= bit; return info; };
            d.BaseOrDiffuseTextureIndex = use(textureInfo(baseTex, true), RTXPT_MATFLAG_UseBaseOrDiffuseTexture);
            d.MetalRoughOrSpecularTextureIndex = use(textureInfo(ormTex, false), RTXPT_MATFLAG_UseMetalRoughOrSpecularTexture);
            const JValue* normalTex = m.find("normalTexture");
            d.NormalTextureIndex = use(textureInfo(normalTex, false), RTXPT_MATFLAG_UseNormalTexture);
            d.EmissiveTextureIndex = use(textureInfo(m.find("emissiveTexture"), true), RTXPT_MATFLAG_UseEmissiveTexture);
            d.TransmissionTextureIndex = 0xFFFFFFFFu; d.OcclusionTextureIndex = 0xFFFFFFFFu;
            // emissive: colour normalised by its largest component, intensity carries the magnitude (GltfImporter.cpp:957-963), times
            // KHR_materials_emissive_strength; FillData multiplies them back together
            float em[3] = { 0, 0, 0 };
            if (const JValue* f = m.find("emissiveFactor")) for (size_t k = 0; k < 3 && k < f->size(); k++) em[k] = float(f->arr[k].num);
            float intensity = std::max(em[0], std::max(em[1], em[2]));
            if (intensity > 0.f) { em[0] /= intensity; em[1] /= intensity; em[2] /= intensity; } else intensity = 1.f;
            if (ext) if (const JValue* es = ext->find("KHR_materials_emissive_strength")) intensity *= float(es->number("emissiveStrength", 1.0));
            float transmission = 0.0f; bool enableTransmission = false;
            if (ext) if (const JValue* tr = ext->find("KHR_materials_transmission")) { transmission = float(tr->number("transmissionFactor", 0.0)); enableTransmission = true; }
            float ior = 1.5f;
            if (ext) if (const JValue* io = ext->find("KHR_materials_ior")) ior = float(io->number("ior", 1.5));
            float volColor[3] = { 1, 1, 1 }, volDist = 3.4e38f; bool thin = true;
            if (ext) if (const JValue* vo = ext->find("KHR_materials_volume"))
            {   // the reference takes these from its .material.json files (VolumeAttenuation*, ThinSurface); a glTF volume with thickness marks a solid
                if (const JValue* c = vo->find("attenuationColor")) for (size_t k = 0; k < 3 && k < c->size(); k++) volColor[k] = float(c->arr[k].num);
                volDist = float(std::min(vo->number("attenuationDistance", 3.4e38), 3.4e38));
                thin = !(vo->number("thicknessFactor", 0.0) > 0.0);
            }
            const std::string alphaMode = m.string("alphaMode", "OPAQUE");
            if (!enableTransmission || thin) flags |= RTXPT_MATFLAG_ThinSurface;        // MaterialsBaker.cpp:543-544
            flags |= RTXPT_MATFLAG_PSDExclude;       // PTMaterialBase default for a material without a .material.json (MaterialsBaker.h:173)
            if (const JValue* extras = m.find("extras")) flags |= (uint32_t(std::min(extras->integer("nestedPriority", 0), 14)) & 0xFu) << RTXPT_MATFLAG_NestedPriorityShift;
            d.Flags = flags;
            d.BaseOrDiffuseColor[0] = base[0]; d.BaseOrDiffuseColor[1] = base[1]; d.BaseOrDiffuseColor[2] = base[2];
            d.EmissiveColor[0] = em[0] * intensity; d.EmissiveColor[1] = em[1] * intensity; d.EmissiveColor[2] = em[2] * intensity;
            d.Roughness = rough; d.Metalness = metal; d.NormalTextureScale = normalTex ? float(normalTex->number("scale", 1.0)) : 1.0f;
            d.TransmissionFactor = enableTransmission ? transmission : 0.0f; d.DiffuseTransmissionFactor = 0.0f;
            d.Opacity = base[3]; d.AlphaCutoff = float(m.number("alphaCutoff", 0.5)); d.IoR = ior;
            d.VolumeAttenuationColor[0] = volColor[0]; d.VolumeAttenuationColor[1] = volColor[1]; d.VolumeAttenuationColor[2] = volColor[2];
            d.VolumeAttenuationDistance = volDist; d.ShadowNoLFadeout = 0.0f;
            d._padding0 = 42; d._padding1 = 42.0f;
            bool alpha = alphaMode == "MASK", noNEE = false, skip = false;
            std::string overrideText;
            if (i < n && findMaterialFile(m.string("name"), overrideText))
            {   // the RTXPT material file wins over the glTF material (MaterialsBaker.cpp:868-917), textures included: a slot the file enables takes
                // the file's texture (PNG or DDS under the media folder); only when that file is absent does the slot keep the glTF texture
                JParser jp{ overrideText.data(), overrideText.data() + overrideText.size() };
                RtxptMaterialJsonInfo info; materialFromJson(jp.parse(), info);
                const uint32_t texIdx[5] = { d.BaseOrDiffuseTextureIndex, d.MetalRoughOrSpecularTextureIndex, d.NormalTextureIndex, d.EmissiveTextureIndex, 0xFFFFFFFFu };
                const uint32_t texBit[5] = { RTXPT_MATFLAG_UseBaseOrDiffuseTexture, RTXPT_MATFLAG_UseMetalRoughOrSpecularTexture, RTXPT_MATFLAG_UseNormalTexture, RTXPT_MATFLAG_UseEmissiveTexture, RTXPT_MATFLAG_UseTransmissionTexture };
                d = info.data;
                uint32_t* slots[5] = { &d.BaseOrDiffuseTextureIndex, &d.MetalRoughOrSpecularTextureIndex, &d.NormalTextureIndex, &d.EmissiveTextureIndex, &d.TransmissionTextureIndex };
                for (int t = 0; t < 5; t++)
                {
                    if (!info.textureEnabled[t]) continue;
                    uint32_t idx = materialFileTexture(info.texturePath[t], info.textureSRGB[t] != 0);
                    if (idx == 0xFFFFFFFFu) idx = texIdx[t];
                    if (idx != 0xFFFFFFFFu) { *slots[t] = idx; d.Flags |= texBit[t]; }
                }
                alpha = info.enableAlphaTesting != 0; noNEE = info.excludeFromNEE != 0; skip = info.skipRender != 0;
                overriddenMaterials++;
            }
            out->materials.push_back(d);
            out->alphaTested.push_back(alpha && d.BaseOrDiffuseTextureIndex != 0xFFFFFFFFu);
            out->excludeFromNEE.push_back(noNEE); out->skipRender.push_back(skip);
        }
    }

    // per-vertex tangents when the asset has none (GltfImporter.cpp:1331-1421): per-triangle tangent/bitangent from the UV gradients, summed
    // per vertex (vertex slot 0 of every triangle first, then slot 1, then slot 2), normalised, handedness from the bitangent
    static std::vector<float> computeTangents(const Primitive& pr)
    {
        const size_t nv = pr.positions.size(), nt = pr.indices.size() / 3;
        std::vector<F3> triT(nt), triB(nt), T(nv, F3{ 0, 0, 0 }), B(nv, F3{ 0, 0, 0 });
        for (size_t t = 0; t < nt; t++)
        {
            const uint32_t i0 = pr.indices[3 * t], i1 = pr.indices[3 * t + 1], i2 = pr.indices[3 * t + 2];
            const F3 dPds = pr.positions[i1] - pr.positions[i0], dPdt = pr.positions[i2] - pr.positions[i0];
            const float ds0 = pr.uvs[2 * i1] - pr.uvs[2 * i0], ds1 = pr.uvs[2 * i1 + 1] - pr.uvs[2 * i0 + 1];
            const float dt0 = pr.uvs[2 * i2] - pr.uvs[2 * i0], dt1 = pr.uvs[2 * i2 + 1] - pr.uvs[2 * i0 + 1];
            const float det = ds0 * dt1 - ds1 * dt0;
            const float r = 1.0f / det;
            F3 tg = (dPds * dt1 - dPdt * ds1) * r, bt = (dPdt * ds0 - dPds * dt0) * r;
            const float tl = norm(tg), bl = norm(bt);
            const bool ok = finitef(tl) && finitef(bl) && tl > 0 && bl > 0;
            triT[t] = ok ? F3{ tg.x / tl, tg.y / tl, tg.z / tl } : F3{ 0, 0, 0 };
            triB[t] = ok ? F3{ bt.x / bl, bt.y / bl, bt.z / bl } : F3{ 0, 0, 0 };
        }
        for (int k = 0; k < 3; k++) for (size_t t = 0; t < nt; t++) { const uint32_t v = pr.indices[3 * t + k]; T[v] = T[v] + triT[t]; B[v] = B[v] + triB[t]; }
        std::vector<float> outT(nv * 4);
        for (size_t v = 0; v < nv; v++)
        {
            const float tl = norm(T[v]), bl = norm(B[v]); const bool ok = tl > 0 && bl > 0;
            const F3 Tn = ok ? F3{ T[v].x / tl, T[v].y / tl, T[v].z / tl } : F3{ 0, 0, 0 }, Bn = ok ? F3{ B[v].x / bl, B[v].y / bl, B[v].z / bl } : F3{ 0, 0, 0 };
            const F3 c = cross(pr.normals[v], Tn);
            const float d = (c.x * Bn.x + c.y * Bn.y) + c.z * Bn.z;
            outT[4 * v] = Tn.x; outT[4 * v + 1] = Tn.y; outT[4 * v + 2] = Tn.z; outT[4 * v + 3] = ok ? (d > 0 ? -1.0f : 1.0f) : 0.0f;
        }
        return outT;
    }

    std::vector<std::vector<Primitive>> meshes;
    void loadMeshes()
    {
        const JValue* ms = root.find("meshes"); if (!ms) return;
        const uint32_t defaultMaterial = uint32_t(out->materials.size() - 1);      // the extra record loadMaterials appended for this model
        for (size_t mi = 0; mi < ms->size(); mi++)
        {
            std::vector<Primitive> prims;
            const JValue& plist = ms->arr[mi].at("primitives");
            for (size_t pi = 0; pi < plist.size(); pi++)
            {
                const JValue& p = plist.arr[pi];
                if (p.integer("mode", 4) != 4) failf("glTF: mesh %zu primitive %zu is not a triangle list", mi, pi);
                if (const JValue* e = p.find("extensions")) if (e->find("KHR_draco_mesh_compression")) failf("glTF: Draco-compressed meshes are not supported");
                const JValue& attrs = p.at("attributes");
                Primitive pr;
                const Accessor pos = accessor(attrs.at("POSITION").type == JValue::Number ? int(attrs.at("POSITION").num) : -1);
                pr.positions.resize(pos.count);
                for (uint32_t v = 0; v < pos.count; v++) pr.positions[v] = { component(pos, v, 0), component(pos, v, 1), component(pos, v, 2) };
                if (const JValue* ia = p.find("indices")) { const Accessor idx = accessor(int(ia->num)); pr.indices.resize(idx.count); for (uint32_t k = 0; k < idx.count; k++) pr.indices[k] = indexAt(idx, k); }
                else { pr.indices.resize(pos.count); for (uint32_t k = 0; k < pos.count; k++) pr.indices[k] = k; }
                if (pr.indices.size() % 3) failf("glTF: mesh %zu primitive %zu has an index count that is not a multiple of 3", mi, pi);
                for (uint32_t ix : pr.indices) if (ix >= pos.count) failf("glTF: mesh %zu primitive %zu indexes past its vertices", mi, pi);
                if (const JValue* na = attrs.find("NORMAL"))
                {
                    const Accessor n = accessor(int(na->num)); if (n.count != pos.count) failf("glTF: NORMAL count differs from POSITION count");
                    pr.normals.resize(pos.count); for (uint32_t v = 0; v < pos.count; v++) pr.normals[v] = { component(n, v, 0), component(n, v, 1), component(n, v, 2) };
                }
                else
                {   // area-weighted vertex normals
                    pr.normals.assign(pos.count, F3{ 0, 0, 0 });
                    for (size_t t = 0; t + 2 < pr.indices.size(); t += 3)
                    {
                        const F3 fn = cross(pr.positions[pr.indices[t + 1]] - pr.positions[pr.indices[t]], pr.positions[pr.indices[t + 2]] - pr.positions[pr.indices[t]]);
                        for (int k = 0; k < 3; k++) pr.normals[pr.indices[t + k]] = pr.normals[pr.indices[t + k]] + fn;
                    }
                    for (F3& n : pr.normals) { const float l = norm(n); n = l > 0 ? F3{ n.x / l, n.y / l, n.z / l } : F3{ 0, 1, 0 }; }
                }
                if (const JValue* ta = attrs.find("TEXCOORD_0"))
                {
                    const Accessor t = accessor(int(ta->num)); if (t.count != pos.count) failf("glTF: TEXCOORD_0 count differs from POSITION count");
                    pr.uvs.resize(size_t(pos.count) * 2); for (uint32_t v = 0; v < pos.count; v++) { pr.uvs[2 * v] = component(t, v, 0); pr.uvs[2 * v + 1] = component(t, v, 1); }
                }
                if (const JValue* ga = attrs.find("TANGENT"))
                {
                    const Accessor t = accessor(int(ga->num)); if (t.count != pos.count || t.components != 4) failf("glTF: TANGENT must be VEC4 per vertex");
                    pr.tangents.resize(size_t(pos.count) * 4); for (uint32_t v = 0; v < pos.count; v++) for (uint32_t c = 0; c < 4; c++) pr.tangents[4 * v + c] = component(t, v, c);
                }
                const int mat = p.integer("material", -1);
                pr.material = (mat >= 0 && materialBase + uint32_t(mat) < defaultMaterial) ? materialBase + uint32_t(mat) : defaultMaterial;
                prims.push_back(std::move(pr));
            }
            meshes.push_back(std::move(prims));
        }
    }

    std::vector<uint32_t> meshFirstGeometry; std::vector<bool> meshHasUv;
    void buildBuffers()
    {
        bufferBase = uint32_t(out->buffers.size());
        for (size_t mi = 0; mi < meshes.size(); mi++)
        {
            std::vector<Primitive>& prims = meshes[mi];
            meshFirstGeometry.push_back(uint32_t(out->geometries.size()));
            size_t nv = 0, ni = 0; bool hasUv = true;
            for (const Primitive& p : prims) { nv += p.positions.size(); ni += p.indices.size(); hasUv = hasUv && !p.uvs.empty(); }
            meshHasUv.push_back(hasUv);
            std::vector<uint8_t> iblob(ni * 4), vblob(nv * 28);
            const size_t offPos = 0, offUv = nv * 12, offNrm = offUv + nv * 8, offTan = offNrm + nv * 4;
            size_t v0 = 0, i0 = 0;
            for (Primitive& p : prims)
            {
                const size_t n = p.positions.size();
                memcpy(&iblob[i0 * 4], p.indices.data(), p.indices.size() * 4);
                std::vector<float> tangents = p.tangents;
                if (tangents.empty()) tangents = p.uvs.empty() ? std::vector<float>(n * 4, 0.0f) : computeTangents(p);
                for (size_t v = 0; v < n; v++)
                {
                    memcpy(&vblob[offPos + (v0 + v) * 12], &p.positions[v], 12);
                    if (hasUv) memcpy(&vblob[offUv + (v0 + v) * 8], &p.uvs[2 * v], 8);
                    const uint32_t pn = packSnorm8x3(p.normals[v]), pt = packSnorm8x4(F3{ tangents[4 * v], tangents[4 * v + 1], tangents[4 * v + 2] }, tangents[4 * v + 3]);
                    memcpy(&vblob[offNrm + (v0 + v) * 4], &pn, 4); memcpy(&vblob[offTan + (v0 + v) * 4], &pt, 4);
                }
                RtxptGeometryData g = {};
                g.numIndices = out->skipRender[p.material] ? 0u : uint32_t(p.indices.size()); g.numVertices = uint32_t(n);      // SkipRender materials keep their slot and draw nothing
                g.indexBufferIndex = int32_t(bufferBase + 2 * mi); g.indexOffset = uint32_t(i0 * 4); g.vertexBufferIndex = int32_t(bufferBase + 2 * mi + 1);
                g.positionOffset = uint32_t(offPos + v0 * 12); g.prevPositionOffset = 0xFFFFFFFFu;
                g.texCoord1Offset = hasUv ? uint32_t(offUv + v0 * 8) : 0xFFFFFFFFu; g.texCoord2Offset = 0xFFFFFFFFu;
                g.normalOffset = uint32_t(offNrm + v0 * 4); g.tangentOffset = uint32_t(offTan + v0 * 4); g.curveRadiusOffset = 0xFFFFFFFFu;
                g.materialIndex = p.material;
                out->geometries.push_back(g);
                v0 += n; i0 += p.indices.size();
            }
            out->blobs.push_back(std::move(iblob)); out->buffers.push_back({ out->blobs.back().data(), out->blobs.back().size() });
            out->blobs.push_back(std::move(vblob)); out->buffers.push_back({ out->blobs.back().data(), out->blobs.back().size() });
        }
    }

    static Mat4 nodeLocal(const JValue& n)
    {
        Mat4 m = identity();
        if (const JValue* mat = n.find("matrix")) { for (size_t k = 0; k < 16 && k < mat->size(); k++) m.m[k] = mat->arr[k].num; return m; }
        double t[3] = { 0, 0, 0 }, r[4] = { 0, 0, 0, 1 }, s[3] = { 1, 1, 1 };
        if (const JValue* v = n.find("translation")) for (size_t k = 0; k < 3 && k < v->size(); k++) t[k] = v->arr[k].num;
        if (const JValue* v = n.find("rotation")) for (size_t k = 0; k < 4 && k < v->size(); k++) r[k] = v->arr[k].num;
        if (const JValue* v = n.find("scale")) for (size_t k = 0; k < 3 && k < v->size(); k++) s[k] = v->arr[k].num;
        const double x = r[0], y = r[1], z = r[2], w = r[3];
        const double R[9] = { 1 - 2 * (y * y + z * z), 2 * (x * y + z * w), 2 * (x * z - y * w),      // column 0
                              2 * (x * y - z * w), 1 - 2 * (x * x + z * z), 2 * (y * z + x * w),      // column 1
                              2 * (x * z + y * w), 2 * (y * z - x * w), 1 - 2 * (x * x + y * y) };    // column 2
        for (int c = 0; c < 3; c++) for (int rr = 0; rr < 3; rr++) m.m[c * 4 + rr] = R[c * 3 + rr] * s[c];
        m.m[12] = t[0]; m.m[13] = t[1]; m.m[14] = t[2];
        return m;
    }
    void visitNode(int index, const Mat4& parent, int depth)
    {
        if (depth > 256) failf("glTF: node hierarchy is deeper than 256 levels (cycle?)");
        const JValue& n = arrayItem("nodes", index);
        const Mat4 world = mul(parent, nodeLocal(n));
        const int meshIndex = n.integer("mesh", -1);
        if (meshIndex >= 0)
        {
            if (size_t(meshIndex) >= meshes.size()) failf("glTF: node %d references a missing mesh", index);
            RtxptInstanceData inst = {};
            inst.flags = 0; inst.firstGeometryInstanceIndex = uint32_t(out->subInstances.size()); inst.firstGeometryIndex = meshFirstGeometry[size_t(meshIndex)];
            inst.numGeometries = uint32_t(meshes[size_t(meshIndex)].size());
            for (int r = 0; r < 3; r++) for (int c = 0; c < 4; c++) inst.transform[r * 4 + c] = inst.prevTransform[r * 4 + c] = float(world.m[c * 4 + r]);
            for (uint32_t k = 0; k < inst.numGeometries; k++)
            {
                const uint32_t gi = inst.firstGeometryIndex + k; const RtxptGeometryData& g = out->geometries[gi];
                const RtxptMaterialData& m = out->materials[g.materialIndex];
                RtxptSubInstanceData s = {};
                uint32_t fl = 0; float cutoff = 0.0f;
                if (out->alphaTested[g.materialIndex] && g.texCoord1Offset != 0xFFFFFFFFu) { fl |= RTXPT_SUBINST_FLAG_ALPHA_TESTED | (m.BaseOrDiffuseTextureIndex & 0xFFFFu); cutoff = m.AlphaCutoff; }
                fl |= uint32_t(int(std::min(std::max(cutoff, 0.0f), 1.0f) * 255.0f + 0.5f)) << 24;
                if (out->excludeFromNEE[g.materialIndex]) fl |= RTXPT_SUBINST_FLAG_EXCLUDE_FROM_NEE;
                s.FlagsAndAlphaInfo = fl;
                s.GlobalGeometryIndex_PTMaterialDataIndex = (gi << 16) | g.materialIndex;
                s.EmissiveLightMappingOffset = 0xFFFFFFFFu; s.AnalyticProxyLightIndex = 0xFFFFFFFFu;
                s.IndexBufferIndex_VertexBufferIndex = (uint32_t(g.indexBufferIndex) << 16) | uint32_t(g.vertexBufferIndex);
                s.IndexOffset = g.indexOffset; s.TexCoord1Offset = g.texCoord1Offset;
                out->subInstances.push_back(s);
                out->triangleCount += g.numIndices / 3;         // instanced triangles, what the BVH will hold
            }
            out->instances.push_back(inst);
        }
        const int camIndex = n.integer("camera", -1);
        if (camIndex >= 0)
        {
            const JValue& cam = arrayItem("cameras", camIndex);
            if (const JValue* persp = cam.find("perspective"))
            {   // glTF cameras look down -Z with +Y up in node space
                RtxptGltfCamera c = {};
                c.position[0] = float(world.m[12]); c.position[1] = float(world.m[13]); c.position[2] = float(world.m[14]);
                for (int k = 0; k < 3; k++) { c.direction[k] = float(-world.m[8 + k]); c.up[k] = float(world.m[4 + k]); }
                c.yfov = float(persp->number("yfov", 1.0)); c.znear = float(persp->number("znear", 0.1)); c.zfar = float(persp->number("zfar", 1e7));
                c.aspectRatio = float(persp->number("aspectRatio", 0.0));
                out->cameras.push_back(c);
            }
        }
        if (const JValue* ext = n.find("extensions")) if (const JValue* lp = ext->find("KHR_lights_punctual"))
        {
            const JValue* rootExt = root.find("extensions"); const JValue* lights = rootExt ? rootExt->find("KHR_lights_punctual") : nullptr;
            const JValue* list = lights ? lights->find("lights") : nullptr; const int li = lp->integer("light", -1);
            if (!list || li < 0 || size_t(li) >= list->size()) failf("glTF: node %d references a missing KHR_lights_punctual light", index);
            const JValue& L = list->arr[size_t(li)]; const std::string type = L.string("type");
            if (type == "point" || type == "spot")
            {   // directional lights are folded into the environment map by the reference (EnvMapBaker); not part of the light list
                RtxptLightDesc d = {};
                d.type = type == "spot" ? RTXPT_LIGHT_SPOT : RTXPT_LIGHT_POINT;
                d.position[0] = float(world.m[12]); d.position[1] = float(world.m[13]); d.position[2] = float(world.m[14]);
                for (int k = 0; k < 3; k++) d.direction[k] = float(-world.m[8 + k]);
                d.color[0] = d.color[1] = d.color[2] = 1.0f;
                if (const JValue* c = L.find("color")) for (size_t k = 0; k < 3 && k < c->size(); k++) d.color[k] = float(c->arr[k].num);
                d.intensity = float(L.number("intensity", 1.0));
                // the sphere radius is RTXPT's extension of Donut's lights (scene.json "radius"); glTF has no such field, so it is read from extras
                if (const JValue* ex = L.find("extras")) d.radius = float(ex->number("radius", 0.0));
                if (const JValue* sp = L.find("spot")) { d.innerAngle = float(sp->number("innerConeAngle", 0.0) * 180.0 / 3.14159265358979323846); d.outerAngle = float(sp->number("outerConeAngle", 0.7853981633974483) * 180.0 / 3.14159265358979323846); }
                out->lights.push_back(d);
            }
        }
        if (const JValue* ch = n.find("children")) for (const JValue& c : ch->arr) visitNode(int(c.num), world, depth + 1);
    }

    // parses one model file and appends its materials, textures, meshes and buffers to the shared scene; instantiate() then places it
    void open(const std::string& path)
    {
        const size_t slash = path.find_last_of("/\\");
        baseDir = (slash == std::string::npos) ? std::string() : path.substr(0, slash + 1);
        modelName = path.substr(slash == std::string::npos ? 0 : slash + 1); { const size_t dot = modelName.find_last_of('.'); if (dot != std::string::npos) modelName.erase(dot); }
        std::vector<uint8_t> file = readFile(path);
        std::string jsonText;
        if (file.size() >= 12 && !memcmp(file.data(), "glTF", 4))
        {   // GLB container: header, JSON chunk, optional BIN chunk
            uint32_t length; memcpy(&length, &file[8], 4);
            size_t off = 12;
            while (off + 8 <= file.size() && off < length)
            {
                uint32_t clen, ctype; memcpy(&clen, &file[off], 4); memcpy(&ctype, &file[off + 4], 4);
                if (off + 8 + clen > file.size()) failf("GLB: truncated chunk");
                if (ctype == 0x4E4F534Au) jsonText.assign(reinterpret_cast<const char*>(&file[off + 8]), clen);
                else if (ctype == 0x004E4942u) glbBin.assign(file.begin() + off + 8, file.begin() + off + 8 + clen);
                off += 8 + size_t(clen);
            }
        }
        else jsonText.assign(file.begin(), file.end());
        JParser jp{ jsonText.data(), jsonText.data() + jsonText.size() };
        root = jp.parse();
        if (root.type != JValue::Object) failf("glTF: top level is not an object");
        if (const JValue* req = root.find("extensionsRequired")) for (const JValue& e : req->arr)
            if (e.str != "KHR_lights_punctual" && e.str != "KHR_materials_transmission" && e.str != "KHR_materials_ior" && e.str != "KHR_materials_volume" && e.str != "KHR_materials_emissive_strength")
                failf("glTF: required extension '%s' is not supported", e.str.c_str());
        if (const JValue* bufs = root.find("buffers")) for (size_t i = 0; i < bufs->size(); i++)
        {
            const JValue& b = bufs->arr[i];
            if (b.find("uri")) bufferData.push_back(resolveUri(b.at("uri").str, baseDir));
            else if (i == 0 && !glbBin.empty()) bufferData.push_back(glbBin);
            else failf("glTF: buffer %zu has no uri", i);
        }
        loadMaterials(); loadMeshes(); buildBuffers();
    }
    // emits the instances (and lights / cameras) of the model's default scene under `parent`; may be called several times (scene-graph instancing)
    void instantiate(const Mat4& parent)
    {
        const JValue* scenes = root.find("scenes");
        if (scenes && scenes->size())
        {
            const JValue& sc = scenes->arr[size_t(std::min<int>(root.integer("scene", 0), int(scenes->size()) - 1))];
            if (const JValue* ns = sc.find("nodes")) for (const JValue& n : ns->arr) visitNode(int(n.num), parent, 0);
        }
        else if (const JValue* nodes = root.find("nodes"))
        {   // no scene: every root node (a node that is nobody's child)
            std::vector<bool> isChild(nodes->size(), false);
            for (const JValue& n : nodes->arr) if (const JValue* ch = n.find("children")) for (const JValue& c : ch->arr) if (size_t(c.num) < isChild.size()) isChild[size_t(c.num)] = true;
            for (size_t i = 0; i < nodes->size(); i++) if (!isChild[i]) visitNode(int(i), parent, 0);
        }
    }
};

void finalizeScene(rtxpt_host_scene* out, const char* what)
{
    if (out->instances.empty()) failf("'%s' contains no mesh instances", what);
    RtxptSceneDesc& d = out->desc;
    d.instances = out->instances.data(); d.instanceCount = uint32_t(out->instances.size());
    d.geometries = out->geometries.data(); d.geometryCount = uint32_t(out->geometries.size());
    d.subInstances = out->subInstances.data(); d.subInstanceCount = uint32_t(out->subInstances.size());
    d.materials = out->materials.data(); d.materialCount = uint32_t(out->materials.size());
    d.buffers = out->buffers.data(); d.bufferCount = uint32_t(out->buffers.size());
    d.textures = out->textures.data(); d.textureCount = uint32_t(out->textures.size());
    d.lights = out->lights.data(); d.lightCount = uint32_t(out->lights.size());
}

thread_local std::string g_loaderError;

} // namespace

extern "C" {

RTXPT_API int rtxpt_b200_load_gltf_ex(const char* path, const char* materialsDir, const char* sceneMaterialsDir, rtxpt_host_scene** outScene, uint32_t* outOverriddenMaterials)
{
    if (!path || !outScene) { g_loaderError = "null argument"; return RTXPT_ERR_INVALID_ARGUMENT; }
    *outScene = nullptr;
    std::unique_ptr<rtxpt_host_scene> scene(new rtxpt_host_scene());
    auto asDir = [](const char* d) { std::string s = d ? d : ""; if (!s.empty() && s.back() != '/' && s.back() != '\\') s += '/'; return s; };
    // ".../Assets/Materials/" -> ".../Assets/": the media folder the texture paths of material files are relative to (MaterialsBaker.cpp:857-860)
    auto parentDir = [](const std::string& dir) { if (dir.size() < 2) return std::string(); const size_t cut = dir.find_last_of("/\\", dir.size() - 2); return cut == std::string::npos ? std::string("./") : dir.substr(0, cut + 1); };
    try
    {
        Loader l; l.out = scene.get(); l.materialsDir = asDir(materialsDir); l.sceneMaterialsDir = asDir(sceneMaterialsDir); l.mediaDir = parentDir(l.materialsDir); l.open(path); l.instantiate(identity()); finalizeScene(scene.get(), path);
        if (outOverriddenMaterials) *outOverriddenMaterials = l.overriddenMaterials;
    }
    catch (const LoadError& e) { g_loaderError = e.msg; return RTXPT_ERR_INVALID_ARGUMENT; }
    catch (const std::exception& e) { g_loaderError = e.what(); return RTXPT_ERR_INVALID_ARGUMENT; }
    *outScene = scene.release();
    return RTXPT_OK;
}
RTXPT_API void rtxpt_b200_loader_keep_block_compression(int enable) { rtxpt_host::g_keepBlockCompression = enable != 0; }
RTXPT_API int rtxpt_b200_load_gltf(const char* path, rtxpt_host_scene** outScene) { return rtxpt_b200_load_gltf_ex(path, nullptr, nullptr, outScene, nullptr); }
RTXPT_API const char* rtxpt_b200_load_gltf_error(void) { return g_loaderError.c_str(); }
RTXPT_API const RtxptSceneDesc* rtxpt_b200_host_scene_desc(const rtxpt_host_scene* scene) { return scene ? &scene->desc : nullptr; }
RTXPT_API int rtxpt_b200_host_scene_cameras(const rtxpt_host_scene* scene, RtxptGltfCamera* outCameras, uint32_t* ioCount)
{
    if (!scene || !ioCount) return RTXPT_ERR_INVALID_ARGUMENT;
    const uint32_t n = uint32_t(scene->cameras.size());
    if (outCameras && *ioCount >= n && n) memcpy(outCameras, scene->cameras.data(), size_t(n) * sizeof(RtxptGltfCamera));
    *ioCount = n;
    return RTXPT_OK;
}
RTXPT_API uint32_t rtxpt_b200_host_scene_triangle_count(const rtxpt_host_scene* scene) { return scene ? scene->triangleCount : 0; }
RTXPT_API void rtxpt_b200_free_host_scene(rtxpt_host_scene* scene) { delete scene; }

} // extern "C"

// ---- RTXPT .scene.json ----------------------------------------------------------------------------------------------------------------------------------
namespace {

// Donut's JSON readers (src/core/json.cpp:200-262): a vector is taken from an array of exactly the right length, a lone number is broadcast, anything else
// keeps the default — e.g. a 3-element "rotation" leaves the identity quaternion
template <int N> void readVec(const JValue* v, double (&out)[N])
{
    if (!v) return;
    if (v->type == JValue::Array && v->size() == size_t(N)) { for (int k = 0; k < N; k++) out[k] = v->arr[size_t(k)].num; }
    else if (v->type == JValue::Number) { for (int k = 0; k < N; k++) out[k] = v->num; }
}
Mat4 trs(const double t[3], const double q[4], const double s[3])
{
    Mat4 m = identity();
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    const double R[9] = { 1 - 2 * (y * y + z * z), 2 * (x * y + z * w), 2 * (x * z - y * w),   2 * (x * y - z * w), 1 - 2 * (x * x + z * z), 2 * (y * z + x * w),
                          2 * (x * z + y * w), 2 * (y * z - x * w), 1 - 2 * (x * x + y * y) };
    for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) m.m[c * 4 + r] = R[c * 3 + r] * s[c];
    m.m[12] = t[0]; m.m[13] = t[1]; m.m[14] = t[2];
    return m;
}

struct SceneFileLoader
{
    rtxpt_host_scene* out = nullptr; std::string mediaDir, materialsDir, sceneMaterialsDir;
    std::vector<std::unique_ptr<Loader>> models;

    void visit(const JValue& list, const Mat4& parent, int depth)
    {
        if (list.type != JValue::Array) return;
        if (depth > 64) failf("scene.json: graph is deeper than 64 levels");
        for (const JValue& src : list.arr)
        {
            if (src.type != JValue::Object) continue;
            double t[3] = { 0, 0, 0 }, q[4] = { 0, 0, 0, 1 }, s[3] = { 1, 1, 1 };
            readVec(src.find("translation"), t);
            if (src.find("rotation")) readVec(src.find("rotation"), q);
            else if (const JValue* e = src.find("euler"))
            {   // rotationQuat(euler) = qZ * qY * qX (donut/core/math/quat.h:398-414)
                double eu[3] = { 0, 0, 0 }; readVec(e, eu);
                const double sx = std::sin(0.5 * eu[0]), cx = std::cos(0.5 * eu[0]), sy = std::sin(0.5 * eu[1]), cy = std::cos(0.5 * eu[1]), sz = std::sin(0.5 * eu[2]), cz = std::cos(0.5 * eu[2]);
                auto qmul = [](const double a[4], const double b[4], double r[4]) {      // xyzw
                    r[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1]; r[1] = a[3] * b[1] - a[0] * b[2] + a[1] * b[3] + a[2] * b[0];
                    r[2] = a[3] * b[2] + a[0] * b[1] - a[1] * b[0] + a[2] * b[3]; r[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2]; };
                const double qx[4] = { sx, 0, 0, cx }, qy[4] = { 0, sy, 0, cy }, qz[4] = { 0, 0, sz, cz }; double zy[4]; qmul(qz, qy, zy); qmul(zy, qx, q);
            }
            readVec(src.find("scaling"), s);
            const Mat4 world = mul(parent, trs(t, q, s));
            if (const JValue* mj = src.find("model"))
            {
                const int mi = (mj->type == JValue::Number) ? int(mj->num) : -1;
                if (mi < 0 || size_t(mi) >= models.size()) failf("scene.json: node '%s' references model %d, which is not in the model array", src.string("name").c_str(), mi);
                models[size_t(mi)]->instantiate(world);
            }
            if (const JValue* ch = src.find("children")) visit(*ch, world, depth + 1);
            const std::string type = src.string("type");
            if (type == "PointLight" || type == "SpotLight")
            {
                RtxptLightDesc d = {};
                d.type = type == "SpotLight" ? RTXPT_LIGHT_SPOT : RTXPT_LIGHT_POINT;
                d.position[0] = float(world.m[12]); d.position[1] = float(world.m[13]); d.position[2] = float(world.m[14]);
                for (int k = 0; k < 3; k++) d.direction[k] = float(-world.m[8 + k]);       // Light::GetDirection = -normalize(localToWorld.row2) (SceneTypes.cpp:67-75)
                double c[3] = { 1, 1, 1 }; readVec(src.find("color"), c); d.color[0] = float(c[0]); d.color[1] = float(c[1]); d.color[2] = float(c[2]);
                d.intensity = float(src.number("intensity", 1.0)); d.radius = float(src.number("radius", 0.0));
                d.innerAngle = float(src.number("innerAngle", 180.0)); d.outerAngle = float(src.number("outerAngle", 180.0));     // degrees (SceneTypes.h defaults)
                out->lights.push_back(d);
            }
            else if (type == "DirectionalLight") out->info.directionalLightCount++;
            else if (type == "EnvironmentLight")
            {
                std::string p = src.string("path"); for (char& ch : p) if (ch == '\\') ch = '/';
                strncpy(out->info.environmentMapPath, p.c_str(), sizeof(out->info.environmentMapPath) - 1);
                double rs[3] = { 1, 1, 1 }; readVec(src.find("radianceScale"), rs); for (int k = 0; k < 3; k++) out->info.environmentRadianceScale[k] = float(rs[k]);
                double rot[1] = { 0 }; readVec(src.find("rotation"), rot); out->info.environmentRotation = float(rot[0]);
            }
            else if (type == "PerspectiveCamera" || type == "PerspectiveCameraEx")
            {   // SceneCamera::GetViewToWorldMatrix flips z: the camera looks down the node's -Z with +Y up (SceneGraph.cpp:133-140, Sample.cpp:459-461)
                RtxptGltfCamera c = {};
                c.position[0] = float(world.m[12]); c.position[1] = float(world.m[13]); c.position[2] = float(world.m[14]);
                for (int k = 0; k < 3; k++) { c.direction[k] = float(-world.m[8 + k]); c.up[k] = float(world.m[4 + k]); }
                c.yfov = float(src.number("verticalFov", 1.0)); c.znear = float(src.number("zNear", 1.0)); c.zfar = float(src.number("zFar", 0.0)); c.aspectRatio = float(src.number("aspectRatio", 0.0));
                out->cameras.push_back(c);
            }
            else if (type == "SampleSettings")
            {
                RtxptSceneFileInfo& i = out->info; i.hasSampleSettings = 1;
                const JValue* rt = src.find("realtimeMode"); i.realtimeMode = (rt && rt->type == JValue::Bool) ? (rt->b ? 1u : 0u) : 1u;
                i.maxBounces = src.integer("maxBounces", -1); i.maxDiffuseBounces = src.integer("maxDiffuseBounces", -1);
                i.realtimeFireflyFilter = float(src.number("realtimeFireflyFilter", 0.0)); i.textureMIPBias = float(src.number("textureMIPBias", 0.0));
                strncpy(i.startingCamera, src.string("startingCamera").c_str(), sizeof(i.startingCamera) - 1);
            }
        }
    }

    void load(const std::string& path, const char* media)
    {
        const size_t slash = path.find_last_of("/\\");
        mediaDir = media && *media ? std::string(media) : (slash == std::string::npos ? std::string() : path.substr(0, slash + 1));
        if (!mediaDir.empty() && mediaDir.back() != '/' && mediaDir.back() != '\\') mediaDir += '/';
        std::string stem = path.substr(slash == std::string::npos ? 0 : slash + 1); { const size_t dot = stem.find_last_of('.'); if (dot != std::string::npos) stem.erase(dot); }     // "x.scene.json" -> "x.scene"
        materialsDir = mediaDir + "Materials/"; sceneMaterialsDir = materialsDir + stem + "/";
        const std::vector<uint8_t> file = readFile(path);
        JParser jp{ reinterpret_cast<const char*>(file.data()), reinterpret_cast<const char*>(file.data()) + file.size() };
        const JValue root = jp.parse();
        if (root.type != JValue::Object) failf("scene.json: top level is not an object");
        if (const JValue* ms = root.find("models")) for (const JValue& m : ms->arr)
        {
            std::string rel = m.type == JValue::String ? m.str : std::string(); for (char& ch : rel) if (ch == '\\') ch = '/';
            std::unique_ptr<Loader> l(new Loader()); l->out = out; l->materialsDir = materialsDir; l->sceneMaterialsDir = sceneMaterialsDir; l->mediaDir = mediaDir;
            l->open(mediaDir + rel);
            models.push_back(std::move(l));
        }
        out->info.modelCount = uint32_t(models.size());
        out->info.environmentRadianceScale[0] = out->info.environmentRadianceScale[1] = out->info.environmentRadianceScale[2] = 1.0f;
        out->info.maxBounces = out->info.maxDiffuseBounces = -1;
        if (const JValue* g = root.find("graph")) visit(*g, identity(), 0);
        finalizeScene(out, path.c_str());
    }
};

} // namespace

extern "C" RTXPT_API int rtxpt_b200_load_scene_json(const char* path, const char* mediaDir, rtxpt_host_scene** outScene)
{
    if (!path || !outScene) { g_loaderError = "null argument"; return RTXPT_ERR_INVALID_ARGUMENT; }
    *outScene = nullptr;
    std::unique_ptr<rtxpt_host_scene> scene(new rtxpt_host_scene());
    try { SceneFileLoader l; l.out = scene.get(); l.load(path, mediaDir); }
    catch (const LoadError& e) { g_loaderError = e.msg; return RTXPT_ERR_INVALID_ARGUMENT; }
    catch (const std::exception& e) { g_loaderError = e.what(); return RTXPT_ERR_INVALID_ARGUMENT; }
    *outScene = scene.release();
    return RTXPT_OK;
}
extern "C" RTXPT_API int rtxpt_b200_host_scene_info(const rtxpt_host_scene* scene, RtxptSceneFileInfo* outInfo)
{
    if (!scene || !outInfo) return RTXPT_ERR_INVALID_ARGUMENT;
    *outInfo = scene->info;
    return RTXPT_OK;
}
