// HDR image files other than DDS, host side: OpenEXR (scan-line files) and Radiance .hdr (RGBE).
//   - the reference enumerates `.exr`, `.hdr` and `.dds` files as environment-map sources (Rtxpt/Sample.cpp:110-118) and reads them through Donut's TextureCache
//     (External/Donut/src/engine/TextureCache.cpp:200-236: tinyexr's LoadEXRFromMemory; stb_image's HDR reader for .hdr) before EnvMapBaker turns them into the cube;
//   - BASELINE.md §3 keeps an EXR import so that an `AccumulatedRadiance` dump of a real RTXPT run, made off this box, can be compared with our accumulation
//     (scripts/compare_hdr_images.py).
// Both readers are written from the file formats' published layouts (OpenEXR "File Layout" document; Ward's RGBE description), not from those libraries.
// EXR support: single-part scan-line files, compression NONE / RLE / ZIPS / ZIP (what tinyexr's SaveEXR and OpenEXR's defaults produce), channel types HALF / FLOAT / UINT,
// channels named R G B A (or a single luminance channel Y); PIZ, PXR24, B44 and DWA files, tiled, deep and multi-part files are refused with a message that says so.
#include "../../include/rtxpt_b200.h"
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <zlib.h>

namespace rtxpt_host {
struct HdrFileImage { uint32_t width = 0, height = 0; std::vector<float> rgba; };
namespace {
struct ImgError { std::string msg; };
[[noreturn]] void fail(const char* fmt, ...)
{
    char buf[512]; va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap); throw ImgError{ buf };
}
struct Reader
{
    const uint8_t* p; size_t n, at = 0;
    void need(size_t k) const { if (k > n - at) fail("EXR: truncated file"); }
    uint8_t u8() { need(1); return p[at++]; }
    uint32_t u32() { need(4); uint32_t v; memcpy(&v, p + at, 4); at += 4; return v; }
    int32_t i32() { return int32_t(u32()); }
    uint64_t u64() { need(8); uint64_t v; memcpy(&v, p + at, 8); at += 8; return v; }
    std::string str(size_t maxLen = 255)
    {
        std::string s;
        for (;;) { const uint8_t c = u8(); if (!c) break; if (s.size() >= maxLen) fail("EXR: attribute name too long"); s.push_back(char(c)); }
        return s;
    }
};
inline float halfToFloat(uint16_t h)
{
    const uint32_t s = uint32_t(h >> 15) << 31, e = (h >> 10) & 31u, m = h & 1023u; uint32_t bits;
    if (e == 0) { if (!m) bits = s; else { int k = 0; uint32_t mm = m; while (!(mm & 1024u)) { mm <<= 1; k++; } bits = s | uint32_t(127 - 15 - k + 1) << 23 | (mm & 1023u) << 13; } }
    else if (e == 31) bits = s | 0x7F800000u | m << 13;
    else bits = s | (e + 112u) << 23 | m << 13;
    float f; memcpy(&f, &bits, 4); return f;
}
struct Channel { std::string name; int type; int xs, ys; };

// undoes OpenEXR's byte predictor and the even / odd byte split that precede RLE and ZIP compression
void unpredictAndInterleave(std::vector<uint8_t>& t, std::vector<uint8_t>& out)
{
    const size_t n = t.size();
    for (size_t i = 1; i < n; i++) t[i] = uint8_t(int(t[i - 1]) + int(t[i]) - 128);
    out.resize(n); const size_t half = (n + 1) / 2;
    for (size_t i = 0; i < n; i++) out[i] = (i & 1) ? t[half + i / 2] : t[i / 2];
}
void inflateBlock(const uint8_t* src, size_t srcBytes, std::vector<uint8_t>& dst, size_t expected)
{
    dst.resize(expected); uLongf got = uLongf(expected);
    if (uncompress(dst.data(), &got, src, uLong(srcBytes)) != Z_OK || got != expected) fail("EXR: a ZIP block does not inflate to its scan lines");
}
void unRle(const uint8_t* src, size_t srcBytes, std::vector<uint8_t>& dst, size_t expected)
{
    dst.clear(); dst.reserve(expected); size_t i = 0;
    while (i < srcBytes)
    {
        const int c = int8_t(src[i++]);
        if (c < 0) { const size_t k = size_t(-c); if (i + k > srcBytes || dst.size() + k > expected) fail("EXR: bad RLE block"); dst.insert(dst.end(), src + i, src + i + k); i += k; }
        else { const size_t k = size_t(c) + 1; if (i >= srcBytes || dst.size() + k > expected) fail("EXR: bad RLE block"); dst.insert(dst.end(), k, src[i++]); }
    }
    if (dst.size() != expected) fail("EXR: an RLE block does not expand to its scan lines");
}
} // namespace

HdrFileImage decodeExr(const uint8_t* data, size_t size)
{
    Reader r{ data, size };
    if (size < 8 || r.u32() != 20000630u) fail("not an OpenEXR file (magic number)");
    const uint32_t version = r.u32();
    if ((version & 0xFF) != 2) fail("EXR: file format version %u is not supported", version & 0xFF);
    if (version & 0x200) fail("EXR: tiled files are not supported (scan-line files only)");
    if (version & 0x800) fail("EXR: deep-data files are not supported");
    if (version & 0x1000) fail("EXR: multi-part files are not supported");
    const size_t maxName = (version & 0x400) ? 255 : 31;
    std::vector<Channel> channels; int compression = -1, lineOrder = 0; int32_t dw[4] = { 0, 0, -1, -1 }; bool haveDw = false;
    for (;;)
    {
        const std::string name = r.str(maxName); if (name.empty()) break;
        const std::string type = r.str(maxName); const uint32_t bytes = r.u32(); r.need(bytes);
        Reader a{ data + r.at, bytes }; r.at += bytes;
        if (name == "channels")
        {
            if (type != "chlist") fail("EXR: 'channels' is not a chlist");
            for (;;)
            {
                Channel c; c.name = a.str(maxName); if (c.name.empty()) break;
                c.type = a.i32(); a.u8(); a.u8(); a.u8(); a.u8(); c.xs = a.i32(); c.ys = a.i32();
                if (c.type < 0 || c.type > 2) fail("EXR: channel '%s' has an unknown pixel type", c.name.c_str());
                if (c.xs != 1 || c.ys != 1) fail("EXR: sub-sampled channels are not supported");
                channels.push_back(c); if (channels.size() > 64) fail("EXR: too many channels");
            }
        }
        else if (name == "compression") { if (bytes != 1) fail("EXR: bad compression attribute"); compression = a.u8(); }
        else if (name == "dataWindow") { if (bytes != 16) fail("EXR: bad dataWindow"); for (int& v : dw) v = a.i32(); haveDw = true; }
        else if (name == "lineOrder") { if (bytes != 1) fail("EXR: bad lineOrder"); lineOrder = a.u8(); }
    }
    if (channels.empty() || compression < 0 || !haveDw) fail("EXR: header lacks channels, compression or dataWindow");
    static const char* const kNames[] = { "none", "RLE", "ZIPS", "ZIP", "PIZ", "PXR24", "B44", "B44A", "DWAA", "DWAB" };
    if (compression > 3) fail("EXR: %s compression is not supported (NONE, RLE, ZIPS, ZIP are); re-save the file with ZIP", compression < 10 ? kNames[compression] : "this");
    const int64_t w64 = int64_t(dw[2]) - dw[0] + 1, h64 = int64_t(dw[3]) - dw[1] + 1;
    if (w64 <= 0 || h64 <= 0 || w64 > 32768 || h64 > 32768 || w64 * h64 > (int64_t(1) << 28)) fail("EXR: bad dataWindow size (at most 32768 per side and 2^28 pixels)");
    const uint32_t W = uint32_t(w64), H = uint32_t(h64);
    // channel -> RGBA slot; a lone Y channel fills R, G and B.  Layer prefixes ("diffuse.R") are not mapped: the first-part colour channels are what RTXPT's dumps hold.
    std::vector<int> slot(channels.size(), -1); size_t pixelBytes = 0; bool haveY = false, haveRGB = false;
    for (size_t i = 0; i < channels.size(); i++)
    {
        const std::string& nm = channels[i].name;
        slot[i] = nm == "R" ? 0 : nm == "G" ? 1 : nm == "B" ? 2 : nm == "A" ? 3 : nm == "Y" ? 4 : -1;
        haveY |= slot[i] == 4; haveRGB |= slot[i] >= 0 && slot[i] <= 2;
        pixelBytes += channels[i].type == 1 ? 2 : 4;
    }
    if (!haveY && !haveRGB) fail("EXR: no R, G, B or Y channel");
    const uint32_t linesPerBlock = compression == 3 ? 16u : 1u, blocks = (H + linesPerBlock - 1) / linesPerBlock;
    r.need(size_t(blocks) * 8);
    std::vector<uint64_t> offsets(blocks); for (auto& o : offsets) o = r.u64();
    HdrFileImage img; img.width = W; img.height = H; img.rgba.assign(size_t(W) * H * 4, 0.0f);
    for (size_t i = 0; i < size_t(W) * H; i++) img.rgba[4 * i + 3] = 1.0f;
    std::vector<uint8_t> tmp, raw;
    (void)lineOrder;        // every chunk carries its own y: the order of the chunks in the file does not matter to this reader
    for (uint32_t b = 0; b < blocks; b++)
    {
        if (offsets[b] > size || size - offsets[b] < 8) fail("EXR: chunk offset outside the file");
        Reader c{ data, size, size_t(offsets[b]) };
        const int64_t y0 = int64_t(c.i32()) - dw[1]; const uint32_t bytes = c.u32(); c.need(bytes);
        if (y0 < 0 || y0 >= int64_t(H) || (y0 % linesPerBlock) != 0) fail("EXR: chunk with a scan line outside the data window");
        const uint32_t lines = std::min<uint32_t>(linesPerBlock, H - uint32_t(y0)); const size_t expected = size_t(lines) * W * pixelBytes;
        const uint8_t* src = data + c.at; const uint8_t* px;
        if (compression == 0 || bytes == expected) { if (bytes != expected) fail("EXR: uncompressed chunk of the wrong size"); px = src; }      // a block that did not shrink is stored as is
        else
        {
            if (compression == 1) unRle(src, bytes, tmp, expected); else inflateBlock(src, bytes, tmp, expected);
            unpredictAndInterleave(tmp, raw); px = raw.data();
        }
        for (uint32_t l = 0; l < lines; l++)
        {
            float* row = img.rgba.data() + (size_t(y0) + l) * W * 4;
            for (size_t ch = 0; ch < channels.size(); ch++)
            {
                const int type = channels[ch].type, s = slot[ch]; const size_t bpp = type == 1 ? 2 : 4;
                if (s >= 0)
                    for (uint32_t x = 0; x < W; x++)
                    {
                        float v;
                        if (type == 1) { uint16_t h; memcpy(&h, px + x * 2, 2); v = halfToFloat(h); }
                        else if (type == 2) memcpy(&v, px + x * 4, 4);
                        else { uint32_t u; memcpy(&u, px + x * 4, 4); v = float(u); }
                        if (s == 4) { if (!haveRGB) row[4 * x] = row[4 * x + 1] = row[4 * x + 2] = v; } else row[4 * x + s] = v;
                    }
                px += bpp * W;
            }
        }
    }
    return img;
}

// Radiance RGBE (.hdr): text header up to an empty line, the resolution line, then scan lines either flat (4 bytes per pixel) or run-length coded per component
HdrFileImage decodeRadianceHdr(const uint8_t* data, size_t size)
{
    size_t at = 0;
    auto line = [&]() { std::string s; while (at < size && data[at] != '\n') { if (s.size() > 4096) fail("HDR: header line too long"); s.push_back(char(data[at++])); } if (at >= size) fail("HDR: truncated header"); at++; return s; };
    if (size < 7 || (memcmp(data, "#?RADIANCE", std::min<size_t>(size, 10)) != 0 && memcmp(data, "#?RGBE", 6) != 0)) fail("neither an OpenEXR, a Radiance HDR nor a DDS file (signature)");
    line();
    bool rgbe = false;
    for (;;) { const std::string s = line(); if (s.empty()) break; if (s.rfind("FORMAT=", 0) == 0) { if (s != "FORMAT=32-bit_rle_rgbe") fail("HDR: format '%s' is not supported (32-bit_rle_rgbe is)", s.c_str()); rgbe = true; } }
    if (!rgbe) fail("HDR: no FORMAT line");
    const std::string res = line(); int H = 0, W = 0;
    if (sscanf(res.c_str(), "-Y %d +X %d", &H, &W) != 2) fail("HDR: resolution line '%s' is not supported (-Y h +X w is)", res.c_str());
    if (W <= 0 || H <= 0 || W > 32768 || H > 32768 || int64_t(W) * H > (int64_t(1) << 28)) fail("HDR: bad dimensions (at most 32768 per side and 2^28 pixels)");
    HdrFileImage img; img.width = uint32_t(W); img.height = uint32_t(H); img.rgba.resize(size_t(W) * H * 4);
    std::vector<uint8_t> scan(size_t(W) * 4);
    auto put = [&](float* o, const uint8_t* p)
    {   // mantissa * 2^(exponent - 128 - 8), no half-step bias: what stb_image's reader (the reference's, through Donut) produces
        if (p[3]) { const float f = ldexpf(1.0f, int(p[3]) - 136); o[0] = p[0] * f; o[1] = p[1] * f; o[2] = p[2] * f; } else o[0] = o[1] = o[2] = 0.0f;
        o[3] = 1.0f;
    };
    for (int y = 0; y < H; y++)
    {
        float* row = img.rgba.data() + size_t(y) * W * 4;
        const bool rle = W >= 8 && W < 32768 && size - at >= 4 && data[at] == 2 && data[at + 1] == 2 && !(data[at + 2] & 0x80);
        if (!rle)
        {
            if (size - at < size_t(W) * 4) fail("HDR: truncated pixel data");
            for (int x = 0; x < W; x++) put(row + 4 * x, data + at + 4 * size_t(x));
            at += size_t(W) * 4; continue;
        }
        if (((int(data[at + 2]) << 8) | data[at + 3]) != W) fail("HDR: scan line of the wrong width");
        at += 4;
        for (int comp = 0; comp < 4; comp++)
            for (int x = 0; x < W;)
            {
                if (at >= size) fail("HDR: truncated pixel data");
                int count = data[at++];
                if (count > 128) { count -= 128; if (count > W - x || at >= size) fail("HDR: corrupt run"); const uint8_t v = data[at++]; for (int k = 0; k < count; k++) scan[size_t(x++) * 4 + comp] = v; }
                else { if (count == 0 || count > W - x || size - at < size_t(count)) fail("HDR: corrupt run"); for (int k = 0; k < count; k++) scan[size_t(x++) * 4 + comp] = data[at++]; }
            }
        for (int x = 0; x < W; x++) put(row + 4 * x, scan.data() + 4 * size_t(x));
    }
    return img;
}
} // namespace rtxpt_host

static thread_local std::string g_hdrImageError;
// OpenEXR / Radiance .hdr / HDR DDS by content: sizes with outRGBA32F == NULL, then RGBA32F rows top to bottom (faces back to back for a DDS cube)
extern "C" RTXPT_API int rtxpt_b200_load_hdr_image(const void* fileBytes, uint64_t fileSize, uint32_t* outWidth, uint32_t* outHeight, uint32_t* outFaces, float* outRGBA32F, uint64_t outCapacityFloats)
{
    if (!fileBytes || !outWidth || !outHeight || !outFaces) return RTXPT_ERR_INVALID_ARGUMENT;
    const uint8_t* d = static_cast<const uint8_t*>(fileBytes);
    if (fileSize >= 4 && !memcmp(d, "DDS ", 4))
    {
        const int rc = rtxpt_b200_load_dds_hdr(fileBytes, fileSize, outWidth, outHeight, outFaces, nullptr, outRGBA32F, outCapacityFloats);
        if (rc != RTXPT_OK) g_hdrImageError = rtxpt_b200_debug_decode_dds_error();
        return rc;
    }
    try
    {
        const bool exr = fileSize >= 4 && d[0] == 0x76 && d[1] == 0x2F && d[2] == 0x31 && d[3] == 0x01;
        const rtxpt_host::HdrFileImage img = exr ? rtxpt_host::decodeExr(d, size_t(fileSize)) : rtxpt_host::decodeRadianceHdr(d, size_t(fileSize));
        *outWidth = img.width; *outHeight = img.height; *outFaces = 1;
        if (outRGBA32F) { if (outCapacityFloats < img.rgba.size()) { g_hdrImageError = "output buffer too small"; return RTXPT_ERR_INVALID_ARGUMENT; } memcpy(outRGBA32F, img.rgba.data(), img.rgba.size() * sizeof(float)); }
    }
    catch (const rtxpt_host::ImgError& e) { g_hdrImageError = e.msg; return RTXPT_ERR_INVALID_ARGUMENT; }
    catch (const std::exception& e) { g_hdrImageError = e.what(); return RTXPT_ERR_INVALID_ARGUMENT; }
    return RTXPT_OK;
}
extern "C" RTXPT_API const char* rtxpt_b200_load_hdr_image_error(void) { return g_hdrImageError.c_str(); }
