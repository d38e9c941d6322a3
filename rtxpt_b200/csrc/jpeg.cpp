// Baseline / extended-sequential JPEG decoder for glTF images (image/jpeg), host side.  The reference reads them through stb_image in Donut's TextureCache
// (External/Donut/src/engine/TextureCache.cpp: stbi_load_from_memory for every non-DDS, non-EXR image); Bistro ships DDS, but glTF 2.0 allows exactly two image formats,
// PNG and JPEG, so a loader that claims glTF has to read both.  Written from ITU-T T.81 (the JPEG standard: marker syntax B.2, Huffman decoding F.2.2, IDCT A.3.3) and JFIF's
// YCbCr conversion; not from stb_image or libjpeg.
// Supported: SOF0 / SOF1, 8 bits per sample, Huffman coding, 1 (grey) or 3 (YCbCr) components, any sampling factors up to 4 (4:4:4, 4:2:2, 4:2:0, 4:1:1 ...), restart
// intervals, multiple scans are not needed for sequential files (one interleaved scan, or one scan per component).  Refused with a message: progressive (SOF2), lossless,
// arithmetic coding, 12-bit samples, CMYK / Adobe transforms.  Chroma is upsampled bilinearly at texel centres (libjpeg's "fancy upsampling" is a triangle filter too; the two
// differ by rounding), the IDCT is a separable double-precision one: decodes agree with libjpeg's to a mean of 0.04 (4:4:4) / 0.33 (subsampled chroma) levels, at most 3
// (tests/test_jpeg.py, against Pillow).
#include "../../include/rtxpt_b200.h"
#include "json_min.h"          // LoadError / failf
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace rtxpt_host {
namespace {
struct Huff { uint8_t bits[17] = {}; uint8_t vals[256] = {}; int mincode[17] = {}, maxcode[18] = {}, valptr[17] = {}; bool present = false; };
void buildHuff(Huff& h)
{   // T.81 Annex C: code lengths -> canonical codes; F.2.2.3: decoding tables
    int code = 0, k = 0;
    for (int l = 1; l <= 16; l++)
    {
        h.valptr[l] = k; h.mincode[l] = code;
        code += h.bits[l]; k += h.bits[l];
        h.maxcode[l] = h.bits[l] ? code - 1 : -1;
        code <<= 1;
    }
    h.maxcode[17] = 0x7FFFFFFF; h.present = true;
}
struct BitReader
{
    const uint8_t* p; const uint8_t* end; uint32_t acc = 0; int count = 0; bool hitMarker = false;
    void fill()
    {
        while (count <= 24)
        {
            uint32_t b = 0;
            if (!hitMarker && p < end)
            {
                b = *p;
                if (b == 0xFF)
                {
                    if (p + 1 < end && p[1] == 0x00) p += 2;            // stuffed zero byte
                    else { hitMarker = true; b = 0; }                   // a marker (RSTn / EOI): feed zeros until the caller deals with it
                }
                else p++;
            }
            acc |= b << (24 - count); count += 8;
        }
    }
    int bit() { if (count == 0) fill(); const int v = int(acc >> 31); acc <<= 1; count--; return v; }
    int bits(int n) { if (n == 0) return 0; if (count < n) fill(); const int v = int(acc >> (32 - n)); acc <<= n; count -= n; return v; }
    void reset() { acc = 0; count = 0; hitMarker = false; }
};
int decodeSymbol(BitReader& br, const Huff& h, const char* name)
{
    int code = 0;
    for (int l = 1; l <= 16; l++)
    {
        code = (code << 1) | br.bit();
        if (h.maxcode[l] >= 0 && code <= h.maxcode[l] && code >= h.mincode[l]) return h.vals[h.valptr[l] + code - h.mincode[l]];
    }
    failf("JPEG '%s': bad Huffman code", name);
}
inline int extend(int v, int t) { return (t && v < (1 << (t - 1))) ? v - (1 << t) + 1 : v; }       // F.2.2.1 EXTEND
const uint8_t kZigZag[64] = { 0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                              35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63 };
struct Component { int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0, pred = 0; int blocksW = 0, blocksH = 0; std::vector<uint8_t> plane; int planeW = 0, planeH = 0; };

void idct8x8(const int* coef, const uint16_t* q, uint8_t* out, int stride)
{   // A.3.3, separable: s(x) = 1/2 sum_u C(u) S(u) cos((2x + 1) u pi / 16)
    static double c[8][8]; static bool init = false;
    if (!init) { for (int x = 0; x < 8; x++) for (int u = 0; u < 8; u++) c[x][u] = (u == 0 ? std::sqrt(0.5) : 1.0) * std::cos((2 * x + 1) * u * 3.14159265358979323846 / 16.0) * 0.5; init = true; }
    double tmp[64];
    for (int v = 0; v < 8; v++)
        for (int x = 0; x < 8; x++)
        {
            double s = 0; for (int u = 0; u < 8; u++) s += c[x][u] * double(coef[v * 8 + u] * int(q[v * 8 + u]));
            tmp[v * 8 + x] = s;
        }
    for (int x = 0; x < 8; x++)
        for (int y = 0; y < 8; y++)
        {
            double s = 0; for (int v = 0; v < 8; v++) s += c[y][v] * tmp[v * 8 + x];
            const int r = int(std::floor(s + 128.5));
            out[y * stride + x] = uint8_t(std::min(255, std::max(0, r)));
        }
}
inline uint16_t be16(const uint8_t* p) { return uint16_t((p[0] << 8) | p[1]); }
} // namespace

struct JpegImage { uint32_t w = 0, h = 0; std::vector<uint8_t> rgba; };

JpegImage decodeJpeg(const uint8_t* d, size_t size, const char* name)
{
    if (size < 4 || d[0] != 0xFF || d[1] != 0xD8) failf("image '%s' is not a JPEG (SOI marker)", name);
    uint16_t qt[4][64] = {}; bool haveQt[4] = {}; Huff dc[4], ac[4];
    std::vector<Component> comps; int width = 0, height = 0, restartInterval = 0; bool haveFrame = false, decoded = false; int adobeTransform = -1;
    size_t pos = 2;
    while (pos + 4 <= size && !decoded)
    {
        if (d[pos] != 0xFF) failf("JPEG '%s': marker expected at byte %zu", name, pos);
        while (pos < size && d[pos] == 0xFF) pos++;                    // fill bytes
        if (pos >= size) break;
        const uint8_t m = d[pos++];
        if (m == 0xD9) break;                                          // EOI
        if (m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;           // standalone markers
        if (pos + 2 > size) failf("JPEG '%s': truncated", name);
        const size_t len = be16(d + pos); if (len < 2 || pos + len > size) failf("JPEG '%s': truncated segment", name);
        const uint8_t* s = d + pos + 2; const size_t n = len - 2;
        switch (m)
        {
        case 0xDB:                                                     // DQT
            for (size_t i = 0; i < n;)
            {
                const int pq = s[i] >> 4, tq = s[i] & 15; i++;
                if (tq > 3 || pq > 1) failf("JPEG '%s': bad quantisation table", name);
                if (i + (pq ? 128 : 64) > n) failf("JPEG '%s': truncated quantisation table", name);
                for (int k = 0; k < 64; k++) { qt[tq][kZigZag[k]] = pq ? be16(s + i + 2 * k) : s[i + k]; }
                i += pq ? 128 : 64; haveQt[tq] = true;
            }
            break;
        case 0xC4:                                                     // DHT
            for (size_t i = 0; i < n;)
            {
                const int tc = s[i] >> 4, th = s[i] & 15; i++;
                if (tc > 1 || th > 3 || i + 16 > n) failf("JPEG '%s': bad Huffman table", name);
                Huff& h = tc ? ac[th] : dc[th]; int total = 0;
                h.bits[0] = 0; for (int l = 1; l <= 16; l++) { h.bits[l] = s[i + l - 1]; total += h.bits[l]; }
                i += 16; if (total > 256 || i + size_t(total) > n) failf("JPEG '%s': bad Huffman table", name);
                memcpy(h.vals, s + i, size_t(total)); i += size_t(total); buildHuff(h);
            }
            break;
        case 0xC0: case 0xC1:                                          // SOF0 / SOF1
        {
            if (haveFrame) failf("JPEG '%s': more than one frame", name);
            if (n < 6) failf("JPEG '%s': truncated frame header", name);
            if (s[0] != 8) failf("JPEG '%s': %u-bit samples are not supported (8-bit are)", name, unsigned(s[0]));
            height = be16(s + 1); width = be16(s + 3); const int nc = s[5];
            if (!width || !height || width > 32768 || height > 32768 || int64_t(width) * height > (int64_t(1) << 28)) failf("JPEG '%s': bad dimensions (at most 32768 per side and 2^28 pixels)", name);
            if (nc != 1 && nc != 3) failf("JPEG '%s': %d colour components are not supported (grey and YCbCr are)", name, nc);
            if (n < size_t(6 + 3 * nc)) failf("JPEG '%s': truncated frame header", name);
            comps.resize(size_t(nc));
            for (int i = 0; i < nc; i++)
            {
                Component& c = comps[size_t(i)]; c.id = s[6 + 3 * i]; c.h = s[7 + 3 * i] >> 4; c.v = s[7 + 3 * i] & 15; c.tq = s[8 + 3 * i];
                if (c.h < 1 || c.h > 4 || c.v < 1 || c.v > 4 || c.tq > 3) failf("JPEG '%s': bad sampling factors", name);
            }
            haveFrame = true;
            break;
        }
        case 0xC2: failf("JPEG '%s': progressive JPEG is not supported (baseline / sequential are); re-save the image as baseline or PNG", name);
        case 0xC3: case 0xC5: case 0xC6: case 0xC7: case 0xC9: case 0xCA: case 0xCB: case 0xCD: case 0xCE: case 0xCF:
            failf("JPEG '%s': coding process 0x%02X (lossless / hierarchical / arithmetic) is not supported", name, unsigned(m));
        case 0xDD: if (n < 2) failf("JPEG '%s': bad restart interval", name); restartInterval = be16(s); break;
        case 0xEE: if (n >= 12 && !memcmp(s, "Adobe", 5)) adobeTransform = s[11]; break;
        case 0xDA:                                                     // SOS: the entropy-coded data follows
        {
            if (!haveFrame) failf("JPEG '%s': scan before frame header", name);
            if (n < 1) failf("JPEG '%s': truncated scan header", name);
            const int ns = s[0]; if (ns < 1 || ns > int(comps.size()) || n < size_t(1 + 2 * ns + 3)) failf("JPEG '%s': bad scan header", name);
            std::vector<Component*> scan;
            for (int i = 0; i < ns; i++)
            {
                Component* c = nullptr; for (Component& k : comps) if (k.id == s[1 + 2 * i]) c = &k;
                if (!c) failf("JPEG '%s': scan names an unknown component", name);
                c->td = s[2 + 2 * i] >> 4; c->ta = s[2 + 2 * i] & 15; if (c->td > 3 || c->ta > 3 || !dc[c->td].present || !ac[c->ta].present || !haveQt[c->tq]) failf("JPEG '%s': scan uses a table that was not defined", name);
                c->pred = 0; scan.push_back(c);
            }
            int hmax = 1, vmax = 1; for (const Component& c : comps) { hmax = std::max(hmax, c.h); vmax = std::max(vmax, c.v); }
            const int mcusX = (width + 8 * hmax - 1) / (8 * hmax), mcusY = (height + 8 * vmax - 1) / (8 * vmax);
            for (Component& c : comps) if (c.plane.empty()) { c.blocksW = mcusX * c.h; c.blocksH = mcusY * c.v; c.planeW = c.blocksW * 8; c.planeH = c.blocksH * 8; c.plane.assign(size_t(c.planeW) * c.planeH, 128); }
            BitReader br{ d + pos + len, d + size };
            const bool interleaved = ns > 1;
            // a non-interleaved scan covers ceil(component size / 8) blocks, not the padded MCU grid (A.2.3)
            const int unitsX = interleaved ? mcusX : (((width * scan[0]->h + hmax - 1) / hmax) + 7) / 8, unitsY = interleaved ? mcusY : (((height * scan[0]->v + vmax - 1) / vmax) + 7) / 8;
            int untilRestart = restartInterval, rstExpected = 0;
            for (int my = 0; my < unitsY; my++) for (int mx = 0; mx < unitsX; mx++)
            {
                if (restartInterval && untilRestart == 0)
                {   // B.2.1: RSTm between intervals; byte-align, check, reset the predictors
                    br.reset();
                    const uint8_t* q = br.p; while (q + 1 < br.end && !(q[0] == 0xFF && q[1] >= 0xD0 && q[1] <= 0xD7)) q++;
                    if (q + 1 >= br.end || q[1] != 0xD0 + rstExpected) failf("JPEG '%s': restart marker missing", name);
                    br.p = q + 2; rstExpected = (rstExpected + 1) & 7; untilRestart = restartInterval;
                    for (Component* c : scan) c->pred = 0;
                }
                for (Component* c : scan)
                {
                    const int bw = interleaved ? c->h : 1, bh = interleaved ? c->v : 1;
                    for (int by = 0; by < bh; by++) for (int bx = 0; bx < bw; bx++)
                    {
                        int coef[64] = {};
                        const int t = decodeSymbol(br, dc[c->td], name); if (t > 11) failf("JPEG '%s': bad DC category", name);
                        c->pred += extend(br.bits(t), t); coef[0] = c->pred;
                        for (int k = 1; k < 64;)
                        {
                            const int rs = decodeSymbol(br, ac[c->ta], name), r = rs >> 4, sz = rs & 15;
                            if (sz == 0) { if (r == 15) { k += 16; continue; } break; }          // ZRL / EOB
                            k += r; if (k > 63) failf("JPEG '%s': coefficient index out of range", name);
                            coef[kZigZag[k]] = extend(br.bits(sz), sz); k++;
                        }
                        const int blockX = interleaved ? mx * c->h + bx : mx, blockY = interleaved ? my * c->v + by : my;
                        if (blockX < c->blocksW && blockY < c->blocksH) idct8x8(coef, qt[c->tq], c->plane.data() + size_t(blockY) * 8 * c->planeW + size_t(blockX) * 8, c->planeW);
                    }
                }
                if (restartInterval) untilRestart--;
            }
            // next marker: skip what is left of the entropy-coded segment
            const uint8_t* q = br.p; while (q + 1 < br.end && !(q[0] == 0xFF && q[1] != 0x00 && !(q[1] >= 0xD0 && q[1] <= 0xD7))) q++;
            pos = size_t(q - d);
            bool all = true; for (const Component& c : comps) all = all && !c.plane.empty();
            static_cast<void>(all);
            // sequential files carry every component exactly once: after the last scan the image is complete; keep reading markers until EOI
            continue;
        }
        default: break;                                                // APPn, COM, ...: skipped
        }
        pos += len;
    }
    if (!haveFrame) failf("JPEG '%s': no frame header", name);
    for (const Component& c : comps) if (c.plane.empty()) failf("JPEG '%s': component %d has no scan", name, c.id);
    if (comps.size() == 3 && adobeTransform == 0) failf("JPEG '%s': Adobe RGB / CMYK transforms are not supported", name);
    JpegImage img; img.w = uint32_t(width); img.h = uint32_t(height); img.rgba.resize(size_t(width) * height * 4);
    int hmax = 1, vmax = 1; for (const Component& c : comps) { hmax = std::max(hmax, c.h); vmax = std::max(vmax, c.v); }
    auto sample = [&](const Component& c, int x, int y) -> float
    {   // full-resolution sample of a (possibly subsampled) component: bilinear between the component's texel centres
        if (c.h == hmax && c.v == vmax) return float(c.plane[size_t(y) * c.planeW + x]);
        const int cw = (width * c.h + hmax - 1) / hmax, ch = (height * c.v + vmax - 1) / vmax;
        const float fx = (x + 0.5f) * float(c.h) / float(hmax) - 0.5f, fy = (y + 0.5f) * float(c.v) / float(vmax) - 0.5f;
        const int x0 = int(std::floor(fx)), y0 = int(std::floor(fy)); const float tx = fx - float(x0), ty = fy - float(y0);
        auto at = [&](int xx, int yy) { xx = std::min(std::max(xx, 0), cw - 1); yy = std::min(std::max(yy, 0), ch - 1); return float(c.plane[size_t(yy) * c.planeW + xx]); };
        return (at(x0, y0) * (1 - tx) + at(x0 + 1, y0) * tx) * (1 - ty) + (at(x0, y0 + 1) * (1 - tx) + at(x0 + 1, y0 + 1) * tx) * ty;
    };
    auto clamp8 = [](float v) { return uint8_t(std::min(255.0f, std::max(0.0f, std::floor(v + 0.5f)))); };
    for (int y = 0; y < height; y++) for (int x = 0; x < width; x++)
    {
        uint8_t* o = img.rgba.data() + (size_t(y) * width + x) * 4; o[3] = 255;
        const float Y = sample(comps[0], x, y);
        if (comps.size() == 1) { o[0] = o[1] = o[2] = clamp8(Y); continue; }
        const float cb = sample(comps[1], x, y) - 128.0f, cr = sample(comps[2], x, y) - 128.0f;          // JFIF: full-range BT.601
        o[0] = clamp8(Y + 1.402f * cr); o[1] = clamp8(Y - 0.344136f * cb - 0.714136f * cr); o[2] = clamp8(Y + 1.772f * cb);
    }
    return img;
}
} // namespace rtxpt_host

static thread_local std::string g_jpegError;
// Inspection hook: decodes a JPEG held in memory into RGBA8 (tests/test_jpeg.py compares it with Pillow's decode); call with outRGBA == NULL for the size
extern "C" RTXPT_API int rtxpt_b200_debug_decode_jpeg(const void* fileBytes, uint64_t fileSize, uint32_t* outWidth, uint32_t* outHeight, uint8_t* outRGBA, uint64_t outCapacity)
{
    if (!fileBytes || !outWidth || !outHeight) return RTXPT_ERR_INVALID_ARGUMENT;
    try
    {
        const rtxpt_host::JpegImage img = rtxpt_host::decodeJpeg(static_cast<const uint8_t*>(fileBytes), size_t(fileSize), "<memory>");
        *outWidth = img.w; *outHeight = img.h;
        if (outRGBA) { if (outCapacity < img.rgba.size()) { g_jpegError = "output buffer too small"; return RTXPT_ERR_INVALID_ARGUMENT; } memcpy(outRGBA, img.rgba.data(), img.rgba.size()); }
    }
    catch (const rtxpt_host::LoadError& e) { g_jpegError = e.msg; return RTXPT_ERR_INVALID_ARGUMENT; }
    catch (const std::exception& e) { g_jpegError = e.what(); return RTXPT_ERR_INVALID_ARGUMENT; }
    return RTXPT_OK;
}
extern "C" RTXPT_API const char* rtxpt_b200_debug_decode_jpeg_error(void) { return g_jpegError.c_str(); }
