// dds.cpp — DDS container + block-compressed texture decoding on the host (BC1/2/3/4/5/7, RGBA8/BGRA8), so that the DDS files the reference's assets and
// `.material.json` files point at (Donut TextureCache / DDSFile.cpp; MSFT_texture_dds in its glTF files) can be handed to the kernels as the RGBA8
// mip chains RtxptTextureDesc carries.  The reference samples these formats through the texture units; decoding to RGBA8 is exact for BC1-5/7
// (their decoders are integer-exact by specification) except for BC1/2/3's colour interpolation, where GPUs may differ from the reference decoder by
// one LSB (documented D3D tolerance).  HDR files (BC6H UF16 / SF16, RGBA16F, RGBA32F; the reference's environment cubes) are decoded to RGBA32F by decodeDdsHdr.
// Checked block by block against an independent decoder (Pillow) in tests/test_dds.py.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/rtxpt_b200.h"
#include "json_min.h"          // LoadError / failf

namespace rtxpt_host {

#include "bc7_tables.inc"

struct DdsImage { uint32_t width = 0, height = 0; bool srgb = false; std::vector<std::vector<uint8_t>> mips; };   // RGBA8, tightly packed

namespace {

inline void color565(uint16_t c, uint8_t out[3])
{
    const uint32_t r = (c >> 11) & 31, g = (c >> 5) & 63, b = c & 31;
    out[0] = uint8_t((r << 3) | (r >> 2)); out[1] = uint8_t((g << 2) | (g >> 4)); out[2] = uint8_t((b << 3) | (b >> 2));
}
// BC1 colour block (also the colour half of BC2/BC3, which never use the 3-colour + transparent mode).  Thirds are truncated - the rounding the
// widely used software decoders (and the test's independent decoder) apply; hardware may differ by one LSB, which D3D allows
void decodeBc1Color(const uint8_t* b, bool allowPunchThrough, uint8_t out[16][4])
{
    const uint16_t c0 = uint16_t(b[0] | (b[1] << 8)), c1 = uint16_t(b[2] | (b[3] << 8));
    uint8_t pal[4][4]; color565(c0, pal[0]); color565(c1, pal[1]); pal[0][3] = pal[1][3] = pal[2][3] = pal[3][3] = 255;
    if (c0 > c1 || !allowPunchThrough)
        for (int k = 0; k < 3; k++) { pal[2][k] = uint8_t((2 * pal[0][k] + pal[1][k]) / 3); pal[3][k] = uint8_t((pal[0][k] + 2 * pal[1][k]) / 3); }
    else
    {
        for (int k = 0; k < 3; k++) { pal[2][k] = uint8_t((pal[0][k] + pal[1][k]) / 2); pal[3][k] = 0; }
        pal[3][3] = 0;
    }
    const uint32_t idx = uint32_t(b[4]) | (uint32_t(b[5]) << 8) | (uint32_t(b[6]) << 16) | (uint32_t(b[7]) << 24);
    for (int i = 0; i < 16; i++) memcpy(out[i], pal[(idx >> (2 * i)) & 3], 4);
}
// BC4-style single-channel block: 2 endpoints + 16 x 3-bit indices
void decodeBc4Channel(const uint8_t* b, uint8_t out[16])
{
    const uint32_t a0 = b[0], a1 = b[1]; uint8_t pal[8]; pal[0] = uint8_t(a0); pal[1] = uint8_t(a1);
    if (a0 > a1) for (uint32_t k = 1; k < 7; k++) pal[k + 1] = uint8_t(((7 - k) * a0 + k * a1) / 7);
    else { for (uint32_t k = 1; k < 5; k++) pal[k + 1] = uint8_t(((5 - k) * a0 + k * a1) / 5); pal[6] = 0; pal[7] = 255; }
    uint64_t bits = 0; for (int k = 0; k < 6; k++) bits |= uint64_t(b[2 + k]) << (8 * k);
    for (int i = 0; i < 16; i++) out[i] = pal[(bits >> (3 * i)) & 7];
}

struct BitReader
{
    const uint8_t* p; uint32_t pos = 0;
    uint32_t get(uint32_t n) { uint32_t v = 0; for (uint32_t i = 0; i < n; i++, pos++) v |= uint32_t((p[pos >> 3] >> (pos & 7)) & 1u) << i; return v; }
};
const uint8_t kW2[4] = { 0, 21, 43, 64 }, kW3[8] = { 0, 9, 18, 27, 37, 46, 55, 64 }, kW4[16] = { 0, 4, 9, 13, 17, 21, 26, 30, 34, 38, 43, 47, 51, 55, 60, 64 };
inline uint8_t lerp7(uint32_t a, uint32_t b, uint32_t w) { return uint8_t(((64 - w) * a + w * b + 32) >> 6); }

void decodeBc7(const uint8_t* blk, uint8_t out[16][4])
{
    uint32_t mode = 0; while (mode < 8 && !((blk[0] >> mode) & 1)) mode++;
    if (mode >= 8) { memset(out, 0, 64); return; }                          // reserved: transparent black
    static const uint8_t subsets[8] = { 3, 2, 3, 2, 1, 1, 1, 2 }, partBits[8] = { 4, 6, 6, 6, 0, 0, 0, 6 }, rotBits[8] = { 0, 0, 0, 0, 2, 2, 0, 0 }, selBits[8] = { 0, 0, 0, 0, 1, 0, 0, 0 };
    static const uint8_t colorBits[8] = { 4, 6, 5, 7, 5, 7, 7, 5 }, alphaBits[8] = { 0, 0, 0, 0, 6, 8, 7, 5 }, pbitMode[8] = { 1, 2, 0, 1, 0, 0, 1, 1 };     // 1: per endpoint, 2: per subset
    static const uint8_t idxBits[8] = { 3, 3, 2, 2, 2, 2, 4, 2 }, idx2Bits[8] = { 0, 0, 0, 0, 3, 2, 0, 0 };
    BitReader br{ blk }; br.get(mode + 1);
    const uint32_t ns = subsets[mode], partition = br.get(partBits[mode]), rotation = br.get(rotBits[mode]), idxSel = br.get(selBits[mode]);
    uint32_t ep[6][4];
    for (int c = 0; c < 3; c++) for (uint32_t e = 0; e < 2 * ns; e++) ep[e][c] = br.get(colorBits[mode]);
    for (uint32_t e = 0; e < 2 * ns; e++) ep[e][3] = alphaBits[mode] ? br.get(alphaBits[mode]) : 255u;
    uint32_t cb = colorBits[mode], ab = alphaBits[mode];
    if (pbitMode[mode])
    {
        uint32_t pb[6];
        if (pbitMode[mode] == 1) for (uint32_t e = 0; e < 2 * ns; e++) pb[e] = br.get(1);
        else for (uint32_t s = 0; s < ns; s++) { pb[2 * s] = br.get(1); pb[2 * s + 1] = pb[2 * s]; }
        for (uint32_t e = 0; e < 2 * ns; e++) { for (int c = 0; c < 3; c++) ep[e][c] = (ep[e][c] << 1) | pb[e]; if (ab) ep[e][3] = (ep[e][3] << 1) | pb[e]; }
        cb++; if (ab) ab++;
    }
    for (uint32_t e = 0; e < 2 * ns; e++)
    {
        for (int c = 0; c < 3; c++) { const uint32_t v = ep[e][c] << (8 - cb); ep[e][c] = v | (v >> cb); }
        if (ab) { const uint32_t v = ep[e][3] << (8 - ab); ep[e][3] = v | (v >> ab); }
    }
    const uint8_t* part = (ns == 2) ? kBc7Partition2[partition] : (ns == 3 ? kBc7Partition3[partition] : nullptr);
    uint32_t anchors[3] = { 0, 0, 0 };
    if (ns == 2) anchors[1] = kBc7Anchor2[partition];
    if (ns == 3) { anchors[1] = kBc7Anchor3a[partition]; anchors[2] = kBc7Anchor3b[partition]; }
    uint32_t i1[16], i2[16];
    const uint32_t b1 = idxBits[mode], b2 = idx2Bits[mode];
    for (uint32_t i = 0; i < 16; i++) { const uint32_t s = part ? part[i] : 0; i1[i] = br.get((i == anchors[s]) ? b1 - 1 : b1); }
    for (uint32_t i = 0; i < 16; i++) i2[i] = b2 ? br.get(i == 0 ? b2 - 1 : b2) : 0;
    auto weight = [](uint32_t bits, uint32_t i) { return bits == 2 ? kW2[i] : (bits == 3 ? kW3[i] : kW4[i]); };
    for (uint32_t i = 0; i < 16; i++)
    {
        const uint32_t s = part ? part[i] : 0; const uint32_t* e0 = ep[2 * s]; const uint32_t* e1 = ep[2 * s + 1];
        uint32_t wc, wa;
        if (!b2) wc = wa = weight(b1, i1[i]);
        else if (!idxSel) { wc = weight(b1, i1[i]); wa = weight(b2, i2[i]); }
        else { wc = weight(b2, i2[i]); wa = weight(b1, i1[i]); }
        uint8_t px[4] = { lerp7(e0[0], e1[0], wc), lerp7(e0[1], e1[1], wc), lerp7(e0[2], e1[2], wc), lerp7(e0[3], e1[3], wa) };
        if (rotation) std::swap(px[3], px[rotation - 1]);
        memcpy(out[i], px, 4);
    }
}


// ---- BC6H (DXGI 95 UF16 / 96 SF16): 16 RGB half-float texels per 128-bit block, 14 modes (D3D11.3 functional spec 19.5.14 / Khronos Data Format 1.3 §21) ----------------------
// A mode's header is a sequence of bit fields of the endpoint components; written here as text, one token per field in bit order:  <component><endpoint>:<hi>-<lo> takes bits
// hi..lo of that component (hi < lo: the listed bits arrive most significant first, modes 13/14), "d" = the 5 partition bits.  Endpoints: 0 = A of subset 0 (the base the others
// are deltas of in transformed modes), 1 = B of subset 0, 2 / 3 = A / B of subset 1.
struct Bc6Mode { uint8_t modeBits, modeValue, regions, transformed, epBits, deltaBits[3]; const char* layout; };
const Bc6Mode kBc6Modes[14] = {
    { 2, 0x00, 2, 1, 10, { 5, 5, 5 }, "g2:4-4 b2:4-4 b3:4-4 r0:9-0 g0:9-0 b0:9-0 r1:4-0 g3:4-4 g2:3-0 g1:4-0 b3:0-0 g3:3-0 b1:4-0 b3:1-1 b2:3-0 r2:4-0 b3:2-2 r3:4-0 b3:3-3 d" },
    { 2, 0x01, 2, 1, 7, { 6, 6, 6 }, "g2:5-5 g3:4-4 g3:5-5 r0:6-0 b3:0-0 b3:1-1 b2:4-4 g0:6-0 b2:5-5 b3:2-2 g2:4-4 b0:6-0 b3:3-3 b3:5-5 b3:4-4 r1:5-0 g2:3-0 g1:5-0 g3:3-0 b1:5-0 b2:3-0 r2:5-0 r3:5-0 d" },
    { 5, 0x02, 2, 1, 11, { 5, 4, 4 }, "r0:9-0 g0:9-0 b0:9-0 r1:4-0 r0:10-10 g2:3-0 g1:3-0 g0:10-10 b3:0-0 g3:3-0 b1:3-0 b0:10-10 b3:1-1 b2:3-0 r2:4-0 b3:2-2 r3:4-0 b3:3-3 d" },
    { 5, 0x06, 2, 1, 11, { 4, 5, 4 }, "r0:9-0 g0:9-0 b0:9-0 r1:3-0 r0:10-10 g3:4-4 g2:3-0 g1:4-0 g0:10-10 g3:3-0 b1:3-0 b0:10-10 b3:1-1 b2:3-0 r2:3-0 b3:0-0 b3:2-2 r3:3-0 g2:4-4 b3:3-3 d" },
    { 5, 0x0A, 2, 1, 11, { 4, 4, 5 }, "r0:9-0 g0:9-0 b0:9-0 r1:3-0 r0:10-10 b2:4-4 g2:3-0 g1:3-0 g0:10-10 b3:0-0 g3:3-0 b1:4-0 b0:10-10 b2:3-0 r2:3-0 b3:1-1 b3:2-2 r3:3-0 b3:4-4 b3:3-3 d" },
    { 5, 0x0E, 2, 1, 9, { 5, 5, 5 }, "r0:8-0 b2:4-4 g0:8-0 g2:4-4 b0:8-0 b3:4-4 r1:4-0 g3:4-4 g2:3-0 g1:4-0 b3:0-0 g3:3-0 b1:4-0 b3:1-1 b2:3-0 r2:4-0 b3:2-2 r3:4-0 b3:3-3 d" },
    { 5, 0x12, 2, 1, 8, { 6, 5, 5 }, "r0:7-0 g3:4-4 b2:4-4 g0:7-0 b3:2-2 g2:4-4 b0:7-0 b3:3-3 b3:4-4 r1:5-0 g2:3-0 g1:4-0 b3:0-0 g3:3-0 b1:4-0 b3:1-1 b2:3-0 r2:5-0 r3:5-0 d" },
    { 5, 0x16, 2, 1, 8, { 5, 6, 5 }, "r0:7-0 b3:0-0 b2:4-4 g0:7-0 g2:5-5 g2:4-4 b0:7-0 g3:5-5 b3:4-4 r1:4-0 g3:4-4 g2:3-0 g1:5-0 g3:3-0 b1:4-0 b3:1-1 b2:3-0 r2:4-0 b3:2-2 r3:4-0 b3:3-3 d" },
    { 5, 0x1A, 2, 1, 8, { 5, 5, 6 }, "r0:7-0 b3:1-1 b2:4-4 g0:7-0 b2:5-5 g2:4-4 b0:7-0 b3:5-5 b3:4-4 r1:4-0 g3:4-4 g2:3-0 g1:4-0 b3:0-0 g3:3-0 b1:5-0 b2:3-0 r2:4-0 b3:2-2 r3:4-0 b3:3-3 d" },
    { 5, 0x1E, 2, 0, 6, { 6, 6, 6 }, "r0:5-0 g3:4-4 b3:0-0 b3:1-1 b2:4-4 g0:5-0 g2:5-5 b2:5-5 b3:2-2 g2:4-4 b0:5-0 g3:5-5 b3:3-3 b3:5-5 b3:4-4 r1:5-0 g2:3-0 g1:5-0 g3:3-0 b1:5-0 b2:3-0 r2:5-0 r3:5-0 d" },
    { 5, 0x03, 1, 0, 10, { 10, 10, 10 }, "r0:9-0 g0:9-0 b0:9-0 r1:9-0 g1:9-0 b1:9-0" },
    { 5, 0x07, 1, 1, 11, { 9, 9, 9 }, "r0:9-0 g0:9-0 b0:9-0 r1:8-0 r0:10-10 g1:8-0 g0:10-10 b1:8-0 b0:10-10" },
    { 5, 0x0B, 1, 1, 12, { 8, 8, 8 }, "r0:9-0 g0:9-0 b0:9-0 r1:7-0 r0:10-11 g1:7-0 g0:10-11 b1:7-0 b0:10-11" },
    { 5, 0x0F, 1, 1, 16, { 4, 4, 4 }, "r0:9-0 g0:9-0 b0:9-0 r1:3-0 r0:10-15 g1:3-0 g0:10-15 b1:3-0 b0:10-15" },
};
inline int32_t signExtend(int32_t v, int bits) { const int32_t m = 1 << (bits - 1); return (v ^ m) - m; }
inline int32_t bc6Unquantize(int32_t v, int bits, bool isSigned)
{
    if (!isSigned)
    {
        if (bits >= 15 || v == 0) return v;
        if (v == (1 << bits) - 1) return 0xFFFF;
        return ((v << 15) + 0x4000) >> (bits - 1);
    }
    if (bits >= 16) return v;
    const bool neg = v < 0; if (neg) v = -v;
    int32_t u;
    if (v == 0) u = 0; else if (v >= (1 << (bits - 1)) - 1) u = 0x7FFF; else u = ((v << 15) + 0x4000) >> (bits - 1);
    return neg ? -u : u;
}
inline uint16_t bc6Finish(int32_t v, bool isSigned)
{
    if (!isSigned) return uint16_t((v * 31) >> 6);
    const bool neg = v < 0; if (neg) v = -v;
    const uint16_t h = uint16_t((v * 31) >> 5);
    return neg ? uint16_t(0x8000u | h) : h;
}
// out: 16 texels x RGB as binary16 bit patterns; a reserved mode decodes to zero (D3D: "the decoder must return 0 in all channels")
void decodeBc6h(const uint8_t* blk, bool isSigned, uint16_t out[16][3])
{
    BitReader br{ blk };
    uint32_t mode = br.get(2); if (mode >= 2) mode |= br.get(3) << 2;
    const Bc6Mode* md = nullptr;
    for (const Bc6Mode& m : kBc6Modes) if (m.modeValue == mode) { md = &m; break; }
    if (!md) { memset(out, 0, 16 * 3 * 2); return; }
    int32_t ep[4][3] = {}; uint32_t partition = 0;
    for (const char* t = md->layout; *t; )
    {
        while (*t == ' ') t++;
        if (*t == 'd') { partition = br.get(5); t++; continue; }
        if (!*t) break;
        const int comp = *t == 'r' ? 0 : (*t == 'g' ? 1 : 2), e = t[1] - '0'; t += 3;
        int hi = 0, lo = 0; while (*t >= '0' && *t <= '9') hi = hi * 10 + (*t++ - '0'); t++; while (*t >= '0' && *t <= '9') lo = lo * 10 + (*t++ - '0');
        if (hi >= lo) { const uint32_t v = br.get(uint32_t(hi - lo + 1)); ep[e][comp] |= int32_t(v << lo); }
        else for (int b = lo; b >= hi; b--) ep[e][comp] |= int32_t(br.get(1) << b);          // listed bits arrive most significant first
    }
    const int n = md->regions * 2;
    if (isSigned) for (int c = 0; c < 3; c++) ep[0][c] = signExtend(ep[0][c], md->epBits);
    if (md->transformed || isSigned)
        for (int e = 1; e < n; e++) for (int c = 0; c < 3; c++) ep[e][c] = signExtend(ep[e][c], md->transformed ? md->deltaBits[c] : md->epBits);
    if (md->transformed)
        for (int e = 1; e < n; e++) for (int c = 0; c < 3; c++)
        {
            ep[e][c] = (ep[0][c] + ep[e][c]) & ((1 << md->epBits) - 1);
            if (isSigned) ep[e][c] = signExtend(ep[e][c], md->epBits);
        }
    for (int e = 0; e < n; e++) for (int c = 0; c < 3; c++) ep[e][c] = bc6Unquantize(ep[e][c], md->epBits, isSigned);
    const uint8_t* part = md->regions == 2 ? kBc7Partition2[partition] : nullptr;
    const uint32_t anchor1 = md->regions == 2 ? kBc7Anchor2[partition] : 0xFFu, ib = md->regions == 2 ? 3u : 4u;
    for (uint32_t i = 0; i < 16; i++)
    {
        const uint32_t s = part ? part[i] : 0u;
        const uint32_t idx = br.get((i == 0 || i == anchor1) ? ib - 1 : ib), w = ib == 3 ? kW3[idx] : kW4[idx];
        for (int c = 0; c < 3; c++) out[i][c] = bc6Finish((ep[2 * s][c] * int32_t(64 - w) + ep[2 * s + 1][c] * int32_t(w) + 32) >> 6, isSigned);
    }
}
float halfBitsToFloat(uint16_t h)
{
    const uint32_t sign = uint32_t(h & 0x8000u) << 16, e = (h >> 10) & 0x1Fu, m = h & 0x3FFu; uint32_t x;
    if (e == 0) { if (m == 0) x = sign; else { int sft = 0; uint32_t mm = m; while (!(mm & 0x400u)) { mm <<= 1; sft++; } x = sign | ((113u - sft) << 23) | ((mm & 0x3FFu) << 13); } }
    else if (e == 31) x = sign | 0x7F800000u | (m << 13);
    else x = sign | ((e + 112u) << 23) | (m << 13);
    float f; memcpy(&f, &x, 4); return f;
}

uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
// the header's mip count is untrusted: at most a full chain of the image and RTXPT_MAX_MIPS (so that `width >> m` never shifts by 32 or more)
uint32_t clampMipCount(uint32_t fromHeader, uint32_t w, uint32_t h) { uint32_t full = 1; for (uint32_t d = std::max(w, h); d > 1; d >>= 1) full++; return std::min(std::min(std::max(1u, fromHeader), full), uint32_t(RTXPT_MAX_MIPS)); }
enum Fmt { FmtNone, FmtBC1, FmtBC2, FmtBC3, FmtBC4, FmtBC5, FmtBC7, FmtRGBA8, FmtBGRA8 };

} // namespace

// Decodes a whole DDS file (first array slice / face only) into RGBA8 mips.  `wantSrgb` is the caller's view of the slot (base / emissive colour);
// the file's own *_SRGB format flag is reported in `srgb`.
DdsImage decodeDds(const uint8_t* data, size_t size, const char* name)
{
    if (size < 128 || memcmp(data, "DDS ", 4) != 0 || rd32(data + 4) != 124) failf("'%s' is not a DDS file", name);
    DdsImage img; img.height = rd32(data + 12); img.width = rd32(data + 16);
    const uint32_t mipCount = clampMipCount(rd32(data + 28), rd32(data + 16), rd32(data + 12)), pfFlags = rd32(data + 80), fourCC = rd32(data + 84);
    size_t off = 128; Fmt fmt = FmtNone;
    auto cc = [](const char* s) { return uint32_t(uint8_t(s[0])) | (uint32_t(uint8_t(s[1])) << 8) | (uint32_t(uint8_t(s[2])) << 16) | (uint32_t(uint8_t(s[3])) << 24); };
    if (pfFlags & 0x4)
    {
        if (fourCC == cc("DX10"))
        {
            if (size < 148) failf("DDS '%s': truncated DX10 header", name);
            const uint32_t dxgi = rd32(data + 128); off = 148;
            switch (dxgi)
            {
            case 70: case 71: fmt = FmtBC1; break; case 72: fmt = FmtBC1; img.srgb = true; break;
            case 73: case 74: fmt = FmtBC2; break; case 75: fmt = FmtBC2; img.srgb = true; break;
            case 76: case 77: fmt = FmtBC3; break; case 78: fmt = FmtBC3; img.srgb = true; break;
            case 79: case 80: fmt = FmtBC4; break; case 82: case 83: fmt = FmtBC5; break;
            case 97: case 98: fmt = FmtBC7; break; case 99: fmt = FmtBC7; img.srgb = true; break;
            case 27: case 28: fmt = FmtRGBA8; break; case 29: fmt = FmtRGBA8; img.srgb = true; break;
            case 87: fmt = FmtBGRA8; break; case 91: fmt = FmtBGRA8; img.srgb = true; break;
            case 94: case 95: case 96: failf("DDS '%s': BC6H is an HDR format: load it with rtxpt_b200_load_dds_hdr (environment maps), not as a material texture", name);
            default: failf("DDS '%s': DXGI format %u is not supported", name, dxgi);
            }
        }
        else if (fourCC == cc("DXT1")) fmt = FmtBC1; else if (fourCC == cc("DXT2") || fourCC == cc("DXT3")) fmt = FmtBC2; else if (fourCC == cc("DXT4") || fourCC == cc("DXT5")) fmt = FmtBC3;
        else if (fourCC == cc("ATI1") || fourCC == cc("BC4U")) fmt = FmtBC4; else if (fourCC == cc("ATI2") || fourCC == cc("BC5U")) fmt = FmtBC5;
        else failf("DDS '%s': FourCC format is not supported", name);
    }
    else if ((pfFlags & 0x40) && rd32(data + 88) == 32)
    {
        const uint32_t rm = rd32(data + 92), bm = rd32(data + 100);
        if (rm == 0x000000FFu && bm == 0x00FF0000u) fmt = FmtRGBA8; else if (rm == 0x00FF0000u && bm == 0x000000FFu) fmt = FmtBGRA8; else failf("DDS '%s': unsupported channel masks", name);
    }
    else failf("DDS '%s': unsupported pixel format", name);
    if (!img.width || !img.height || img.width > 32768 || img.height > 32768) failf("DDS '%s': bad dimensions", name);
    for (uint32_t m = 0; m < mipCount; m++)
    {
        const uint32_t w = std::max(1u, img.width >> m), h = std::max(1u, img.height >> m);
        std::vector<uint8_t> rgba(size_t(w) * h * 4);
        if (fmt == FmtRGBA8 || fmt == FmtBGRA8)
        {
            const size_t bytes = size_t(w) * h * 4; if (off + bytes > size) failf("DDS '%s': truncated", name);
            memcpy(rgba.data(), data + off, bytes); off += bytes;
            if (fmt == FmtBGRA8) for (size_t i = 0; i < size_t(w) * h; i++) std::swap(rgba[4 * i], rgba[4 * i + 2]);
        }
        else
        {
            const uint32_t bw = (w + 3) / 4, bh = (h + 3) / 4; const size_t blockBytes = (fmt == FmtBC1 || fmt == FmtBC4) ? 8 : 16;
            if (off + size_t(bw) * bh * blockBytes > size) failf("DDS '%s': truncated", name);
            for (uint32_t by = 0; by < bh; by++) for (uint32_t bx = 0; bx < bw; bx++)
            {
                const uint8_t* b = data + off + (size_t(by) * bw + bx) * blockBytes; uint8_t px[16][4];
                switch (fmt)
                {
                case FmtBC1: decodeBc1Color(b, true, px); break;
                case FmtBC2: decodeBc1Color(b + 8, false, px); for (int i = 0; i < 16; i++) { const uint32_t a = (b[i >> 1] >> ((i & 1) * 4)) & 15; px[i][3] = uint8_t(a * 17); } break;
                case FmtBC3: { decodeBc1Color(b + 8, false, px); uint8_t a[16]; decodeBc4Channel(b, a); for (int i = 0; i < 16; i++) px[i][3] = a[i]; break; }
                case FmtBC4: { uint8_t r[16]; decodeBc4Channel(b, r); for (int i = 0; i < 16; i++) { px[i][0] = px[i][1] = px[i][2] = r[i]; px[i][3] = 255; } break; }
                case FmtBC5: { uint8_t r[16], g[16]; decodeBc4Channel(b, r); decodeBc4Channel(b + 8, g); for (int i = 0; i < 16; i++) { px[i][0] = r[i]; px[i][1] = g[i]; px[i][2] = 0; px[i][3] = 255; } break; }
                case FmtBC7: decodeBc7(b, px); break;
                default: break;
                }
                for (uint32_t y = 0; y < 4 && by * 4 + y < h; y++) for (uint32_t x = 0; x < 4 && bx * 4 + x < w; x++) memcpy(&rgba[(size_t(by * 4 + y) * w + bx * 4 + x) * 4], px[y * 4 + x], 4);
            }
            off += size_t(bw) * bh * blockBytes;
        }
        img.mips.push_back(std::move(rgba));
    }
    return img;
}

// BC1 / BC2 / BC3 / BC7 files handed through without decoding (RTXPT_FORMAT_BC*): the raw blocks of every mip; false for every other format (the caller decodes those)
struct DdsBlocks { uint32_t width = 0, height = 0, format = 0; bool srgb = false; std::vector<std::vector<uint8_t>> mips; };
bool extractDdsBlocks(const uint8_t* data, size_t size, const char* name, DdsBlocks& out)
{
    if (size < 128 || memcmp(data, "DDS ", 4) != 0 || rd32(data + 4) != 124) return false;
    out.height = rd32(data + 12); out.width = rd32(data + 16);
    const uint32_t mipCount = clampMipCount(rd32(data + 28), rd32(data + 16), rd32(data + 12)), pfFlags = rd32(data + 80), fourCC = rd32(data + 84);
    auto cc = [](const char* s) { return uint32_t(uint8_t(s[0])) | (uint32_t(uint8_t(s[1])) << 8) | (uint32_t(uint8_t(s[2])) << 16) | (uint32_t(uint8_t(s[3])) << 24); };
    if (!(pfFlags & 0x4)) return false;
    size_t off = 128; uint32_t fmt = 0;
    if (fourCC == cc("DX10"))
    {
        if (size < 148) return false;
        off = 148;
        switch (rd32(data + 128)) { case 70: case 71: fmt = RTXPT_FORMAT_BC1_UNORM; break; case 72: fmt = RTXPT_FORMAT_BC1_UNORM; out.srgb = true; break; case 73: case 74: fmt = RTXPT_FORMAT_BC2_UNORM; break;
                                    case 75: fmt = RTXPT_FORMAT_BC2_UNORM; out.srgb = true; break; case 76: case 77: fmt = RTXPT_FORMAT_BC3_UNORM; break; case 78: fmt = RTXPT_FORMAT_BC3_UNORM; out.srgb = true; break;
                                    case 97: case 98: fmt = RTXPT_FORMAT_BC7_UNORM; break; case 99: fmt = RTXPT_FORMAT_BC7_UNORM; out.srgb = true; break; default: return false; }
    }
    else if (fourCC == cc("DXT1")) fmt = RTXPT_FORMAT_BC1_UNORM; else if (fourCC == cc("DXT2") || fourCC == cc("DXT3")) fmt = RTXPT_FORMAT_BC2_UNORM; else if (fourCC == cc("DXT4") || fourCC == cc("DXT5")) fmt = RTXPT_FORMAT_BC3_UNORM;
    else return false;
    if (!out.width || !out.height || out.width > 32768 || out.height > 32768) failf("DDS '%s': bad dimensions", name);
    if ((out.width & 3u) || (out.height & 3u)) return false;              // CUDA's block-compressed arrays want whole blocks at mip 0: odd sizes take the decoded path
    out.format = fmt; const size_t blockBytes = fmt == RTXPT_FORMAT_BC1_UNORM ? 8 : 16;
    for (uint32_t m = 0; m < mipCount && m < RTXPT_MAX_MIPS; m++)
    {
        const uint32_t w = std::max(1u, out.width >> m), h = std::max(1u, out.height >> m); const size_t bytes = size_t((w + 3) / 4) * ((h + 3) / 4) * blockBytes;
        if (off + bytes > size) failf("DDS '%s': truncated", name);
        out.mips.emplace_back(data + off, data + off + bytes); off += bytes;
    }
    return true;
}
// one mip of raw blocks -> RGBA8 (the opacity-mask baker reads the alpha of mip 0 of a texture that stays compressed on the device)
void decodeBlocksToRgba8(uint32_t format, const uint8_t* blocks, uint32_t w, uint32_t h, std::vector<uint8_t>& rgba)
{
    rgba.assign(size_t(w) * h * 4, 0);
    const uint32_t bw = (w + 3) / 4, bh = (h + 3) / 4; const size_t blockBytes = (format == RTXPT_FORMAT_BC1_UNORM || format == RTXPT_FORMAT_BC1_SRGB) ? 8 : 16;
    for (uint32_t by = 0; by < bh; by++) for (uint32_t bx = 0; bx < bw; bx++)
    {
        const uint8_t* b = blocks + (size_t(by) * bw + bx) * blockBytes; uint8_t px[16][4];
        switch (format)
        {
        case RTXPT_FORMAT_BC1_UNORM: case RTXPT_FORMAT_BC1_SRGB: decodeBc1Color(b, true, px); break;
        case RTXPT_FORMAT_BC2_UNORM: case RTXPT_FORMAT_BC2_SRGB: decodeBc1Color(b + 8, false, px); for (int i = 0; i < 16; i++) { const uint32_t a = (b[i >> 1] >> ((i & 1) * 4)) & 15; px[i][3] = uint8_t(a * 17); } break;
        case RTXPT_FORMAT_BC3_UNORM: case RTXPT_FORMAT_BC3_SRGB: { decodeBc1Color(b + 8, false, px); uint8_t a[16]; decodeBc4Channel(b, a); for (int i = 0; i < 16; i++) px[i][3] = a[i]; break; }
        default: decodeBc7(b, px); break;
        }
        for (uint32_t y = 0; y < 4 && by * 4 + y < h; y++) for (uint32_t x = 0; x < 4 && bx * 4 + x < w; x++) memcpy(&rgba[(size_t(by * 4 + y) * w + bx * 4 + x) * 4], px[y * 4 + x], 4);
    }
}

// HDR DDS files - what the reference's environment maps are (Assets/EnvironmentMaps/*_cube_bc6u.dds: BC6H_UF16 cubes, loaded by Donut's DDSFile.cpp and fed to
// EnvMapBaker::Update, Rtxpt/Lighting/Distant/EnvMapBaker.cpp:164-169, :372-375): BC6H UF16 / SF16, RGBA16F, RGBA32F; 2-D or cube; mip 0 of every face as RGBA32F.
struct HdrImage { uint32_t width = 0, height = 0, faces = 1, mipCount = 1; std::vector<float> rgba; };      // faces back to back, D3D order +x -x +y -y +z -z
HdrImage decodeDdsHdr(const uint8_t* data, size_t size, const char* name)
{
    if (size < 128 || memcmp(data, "DDS ", 4) != 0 || rd32(data + 4) != 124) failf("'%s' is not a DDS file", name);
    HdrImage img; img.height = rd32(data + 12); img.width = rd32(data + 16); img.mipCount = clampMipCount(rd32(data + 28), img.width, img.height);
    const uint32_t pfFlags = rd32(data + 80), fourCC = rd32(data + 84), caps2 = rd32(data + 112);
    auto cc = [](const char* s) { return uint32_t(uint8_t(s[0])) | (uint32_t(uint8_t(s[1])) << 8) | (uint32_t(uint8_t(s[2])) << 16) | (uint32_t(uint8_t(s[3])) << 24); };
    enum { BC6U, BC6S, F16, F32 } fmt; size_t off = 128;
    if (!(pfFlags & 0x4)) failf("DDS '%s': not an HDR format (no FourCC)", name);
    if (fourCC == cc("DX10"))
    {
        if (size < 148) failf("DDS '%s': truncated DX10 header", name);
        const uint32_t dxgi = rd32(data + 128), misc = rd32(data + 136), arraySize = std::max(1u, rd32(data + 140)); off = 148;
        switch (dxgi) { case 94: case 95: fmt = BC6U; break; case 96: fmt = BC6S; break; case 10: fmt = F16; break; case 2: fmt = F32; break; default: failf("DDS '%s': DXGI format %u is not an HDR format this loader reads", name, dxgi); }
        img.faces = arraySize * ((misc & 0x4u) ? 6u : 1u);
    }
    else
    {
        if (fourCC == 113) fmt = F16; else if (fourCC == 116) fmt = F32; else failf("DDS '%s': FourCC is not an HDR format this loader reads", name);      // D3DFMT_A16B16G16R16F / A32B32G32R32F
        if (caps2 & 0x200u) img.faces = 6;
    }
    if (!img.width || !img.height || img.width > 16384 || img.height > 16384 || img.faces > 6 * 64) failf("DDS '%s': bad dimensions", name);
    const uint32_t mips = std::min(img.mipCount, 15u);
    const size_t texels = size_t(img.width) * img.height; img.rgba.resize(texels * 4 * img.faces);
    for (uint32_t f = 0; f < img.faces; f++)
    {
        float* dst = img.rgba.data() + size_t(f) * texels * 4;
        for (uint32_t m = 0; m < mips; m++)
        {
            const uint32_t w = std::max(1u, img.width >> m), h = std::max(1u, img.height >> m);
            const size_t bytes = (fmt == BC6U || fmt == BC6S) ? size_t((w + 3) / 4) * ((h + 3) / 4) * 16 : size_t(w) * h * (fmt == F16 ? 8 : 16);
            if (off + bytes > size) failf("DDS '%s': truncated", name);
            if (m == 0)
            {
                if (fmt == F32) memcpy(dst, data + off, bytes);
                else if (fmt == F16) for (size_t i = 0; i < texels * 4; i++) { uint16_t hv; memcpy(&hv, data + off + i * 2, 2); dst[i] = halfBitsToFloat(hv); }
                else
                {
                    const uint32_t bw = (w + 3) / 4, bh = (h + 3) / 4;
                    for (uint32_t by = 0; by < bh; by++) for (uint32_t bx = 0; bx < bw; bx++)
                    {
                        uint16_t px[16][3]; decodeBc6h(data + off + (size_t(by) * bw + bx) * 16, fmt == BC6S, px);
                        for (uint32_t y = 0; y < 4 && by * 4 + y < h; y++) for (uint32_t x = 0; x < 4 && bx * 4 + x < w; x++)
                        { float* t = dst + (size_t(by * 4 + y) * w + bx * 4 + x) * 4; for (int c = 0; c < 3; c++) t[c] = halfBitsToFloat(px[y * 4 + x][c]); t[3] = 1.0f; }
                    }
                }
            }
            off += bytes;
        }
    }
    return img;
}

} // namespace rtxpt_host

static thread_local std::string g_ddsError;
// HDR DDS (environment maps): sizes with outRGBA32F == NULL, then mip 0 of every face as RGBA32F, faces back to back - the `source` of RtxptEnvBakeDesc (sourceType 2 for a cube)
extern "C" RTXPT_API int rtxpt_b200_load_dds_hdr(const void* fileBytes, uint64_t fileSize, uint32_t* outWidth, uint32_t* outHeight, uint32_t* outFaces, uint32_t* outMipCount,
                                                 float* outRGBA32F, uint64_t outCapacityFloats)
{
    if (!fileBytes || !outWidth || !outHeight || !outFaces) return RTXPT_ERR_INVALID_ARGUMENT;
    try
    {
        const rtxpt_host::HdrImage img = rtxpt_host::decodeDdsHdr(static_cast<const uint8_t*>(fileBytes), size_t(fileSize), "<memory>");
        *outWidth = img.width; *outHeight = img.height; *outFaces = img.faces; if (outMipCount) *outMipCount = img.mipCount;
        if (outRGBA32F) { if (outCapacityFloats < img.rgba.size()) { g_ddsError = "output buffer too small"; return RTXPT_ERR_INVALID_ARGUMENT; } memcpy(outRGBA32F, img.rgba.data(), img.rgba.size() * sizeof(float)); }
    }
    catch (const rtxpt_host::LoadError& e) { g_ddsError = e.msg; return RTXPT_ERR_INVALID_ARGUMENT; }
    return RTXPT_OK;
}
// Inspection hook: decodes mip `mip` of a DDS file image held in memory into RGBA8 (tests/test_dds.py compares it with an independent decoder)
extern "C" RTXPT_API int rtxpt_b200_debug_decode_dds(const void* fileBytes, uint64_t fileSize, uint32_t mip, uint32_t* outWidth, uint32_t* outHeight, uint32_t* outMipCount, uint32_t* outSrgb,
                                                     uint8_t* outRGBA, uint64_t outCapacity)
{
    if (!fileBytes || !outWidth || !outHeight) return RTXPT_ERR_INVALID_ARGUMENT;
    try
    {
        const rtxpt_host::DdsImage img = rtxpt_host::decodeDds(static_cast<const uint8_t*>(fileBytes), size_t(fileSize), "<memory>");
        if (mip >= img.mips.size()) { g_ddsError = "mip out of range"; return RTXPT_ERR_INVALID_ARGUMENT; }
        *outWidth = std::max(1u, img.width >> mip); *outHeight = std::max(1u, img.height >> mip);
        if (outMipCount) *outMipCount = uint32_t(img.mips.size());
        if (outSrgb) *outSrgb = img.srgb ? 1u : 0u;
        if (outRGBA) { if (outCapacity < img.mips[mip].size()) { g_ddsError = "output buffer too small"; return RTXPT_ERR_INVALID_ARGUMENT; } memcpy(outRGBA, img.mips[mip].data(), img.mips[mip].size()); }
    }
    catch (const rtxpt_host::LoadError& e) { g_ddsError = e.msg; return RTXPT_ERR_INVALID_ARGUMENT; }
    return RTXPT_OK;
}
extern "C" RTXPT_API const char* rtxpt_b200_debug_decode_dds_error(void) { return g_ddsError.c_str(); }
