// device_math.cuh — vector helpers, fp16 payload packing and the stateless integer sample generators used by every kernel.
// Follows Rtxpt/Shaders/PathTracer/Utils/{Packing,Utils,NoiseAndSequences,StatelessSampleGenerators,SampleGenerators}.hlsli and
// Utils/Math/MathHelpers.hlsli, Utils/Geometry.hlsli (citations at each function).
#pragma once
#include <string.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

#ifndef PT_DEVICE      // tests/emu/shade_host_emu.cu builds the shading functions for the host (test infrastructure): it defines PT_DEVICE as __host__ __device__ before this header
#define PT_DEVICE __device__ __forceinline__
#endif
#define PT_HD __host__ __device__ __forceinline__

namespace pt {

typedef uint32_t uint;

constexpr float kPi = 3.14159265358979323846f;
constexpr float k2Pi = 6.28318530717958647692f;
constexpr float k1OverPi = 0.31830988618379067153f;
constexpr float k2OverPi = 0.63661977236758134308f;
constexpr float kPiOver2 = 1.57079632679489661923f;
constexpr float kPiOver4 = 0.78539816339744830961f;
constexpr float kHalfMax = 65504.0f;
constexpr float kFltMax = 3.402823466e+38f;
constexpr float kFltMin = 1.175494351e-38f;
constexpr float kMaxRayTravel = 1e15f;                      // Rtxpt/Shaders/PathTracer/Config.h:86

// ---- float3 --------------------------------------------------------------------------------------------------------------
PT_HD float3 mk3(float x, float y, float z) { return make_float3(x, y, z); }
PT_HD float3 mk3(float s) { return make_float3(s, s, s); }
PT_HD float3 operator+(float3 a, float3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
PT_HD float3 operator-(float3 a, float3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
PT_HD float3 operator-(float3 a) { return mk3(-a.x, -a.y, -a.z); }
PT_HD float3 operator*(float3 a, float3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
PT_HD float3 operator*(float3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
PT_HD float3 operator*(float s, float3 a) { return mk3(a.x * s, a.y * s, a.z * s); }
#if defined(PT_FAST_MATH)
PT_HD float3 operator/(float3 a, float s) { const float r = 1.0f / s; return mk3(a.x * r, a.y * r, a.z * r); }      // one MUFU.RCP instead of three scaled divisions
#else
PT_HD float3 operator/(float3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }
#endif
PT_HD float3 operator/(float3 a, float3 b) { return mk3(a.x / b.x, a.y / b.y, a.z / b.z); }
PT_HD float dot3(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
PT_HD float3 cross3(float3 a, float3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
PT_HD float len3(float3 a) { return sqrtf(dot3(a, a)); }
PT_HD float3 norm3(float3 a) { return a / len3(a); }
PT_HD float2 mk2(float x, float y) { return make_float2(x, y); }
PT_HD float sat(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }         // NaN -> 0, like HLSL saturate
PT_HD float3 sat3(float3 v) { return mk3(sat(v.x), sat(v.y), sat(v.z)); }
PT_HD float clampf(float x, float a, float b) { return fminf(fmaxf(x, a), b); }
PT_HD float lerpf(float a, float b, float t) { return a + (b - a) * t; }
PT_HD float3 lerp3(float3 a, float3 b, float t) { return a + (b - a) * t; }
PT_HD float sgn(float x) { return (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f); }
PT_HD bool anyPositive(float3 v) { return v.x > 0 || v.y > 0 || v.z > 0; }
PT_HD float luminance(float3 c) { return dot3(c, mk3(0.2126f, 0.7152f, 0.0722f)); }      // Utils/Utils.hlsli:51
PT_HD float average(float3 c) { return (c.x + c.y + c.z) / 3.0f; }                      // Utils/Utils.hlsli:57
PT_HD float maxComp(float3 c) { return fmaxf(fmaxf(c.x, c.y), c.z); }                   // Utils/ColorHelpers.hlsli:19-27

// bit casts usable from __host__ __device__ bodies (the host builds of tests/emu)
#ifdef __CUDA_ARCH__
PT_HD uint floatBits(float f) { return __float_as_uint(f); }
PT_HD float bitsToFloat(uint u) { return __uint_as_float(u); }
#else
PT_HD uint floatBits(float f) { uint u; memcpy(&u, &f, 4); return u; }
PT_HD float bitsToFloat(uint u) { float f; memcpy(&f, &u, 4); return f; }
#endif
// OctToNDirUnorm32 (Utils.hlsli:128-153; the [0,1] mapping is applied twice on both sides, see lights_bake.cpp)
PT_HD float3 octUnorm32ToDir(uint p)
{
    float fx = sat(float(p & 0xffffu) / float(0xfffe)) * 2.0f - 1.0f, fy = sat(float(p >> 16) / float(0xfffe)) * 2.0f - 1.0f;
    fx = fx * 2.0f - 1.0f; fy = fy * 2.0f - 1.0f;
    float3 n = mk3(fx, fy, 1.0f - fabsf(fx) - fabsf(fy));
    const float t = sat(-n.z);
    n.x += (n.x >= 0.0f) ? -t : t; n.y += (n.y >= 0.0f) ? -t : t;
    return norm3(n);
}

// ---- fp16 storage (RNE; equals the oracle's f32tof16) ---------------------------------------------------------------------
PT_HD uint f32tof16(float v) { return (uint)__half_as_ushort(__float2half_rn(v)); }
PT_HD float f16tof32(uint h) { return __half2float(__ushort_as_half((unsigned short)(h & 0xFFFF))); }
PT_DEVICE float lp(float v) { return __half2float(__float2half_rn(v)); }                 // one "lpfloat" store
PT_DEVICE float3 lp3(float3 v) { return mk3(lp(v.x), lp(v.y), lp(v.z)); }
PT_DEVICE uint packHalf2NoClamp(float a, float b) { return f32tof16(a) | (f32tof16(b) << 16); }          // Fp32ToFp16NoClamp, Packing.hlsli:212
PT_DEVICE uint packHalf2Clamp(float a, float b) { return packHalf2NoClamp(clampf(a, -kHalfMax, kHalfMax), clampf(b, -kHalfMax, kHalfMax)); }  // Packing.hlsli:206

PT_HD float unpackSnorm8(uint v) { int s = int(v << 24) >> 24; return clampf(float(s) / 127.0f, -1.0f, 1.0f); }  // Packing.hlsli:127
PT_HD float unpackUnorm8(uint v) { return float(v & 0xFFu) / 255.0f; }

// ---- fast approximations the reference uses on purpose (Utils/Utils.hlsli:486-499) ------------------------------------------
PT_DEVICE float fastSqrt(float x) { return __int_as_float(0x1fbd1df5 + (__float_as_int(x) >> 1)); }
PT_DEVICE float fastACos(float inX)
{
    float x = fabsf(inX);
    float res = -0.156583f * x + 1.570796f;
    res *= fastSqrt(1.0f - x);
    return (inX >= 0) ? res : 3.141593f - res;
}

// ---- integer hashing / Owen-scrambled Sobol (NoiseAndSequences.hlsli:58-84, :130-229) --------------------------------------
PT_HD uint hash32(uint x) { x ^= x >> 16; x *= 0x21f0aaadu; x ^= x >> 15; x *= 0xf35a2d97u; x ^= x >> 15; return x; }
PT_HD uint hash32Combine(uint seed, uint value) { return seed ^ (hash32(value) + 0x9e3779b9u + (seed << 6) + (seed >> 2)); }
PT_HD float hashToFloat(uint h) { return float(h >> 8) * 5.9604644775390625e-8f; }      // / 2^24, exact

// Sobol direction numbers for dimensions 1..4 (dimension 0 is bit reversal); NoiseAndSequences.hlsli:135-180
static __constant__ uint cSobolDirections[4][32] = {
    { 0x80000000, 0xc0000000, 0xa0000000, 0xf0000000, 0x88000000, 0xcc000000, 0xaa000000, 0xff000000,
      0x80800000, 0xc0c00000, 0xa0a00000, 0xf0f00000, 0x88880000, 0xcccc0000, 0xaaaa0000, 0xffff0000,
      0x80008000, 0xc000c000, 0xa000a000, 0xf000f000, 0x88008800, 0xcc00cc00, 0xaa00aa00, 0xff00ff00,
      0x80808080, 0xc0c0c0c0, 0xa0a0a0a0, 0xf0f0f0f0, 0x88888888, 0xcccccccc, 0xaaaaaaaa, 0xffffffff },
    { 0x80000000, 0xc0000000, 0x60000000, 0x90000000, 0xe8000000, 0x5c000000, 0x8e000000, 0xc5000000,
      0x68800000, 0x9cc00000, 0xee600000, 0x55900000, 0x80680000, 0xc09c0000, 0x60ee0000, 0x90550000,
      0xe8808000, 0x5cc0c000, 0x8e606000, 0xc5909000, 0x6868e800, 0x9c9c5c00, 0xeeee8e00, 0x5555c500,
      0x8000e880, 0xc0005cc0, 0x60008e60, 0x9000c590, 0xe8006868, 0x5c009c9c, 0x8e00eeee, 0xc5005555 },
    { 0x80000000, 0xc0000000, 0x20000000, 0x50000000, 0xf8000000, 0x74000000, 0xa2000000, 0x93000000,
      0xd8800000, 0x25400000, 0x59e00000, 0xe6d00000, 0x78080000, 0xb40c0000, 0x82020000, 0xc3050000,
      0x208f8000, 0x51474000, 0xfbea2000, 0x75d93000, 0xa0858800, 0x914e5400, 0xdbe79e00, 0x25db6d00,
      0x58800080, 0xe54000c0, 0x79e00020, 0xb6d00050, 0x800800f8, 0xc00c0074, 0x200200a2, 0x50050093 },
    { 0x80000000, 0x40000000, 0x20000000, 0xb0000000, 0xf8000000, 0xdc000000, 0x7a000000, 0x9d000000,
      0x5a800000, 0x2fc00000, 0xa1600000, 0xf0b00000, 0xda880000, 0x6fc40000, 0x81620000, 0x40bb0000,
      0x22878000, 0xb3c9c000, 0xfb65a000, 0xddb2d000, 0x78022800, 0x9c0b3c00, 0x5a0fb600, 0x2d0ddb00,
      0xa2878080, 0xf3c9c040, 0xdb65a020, 0x6db2d0b0, 0x800228f8, 0x400b3cdc, 0x200fb67a, 0xb00ddb9d },
};
PT_DEVICE uint sobolDimBitwise(uint index, uint dim /*1..4*/)
{
    uint X = 0;
    #pragma unroll 8
    for (uint bit = 0; bit < 32; bit++)
        X ^= ((index >> bit) & 1u) ? cSobolDirections[dim - 1][bit] : 0u;
    return X;
}
#if defined(PT_SOBOL_TABLES)
// The Sobol matrix product is linear over GF(2): four byte-indexed partial products replace the 32-step loop (same bits out).
// Filled once per context by k_init_sobol_tables (shade_kernels.cu).
static __device__ uint gSobolByte[4][4][256];
PT_DEVICE uint sobolDim(uint index, uint dim /*1..4*/)
{
    const uint* T = &gSobolByte[dim - 1][0][0];
    return __ldg(T + (index & 0xFFu)) ^ __ldg(T + 256 + ((index >> 8) & 0xFFu)) ^ __ldg(T + 512 + ((index >> 16) & 0xFFu)) ^ __ldg(T + 768 + (index >> 24));
}
#else
PT_DEVICE uint sobolDim(uint index, uint dim) { return sobolDimBitwise(index, dim); }
#endif
PT_DEVICE uint owenHash(uint x, uint seed) { x ^= x * 0x3d20adeau; x += seed; x *= (seed >> 16) | 1u; x ^= x * 0x05526c56u; x ^= x * 0x53a22864u; return x; }
PT_DEVICE uint owenScramble(uint x, uint seed) { return __brev(owenHash(__brev(x), seed)); }

// SampleGeneratorVertexBase::make (StatelessSampleGenerators.hlsli:27-49)
PT_HD uint vertexBaseHash(uint packedPixel, uint vertexIndex) { return hash32Combine(hash32(vertexIndex + 0x035F9F29u), packedPixel); }

// UniformSampleSequenceGenerator (StatelessSampleGenerators.hlsli:187-232): state is one u32
struct UniformSeq
{
    uint h;
    PT_HD static UniformSeq make(uint baseHash, uint sampleIndex, uint effectSeed)
    {
        UniformSeq s; s.h = hash32Combine(hash32Combine(baseHash, effectSeed), sampleIndex); return s;
    }
    PT_HD float next() { h = hash32(h); return hashToFloat(h); }
    PT_DEVICE uint nextBits() { h = hash32(h); return h; }
};
// SampleSequenceGenerator::Generate, low-discrepancy branch (StatelessSampleGenerators.hlsli:150-181), dimension `dim` of sample `sampleIndex`
PT_DEVICE uint ldSampleBits(uint baseHash, uint sampleIndex, uint effectSeed, uint dim)
{
    uint currentHash = hash32Combine(baseHash, effectSeed);
    uint shuffleSeed = hash32Combine(currentHash, 0);
    uint dimSeed = hash32Combine(currentHash, 1 + dim);
    uint shuffled = owenScramble(sampleIndex, shuffleSeed);
    uint s = (dim == 0) ? __brev(shuffled) : sobolDim(shuffled, dim);
    return owenScramble(s, dimSeed);
}

// ---- mappings / sampling (Utils/Math/MathHelpers.hlsli, Utils/Geometry.hlsli) --------------------------------------------------
PT_DEVICE float2 dirToOctEqualArea(float3 n)      // ndir_to_oct_equal_area_unorm, MathHelpers.hlsli:185-200
{
    float r = sqrtf(1.f - fabsf(n.z));
    float phi = atan2f(fabsf(n.y), fabsf(n.x));
    float py = r * phi * k2OverPi;
    float px = r - py;
    if (n.z < 0.f) { float t = px; px = 1.f - py; py = 1.f - t; }
    px *= sgn(n.x); py *= sgn(n.y);
    return mk2(sat(px * 0.5f + 0.5f), sat(py * 0.5f + 0.5f));
}
PT_DEVICE float3 octEqualAreaToDir(float2 p)      // oct_to_ndir_equal_area_unorm, MathHelpers.hlsli:207-227
{
    float px = p.x * 2.f - 1.f, py = p.y * 2.f - 1.f;
    float d = 1.f - (fabsf(px) + fabsf(py));
    float r = 1.f - fabsf(d);
    float phi = (r > 0.f) ? ((fabsf(py) - fabsf(px)) / r + 1.f) * kPiOver4 : 0.f;
    float f = r * sqrtf(2.f - r * r);
    return mk3(f * sgn(px) * cosf(phi), f * sgn(py) * sinf(phi), sgn(d) * (1.f - r * r));
}
PT_HD float2 sampleDiskPolar(float u0, float u1) { float r = sqrtf(u0); float phi = k2Pi * u1; return mk2(r * cosf(phi), r * sinf(phi)); }   // MathHelpers.hlsli:238
PT_DEVICE float3 sampleCosineHemisphereConcentric(float u0, float u1, float& pdf)       // MathHelpers.hlsli:288-320
{
    float ux = 2.f * u0 - 1.f, uy = 2.f * u1 - 1.f;
    float dx, dy;
    if (ux == 0.f && uy == 0.f) { dx = ux; dy = uy; }
    else
    {
        float phi, r;
        if (fabsf(ux) > fabsf(uy)) { r = ux; phi = (uy / ux) * kPiOver4; }
        else                       { r = uy; phi = kPiOver2 - (ux / uy) * kPiOver4; }
        dx = r * cosf(phi); dy = r * sinf(phi);
    }
    float z = sqrtf(fmaxf(0.f, 1.f - (dx * dx + dy * dy)));
    pdf = z * k1OverPi;
    return mk3(dx, dy, z);
}
PT_DEVICE float3 perpStark(float3 u)              // MathHelpers.hlsli:436-448
{
    float ax = fabsf(u.x), ay = fabsf(u.y), az = fabsf(u.z);
    uint uyx = (ax - ay) < 0 ? 1 : 0, uzx = (ax - az) < 0 ? 1 : 0, uzy = (ay - az) < 0 ? 1 : 0;
    uint xm = uyx & uzx, ym = (1 ^ xm) & uzy, zm = 1 ^ (xm | ym);
    return norm3(cross3(u, mk3(float(xm), float(ym), float(zm))));
}
PT_DEVICE float pdfAreaToSolidAngle(float pdfA, float dist, float cosTheta) { return pdfA * (dist * dist) / fmaxf(cosTheta, 2e-9f); }   // Geometry.hlsli:79

// row-major float3x4 (HLSL mul(M, float4(v,1|0)) and mul(v, (float3x3)M))
PT_HD float3 xfPoint(const float* m, float3 v) { return mk3(m[0] * v.x + m[1] * v.y + m[2] * v.z + m[3], m[4] * v.x + m[5] * v.y + m[6] * v.z + m[7], m[8] * v.x + m[9] * v.y + m[10] * v.z + m[11]); }
PT_HD float3 xfVector(const float* m, float3 v) { return mk3(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[4] * v.x + m[5] * v.y + m[6] * v.z, m[8] * v.x + m[9] * v.y + m[10] * v.z); }
PT_HD float3 rowVecTimes3x3(float3 v, const float* m) { return mk3(v.x * m[0] + v.y * m[4] + v.z * m[8], v.x * m[1] + v.y * m[5] + v.z * m[9], v.x * m[2] + v.y * m[6] + v.z * m[10]); }

// Wächter & Binder self-intersection offset (PathTracerHelpers.hlsli:29-42)
PT_DEVICE float3 offsetRayOrigin(float3 p, float3 n)
{
    const float origin = 1.f / 16.f, fScale = 3.f / 65536.f, iScale = 3 * 256.f;
    int ox = int(n.x * iScale), oy = int(n.y * iScale), oz = int(n.z * iScale);
    float ix = __int_as_float(__float_as_int(p.x) + ((p.x < 0.f) ? -ox : ox));
    float iy = __int_as_float(__float_as_int(p.y) + ((p.y < 0.f) ? -oy : oy));
    float iz = __int_as_float(__float_as_int(p.z) + ((p.z < 0.f) ? -oz : oz));
    return mk3(fabsf(p.x) < origin ? p.x + n.x * fScale : ix, fabsf(p.y) < origin ? p.y + n.y * fScale : iy, fabsf(p.z) < origin ? p.z + n.z * fScale : iz);
}

} // namespace pt
