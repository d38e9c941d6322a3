// ORACLE — test infrastructure only.  CPU restatement of RTXPT's PathTrace hot path; never linked into the product.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may build or call this.
//
// pt_math.h: vector types, HLSL intrinsics and the packing helpers the reference shaders rely on.
//   Rtxpt/Shaders/PathTracer/Utils/Packing.hlsli, Utils/Utils.hlsli, Utils/Math/MathHelpers.hlsli, Utils/Geometry.hlsli
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>

namespace orc {

typedef uint32_t uint;

static const float K_PI   = 3.14159265358979323846f;
static const float K_2PI  = 6.28318530717958647692f;
static const float K_1_PI = 0.31830988618379067153f;
static const float K_2_PI = 0.63661977236758134308f;
static const float K_PI_2 = 1.57079632679489661923f;
static const float K_PI_4 = 0.78539816339744830961f;
static const float HLF_MAX = 65504.0f;
static const float FLT_MAX_ = 3.402823466e+38f;
static const float FLT_MIN_ = 1.175494351e-38f;
static const float kMaxRayTravel = 1e15f;               // Rtxpt/Shaders/PathTracer/Config.h:86

struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct float4 { float x, y, z, w; };

inline float2 f2(float x, float y) { return {x, y}; }
inline float3 f3(float x, float y, float z) { return {x, y, z}; }
inline float3 f3(float s) { return {s, s, s}; }
inline float4 f4(float x, float y, float z, float w) { return {x, y, z, w}; }
inline float4 f4(float3 v, float w) { return {v.x, v.y, v.z, w}; }

inline float2 operator+(float2 a, float2 b) { return {a.x + b.x, a.y + b.y}; }
inline float2 operator-(float2 a, float2 b) { return {a.x - b.x, a.y - b.y}; }
inline float2 operator*(float2 a, float b)  { return {a.x * b, a.y * b}; }
inline float2 operator*(float b, float2 a)  { return {a.x * b, a.y * b}; }
inline float2 operator*(float2 a, float2 b) { return {a.x * b.x, a.y * b.y}; }
inline float2 operator/(float2 a, float b)  { return {a.x / b, a.y / b}; }
inline float  dot(float2 a, float2 b)       { return a.x * b.x + a.y * b.y; }
inline float  length(float2 a)              { return sqrtf(dot(a, a)); }

inline float3 operator+(float3 a, float3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline float3 operator-(float3 a, float3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline float3 operator-(float3 a)           { return {-a.x, -a.y, -a.z}; }
inline float3 operator*(float3 a, float3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline float3 operator*(float3 a, float b)  { return {a.x * b, a.y * b, a.z * b}; }
inline float3 operator*(float b, float3 a)  { return {a.x * b, a.y * b, a.z * b}; }
inline float3 operator/(float3 a, float b)  { return {a.x / b, a.y / b, a.z / b}; }
inline float3 operator/(float3 a, float3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
inline float3& operator+=(float3& a, float3 b) { a = a + b; return a; }
inline float3& operator*=(float3& a, float3 b) { a = a * b; return a; }
inline float3& operator*=(float3& a, float b)  { a = a * b; return a; }
inline float3& operator/=(float3& a, float b)  { a = a / b; return a; }
inline float  dot(float3 a, float3 b)  { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float3 cross(float3 a, float3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline float  length(float3 a)          { return sqrtf(dot(a, a)); }
inline float3 normalize(float3 a)       { return a / length(a); }       // HLSL normalize = v * rsqrt(dot(v,v)); ulp-level difference accepted
inline float3 abs3(float3 a)            { return {fabsf(a.x), fabsf(a.y), fabsf(a.z)}; }
inline float3 max3v(float3 a, float3 b) { return {std::max(a.x, b.x), std::max(a.y, b.y), std::max(a.z, b.z)}; }
inline float3 min3v(float3 a, float3 b) { return {std::min(a.x, b.x), std::min(a.y, b.y), std::min(a.z, b.z)}; }
inline float4 operator+(float4 a, float4 b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
inline float4 operator*(float4 a, float b)  { return {a.x * b, a.y * b, a.z * b, a.w * b}; }
inline float3 xyz(float4 a) { return {a.x, a.y, a.z}; }

inline float saturate(float x) { return std::min(std::max(0.0f, x), 1.0f); }     // NaN -> 0 like HLSL (std::max keeps its FIRST argument when the comparison fails)
inline float3 saturate(float3 v) { return {saturate(v.x), saturate(v.y), saturate(v.z)}; }
inline float clampf(float x, float a, float b) { return std::min(std::max(x, a), b); }
inline float3 clamp3(float3 v, float a, float b) { return {clampf(v.x, a, b), clampf(v.y, a, b), clampf(v.z, a, b)}; }
inline float lerp(float a, float b, float t) { return a + (b - a) * t; }
inline float3 lerp(float3 a, float3 b, float t) { return a + (b - a) * t; }
inline float sq(float x) { return x * x; }
inline float signf(float x) { return (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f); }
inline bool any_gt0(float3 v) { return v.x > 0 || v.y > 0 || v.z > 0; }

inline uint asuint(float f) { uint u; memcpy(&u, &f, 4); return u; }
inline float asfloat(uint u) { float f; memcpy(&f, &u, 4); return f; }
inline int asint(float f) { int i; memcpy(&i, &f, 4); return i; }
inline float asfloat_i(int i) { float f; memcpy(&f, &i, 4); return f; }

// ---- float16 (round-to-nearest-even, the D3D11+ f32tof16 rule; equals CUDA __float2half_rn) ------------------------
inline uint f32tof16(float value)
{
    uint x = asuint(value);
    uint sign = (x >> 16) & 0x8000u;
    uint mant = x & 0x007FFFFFu;
    int  exp  = int((x >> 23) & 0xFF);
    if (exp == 0xFF) return sign | 0x7C00u | (mant ? (0x200u | (mant >> 13)) : 0u);    // inf / nan
    int e = exp - 127 + 15;
    if (e >= 31) return sign | 0x7C00u;                                                 // overflow -> inf
    if (e <= 0)
    {
        if (e < -10) return sign;                                                       // underflow -> signed zero
        mant |= 0x00800000u;
        uint shift = uint(14 - e);
        uint h = mant >> shift;
        uint rem = mant & ((1u << shift) - 1u);
        uint half = 1u << (shift - 1);
        if (rem > half || (rem == half && (h & 1u))) h++;
        return sign | h;
    }
    uint h = (uint(e) << 10) | (mant >> 13);
    uint rem = mant & 0x1FFFu;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) h++;                             // may carry into exponent / inf: correct
    return sign | h;
}
inline float f16tof32(uint h)
{
    uint sign = (h & 0x8000u) << 16;
    uint exp = (h >> 10) & 0x1Fu;
    uint mant = h & 0x3FFu;
    if (exp == 0)
    {
        if (mant == 0) return asfloat(sign);
        float f = float(mant) * (1.0f / 16777216.0f);                                   // mant * 2^-24
        return asfloat(asuint(f) | sign);
    }
    if (exp == 31) return asfloat(sign | 0x7F800000u | (mant << 13));
    return asfloat(sign | ((exp + 112u) << 23) | (mant << 13));
}
// "lpfloat" (float16_t when RTXPT_LP_TYPES_USE_16BIT_PRECISION, the default: SampleUI.h:182) — this restatement evaluates
// lpfloat expressions in fp32 and rounds to fp16 where the reference stores into an lpfloat variable.
inline float lp(float v) { return f16tof32(f32tof16(v)); }
inline float3 lp(float3 v) { return {lp(v.x), lp(v.y), lp(v.z)}; }

// Utils/Packing.hlsli:206-245
inline uint Fp32ToFp16(float2 v) { return (f32tof16(clampf(v.y, -HLF_MAX, HLF_MAX)) << 16) | (f32tof16(clampf(v.x, -HLF_MAX, HLF_MAX)) & 0xFFFF); }
inline uint Fp32ToFp16NoClamp(float2 v) { return (f32tof16(v.y) << 16) | (f32tof16(v.x) & 0xFFFF); }
inline float2 Fp16ToFp32(uint r) { return {f16tof32(r & 0xFFFF), f16tof32(r >> 16)}; }

// Utils/Packing.hlsli:122-170
inline float Unpack_R8_SNORM(uint value) { int s = int(value << 24) >> 24; return clampf(float(s) / 127.0f, -1.0f, 1.0f); }
inline float3 Unpack_RGB8_SNORM(uint v) { return {Unpack_R8_SNORM(v), Unpack_R8_SNORM(v >> 8), Unpack_R8_SNORM(v >> 16)}; }
inline float4 Unpack_RGBA8_SNORM(uint v) { return {Unpack_R8_SNORM(v), Unpack_R8_SNORM(v >> 8), Unpack_R8_SNORM(v >> 16), Unpack_R8_SNORM(v >> 24)}; }
// Utils/Packing.hlsli:17-52
inline uint Pack_R8_UFLOAT(float r, float d = 0.5f) { return uint(floorf(r * 255.0f + d)) & 0xFFu; }
inline float Unpack_R8_UFLOAT(uint r) { return float(r & 0xFFu) / 255.0f; }
inline uint Pack_R8G8B8_UFLOAT(float3 rgb) { return Pack_R8_UFLOAT(rgb.x) | (Pack_R8_UFLOAT(rgb.y) << 8) | (Pack_R8_UFLOAT(rgb.z) << 16); }
inline float3 Unpack_R8G8B8_UFLOAT(uint rgb) { return {Unpack_R8_UFLOAT(rgb), Unpack_R8_UFLOAT(rgb >> 8), Unpack_R8_UFLOAT(rgb >> 16)}; }

// Utils/Utils.hlsli:51-61, Utils/ColorHelpers.hlsli:19-27
inline float Luminance(float3 rgb) { return dot(rgb, f3(0.2126f, 0.7152f, 0.0722f)); }
inline float Average(float3 rgb) { return (rgb.x + rgb.y + rgb.z) / 3.0f; }
inline float max3(float3 v) { return std::max(std::max(v.x, v.y), v.z); }

// Utils/Utils.hlsli:486-499
inline float FastSqrt(float x) { return asfloat_i(0x1fbd1df5 + (asint(x) >> 1)); }
inline float FastACos(float inX)
{
    const float PI = 3.141593f, HALF_PI = 1.570796f;
    float x = fabsf(inX);
    float res = -0.156583f * x + HALF_PI;
    res *= FastSqrt(1.0f - x);
    return (inX >= 0) ? res : PI - res;
}

// Utils/Math/MathHelpers.hlsli:185-227 (equal-area octahedral mapping)
inline float2 ndir_to_oct_equal_area_unorm(float3 n)
{
    float r = sqrtf(1.f - fabsf(n.z));
    float phi = atan2f(fabsf(n.y), fabsf(n.x));
    float2 p;
    p.y = r * phi * K_2_PI;
    p.x = r - p.y;
    if (n.z < 0.f) p = f2(1.f - p.y, 1.f - p.x);
    p = f2(p.x * signf(n.x), p.y * signf(n.y));
    return f2(saturate(p.x * 0.5f + 0.5f), saturate(p.y * 0.5f + 0.5f));
}
inline float3 oct_to_ndir_equal_area_unorm(float2 p)
{
    p = f2(p.x * 2.f - 1.f, p.y * 2.f - 1.f);
    float d = 1.f - (fabsf(p.x) + fabsf(p.y));
    float r = 1.f - fabsf(d);
    float phi = (r > 0.f) ? ((fabsf(p.y) - fabsf(p.x)) / r + 1.f) * K_PI_4 : 0.f;
    float f = r * sqrtf(2.f - r * r);
    float x = f * signf(p.x) * cosf(phi);
    float y = f * signf(p.y) * sinf(phi);
    float z = signf(d) * (1.f - r * r);
    return f3(x, y, z);
}

// Utils/Math/MathHelpers.hlsli:238-320
inline float2 sample_disk(float2 u)
{
    float r = sqrtf(u.x);
    float phi = K_2PI * u.y;
    return f2(r * cosf(phi), r * sinf(phi));
}
inline float2 sample_disk_concentric(float2 u)
{
    u = f2(2.f * u.x - 1.f, 2.f * u.y - 1.f);
    if (u.x == 0.f && u.y == 0.f) return u;
    float phi, r;
    if (fabsf(u.x) > fabsf(u.y)) { r = u.x; phi = (u.y / u.x) * K_PI_4; }
    else                         { r = u.y; phi = K_PI_2 - (u.x / u.y) * K_PI_4; }
    return f2(r * cosf(phi), r * sinf(phi));
}
inline float3 sample_cosine_hemisphere_concentric(float2 u, float& pdf)
{
    float2 d = sample_disk_concentric(u);
    float z = sqrtf(std::max(0.f, 1.f - dot(d, d)));
    pdf = z * K_1_PI;
    return f3(d.x, d.y, z);
}
// Utils/Math/MathHelpers.hlsli:436-448
inline float3 perp_stark(float3 u)
{
    float3 a = abs3(u);
    uint uyx = (a.x - a.y) < 0 ? 1 : 0;
    uint uzx = (a.x - a.z) < 0 ? 1 : 0;
    uint uzy = (a.y - a.z) < 0 ? 1 : 0;
    uint xm = uyx & uzx;
    uint ym = (1 ^ xm) & uzy;
    uint zm = 1 ^ (xm | ym);
    return normalize(cross(u, f3(float(xm), float(ym), float(zm))));
}
// Utils/Geometry.hlsli:33-41, 79-82
inline float3 SampleTriangleUniform(float2 rnd)
{
    float sqrtx = sqrtf(rnd.x);
    return f3(1 - sqrtx, sqrtx * (1 - rnd.y), sqrtx * rnd.y);
}
inline float pdfAtoW(float pdfA, float distance_, float cosTheta) { return pdfA * sq(distance_) / std::max(cosTheta, 2e-9f); }

// row-major float3x4 helpers (HLSL mul(M, float4(v,w)))
inline float3 mul34_point(const float* m, float3 v) { return f3(m[0] * v.x + m[1] * v.y + m[2] * v.z + m[3], m[4] * v.x + m[5] * v.y + m[6] * v.z + m[7], m[8] * v.x + m[9] * v.y + m[10] * v.z + m[11]); }
inline float3 mul34_vec(const float* m, float3 v)   { return f3(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[4] * v.x + m[5] * v.y + m[6] * v.z, m[8] * v.x + m[9] * v.y + m[10] * v.z); }
// HLSL mul(v, (float3x3)M): row vector times matrix
inline float3 mul_vec_33of34(float3 v, const float* m) { return f3(v.x * m[0] + v.y * m[4] + v.z * m[8], v.x * m[1] + v.y * m[5] + v.z * m[9], v.x * m[2] + v.y * m[6] + v.z * m[10]); }

} // namespace orc
