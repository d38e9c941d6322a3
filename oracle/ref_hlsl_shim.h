// ORACLE/_ref — TEST INFRASTRUCTURE.  A minimal C++ stand-in for the HLSL language features the reference's material headers use, so that the UNMODIFIED reference sources
//   Rtxpt/Shaders/PathTracer/Rendering/Materials/{Fresnel,Microfacet,BxDF,StandardBSDF,IBSDF,LobeType,BxDFConfig}.hlsli, Utils/Math/{MathHelpers,MathConstants}.hlsli,
//   Utils/ColorHelpers.hlsli, Scene/Material/MaterialData.hlsli
// compile with g++ from where they lie under /root/reference (oracle/ref_hlsl_tu.sh streams them to the compiler; nothing of the reference is copied into this repository).
// This file holds NO reference code: vector types with the swizzles those headers use, HLSL intrinsics by their documented meaning, float16_t as "float rounded to binary16
// after every operation" (what a native 16-bit type does), f32tof16 / f16tof32 as IEEE round-to-nearest-even conversions.  Scalar arithmetic is IEEE binary32 without
// contraction (-ffp-contract=off), the same contract as oracle/*.h, so that agreement between the oracle's restatement and this build is agreement of the formulas.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <type_traits>

typedef uint32_t uint;

// ---- binary16 -----------------------------------------------------------------------------------------------------------------------------------------------
inline uint f32tof16(float f)
{
    uint32_t x; std::memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u; x &= 0x7FFFFFFFu;
    if (x >= 0x7F800000u) return sign | 0x7C00u | ((x > 0x7F800000u) ? 0x200u : 0u);
    if (x >= 0x477FF000u) return sign | 0x7C00u;                                     // rounds to infinity
    if (x < 0x33000001u) return sign;                                                // rounds to zero
    if (x < 0x38800000u)
    {   // subnormal half
        const int shift = 126 - int(x >> 23);                                        // 14..24
        const uint32_t mant = (x & 0x7FFFFFu) | 0x800000u;
        uint32_t h = mant >> shift; const uint32_t rem = mant & ((1u << shift) - 1u), halfway = 1u << (shift - 1);
        if (rem > halfway || (rem == halfway && (h & 1u))) h++;
        return sign | h;
    }
    uint32_t h = ((x - 0x38000000u) >> 13); const uint32_t rem = x & 0x1FFFu;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) h++;
    return sign | h;
}
inline float f16tof32(uint h)
{
    const uint32_t sign = (h & 0x8000u) << 16, e = (h >> 10) & 0x1Fu, m = h & 0x3FFu; uint32_t x;
    if (e == 0) { if (m == 0) x = sign; else { int s = 0; uint32_t mm = m; while (!(mm & 0x400u)) { mm <<= 1; s++; } x = sign | ((113u - s) << 23) | ((mm & 0x3FFu) << 13); } }
    else if (e == 31) x = sign | 0x7F800000u | (m << 13);
    else x = sign | ((e + 112u) << 23) | (m << 13);
    float f; std::memcpy(&f, &x, 4); return f;
}
struct float16_t
{
    float v;                                                                         // always exactly representable in binary16
    float16_t() : v(0) {}
    float16_t(float f) : v(f16tof32(f32tof16(f))) {}
    float16_t(double f) : v(f16tof32(f32tof16(float(f)))) {}
    float16_t(int i) : v(f16tof32(f32tof16(float(i)))) {}
    float16_t(uint i) : v(f16tof32(f32tof16(float(i)))) {}
    operator float() const { return v; }
};
inline float16_t operator+(float16_t a, float16_t b) { return float16_t(a.v + b.v); }
inline float16_t operator-(float16_t a, float16_t b) { return float16_t(a.v - b.v); }
inline float16_t operator*(float16_t a, float16_t b) { return float16_t(a.v * b.v); }
inline float16_t operator/(float16_t a, float16_t b) { return float16_t(a.v / b.v); }
inline float16_t operator-(float16_t a) { return float16_t(-a.v); }
// float16_t with a wider operand: HLSL promotes to binary32
#define RTXPT_SHIM_HALF_MIXED(OP, R) \
    inline R operator OP(float16_t a, float b) { return a.v OP b; } inline R operator OP(float a, float16_t b) { return a OP b.v; } \
    inline R operator OP(float16_t a, int b) { return a.v OP float(b); } inline R operator OP(int a, float16_t b) { return float(a) OP b.v; } \
    inline R operator OP(float16_t a, double b) { return a.v OP float(b); } inline R operator OP(double a, float16_t b) { return float(a) OP b.v; }
RTXPT_SHIM_HALF_MIXED(+, float) RTXPT_SHIM_HALF_MIXED(-, float) RTXPT_SHIM_HALF_MIXED(*, float) RTXPT_SHIM_HALF_MIXED(/, float)
RTXPT_SHIM_HALF_MIXED(<, bool) RTXPT_SHIM_HALF_MIXED(>, bool) RTXPT_SHIM_HALF_MIXED(<=, bool) RTXPT_SHIM_HALF_MIXED(>=, bool) RTXPT_SHIM_HALF_MIXED(==, bool) RTXPT_SHIM_HALF_MIXED(!=, bool)
inline bool operator<(float16_t a, float16_t b) { return a.v < b.v; } inline bool operator>(float16_t a, float16_t b) { return a.v > b.v; } inline bool operator==(float16_t a, float16_t b) { return a.v == b.v; }
inline bool operator<=(float16_t a, float16_t b) { return a.v <= b.v; } inline bool operator>=(float16_t a, float16_t b) { return a.v >= b.v; } inline bool operator!=(float16_t a, float16_t b) { return a.v != b.v; }

// ---- vectors -----------------------------------------------------------------------------------------------------------------------------------------------------
template <typename T> struct vec2; template <typename T> struct vec3; template <typename T> struct vec4;
// a swizzle is a view of the parent's storage: readable as a vector, and (for the non-repeating ones) assignable
template <typename T, int A, int B> struct swz2 { T d[4]; operator vec2<T>() const; swz2& operator=(const vec2<T>& v); swz2& operator+=(const vec2<T>& v) { d[A] = d[A] + v.x; d[B] = d[B] + v.y; return *this; } };
template <typename T, int A, int B, int C> struct swz3 { T d[4]; operator vec3<T>() const; swz3& operator=(const vec3<T>& v); swz3& operator+=(const vec3<T>& v) { d[A] = d[A] + v.x; d[B] = d[B] + v.y; d[C] = d[C] + v.z; return *this; } };

template <typename T> struct vec2
{
    union { struct { T x, y; }; struct { T r, g; }; swz2<T, 0, 1> xy; swz2<T, 1, 0> yx; swz2<T, 0, 0> xx; };
    vec2() : x(T(0)), y(T(0)) {}
    vec2(T s) : x(s), y(s) {}
    template <typename U, typename std::enable_if<std::is_arithmetic<U>::value && !std::is_same<U, T>::value, int>::type = 0> vec2(U s) : x(T(s)), y(T(s)) {}
    vec2(T a, T b) : x(a), y(b) {}
    template <typename U, typename std::enable_if<!(std::is_integral<U>::value && std::is_integral<T>::value), int>::type = 0> explicit vec2(const vec2<U>& o) : x(T(o.x)), y(T(o.y)) {}
    template <typename U, typename std::enable_if<std::is_integral<U>::value && std::is_integral<T>::value && !std::is_same<U, T>::value, int>::type = 0> vec2(const vec2<U>& o) : x(T(o.x)), y(T(o.y)) {}      // int2 <-> uint2: implicit in HLSL
    template <typename U, int A, int B> explicit vec2(const swz2<U, A, B>& s) : x(T(s.d[A])), y(T(s.d[B])) {}
    vec2(const vec2& o) : x(o.x), y(o.y) {}
    vec2& operator=(const vec2& o) { x = o.x; y = o.y; return *this; }
    T& operator[](int i) { return i == 0 ? x : y; } const T& operator[](int i) const { return i == 0 ? x : y; }
};
template <typename T> struct vec3
{
    union { struct { T x, y, z; }; struct { T r, g, b; }; swz2<T, 0, 1> xy; swz2<T, 1, 0> yx; swz2<T, 1, 2> yz; swz3<T, 0, 1, 2> xyz; swz3<T, 0, 1, 2> rgb; };
    vec3() : x(T(0)), y(T(0)), z(T(0)) {}
    vec3(T s) : x(s), y(s), z(s) {}
    template <typename U, typename std::enable_if<std::is_arithmetic<U>::value && !std::is_same<U, T>::value, int>::type = 0> vec3(U s) : x(T(s)), y(T(s)), z(T(s)) {}      // float3 v = 0;
    vec3(T a, T b, T c) : x(a), y(b), z(c) {}
    vec3(const vec2<T>& a, T c) : x(a.x), y(a.y), z(c) {}
    vec3(T a, const vec2<T>& b) : x(a), y(b.x), z(b.y) {}
    template <typename U> vec3(const vec3<U>& o) : x(T(float(o.x))), y(T(float(o.y))), z(T(float(o.z))) {}     // float3 <-> float16_t3 convert implicitly in HLSL
    vec3(const vec3& o) : x(o.x), y(o.y), z(o.z) {}
    vec3& operator=(const vec3& o) { x = o.x; y = o.y; z = o.z; return *this; }
    T& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); } const T& operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
template <typename T> struct vec4
{
    union { struct { T x, y, z, w; }; struct { T r, g, b, a; }; swz2<T, 0, 1> xy; swz2<T, 2, 3> zw; swz3<T, 0, 1, 2> xyz; swz3<T, 0, 1, 2> rgb; swz3<T, 0, 1, 3> xyw; swz3<T, 0, 2, 3> xzw; };
    vec4() : x(T(0)), y(T(0)), z(T(0)), w(T(0)) {}
    vec4(T s) : x(s), y(s), z(s), w(s) {}
    vec4(T a, T b, T c, T d) : x(a), y(b), z(c), w(d) {}
    vec4(const vec3<T>& a, T d) : x(a.x), y(a.y), z(a.z), w(d) {}
    vec4(const vec2<T>& a, const vec2<T>& b) : x(a.x), y(a.y), z(b.x), w(b.y) {}
    vec4(const vec2<T>& a, T c, T d) : x(a.x), y(a.y), z(c), w(d) {}
    vec4(const vec4& o) : x(o.x), y(o.y), z(o.z), w(o.w) {}
    vec4& operator=(const vec4& o) { x = o.x; y = o.y; z = o.z; w = o.w; return *this; }
    const vec4& xyzw_() const { return *this; }
    T& operator[](int i) { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); } const T& operator[](int i) const { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
};
template <typename T, int A, int B> swz2<T, A, B>::operator vec2<T>() const { return vec2<T>(d[A], d[B]); }
template <typename T, int A, int B> swz2<T, A, B>& swz2<T, A, B>::operator=(const vec2<T>& v) { d[A] = v.x; d[B] = v.y; return *this; }
template <typename T, int A, int B, int C> swz3<T, A, B, C>::operator vec3<T>() const { return vec3<T>(d[A], d[B], d[C]); }
template <typename T, int A, int B, int C> swz3<T, A, B, C>& swz3<T, A, B, C>::operator=(const vec3<T>& v) { d[A] = v.x; d[B] = v.y; d[C] = v.z; return *this; }

typedef vec2<float> float2; typedef vec3<float> float3; typedef vec4<float> float4;
typedef vec2<float16_t> float16_t2; typedef vec3<float16_t> float16_t3; typedef vec4<float16_t> float16_t4;
typedef vec2<float16_t> half2; typedef vec3<float16_t> half3; typedef float16_t half;
typedef vec2<uint> uint2; typedef vec3<uint> uint3; typedef vec4<uint> uint4;
typedef vec2<int> int2; typedef vec3<int> int3;
typedef vec2<bool> bool2; typedef vec3<bool> bool3; typedef vec4<bool> bool4;
typedef uint16_t uint16_t1; typedef vec2<uint> uint16_t2;      /* lpuint2 carries pixel coordinates only: kept 32-bit wide here */ typedef vec3<uint16_t> uint16_t3; typedef vec4<uint16_t> uint16_t4;

// component-wise operators, non-template per type so that swizzle views convert implicitly
#define RTXPT_SHIM_VEC_OPS(V2, V3, V4, S) \
    inline V2 operator+(V2 a, V2 b) { return V2(a.x + b.x, a.y + b.y); } inline V2 operator-(V2 a, V2 b) { return V2(a.x - b.x, a.y - b.y); } \
    inline V2 operator*(V2 a, V2 b) { return V2(a.x * b.x, a.y * b.y); } inline V2 operator/(V2 a, V2 b) { return V2(a.x / b.x, a.y / b.y); } \
    inline V2 operator*(V2 a, S s) { return V2(a.x * s, a.y * s); } inline V2 operator*(S s, V2 a) { return V2(s * a.x, s * a.y); } inline V2 operator/(V2 a, S s) { return V2(a.x / s, a.y / s); } \
    inline V2 operator+(V2 a, S s) { return V2(a.x + s, a.y + s); } inline V2 operator-(V2 a, S s) { return V2(a.x - s, a.y - s); } inline V2 operator-(S s, V2 a) { return V2(s - a.x, s - a.y); } inline V2 operator+(S s, V2 a) { return V2(s + a.x, s + a.y); } \
    inline V2 operator-(V2 a) { return V2(-a.x, -a.y); } \
    inline V3 operator+(V3 a, V3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); } inline V3 operator-(V3 a, V3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); } \
    inline V3 operator*(V3 a, V3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); } inline V3 operator/(V3 a, V3 b) { return V3(a.x / b.x, a.y / b.y, a.z / b.z); } \
    inline V3 operator*(V3 a, S s) { return V3(a.x * s, a.y * s, a.z * s); } inline V3 operator*(S s, V3 a) { return V3(s * a.x, s * a.y, s * a.z); } inline V3 operator/(V3 a, S s) { return V3(a.x / s, a.y / s, a.z / s); } \
    inline V3 operator/(S s, V3 a) { return V3(s / a.x, s / a.y, s / a.z); } \
    inline V3 operator+(V3 a, S s) { return V3(a.x + s, a.y + s, a.z + s); } inline V3 operator+(S s, V3 a) { return V3(s + a.x, s + a.y, s + a.z); } inline V3 operator-(V3 a, S s) { return V3(a.x - s, a.y - s, a.z - s); } inline V3 operator-(S s, V3 a) { return V3(s - a.x, s - a.y, s - a.z); } \
    inline V3 operator-(V3 a) { return V3(-a.x, -a.y, -a.z); } \
    inline V3& operator+=(V3& a, V3 b) { a = a + b; return a; } inline V3& operator-=(V3& a, V3 b) { a = a - b; return a; } inline V3& operator*=(V3& a, V3 b) { a = a * b; return a; } inline V3& operator/=(V3& a, V3 b) { a = a / b; return a; } \
    inline V3& operator*=(V3& a, S s) { a = a * s; return a; } inline V3& operator/=(V3& a, S s) { a = a / s; return a; } \
    inline V2& operator+=(V2& a, V2 b) { a = a + b; return a; } inline V2& operator*=(V2& a, V2 b) { a = a * b; return a; } inline V2& operator*=(V2& a, S s) { a = a * s; return a; } inline V2& operator/=(V2& a, S s) { a = a / s; return a; } \
    inline V4 operator+(V4 a, V4 b) { return V4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); } inline V4 operator-(V4 a, V4 b) { return V4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); } \
    inline V4 operator*(V4 a, V4 b) { return V4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); } inline V4 operator*(V4 a, S s) { return V4(a.x * s, a.y * s, a.z * s, a.w * s); } inline V4 operator*(S s, V4 a) { return V4(s * a.x, s * a.y, s * a.z, s * a.w); } \
    inline V4 operator/(V4 a, S s) { return V4(a.x / s, a.y / s, a.z / s, a.w / s); }
RTXPT_SHIM_VEC_OPS(float2, float3, float4, float)
RTXPT_SHIM_VEC_OPS(float16_t2, float16_t3, float16_t4, float16_t)
// mixed precision: float16_t3 op float -> float3 (HLSL promotes the narrower operand)
inline float3 operator*(float16_t3 a, float s) { return float3(a) * s; } inline float3 operator*(float s, float16_t3 a) { return s * float3(a); }
inline float3 operator*(float16_t3 a, float3 b) { return float3(a) * b; } inline float3 operator*(float3 a, float16_t3 b) { return a * float3(b); }
inline float3 operator+(float16_t3 a, float3 b) { return float3(a) + b; } inline float3 operator+(float3 a, float16_t3 b) { return a + float3(b); }
inline bool3 operator>(float3 a, float s) { return bool3(a.x > s, a.y > s, a.z > s); } inline bool3 operator>(float16_t3 a, float s) { return bool3(float(a.x) > s, float(a.y) > s, float(a.z) > s); }
inline bool3 operator<(float3 a, float s) { return bool3(a.x < s, a.y < s, a.z < s); }
inline bool3 operator>(float3 a, float3 b) { return bool3(a.x > b.x, a.y > b.y, a.z > b.z); }
inline bool2 operator==(vec2<uint> a, vec2<uint> b) { return bool2(a.x == b.x, a.y == b.y); }
inline bool2 operator>=(float2 a, float s) { return bool2(a.x >= s, a.y >= s); }
inline bool3 operator==(float3 a, float s) { return bool3(a.x == s, a.y == s, a.z == s); } inline bool4 operator>(float4 a, int s) { return bool4(a.x > float(s), a.y > float(s), a.z > float(s), a.w > float(s)); } inline bool any(bool4 b) { return b.x || b.y || b.z || b.w; }
inline bool any(bool3 b) { return b.x || b.y || b.z; } inline bool all(bool3 b) { return b.x && b.y && b.z; } inline bool any(bool2 b) { return b.x || b.y; } inline bool all(bool2 b) { return b.x && b.y; }
inline bool any(float3 v) { return v.x != 0 || v.y != 0 || v.z != 0; }

// ---- matrices (row-major; mul(M, v) = M v) ------------------------------------------------------------------------------------------------------------------
struct float2x2 { union { float m[2][2]; struct { float _m00, _m01, _m10, _m11; }; }; float2x2() {} float2x2(float a, float b, float c, float d) { m[0][0] = a; m[0][1] = b; m[1][0] = c; m[1][1] = d; } float* operator[](int r) { return m[r]; } const float* operator[](int r) const { return m[r]; } };
struct float3x3 { union { float m[3][3]; struct { float _m00, _m01, _m02, _m10, _m11, _m12, _m20, _m21, _m22; }; }; float3x3() {} float3x3(float a, float b, float c, float d, float e, float f, float g, float h, float i) { m[0][0] = a; m[0][1] = b; m[0][2] = c; m[1][0] = d; m[1][1] = e; m[1][2] = f; m[2][0] = g; m[2][1] = h; m[2][2] = i; }
                  float3x3(float3 r0, float3 r1, float3 r2) { m[0][0] = r0.x; m[0][1] = r0.y; m[0][2] = r0.z; m[1][0] = r1.x; m[1][1] = r1.y; m[1][2] = r1.z; m[2][0] = r2.x; m[2][1] = r2.y; m[2][2] = r2.z; }
                  float* operator[](int r) { return m[r]; } const float* operator[](int r) const { return m[r]; } };
struct float3x4 { float m[3][4]; explicit operator float3x3() const { return float3x3(m[0][0], m[0][1], m[0][2], m[1][0], m[1][1], m[1][2], m[2][0], m[2][1], m[2][2]); } };          // constant-buffer member of ToneMappingConstants (not used by the pinned operators)
struct float2x3 { union { float m[2][3]; struct { float _m00, _m01, _m02, _m10, _m11, _m12; }; }; float2x3() {} float2x3(float a, float b, float c, float d, float e, float f) { _m00 = a; _m01 = b; _m02 = c; _m10 = d; _m11 = e; _m12 = f; } float* operator[](int r) { return m[r]; } const float* operator[](int r) const { return m[r]; } };
// a half matrix: elements are binary16 values.  How DXC lowers a product of two half matrices (which adds are fused, in which order, at which width) cannot be observed here;
// the shim computes the products in binary32 and rounds each element of the result once - the convention the oracle and the product follow (pt_path.h: SplitDeltaPath)
struct float16_t3x3
{
    float m[3][3];
    float16_t3x3() {}
    explicit float16_t3x3(const float3x3& a) { for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) m[r][c] = f16tof32(f32tof16(a.m[r][c])); }
    float16_t3x3(float3 r0, float3 r1, float3 r2) { const float3 rows[3] = { r0, r1, r2 }; for (int r = 0; r < 3; r++) { m[r][0] = f16tof32(f32tof16(rows[r].x)); m[r][1] = f16tof32(f32tof16(rows[r].y)); m[r][2] = f16tof32(f32tof16(rows[r].z)); } }
    float16_t3x3(float a, float b, float c, float d, float e, float f, float g, float h, float i) { const float v[9] = { a, b, c, d, e, f, g, h, i }; for (int k = 0; k < 9; k++) m[k / 3][k % 3] = f16tof32(f32tof16(v[k])); }
    operator float3x3() const { float3x3 o; for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) o.m[r][c] = m[r][c]; return o; }
};
inline float16_t3x3 mul(const float16_t3x3& A, const float16_t3x3& B)
{
    float16_t3x3 o; for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) o.m[r][c] = f16tof32(f32tof16(A.m[r][0] * B.m[0][c] + A.m[r][1] * B.m[1][c] + A.m[r][2] * B.m[2][c]));
    return o;
}
inline float3x3 mul(const float3x3& A, const float3x3& B) { float3x3 o; for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) o.m[r][c] = A.m[r][0] * B.m[0][c] + A.m[r][1] * B.m[1][c] + A.m[r][2] * B.m[2][c]; return o; }
inline float3x3 mul(const float3x3& A, const float16_t3x3& B) { return mul(A, float3x3(B)); }      // the half matrix is promoted
inline float16_t3x3 transpose(const float16_t3x3& A) { float16_t3x3 o; for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) o.m[r][c] = A.m[c][r]; return o; }
inline float3 shimRow(const float3x3& M, int r) { return float3(M.m[r][0], M.m[r][1], M.m[r][2]); } inline void shimSetRow(float3x3& M, int r, float3 v) { M.m[r][0] = v.x; M.m[r][1] = v.y; M.m[r][2] = v.z; }
inline float2 mul(const float2x2& M, float2 v) { return float2(M.m[0][0] * v.x + M.m[0][1] * v.y, M.m[1][0] * v.x + M.m[1][1] * v.y); }
inline float3 mul(const float3x3& M, float3 v) { return float3(M.m[0][0] * v.x + M.m[0][1] * v.y + M.m[0][2] * v.z, M.m[1][0] * v.x + M.m[1][1] * v.y + M.m[1][2] * v.z, M.m[2][0] * v.x + M.m[2][1] * v.y + M.m[2][2] * v.z); }
inline float3 mul(float3 v, const float3x3& M) { return float3(v.x * M.m[0][0] + v.y * M.m[1][0] + v.z * M.m[2][0], v.x * M.m[0][1] + v.y * M.m[1][1] + v.z * M.m[2][1], v.x * M.m[0][2] + v.y * M.m[1][2] + v.z * M.m[2][2]); }
inline float determinant(const float2x2& M) { return M.m[0][0] * M.m[1][1] - M.m[0][1] * M.m[1][0]; }
inline float determinant(const float3x3& M) { return M.m[0][0] * (M.m[1][1] * M.m[2][2] - M.m[1][2] * M.m[2][1]) - M.m[0][1] * (M.m[1][0] * M.m[2][2] - M.m[1][2] * M.m[2][0]) + M.m[0][2] * (M.m[1][0] * M.m[2][1] - M.m[1][1] * M.m[2][0]); }
inline float2x2 operator*(const float2x2& M, float s) { return float2x2(M.m[0][0] * s, M.m[0][1] * s, M.m[1][0] * s, M.m[1][1] * s); }
inline float2x2 operator/(const float2x2& M, float s) { return float2x2(M.m[0][0] / s, M.m[0][1] / s, M.m[1][0] / s, M.m[1][1] / s); }
inline float3x3 operator*(const float3x3& M, float s) { float3x3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = M.m[i][j] * s; return r; }
inline float3x3 operator/(const float3x3& M, float s) { float3x3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = M.m[i][j] / s; return r; }
inline float3x3 transpose(const float3x3& M) { float3x3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = M.m[j][i]; return r; }
inline float2x2 mul(const float2x3&, const float3x3&);      // declared for the generic inverse( float2x3 ), which nothing pinned here calls
inline float2x3 mul(const float2x2&, const float2x3&);

// ---- intrinsics ----------------------------------------------------------------------------------------------------------------------------------------------------
// scalar overloads spelled out so that HLSL's mixed int / float / double literals resolve to binary32 as they do in the shader
#define RTXPT_SHIM_MINMAX(F, OP) \
    inline float F(float a, float b) { return (b != b) ? a : ((a != a) ? b : ((a OP b) ? a : b)); } /* DXIL FMin / FMax: a NaN operand loses */ inline float F(float a, int b) { return F(a, float(b)); } inline float F(int a, float b) { return F(float(a), b); } \
    inline float F(float a, double b) { return F(a, float(b)); } inline float F(double a, float b) { return F(float(a), b); } inline int F(int a, int b) { return (a OP b) ? a : b; } inline uint F(uint a, uint b) { return (a OP b) ? a : b; } \
    inline float F(float16_t a, float b) { return F(float(a), b); } inline float F(float a, float16_t b) { return F(a, float(b)); } inline float16_t F(float16_t a, float16_t b) { return (a.v OP b.v) ? a : b; }
RTXPT_SHIM_MINMAX(min, <)
RTXPT_SHIM_MINMAX(max, >)
inline float2 min(float2 a, float2 b) { return float2(min(a.x, b.x), min(a.y, b.y)); } inline float2 max(float2 a, float2 b) { return float2(max(a.x, b.x), max(a.y, b.y)); }
inline float3 min(float3 a, float3 b) { return float3(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z)); } inline float3 max(float3 a, float3 b) { return float3(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)); }
inline float3 max(float3 a, float b) { return max(a, float3(b)); } inline float3 min(float3 a, float b) { return min(a, float3(b)); }
inline float2 max(float2 a, float b) { return max(a, float2(b)); } inline float2 min(float2 a, float b) { return min(a, float2(b)); }
inline float saturate(float x) { return (x != x) ? 0.0f : (x < 0.0f ? 0.0f : (x > 1.0f ? 1.0f : x)); } /* DXIL Saturate: NaN -> 0 */ inline float saturate(double x) { return saturate(float(x)); }
inline float3 saturate(float3 v) { return float3(saturate(v.x), saturate(v.y), saturate(v.z)); } inline float2 saturate(float2 v) { return float2(saturate(v.x), saturate(v.y)); }
inline float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); } inline float clamp(float x, int lo, float hi) { return clamp(x, float(lo), hi); } inline float clamp(float x, float lo, int hi) { return clamp(x, lo, float(hi)); }
inline float clamp(float x, int lo, int hi) { return clamp(x, float(lo), float(hi)); }
inline float3 clamp(float3 v, float lo, float hi) { return float3(clamp(v.x, lo, hi), clamp(v.y, lo, hi), clamp(v.z, lo, hi)); }
inline float3 clamp(float3 v, float3 lo, float3 hi) { return float3(clamp(v.x, lo.x, hi.x), clamp(v.y, lo.y, hi.y), clamp(v.z, lo.z, hi.z)); }
inline float4 clamp(float4 v, float4 lo, float4 hi) { return float4(clamp(v.x, lo.x, hi.x), clamp(v.y, lo.y, hi.y), clamp(v.z, lo.z, hi.z), clamp(v.w, lo.w, hi.w)); }
inline float2 clamp(float2 v, float lo, float hi) { return float2(clamp(v.x, lo, hi), clamp(v.y, lo, hi)); }
inline float lerp(float a, float b, float t) { return a + (b - a) * t; }            // HLSL lerp: x + s ( y - x )
inline float smoothstep(float a, float b, float x) { const float t = saturate((x - a) / (b - a)); return t * t * (3.0f - 2.0f * t); }      // HLSL smoothstep
inline float3 lerp(float3 a, float3 b, float t) { return a + (b - a) * t; } inline float3 lerp(float3 a, float3 b, float3 t) { return a + (b - a) * t; } inline float2 lerp(float2 a, float2 b, float t) { return a + (b - a) * t; }
inline float mad(float a, float b, float c) { return a * b + c; } inline float3 mad(float3 a, float3 b, float3 c) { return a * b + c; }                   // not fused: "mad" leaves fusing to the compiler; the oracle and product use the unfused form
inline float rcp(float x) { return 1.0f / x; } inline float rsqrt(float x) { return 1.0f / std::sqrt(x); }
inline float sqrt(float x) { return std::sqrt(x); } inline float sqrt(int x) { return std::sqrt(float(x)); }
inline float3 sqrt(float3 v) { return float3(std::sqrt(v.x), std::sqrt(v.y), std::sqrt(v.z)); } inline float3 sqrt(float16_t3 v) { return sqrt(float3(v)); } inline float2 sqrt(float2 v) { return float2(std::sqrt(v.x), std::sqrt(v.y)); }
inline float abs(float x) { return std::fabs(x); } inline float3 abs(float3 v) { return float3(std::fabs(v.x), std::fabs(v.y), std::fabs(v.z)); } inline float2 abs(float2 v) { return float2(std::fabs(v.x), std::fabs(v.y)); }
inline int abs(int x) { return x < 0 ? -x : x; }
inline float pow(float x, float y) { return std::pow(x, y); } inline float pow(float x, int y) { return std::pow(x, float(y)); } inline float pow(float x, double y) { return std::pow(x, float(y)); }
inline float3 pow(float3 v, float y) { return float3(pow(v.x, y), pow(v.y, y), pow(v.z, y)); } inline float3 pow(float3 v, float3 y) { return float3(pow(v.x, y.x), pow(v.y, y.y), pow(v.z, y.z)); }
inline float exp(float x) { return std::exp(x); } inline float exp2(float x) { return std::exp2(x); } inline float log(float x) { return std::log(x); } inline float log2(float x) { return std::log2(x); }
inline float3 exp(float3 v) { return float3(std::exp(v.x), std::exp(v.y), std::exp(v.z)); } inline float3 log(float3 v) { return float3(std::log(v.x), std::log(v.y), std::log(v.z)); }
inline float sin(float x) { return std::sin(x); } inline float cos(float x) { return std::cos(x); } inline float tan(float x) { return std::tan(x); }
inline float acos(float x) { return std::acos(x); } inline float asin(float x) { return std::asin(x); } inline float atan(float x) { return std::atan(x); } inline float atan2(float y, float x) { return std::atan2(y, x); }
inline void sincos(float x, float& s, float& c) { s = std::sin(x); c = std::cos(x); }
inline float floor(float x) { return std::floor(x); } inline float ceil(float x) { return std::ceil(x); } inline float frac(float x) { return x - std::floor(x); } inline float round(float x) { return std::nearbyint(x); }
inline float2 floor(float2 v) { return float2(std::floor(v.x), std::floor(v.y)); } inline float3 floor(float3 v) { return float3(std::floor(v.x), std::floor(v.y), std::floor(v.z)); }
inline float sign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); } inline float2 sign(float2 v) { return float2(sign(v.x), sign(v.y)); }
inline float step(float e, float x) { return x >= e ? 1.0f : 0.0f; }
inline float fmod(float a, float b) { return std::fmod(a, b); }
inline bool isnan(float x) { return std::isnan(x); } inline bool isinf(float x) { return std::isinf(x); } inline bool isfinite(float x) { return std::isfinite(x); }
inline float dot(float2 a, float2 b) { return a.x * b.x + a.y * b.y; }
inline float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float dot(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
inline float16_t dot(float16_t3 a, float16_t3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float3 cross(float3 a, float3 b) { return float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline float16_t3 cross(float16_t3 a, float16_t3 b) { return float16_t3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline float length(float2 v) { return std::sqrt(dot(v, v)); } inline float length(float3 v) { return std::sqrt(dot(v, v)); }
inline float3 normalize(float3 v) { return v / length(v); }      // the documented meaning, x / length( x ); a shader compiler may emit x * rsqrt( dot( x, x ) ), an ulp-level difference
inline float2 normalize(float2 v) { return v / length(v); }
inline float3 reflect(float3 i, float3 n) { return i - 2.0f * dot(n, i) * n; }
inline uint asuint(float f) { uint u; std::memcpy(&u, &f, 4); return u; } inline float asfloat(uint u) { float f; std::memcpy(&f, &u, 4); return f; } inline int asint(float f) { int i; std::memcpy(&i, &f, 4); return i; } inline float asfloat(int i) { float f; std::memcpy(&f, &i, 4); return f; }
inline uint3 asuint(float3 v) { return uint3(asuint(v.x), asuint(v.y), asuint(v.z)); }
inline uint firstbithigh(uint v) { return v ? 31u - uint(__builtin_clz(v)) : 0xFFFFFFFFu; } inline uint firstbitlow(uint v) { return v ? uint(__builtin_ctz(v)) : 0xFFFFFFFFu; }
inline uint countbits(uint v) { return uint(__builtin_popcount(v)); } inline uint reversebits(uint v) { uint r = 0; for (int i = 0; i < 32; i++) r |= ((v >> i) & 1u) << (31 - i); return r; }
// uint3 pieces of TriangleLight::Create / Store (PolymorphicLight.hlsli:478-515): masks, shifts, half conversion per component
inline uint3 operator&(uint3 a, uint m) { return uint3(a.x & m, a.y & m, a.z & m); } inline uint3 operator>>(uint3 a, int n) { return uint3(a.x >> n, a.y >> n, a.z >> n); }
inline uint3 operator<<(uint3 a, int n) { return uint3(a.x << n, a.y << n, a.z << n); } inline uint3 operator|(uint3 a, uint3 b) { return uint3(a.x | b.x, a.y | b.y, a.z | b.z); }
inline float2 f16tof32(uint2 h) { return float2(f16tof32(h.x), f16tof32(h.y)); } inline uint2 f32tof16(float2 f) { return uint2(f32tof16(f.x), f32tof16(f.y)); }
inline float4& operator+=(float4& a, float4 b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; return a; }
inline uint2& operator+=(uint2& a, uint2 b) { a.x += b.x; a.y += b.y; return a; }
inline float3 f16tof32(uint3 h) { return float3(f16tof32(h.x), f16tof32(h.y), f16tof32(h.z)); } inline uint3 f32tof16(float3 f) { return uint3(f32tof16(f.x), f32tof16(f.y), f32tof16(f.z)); }
inline float3 asfloat(uint3 v) { float3 r; std::memcpy(&r.x, &v.x, 4); std::memcpy(&r.y, &v.y, 4); std::memcpy(&r.z, &v.z, 4); return r; }
// int3 pieces of ComputeRayOrigin (PathTracerHelpers.hlsli:29-42): bit casts per component, integer add / negate, per-component select
inline int3 asint(float3 v) { return int3(asint(v.x), asint(v.y), asint(v.z)); } inline float3 asfloat(int3 v) { return float3(asfloat(v.x), asfloat(v.y), asfloat(v.z)); }
inline int3 operator+(int3 a, int3 b) { return int3(a.x + b.x, a.y + b.y, a.z + b.z); } inline int3 operator-(int3 a) { return int3(-a.x, -a.y, -a.z); }
inline int3 select(bool3 c, int3 a, int3 b) { return int3(c.x ? a.x : b.x, c.y ? a.y : b.y, c.z ? a.z : b.z); }
inline float select(bool c, float a, float b) { return c ? a : b; }
inline int2 select(bool2 c, int2 a, int2 b) { return int2(c.x ? a.x : b.x, c.y ? a.y : b.y); }
inline float2 select(bool2 c, float a, float b) { return float2(c.x ? a : b, c.y ? a : b); } inline float3 select(bool3 c, float3 a, float3 b) { return float3(c.x ? a.x : b.x, c.y ? a.y : b.y, c.z ? a.z : b.z); }

#define row_major        /* HLSL matrix layout qualifier */
inline void DebugCross(float3, float, float4) {}      // debug drawing hooks of the lighting headers: no-ops
// ---- resource views over host arrays (Lighting/LightSampler.hlsli, LightingTypes.hlsli): reads outside the bound range return 0 as a D3D buffer / texture does, writes outside
// are dropped ----------------------------------------------------------------------------------------------------------------------------------------------------------------------
inline uint InstanceIndex() { return 0u; } inline uint GeometryIndex() { return 0u; } inline uint PrimitiveIndex() { return 0u; }      // DXR intrinsics (only named by a TriangleHit::make overload the path does not call)
typedef vec4<uint> uint4_shim_;
inline vec4<uint> operator&(vec4<uint> a, uint m) { return vec4<uint>(a.x & m, a.y & m, a.z & m, a.w & m); } inline vec4<uint> operator>>(vec4<uint> a, int n) { return vec4<uint>(a.x >> n, a.y >> n, a.z >> n, a.w >> n); }
inline vec4<uint> operator<<(vec4<uint> a, int n) { return vec4<uint>(a.x << n, a.y << n, a.z << n, a.w << n); } inline vec4<uint> operator|(vec4<uint> a, vec4<uint> b) { return vec4<uint>(a.x | b.x, a.y | b.y, a.z | b.z, a.w | b.w); }
inline float4 f16tof32(vec4<uint> h) { return float4(f16tof32(h.x), f16tof32(h.y), f16tof32(h.z), f16tof32(h.w)); } inline vec4<uint> f32tof16(float4 f) { return vec4<uint>(f32tof16(f.x), f32tof16(f.y), f32tof16(f.z), f32tof16(f.w)); }
inline float4 clamp(float4 v, float lo, float hi) { return float4(clamp(v.x, lo, hi), clamp(v.y, lo, hi), clamp(v.z, lo, hi), clamp(v.w, lo, hi)); }
template <typename T> struct RWStructuredBuffer { T* p = nullptr; uint n = 0; mutable T sink{}; T& operator[](uint i) const { if (i < n) return p[i]; sink = T{}; return sink; } };
template <typename T> struct RWTexture2DArray { T* p = nullptr; uint w = 0, h = 0, d = 0; mutable T sink = T(0); T& operator[](uint3 c) const { if (c.x < w && c.y < h && c.z < d) return p[(size_t(c.z) * h + c.y) * w + c.x]; sink = T(0); return sink; } };
struct SamplerState {};
// a cube map the generator defines analytically: g_shimCubeSample( direction, lod ) (both sides of a golden evaluate the same closed form)
extern float4 (*g_shimCubeSample)(float3 dir, float lod);
template <typename T> struct TextureCube { T SampleLevel(SamplerState, float3 dir, float lod) const { return g_shimCubeSample(dir, lod); } };
// compute-shader resources and intrinsics of the baker passes (Rtxpt/Lighting/LightsBaker.hlsl).  A pass whose threads do not talk to each other is run one thread after the other:
// barriers are then no-ops and a wave is one lane wide (the generator only runs such passes - the others are compiled, not called)
template <typename T> struct RWBuffer { T* p = nullptr; uint n = 0; mutable T sink = T(0); T& operator[](uint i) const { if (i < n) return p[i]; sink = T(0); return sink; } };
struct RWByteAddressBuffer
{
    uint* p = nullptr; uint bytes = 0;
    uint Load(uint a) const { return a + 4 <= bytes ? p[a >> 2] : 0u; } void Store(uint a, uint v) const { if (a + 4 <= bytes) p[a >> 2] = v; }
    template <typename T> T Load(uint a) const { T t; std::memset(&t, 0, sizeof(T)); if (a + sizeof(T) <= bytes) std::memcpy(&t, reinterpret_cast<const char*>(p) + a, sizeof(T)); return t; }
    template <typename T> void Store(uint a, const T& t) const { if (a + sizeof(T) <= bytes) std::memcpy(reinterpret_cast<char*>(p) + a, &t, sizeof(T)); }
    void InterlockedAdd(uint a, uint v, uint& old) const { old = Load(a); Store(a, old + v); } void InterlockedAdd(uint a, uint v) const { Store(a, Load(a) + v); }
    void InterlockedCompareExchange(uint a, uint cmp, uint v, uint& old) const { old = Load(a); if (old == cmp) Store(a, v); }
};
// a pass that synchronises its thread group is run on real threads by the generator (one group at a time, `groupshared` = statics): it points g_shimGroupBarrier at a barrier for the
// group's thread count; everywhere else the pointer is null and the barrier is a no-op
#include <pthread.h>
static pthread_barrier_t* g_shimGroupBarrier = nullptr;
inline void GroupMemoryBarrierWithGroupSync() { if (g_shimGroupBarrier) pthread_barrier_wait(g_shimGroupBarrier); } inline void GroupMemoryBarrier() {} inline void DeviceMemoryBarrier() {} inline void AllMemoryBarrier() {}
template <typename T, typename U> inline void InterlockedAdd(T& dst, U v) { dst = dst + T(v); } template <typename T, typename U> inline void InterlockedAdd(T& dst, U v, T& old) { old = dst; dst = dst + T(v); }
template <typename T, typename U> inline void InterlockedMax(T& dst, U v) { if (T(v) > dst) dst = T(v); } template <typename T, typename U> inline void InterlockedMin(T& dst, U v) { if (T(v) < dst) dst = T(v); }
template <typename T, typename U> inline void InterlockedOr(T& dst, U v) { dst = dst | T(v); }
template <typename T, typename U, typename V> inline void InterlockedCompareExchange(T& dst, U cmp, V v, T& old) { old = dst; if (dst == T(cmp)) dst = T(v); }
template <typename T> inline T WaveActiveMax(T v) { return v; } template <typename T> inline T WaveActiveMin(T v) { return v; } template <typename T> inline T WaveActiveSum(T v) { return v; }
template <typename T> inline T WavePrefixSum(T) { return T(0); } template <typename T> inline T WaveReadLaneFirst(T v) { return v; } template <typename T> inline T WaveReadLaneAt(T v, uint) { return v; }
inline bool WaveIsFirstLane() { return true; } inline uint WaveGetLaneIndex() { return 0u; } inline uint WaveGetLaneCount() { return 1u; } inline uint WaveActiveCountBits(bool b) { return b ? 1u : 0u; }
inline uint WavePrefixCountBits(bool) { return 0u; } inline bool WaveActiveAnyTrue(bool b) { return b; } inline bool WaveActiveAllTrue(bool b) { return b; }
inline vec4<uint> WaveActiveBallot(bool b) { return vec4<uint>(b ? 1u : 0u, 0u, 0u, 0u); } template <typename T> inline vec4<uint> WaveMatch(T) { return vec4<uint>(1u, 0u, 0u, 0u); }
inline uint WaveMultiPrefixCountBits(bool, vec4<uint>) { return 0u; } inline vec4<uint> countbits(vec4<uint> m) { return vec4<uint>(uint(__builtin_popcount(m.x)), uint(__builtin_popcount(m.y)), uint(__builtin_popcount(m.z)), uint(__builtin_popcount(m.w))); }
struct RayDesc { float3 Origin; float TMin; float3 Direction; float TMax; };     // the D3D built-in
template <typename T> struct StructuredBuffer { const T* p = nullptr; uint n = 0; T operator[](uint i) const { return i < n ? p[i] : T{}; } };
template <typename T> struct Buffer { const T* p = nullptr; uint n = 0; T operator[](uint i) const { return i < n ? p[i] : T(0); } };
template <typename T> struct Texture2D { const T* p = nullptr; uint w = 0, h = 0; T operator[](uint2 c) const { return (c.x < w && c.y < h) ? p[size_t(c.y) * w + c.x] : T(0); } T operator[](int2 c) const { return (*this)[uint2(uint(c.x), uint(c.y))]; }
    void GetDimensions(uint& ow, uint& oh) const { ow = w; oh = h; } T Load(int3 c) const { return (c.x >= 0 && c.y >= 0 && uint(c.x) < w && uint(c.y) < h) ? p[size_t(c.y) * w + c.x] : T(0); } };
template <typename T> struct RWTexture3D { T* p = nullptr; mutable T sink = T(0); T& operator[](uint3) const { sink = T(0); return sink; } };
template <typename T> struct RWTexture2D
{
    T* p = nullptr; uint w = 0, h = 0; mutable T sink = T(0);
    T& operator[](uint2 c) const { if (c.x < w && c.y < h) return p[size_t(c.y) * w + c.x]; sink = T(0); return sink; }
    T& operator[](int2 c) const { return (*this)[uint2(uint(c.x), uint(c.y))]; }
};
inline uint2 operator+(uint2 a, uint2 b) { return uint2(a.x + b.x, a.y + b.y); } inline uint2 operator/(uint2 a, uint2 b) { return uint2(a.x / b.x, a.y / b.y); } inline uint2 operator/(uint2 a, uint b) { return uint2(a.x / b, a.y / b); }
inline uint2 operator*(uint2 a, uint b) { return uint2(a.x * b, a.y * b); } inline uint2 operator*(uint2 a, uint2 b) { return uint2(a.x * b.x, a.y * b.y); } inline uint2 operator-(uint2 a, uint2 b) { return uint2(a.x - b.x, a.y - b.y); }
inline uint min(uint a, int b) { return a < uint(b) ? a : uint(b); } inline uint max(uint a, int b) { return a > uint(b) ? a : uint(b); } inline uint min(int a, uint b) { return uint(a) < b ? uint(a) : b; }
// int2 / uint2 mix freely in HLSL (the baker's pixel arithmetic): results take the left operand's type here, comparisons are per component
inline int2 toInt2(uint2 v) { return int2(int(v.x), int(v.y)); } inline uint2 toUint2(int2 v) { return uint2(uint(v.x), uint(v.y)); }
inline int2 operator+(int2 a, int2 b) { return int2(a.x + b.x, a.y + b.y); } inline int2 operator-(int2 a, int2 b) { return int2(a.x - b.x, a.y - b.y); } inline int2 operator*(int2 a, int2 b) { return int2(a.x * b.x, a.y * b.y); }
inline int2 operator/(int2 a, int2 b) { return int2(a.x / b.x, a.y / b.y); } inline int2 operator*(int2 a, int b) { return int2(a.x * b, a.y * b); } inline int2 operator*(int a, int2 b) { return int2(a * b.x, a * b.y); }
inline int2 operator/(int2 a, int b) { return int2(a.x / b, a.y / b); } inline int2 operator+(int2 a, int b) { return int2(a.x + b, a.y + b); } inline int2 operator-(int2 a, int b) { return int2(a.x - b, a.y - b); }
inline int2 operator-(int2 a) { return int2(-a.x, -a.y); } inline int2 operator+(int2 a, uint2 b) { return a + toInt2(b); } inline int2 operator-(int2 a, uint2 b) { return a - toInt2(b); }
inline uint2 operator-(uint2 a, int2 b) { return uint2(a.x - uint(b.x), a.y - uint(b.y)); } inline uint2 operator+(uint2 a, int2 b) { return uint2(a.x + uint(b.x), a.y + uint(b.y)); }
inline uint2 operator+(uint2 a, uint b) { return uint2(a.x + b, a.y + b); } inline uint2 operator-(uint2 a, uint b) { return uint2(a.x - b, a.y - b); } inline uint2 operator%(uint2 a, uint b) { return uint2(a.x % b, a.y % b); }
inline uint2 operator>>(uint2 a, uint b) { return uint2(a.x >> b, a.y >> b); } inline uint2 operator<<(uint2 a, uint b) { return uint2(a.x << b, a.y << b); } inline uint2 operator&(uint2 a, uint b) { return uint2(a.x & b, a.y & b); }
inline bool2 operator>=(int2 a, int b) { return bool2(a.x >= b, a.y >= b); } inline bool2 operator<(int2 a, int2 b) { return bool2(a.x < b.x, a.y < b.y); } inline bool2 operator>=(int2 a, int2 b) { return bool2(a.x >= b.x, a.y >= b.y); }
inline bool2 operator<(int2 a, uint2 b) { return bool2(a.x < int(b.x), a.y < int(b.y)); } inline bool2 operator>=(uint2 a, int2 b) { return bool2(int(a.x) >= b.x, int(a.y) >= b.y); }
inline bool2 operator>=(uint2 a, uint2 b) { return bool2(a.x >= b.x, a.y >= b.y); } inline bool2 operator<(uint2 a, uint2 b) { return bool2(a.x < b.x, a.y < b.y); } inline bool2 operator<(uint2 a, int2 b) { return bool2(int(a.x) < b.x, int(a.y) < b.y); }
inline bool2 operator==(int2 a, int2 b) { return bool2(a.x == b.x, a.y == b.y); } inline bool2 operator!=(uint2 a, uint2 b) { return bool2(a.x != b.x, a.y != b.y); }
inline bool2 operator&&(bool2 a, bool2 b) { return bool2(a.x && b.x, a.y && b.y); } inline bool2 operator&(bool2 a, bool2 b) { return bool2(a.x && b.x, a.y && b.y); }
inline float max(float a, uint b) { return max(a, float(b)); } inline float min(float a, uint b) { return min(a, float(b)); }
inline int2 min(int2 a, int2 b) { return int2(min(a.x, b.x), min(a.y, b.y)); } inline int2 max(int2 a, int2 b) { return int2(max(a.x, b.x), max(a.y, b.y)); } inline int2 clamp(int2 v, int2 lo, int2 hi) { return min(max(v, lo), hi); }
inline uint clamp(uint x, int a, uint b) { const uint lo = uint(a); return x < lo ? lo : (x > b ? b : x); } inline uint clamp(uint x, uint a, uint b) { return x < a ? a : (x > b ? b : x); }
