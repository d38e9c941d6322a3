// ORACLE/_ref — compiles the C++ half of the UNMODIFIED reference header
//   /root/reference/Rtxpt/Shaders/PathTracer/Utils/NoiseAndSequences.hlsli (Hash32, Hash32Combine, Hash32ToFloat :58-84, SobolC :461-525)
// from where it lies (never copied into this repository) and prints known answers as JSON.  tests/golden/make_rng_golden.py runs
// this binary and commits the vectors that pin the oracle's and the product's integer sample generators.
#include <cstdint>
#include <cstdio>
#include <cmath>
typedef uint32_t uint;
// minimal stand-ins for the donut::math names the header's C++ half mentions (declarations only; no reference code)
struct float2 { float x, y; float2(float a = 0, float b = 0) : x(a), y(b) {} };
struct float3 { float x, y, z; float3(float a = 0, float b = 0, float c = 0) : x(a), y(b), z(c) {} };
struct uint2 { uint x, y; };
struct uint3 { uint x, y, z; };
// NoiseAndSequences.hlsli pulls in Utils.hlsli, whose C++ half needs the whole donut math library; none of it is used by the
// functions pinned here, so its include guard is pre-defined to skip it.
#define __UTILS_HLSLI__
#include "NoiseAndSequences.hlsli"

int main()
{
    printf("{\n \"hash32\": [");
    const uint32_t xs[] = { 0u, 1u, 2u, 0x035F9F29u, 0xFFFFFFFFu, 0x80000000u, 12345u, 0xDEADBEEFu, 65536u, 0x00010001u, 1920u * 1080u, 4096u };
    const int nx = int(sizeof(xs) / sizeof(xs[0]));
    for (int i = 0; i < nx; i++) printf("%s[%u, %u]", i ? ", " : "", xs[i], Hash32(xs[i]));
    printf("],\n \"hash32_combine\": [");
    for (int i = 0; i < nx; i++) for (int j = 0; j < 4; j++) printf("%s[%u, %u, %u]", (i || j) ? ", " : "", xs[i], xs[(i + j + 3) % nx], Hash32Combine(xs[i], xs[(i + j + 3) % nx]));
    printf("],\n \"hash32_to_float\": [");
    for (int i = 0; i < nx; i++) printf("%s[%u, %.9g]", i ? ", " : "", Hash32(xs[i]), (double)Hash32ToFloat(Hash32(xs[i])));
    printf("],\n \"sobol\": [");
    bool first = true;
    for (uint32_t dim = 0; dim < 5; dim++) for (uint32_t k = 0; k < 24; k++)
    {
        uint32_t index = (k < 16) ? k : Hash32(k * 7919u + dim);
        printf("%s[%u, %u, %u]", first ? "" : ", ", index, dim, SobolC(index, dim)); first = false;
    }
    printf("]\n}\n");
    return 0;
}
