// ORACLE — test infrastructure only (see pt_math.h).
// pt_rng.h: stateless integer sample generators.  Bit-exact restatement of
//   Rtxpt/Shaders/PathTracer/Utils/NoiseAndSequences.hlsli:58-84 (Hash32, Hash32Combine, Hash32ToFloat), :130-229 (bhos_*)
//   Rtxpt/Shaders/PathTracer/Utils/StatelessSampleGenerators.hlsli:18-49 (vertex base), :60-183 (LD sequence), :187-232 (uniform)
//   Rtxpt/Shaders/PathTracer/Utils/SampleGenerators.hlsli:45-52 (sampleNext1D)
#pragma once
#include "pt_math.h"

namespace orc {

inline uint Hash32(uint x)
{
    x ^= x >> 16; x *= 0x21f0aaadu; x ^= x >> 15; x *= 0xf35a2d97u; x ^= x >> 15;
    return x;
}
inline uint Hash32Combine(uint seed, uint value) { return seed ^ (Hash32(value) + 0x9e3779b9u + (seed << 6) + (seed >> 2)); }
inline float Hash32ToFloat(uint hash) { return float(hash >> 8) / 16777216.0f; }

// Sobol direction numbers, dimensions 0..4 (NoiseAndSequences.hlsli:135-180; same table as the C++ half :463-508)
static const uint kSobolDirections[5][32] = {
    { 0x80000000, 0x40000000, 0x20000000, 0x10000000, 0x08000000, 0x04000000, 0x02000000, 0x01000000,
      0x00800000, 0x00400000, 0x00200000, 0x00100000, 0x00080000, 0x00040000, 0x00020000, 0x00010000,
      0x00008000, 0x00004000, 0x00002000, 0x00001000, 0x00000800, 0x00000400, 0x00000200, 0x00000100,
      0x00000080, 0x00000040, 0x00000020, 0x00000010, 0x00000008, 0x00000004, 0x00000002, 0x00000001 },
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
inline uint bhos_sobol(uint index, uint dimension)
{
    uint X = 0;
    for (uint bit = 0; bit < 32; bit++)
        if ((index >> bit) & 1u) X ^= kSobolDirections[dimension][bit];
    return X;
}
inline uint bhos_reverse_bits(uint x)
{
    x = ((x & 0xaaaaaaaau) >> 1) | ((x & 0x55555555u) << 1);
    x = ((x & 0xccccccccu) >> 2) | ((x & 0x33333333u) << 2);
    x = ((x & 0xf0f0f0f0u) >> 4) | ((x & 0x0f0f0f0fu) << 4);
    x = ((x & 0xff00ff00u) >> 8) | ((x & 0x00ff00ffu) << 8);
    return (x >> 16) | (x << 16);
}
inline uint bhos_owen_hash(uint x, uint seed)
{
    x ^= x * 0x3d20adeau; x += seed; x *= (seed >> 16) | 1u; x ^= x * 0x05526c56u; x ^= x * 0x53a22864u;
    return x;
}
inline uint bhos_owen_scramble(uint x, uint seed) { return bhos_reverse_bits(bhos_owen_hash(bhos_reverse_bits(x), seed)); }

enum SampleGeneratorEffectSeed : uint { SeedBase = 0, SeedScatterBSDF = 1 };

struct SampleGeneratorVertexBase
{
    uint baseHash, sampleIndex;
    static SampleGeneratorVertexBase make(uint packedPixel, uint vertexIndex, uint sampleIndex)
    {
        SampleGeneratorVertexBase r;
        r.sampleIndex = sampleIndex;
        r.baseHash = Hash32Combine(Hash32(vertexIndex + 0x035F9F29u), packedPixel);
        return r;
    }
};

// StatelessSampleGenerators.hlsli:187-232
struct UniformSampleSequenceGenerator
{
    uint currentHash;
    static UniformSampleSequenceGenerator make(const SampleGeneratorVertexBase& base, uint effectSeed, int subSampleCount = 1)
    {
        UniformSampleSequenceGenerator r;
        uint activeIndex = base.sampleIndex * uint(subSampleCount);
        r.currentHash = Hash32Combine(base.baseHash, effectSeed);
        r.currentHash = Hash32Combine(r.currentHash, activeIndex);
        return r;
    }
    uint Next() { currentHash = Hash32(currentHash); return currentHash; }
    float Next1D() { return Hash32ToFloat(Next()); }
};

// StatelessSampleGenerators.hlsli:60-183 (the default `SampleGenerator`, used non-LD by computeCameraRay)
struct SampleSequenceGenerator
{
    static const uint cLDDisabled = 0xFFFFFFFEu, cLDDisabled_RanOut = 0xFFFFFFFFu;
    uint startingHash, currentHash, sampleIndex, dimension, activeIndex;
    static SampleSequenceGenerator make(const SampleGeneratorVertexBase& base, uint effectSeed = SeedBase, bool lowDiscrepancy = false, int subSampleCount = 1)
    {
        SampleSequenceGenerator r;
        r.sampleIndex = base.sampleIndex;
        r.activeIndex = r.sampleIndex * uint(subSampleCount);
        r.currentHash = Hash32Combine(base.baseHash, effectSeed);
        r.startingHash = r.currentHash;
        if (lowDiscrepancy) r.dimension = 0;
        else { r.currentHash = Hash32Combine(r.currentHash, r.activeIndex); r.dimension = cLDDisabled; }
        return r;
    }
    uint Next()
    {
        const uint maxSupportedDimensionIndex = 5;
        if (dimension >= cLDDisabled) { currentHash = Hash32(currentHash); return currentHash; }
        uint shuffle_seed = Hash32Combine(currentHash, 0);
        uint dim_seed = Hash32Combine(currentHash, 1 + dimension);
        uint shuffled_index = bhos_owen_scramble(activeIndex, shuffle_seed);
        uint dim_sample = (dimension == 0) ? bhos_reverse_bits(shuffled_index) : bhos_sobol(shuffled_index, dimension);
        dim_sample = bhos_owen_scramble(dim_sample, dim_seed);
        dimension++;
        if (dimension >= maxSupportedDimensionIndex) { currentHash = Hash32Combine(currentHash, activeIndex); dimension = cLDDisabled_RanOut; }
        return dim_sample;
    }
    float Next1D() { return Hash32ToFloat(Next()); }
};

// SampleSequenceGenerator::Generate (StatelessSampleGenerators.hlsli:150-181): up to 4 LD values
inline void GenerateLD(uint count, const SampleGeneratorVertexBase& base, uint effectSeed, float out[4])
{
    uint activeIndex = base.sampleIndex;
    uint currentHash = Hash32Combine(base.baseHash, effectSeed);
    for (uint dim = 0; dim < count && dim < 4; dim++)
    {
        uint shuffle_seed = Hash32Combine(currentHash, 0);
        uint dim_seed = Hash32Combine(currentHash, 1 + dim);
        uint shuffled_index = bhos_owen_scramble(activeIndex, shuffle_seed);
        uint s = (dim == 0) ? bhos_reverse_bits(shuffled_index) : bhos_sobol(shuffled_index, dim);
        out[dim] = Hash32ToFloat(bhos_owen_scramble(s, dim_seed));
    }
}
// UniformSampleSequenceGenerator::Generate (StatelessSampleGenerators.hlsli:213-230)
inline void GenerateUniform(uint count, const SampleGeneratorVertexBase& base, uint effectSeed, float out[4])
{
    uint h = Hash32Combine(base.baseHash, effectSeed);
    h = Hash32Combine(h, base.sampleIndex);
    for (uint i = 0; i < count && i < 4; i++) { h = Hash32(h); out[i] = Hash32ToFloat(h); }
}

} // namespace orc
