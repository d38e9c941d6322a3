// Block-compressed cudaArray probe: which texture descriptor settings the runtime accepts for cudaChannelFormatKindUnsignedBlockCompressed7, and what a fetch returns
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <vector>
__global__ void k(cudaTextureObject_t t, float4* out) { const int i = threadIdx.x; out[i] = tex2DLod<float4>(t, (i % 4 + 0.5f) / 8.0f, (i / 4 + 0.5f) / 8.0f, 0.0f); }
int main()
{
    // 8x8 BC7: four mode-6 blocks with constant colours (endpoints equal): block b -> r = g = b = 32 * (b + 1) / 255ish
    std::vector<unsigned char> blocks(4 * 16, 0);
    for (int b = 0; b < 4; b++)
    {   // mode 6: bit 6 set; 7-bit endpoints R0 R1 G0 G1 B0 B1 A0 A1, p-bits, 4-bit indices
        unsigned long long lo = 0x40, hi = 0; const unsigned v = 16 * (b + 1), a = 127; int pos = 7;
        auto put = [&](unsigned val, int bits) { for (int i = 0; i < bits; i++, pos++) { if ((val >> i) & 1) { if (pos < 64) lo |= 1ull << pos; else hi |= 1ull << (pos - 64); } } };
        put(v, 7); put(v, 7); put(v, 7); put(v, 7); put(v, 7); put(v, 7); put(a, 7); put(a, 7); put(1, 1); put(1, 1);
        memcpy(&blocks[b * 16], &lo, 8); memcpy(&blocks[b * 16 + 8], &hi, 8);
    }
    for (int srgb = 0; srgb < 2; srgb++) for (int mode = 1; mode < 4; mode++)      // mode bit 0: NormalizedFloat (ElementType is refused for these kinds), bit 1: td.sRGB
    {
        cudaChannelFormatDesc fmt = srgb ? cudaCreateChannelDesc<cudaChannelFormatKindUnsignedBlockCompressed7SRGB>() : cudaCreateChannelDesc<cudaChannelFormatKindUnsignedBlockCompressed7>();
        cudaMipmappedArray_t arr = nullptr; cudaError_t e = cudaMallocMipmappedArray(&arr, &fmt, make_cudaExtent(8, 8, 0), 1);
        printf("kind %s, readMode %s, td.sRGB %d: malloc %s", srgb ? "BC7SRGB" : "BC7", (mode & 1) ? "NormalizedFloat" : "ElementType", mode >> 1, cudaGetErrorString(e)); if (e) { printf("\n"); cudaGetLastError(); continue; }
        cudaArray_t lvl; cudaGetMipmappedArrayLevel(&lvl, arr, 0);
        e = cudaMemcpy2DToArray(lvl, 0, 0, blocks.data(), 32, 32, 2, cudaMemcpyHostToDevice); printf(", copy %s", cudaGetErrorString(e));
        cudaResourceDesc res{}; res.resType = cudaResourceTypeMipmappedArray; res.res.mipmap.mipmap = arr;
        cudaTextureDesc td{}; td.addressMode[0] = td.addressMode[1] = cudaAddressModeWrap; td.filterMode = cudaFilterModePoint; td.mipmapFilterMode = cudaFilterModePoint;
        td.readMode = (mode & 1) ? cudaReadModeNormalizedFloat : cudaReadModeElementType; td.sRGB = mode >> 1; td.normalizedCoords = 1; td.maxAnisotropy = 1; td.maxMipmapLevelClamp = 0;
        cudaTextureObject_t t = 0; e = cudaCreateTextureObject(&t, &res, &td, nullptr); printf(", texture %s", cudaGetErrorString(e));
        if (!e)
        {
            float4* out; cudaMalloc(&out, 16 * 16); k<<<1, 16>>>(t, out); e = cudaDeviceSynchronize(); float4 h[16]; cudaMemcpy(h, out, 256, cudaMemcpyDeviceToHost);
            printf(", fetch %s: texel(0,0) %.5f %.5f %.5f %.5f  texel(5,0) %.5f  texel(0,5) %.5f", cudaGetErrorString(e), h[0].x, h[0].y, h[0].z, h[0].w, h[1 * 0 + 3].x, h[12].x);
            cudaDestroyTextureObject(t);
        }
        printf("\n"); cudaGetLastError(); cudaFreeMipmappedArray(arr);
    }
    return 0;
}
