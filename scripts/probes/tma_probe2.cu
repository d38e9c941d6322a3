// TMA 2-D tile load probe, parametrised by environment: BOXW BOXH IMGW IMGH DLSYM(0/1) CTAGROUP(0/1) X Y
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
typedef unsigned int uint;
struct __align__(1024) Tile { uint v[4096]; unsigned long long mbar; };
template <int CG> __global__ void k(const __grid_constant__ CUtensorMap map, uint* out, int x, int y, uint bytes, uint n)
{
    __shared__ Tile t;
    const uint mb = (uint)__cvta_generic_to_shared(&t.mbar), dst = (uint)__cvta_generic_to_shared(t.v);
    if (threadIdx.x == 0)
    {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(mb));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mb), "r"(bytes) : "memory");
        if (CG) asm volatile("cp.async.bulk.tensor.2d.cta_group::1.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" :: "r"(dst), "l"(&map), "r"(x), "r"(y), "r"(mb) : "memory");
        else    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" :: "r"(dst), "l"(&map), "r"(x), "r"(y), "r"(mb) : "memory");
    }
    __syncthreads();
    uint done = 0;
    while (!done) asm volatile("{ .reg .pred q; mbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0; selp.u32 %0, 1, 0, q; }" : "=r"(done) : "r"(mb) : "memory");
    for (uint i = threadIdx.x; i < n; i += blockDim.x) out[i] = t.v[i];
}
static int env(const char* n, int d) { const char* e = getenv(n); return e ? atoi(e) : d; }
int main()
{
    const uint W = env("IMGW", 96), H = env("IMGH", 80), bw = env("BOXW", 20), bh = env("BOXH", 20); const int x = env("X", 30), y = env("Y", 20);
    std::vector<uint> img(W * H); for (uint i = 0; i < W * H; i++) img[i] = i * 2654435761u;
    uint* d; cudaMalloc(&d, W * H * 4); cudaMemcpy(d, img.data(), W * H * 4, cudaMemcpyHostToDevice);
    uint* out; cudaMalloc(&out, 4096 * 4); cudaMemset(out, 0, 4096 * 4);
    typedef CUresult (*Encode)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* f = nullptr;
    if (env("DLSYM", 0)) { void* h = dlopen("libcuda.so.1", RTLD_NOW); f = h ? dlsym(h, "cuTensorMapEncodeTiled") : nullptr; }
    else { cudaDriverEntryPointQueryResult q; cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q); }
    alignas(64) CUtensorMap m; memset(&m, 0, sizeof(m));
    const cuuint64_t dims[2] = { W, H }, strides[1] = { W * 4 }; const cuuint32_t box[2] = { bw, bh }, es[2] = { 1, 1 };
    CUresult r = ((Encode)f)(&m, env("F32", 0) ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     (CUtensorMapL2promotion)env("L2P", 0), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (env("CTAGROUP", 0)) k<1><<<1, 128>>>(m, out, x, y, bw * bh * 4, bw * bh); else k<0><<<1, 128>>>(m, out, x, y, bw * bh * 4, bw * bh);
    cudaError_t e = cudaDeviceSynchronize(); std::vector<uint> h(bw * bh); cudaMemcpy(h.data(), out, bw * bh * 4, cudaMemcpyDeviceToHost); int bad = 0;
    for (uint j = 0; j < bh; j++) for (uint i = 0; i < bw; i++) if (h[j * bw + i] != img[(y + j) * W + x + i]) bad++;
    printf("img %ux%u box %ux%u at (%d,%d) f32=%d dlsym=%d ctagroup=%d l2p=%d: encode %d, %s, mismatches %d of %u\n", W, H, bw, bh, x, y, env("F32", 0), env("DLSYM", 0), env("CTAGROUP", 0), env("L2P", 0), int(r), cudaGetErrorString(e), bad, bw * bh);
    return 0;
}
