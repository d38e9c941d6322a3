// TMA 2-D tile load probe: which way of handing the tensor map to the kernel works on this driver (top-level __grid_constant__ parameter, struct member, global memory)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <vector>
typedef unsigned int uint;
struct Maps { CUtensorMap a, b; uint flag; };
struct __align__(128) Tile { uint v[400]; unsigned long long mbar; };
__device__ __forceinline__ void load(void* dst, const CUtensorMap* map, int x, int y, unsigned long long* mbar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"((uint)__cvta_generic_to_shared(dst)), "l"(map), "r"(x), "r"(y), "r"((uint)__cvta_generic_to_shared(mbar)) : "memory");
}
__device__ void body(const CUtensorMap* map, uint* out)
{
    __shared__ Tile t;
    const uint mb = (uint)__cvta_generic_to_shared(&t.mbar);
    if (threadIdx.x == 0)
    {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(mb));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mb), "r"(1600u) : "memory");
        load(t.v, map, 30, 20, &t.mbar);
    }
    __syncthreads();
    uint done = 0;
    while (!done) asm volatile("{ .reg .pred q; mbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0; selp.u32 %0, 1, 0, q; }" : "=r"(done) : "r"(mb) : "memory");
    for (int i = threadIdx.x; i < 400; i += blockDim.x) out[i] = t.v[i];
}
__global__ void k_top(const __grid_constant__ CUtensorMap map, uint* out) { body(&map, out); }
__global__ void k_struct(const __grid_constant__ Maps m, uint* out) { body(&m.b, out); }
__global__ void k_global(const CUtensorMap* map, uint* out) { body(map, out); }
int main()
{
    const uint W = 96, H = 80; std::vector<uint> img(W * H); for (uint i = 0; i < W * H; i++) img[i] = i * 2654435761u;
    uint* d; cudaMalloc(&d, W * H * 4); cudaMemcpy(d, img.data(), W * H * 4, cudaMemcpyHostToDevice);
    uint* out; cudaMalloc(&out, 1600);
    typedef CUresult (*Encode)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* f = nullptr; cudaDriverEntryPointQueryResult q; cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
    printf("entry point: %s, query %d, f %p\n", cudaGetErrorString(e), int(q), f);
    Maps m; memset(&m, 0, sizeof(m));
    const cuuint64_t dims[2] = { W, H }, strides[1] = { W * 4 }; const cuuint32_t box[2] = { 20, 20 }, es[2] = { 1, 1 };
    for (CUtensorMap* t : { &m.a, &m.b })
    { CUresult r = ((Encode)f)(t, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); printf("encode: %d\n", int(r)); }
    CUtensorMap* dm; cudaMalloc(&dm, sizeof(CUtensorMap)); cudaMemcpy(dm, &m.a, sizeof(CUtensorMap), cudaMemcpyHostToDevice);
    auto check = [&](const char* name) {
        cudaError_t e = cudaDeviceSynchronize(); std::vector<uint> h(400); cudaMemcpy(h.data(), out, 1600, cudaMemcpyDeviceToHost); int bad = 0;
        for (int j = 0; j < 20; j++) for (int i = 0; i < 20; i++) if (h[j * 20 + i] != img[(20 + j) * W + 30 + i]) bad++;
        printf("%-10s: %s, mismatches %d\n", name, cudaGetErrorString(e), bad); cudaMemset(out, 0, 1600); };
    int which = 0; if (const char* w = getenv("WHICH")) which = atoi(w);
    if (which == 0 || which == 1) { k_top<<<1, 128>>>(m.a, out); check("top-level"); }
    if (which == 0 || which == 2) { k_struct<<<1, 128>>>(m, out); check("struct"); }
    if (which == 0 || which == 3) { k_global<<<1, 128>>>(dm, out); check("global"); }
    return 0;
}
