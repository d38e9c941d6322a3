// reblur_kernels.cu — NRD's REBLUR_DIFFUSE_SPECULAR chain for one stable plane, one kernel per pass of the dispatch graph
// (the pass bodies live in reblur_passes.cuh; see there for the per-pass citations).  One thread per pixel, 16x16 CTAs aligned to NRD's 16x16 sky tiles, taps fetched through L2
// (the working set of one plane at 1080p, ~90 MB, fits the 126 MB L2); shared-memory tiling of the 3x3 / 5x5 / 9x9 neighbourhoods and pass fusion are round-2 work, to be driven
// by ncu.  Compiled, NOT yet run on a GPU; the same source passes tests/test_reblur_port.py on the CPU.
#include "reblur_passes.cuh"
#include "kernels.h"

namespace pt { namespace rb {

#define RB_XY const int x = int(blockIdx.x * 16 + threadIdx.x), y = int(blockIdx.y * 16 + threadIdx.y); if (x >= int(p.W) || y >= int(p.H)) return

__global__ void __launch_bounds__(256) k_rb_classify_tiles(const __grid_constant__ Params p)
{
    const int sky = __syncthreads_count(beyondDenoisingRange(p, int(blockIdx.x * 16 + threadIdx.x), int(blockIdx.y * 16 + threadIdx.y)));
    if (threadIdx.x == 0 && threadIdx.y == 0) p.tiles[blockIdx.y * p.tilesW + blockIdx.x] = sky == 256 ? 1 : 0;
}
__global__ void __launch_bounds__(256) k_rb_hit_dist_reconstruction(const __grid_constant__ Params p) { RB_XY; hitDistReconstructionPixel(p, x, y); }
template <int MODE> __global__ void __launch_bounds__(256) k_rb_spatial(const __grid_constant__ Params p) { RB_XY; spatialPixel<MODE>(p, x, y); }
__global__ void __launch_bounds__(256) k_rb_temporal_accumulation(const __grid_constant__ Params p) { RB_XY; temporalAccumulationPixel(p, x, y); }
__global__ void __launch_bounds__(256) k_rb_history_fix(const __grid_constant__ Params p) { RB_XY; historyFixPixel(p, x, y); }
__global__ void __launch_bounds__(256) k_rb_temporal_stabilization(const __grid_constant__ Params p) { RB_XY; temporalStabilizationPixel(p, x, y); }

} // namespace rb

void launchReblurFrame(const rb::Params& p, cudaStream_t s)
{
    const dim3 grid((p.W + 15) / 16, (p.H + 15) / 16), block(16, 16);
    rb::k_rb_classify_tiles<<<grid, block, 0, s>>>(p);
    rb::k_rb_hit_dist_reconstruction<<<grid, block, 0, s>>>(p);
    rb::k_rb_spatial<0><<<grid, block, 0, s>>>(p);
    rb::k_rb_temporal_accumulation<<<grid, block, 0, s>>>(p);
    rb::k_rb_history_fix<<<grid, block, 0, s>>>(p);
    rb::k_rb_spatial<1><<<grid, block, 0, s>>>(p);
    rb::k_rb_spatial<2><<<grid, block, 0, s>>>(p);
    rb::k_rb_temporal_stabilization<<<grid, block, 0, s>>>(p);
}

} // namespace pt
