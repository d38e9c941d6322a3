/* rtxpt_b200_mgpu.h - C++-side multi-GPU host of the PathTrace boundary: N contexts, one per GPU, each tracing the interleaved screen tiles it owns
 * (RtxptConfig.tileRank / tileWorld), and ONE ncclAllGather per frame of the accumulated radiance tiles over NVLink, after which every GPU holds the whole frame.
 * The reference is single-GPU; this is the host a multi-GPU RTXPT would put around Sample::PathTrace (Rtxpt/Sample.cpp:2438-2559).
 *
 * Two ways to stand it up, same frame call afterwards:
 *   single process, all GPUs of the node ..... rtxpt_b200_mgpu_create          (ncclCommInitAll; calls may come from one host thread, the work runs asynchronously per device)
 *   one process per GPU (mpirun / torchrun) .. rtxpt_b200_mgpu_unique_id on rank 0, broadcast the 128 bytes by whatever the launcher offers, rtxpt_b200_mgpu_create_rank everywhere
 * Lives in its own shared library (librtxpt_b200_mgpu.so, links libnccl) so that librtxpt_b200.so itself carries no NCCL dependency; bench.py's Python host does the same exchange
 * with torch.distributed. */
#ifndef RTXPT_B200_MGPU_H_
#define RTXPT_B200_MGPU_H_
#include "rtxpt_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct rtxpt_mgpu rtxpt_mgpu;
#define RTXPT_MGPU_UNIQUE_ID_BYTES 128          /* sizeof(ncclUniqueId) */

/* `base`: maxSubSamplesPerLaunch, tileSize and flags are taken from it; deviceOrdinal / tileRank / tileWorld are set per context.  deviceOrdinals == NULL: devices 0..deviceCount-1. */
RTXPT_API int rtxpt_b200_mgpu_create(const RtxptConfig* base, uint32_t deviceCount, const int32_t* deviceOrdinals, rtxpt_mgpu** out);
RTXPT_API int rtxpt_b200_mgpu_unique_id(void* outId128);
RTXPT_API int rtxpt_b200_mgpu_create_rank(const RtxptConfig* base, int32_t deviceOrdinal, uint32_t rank, uint32_t world, const void* id128, rtxpt_mgpu** out);
RTXPT_API int rtxpt_b200_mgpu_destroy(rtxpt_mgpu* m);
RTXPT_API const char* rtxpt_b200_mgpu_last_error(void);

/* Scene and constants go to every local context (the scene and its BVH are replicated per GPU, §8e). */
RTXPT_API int rtxpt_b200_mgpu_upload_scene(rtxpt_mgpu* m, const RtxptSceneDesc* scene);
RTXPT_API int rtxpt_b200_mgpu_set_constants(rtxpt_mgpu* m, const RtxptPathTracerConstants* constants);
/* One frame: path_trace on every local context -> pack the owned tiles -> ncclAllGather (grouped over the local devices) -> unpack into every context's full-frame accumulated
 * image.  Asynchronous; exchange = 0 skips the all-gather (1024 spp reference accumulation gathers only at the end, §8e). */
RTXPT_API int rtxpt_b200_mgpu_render_frame(rtxpt_mgpu* m, uint32_t firstSubSampleIndex, uint32_t subSampleCount, int accumulate, int exchange);
/* Realtime mode (BASELINE configs[2]) on the same contexts: view / realtime constants go to every local context, then one call renders a frame the way rtxpt_b200.h's "The realtime
 * frame on several GPUs" lays out - BUILD + FILL on each device's tiles; ncclAllGather (grouped over the local devices) of the guides, of the seven NRD inputs per plane and of
 * the output colour, packed by tile ownership (rtxpt_b200_exchange_*); ReBLUR on the whole frame on every device; tone mapping when toneMapping != NULL.  neeatFeedback != 0 runs
 * rtxpt_b200_neeat_update_begin first (the constants must carry NEEATFeedback = 1); every device adapts on its own tiles.  rtxpt_b200_mgpu_last_frame_ms: trace / the rest. */
RTXPT_API int rtxpt_b200_mgpu_set_view(rtxpt_mgpu* m, const RtxptViewConstants* view);
RTXPT_API int rtxpt_b200_mgpu_set_realtime(rtxpt_mgpu* m, const RtxptRealtimeConstants* realtime);
RTXPT_API int rtxpt_b200_mgpu_render_realtime_frame(rtxpt_mgpu* m, const RtxptDenoiserConstants* denoiser, const RtxptReblurFrame* frame, const RtxptToneMappingParams* toneMapping, int neeatFeedback);
RTXPT_API int rtxpt_b200_mgpu_synchronize(rtxpt_mgpu* m);
/* Device time of the last render_frame per local device: trace, pack + all-gather + unpack (CUDA events on each device's stream). */
RTXPT_API int rtxpt_b200_mgpu_last_frame_ms(rtxpt_mgpu* m, uint32_t localIndex, float* outTraceMs, float* outExchangeMs);
RTXPT_API uint32_t rtxpt_b200_mgpu_local_count(const rtxpt_mgpu* m);
RTXPT_API rtxpt_ctx* rtxpt_b200_mgpu_context(rtxpt_mgpu* m, uint32_t localIndex);      /* for readback / stats through the single-GPU API */

#ifdef __cplusplus
}
#endif
#endif
