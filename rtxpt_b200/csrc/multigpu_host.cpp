// multigpu_host.cpp - librtxpt_b200_mgpu.so: the C++ multi-GPU host declared in include/rtxpt_b200_mgpu.h.  Tile-parallel path tracing over N contexts of librtxpt_b200.so with one
// ncclAllGather of the accumulated radiance tiles per frame (the reference is single-GPU, its frame loop is Sample::Render -> PathTrace, Rtxpt/Sample.cpp:2184, :2438).
#include "../../include/rtxpt_b200_mgpu.h"
#include <cuda_runtime.h>
#include <nccl.h>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

static_assert(sizeof(ncclUniqueId) == RTXPT_MGPU_UNIQUE_ID_BYTES, "ncclUniqueId size");

namespace {
thread_local std::string g_error;
int fail(int code, const char* fmt, ...)
{
    char buf[512]; va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    g_error = buf; return code;
}
struct Local
{
    int device = 0; rtxpt_ctx* ctx = nullptr; ncclComm_t comm = nullptr; cudaStream_t stream = nullptr;
    float4* send = nullptr; float4* recv = nullptr; cudaEvent_t ev[3] = { nullptr, nullptr, nullptr }; bool timed = false;
    uint8_t* xsend = nullptr; uint8_t* xrecv = nullptr; size_t xbytes = 0;        // realtime frame: send / receive blocks of the per-pixel image exchange, grown on demand
};
}
struct rtxpt_mgpu { std::vector<Local> local; uint32_t world = 0; uint32_t padded = 0; bool haveConstants = false; uint32_t activePlanes = RTXPT_STABLE_PLANE_COUNT; };

#define CU_(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail(RTXPT_ERR_CUDA, "%s: %s", #x, cudaGetErrorString(e_)); } while (0)
#define NC_(x) do { ncclResult_t r_ = (x); if (r_ != ncclSuccess) return fail(RTXPT_ERR_CUDA, "%s: %s", #x, ncclGetErrorString(r_)); } while (0)
#define RT_(x) do { int r_ = (x); if (r_ != RTXPT_OK) return fail(r_, "%s: %s", #x, rtxpt_b200_last_error()); } while (0)

static int createLocal(const RtxptConfig* base, Local& l, int device, uint32_t rank, uint32_t world)
{
    l.device = device;
    CU_(cudaSetDevice(device));
    RtxptConfig cfg = *base; cfg.deviceOrdinal = device; cfg.tileRank = rank; cfg.tileWorld = world; if (cfg.tileSize == 0) cfg.tileSize = 64; if (cfg.maxSubSamplesPerLaunch == 0) cfg.maxSubSamplesPerLaunch = 4;
    RT_(rtxpt_b200_create(&cfg, &l.ctx));
    CU_(cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking));
    for (cudaEvent_t& e : l.ev) CU_(cudaEventCreate(&e));
    return RTXPT_OK;
}

extern "C" {

RTXPT_API const char* rtxpt_b200_mgpu_last_error(void) { return g_error.c_str(); }

RTXPT_API int rtxpt_b200_mgpu_create(const RtxptConfig* base, uint32_t deviceCount, const int32_t* deviceOrdinals, rtxpt_mgpu** out)
{
    if (!base || !out || deviceCount == 0) return fail(RTXPT_ERR_INVALID_ARGUMENT, "null argument or no devices");
    int present = 0; if (cudaGetDeviceCount(&present) != cudaSuccess || present == 0) return fail(RTXPT_ERR_NO_DEVICE, "no CUDA device (there is no CPU fallback)");
    if (int(deviceCount) > present && !deviceOrdinals) return fail(RTXPT_ERR_INVALID_ARGUMENT, "%u devices requested, %d present", deviceCount, present);
    rtxpt_mgpu* m = new rtxpt_mgpu(); m->world = deviceCount; m->local.resize(deviceCount);
    std::vector<int> devs(deviceCount); for (uint32_t i = 0; i < deviceCount; i++) devs[i] = deviceOrdinals ? deviceOrdinals[i] : int(i);
    for (uint32_t i = 0; i < deviceCount; i++) { int rc = createLocal(base, m->local[i], devs[i], i, deviceCount); if (rc != RTXPT_OK) { rtxpt_b200_mgpu_destroy(m); return rc; } }
    std::vector<ncclComm_t> comms(deviceCount);
    { ncclResult_t r = ncclCommInitAll(comms.data(), int(deviceCount), devs.data()); if (r != ncclSuccess) { rtxpt_b200_mgpu_destroy(m); return fail(RTXPT_ERR_CUDA, "ncclCommInitAll: %s", ncclGetErrorString(r)); } }
    for (uint32_t i = 0; i < deviceCount; i++) m->local[i].comm = comms[i];
    *out = m; return RTXPT_OK;
}

RTXPT_API int rtxpt_b200_mgpu_unique_id(void* outId128)
{
    if (!outId128) return fail(RTXPT_ERR_INVALID_ARGUMENT, "null argument");
    ncclUniqueId id; NC_(ncclGetUniqueId(&id)); memcpy(outId128, &id, sizeof(id)); return RTXPT_OK;
}

RTXPT_API int rtxpt_b200_mgpu_create_rank(const RtxptConfig* base, int32_t deviceOrdinal, uint32_t rank, uint32_t world, const void* id128, rtxpt_mgpu** out)
{
    if (!base || !out || !id128 || world == 0 || rank >= world) return fail(RTXPT_ERR_INVALID_ARGUMENT, "bad rank / world / id");
    rtxpt_mgpu* m = new rtxpt_mgpu(); m->world = world; m->local.resize(1);
    int rc = createLocal(base, m->local[0], deviceOrdinal, rank, world); if (rc != RTXPT_OK) { rtxpt_b200_mgpu_destroy(m); return rc; }
    ncclUniqueId id; memcpy(&id, id128, sizeof(id));
    { ncclResult_t r = ncclCommInitRank(&m->local[0].comm, int(world), id, int(rank)); if (r != ncclSuccess) { rtxpt_b200_mgpu_destroy(m); return fail(RTXPT_ERR_CUDA, "ncclCommInitRank: %s", ncclGetErrorString(r)); } }
    *out = m; return RTXPT_OK;
}

RTXPT_API int rtxpt_b200_mgpu_destroy(rtxpt_mgpu* m)
{
    if (!m) return RTXPT_OK;
    for (Local& l : m->local)
    {
        cudaSetDevice(l.device);
        if (l.stream) cudaStreamSynchronize(l.stream);
        if (l.comm) ncclCommDestroy(l.comm);
        if (l.send) cudaFree(l.send); if (l.recv) cudaFree(l.recv); if (l.xsend) cudaFree(l.xsend); if (l.xrecv) cudaFree(l.xrecv);
        for (cudaEvent_t e : l.ev) if (e) cudaEventDestroy(e);
        if (l.ctx) rtxpt_b200_destroy(l.ctx);
        if (l.stream) cudaStreamDestroy(l.stream);
    }
    delete m; return RTXPT_OK;
}

RTXPT_API int rtxpt_b200_mgpu_upload_scene(rtxpt_mgpu* m, const RtxptSceneDesc* scene)
{
    if (!m || !scene) return fail(RTXPT_ERR_INVALID_ARGUMENT, "null argument");
    for (Local& l : m->local) RT_(rtxpt_b200_upload_scene(l.ctx, scene));          // the BVH is built per context: the build is deterministic, every GPU gets the same tree
    return RTXPT_OK;
}

RTXPT_API int rtxpt_b200_mgpu_set_constants(rtxpt_mgpu* m, const RtxptPathTracerConstants* k)
{
    if (!m || !k) return fail(RTXPT_ERR_INVALID_ARGUMENT, "null argument");
    for (Local& l : m->local)
    {
        RT_(rtxpt_b200_set_constants(l.ctx, k));
        uint32_t owned = 0, padded = 0; RT_(rtxpt_b200_tile_layout(l.ctx, &owned, &padded));
        if (padded != m->padded || !l.send)
        {   // image size changed: the exchange buffers follow the tile layout (the same padded length on every rank: it is a function of W, H, tile size and world only)
            CU_(cudaSetDevice(l.device)); CU_(cudaStreamSynchronize(l.stream));
            if (l.send) cudaFree(l.send); if (l.recv) cudaFree(l.recv); l.send = l.recv = nullptr;
            CU_(cudaMalloc(&l.send, size_t(padded) * sizeof(float4))); CU_(cudaMalloc(&l.recv, size_t(padded) * m->world * sizeof(float4)));
        }
        if (&l == &m->local.back()) m->padded = padded;
    }
    m->haveConstants = true; return RTXPT_OK;
}

RTXPT_API int rtxpt_b200_mgpu_render_frame(rtxpt_mgpu* m, uint32_t firstSubSampleIndex, uint32_t subSampleCount, int accumulate, int exchange)
{
    if (!m) return fail(RTXPT_ERR_INVALID_ARGUMENT, "null argument");
    if (!m->haveConstants) return fail(RTXPT_ERR_INVALID_ARGUMENT, "constants not set");
    for (Local& l : m->local)
    {
        CU_(cudaSetDevice(l.device)); CU_(cudaEventRecord(l.ev[0], l.stream));
        RT_(rtxpt_b200_path_trace(l.ctx, firstSubSampleIndex, subSampleCount, accumulate, l.stream));
        CU_(cudaEventRecord(l.ev[1], l.stream));
        if (exchange && m->world > 1) RT_(rtxpt_b200_pack_owned(l.ctx, l.send, l.stream));
    }
    if (exchange && m->world > 1)
    {
        NC_(ncclGroupStart());                                                          // one group: the local devices' all-gathers must be issued together from a single thread
        for (Local& l : m->local) NC_(ncclAllGather(l.send, l.recv, size_t(m->padded) * 4, ncclFloat, l.comm, l.stream));
        NC_(ncclGroupEnd());
        for (Local& l : m->local) { CU_(cudaSetDevice(l.device)); RT_(rtxpt_b200_unpack_all(l.ctx, l.recv, l.stream)); }
    }
    for (Local& l : m->local) { CU_(cudaSetDevice(l.device)); CU_(cudaEventRecord(l.ev[2], l.stream)); l.timed = true; }
    return RTXPT_OK;
}

// ---- realtime mode (BASELINE configs[2]) on the local devices: the frame recipe of rtxpt_b200.h ("The realtime frame on several GPUs") -----------------------------------------
RTXPT_API int rtxpt_b200_mgpu_set_view(rtxpt_mgpu* m, const RtxptViewConstants* view)
{
    if (!m || !view) return fail(RTXPT_ERR_INVALID_ARGUMENT, "null argument");
    for (Local& l : m->local) { CU_(cudaSetDevice(l.device)); RT_(rtxpt_b200_set_view(l.ctx, view)); }
    return RTXPT_OK;
}
RTXPT_API int rtxpt_b200_mgpu_set_realtime(rtxpt_mgpu* m, const RtxptRealtimeConstants* realtime)
{
    if (!m || !realtime) return fail(RTXPT_ERR_INVALID_ARGUMENT, "null argument");
    for (Local& l : m->local) { CU_(cudaSetDevice(l.device)); RT_(rtxpt_b200_set_realtime(l.ctx, realtime)); }
    m->activePlanes = realtime->activeStablePlaneCount;
    return RTXPT_OK;
}
static int exchangeImages(rtxpt_mgpu* m, const int* ids, uint32_t count)
{
    if (m->world <= 1) return RTXPT_OK;
    size_t bytes = 0;
    for (Local& l : m->local)
    {
        CU_(cudaSetDevice(l.device)); RT_(rtxpt_b200_exchange_bytes(l.ctx, ids, count, &bytes));
        if (l.xbytes < bytes)
        {
            CU_(cudaStreamSynchronize(l.stream)); if (l.xsend) cudaFree(l.xsend); if (l.xrecv) cudaFree(l.xrecv); l.xsend = l.xrecv = nullptr; l.xbytes = 0;
            CU_(cudaMalloc(&l.xsend, bytes)); CU_(cudaMalloc(&l.xrecv, bytes * m->world)); l.xbytes = bytes;
        }
        RT_(rtxpt_b200_exchange_pack(l.ctx, ids, count, l.xsend, l.stream));
    }
    NC_(ncclGroupStart());
    for (Local& l : m->local) NC_(ncclAllGather(l.xsend, l.xrecv, bytes, ncclUint8, l.comm, l.stream));
    NC_(ncclGroupEnd());
    for (Local& l : m->local) { CU_(cudaSetDevice(l.device)); RT_(rtxpt_b200_exchange_unpack(l.ctx, ids, count, l.xrecv, l.stream)); }
    return RTXPT_OK;
}
RTXPT_API int rtxpt_b200_mgpu_render_realtime_frame(rtxpt_mgpu* m, const RtxptDenoiserConstants* denoiser, const RtxptReblurFrame* frame, const RtxptToneMappingParams* toneMapping, int neeatFeedback)
{
    if (!m || !denoiser || !frame) return fail(RTXPT_ERR_INVALID_ARGUMENT, "null argument");
    if (!m->haveConstants) return fail(RTXPT_ERR_INVALID_ARGUMENT, "constants not set");
    static const int guides[] = { RTXPT_BUFFER_DEPTH_F32, RTXPT_BUFFER_SPECULAR_HITT_F32, RTXPT_BUFFER_STABLE_PLANE_NEIGHBOUR_GUIDES };
    static const int nrdIn[] = { RTXPT_BUFFER_DENOISER_VIEWSPACE_Z_F32, RTXPT_BUFFER_DENOISER_MOTION_VECTORS_F16, RTXPT_BUFFER_DENOISER_NORMAL_ROUGHNESS_R10G10B10A2, RTXPT_BUFFER_DENOISER_DIFF_RADIANCE_HITDIST_F16,
                                 RTXPT_BUFFER_DENOISER_SPEC_RADIANCE_HITDIST_F16, RTXPT_BUFFER_DENOISER_DISOCCLUSION_MIX_R8, RTXPT_BUFFER_COMBINED_HISTORY_CLAMP_RELAX_R8 };
    static const int color[] = { RTXPT_BUFFER_OUTPUT_COLOR_F16 };
    for (Local& l : m->local)
    {
        CU_(cudaSetDevice(l.device)); CU_(cudaEventRecord(l.ev[0], l.stream));
        if (neeatFeedback) RT_(rtxpt_b200_neeat_update_begin(l.ctx, l.stream));
        RT_(rtxpt_b200_path_trace_realtime(l.ctx, 0, l.stream));
        CU_(cudaEventRecord(l.ev[1], l.stream));
    }
    int rc = exchangeImages(m, guides, 3); if (rc != RTXPT_OK) return rc;
    for (Local& l : m->local) { CU_(cudaSetDevice(l.device)); RT_(rtxpt_b200_denoise_spec_hit_t(l.ctx, l.stream)); }
    for (int plane = int(m->activePlanes) - 1; plane >= 0; plane--)
    {
        for (Local& l : m->local) { CU_(cudaSetDevice(l.device)); RT_(rtxpt_b200_denoiser_prepare_inputs(l.ctx, uint32_t(plane), plane == int(m->activePlanes) - 1, denoiser, l.stream)); }
        rc = exchangeImages(m, nrdIn, 7); if (rc != RTXPT_OK) return rc;
        for (Local& l : m->local)
        {
            CU_(cudaSetDevice(l.device));
            RT_(rtxpt_b200_reblur_denoise(l.ctx, uint32_t(plane), frame, l.stream));                 // whole frame on every device: identical histories
            RT_(rtxpt_b200_denoiser_final_merge(l.ctx, uint32_t(plane), nullptr, nullptr, l.stream));
        }
    }
    rc = exchangeImages(m, color, 1); if (rc != RTXPT_OK) return rc;
    for (Local& l : m->local)
    {
        CU_(cudaSetDevice(l.device));
        if (toneMapping) RT_(rtxpt_b200_tone_map(l.ctx, toneMapping, RTXPT_BUFFER_OUTPUT_COLOR_F16, l.stream));
        CU_(cudaEventRecord(l.ev[2], l.stream)); l.timed = true;
    }
    return RTXPT_OK;
}

RTXPT_API int rtxpt_b200_mgpu_synchronize(rtxpt_mgpu* m)
{
    if (!m) return fail(RTXPT_ERR_INVALID_ARGUMENT, "null argument");
    for (Local& l : m->local) { CU_(cudaSetDevice(l.device)); CU_(cudaStreamSynchronize(l.stream)); }
    return RTXPT_OK;
}

RTXPT_API int rtxpt_b200_mgpu_last_frame_ms(rtxpt_mgpu* m, uint32_t i, float* outTraceMs, float* outExchangeMs)
{
    if (!m || i >= m->local.size() || !outTraceMs || !outExchangeMs) return fail(RTXPT_ERR_INVALID_ARGUMENT, "bad argument");
    Local& l = m->local[i]; if (!l.timed) return fail(RTXPT_ERR_INVALID_ARGUMENT, "no frame rendered yet");
    CU_(cudaSetDevice(l.device)); CU_(cudaEventSynchronize(l.ev[2]));
    CU_(cudaEventElapsedTime(outTraceMs, l.ev[0], l.ev[1])); CU_(cudaEventElapsedTime(outExchangeMs, l.ev[1], l.ev[2]));
    return RTXPT_OK;
}

RTXPT_API uint32_t rtxpt_b200_mgpu_local_count(const rtxpt_mgpu* m) { return m ? uint32_t(m->local.size()) : 0u; }
RTXPT_API rtxpt_ctx* rtxpt_b200_mgpu_context(rtxpt_mgpu* m, uint32_t i) { return (m && i < m->local.size()) ? m->local[i].ctx : nullptr; }

} // extern "C"
