// multigpu_gltf.cpp - C++ multi-GPU host on top of include/rtxpt_b200_mgpu.h: one context per GPU of the node, interleaved screen tiles, one ncclAllGather of the radiance tiles
// per frame.  Loads a glTF / .scene.json with the library's loader, accumulates N samples over all GPUs, prints per-GPU trace / exchange times and writes the
// frame GPU 0 ends up with (every GPU holds the whole frame after the all-gather) as a PFM.
//   multigpu_gltf scene.gltf out.pfm [gpus width height samples bounces]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "rtxpt_b200_mgpu.h"

int main(int argc, char** argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s scene.gltf out.pfm [gpus width height samples bounces]\n", argv[0]); return 2; }
    const uint32_t gpus = argc > 3 ? uint32_t(atoi(argv[3])) : 1, width = argc > 4 ? uint32_t(atoi(argv[4])) : 512, height = argc > 5 ? uint32_t(atoi(argv[5])) : 512;
    const uint32_t samples = argc > 6 ? uint32_t(atoi(argv[6])) : 64, bounces = argc > 7 ? uint32_t(atoi(argv[7])) : 6;
    rtxpt_host_scene* scene = nullptr;
    const size_t len = strlen(argv[1]); const bool sceneFile = len > 11 && !strcmp(argv[1] + len - 11, ".scene.json");
    if ((sceneFile ? rtxpt_b200_load_scene_json(argv[1], nullptr, &scene) : rtxpt_b200_load_gltf(argv[1], &scene)) != RTXPT_OK) { fprintf(stderr, "load: %s\n", rtxpt_b200_load_gltf_error()); return 1; }
    uint32_t cameraCount = 1; RtxptGltfCamera gcam = {}; rtxpt_b200_host_scene_cameras(scene, &gcam, &cameraCount);
    if (cameraCount == 0) { fprintf(stderr, "the file has no perspective camera\n"); return 1; }
    RtxptCameraData cam; const float jitter[2] = { 0, 0 };
    rtxpt_b200_bridge_camera(width, height, gcam.position, gcam.direction, gcam.up, gcam.yfov, gcam.znear, gcam.zfar > 0 ? gcam.zfar : 1e7f, 10000.0f, 0.0f, jitter, &cam);
    RtxptPathTracerConstants consts; rtxpt_b200_default_constants(&cam, 0, &consts); consts.bounceCount = bounces; consts.diffuseBounceCount = bounces;

    RtxptConfig cfg = {}; cfg.maxSubSamplesPerLaunch = 4; cfg.tileSize = 64;
    rtxpt_mgpu* m = nullptr;
    if (rtxpt_b200_mgpu_create(&cfg, gpus, nullptr, &m) != RTXPT_OK) { fprintf(stderr, "create: %s\n", rtxpt_b200_mgpu_last_error()); return 1; }
    if (rtxpt_b200_mgpu_upload_scene(m, rtxpt_b200_host_scene_desc(scene)) != RTXPT_OK) { fprintf(stderr, "upload: %s\n", rtxpt_b200_mgpu_last_error()); return 1; }
    rtxpt_b200_free_host_scene(scene);
    for (uint32_t done = 0; done < samples; done += 4)
    {
        consts.sampleBaseIndex = done; const uint32_t n = samples - done < 4 ? samples - done : 4;
        if (rtxpt_b200_mgpu_set_constants(m, &consts) != RTXPT_OK || rtxpt_b200_mgpu_render_frame(m, 0, n, 1, done + 4 >= samples /* gather once, at the end */) != RTXPT_OK)
        { fprintf(stderr, "frame: %s\n", rtxpt_b200_mgpu_last_error()); return 1; }
    }
    rtxpt_b200_mgpu_synchronize(m);
    for (uint32_t i = 0; i < rtxpt_b200_mgpu_local_count(m); i++)
    {
        float t = 0, x = 0; rtxpt_b200_mgpu_last_frame_ms(m, i, &t, &x); RtxptStats st; rtxpt_b200_get_stats(rtxpt_b200_mgpu_context(m, i), &st);
        fprintf(stderr, "gpu %u: last frame trace %.3f ms, pack + all-gather + unpack %.3f ms, %llu rays\n", i, t, x, (unsigned long long)(st.scatterRays + st.shadowRays));
    }
    std::vector<float> frame(size_t(width) * height * 4);
    if (rtxpt_b200_readback(rtxpt_b200_mgpu_context(m, 0), RTXPT_BUFFER_ACCUMULATED_F32, frame.data(), frame.size() * sizeof(float)) != RTXPT_OK) { fprintf(stderr, "readback: %s\n", rtxpt_b200_last_error()); return 1; }
    rtxpt_b200_mgpu_destroy(m);
    FILE* f = fopen(argv[2], "wb"); if (!f) { perror(argv[2]); return 1; }
    fprintf(f, "PF\n%u %u\n-1.0\n", width, height);
    for (uint32_t y = height; y-- > 0;) for (uint32_t x = 0; x < width; x++) fwrite(&frame[(size_t(y) * width + x) * 4], sizeof(float), 3, f);
    fclose(f);
    return 0;
}
