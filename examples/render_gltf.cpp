// render_gltf.cpp — a complete C++ host on top of the C ABI (include/rtxpt_b200.h): load a glTF (or an RTXPT .scene.json) with the library's loader, bridge its camera,
// fill the reference-mode constants, accumulate N samples on the GPU and write the RGBA32F accumulation as a PFM (and a tone-mapped PPM).
// This is the standalone equivalent of Sample::Render -> PathTrace -> AccumulationPass for a static scene (Rtxpt/Sample.cpp:2184, :2438-2559).
//   render_gltf scene.gltf out.pfm [width height samples bounces [materialsDir [sceneMaterialsDir]]]      (RTXPT .material.json overrides)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "rtxpt_b200.h"

static int fail(const char* what, rtxpt_ctx* ctx) { fprintf(stderr, "%s: %s\n", what, ctx ? rtxpt_b200_last_error() : rtxpt_b200_load_gltf_error()); return 1; }

int main(int argc, char** argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s scene.gltf out.pfm [width height samples bounces [materialsDir [sceneMaterialsDir]]]\n", argv[0]); return 2; }
    const uint32_t width = argc > 3 ? uint32_t(atoi(argv[3])) : 512, height = argc > 4 ? uint32_t(atoi(argv[4])) : 512;
    const uint32_t samples = argc > 5 ? uint32_t(atoi(argv[5])) : 64, bounces = argc > 6 ? uint32_t(atoi(argv[6])) : 6;

    rtxpt_host_scene* scene = nullptr;
    uint32_t overridden = 0;
    const size_t len = strlen(argv[1]); const bool sceneFile = len > 11 && !strcmp(argv[1] + len - 11, ".scene.json");
    if (sceneFile ? rtxpt_b200_load_scene_json(argv[1], nullptr, &scene) != RTXPT_OK      // RTXPT scene file: models, graph, lights, cameras, Materials/ overrides
                  : rtxpt_b200_load_gltf_ex(argv[1], argc > 7 ? argv[7] : nullptr, argc > 8 ? argv[8] : nullptr, &scene, &overridden) != RTXPT_OK) return fail("load", nullptr);
    if (overridden) fprintf(stderr, "%u materials taken from .material.json files\n", overridden);
    uint32_t cameraCount = 1; RtxptGltfCamera gcam = {};
    rtxpt_b200_host_scene_cameras(scene, &gcam, &cameraCount);
    if (cameraCount == 0) { fprintf(stderr, "the file has no perspective camera\n"); return 1; }

    RtxptCameraData cam; const float jitter[2] = { 0, 0 };
    rtxpt_b200_bridge_camera(width, height, gcam.position, gcam.direction, gcam.up, gcam.yfov, gcam.znear, gcam.zfar > 0 ? gcam.zfar : 1e7f, 10000.0f, 0.0f, jitter, &cam);
    RtxptPathTracerConstants consts; rtxpt_b200_default_constants(&cam, 0, &consts);
    consts.bounceCount = bounces; consts.diffuseBounceCount = bounces;

    RtxptConfig cfg = {}; cfg.deviceOrdinal = -1; cfg.maxSubSamplesPerLaunch = 4; cfg.tileWorld = 1; cfg.tileSize = 64;
    rtxpt_ctx* ctx = nullptr;
    if (rtxpt_b200_create(&cfg, &ctx) != RTXPT_OK) return fail("create (a CUDA device is required; there is no CPU fallback)", ctx);
    // EnvironmentLight of the scene file (Sample::SceneLoaded -> EnvMapBaker, Rtxpt/Sample.cpp:1364-1388, Lighting/Distant/EnvMapBaker.cpp:164-169): the HDR file - .dds (the reference ships
    // BC6H_UF16 cubes, *_cube_bc6u.dds), .exr or .hdr lat-long (Sample.cpp:110-118) - is decoded on the host, baked on the GPU into the cube + MIP chain the path tracer samples, and goes up with the scene
    RtxptSceneDesc desc = *rtxpt_b200_host_scene_desc(scene);
    std::vector<float> envMips; RtxptSceneFileInfo info = {}; rtxpt_b200_host_scene_info(scene, &info);
    if (sceneFile && info.environmentMapPath[0])
    {
        std::string dir(argv[1]); const size_t slash = dir.find_last_of('/'); dir = slash == std::string::npos ? std::string(".") : dir.substr(0, slash);
        const std::string envPath = dir + "/" + info.environmentMapPath;
        std::vector<unsigned char> bytes;
        if (FILE* ef = fopen(envPath.c_str(), "rb")) { fseek(ef, 0, SEEK_END); bytes.resize(size_t(ftell(ef))); fseek(ef, 0, SEEK_SET); if (fread(bytes.data(), 1, bytes.size(), ef) != bytes.size()) bytes.clear(); fclose(ef); }
        uint32_t ew = 0, eh = 0, faces = 0;
        if (bytes.empty() || rtxpt_b200_load_hdr_image(bytes.data(), bytes.size(), &ew, &eh, &faces, nullptr, 0) != RTXPT_OK)
            fprintf(stderr, "environment map '%s' not loaded (%s); rendering without it\n", envPath.c_str(), bytes.empty() ? "file missing" : rtxpt_b200_load_hdr_image_error());
        else
        {
            std::vector<float> src(size_t(ew) * eh * 4 * faces);
            rtxpt_b200_load_hdr_image(bytes.data(), bytes.size(), &ew, &eh, &faces, src.data(), src.size());
            RtxptEnvBakeDesc bake = {}; bake.cubeDim = 1024; bake.sourceType = faces >= 6 ? 2u : 1u; bake.sourceWidth = ew; bake.sourceHeight = eh; bake.source = src.data();
            for (int k = 0; k < 3; k++) bake.scaleColor[k] = info.environmentRadianceScale[k];
            envMips.resize(rtxpt_b200_env_bake_floats(bake.cubeDim));
            if (rtxpt_b200_bake_env_map(ctx, &bake, envMips.data(), envMips.size()) != RTXPT_OK) return fail("bake_env_map", ctx);
            desc.envCube.faceSize = bake.cubeDim; desc.envCube.mipLevels = rtxpt_b200_env_bake_mip_count(bake.cubeDim);
            const float* p = envMips.data();
            for (uint32_t m = 0; m < desc.envCube.mipLevels && m < RTXPT_MAX_MIPS; m++) { const size_t n = size_t(bake.cubeDim >> m) * (bake.cubeDim >> m) * 4; for (int f = 0; f < 6; f++) { desc.envCube.faces[f][m] = p; p += n; } }
            if (desc.envCube.mipLevels > RTXPT_MAX_MIPS) desc.envCube.mipLevels = RTXPT_MAX_MIPS;
            rtxpt_b200_default_constants(&cam, 1, &consts); consts.bounceCount = bounces; consts.diffuseBounceCount = bounces;
            fprintf(stderr, "environment map %s: %u x %u x %u faces -> %u^2 cube\n", info.environmentMapPath, ew, eh, faces, bake.cubeDim);
        }
    }
    if (rtxpt_b200_upload_scene(ctx, &desc) != RTXPT_OK) return fail("upload_scene", ctx);
    rtxpt_b200_free_host_scene(scene);                          // the library keeps device copies; host memory can go
    std::vector<float> frame(size_t(width) * height * 4);
    for (uint32_t done = 0; done < samples; done += 4)
    {   // sampleBaseIndex advances like m_sampleIndex * ActualSamplesPerPixel (Sample.cpp:1507)
        consts.sampleBaseIndex = done;
        if (rtxpt_b200_set_constants(ctx, &consts) != RTXPT_OK) return fail("set_constants", ctx);
        const uint32_t n = samples - done < 4 ? samples - done : 4;
        if (rtxpt_b200_path_trace(ctx, 0, n, 1, nullptr) != RTXPT_OK) return fail("path_trace", ctx);
    }
    if (rtxpt_b200_readback(ctx, RTXPT_BUFFER_ACCUMULATED_F32, frame.data(), frame.size() * sizeof(float)) != RTXPT_OK) return fail("readback", ctx);
    RtxptStats st; rtxpt_b200_get_stats(ctx, &st);
    fprintf(stderr, "%u x %u, %u spp, %u triangles, BVH %.2f s, last batch %.2f ms, %.1f Mrays/s\n", width, height, samples, st.bvhTriangleCount, st.bvhBuildSeconds,
            st.msTotal, st.msTotal > 0 ? double(st.scatterRays + st.shadowRays) / (st.msTotal * 1e3) : 0.0);
    rtxpt_b200_destroy(ctx);

    FILE* f = fopen(argv[2], "wb"); if (!f) { perror(argv[2]); return 1; }
    fprintf(f, "PF\n%u %u\n-1.0\n", width, height);             // PFM: RGB float, little endian, bottom row first
    for (uint32_t y = height; y-- > 0;) for (uint32_t x = 0; x < width; x++) fwrite(&frame[(size_t(y) * width + x) * 4], sizeof(float), 3, f);
    fclose(f);
    return 0;
}
