// realtime.cuh — realtime mode's path-space decomposition into "stable planes" as device functions of the shade kernels (included by shade.cuh
// after the Surface / path helpers it builds on).  What it restates:
//   Rtxpt/Shaders/PathTracer/StablePlanes.hlsli:82-274 (StablePlanesContext: header, plane records, stable radiance), :277-315 (branch IDs)
//   Rtxpt/Shaders/PathTracer/PathTracerStablePlanes.hlsli:25-99 (SplitDeltaPath), :102-326 (StablePlanesHandleHit), :329-380 (StablePlanesOnScatter),
//   :382-412 (StablePlanesHandleMiss); Rendering/Materials/BxDF.hlsli:972-1053 (evalDeltaLobes), StandardBSDF.hlsli:93-121 (estimateSpecDiffBSDF),
//   Microfacet.hlsli:282-355 (approxSpecularIntegralGGX); Utils/Utils.hlsli:118-189 (octahedral packing, PackOrthoMatrix), :262-352 (GenericTS
//   addressing); PathTracerHelpers.hlsli:227-262 (MatrixRotateFromTo); PathTracerBridgeDonut.hlsli:890-909 (computeMotionVector), :1096-1175.
// One path per pixel is in flight in either pass, so the read-modify-write sequences on a pixel's header, plane records, stable radiance and
// specular hit distance need no atomics - the same exclusivity the reference's one-thread-per-pixel raygen has.
#pragma once

namespace pt {

constexpr uint kStablePlaneCount = 3, kStablePlaneMaxVertexIndex = 15, kMaxDeltaLobes = 3;
constexpr uint kEnqueuedBranchID = 0xFFFFFFFEu, kJustStartedBranchID = 0u;          // kInvalidBranchID, vertexIndexFromBranchID: wavefront.cuh (shared with the denoiser interface's host+device bodies)
constexpr float kEnvironmentMapSceneDistance = 50000.0f * 100.0f;           // Config.h:84-85

PT_DEVICE uint advanceBranchID(uint prev, uint deltaLobe) { return (prev << 2) | deltaLobe; }
PT_DEVICE bool isOnStablePath(uint planeBranchID, uint planeVertexIndex, uint vertexBranchID, uint vertexIndex)
{
    if (vertexIndex > planeVertexIndex) return false;
    return (planeBranchID >> ((planeVertexIndex - vertexIndex) * 2u)) == vertexBranchID;
}

// ---- addressing -------------------------------------------------------------------------------------------------------------------------------
// ---- packing ------------------------------------------------------------------------------------------------------------------------------------
PT_DEVICE uint packTwoHalf(float hi, float lo) { return (f32tof16(clampf(hi, -kHalfMax, kHalfMax)) << 16) | f32tof16(clampf(lo, -kHalfMax, kHalfMax)); }     // PackTwoFp32ToFp16
PT_DEVICE float2 encodeOct(float3 n)
{
    n = n / (fabsf(n.x) + fabsf(n.y) + fabsf(n.z));
    float x = n.x, y = n.y;
    if (!(n.z >= 0.0f)) { const float wx = (1.0f - fabsf(n.y)) * (n.x >= 0.0f ? 1.0f : -1.0f), wy = (1.0f - fabsf(n.x)) * (n.y >= 0.0f ? 1.0f : -1.0f); x = wx; y = wy; }
    return mk2(x * 0.5f + 0.5f, y * 0.5f + 0.5f);
}
PT_DEVICE float3 decodeOct(float2 f)
{
    const float fx = f.x * 2.0f - 1.0f, fy = f.y * 2.0f - 1.0f;
    float3 n = mk3(fx, fy, 1.0f - fabsf(fx) - fabsf(fy));
    const float t = sat(-n.z);
    n.x += (n.x >= 0.0f) ? -t : t; n.y += (n.y >= 0.0f) ? -t : t;
    return norm3(n);
}
// the [0,1] mapping is applied twice on both sides, as in the reference (Utils.hlsli:139-153)
PT_DEVICE uint dirToOctUnorm32(float3 n) { const float2 e = encodeOct(n); return uint(sat(e.x * 0.5f + 0.5f) * float(0xfffe)) | (uint(sat(e.y * 0.5f + 0.5f) * float(0xfffe)) << 16); }
PT_DEVICE uint dirToOctUnorm30(float3 n)
{
    const float2 e = encodeOct(n);
    return (uint(sat(e.x * 0.5f + 0.5f) * float(0x7fff) + 0.5f) & 0x7fffu) | ((uint(sat(e.y * 0.5f + 0.5f) * float(0x7fff) + 0.5f) & 0x7fffu) << 15);
}
PT_DEVICE float3 octUnorm30ToDir(uint u)
{
    const float px = sat(float(u & 0x7fffu) / float(0x7fff)), py = sat(float(u >> 15) / float(0x7fff));
    return decodeOct(mk2(px * 2.0f - 1.0f, py * 2.0f - 1.0f));
}
struct Mat3 { float3 r0, r1, r2; };     // rows, like an HLSL float3x3
PT_DEVICE float3 matCol(const Mat3& m, int c) { return c == 0 ? mk3(m.r0.x, m.r1.x, m.r2.x) : (c == 1 ? mk3(m.r0.y, m.r1.y, m.r2.y) : mk3(m.r0.z, m.r1.z, m.r2.z)); }
PT_DEVICE Mat3 matMul(const Mat3& a, const Mat3& b)
{
    const float3 c0 = matCol(b, 0), c1 = matCol(b, 1), c2 = matCol(b, 2);
    Mat3 o; o.r0 = mk3(dot3(a.r0, c0), dot3(a.r0, c1), dot3(a.r0, c2)); o.r1 = mk3(dot3(a.r1, c0), dot3(a.r1, c1), dot3(a.r1, c2)); o.r2 = mk3(dot3(a.r2, c0), dot3(a.r2, c1), dot3(a.r2, c2));
    return o;
}
PT_DEVICE float3 matVec(const Mat3& m, float3 v) { return mk3(dot3(m.r0, v), dot3(m.r1, v), dot3(m.r2, v)); }
PT_DEVICE Mat3 matTranspose(const Mat3& m) { Mat3 o; o.r0 = matCol(m, 0); o.r1 = matCol(m, 1); o.r2 = matCol(m, 2); return o; }
PT_DEVICE Mat3 matLp(const Mat3& m) { Mat3 o; o.r0 = lp3(m.r0); o.r1 = lp3(m.r1); o.r2 = lp3(m.r2); return o; }
PT_DEVICE void packOrthoMatrix(const Mat3& x, uint& w0, uint& w1)
{
    const uint handedness = dot3(cross3(x.r0, x.r1), x.r2) > 0 ? 1u : 0u;
    w0 = dirToOctUnorm30(x.r0); w1 = dirToOctUnorm30(x.r1) | (handedness << 31);
}
PT_DEVICE Mat3 unpackOrthoMatrix(uint w0, uint w1)
{
    Mat3 x; x.r0 = octUnorm30ToDir(w0); x.r1 = octUnorm30ToDir(w1 & 0x7FFFFFFFu);
    x.r2 = (w1 >> 31) ? cross3(x.r0, x.r1) : cross3(x.r1, x.r0);
    return x;
}
PT_DEVICE Mat3 rotateFromTo(float3 from, float3 to)
{
    Mat3 m; const float e = dot3(from, to);
    if (fabsf(e) > float(1.0f - 1e-10f)) { m.r0 = mk3(1, 0, 0); m.r1 = mk3(0, 1, 0); m.r2 = mk3(0, 0, 1); return m; }
    const float3 v = cross3(from, to);
    const float h = 1.0f / (1.0f + e), hvx = h * v.x, hvz = h * v.z, hvxy = hvx * v.y, hvxz = hvx * v.z, hvyz = hvz * v.y;
    m.r0 = mk3(e + hvx * v.x, hvxy - v.z, hvxz + v.y);
    m.r1 = mk3(hvxy + v.z, e + h * v.y * v.y, hvyz - v.x);
    m.r2 = mk3(hvxz - v.y, hvyz + v.x, e + hvz * v.z);
    return m;
}

// ---- delta lobes and demodulation estimates ---------------------------------------------------------------------------------------------------
struct DeltaLobe { float3 thp; float3 dir; bool transmission; };
// StandardBSDF::evalDeltaLobes: lobe 0 = delta transmission, lobe 1 = delta reflection (world-space directions); nonDeltaPart = selection
// probability mass of the lobes that are not delta
PT_DEVICE void evalDeltaLobes(const BsdfSetup& b, bool psdExclude, DeltaLobe lobes[2], float& nonDeltaPart)
{
    lobes[0].thp = lobes[1].thp = mk3(0.f); lobes[0].dir = lobes[1].dir = mk3(0.f); lobes[0].transmission = true; lobes[1].transmission = false;
    nonDeltaPart = b.pDR + b.pDT;
    if (b.alphaRefl > 0) nonDeltaPart += b.pSR;
    if (b.alphaTrans > 0) nonDeltaPart += b.pSRT;
    if ((b.pSR + b.pSRT) == 0 || psdExclude) return;
    const float3 wi = b.wi;
    float3 reflDir = mk3(-wi.x, -wi.y, wi.z), transDir = mk3(0.f);
    if (b.alphaRefl == 0) lobes[1].thp = (1 - b.pSRT) * fresnelSchlick3(b.specAlbedo, 1.f, wi.z);
    if (b.alphaTrans == 0.f)
    {
        float cosThetaT;
        float F = fresnelDielectric(b.eta, wi.z, cosThetaT);
        lobes[1].thp = lobes[1].thp + mk3(b.pSRT * F);
        float actualEta = b.eta;
        if (b.thin) { actualEta = 1.0f; F = fresnelDielectric(actualEta, wi.z, cosThetaT); }
        transDir = mk3(-wi.x * actualEta, -wi.y * actualEta, -cosThetaT);
        lobes[0].thp = b.transAlbedo * (b.pSRT * (1.0f - F));
    }
    lobes[0].dir = b.fromLocal(transDir); lobes[1].dir = b.fromLocal(reflDir);
}
PT_DEVICE float3 approxSpecularIntegralGGX(float3 specularReflectance, float alpha, float cosTheta)        // SmithGGXCorrelated coefficients
{
    cosTheta = fabsf(cosTheta);
    const float X0 = 1.f, X1 = cosTheta, X2 = cosTheta * cosTheta, X3 = cosTheta * (cosTheta * cosTheta);
    const float Y0 = 1.f, Y1 = alpha, Y3 = alpha * (alpha * alpha);
    const float b1 = (0.995367f * X0 + -1.38839f * X1) * Y0 + (-0.24751f * X0 + 1.97442f * X1) * Y1;
    const float b2 = ((1.0f * X0 + 2.68132f * X1) + 52.366f * X3) * Y0 + ((16.0932f * X0 + -3.98452f * X1) + 59.3013f * X3) * Y1 + ((-5.18731f * X0 + 255.259f * X1) + 2544.07f * X3) * Y3;
    const float s1 = (-0.0564526f * X0 + 3.82901f * X1) * Y0 + (16.91f * X0 + -11.0303f * X1) * Y1;
    const float s2 = ((1.0f * X0 + 4.11118f * X2) + -1.37886f * X3) * Y0 + ((19.3254f * X0 + -28.9947f * X2) + 16.9514f * X3) * Y1 + ((0.545386f * X0 + 96.0994f * X2) + -79.4492f * X3) * Y3;
    float bias = b1 * (1.0f / b2);
    const float scale = s1 * (1.0f / s2);
    bias *= sat(dot3(specularReflectance, mk3(1.f / 3.f)) * 50.0f);
    return specularReflectance * fmaxf(0.0f, scale) + mk3(fmaxf(0.0f, bias));
}
PT_DEVICE void estimateSpecDiffBSDF(const BsdfParams& d, float3 normal, float3 view, float3& outDiff, float3& outSpec)
{
    const float alpha = d.roughness * d.roughness;
    const float roughness = alpha < kMinGGXAlpha ? 0.f : d.roughness;
    const float dT = d.diffuseTransmission, sT = d.specularTransmission;
    const float3 diffuseReflectionAlbedo = lp3(lp(lp(1.f - dT) * lp(1.f - sT)) * d.diffuse);
    const float3 diffuseTransmissionAlbedo = lp3(lp3(dT * d.transmission) * lp(1.f - sT));
    const float3 specularReflectionAlbedo = lp3(lp(1.f - sT) * d.specular);
    const float3 specularTransmissionAlbedo = lp3(sT * d.transmission);
    outDiff = lp3(diffuseReflectionAlbedo + diffuseTransmissionAlbedo);
    outSpec = approxSpecularIntegralGGX(specularReflectionAlbedo, roughness * roughness, sat(dot3(normal, view))) + specularTransmissionAlbedo;
}

// ---- StablePlanesContext ---------------------------------------------------------------------------------------------------------------------------
PT_DEVICE void accumulateStableRadiance(const LaunchParams& p, uint id, float3 r)       // StableRadianceUAV[pixelPos].xyz += radiance on an RGBA16F target
{
    uint2& w = p.rt.stableRadiance[pixelOffset(p, id)];
    const uint2 v = w;
    w = make_uint2(f32tof16(f16tof32(v.x) + r.x) | (f32tof16(f16tof32(v.x >> 16) + r.y) << 16), f32tof16(f16tof32(v.y) + r.z) | (v.y & 0xFFFF0000u));
}
PT_DEVICE float3 computeMotionVector(const RealtimeParams& rt, float3 posW, float3 prevPosW)
{
#ifdef PT_HOST_EMU
    return emuMotionVector(posW, prevPosW);
#endif
    const float* M = rt.worldToClipNoOffset; const float* Q = rt.prevWorldToClipNoOffset;
    const float cx = ((posW.x * M[0] + posW.y * M[4]) + posW.z * M[8]) + M[12], cy = ((posW.x * M[1] + posW.y * M[5]) + posW.z * M[9]) + M[13], cw = ((posW.x * M[3] + posW.y * M[7]) + posW.z * M[11]) + M[15];
    const float qx = ((prevPosW.x * Q[0] + prevPosW.y * Q[4]) + prevPosW.z * Q[8]) + Q[12], qy = ((prevPosW.x * Q[1] + prevPosW.y * Q[5]) + prevPosW.z * Q[9]) + Q[13], qw = ((prevPosW.x * Q[3] + prevPosW.y * Q[7]) + prevPosW.z * Q[11]) + Q[15];
    if (cw <= 0 || qw <= 0) return mk3(0.f);
    return mk3((qx / qw - cx / cw) * rt.clipToWindowScale[0], (qy / qw - cy / cw) * rt.clipToWindowScale[1], qw - cw);
}
PT_DEVICE void exportDominantGuides(const LaunchParams& p, uint id, float3 virtualWorldPos, float3 motion, uint packedThroughput)       // Bridge::ExportSurface / ExportNonSurface
{
    const float* M = p.worldToClip;
    const float z = virtualWorldPos.x * M[2] + virtualWorldPos.y * M[6] + virtualWorldPos.z * M[10] + M[14], w = virtualWorldPos.x * M[3] + virtualWorldPos.y * M[7] + virtualWorldPos.z * M[11] + M[15];
    const size_t o = pixelOffset(p, id);
    p.depth[o] = z / w; p.throughput[o] = packedThroughput;
    p.motionVectors[o] = make_uint2(f32tof16(motion.x) | (f32tof16(motion.y) << 16), f32tof16(motion.z));
}
PT_DEVICE void storeStablePlane(const LaunchParams& p, uint id, uint planeIndex, uint vertexIndex, float3 rayOrigin, float3 rayDir, uint stableBranchID, float sceneLength, float rayT,
                                float3 thp, float3 motion, float roughness, float3 worldNormal, float3 diffEstimate, float3 specEstimate, bool dominant)
{
    uint4* rec = reinterpret_cast<uint4*>(p.rt.planes + planeAddress(p.rt, id, planeIndex));
    rec[0] = make_uint4(__float_as_uint(rayOrigin.x), __float_as_uint(rayOrigin.y), __float_as_uint(rayOrigin.z), __float_as_uint(rayT));
    rec[1] = make_uint4(__float_as_uint(rayDir.x), __float_as_uint(rayDir.y), __float_as_uint(rayDir.z), __float_as_uint(sceneLength));
    rec[2] = make_uint4(packTwoHalf(thp.x, motion.x), packTwoHalf(thp.y, motion.y), packTwoHalf(thp.z, motion.z), (vertexIndex << 16) | f32tof16(roughness));
    const float lo = 0.04f, hi = 6.5504e+4f;        // kNRDMinReflectance / kNRDMaxReflectance
    rec[3] = make_uint4(packTwoHalf(clampf(diffEstimate.x, lo, hi), clampf(specEstimate.x, lo, hi)), packTwoHalf(clampf(diffEstimate.y, lo, hi), clampf(specEstimate.y, lo, hi)),
                        packTwoHalf(clampf(diffEstimate.z, lo, hi), clampf(specEstimate.z, lo, hi)), dirToOctUnorm32(worldNormal));
    rec[4] = make_uint4(0u, 0u, 0u, 0u);            // no noisy radiance yet; FlagsAndVertexIndex / PackedCounters are written as 0 by the reference
    headerWord(p, id, planeIndex) = stableBranchID;
    if (dominant && planeIndex != 0) { uint& h = headerWord(p, id, 3); h = (h & 0xFFFFFFFCu) | (3u & planeIndex); }
}
// CommitDenoiserRadiance (StablePlanes.hlsli:232-252): the path's L (radiance rgb + specular average, fp16) is added to its plane's record
PT_DEVICE void commitDenoiserRadiance(const LaunchParams& p, PathRegs& path)
{
    uint2* w = reinterpret_cast<uint2*>(p.rt.planes + planeAddress(p.rt, path.id, path.stablePlaneIndex())) + 8;      // PackedNoisyRadianceAndSpecAvg at byte 64
    const uint2 e = *w;
    float4 a = path.L();
    if (e.x != 0 && e.y != 0) a = make_float4(a.x + f16tof32(e.x), a.y + f16tof32(e.x >> 16), a.z + f16tof32(e.y), a.w + f16tof32(e.y >> 16));
    *w = make_uint2(packHalf2Clamp(a.x, a.y), packHalf2Clamp(a.z, a.w));
    path.setL(make_float4(0.f, 0.f, 0.f, 0.f));
}
PT_DEVICE void exportSpecHitTStop(const LaunchParams& p, const PathRegs& path)
{
    float& t = p.rt.specularHitT[pixelOffset(p, path.id)];
    if (t < 0) t = fmaxf(0.0f, path.sceneLength + t);
}

// ---- BUILD pass ------------------------------------------------------------------------------------------------------------------------------------------
// the enqueued branch's path state is parked in its plane record in the reference's payload order (PathPayload.hlsli:29-65, StablePlanes.hlsli:341-369)
PT_DEVICE void storeExplorationStart(const LaunchParams& p, const PathRegs& q, uint planeIndex)
{
    uint4* rec = reinterpret_cast<uint4*>(p.rt.planes + planeAddress(p.rt, q.id, planeIndex));
    rec[0] = make_uint4(__float_as_uint(q.origin.x), __float_as_uint(q.origin.y), __float_as_uint(q.origin.z), q.id);
    rec[1] = make_uint4(__float_as_uint(q.dir.x), __float_as_uint(q.dir.y), __float_as_uint(q.dir.z), __float_as_uint(q.sceneLength));
    rec[2] = make_uint4(q.thpXY, q.thpZ, q.lXY, q.lZW);
    rec[3] = make_uint4(q.interior0, q.interior1, q.packedCounters, q.sampleIndex);
    rec[4] = make_uint4(q.rayCone, q.pack0, q.pack1, q.flagsAndVertexIndex);
    headerWord(p, q.id, planeIndex) = kEnqueuedBranchID;
}
PT_DEVICE void explorationStart(const LaunchParams& p, PathRegs& q, uint id, uint planeIndex)
{
    const uint4* rec = reinterpret_cast<const uint4*>(p.rt.planes + planeAddress(p.rt, id, planeIndex));
    const uint4 a = rec[0], b = rec[1], c = rec[2], d = rec[3], e = rec[4];
    q.origin = mk3(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z)); q.id = a.w;
    q.dir = mk3(__uint_as_float(b.x), __uint_as_float(b.y), __uint_as_float(b.z)); q.sceneLength = __uint_as_float(b.w);
    q.thpXY = c.x; q.thpZ = c.y; q.lXY = c.z; q.lZW = c.w;
    q.interior0 = d.x; q.interior1 = d.y; q.packedCounters = d.z; q.sampleIndex = d.w;
    q.rayCone = e.x; q.pack0 = e.y; q.pack1 = e.z; q.flagsAndVertexIndex = e.w;
    headerWord(p, id, planeIndex) = kJustStartedBranchID;
}

PT_DEVICE PathRegs splitDeltaPath(const LaunchParams& p, const PathRegs& oldPath, float3 rayDir, const Surface& s, const DeltaLobe& lobe, uint deltaLobeIndex, bool verifyDominantFlag)
{
    PathRegs q = oldPath;
    q.dir = lobe.dir;
    q.setThp(q.thp() * lobe.thp);
    q.origin = offsetRayOrigin(s.posW, lobe.transmission ? -s.faceN : s.faceN);
    q.stableBranchID() = advanceBranchID(oldPath.sampleIndex, deltaLobeIndex);
    q.setFlag(kPFDelta, true);
    if (!lobe.transmission) q.setFlag(kPFSpecular, true);
    else
    {
        q.setFlag(kPFTransmission, true);
        if (p.c.nestedDielectricsQuality > 0 && !s.thin)
        {
            interiorHandleIntersection(q, s.materialID, s.nestedPriority, s.frontFacing);
            q.setFlag(kPFInsideDielectric, q.interior0 != 0);
        }
    }
    if (__uint_as_float(q.pack0) == 0)      // GetMotionVectorSceneLength() == 0: not behind a surface that blocks motion vectors
    {
        Mat3 localT;        // lpfloat3x3 in the reference: fp16 elements; products accumulated in fp32 and rounded once per element
        if (lobe.transmission) localT = matLp(rotateFromTo(lobe.dir, rayDir));
        else
        {
            Mat3 toTangent; toTangent.r0 = lp3(s.T); toTangent.r1 = lp3(s.B); toTangent.r2 = lp3(s.N);
            Mat3 mirrored = toTangent; mirrored.r2 = -toTangent.r2;            // mul(mirror, toTangent): the z row changes sign
            localT = matLp(matMul(matTranspose(toTangent), mirrored));
        }
        const Mat3 x = matMul(unpackOrthoMatrix(q.lXY, q.lZW), localT);
        packOrthoMatrix(x, q.lXY, q.lZW);
    }
    if (verifyDominantFlag && q.hasFlag(kPFStablePlaneOnDominantBranch))
        if (int(deltaLobeIndex) != int(s.psdDominantDeltaLobeP1) - 1) q.setFlag(kPFStablePlaneOnDominantBranch, false);
    return q;
}

PT_DEVICE void stablePlanesHandleHit(const LaunchParams& p, PathRegs& path, float3 rayOrigin, float3 rayDir, float rayT, const Surface& s, const BsdfSetup& bsdf, bool pathStopping)
{
    const uint vertexIndex = path.vertexIndex(), currentSPIndex = path.stablePlaneIndex(), id = path.id;
    if (s.psdBlockMVs && __uint_as_float(path.pack0) == 0) path.pack0 = __float_as_uint(path.sceneLength);
    if (vertexIndex == 1) headerWord(p, id, 3) = __float_as_uint(fminf(kMaxRayTravel, path.sceneLength)) & 0xFFFFFFFCu;     // StoreFirstHitRayLengthAndClearDominantToZero
    bool setAsBase = true;
    if (vertexIndex < p.rt.maxVertexDepth && !pathStopping)
    {
        DeltaLobe lobes[2]; float nonDeltaPart;
        evalDeltaLobes(bsdf, s.psdExclude, lobes, nonDeltaPart);
        const bool hasNonDeltaLobes = nonDeltaPart > 1e-5f;
        int nonZero[2] = { 0, 0 }; int nonZeroCount = 0; bool potentiallyVolumeTransmission = false;
        #pragma unroll
        for (int k = 0; k < 2; k++) if (average(lobes[k].thp) > 0.001f) { nonZero[nonZeroCount++] = k; potentiallyVolumeTransmission |= lobes[k].transmission; }
        if (nonZeroCount > 0)
        {
            bool allowPSR = p.rt.allowPSR && (nonZeroCount == 1) && (currentSPIndex == 0) && !potentiallyVolumeTransmission;
            allowPSR = allowPSR && !s.psdBlockMVs;
            bool canReuseExisting = (currentSPIndex != 0);
            canReuseExisting = canReuseExisting || allowPSR;
            canReuseExisting = canReuseExisting && !hasNonDeltaLobes;
            int available[2]; int availableCount = 0;       // GetAvailableEmptyPlanes: planes 1..active-1 whose branch ID is still invalid
            for (uint i = 1; i < min(p.rt.activePlaneCount, kStablePlaneCount); i++) if (headerWord(p, id, i) == kInvalidBranchID) available[availableCount++] = int(i);
            canReuseExisting = canReuseExisting && ((currentSPIndex == 0) || (s.psdDominantDeltaLobeP1 > 0));
            nonZeroCount = min(nonZeroCount, availableCount + (canReuseExisting ? 1 : 0));
            int lobeForReuse = -1;
            if (canReuseExisting) { lobeForReuse = nonZero[nonZeroCount - 1]; nonZeroCount--; }
            for (int i = 0; i < nonZeroCount; i++)
            {
                const int k = nonZero[i];
                PathRegs split = splitDeltaPath(p, path, rayDir, s, k == 0 ? lobes[0] : lobes[1], uint(k), true);
                split.setStablePlaneIndex(uint(available[i]));
                storeExplorationStart(p, split, uint(available[i]));
            }
            if (lobeForReuse != -1)
            {
                setAsBase = false;
                path = splitDeltaPath(p, path, rayDir, s, lobeForReuse == 0 ? lobes[0] : lobes[1], uint(lobeForReuse), nonZeroCount > 0);
            }
        }
    }
    if (setAsBase)
    {
        float3 co, cd; computeCameraRay(p.c, id, p.firstSampleIndex, co, cd);
        const Mat3 imageXform = unpackOrthoMatrix(path.lXY, path.lZW);
        const float mvLength = __uint_as_float(path.pack0);
        const bool blockedAtSurface = mvLength != 0;
        const float sceneLengthForMVs = blockedAtSurface ? mvLength : path.sceneLength;
        const float3 virtualWorldPos = co + cd * sceneLengthForMVs;
        const float3 virtualWorldMotion = matVec(imageXform, s.prevPosW - s.posW);      // actual world-space motion (instance.prevTransform / previous-position stream) seen through the stacked reflections (PathTracerStablePlanes.hlsli:286-288)
        const float3 motion = computeMotionVector(p.rt, virtualWorldPos, virtualWorldPos + virtualWorldMotion);
        float roughness = sat(s.bsdf.roughness);
        const float3 worldNormal = norm3(matVec(imageXform, s.N));
        float3 diffEstimate, specEstimate; estimateSpecDiffBSDF(s.bsdf, s.N, s.V, diffEstimate, specEstimate);
        if (blockedAtSurface) roughness *= 0.25f * 0.95f;
        const bool isDominant = path.hasFlag(kPFStablePlaneOnDominantBranch);
        storeStablePlane(p, id, currentSPIndex, vertexIndex, rayOrigin, rayDir, path.sampleIndex, path.sceneLength, rayT, path.thp(), motion, roughness, worldNormal, diffEstimate, specEstimate, isDominant);
        if (isDominant) { const float3 t = path.thp(); exportDominantGuides(p, id, virtualWorldPos, motion, packR11G11B10(mk3(sat(t.x), sat(t.y), sat(t.z)))); }
        path.setFlag(kPFActive, false);
    }
}

PT_DEVICE void stablePlanesHandleMiss(const LaunchParams& p, PathRegs& path, float3 emission, float3 rayOrigin, float3 rayDir)
{
    const uint id = path.id, vertexIndex = path.vertexIndex();
    if (vertexIndex == 1) headerWord(p, id, 3) = __float_as_uint(kMaxRayTravel) & 0xFFFFFFFCu;
    float3 co, cd; computeCameraRay(p.c, id, p.firstSampleIndex, co, cd);
    const float mvLength = __uint_as_float(path.pack0);
    const bool blockedAtSurface = mvLength != 0;
    const float sceneLengthForMVs = blockedAtSurface ? mvLength : kEnvironmentMapSceneDistance;
    const float3 virtualWorldPos = co + cd * sceneLengthForMVs;
    const float3 motion = computeMotionVector(p.rt, virtualWorldPos, virtualWorldPos);
    const bool isDominant = path.hasFlag(kPFStablePlaneOnDominantBranch);
    const float lum = fmaxf(1e-7f, fmaxf(fmaxf(emission.x, emission.y), emission.z));        // ReinhardMax
    const float3 r = emission * ((lum / (lum + 1)) / lum);
    const float3 skyAlbedo = mk3(sqrtf(r.x), sqrtf(r.y), sqrtf(r.z));
    storeStablePlane(p, id, path.stablePlaneIndex(), vertexIndex, rayOrigin, rayDir, path.sampleIndex, blockedAtSurface ? sceneLengthForMVs : __uint_as_float(0x7F800000u), 0.0f, path.thp(), motion,
                     blockedAtSurface ? 0.1f : 1.0f, -rayDir, skyAlbedo, blockedAtSurface ? mk3(0.5f) : mk3(0.f), isDominant);
    if (isDominant) exportDominantGuides(p, id, virtualWorldPos, motion, 0u);
}

// ---- FILL pass --------------------------------------------------------------------------------------------------------------------------------------------
PT_DEVICE void stablePlanesOnScatter(const LaunchParams& p, PathRegs& path, uint lobe)
{
    const uint id = path.id;
    if (path.hasFlag(kPFStablePlaneOnPlane)) path.setFlag(kPFStablePlaneBaseScatterDiff, (lobe & (kLobeDiffuseReflection | kLobeDiffuseTransmission)) != 0);
    path.setFlag(kPFStablePlaneOnPlane, false);
    const uint nextVertexIndex = path.vertexIndex() + 1;
    if (path.hasFlag(kPFStablePlaneOnBranch) && nextVertexIndex <= kStablePlaneMaxVertexIndex)
    {
        const uint deltaLobeIndex = ((lobe & kLobeDelta) == 0u) ? 0xFFFFFFFFu : (((lobe & kLobeTransmission) == 0u) ? 1u : 0u);      // BSDFSample::getDeltaLobeIndex
        path.stableBranchID() = advanceBranchID(path.sampleIndex, deltaLobeIndex);
        bool onStablePath = false;
        for (uint spi = 0; spi < kStablePlaneCount; spi++)
        {
            const uint planeBranchID = headerWord(p, id, spi);
            if (planeBranchID == kInvalidBranchID) continue;
            if (planeBranchID == path.sampleIndex)
            {
                commitDenoiserRadiance(p, path);
                path.setStablePlaneIndex(spi);
                path.setFlag(kPFStablePlaneOnDominantBranch, spi == (headerWord(p, id, 3) & 3u));
                path.setFlag(kPFStablePlaneOnPlane, true);
                path.setCounter(kCtrBouncesFromStablePlane, 0);
                onStablePath = true;
                break;
            }
            onStablePath = onStablePath || isOnStablePath(planeBranchID, vertexIndexFromBranchID(planeBranchID), path.sampleIndex, nextVertexIndex);
        }
        path.setFlag(kPFStablePlaneOnBranch, onStablePath);
    }
    else
    {
        path.stableBranchID() = kInvalidBranchID;
        path.setFlag(kPFStablePlaneOnBranch, false);
        path.incrementCounter(kCtrBouncesFromStablePlane);
    }
    if (!path.hasFlag(kPFStablePlaneOnPlane)) path.incrementCounter(kCtrBouncesFromStablePlane);
}

} // namespace pt
