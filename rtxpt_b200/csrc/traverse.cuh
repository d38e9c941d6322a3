// traverse.cuh — software ray queries: CWBVH8 traversal + watertight ray/triangle test + any-hit alpha test.
// Replaces what the reference delegates to RT cores through RayQuery::TraceRayInline:
//   Bridge::traceScatterRay  (Rtxpt/Shaders/PathTracerBridgeDonut.hlsli:1029-1055)  -> traceRay<false>
//   Bridge::traceVisibilityRay (…:993-1027, RAY_FLAG_ACCEPT_FIRST_HIT_AND_END_SEARCH)  -> traceRay<true>
//   AlphaTestImpl (…:929-971), AlphaTestVisibilityRay ExcludeFromNEE (…:980-989)
// Algorithms: Ylitie/Karras/Laine HPG 2017 (node-group / triangle-group stack, octant-ordered hit masks);
//             Woop/Benthin/Wald JCGT 2013 (watertight test; fp64 fallback when an edge function is exactly zero).
// Conventions shared with the oracle so hit records are bit-identical: no FMA contraction in the triangle test, t = T/det,
// (u,v) = (V/det, W/det), accept tMin < t < tMax, equal-t ties go to the smaller global triangle id.
#pragma once
#include "device_math.cuh"
#include "scene_device.cuh"

namespace pt {

struct HitRecord { float t, u, v; uint gid; };      // gid == 0xFFFFFFFF: miss

constexpr int kTraversalStackSize = 32;

struct WatertightRay
{
    int kx, ky, kz; float Sx, Sy, Sz;
    PT_DEVICE void setup(float3 d)
    {
        float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
        kz = (ax >= ay) ? ((ax >= az) ? 0 : 2) : ((ay >= az) ? 1 : 2);
        kx = (kz + 1) % 3; ky = (kx + 1) % 3;
        float dk = (kz == 0) ? d.x : ((kz == 1) ? d.y : d.z);
        if (dk < 0.0f) { int t = kx; kx = ky; ky = t; }
        float dx = (kx == 0) ? d.x : ((kx == 1) ? d.y : d.z);
        float dy = (ky == 0) ? d.x : ((ky == 1) ? d.y : d.z);
        Sx = __fdiv_rn(dx, dk); Sy = __fdiv_rn(dy, dk); Sz = __fdiv_rn(1.0f, dk);
    }
};

PT_DEVICE float comp(float3 v, int k) { return (k == 0) ? v.x : ((k == 1) ? v.y : v.z); }

PT_DEVICE bool intersectTriangleWatertight(const WatertightRay& wr, float3 org, float3 v0, float3 v1, float3 v2, float tMin, float tMax,
                                           float& tOut, float& uOut, float& vOut)
{
    const float3 A = mk3(__fsub_rn(v0.x, org.x), __fsub_rn(v0.y, org.y), __fsub_rn(v0.z, org.z));
    const float3 B = mk3(__fsub_rn(v1.x, org.x), __fsub_rn(v1.y, org.y), __fsub_rn(v1.z, org.z));
    const float3 C = mk3(__fsub_rn(v2.x, org.x), __fsub_rn(v2.y, org.y), __fsub_rn(v2.z, org.z));
    const float Akz = comp(A, wr.kz), Bkz = comp(B, wr.kz), Ckz = comp(C, wr.kz);
    const float Ax = __fsub_rn(comp(A, wr.kx), __fmul_rn(wr.Sx, Akz)), Ay = __fsub_rn(comp(A, wr.ky), __fmul_rn(wr.Sy, Akz));
    const float Bx = __fsub_rn(comp(B, wr.kx), __fmul_rn(wr.Sx, Bkz)), By = __fsub_rn(comp(B, wr.ky), __fmul_rn(wr.Sy, Bkz));
    const float Cx = __fsub_rn(comp(C, wr.kx), __fmul_rn(wr.Sx, Ckz)), Cy = __fsub_rn(comp(C, wr.ky), __fmul_rn(wr.Sy, Ckz));
    float U = __fsub_rn(__fmul_rn(Cx, By), __fmul_rn(Cy, Bx));
    float V = __fsub_rn(__fmul_rn(Ax, Cy), __fmul_rn(Ay, Cx));
    float W = __fsub_rn(__fmul_rn(Bx, Ay), __fmul_rn(By, Ax));
    if (U == 0.0f || V == 0.0f || W == 0.0f)
    {
        U = float(__dsub_rn(__dmul_rn(double(Cx), double(By)), __dmul_rn(double(Cy), double(Bx))));
        V = float(__dsub_rn(__dmul_rn(double(Ax), double(Cy)), __dmul_rn(double(Ay), double(Cx))));
        W = float(__dsub_rn(__dmul_rn(double(Bx), double(Ay)), __dmul_rn(double(By), double(Ax))));
    }
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
    const float det = __fadd_rn(__fadd_rn(U, V), W);
    if (det == 0.0f) return false;
    const float Az = __fmul_rn(wr.Sz, Akz), Bz = __fmul_rn(wr.Sz, Bkz), Cz = __fmul_rn(wr.Sz, Ckz);
    const float T = __fadd_rn(__fadd_rn(__fmul_rn(U, Az), __fmul_rn(V, Bz)), __fmul_rn(W, Cz));
    const float t = __fdiv_rn(T, det);
    if (!(t > tMin && t < tMax)) return false;
    tOut = t; uOut = __fdiv_rn(V, det); vOut = __fdiv_rn(W, det);
    return true;
}

// AlphaTestImpl (BridgeDonut:929-971): true when the candidate is opaque at (u,v)
PT_DEVICE bool alphaTestPasses(const SceneView& sc, const RtxptSubInstanceData& s, uint primitiveIndex, float u, float v)
{
    const uint ib = s.IndexBufferIndex_VertexBufferIndex >> 16, vb = s.IndexBufferIndex_VertexBufferIndex & 0xFFFF;
    const uint3 idx = loadIndex3(sc, ib, s.IndexOffset + primitiveIndex * 12);
    const float2 t0 = loadFloat2(sc, vb, s.TexCoord1Offset + idx.x * 8), t1 = loadFloat2(sc, vb, s.TexCoord1Offset + idx.y * 8), t2 = loadFloat2(sc, vb, s.TexCoord1Offset + idx.z * 8);
    const float b0 = 1.0f - (u + v);
    const float2 uv = mk2(t0.x * b0 + t1.x * u + t2.x * v, t0.y * b0 + t1.y * u + t2.y * v);
    const float opacity = tex2DLod<float4>(sc.textures[s.FlagsAndAlphaInfo & 0xFFFF], uv.x, uv.y, 0.0f).w;
    return opacity >= float(s.FlagsAndAlphaInfo >> 24) / 255.0f;
}

struct TraversalCounters { uint nodeVisits, triTests; };

// Extracts byte j of a packed word as float
PT_DEVICE float byteToFloat(uint w, int j) { return float((w >> (8 * j)) & 0xFFu); }

// Resumable per-lane traversal state.  A persistent warp keeps one Traverser per lane; run() advances the lane's ray until it finishes
// or until so few lanes of the warp are still working that the warp should return to its caller to fetch new rays for the idle lanes
// (dynamic fetch, Aila & Laine HPG 2009) — the caller then re-enters run() on the unfinished lanes.
// The traversal stack is a separate local array owned by the kernel so that the scalar state below stays in registers.
template <bool ANY_HIT, bool COUNT>
struct Traverser
{
    float3 org, dir;
    float idx, idy, idz, tMin, tMax;
    WatertightRay wr;
    HitRecord best; uint bestSubInstance;
    uint2 nodeGroup, triGroup;
    uint octinv;
    int sp;
    bool done;

    PT_DEVICE void init(const SceneView& sc, float3 o, float3 d, float tmin, float tmax)
    {
        org = o; dir = d; tMin = tmin; tMax = tmax;
        best.t = tmax; best.u = 0; best.v = 0; best.gid = 0xFFFFFFFFu; bestSubInstance = 0;
        wr.setup(d);
        const float eps = 1.0e-30f;
        idx = 1.0f / (fabsf(d.x) > eps ? d.x : copysignf(eps, d.x));
        idy = 1.0f / (fabsf(d.y) > eps ? d.y : copysignf(eps, d.y));
        idz = 1.0f / (fabsf(d.z) > eps ? d.z : copysignf(eps, d.z));
        octinv = 7u - ((d.x < 0.0f ? 4u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 1u : 0u));
        sp = 0;
        nodeGroup = make_uint2(0u, 0x80000000u);        // virtual parent of the root: one internal child in slot 7^octinv
        triGroup = make_uint2(0u, 0u);
        done = (sc.bvhTriCount == 0);
    }

    PT_DEVICE HitRecord result() const { HitRecord r = best; if (r.gid == 0xFFFFFFFFu) r.t = -1.0f; return r; }

    PT_DEVICE void run(const SceneView& sc, const uint4* __restrict__ nodes, const uint4* __restrict__ smemNodes, uint smemNodeCount, int minActiveLanes,
                       TraversalCounters* counters, uint2* __restrict__ stack)
    {
        const uint octinv4 = octinv * 0x01010101u;
        while (!done)
        {
            if (nodeGroup.y & 0xFF000000u)
            {
                const uint hits = nodeGroup.y;
                const uint bitIndex = 31u - __clz(hits & 0xFF000000u);
                nodeGroup.y &= ~(1u << bitIndex);
                if (nodeGroup.y & 0xFF000000u) { if (sp < kTraversalStackSize) stack[sp++] = nodeGroup; }
                const uint slot = (bitIndex - 24u) ^ octinv;
                const uint rel = __popc(hits & ~(0xFFFFFFFFu << slot) & 0xFFu);
                const uint nodeIndex = nodeGroup.x + rel;
                const uint4* np = (nodeIndex < smemNodeCount) ? (smemNodes + nodeIndex * 5) : (nodes + size_t(nodeIndex) * 5);
                const uint4 n0 = np[0], n1 = np[1], n2 = np[2], n3 = np[3], n4 = np[4];
                if (COUNT) counters->nodeVisits++;

                const float px = __uint_as_float(n0.x), py = __uint_as_float(n0.y), pz = __uint_as_float(n0.z);
                const float sx = __uint_as_float((n0.w & 0xFFu) << 23), sy = __uint_as_float(((n0.w >> 8) & 0xFFu) << 23), sz = __uint_as_float(((n0.w >> 16) & 0xFFu) << 23);
                const uint imask = n0.w >> 24;
                nodeGroup.x = n1.x; triGroup.x = n1.y;
                const float adx = sx * idx, ady = sy * idy, adz = sz * idz;
                const float ox = (px - org.x) * idx, oy = (py - org.y) * idy, oz = (pz - org.z) * idz;
                uint hitmask = 0;
                #pragma unroll
                for (int half = 0; half < 2; half++)
                {
                    const uint meta4 = half ? n1.w : n1.z;
                    const uint isInner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
                    const uint innerMask4 = (isInner4 >> 4) * 0xFFu;
                    const uint bitIndex4 = (meta4 ^ (octinv4 & innerMask4)) & 0x1F1F1F1Fu;
                    const uint childBits4 = (meta4 >> 5) & 0x07070707u;
                    const uint qlox = half ? n2.y : n2.x, qloy = half ? n2.w : n2.z, qloz = half ? n3.y : n3.x;
                    const uint qhix = half ? n3.w : n3.z, qhiy = half ? n4.y : n4.x, qhiz = half ? n4.w : n4.z;
                    const uint nearx = (dir.x < 0.0f) ? qhix : qlox, farx = (dir.x < 0.0f) ? qlox : qhix;
                    const uint neary = (dir.y < 0.0f) ? qhiy : qloy, fary = (dir.y < 0.0f) ? qloy : qhiy;
                    const uint nearz = (dir.z < 0.0f) ? qhiz : qloz, farz = (dir.z < 0.0f) ? qloz : qhiz;
                    #pragma unroll
                    for (int j = 0; j < 4; j++)
                    {
                        // slab distances with one FMA each; the box test only has to be conservative (the triangle test decides), see the slack below
                        const float t0x = __fmaf_rn(byteToFloat(nearx, j), adx, ox), t1x = __fmaf_rn(byteToFloat(farx, j), adx, ox);
                        const float t0y = __fmaf_rn(byteToFloat(neary, j), ady, oy), t1y = __fmaf_rn(byteToFloat(fary, j), ady, oy);
                        const float t0z = __fmaf_rn(byteToFloat(nearz, j), adz, oz), t1z = __fmaf_rn(byteToFloat(farz, j), adz, oz);
                        float cmin = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, tMin));
                        float cmax = fminf(fminf(t1x, t1y), fminf(t1z, best.t));
                        cmin = __fmaf_rn(-fabsf(cmin), 6.0e-7f, cmin); cmax = __fmaf_rn(fabsf(cmax), 6.0e-7f, cmax);
                        if (cmin <= cmax)
                            hitmask |= ((childBits4 >> (8 * j)) & 0xFFu) << ((bitIndex4 >> (8 * j)) & 0xFFu);
                    }
                }
                nodeGroup.y = (hitmask & 0xFF000000u) | imask;
                triGroup.y = hitmask & 0x00FFFFFFu;
            }
            else
            {
                triGroup = nodeGroup;
                nodeGroup = make_uint2(0u, 0u);
            }

            while (triGroup.y != 0)
            {
                const uint k = __ffs(triGroup.y) - 1;
                triGroup.y &= triGroup.y - 1;
                const float4* tp = sc.bvhTris + size_t(triGroup.x + k) * 3;
                const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
                if (COUNT) counters->triTests++;
                float t, u, v;
                if (!intersectTriangleWatertight(wr, org, mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), mk3(c.x, c.y, c.z), tMin, tMax, t, u, v)) continue;
                const uint gid = __float_as_uint(a.w);
                if (best.gid != 0xFFFFFFFFu ? !(t < best.t || (t == best.t && gid < best.gid)) : !(t < best.t)) continue;
                const uint sub = __float_as_uint(b.w);
                if (sub & (kTriFlagAlphaTested | kTriFlagExcludeFromNEE))
                {   // non-opaque geometry (SampleCommon/AccelerationStructureUtil.h:88-89)
                    if (ANY_HIT && (sub & kTriFlagExcludeFromNEE)) continue;
                    if ((sub & kTriFlagAlphaTested) && !alphaTestPasses(sc, sc.subInstances[sub & kTriSubInstanceMask], __float_as_uint(c.w), u, v)) continue;
                }
                best.t = t; best.u = u; best.v = v; best.gid = gid; bestSubInstance = sub & kTriSubInstanceMask;
                if (ANY_HIT) { done = true; return; }
            }

            if ((nodeGroup.y & 0xFF000000u) == 0)
            {
                if (sp == 0) { done = true; return; }
                nodeGroup = stack[--sp];
            }
            if (__popc(__activemask()) < minActiveLanes) return;        // let the warp refill its idle lanes
        }
    }
};

} // namespace pt
