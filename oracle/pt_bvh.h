// ORACLE — test infrastructure only (see pt_math.h).
// pt_bvh.h: ray queries.  The reference has NO source for this part: BVH build, traversal and the ray/triangle test live in the
// DXR runtime + driver behind RayQuery::TraceRayInline (Rtxpt/Shaders/PathTracerBridgeDonut.hlsli:995-996,1029-1034) and
// buildBottomLevelAccelStruct (Rtxpt/Sample.cpp:1061-1079).  "parity unpinned" at this boundary: DXR only guarantees
// watertightness and closest-hit semantics.  What is restated here is the published algorithm the product also follows:
//   S. Woop, C. Benthin, I. Wald, "Watertight Ray/Triangle Intersection", JCGT 2(1), 2013   (shear + edge functions, fp64 fallback)
// over a plain binned-SAH BVH2 (independent of the product's compressed 8-wide BVH), with the any-hit alpha test of
//   PathTracerBridgeDonut.hlsli:929-989, :993-1055.
// Conventions shared with the product so that hit records are bit-identical: no FMA contraction, t = T/det, (u,v) = (V/det, W/det),
// hits accepted for tMin < t < tMax, equal-t ties resolved towards the smaller global triangle id.
#pragma once
#include "pt_scene.h"
#include <vector>
#include <numeric>

namespace orc {

struct Tri { float3 v0, v1, v2; uint instanceIndex, geometryIndex, primitiveIndex, subInstanceIndex; };
struct Hit { float t = -1.0f; float u = 0, v = 0; uint triId = 0xFFFFFFFFu; bool valid() const { return triId != 0xFFFFFFFFu; } };

struct RayShear     // per-ray constants of the watertight test
{
    int kx, ky, kz; float Sx, Sy, Sz;
    explicit RayShear(float3 d)
    {
        float a[3] = { fabsf(d.x), fabsf(d.y), fabsf(d.z) };
        kz = (a[0] >= a[1]) ? ((a[0] >= a[2]) ? 0 : 2) : ((a[1] >= a[2]) ? 1 : 2);
        kx = (kz + 1) % 3; ky = (kx + 1) % 3;
        float dd[3] = { d.x, d.y, d.z };
        if (dd[kz] < 0.0f) std::swap(kx, ky);
        Sx = dd[kx] / dd[kz]; Sy = dd[ky] / dd[kz]; Sz = 1.0f / dd[kz];
    }
};

inline bool intersectTriWatertight(const RayShear& rs, float3 org, const Tri& tr, float tMin, float tMax, float& tOut, float& uOut, float& vOut)
{
    const float A[3] = { tr.v0.x - org.x, tr.v0.y - org.y, tr.v0.z - org.z };
    const float B[3] = { tr.v1.x - org.x, tr.v1.y - org.y, tr.v1.z - org.z };
    const float C[3] = { tr.v2.x - org.x, tr.v2.y - org.y, tr.v2.z - org.z };
    const float Ax = A[rs.kx] - rs.Sx * A[rs.kz], Ay = A[rs.ky] - rs.Sy * A[rs.kz];
    const float Bx = B[rs.kx] - rs.Sx * B[rs.kz], By = B[rs.ky] - rs.Sy * B[rs.kz];
    const float Cx = C[rs.kx] - rs.Sx * C[rs.kz], Cy = C[rs.ky] - rs.Sy * C[rs.kz];
    float U = Cx * By - Cy * Bx, V = Ax * Cy - Ay * Cx, W = Bx * Ay - By * Ax;
    if (U == 0.0f || V == 0.0f || W == 0.0f)
    {
        double CxBy = double(Cx) * double(By), CyBx = double(Cy) * double(Bx); U = float(CxBy - CyBx);
        double AxCy = double(Ax) * double(Cy), AyCx = double(Ay) * double(Cx); V = float(AxCy - AyCx);
        double BxAy = double(Bx) * double(Ay), ByAx = double(By) * double(Ax); W = float(BxAy - ByAx);
    }
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
    const float det = U + V + W;
    if (det == 0.0f) return false;
    const float Az = rs.Sz * A[rs.kz], Bz = rs.Sz * B[rs.kz], Cz = rs.Sz * C[rs.kz];
    const float T = U * Az + V * Bz + W * Cz;
    const float t = T / det;
    if (!(t > tMin && t < tMax)) return false;
    tOut = t; uOut = V / det; vOut = W / det;
    return true;
}

struct Bvh2
{
    struct Node { float lo[3], hi[3]; uint left, count; };      // count>0: leaf [left, left+count) into order[]; else children left, left+1
    std::vector<Tri> tris;
    std::vector<uint> order;
    std::vector<Node> nodes;

    void gather(const Scene& sc)
    {
        const RtxptSceneDesc& d = *sc.desc;
        for (uint ii = 0; ii < d.instanceCount; ii++)
        {
            const RtxptInstanceData& inst = d.instances[ii];
            for (uint gi = 0; gi < inst.numGeometries; gi++)
            {
                const RtxptGeometryData& g = d.geometries[inst.firstGeometryIndex + gi];
                for (uint t = 0; t < g.numIndices / 3; t++)
                {
                    Tri tr; float3 p[3];
                    for (int k = 0; k < 3; k++)
                    {
                        uint idx = sc.load32(g.indexBufferIndex, g.indexOffset + t * 12 + k * 4);
                        p[k] = mul34_point(inst.transform, sc.loadFloat3(g.vertexBufferIndex, g.positionOffset + idx * 12));
                    }
                    tr.v0 = p[0]; tr.v1 = p[1]; tr.v2 = p[2];
                    tr.instanceIndex = ii; tr.geometryIndex = gi; tr.primitiveIndex = t; tr.subInstanceIndex = inst.firstGeometryInstanceIndex + gi;
                    tris.push_back(tr);
                }
            }
        }
    }

    void build(const Scene& sc)
    {
        gather(sc);
        const uint n = uint(tris.size());
        order.resize(n); std::iota(order.begin(), order.end(), 0u);
        std::vector<float> cen(size_t(n) * 3), blo(size_t(n) * 3), bhi(size_t(n) * 3);
        for (uint i = 0; i < n; i++)
        {
            const Tri& t = tris[i];
            float3 lo = min3v(min3v(t.v0, t.v1), t.v2), hi = max3v(max3v(t.v0, t.v1), t.v2);
            blo[i * 3 + 0] = lo.x; blo[i * 3 + 1] = lo.y; blo[i * 3 + 2] = lo.z;
            bhi[i * 3 + 0] = hi.x; bhi[i * 3 + 1] = hi.y; bhi[i * 3 + 2] = hi.z;
            for (int a = 0; a < 3; a++) cen[i * 3 + a] = 0.5f * (blo[i * 3 + a] + bhi[i * 3 + a]);
        }
        nodes.clear(); nodes.reserve(size_t(n) * 2 + 2);
        nodes.push_back(Node());
        if (n == 0) { nodes[0].count = 0; nodes[0].left = 0; for (int a = 0; a < 3; a++) { nodes[0].lo[a] = 1; nodes[0].hi[a] = -1; } return; }
        struct Job { uint node, first, count; };
        std::vector<Job> stack; stack.push_back({ 0, 0, n });
        while (!stack.empty())
        {
            Job j = stack.back(); stack.pop_back();
            float lo[3] = { 1e30f, 1e30f, 1e30f }, hi[3] = { -1e30f, -1e30f, -1e30f }, clo[3] = { 1e30f, 1e30f, 1e30f }, chi[3] = { -1e30f, -1e30f, -1e30f };
            for (uint i = j.first; i < j.first + j.count; i++)
            {
                uint id = order[i];
                for (int a = 0; a < 3; a++)
                {
                    lo[a] = std::min(lo[a], blo[id * 3 + a]); hi[a] = std::max(hi[a], bhi[id * 3 + a]);
                    clo[a] = std::min(clo[a], cen[id * 3 + a]); chi[a] = std::max(chi[a], cen[id * 3 + a]);
                }
            }
            Node nd;
            for (int a = 0; a < 3; a++)
            {   // pad: the box test only has to be conservative, the triangle test decides
                float pad = 1e-5f * std::max(std::max(fabsf(lo[a]), fabsf(hi[a])), 1.0f);
                nd.lo[a] = lo[a] - pad; nd.hi[a] = hi[a] + pad;
            }
            int axis = 0; float ext = chi[0] - clo[0];
            for (int a = 1; a < 3; a++) if (chi[a] - clo[a] > ext) { ext = chi[a] - clo[a]; axis = a; }
            if (j.count <= 4 || ext <= 0.0f) { nd.left = j.first; nd.count = j.count; nodes[j.node] = nd; continue; }
            // binned SAH on the widest centroid axis
            const int NB = 16;
            struct Bin { float lo[3], hi[3]; uint n; } bins[NB];
            for (int b = 0; b < NB; b++) { bins[b].n = 0; for (int a = 0; a < 3; a++) { bins[b].lo[a] = 1e30f; bins[b].hi[a] = -1e30f; } }
            float scale = float(NB) / ext;
            auto binOf = [&](uint id) { int b = int((cen[id * 3 + axis] - clo[axis]) * scale); return std::min(std::max(b, 0), NB - 1); };
            for (uint i = j.first; i < j.first + j.count; i++)
            {
                uint id = order[i]; Bin& b = bins[binOf(id)]; b.n++;
                for (int a = 0; a < 3; a++) { b.lo[a] = std::min(b.lo[a], blo[id * 3 + a]); b.hi[a] = std::max(b.hi[a], bhi[id * 3 + a]); }
            }
            auto area = [](const float* l, const float* h) { float dx = h[0] - l[0], dy = h[1] - l[1], dz = h[2] - l[2]; return (dx < 0) ? 0.0f : 2.0f * (dx * dy + dy * dz + dz * dx); };
            float rightArea[NB]; uint rightN[NB];
            { float l[3] = { 1e30f, 1e30f, 1e30f }, h[3] = { -1e30f, -1e30f, -1e30f }; uint c = 0;
              for (int b = NB - 1; b > 0; b--) { c += bins[b].n; for (int a = 0; a < 3; a++) { l[a] = std::min(l[a], bins[b].lo[a]); h[a] = std::max(h[a], bins[b].hi[a]); } rightArea[b] = area(l, h); rightN[b] = c; } }
            float best = 1e30f; int bestSplit = -1;
            { float l[3] = { 1e30f, 1e30f, 1e30f }, h[3] = { -1e30f, -1e30f, -1e30f }; uint c = 0;
              for (int b = 0; b < NB - 1; b++) { c += bins[b].n; for (int a = 0; a < 3; a++) { l[a] = std::min(l[a], bins[b].lo[a]); h[a] = std::max(h[a], bins[b].hi[a]); }
                  if (c == 0 || rightN[b + 1] == 0) continue;
                  float cost = area(l, h) * float(c) + rightArea[b + 1] * float(rightN[b + 1]);
                  if (cost < best) { best = cost; bestSplit = b; } } }
            uint mid;
            if (bestSplit < 0) mid = j.first + j.count / 2;
            else mid = uint(std::partition(order.begin() + j.first, order.begin() + j.first + j.count, [&](uint id) { return binOf(id) <= bestSplit; }) - order.begin());
            if (mid == j.first || mid == j.first + j.count) mid = j.first + j.count / 2;
            nd.left = uint(nodes.size()); nd.count = 0;
            nodes[j.node] = nd;
            nodes.push_back(Node()); nodes.push_back(Node());
            stack.push_back({ nd.left, j.first, mid - j.first });
            stack.push_back({ nd.left + 1, mid, j.first + j.count - mid });
        }
    }

    // anyHit: stop at the first accepted hit (RAY_FLAG_ACCEPT_FIRST_HIT_AND_END_SEARCH) and apply ExcludeFromNEE (BridgeDonut:980-989)
    Hit trace(const Scene& sc, float3 org, float3 dir, float tMin, float tMax, bool anyHit, uint64_t* nodeVisits = nullptr, uint64_t* triTests = nullptr) const
    {
        Hit best; best.t = tMax;
        if (tris.empty()) { best.t = -1; return best; }
        RayShear rs(dir);
        float inv[3] = { 1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z };
        float o[3] = { org.x, org.y, org.z };
        uint stack[96]; int sp = 0; stack[sp++] = 0;
        while (sp > 0)
        {
            const Node& nd = nodes[stack[--sp]];
            if (nodeVisits) (*nodeVisits)++;
            float t0 = tMin, t1 = best.t;
            bool miss = false;
            for (int a = 0; a < 3; a++)
            {
                float ta = (nd.lo[a] - o[a]) * inv[a], tb = (nd.hi[a] - o[a]) * inv[a];
                if (ta > tb) std::swap(ta, tb);
                if (ta != ta || tb != tb) continue;     // 0 * inf: ray parallel to and inside the slab plane
                t0 = std::max(t0, ta); t1 = std::min(t1, tb * 1.0000004f);
                if (t0 > t1) { miss = true; break; }
            }
            if (miss) continue;
            if (nd.count > 0)
            {
                for (uint i = nd.left; i < nd.left + nd.count; i++)
                {
                    uint id = order[i];
                    const Tri& tr = tris[id];
                    float t, u, v;
                    if (triTests) (*triTests)++;
                    // tMax passed as +inf-like upper bound then compared manually so equal-t ties can be resolved by id
                    if (!intersectTriWatertight(rs, org, tr, tMin, tMax, t, u, v)) continue;
                    if (best.valid() ? !(t < best.t || (t == best.t && id < best.triId)) : !(t < best.t)) continue;
                    const RtxptSubInstanceData& s = sc.subInstances[tr.subInstanceIndex];
                    // geometry is non-opaque when alpha tested or excluded from NEE (SampleCommon/AccelerationStructureUtil.h:88-89)
                    if (anyHit && (s.FlagsAndAlphaInfo & RTXPT_SUBINST_FLAG_EXCLUDE_FROM_NEE)) continue;
                    if ((s.FlagsAndAlphaInfo & RTXPT_SUBINST_FLAG_ALPHA_TESTED) && !alphaTest(sc, s, tr.primitiveIndex, f2(u, v))) continue;
                    best.t = t; best.u = u; best.v = v; best.triId = id;
                    if (anyHit) return best;
                }
            }
            else
            {
                if (sp + 2 > 96) continue;
                stack[sp++] = nd.left; stack[sp++] = nd.left + 1;
            }
        }
        if (!best.valid()) best.t = -1.0f;
        return best;
    }
};

} // namespace orc
