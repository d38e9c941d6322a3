"""CPU-only: SAH statistics of the product's BVH over the bench scene (or a smaller city), to judge builder changes without a GPU."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, ctypes as C
from rtxpt_b200 import lib, scenes


def scene_triangles(scene):
    d = scene.desc; out = []
    for ii in range(d.instanceCount):
        inst = d.instances[ii]; xf = np.array(inst.transform[:], np.float32).reshape(3, 4)
        for gi in range(inst.numGeometries):
            g = d.geometries[inst.firstGeometryIndex + gi]
            ib = d.buffers[g.indexBufferIndex]; vb = d.buffers[g.vertexBufferIndex]
            idx = np.frombuffer((C.c_uint8 * ib.sizeBytes).from_address(ib.data), np.uint32, g.numIndices, g.indexOffset)
            pos = np.frombuffer((C.c_uint8 * vb.sizeBytes).from_address(vb.data), np.float32, g.numVertices * 3, g.positionOffset).reshape(-1, 3)
            w = pos @ xf[:, :3].T + xf[:, 3]
            out.append(w[idx].reshape(-1, 3, 3))
    return np.concatenate(out).astype(np.float32)


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_800_000
    scene, cam = scenes.city_block(target_triangles=n)
    tris = scene_triangles(scene)
    st = lib.bvh_stats(tris)
    print("triangles %d  refs %d  nodes %d  leaves %d  depth %d  build %.2f s  E[node visits] %.2f  E[tri tests] %.2f" %
          (len(tris), st.triangleReferenceCount, st.nodeCount, st.leafCount, st.maxDepth, st.buildSeconds, st.expectedNodeVisits, st.expectedTriangleTests))
