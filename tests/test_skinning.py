"""CPU: skinned-mesh animation: the oracle's restatement of Donut's skinning pass (oracle/pt_skinning.h) held to what linear blend skinning guarantees, and the
product's bodies (rtxpt_b200/csrc/skinning.cuh, host build in tests/emu) equal to the oracle.  GPU: tests/test_gpu_skinning.py (-m gpu)."""
import ctypes as C
import numpy as np
import pytest


def _snorm(v): return (np.clip(v, -1, 1) * 127).astype(np.int32) & 0xFF           # Pack_R8_SNORM truncates


def _pack4(n): q = _snorm(n); return (q[:, 0] | (q[:, 1] << 8) | (q[:, 2] << 16) | (q[:, 3] << 24)).astype(np.uint32)


def _unpack3(u): s = np.stack([(u << 24).astype(np.int32) >> 24, (u << 16).astype(np.int32) >> 24, (u << 8).astype(np.int32) >> 24], -1); return np.clip(s / 127.0, -1, 1)


def _mesh(rng, nv=300, nt=500, joints=6):
    pos = rng.normal(0, 1, (nv, 3)).astype(np.float32)
    n = rng.normal(0, 1, (nv, 3)); n /= np.linalg.norm(n, axis=1, keepdims=True); t = np.cross(n, rng.normal(0, 1, (nv, 3))); t /= np.linalg.norm(t, axis=1, keepdims=True)
    nrm = _pack4(np.concatenate([n, np.zeros((nv, 1))], 1)); tan = _pack4(np.concatenate([t, np.where(rng.random((nv, 1)) < 0.5, -1.0, 1.0)], 1))
    ji = rng.integers(0, joints, (nv, 4)).astype(np.uint16)
    jw = rng.random((nv, 4)).astype(np.float32); jw[rng.random((nv, 4)) < 0.4] = 0; jw[:, 0] += 0.05; jw /= jw.sum(1, keepdims=True)
    idx = rng.integers(0, nv, (nt, 3)).astype(np.uint32)
    return pos, nrm, tan, ji, jw, idx


def _mats(rng, joints=6):
    out = []
    for k in range(joints):
        a, ax = rng.uniform(-1, 1), rng.normal(0, 1, 3); ax /= np.linalg.norm(ax)
        K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]]); R = np.eye(3) + np.sin(a) * K + (1 - np.cos(a)) * K @ K
        M = np.eye(4); M[:3, :3] = R.T; M[3, :3] = rng.uniform(-2, 2, 3)            # row vector x matrix: v' = v R^T + t
        out.append(M)
    return np.float32(out)


def _run(L, fn, mesh, mats, first_gid=7, with_nt=True):
    pos, nrm, tan, ji, jw, idx = mesh
    op = np.zeros_like(pos); on = np.zeros(len(pos), np.uint32); ot = np.zeros(len(pos), np.uint32)
    rec = np.arange((first_gid + len(idx)) * 24, dtype=np.uint32).reshape(-1, 24).copy()           # recognisable filler: untouched words must survive
    f = getattr(L, fn); f.argtypes = [C.c_uint32] * 3 + [C.c_void_p] * 11
    p = lambda a: None if a is None else a.ctypes.data
    assert f(len(pos), len(idx), first_gid, p(pos), p(nrm if with_nt else None), p(tan if with_nt else None), p(ji), p(jw), p(np.ascontiguousarray(mats).reshape(-1, 16)), p(idx), p(op), p(on), p(ot), p(rec)) == 0
    return op, on, ot, rec


def test_oracle_skinning_properties(oracle):
    L = oracle.lib(); rng = np.random.default_rng(12)
    mesh = _mesh(rng); pos, nrm, tan, ji, jw, idx = mesh
    # identity joints: positions unchanged, normals / tangents survive their own quantisation, records point at the right vertices, everything else untouched
    I = np.tile(np.eye(4, dtype=np.float32), (6, 1, 1))
    op, on, ot, rec = _run(L, "oracle_skin", mesh, I)
    assert np.allclose(op, pos, atol=1e-6) and np.abs(_unpack3(on) - _unpack3(nrm)).max() < 0.02 and np.array_equal(ot >> 24, tan >> 24)
    filler = np.arange(rec.size, dtype=np.uint32).reshape(rec.shape)
    assert np.array_equal(rec[:7], filler[:7])                                                     # triangles before firstGid
    for k in range(3): assert np.array_equal(rec[7:, 4 * k:4 * k + 3].view(np.float32), op[idx[:, k]]) and np.array_equal(rec[7:, 4 * k + 3], on[idx[:, k]])
    assert np.array_equal(rec[7:, 18], ot[idx[:, 0]]) and np.array_equal(rec[7:, 19], ot[idx[:, 1]]) and np.array_equal(rec[7:, 20], ot[idx[:, 2]])
    assert np.array_equal(rec[7:, 12:18], filler[7:, 12:18]) and np.array_equal(rec[7:, 21:], filler[7:, 21:])          # uvs and ids are not the skin's business
    # one rigid joint for everyone: positions follow the transform exactly, normals rotate with it, lengths stay 1
    M = _mats(rng, 1)[0]; allone = (pos, nrm, tan, np.zeros_like(ji), np.tile(np.float32([1, 0, 0, 0]), (len(pos), 1)), idx)
    op, on, ot, _ = _run(L, "oracle_skin", allone, M[None])
    assert np.allclose(op, pos @ M[:3, :3] + M[3, :3], atol=1e-5)
    assert np.abs(_unpack3(on) - _unpack3(nrm) @ M[:3, :3]).max() < 0.03 and np.abs(np.linalg.norm(_unpack3(on), axis=1) - 1).max() < 0.03
    # blending is linear in the weights: the skinned position is the weighted mean of the per-joint positions
    mats = _mats(rng)
    op, _, _, _ = _run(L, "oracle_skin", mesh, mats)
    per_joint = np.einsum("vi,jik->vjk", np.concatenate([pos, np.ones((len(pos), 1), np.float32)], 1), mats)[..., :3]
    want = (per_joint[np.arange(len(pos))[:, None], ji] * jw[..., None]).sum(1)
    assert np.allclose(op, want, atol=1e-4)
    # no normals / tangents supplied: only positions are written
    op2, _, _, rec2 = _run(L, "oracle_skin", mesh, mats, with_nt=False)
    assert np.array_equal(op2, op) and np.array_equal(rec2[7:, 3], filler[7:, 3]) and np.array_equal(rec2[7:, 18:21], filler[7:, 18:21])


def test_product_bodies_equal_the_oracle(oracle):
    import host_build_lib as emu
    rng = np.random.default_rng(13)
    for nv, nt, with_nt in ((300, 500, True), (17, 40, False), (1, 1, True)):
        mesh = _mesh(rng, nv, nt); mats = _mats(rng)
        a = _run(oracle.lib(), "oracle_skin", mesh, mats, with_nt=with_nt); b = _run(emu.lib(), "emu_skin", mesh, mats, with_nt=with_nt)
        assert np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32)) and np.array_equal(a[3], b[3])
        if with_nt: assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
