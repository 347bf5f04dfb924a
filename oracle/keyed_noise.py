"""
TEST INFRASTRUCTURE -- numpy restatement of libvqvs' keyed Gaussian noise (vqvs_keyed_normal, csrc/simt_ops.cu).

The reference draws step noise with `torch.randn_like` (diffusion/diffusion.py:62-63); batch-sharded sampling needs noise
that depends only on (seed, GLOBAL sample index, step), which is this repository's own addition, so the
"reference" here is the published Philox4x32-10 algorithm (Salmon et al., SC'11; the same generator behind
torch.cuda's and cuRAND's Philox engines) followed by Box-Muller.  Integer stage: bit-exact against the kernel.  Float
stage (log / sincos): float32 libm vs CUDA intrinsic-free sincosf/logf, compared at 1e-5.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3))
    k0, k1 = np.uint32(k0), np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            h0, l0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
            h1, l1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
            c0, c1, c2, c3 = h1 ^ c1 ^ k0, l1, h0 ^ c3 ^ k1, l0
            k0, k1 = np.uint32(k0 + W0), np.uint32(k1 + W1)
    return c0, c1, c2, c3


def keyed_words(seed: int, row: int, step: int, length: int):
    """The four Philox words of every quad of one row: uint32 [quads, 4]."""
    quads = (length + 3) // 4
    q = np.arange(quads, dtype=np.uint64)
    st = np.uint32(step & 0xFFFFFFFF)
    with np.errstate(over="ignore"):
        c1 = np.uint32((row >> 32) & 0xFFFFFFFF) ^ np.uint32(st * np.uint32(0x85EBCA6B))
        c3 = (q >> np.uint64(32)).astype(np.uint32) + st
    w = philox4x32_10(np.full(quads, row & 0xFFFFFFFF, dtype=np.uint32), np.full(quads, c1, dtype=np.uint32),
                      q.astype(np.uint32), c3, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return np.stack(w, axis=1)


def keyed_normal(seed: int, rows, step: int, length: int) -> np.ndarray:
    """float32 [len(rows), length]; row r depends only on (seed, rows[r], step)."""
    out = np.empty((len(rows), length), dtype=np.float32)
    for i, row in enumerate(rows):
        w = keyed_words(seed, int(row), step, length)
        u = ((w >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)
        z = np.empty((w.shape[0], 4), dtype=np.float32)
        for p in range(2):
            rad = np.sqrt(np.float32(-2.0) * np.log(u[:, 2 * p]))
            ang = np.float32(6.283185307179586) * u[:, 2 * p + 1]
            z[:, 2 * p] = rad * np.cos(ang)
            z[:, 2 * p + 1] = rad * np.sin(ang)
        out[i] = z.reshape(-1)[:length]
    return out
