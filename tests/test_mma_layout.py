"""Host-side check of the fragment bookkeeping of gemm_mma_kernel (tacotron_b200/csrc/train.cu): the shared-memory
indices each lane reads for its A/B fragments and the (row, column) each lane writes its accumulators to are replayed
here in numpy against the PTX m16n8k8 .tf32 fragment definition, for all 8 warps of a 64x64x16 tile.  It pins the
index algebra (which is what a blind kernel gets wrong), not the hardware instruction -- that one is already exercised
by the forward decoder kernel and by tests/test_gpu_z_gemm_mma.py on a GPU."""
import numpy as np

GBM = GBN = 64
GBK = 16


def _mma_m16n8k8(a_regs, b_regs):
    """PTX semantics: per-lane registers -> per-lane D registers.  a_regs [32][4], b_regs [32][2] -> d [32][4]."""
    A = np.zeros((16, 8)); Bm = np.zeros((8, 8))
    for lane in range(32):
        g, tg = lane >> 2, lane & 3
        A[g, tg], A[g + 8, tg], A[g, tg + 4], A[g + 8, tg + 4] = a_regs[lane]
        Bm[tg, g], Bm[tg + 4, g] = b_regs[lane]
    D = A @ Bm
    d = np.zeros((32, 4))
    for lane in range(32):
        g, tg = lane >> 2, lane & 3
        d[lane] = [D[g, 2 * tg], D[g, 2 * tg + 1], D[g + 8, 2 * tg], D[g + 8, 2 * tg + 1]]
    return d


def test_fragment_indices_reconstruct_the_tile_product():
    rng = np.random.default_rng(0)
    As = rng.standard_normal((GBK, GBM))            # As[k][m]  (k-major, as the loaders store it)
    Bs = rng.standard_normal((GBK, GBN))            # Bs[k][n]
    C = np.full((GBM, GBN), np.nan)
    for warp in range(8):
        wm, wn = (warp & 3) * 16, (warp >> 2) * 32
        acc = np.zeros((4, 32, 4))                  # [nt][lane][e]
        for ks in (0, 8):
            a = np.zeros((32, 4))
            for lane in range(32):
                g, tg = lane >> 2, lane & 3
                a[lane] = [As[ks + tg, wm + g], As[ks + tg, wm + g + 8], As[ks + tg + 4, wm + g], As[ks + tg + 4, wm + g + 8]]
            for nt in range(4):
                b = np.zeros((32, 2))
                for lane in range(32):
                    g, tg = lane >> 2, lane & 3
                    b[lane] = [Bs[ks + tg, wn + nt * 8 + g], Bs[ks + tg + 4, wn + nt * 8 + g]]
                acc[nt] += _mma_m16n8k8(a, b)
        for nt in range(4):
            for lane in range(32):
                g, tg = lane >> 2, lane & 3
                for e in range(4):
                    m = wm + g + (8 if (e >> 1) else 0)
                    n = wn + nt * 8 + 2 * tg + (e & 1)
                    assert np.isnan(C[m, n]), "two lanes write the same output"
                    C[m, n] = acc[nt, lane, e]
    assert not np.isnan(C).any(), "some outputs are never written"
    assert np.allclose(C, As.T @ Bs, atol=1e-12)


def test_smem_row_stride_is_conflict_free():
    """bank of As[ks+tg][wm+g] with a 72-float row: 8*tg + g (mod 32) is a permutation of 0..31 over the warp"""
    banks = {((tg * 72) + g) % 32 for tg in range(4) for g in range(8)}
    assert len(banks) == 32
    banks68 = {((tg * 68) + g) % 32 for tg in range(4) for g in range(8)}
    assert len(banks68) < 32                        # the FFMA kernel's 68-float row would conflict here


def test_3xtf32_split_error():
    """hi = top 11 significant bits, lo = x - hi (exact); the three products kept reproduce x*y to ~2^-21"""
    rng = np.random.default_rng(1)
    x = rng.standard_normal(10000).astype(np.float32)
    y = rng.standard_normal(10000).astype(np.float32)

    def split(v):
        hi = (v.view(np.uint32) & np.uint32(0xffffe000)).view(np.float32)
        lo = (v - hi).astype(np.float32)
        lo_t = (lo.view(np.uint32) & np.uint32(0xffffe000)).view(np.float32)     # the tensor core reads lo as TF32 too
        return hi.astype(np.float64), lo_t.astype(np.float64)
    xh, xl = split(x); yh, yl = split(y)
    approx = xl * yh + xh * yl + xh * yh
    exact = x.astype(np.float64) * y.astype(np.float64)
    rel = np.abs(approx - exact) / np.maximum(np.abs(exact), 1e-30)
    assert rel.max() < 4e-6 and np.median(rel) < 5e-7
