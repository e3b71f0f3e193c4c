"""Pins the TMA im2col addressing the conv kernels rely on (zero padding, stride, dilation, row/image wrap)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _unswizzle(raw, rows, row_bytes=128):
    """raw uint8 [rows*128] in SWIZZLE_128B layout -> logical [rows][128] bytes."""
    raw = raw.reshape(rows, row_bytes // 16, 16)
    out = np.empty_like(raw)
    for r in range(rows):
        for j in range(row_bytes // 16):
            out[r, j] = raw[r, j ^ (r % 8)]
    return out.reshape(rows, row_bytes)


def _expected(x, ppc, c0, stride, pad, dil, R, tap_r, tap_s, m0):
    n_, h_, w_, _ = x.shape
    ho = (h_ + 2 * pad - dil * (R - 1) - 1) // stride + 1
    wo = (w_ + 2 * pad - dil * (R - 1) - 1) // stride + 1
    exp = np.zeros((ppc, 64), dtype=np.float32)
    for i in range(ppc):
        m = m0 + i
        n = m // (ho * wo)
        p = (m % (ho * wo)) // wo
        q = m % wo
        ih = p * stride - pad + tap_r * dil
        iw = q * stride - pad + tap_s * dil
        if n < n_ and 0 <= ih < h_ and 0 <= iw < w_:
            exp[i] = x[n, ih, iw, c0:c0 + 64]
    return exp, ho, wo


@pytest.mark.parametrize("cfg", [
    # N, H, W, C, R, stride, pad, dil, ppc, m0
    (2, 9, 7, 64, 3, 1, 1, 1, 128, 0),
    (2, 9, 7, 128, 3, 1, 1, 1, 128, 0),
    (3, 9, 7, 64, 3, 1, 2, 2, 64, 64),
    (2, 13, 13, 64, 3, 2, 1, 1, 64, 0),
    (4, 9, 9, 64, 1, 2, 0, 1, 64, 0),
    (2, 11, 11, 64, 3, 1, 6, 6, 128, 128),
    (1, 5, 5, 64, 1, 1, 0, 1, 64, 0),   # < 128 KiB tensor, M tail beyond the batch
])
def test_im2col_tile(cfg):
    from zs3_b200 import kernels as K
    N, H, W, Cc, R, stride, pad, dil, ppc, m0 = cfg
    g = torch.Generator().manual_seed(7)
    x = torch.randn(N, H, W, Cc, generator=g).to(torch.bfloat16)
    xd = x.cuda()
    xn = x.float().numpy()
    upper = pad - (R - 1) * dil
    bad = []
    for tap_r in range(R):
        for tap_s in range(R):
            for c0 in range(0, Cc, 64):
                exp, ho, wo = _expected(xn, ppc, c0, stride, pad, dil, R, tap_r, tap_s, m0)
                n = m0 // (ho * wo)
                p = (m0 % (ho * wo)) // wo
                q = m0 % wo
                raw = K.im2col_probe(xd, pad, upper, stride, 64, ppc, c0, q * stride - pad, p * stride - pad, n,
                                     tap_s * dil, tap_r * dil)
                torch.cuda.synchronize()
                got = _unswizzle(raw.cpu().numpy(), ppc).copy().view(np.uint16).astype(np.uint32) << 16
                got = got.view(np.float32).reshape(ppc, 64)
                if not np.array_equal(got, exp):
                    nbad = int((got != exp).any(axis=1).sum())
                    bad.append((tap_r, tap_s, c0, nbad))
                    if len(bad) == 1:
                        # diagnose: where did each of the first rows come from?
                        flat = xn[..., c0:c0 + 64].reshape(-1, 64)
                        for i in range(min(ppc, 24)):
                            hit = np.where((flat == got[i]).all(axis=1))[0]
                            print(f"row {i}: src pixel {hit[:3]} zero={not got[i].any()} expect_zero={not exp[i].any()}")
    assert not bad, f"im2col mismatch (tap_r, tap_s, c0, bad_rows): {bad[:8]}"
