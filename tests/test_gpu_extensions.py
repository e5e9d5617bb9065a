"""The two operations BASELINE.json's north_star names that the reference does not contain: RBOX geometry
decode and the TPS rectification grid_sample.  There is no reference behaviour to be identical to ("parity unpinned"):
RBOX is checked against the oracle's closed form AND an independent numpy statement of the published algorithm, within
north_star's coordinate tolerance; TPS against torch.nn.functional.grid_sample, within a float32 tolerance stated here."""
import numpy as np
import pytest

from oracle import cpu

pytestmark = pytest.mark.gpu

COORD_RTOL = 1e-4  # north_star: "1e-4 relative on coordinates"


def numpy_restore_rbox(origin, geometry):
    """restore_rectangle_rbox of the public EAST implementations, for points `origin` (n,2) in image pixels and
    geometry (n,5) = top, right, bottom, left distances (image pixels) + angle."""
    out = np.zeros((len(origin), 4, 2))
    for i, (o, g) in enumerate(zip(origin, geometry)):
        d, a = g[:4], g[4]
        if a >= 0:
            p = np.array([[0, -d[0] - d[2]], [d[1] + d[3], -d[0] - d[2]], [d[1] + d[3], 0], [0, 0], [d[3], -d[2]]])
            rx, ry = np.array([np.cos(a), np.sin(a)]), np.array([-np.sin(a), np.cos(a)])
        else:
            p = np.array([[-d[1] - d[3], -d[0] - d[2]], [0, -d[0] - d[2]], [0, 0], [-d[1] - d[3], 0], [-d[1], -d[2]]])
            rx, ry = np.array([np.cos(-a), -np.sin(-a)]), np.array([np.sin(-a), np.cos(-a)])
        pr = np.stack([p @ rx, p @ ry], axis=1)
        out[i] = pr[:4] + (o - pr[4])
    return out


@pytest.mark.parametrize("q", [1, 2])
def test_rbox_decode(q):
    import manuscript_b200 as mb

    rng = np.random.default_rng(3 + q)
    H, W = 96, 128
    score = rng.uniform(0, 0.55, (H, W)).astype(np.float32)
    hot = rng.random((H, W)) < 0.08
    score[hot] = rng.uniform(0.61, 0.99, hot.sum()).astype(np.float32)
    geo = np.concatenate([rng.uniform(0.5, 12.0, (4, H, W)), rng.uniform(-0.6, 0.6, (1, H, W))]).astype(np.float32)
    geo[4][rng.random((H, W)) < 0.1] = 0.0
    got = mb.decode_rbox_from_maps(score, geo, 0.6, 4.0, q)
    want = cpu.decode_rbox_from_maps(score, geo, 0.6, 4.0, q)
    assert got.shape == want.shape and len(want) > 300
    np.testing.assert_array_equal(got[:, 8], want[:, 8])  # same pixels, same order, same scores
    np.testing.assert_allclose(got[:, :8], want[:, :8], rtol=COORD_RTOL, atol=1e-3)
    # (H, W, 5) layout accepted like the QUAD decode accepts (H, W, 8)
    np.testing.assert_array_equal(mb.decode_rbox_from_maps(score, geo.transpose(1, 2, 0), 0.6, 4.0, q), got)
    # the oracle's closed form against the published algorithm, on the same candidate pixels
    ys, xs = np.nonzero(score > np.float32(0.6))
    if q > 1:
        cells = np.unique(np.stack([ys // q * q + q // 2, xs // q * q + q // 2], 1), axis=0)
        ys, xs = cells[:, 0], cells[:, 1]
    g = geo[:, ys, xs].T.astype(np.float64)
    g[:, :4] *= 4.0
    ref = numpy_restore_rbox(np.stack([xs * 4.0, ys * 4.0], 1), g).reshape(-1, 8)
    np.testing.assert_allclose(want[:, :8], ref, rtol=1e-6, atol=1e-4)
    # the decoded rectangles go through the same NMS as QUAD candidates
    kept = mb.locality_aware_nms(got, 0.2)
    np.testing.assert_array_equal(kept, cpu.locality_aware_nms(got, 0.2))
    # odd map with q = 2: the IndexError of the QUAD decode
    bad = np.zeros((7, 8), np.float32)
    bad[6, 0] = 0.9
    with pytest.raises(IndexError):
        mb.decode_rbox_from_maps(bad, np.zeros((5, 7, 8), np.float32), 0.6, 4.0, 2)
    assert mb.decode_rbox_from_maps(np.zeros((8, 8), np.float32), np.zeros((5, 8, 8), np.float32), 0.6, 4.0, 2).shape == (0, 9)


@pytest.mark.parametrize("n_fid,out_hw,in_hw", [(20, (32, 100), (32, 128)), (20, (32, 128), (32, 128)), (10, (24, 60), (48, 90))])
def test_tps_rectify_matches_torch_grid_sample(n_fid, out_hw, in_hw):
    """ms_tps_rectify against the TPS-STN written in torch (GridGenerator matmuls + F.grid_sample, padding_mode="border",
    align_corners=True).  Tolerances: (i) against the same computation in float64 (grid and interpolation), on white
    noise in [-1, 1]: the kernel accumulates the grid in float64 and interpolates in float32 (pixel coordinates up to 127 carry
    8e-6 of rounding, times a value difference of up to 2) -> <= 5e-5; (ii) against
    torch's own float32 pipeline the sampling POSITIONS differ by float32 rounding of sums whose terms are ~10 (about
    1e-5 of the image, 1e-3 pixel), so on a smooth image (|gradient| <= 0.05 / pixel) the samples agree to <= 1e-4."""
    import torch
    import torch.nn.functional as F

    import manuscript_b200 as mb

    torch.manual_seed(n_fid)
    B, C = 37, 3
    tps = mb.TPSGrid(n_fid=n_fid, out_hw=out_hw)
    noise = torch.rand((B, C, *in_hw), device="cuda") * 2 - 1
    yy, xx = torch.meshgrid(torch.arange(in_hw[0], device="cuda"), torch.arange(in_hw[1], device="cuda"), indexing="ij")
    smooth = (torch.sin(xx * 0.03 + torch.arange(B, device="cuda")[:, None, None, None] * 0.1) *
              torch.cos(yy * 0.04)).expand(B, C, *in_hw).contiguous().float()
    base = torch.from_numpy(tps.base_points).cuda()
    c_prime = (base[None] + 0.15 * torch.randn((B, n_fid, 2), device="cuda")).contiguous()
    c_prime[0] = base                      # identity warp
    c_prime[1] = base * 1.4                # reaches outside the image: border padding
    zeros = torch.zeros((B, 3, 2), device="cuda")

    def torch_tps(img, dtype):
        T = torch.matmul(tps.inv_delta_c.to(dtype)[None], torch.cat([c_prime, zeros], dim=1).to(dtype))
        grid = torch.matmul(tps.p_hat.to(dtype)[None], T).reshape(B, out_hw[0], out_hw[1], 2)
        return F.grid_sample(img.to(dtype), grid, mode="bilinear", padding_mode="border", align_corners=True)

    got = tps.rectify(noise, c_prime)
    assert tuple(got.shape) == (B, C, *out_hw)
    err64 = (got.double() - torch_tps(noise, torch.float64)).abs().max().item()
    assert err64 <= 5e-5, err64
    err32 = (tps.rectify(smooth, c_prime) - torch_tps(smooth, torch.float32)).abs().max().item()
    assert err32 <= 1e-4, err32
