"""TPS rectification of a recogniser batch on the B200 (ms_tps_rectify).

An EXTENSION named by BASELINE.json's north_star; the reference has no transformation stage, so there is no
reference behaviour to match: the operation is the TPS-STN of the TRBA literature (Baek et al. 2019) -- GridGenerator
followed by F.grid_sample(padding_mode="border", align_corners=True) -- and is checked against
torch.nn.functional.grid_sample in tests/test_gpu_extensions.py.

`TPSGrid` holds what depends only on (F, out_h, out_w): inv_delta_C and P_hat, computed once in float64 and stored as
float32 on the device like the registered buffers of the public implementation.  The localisation network that predicts
the fiducial points C' is a network and stays outside this package.
"""
import ctypes as C

import numpy as np

from ._cabi import Context, check


def build_tps_constants(n_fid, out_h, out_w, eps=1e-6):
    """(C (F,2), inv_delta_C (F+3,F+3), P_hat (n,F+3)) in float64: fiducial base points on the rectified image (half on
    the top edge, half on the bottom edge), the inverse of the TPS system matrix, and [1, P, rbf(P, C)] for the output
    pixel centres P."""
    F = int(n_fid)
    if F < 2 or F % 2:
        raise ValueError("the number of fiducial points must be even")
    xs = np.linspace(-1.0, 1.0, F // 2)
    Cb = np.concatenate([np.stack([xs, -np.ones(F // 2)], 1), np.stack([xs, np.ones(F // 2)], 1)], 0)
    hat = np.linalg.norm(Cb[:, None, :] - Cb[None, :, :], axis=2)
    np.fill_diagonal(hat, 1.0)
    hat = (hat ** 2) * np.log(hat)
    delta = np.concatenate([
        np.concatenate([np.ones((F, 1)), Cb, hat], axis=1),
        np.concatenate([np.zeros((2, 3)), Cb.T], axis=1),
        np.concatenate([np.zeros((1, 3)), np.ones((1, F))], axis=1)], axis=0)
    inv_delta = np.linalg.inv(delta)
    gx = (np.arange(-out_w, out_w, 2) + 1.0) / out_w
    gy = (np.arange(-out_h, out_h, 2) + 1.0) / out_h
    P = np.stack(np.meshgrid(gx, gy), axis=2).reshape(-1, 2)  # row-major over (y, x)
    r = np.linalg.norm(P[:, None, :] - Cb[None, :, :], axis=2)
    rbf = np.square(r) * np.log(r + eps)
    P_hat = np.concatenate([np.ones((len(P), 1)), P, rbf], axis=1)
    return Cb, inv_delta, P_hat


class TPSGrid:
    def __init__(self, n_fid=20, out_hw=(32, 100), device=0):
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError("manuscript_b200.TPSGrid needs a CUDA device: there is no CPU fallback")
        self.torch = torch
        self.device = torch.device("cuda", int(device)) if not isinstance(device, torch.device) else device
        self.n_fid, (self.out_h, self.out_w) = int(n_fid), (int(out_hw[0]), int(out_hw[1]))
        Cb, inv_delta, P_hat = build_tps_constants(n_fid, self.out_h, self.out_w)
        self.base_points = Cb.astype(np.float32)
        self.inv_delta_c = torch.from_numpy(inv_delta.astype(np.float32)).to(self.device)
        self.p_hat = torch.from_numpy(P_hat.astype(np.float32)).to(self.device)            # (n, F+3), as the module keeps it
        self.p_hat_t = self.p_hat.t().contiguous()                                         # (F+3, n) for the kernel
        self.ctx = Context(self.device.index or 0)

    def rectify(self, batch, c_prime, out=None):
        """batch (B,C,H,W) f32 CUDA, c_prime (B,F,2) f32 CUDA -> (B,C,out_h,out_w) f32 CUDA."""
        torch = self.torch
        assert batch.is_cuda and c_prime.is_cuda and batch.dtype == torch.float32 and c_prime.dtype == torch.float32
        B, Cn, H, W = batch.shape
        assert tuple(c_prime.shape) == (B, self.n_fid, 2), c_prime.shape
        batch, c_prime = batch.contiguous(), c_prime.contiguous()
        if out is None:
            out = torch.empty((B, Cn, self.out_h, self.out_w), dtype=torch.float32, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        for b0 in range(0, B, 65535):
            nb = min(65535, B - b0)
            with torch.cuda.device(self.device):
                check(self.ctx.lib.ms_tps_rectify(self.ctx.handle, batch[b0:].data_ptr(), c_prime[b0:].data_ptr(),
                                                  self.inv_delta_c.data_ptr(), self.p_hat_t.data_ptr(), nb, self.n_fid, Cn,
                                                  H, W, self.out_h, self.out_w, out[b0:].data_ptr(), C.c_void_p(stream)))
        return out
