"""The shrink-and-contain bound the NMS resolve kernel uses before it clips (csrc/lanms.cu,
iou_above_by_containment): a numpy restatement of the same arithmetic, checked against the oracle's float64
Sutherland-Hodgman IoU (lanms.py:80-91) on random pairs of regular quads.  The bound may answer "IoU > thr" only
when the oracle agrees; it is allowed to stay silent.  CPU only -- the device code itself is pinned by the
bit-exact LANMS parity tests in test_gpu_parity.py."""
import numpy as np
import pytest

from oracle import cpu


def _shoelace(q):  # (n, 4, 2) -> |area|, sequential sum like ms_shoelace
    acc = np.zeros(len(q))
    for i in range(4):
        j = (i + 1) % 4
        acc = acc + (q[:, i, 0] * q[:, j, 1] - q[:, j, 0] * q[:, i, 1])
    return np.abs(acc) / 2.0


def _shrunk_inside(a, b, lam):
    c = 0.25 * (a[:, 0] + a[:, 1] + a[:, 2] + a[:, 3])
    ok = np.ones(len(a), bool)
    for k in range(4):
        v = c + lam[:, None] * (a[:, k] - c)
        for e in range(4):
            e1 = (e + 1) % 4
            t1 = (b[:, e1, 0] - b[:, e, 0]) * (v[:, 1] - b[:, e, 1])
            t2 = (b[:, e1, 1] - b[:, e, 1]) * (v[:, 0] - b[:, e, 0])
            ok &= (t1 - t2) > 1e-9 * (np.abs(t1) + np.abs(t2))
    return ok


def containment_says_above(a, b, thr):
    a1, a2 = _shoelace(a), _shoelace(b)
    need = 1.06 * thr * (a1 + a2) / (1.0 + thr)
    la, lb = np.maximum(need / a1, 0.0025), np.maximum(need / a2, 0.0025)
    r = (la < 0.9) & _shrunk_inside(a, b, np.sqrt(la) * 1.0000001)
    r |= (lb < 0.9) & _shrunk_inside(b, a, np.sqrt(lb) * 1.0000001)
    return r & (a1 > 0) & (a2 > 0)


def _random_regular_pairs(rng, n):
    rect = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], float)  # positively oriented in the reference's convention
    w, h = rng.uniform(8, 300, (n, 1, 1)), rng.uniform(6, 80, (n, 1, 1))
    org = rng.uniform(0, 2000, (n, 1, 2))

    def quad(shift, scale):
        q = rect[None] * np.concatenate([w * scale, h * scale], axis=2) + org + shift
        q = q + rng.normal(0, 0.02, (n, 4, 1)) * np.concatenate([w, h], axis=2)  # not quite rectangles
        th = rng.uniform(-0.3, 0.3, n)
        ctr = q.mean(1, keepdims=True)
        rot = np.stack([np.stack([np.cos(th), -np.sin(th)], -1), np.stack([np.sin(th), np.cos(th)], -1)], -2)
        return np.einsum("nij,nkj->nki", rot, q - ctr) + ctr

    frac = rng.uniform(-1.0, 1.0, (n, 1, 2)) * rng.choice([0.02, 0.2, 0.6, 1.0], (n, 1, 1))
    shift = frac * np.concatenate([w, h], axis=2)
    return quad(0.0, 1.0), quad(shift, rng.uniform(0.5, 1.6, (n, 1, 1)))


@pytest.mark.parametrize("thr", [0.0, 0.05, 0.2, 0.5, 0.8, 0.95])
def test_bound_never_contradicts_the_reference_iou(thr):
    rng = np.random.default_rng(int(thr * 1000) + 1)
    a, b = _random_regular_pairs(rng, 4000)
    says = containment_says_above(a, b, thr)
    iou = np.array([cpu.polygon_iou(x, y) for x, y in zip(a, b)])
    assert not np.any(says & ~(iou > thr)), "the bound claimed IoU > thr where the reference's clip says otherwise"
    # it is useful, not just safe: most clearly overlapping pairs are decided without the clip
    clear = iou > min(0.97, thr + 0.4 * (1 - thr) + 0.1)
    if clear.sum() > 50:
        assert says[clear].mean() > 0.6
    # and it never fires on the margin it promises to leave (6 %)
    assert not np.any(says & (iou <= thr * 1.02))
