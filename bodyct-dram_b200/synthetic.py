"""Synthetic inputs of SURVEY §8d for the entry points, the benchmark and smoke runs (there is no dataset on this path):
lobe-chunk training batches (`synthetic_loader`, re-exported from train.py) and a full CT scan with a 5-lobe label volume."""
import numpy as np

from train import synthetic_loader  # noqa: F401


def synthetic_scan(shape=(400, 512, 512), spacing=(1.0, 0.7, 0.7), seed=0):
    """int16 HU volume [D,H,W]: air -1000, body ellipse +40, two lungs -850 + N(0,50) with -300 blobs; uint8 lobe labels
    1..5 (3 right, 2 left, split by axial planes); lesion mask; spacing (z,y,x) in mm.  Deterministic in `seed`."""
    g = np.random.RandomState(seed)
    D, H, W = shape
    zz = np.arange(D, dtype=np.float32)[:, None, None]
    yy = np.arange(H, dtype=np.float32)[None, :, None]
    xx = np.arange(W, dtype=np.float32)[None, None, :]
    scan = np.full(shape, -1000, dtype=np.int16)
    body = ((yy - H / 2) / (0.42 * H)) ** 2 + ((xx - W / 2) / (0.46 * W)) ** 2 <= 1.0
    scan[np.broadcast_to(body, shape)] = 40
    lobe = np.zeros(shape, dtype=np.uint8)
    for side, cx in ((0, 0.30 * W), (1, 0.70 * W)):
        lung = (((zz - D / 2) / (0.40 * D)) ** 2 + ((yy - H / 2) / (0.28 * H)) ** 2 + ((xx - cx) / (0.15 * W)) ** 2) <= 1.0
        if side == 0:
            lab = np.where(zz < 0.38 * D, 1, np.where(zz < 0.6 * D, 2, 3)).astype(np.uint8)
        else:
            lab = np.where(zz < 0.5 * D, 4, 5).astype(np.uint8)
        lobe = np.where(lung, np.broadcast_to(lab, shape), lobe)
    lungs = lobe > 0
    noise = (g.randn(*shape).astype(np.float32) * 50.0 - 850.0)
    scan = np.where(lungs, noise, scan).astype(np.int16)
    lesion = np.zeros(shape, dtype=bool)
    nz = np.argwhere(lungs)
    for _ in range(12):
        c = nz[g.randint(len(nz))]
        r = g.uniform(6, 18)
        blob = ((zz - c[0]) ** 2 + (yy - c[1]) ** 2 + (xx - c[2]) ** 2) <= r * r
        lesion |= blob & lungs
    scan = np.where(lesion, np.int16(-300), scan).astype(np.int16)
    return scan, lobe, lesion.astype(np.uint8), np.asarray(spacing, dtype=np.float64)
