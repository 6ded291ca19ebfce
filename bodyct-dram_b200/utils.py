"""Host-side helpers the job runners and entry points need, mirroring the names in the reference's `dram/utils.py`
(Settings :42-69, get_callable_by_name :280-283, windowing :189-198, find_crops :244-254, binary_cam :226-242,
expand_dims/squeeze_dims :127-140, write_array_to_mha_itk :142-159, IOU/Dice :437-446, AverageMeter :98-114).  Pure numpy/Python — no SimpleITK,
skimage or OpenCV; the GPU versions of windowing / resampling / Otsu live in libdram_b200 (dram_native.ops)."""
import importlib
import importlib.util
import math
import os
import zlib

import numpy as np


class Settings:
    """Execute a settings .py file; every UPPERCASE name becomes an attribute (same contract as the reference)."""

    def __init__(self, settings_module_path, settings_name="settings"):
        self.settings_module_path = settings_module_path
        spec = importlib.util.spec_from_file_location(settings_name, settings_module_path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        self._explicit_settings = set()
        for name in dir(mod):
            if name.isupper():
                setattr(self, name, getattr(mod, name))
                self._explicit_settings.add(name)

    def is_overridden(self, setting):
        return setting in self._explicit_settings

    def __str__(self):
        return "\n".join(f"{k} = {getattr(self, k)!r}" for k in sorted(self._explicit_settings))


def get_callable_by_name(dotted):
    module_name, _, attr = dotted.rpartition('.')
    return getattr(importlib.import_module(module_name), attr)


class AverageMeter:
    def __init__(self):
        self.reset()

    def reset(self):
        self.val = self.avg = self.sum = 0.0
        self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / max(self.count, 1)


def expand_dims(t, expected_dim):
    while t.dim() < expected_dim:
        t = t.unsqueeze(0)
    return t


def squeeze_dims(t, expected_dim, squeeze_start_index=0):
    while t.dim() > expected_dim:
        t = t.squeeze(squeeze_start_index)
    return t


def windowing(image, from_span=(-1150, 350), to_span=(0, 255)):
    lo, hi = (np.min(image), np.max(image)) if from_span is None else from_span
    image = np.clip(image, a_min=lo, a_max=hi)
    return ((image - lo) / float(hi - lo)) * (to_span[1] - to_span[0]) + to_span[0]


def find_crops(mask, spacing, border):
    """Bounding box of mask > 0, padded by ceil(border / spacing) voxels per axis and clipped to the volume."""
    nz = np.nonzero(np.asarray(mask) > 0)
    out = []
    for ax, (size, sp) in enumerate(zip(mask.shape, spacing)):
        pad = int(math.ceil(border / sp)) if border > 0 else 0
        out.append(slice(max(0, int(nz[ax].min()) - pad), min(size, int(nz[ax].max()) + 1 + pad)))
    return tuple(out)


def otsu_threshold_from_histogram(hist_by_value):
    """Otsu threshold of uint8 data given its 256-entry value histogram (what skimage computes for integer images:
    one bin per integer between min and max, first maximum of the between-class variance, returns the bin centre)."""
    hist_by_value = np.asarray(hist_by_value, dtype=np.float64)
    nz = np.nonzero(hist_by_value)[0]
    lo, hi = int(nz[0]), int(nz[-1])
    hist = hist_by_value[lo:hi + 1]
    centers = np.arange(lo, hi + 1, dtype=np.float64)
    w1 = np.cumsum(hist)
    w2 = np.cumsum(hist[::-1])[::-1]
    m1 = np.cumsum(hist * centers) / np.maximum(w1, 1e-300)
    m2 = (np.cumsum((hist * centers)[::-1]) / np.maximum(w2[::-1], 1e-300))[::-1]
    return centers[:-1][int(np.argmax(w1[:-1] * w2[1:] * (m1[:-1] - m2[1:]) ** 2))]


def binary_cam(cam_probs, scaler=1.0, from_span=(0, 1)):
    """(mask, threshold in [0,1]) — window to uint8, Otsu on the 8-bit values (utils.py:226-242)."""
    cam = np.asarray(cam_probs.detach().cpu().numpy() if hasattr(cam_probs, "detach") else cam_probs)
    if cam.size == 0:
        raise ValueError("empty array encountered! cam_probs.size == 0.")
    w = windowing(cam, from_span=from_span).astype(np.uint8)
    hist = np.bincount(w.ravel(), minlength=256)
    if np.count_nonzero(hist) < 2:
        return np.ones_like(w).astype(bool), float(np.nonzero(hist)[0][0]) / 255.0
    th = min(otsu_threshold_from_histogram(hist) * scaler, 255.0)
    return w >= th, th / 255.0


def IOU(predict, target, smooth):
    inter = np.sum(np.logical_and(predict, target))
    return (inter + smooth) / (np.sum(np.logical_or(predict, target)) + smooth)


def Dice(predict, target, smooth):
    inter = np.sum(np.logical_and(predict, target))
    return (2. * inter + smooth) / (predict.sum() + target.sum() + smooth)


# ------------------------------------------------------------------------------------------------ MetaImage (.mha) I/O
# The reference writes its outputs with SimpleITK's ImageFileWriter (utils.py:142-159, compressed .mha) and reads scans /
# lobe masks through SimpleITK.  SimpleITK is not a dependency here: MetaImage is a text header followed by (zlib-
# compressed) little-endian voxels in x-fastest order, which numpy + zlib cover.  Arrays are z-y-x; spacing / origin are in
# ITK's x-y-z order exactly as the reference hands them to SetSpacing / SetOrigin (callers reverse them,
# job_runner.py:873-890); direction is the row-major 3x3 matrix.
_MET_TYPES = {"MET_UCHAR": np.uint8, "MET_CHAR": np.int8, "MET_USHORT": np.uint16, "MET_SHORT": np.int16,
              "MET_UINT": np.uint32, "MET_INT": np.int32, "MET_FLOAT": np.float32, "MET_DOUBLE": np.float64}
_MET_NAMES = {np.dtype(v): k for k, v in _MET_TYPES.items()}


def write_mha(path, arr, spacing=(1.0, 1.0, 1.0), origin=(0.0, 0.0, 0.0), direction=None, compress=True,
              orientation="RAI"):
    arr = np.ascontiguousarray(arr)
    if arr.ndim != 3 or arr.dtype not in _MET_NAMES:
        raise ValueError(f"write_mha: need a 3-d array of a MetaImage element type, got {arr.dtype} {arr.shape}")
    direction = np.eye(3).flatten().tolist() if direction is None else list(np.asarray(direction, np.float64).flatten())
    data = arr.astype(arr.dtype.newbyteorder("<"), copy=False).tobytes()
    if compress:
        data = zlib.compress(data, 2)
    fmt = lambda xs: " ".join(repr(float(x)) if float(x) != int(float(x)) else str(int(float(x))) for x in xs)
    header = ["ObjectType = Image", "NDims = 3", "BinaryData = True", "BinaryDataByteOrderMSB = False",
              f"CompressedData = {'True' if compress else 'False'}"]
    if compress:
        header.append(f"CompressedDataSize = {len(data)}")
    header += [f"TransformMatrix = {fmt(direction)}", f"Offset = {fmt(origin)}", "CenterOfRotation = 0 0 0",
               f"AnatomicalOrientation = {orientation}", f"ElementSpacing = {fmt(spacing)}",
               f"DimSize = {arr.shape[2]} {arr.shape[1]} {arr.shape[0]}", f"ElementType = {_MET_NAMES[arr.dtype]}",
               "ElementDataFile = LOCAL"]
    with open(path, "wb") as f:
        f.write(("\n".join(header) + "\n").encode("ascii"))
        f.write(data)


def read_mha(path):
    """-> (array z-y-x, {"spacing", "origin", "direction"} in ITK x-y-z order)"""
    with open(path, "rb") as f:
        raw = f.read()
    meta, pos = {}, 0
    while True:
        end = raw.index(b"\n", pos)
        key, _, val = raw[pos:end].decode("ascii", "replace").partition("=")
        meta[key.strip()] = val.strip()
        pos = end + 1
        if key.strip() == "ElementDataFile":
            break
    if meta.get("ElementDataFile") != "LOCAL" or int(meta.get("NDims", 3)) != 3:
        raise ValueError(f"read_mha: only 3-d images with local data are supported ({path})")
    if meta.get("BinaryDataByteOrderMSB", "False") == "True" or meta.get("ElementByteOrderMSB", "False") == "True":
        raise ValueError("read_mha: big-endian MetaImages are not supported")
    dtype = _MET_TYPES[meta["ElementType"]]
    nx, ny, nz = (int(v) for v in meta["DimSize"].split())
    data = raw[pos:]
    if meta.get("CompressedData", "False") == "True":
        data = zlib.decompress(data)
    arr = np.frombuffer(data, dtype=np.dtype(dtype).newbyteorder("<"), count=nx * ny * nz).reshape(nz, ny, nx).astype(dtype)
    floats = lambda k, d: [float(v) for v in meta.get(k, d).split()]
    return arr, {"spacing": floats("ElementSpacing", "1 1 1"), "origin": floats("Offset", meta.get("Position", "0 0 0")),
                 "direction": floats("TransformMatrix", "1 0 0 0 1 0 0 0 1")}


def write_array_to_mha_itk(target_path, arrs, names, type=np.int16, origin=[0.0, 0.0, 0.0],
                           direction=np.eye(3, dtype=np.float64).flatten().tolist(), spacing=[1.0, 1.0, 1.0],
                           orientation='RAI'):
    """Same signature and files as utils.py:142-159: `<target_path>/<name>.mha`, compressed, one per array (z-y-x)."""
    for arr, name in zip(arrs, names):
        write_mha(os.path.join(target_path, '{}.mha'.format(name)), np.asarray(arr).astype(type), spacing, origin, direction,
                  compress=True, orientation=orientation)
