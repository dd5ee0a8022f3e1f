"""ctypes binding of liborbfe.so (include/orbfe.h) + Python mirrors of the reference's C++ front-end
interface, used by tests/ and bench.py.

The reference is C++ (its drop-in shim is include/orbfe_shim.hpp); this module mirrors the same
call surface for Python callers:

    ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST).Compute(image, mask)
                                          -> src/orb_features/orb_extractor.h:25-93
    ComputeStereoMatches(left, right, ...) -> Frame::ComputeStereoMatches, src/data/frame.cpp:406-577
    OrbMatcher(nnratio, checkOri).Search*  -> src/orb_features/orb_matcher.h:14-119

Everything computes on the GPU through the C ABI.  There is NO CPU fallback: if the CUDA library
cannot be loaded, or no CUDA device is visible, calls raise OrbfeError.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28  # cv::KeyPoint

TH_LOW, TH_HIGH, HISTO_LENGTH = 50, 100, 30  # orb_matcher.cpp:5-7


class OrbfeError(RuntimeError):
    pass


class _Params(C.Structure):
    _fields_ = [("nfeatures", C.c_int32), ("scale_factor", C.c_float), ("nlevels", C.c_int32),
                ("ini_th_fast", C.c_int32), ("min_th_fast", C.c_int32), ("max_width", C.c_int32),
                ("max_height", C.c_int32), ("max_images", C.c_int32)]


EXPORTS = [
    "orbfe_last_error", "orbfe_version", "orbfe_device_count", "orbfe_extractor_create", "orbfe_extractor_destroy",
    "orbfe_extractor_tables", "orbfe_extractor_max_keypoints", "orbfe_extract", "orbfe_extract_batch",
    "orbfe_pyramid_level", "orbfe_upload", "orbfe_upload_color", "orbfe_run", "orbfe_run_stereo", "orbfe_download", "orbfe_download_async",
    "orbfe_sync", "orbfe_pinned_alloc", "orbfe_pinned_free",
    "orbfe_event_record", "orbfe_event_elapsed_ms", "orbfe_set_stage_timing", "orbfe_stage_summary",
    "orbfe_launch_count",
    "orbfe_debug_candidates", "orbfe_debug_level_keypoints", "orbfe_debug_blurred", "orbfe_stereo_match",
    "orbfe_descriptor_distance", "orbfe_frame_create", "orbfe_frame_destroy", "orbfe_frame_pool_trim", "orbfe_features_in_area",
    "orbfe_search_for_initialization", "orbfe_search_by_projection_mappoints",
    "orbfe_search_by_projection_lastframe", "orbfe_search_by_bow",
    "orbfe_search_by_projection_sim3", "orbfe_search_by_projection_keyframe", "orbfe_fuse", "orbfe_fuse_sim3",
    "orbfe_search_by_sim3", "orbfe_search_by_bow_keyframes", "orbfe_search_for_triangulation",
    "orbfe_vocabulary_create", "orbfe_vocabulary_load_text", "orbfe_vocabulary_destroy", "orbfe_vocabulary_info",
    "orbfe_bow_transform", "orbfe_undistort_keypoints", "orbfe_is_in_frustum", "orbfe_debug_logf", "orbfe_search_local_points",
    "orbfe_frame_from_extractor", "orbfe_frame_refresh_from_extractor", "orbfe_frame_num_keypoints",
]

_libs = {}
vp = C.c_void_p


def load(path=None, _test_emulation=False):
    """Loads the C-ABI library.  `path=None` = the in-tree product build (compiled on demand with
    nvcc).  The emulated test build under tests/emu is refused unless a test asks for it."""
    if path is None:
        path = os.environ.get("ORBFE_LIB") or _build.LIB  # ORBFE_LIB: A/B-test another nvcc build of the same sources
        if not os.path.exists(path):
            _build.build()
    path = os.path.abspath(path)
    if path in _libs:
        if b"EMULATED" in _libs[path].orbfe_version() and not _test_emulation:
            raise OrbfeError("refusing to load the emulated TEST build as the product library")
        return _libs[path]
    try:
        L = C.CDLL(path)
    except OSError as e:  # fail loudly: no fallback
        raise OrbfeError(f"cannot load the CUDA library {path}: {e}") from e
    L.orbfe_last_error.restype = C.c_char_p
    L.orbfe_version.restype = C.c_char_p
    L.orbfe_launch_count.restype = C.c_longlong
    L.orbfe_launch_count.argtypes = [vp]
    if b"EMULATED" in L.orbfe_version() and not _test_emulation:
        raise OrbfeError("refusing to load the emulated TEST build as the product library")
    i, f, sz = C.c_int, C.c_float, C.c_size_t
    L.orbfe_extractor_create.argtypes = [C.POINTER(_Params), i, C.POINTER(vp)]
    L.orbfe_extractor_destroy.argtypes = [vp]
    L.orbfe_extractor_tables.argtypes = [vp] * 7
    L.orbfe_extractor_max_keypoints.argtypes = [vp]
    L.orbfe_extract.argtypes = [vp, vp, i, i, sz, vp, vp, i, vp]
    L.orbfe_extract_batch.argtypes = [vp, vp, i, i, i, sz, vp, vp, i, vp]
    L.orbfe_pyramid_level.argtypes = [vp, i, i, vp, sz, vp, vp]
    L.orbfe_upload.argtypes = [vp, i, vp, i, i, i, sz]
    L.orbfe_upload_color.argtypes = [vp, i, vp, i, i, i, sz, i, i]
    L.orbfe_run.argtypes = [vp, i]
    L.orbfe_run_stereo.argtypes = [vp, i, f, f]
    L.orbfe_download.argtypes = [vp, i, vp, vp, i, vp, vp, vp]
    L.orbfe_download_async.argtypes = [vp, i, vp, vp, i, vp, vp, vp]
    L.orbfe_sync.argtypes = [vp]
    L.orbfe_event_record.argtypes = [vp, i]
    L.orbfe_event_elapsed_ms.argtypes = [vp, i, i, vp]
    L.orbfe_set_stage_timing.argtypes = [vp, i]
    L.orbfe_stage_summary.argtypes = [vp, vp, vp]
    L.orbfe_debug_candidates.argtypes = [vp, i, i, vp, i, vp]
    L.orbfe_debug_level_keypoints.argtypes = [vp, i, i, vp, i, vp]
    L.orbfe_debug_blurred.argtypes = [vp, i, i, vp, sz, vp, vp]
    L.orbfe_stereo_match.argtypes = [vp, vp, i, vp, vp, i, vp, vp, f, f, vp, vp, vp]
    L.orbfe_descriptor_distance.argtypes = [i, vp, vp, i, vp]
    L.orbfe_frame_create.argtypes = [i, i, vp, vp, vp, f, f, f, f, i, vp, C.POINTER(vp)]
    L.orbfe_frame_destroy.argtypes = [vp]
    L.orbfe_frame_pool_trim.argtypes = []
    L.orbfe_pinned_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    L.orbfe_pinned_free.argtypes = [vp]
    L.orbfe_frame_from_extractor.argtypes = [vp, i, i, f, f, f, f, C.POINTER(vp)]
    L.orbfe_frame_refresh_from_extractor.argtypes = [vp, vp, i, i, f, f, f, f]
    L.orbfe_frame_num_keypoints.argtypes = [vp]
    L.orbfe_features_in_area.argtypes = [vp, f, f, f, i, i, vp, i, vp]
    L.orbfe_search_for_initialization.argtypes = [vp, vp, vp, vp, i, f, i, vp]
    L.orbfe_search_by_projection_mappoints.argtypes = [vp, i] + [vp] * 9 + [i, f, vp, vp]
    L.orbfe_search_by_projection_lastframe.argtypes = [vp, i] + [vp] * 8 + [f, i, i, vp, f, i, vp, vp]
    L.orbfe_search_by_bow.argtypes = [vp, i, vp, vp, vp, i, vp, vp, vp, i, vp, vp, vp, f, i, vp, vp]
    L.orbfe_search_by_projection_sim3.argtypes = [vp, i] + [vp] * 6 + [i, vp, vp]
    L.orbfe_search_by_projection_keyframe.argtypes = [vp, i] + [vp] * 7 + [f, i, i, vp, vp]
    L.orbfe_fuse.argtypes = [vp, i] + [vp] * 6 + [f, vp, vp]
    L.orbfe_fuse_sim3.argtypes = [vp, i] + [vp] * 5 + [f, vp, vp]
    L.orbfe_search_by_sim3.argtypes = [vp, vp] + [vp] * 10 + [f, vp, vp]
    L.orbfe_search_by_bow_keyframes.argtypes = [vp, i, vp, vp, vp, vp, i, vp, vp, vp, i, vp, vp, vp, f, i, vp, vp]
    L.orbfe_search_for_triangulation.argtypes = [vp, i, vp, vp, vp, vp, vp, i, vp, vp, vp, i, vp, vp, vp, vp, f, f, i, i, vp, vp]
    L.orbfe_vocabulary_create.argtypes = [i, i, i, i, i, i, vp, vp, vp, vp, C.POINTER(vp)]
    L.orbfe_vocabulary_load_text.argtypes = [C.c_char_p, i, C.POINTER(vp)]
    L.orbfe_vocabulary_destroy.argtypes = [vp]
    L.orbfe_vocabulary_info.argtypes = [vp] + [vp] * 6
    L.orbfe_bow_transform.argtypes = [vp, i, vp, i] + [vp] * 9
    L.orbfe_undistort_keypoints.argtypes = [i, i, vp, f, f, f, f, vp, i, vp]
    L.orbfe_is_in_frustum.argtypes = [i, i] + [vp] * 8 + [f] * 10 + [i, f] + [vp] * 7
    L.orbfe_debug_logf.argtypes = [i, i, vp, vp]
    L.orbfe_search_local_points.argtypes = [vp, i] + [vp] * 8 + [f] * 7 + [vp] * 3 + [i, f] + [vp] * 5
    _libs[path] = L
    return L


def _p(a):
    return a.ctypes.data_as(vp) if a is not None else None


def _check(L, rc, allow=()):
    if rc != 0 and rc not in allow:
        raise OrbfeError(f"orbfe error {rc}: {L.orbfe_last_error().decode(errors='replace')}")
    return rc


class _PinnedBlock:
    """owner of one orbfe_pinned_alloc block: freed when the last numpy view of it is gone"""

    def __init__(self, L, nbytes):
        self.L, self.p = L, vp()
        _check(L, L.orbfe_pinned_alloc(nbytes, C.byref(self.p)))
        self.buf = (C.c_uint8 * max(nbytes, 1)).from_address(self.p.value)

    def __del__(self):
        if getattr(self, "p", None) and self.p.value:
            self.L.orbfe_pinned_free(self.p)
            self.p = vp()


def pinned_empty(shape, dtype=np.uint8, lib=None):
    """numpy array in PORTABLE page-locked host memory (orbfe_pinned_alloc): pinned for every device of the process, so frames
    and result buffers can be handed to the copy engines of several GPUs from one process"""
    L = lib or load()
    dt = np.dtype(dtype)
    n = int(np.prod(shape)) * dt.itemsize
    blk = _PinnedBlock(L, n)
    a = np.frombuffer(blk.buf, dtype=np.uint8, count=n).view(dt).reshape(shape)
    _pinned_owners[id(blk.buf)] = blk   # np.frombuffer keeps blk.buf alive; the block object is parked until release_pinned()
    return a


_pinned_owners = {}


def release_pinned():
    """frees every pinned_empty block (call when no array from pinned_empty is in use any more)"""
    _pinned_owners.clear()


def device_count(lib=None):
    return (lib or load()).orbfe_device_count()


class ORBextractor:
    """Mirror of ORBextractor (src/orb_features/orb_extractor.h:25-93)."""

    def __init__(self, nfeatures=2000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7, device=0,
                 max_images=1, max_size=None, lib=None):
        self.L = lib or load()
        self.nfeatures, self.nlevels, self.max_images = nfeatures, nlevels, max_images
        w, h = (max_size or (0, 0))
        prm = _Params(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, w, h, max_images)
        self.h = vp()
        _check(self.L, self.L.orbfe_extractor_create(C.byref(prm), device, C.byref(self.h)))
        n = nlevels
        self._scale, self._inv, self._s2, self._is2 = (np.zeros(n, np.float32) for _ in range(4))
        self._fpl = np.zeros(n, np.int32)
        nl = C.c_int()
        _check(self.L, self.L.orbfe_extractor_tables(self.h, C.byref(nl), _p(self._scale), _p(self._inv), _p(self._s2),
                                                    _p(self._is2), _p(self._fpl)))

    def close(self):
        if getattr(self, "h", None):
            self.L.orbfe_extractor_destroy(self.h)
            self.h = None

    __del__ = close

    # -- orb_extractor.h:46-58
    def GetLevels(self): return self.nlevels
    def GetScaleFactor(self): return float(self._scale[1]) if self.nlevels > 1 else 1.0
    def GetScaleFactors(self): return self._scale.copy()
    def GetInverseScaleFactors(self): return self._inv.copy()
    def GetScaleSigmaSquares(self): return self._s2.copy()
    def GetInverseScaleSigmaSquares(self): return self._is2.copy()
    def features_per_level(self): return self._fpl.copy()

    def max_keypoints(self):
        return self.L.orbfe_extractor_max_keypoints(self.h)

    def Compute(self, image, mask=None):
        """ORBextractor::Compute(image, mask, keypoints, descriptors) (orb_extractor.cpp:985-1049).
        Returns (keypoints[KP_DTYPE], descriptors[N,32] u8); mask is ignored as in the reference."""
        if image is None or image.size == 0:
            return np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8)
        if image.dtype != np.uint8 or image.ndim != 2:
            raise OrbfeError("image must be CV_8UC1 (2-D uint8)")  # assert at orb_extractor.cpp:994
        if image.strides[1] != 1:
            image = np.ascontiguousarray(image)
        cap = self.max_keypoints()
        while True:
            kps = np.zeros(cap, KP_DTYPE)
            desc = np.zeros((cap, 32), np.uint8)
            n = C.c_int()
            rc = _check(self.L, self.L.orbfe_extract(self.h, _p(image), image.shape[1], image.shape[0], image.strides[0],
                                                    _p(kps), _p(desc), cap, C.byref(n)), allow=(-3,))
            if rc == 0:
                return kps[:n.value].copy(), desc[:n.value].copy()
            cap = max(n.value, self.max_keypoints())

    __call__ = Compute  # upstream ORB-SLAM2 spells the entry point operator()

    def GetImagePyramid(self, slot=0):
        """mvImagePyramid ROI views (orb_extractor.h:58), copied from the device."""
        out = []
        for l in range(self.nlevels):
            out.append(self.pyramid_level(l, slot))
        return out

    def pyramid_level(self, level, slot=0):
        w, h = C.c_int(), C.c_int()
        _check(self.L, self.L.orbfe_pyramid_level(self.h, slot, level, None, 0, C.byref(w), C.byref(h)))
        a = np.empty((h.value, w.value), np.uint8)
        _check(self.L, self.L.orbfe_pyramid_level(self.h, slot, level, _p(a), a.strides[0], C.byref(w), C.byref(h)))
        return a

    # -- stage taps for parity tests
    def debug_candidates(self, level, slot=0):
        out = np.zeros(1 << 18, KP_DTYPE)
        n = C.c_int()
        _check(self.L, self.L.orbfe_debug_candidates(self.h, slot, level, _p(out), len(out), C.byref(n)))
        return out[:n.value].copy()

    def debug_level_keypoints(self, level, slot=0):
        out = np.zeros(1 << 16, KP_DTYPE)
        n = C.c_int()
        _check(self.L, self.L.orbfe_debug_level_keypoints(self.h, slot, level, _p(out), len(out), C.byref(n)))
        return out[:n.value].copy()

    def debug_blurred(self, level, slot=0):
        w, h = C.c_int(), C.c_int()
        _check(self.L, self.L.orbfe_debug_blurred(self.h, slot, level, None, 0, C.byref(w), C.byref(h)))
        a = np.empty((h.value, w.value), np.uint8)
        _check(self.L, self.L.orbfe_debug_blurred(self.h, slot, level, _p(a), a.strides[0], C.byref(w), C.byref(h)))
        return a

    # -- batched / device-resident path (bench, offline loop of examples/main_stereo.cpp:102-143)
    def _ptr_array(self, imgs):
        arr = (vp * len(imgs))()
        for i, im in enumerate(imgs):
            assert im.dtype == np.uint8 and im.ndim == 2 and im.strides[1] == 1
            assert im.shape == imgs[0].shape and im.strides[0] == imgs[0].strides[0]
            arr[i] = im.ctypes.data
        return arr

    def upload(self, imgs, first_slot=0):
        arr = self._ptr_array(imgs)
        im = imgs[0]
        _check(self.L, self.L.orbfe_upload(self.h, first_slot, arr, len(imgs), im.shape[1], im.shape[0], im.strides[0]))

    def upload_color(self, imgs, rgb=True, first_slot=0):
        """colour frames (h, w, 3|4) u8: gray conversion on the device (cv::cvtColor, tracker.cpp:110-127)"""
        arr = (vp * len(imgs))()
        im = imgs[0]
        for i, a in enumerate(imgs):
            assert a.dtype == np.uint8 and a.ndim == 3 and a.shape == im.shape and a.strides[2] == 1 and a.strides[1] == a.shape[2]
            arr[i] = a.ctypes.data
        _check(self.L, self.L.orbfe_upload_color(self.h, first_slot, arr, len(imgs), im.shape[1], im.shape[0], im.strides[0],
                                                im.shape[2], int(rgb)))

    def upload_ptrs(self, ptr_array, n, w, h, stride, first_slot=0):
        _check(self.L, self.L.orbfe_upload(self.h, first_slot, ptr_array, n, w, h, stride))

    def run(self, n_imgs):
        _check(self.L, self.L.orbfe_run(self.h, n_imgs))

    def run_stereo(self, n_pairs, bf, baseline):
        _check(self.L, self.L.orbfe_run_stereo(self.h, n_pairs, bf, baseline))

    def sync(self):
        _check(self.L, self.L.orbfe_sync(self.h))

    def make_buffers(self, n_imgs, stereo=False):
        cap = self.max_keypoints()
        b = dict(kps=np.zeros((n_imgs, cap), KP_DTYPE), desc=np.zeros((n_imgs, cap, 32), np.uint8),
                 n=np.zeros(n_imgs, np.int32), cap=cap, ur=None, depth=None)
        if stereo:
            b["ur"] = np.zeros((n_imgs, cap), np.float32)
            b["depth"] = np.zeros((n_imgs, cap), np.float32)
        return b

    def download(self, n_imgs, buf):
        _check(self.L, self.L.orbfe_download(self.h, n_imgs, _p(buf["kps"]), _p(buf["desc"]), buf["cap"], _p(buf["n"]),
                                            _p(buf["ur"]), _p(buf["depth"])))
        return buf

    def download_async(self, n_imgs, buf):
        """enqueue the D2H copies straight into `buf` (make_buffers layout, ideally pinned); valid after sync()"""
        _check(self.L, self.L.orbfe_download_async(self.h, n_imgs, _p(buf["kps"]), _p(buf["desc"]), buf["cap"],
                                                  _p(buf["n"]), _p(buf["ur"]), _p(buf["depth"])))
        return buf

    def extract_batch(self, imgs):
        buf = self.make_buffers(len(imgs))
        arr = self._ptr_array(imgs)
        im = imgs[0]
        _check(self.L, self.L.orbfe_extract_batch(self.h, arr, len(imgs), im.shape[1], im.shape[0], im.strides[0],
                                                 _p(buf["kps"]), _p(buf["desc"]), buf["cap"], _p(buf["n"])))
        return [(buf["kps"][i, :buf["n"][i]].copy(), buf["desc"][i, :buf["n"][i]].copy()) for i in range(len(imgs))]

    def event_record(self, slot):
        _check(self.L, self.L.orbfe_event_record(self.h, slot))

    def event_elapsed_ms(self, a, b):
        ms = C.c_float()
        _check(self.L, self.L.orbfe_event_elapsed_ms(self.h, a, b, C.byref(ms)))
        return ms.value

    def set_stage_timing(self, on):
        _check(self.L, self.L.orbfe_set_stage_timing(self.h, int(on)))

    STAGES = ("pyramid", "fast", "quadtree", "blur", "describe", "stereo_search", "stereo_median")

    def stage_summary(self):
        """(dict stage -> summed ms, runs) since the last call (needs set_stage_timing(True))."""
        ms = (C.c_float * 7)()
        n = C.c_int()
        _check(self.L, self.L.orbfe_stage_summary(self.h, ms, C.byref(n)))
        return dict(zip(self.STAGES, [float(x) for x in ms])), n.value

    def launch_count(self):
        return self.L.orbfe_launch_count(self.h)


def ComputeStereoMatches(left, right, kps_left, desc_left, kps_right, desc_right, bf, baseline):
    """Frame::ComputeStereoMatches (src/data/frame.cpp:406-577).  `left`/`right` are the two
    ORBextractor objects that produced the keypoints (their device pyramids are read).
    Returns (n_matched, stereo_coords_[N], depths_[N])."""
    L = left.L
    kl = np.ascontiguousarray(kps_left, KP_DTYPE); kr = np.ascontiguousarray(kps_right, KP_DTYPE)
    dl = np.ascontiguousarray(desc_left, np.uint8); dr = np.ascontiguousarray(desc_right, np.uint8)
    ur = np.zeros(len(kl), np.float32)
    dp = np.zeros(len(kl), np.float32)
    n = C.c_int()
    _check(L, L.orbfe_stereo_match(left.h, right.h, len(kl), _p(kl), _p(dl), len(kr), _p(kr), _p(dr), bf, baseline,
                                   _p(ur), _p(dp), C.byref(n)))
    return n.value, ur, dp


class Frame:
    """The matcher's view of a Frame: undistorted keypoints, descriptors, stereo coordinates, image
    bounds, scale table; owns the 64x48 feature grid (frame.cpp:234-248, 339-403) on the GPU."""

    def __init__(self, kps_un, desc, scale_factors, bounds, u_right=None, device=0, lib=None):
        self.L = lib or load()
        self.kps = np.ascontiguousarray(kps_un, KP_DTYPE)
        self.desc = np.ascontiguousarray(desc, np.uint8)
        self.scale = np.ascontiguousarray(scale_factors, np.float32)
        self.ur = None if u_right is None else np.ascontiguousarray(u_right, np.float32)
        self.bounds = tuple(float(b) for b in bounds)
        self.h = vp()
        _check(self.L, self.L.orbfe_frame_create(device, len(self.kps), _p(self.kps), _p(self.desc), _p(self.ur),
                                                self.bounds[0], self.bounds[1], self.bounds[2], self.bounds[3],
                                                len(self.scale), _p(self.scale), C.byref(self.h)))

    @classmethod
    def from_extractor(cls, ex, bounds, slot=0, stereo=False):
        """the matcher view of the frame whose extraction results sit in slot `slot` of extractor `ex`, built device to
        device (orbfe_frame_from_extractor); `kps` is a placeholder of the right length, the features never reach the host"""
        self = cls.__new__(cls)
        self.L = ex.L
        self.bounds = tuple(float(b) for b in bounds)
        self.h = vp()
        _check(self.L, self.L.orbfe_frame_from_extractor(ex.h, int(slot), int(stereo), *self.bounds, C.byref(self.h)))
        self.kps = np.zeros(self.L.orbfe_frame_num_keypoints(self.h), KP_DTYPE)
        self.desc, self.ur, self.scale = None, None, ex.GetScaleFactors()
        return self

    def refresh_from_extractor(self, ex, slot=0, stereo=False):
        """the next frame's results into the same handle (orbfe_frame_refresh_from_extractor): nothing is allocated in steady state"""
        _check(self.L, self.L.orbfe_frame_refresh_from_extractor(self.h, ex.h, int(slot), int(stereo), *self.bounds))
        self.kps = np.zeros(self.L.orbfe_frame_num_keypoints(self.h), KP_DTYPE)
        return self

    def close(self):
        if getattr(self, "h", None):
            self.L.orbfe_frame_destroy(self.h)
            self.h = None

    __del__ = close

    def GetFeaturesInArea(self, x, y, r, minLevel=-1, maxLevel=-1):
        out = np.zeros(len(self.kps) + 1, np.int32)
        n = C.c_int()
        _check(self.L, self.L.orbfe_features_in_area(self.h, x, y, r, minLevel, maxLevel, _p(out), len(out), C.byref(n)))
        return out[:n.value].copy()


def DescriptorDistance(a, b, device=0, lib=None):
    """static OrbMatcher::DescriptorDistance (orb_matcher.cpp:1630-1646), batched over rows."""
    L = lib or load()
    a = np.ascontiguousarray(a, np.uint8).reshape(-1, 32)
    b = np.ascontiguousarray(b, np.uint8).reshape(-1, 32)
    d = np.zeros(len(a), np.int32)
    _check(L, L.orbfe_descriptor_distance(device, _p(a), _p(b), len(a), _p(d)))
    return d


class OrbMatcher:
    """Mirror of OrbMatcher (src/orb_features/orb_matcher.h:14-119) for the hot-path routines."""
    TH_LOW, TH_HIGH, HISTO_LENGTH = TH_LOW, TH_HIGH, HISTO_LENGTH

    def __init__(self, nnratio=0.6, checkOri=True):
        self.nnratio, self.checkOri = float(nnratio), bool(checkOri)

    DescriptorDistance = staticmethod(DescriptorDistance)

    def SearchForInitialization(self, F1, F2, vbPrevMatched, windowSize=10):
        """orb_matcher.cpp:264-382 -> (nmatches, vnMatches12, updated vbPrevMatched)."""
        pm = np.ascontiguousarray(vbPrevMatched, np.float32).copy()
        m12 = np.zeros(len(F1.kps), np.int32)
        n = C.c_int()
        _check(F1.L, F1.L.orbfe_search_for_initialization(F1.h, F2.h, _p(pm), _p(m12), windowSize, self.nnratio,
                                                         int(self.checkOri), C.byref(n)))
        return n.value, m12, pm

    def SearchByProjectionMapPoints(self, F, valid, proj_x, proj_y, proj_xr, pred_level, view_cos, mp_desc, has_obs,
                                    occupied, th=1):
        """SearchByProjection(Frame&, const vector<MapPoint*>&, th) (orb_matcher.cpp:13-111)."""
        a = lambda v, t: np.ascontiguousarray(v, t)
        assigned = np.zeros(len(F.kps), np.int32)
        n = C.c_int()
        args = [a(valid, np.uint8), a(proj_x, np.float32), a(proj_y, np.float32), a(proj_xr, np.float32),
                a(pred_level, np.int32), a(view_cos, np.float32), a(mp_desc, np.uint8), a(has_obs, np.uint8),
                a(occupied, np.uint8)]
        _check(F.L, F.L.orbfe_search_by_projection_mappoints(F.h, len(args[0]), *[_p(x) for x in args], int(th),
                                                            self.nnratio, _p(assigned), C.byref(n)))
        return n.value, assigned

    def SearchByProjectionLastFrame(self, Cur, valid, u, v, invzc, last_octave, last_angle, mp_desc, has_obs, bf,
                                    forward, backward, occupied, th):
        """SearchByProjection(Frame& Cur, const Frame& Last, th, bMono) (orb_matcher.cpp:1312-1453)."""
        a = lambda x, t: np.ascontiguousarray(x, t)
        assigned = np.zeros(len(Cur.kps), np.int32)
        n = C.c_int()
        args = [a(valid, np.uint8), a(u, np.float32), a(v, np.float32), a(invzc, np.float32),
                a(last_octave, np.int32), a(last_angle, np.float32), a(mp_desc, np.uint8), a(has_obs, np.uint8)]
        _check(Cur.L, Cur.L.orbfe_search_by_projection_lastframe(Cur.h, len(args[0]), *[_p(x) for x in args], bf,
                                                                int(forward), int(backward),
                                                                _p(a(occupied, np.uint8)), th, int(self.checkOri),
                                                                _p(assigned), C.byref(n)))
        return n.value, assigned


def flatten_feature_vector(fv):
    """DBoW2::FeatureVector (dict node id -> list of feature indices) -> (ids ascending, starts, indices); a tuple that is
    already in that form passes through (callers that keep their feature vectors flat, as the C++ shim does)"""
    if isinstance(fv, tuple):
        return fv
    ids = np.array(sorted(fv), np.uint32)
    starts = np.zeros(len(ids) + 1, np.int32)
    idx = []
    for k, n in enumerate(ids):
        idx.extend(fv[int(n)])
        starts[k + 1] = len(idx)
    return ids, starts, np.array(idx, np.uint32)


def SearchByBoW(F, kf_desc, kf_angle, kf_valid, kf_fv, f_fv, nnratio=0.7, checkOri=True):
    """OrbMatcher::SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches) (orb_matcher.cpp:133-262).  kf_fv / f_fv are
    the DBoW2 feature vectors as dicts; returns (nmatches, matched_kf_idx[NumKeypoints])."""
    kd = np.ascontiguousarray(kf_desc, np.uint8)
    ka = np.ascontiguousarray(kf_angle, np.float32)
    kv = np.ascontiguousarray(kf_valid, np.uint8)
    ki, ks, kx = flatten_feature_vector(kf_fv)
    fi, fs, fx = flatten_feature_vector(f_fv)
    out = np.zeros(len(F.kps), np.int32)
    n = C.c_int()
    _check(F.L, F.L.orbfe_search_by_bow(F.h, len(kd), _p(kd), _p(ka), _p(kv), len(ki), _p(ki), _p(ks), _p(kx), len(fi), _p(fi),
                                        _p(fs), _p(fx), nnratio, int(checkOri), _p(out), C.byref(n)))
    return n.value, out


# ---- the remaining OrbMatcher searches; argument conventions of include/orbfe.h ---------------
def _a(x, t):
    return np.ascontiguousarray(x, t)


def SearchByProjectionSim3(KF, valid, u, v, pred_level, mp_desc, matched_in, th):
    """SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th) (orb_matcher.cpp:384-497) -> (nmatches, matched[N])"""
    out = np.zeros(len(KF.kps), np.int32)
    n = C.c_int()
    args = [_a(valid, np.uint8), _a(u, np.float32), _a(v, np.float32), _a(pred_level, np.int32), _a(mp_desc, np.uint8),
            _a(matched_in, np.uint8)]
    _check(KF.L, KF.L.orbfe_search_by_projection_sim3(KF.h, len(args[0]), *[_p(x) for x in args], int(th), _p(out), C.byref(n)))
    return n.value, out


def SearchByProjectionKeyFrame(Cur, valid, u, v, pred_level, kf_angle, mp_desc, occupied, th, ORBdist, checkOri=True):
    """SearchByProjection(Frame& Cur, KeyFrame*, sAlreadyFound, th, ORBdist) (orb_matcher.cpp:1455-1582)"""
    out = np.zeros(len(Cur.kps), np.int32)
    n = C.c_int()
    args = [_a(valid, np.uint8), _a(u, np.float32), _a(v, np.float32), _a(pred_level, np.int32), _a(kf_angle, np.float32),
            _a(mp_desc, np.uint8), _a(occupied, np.uint8)]
    _check(Cur.L, Cur.L.orbfe_search_by_projection_keyframe(Cur.h, len(args[0]), *[_p(x) for x in args], float(th), int(ORBdist),
                                                           int(checkOri), _p(out), C.byref(n)))
    return n.value, out


def Fuse(KF, valid, u, v, ur, pred_level, mp_desc, th):
    """the search of Fuse(KeyFrame*, vpMapPoints, th) (orb_matcher.cpp:804-954); ur=None: Fuse(KeyFrame*, Scw, ...)
    (:956-1079) -> (nFused, best_idx[n_mp])"""
    n_mp = len(valid)
    out = np.zeros(n_mp, np.int32)
    n = C.c_int()
    args = [_a(valid, np.uint8), _a(u, np.float32), _a(v, np.float32)]
    tail = [_a(pred_level, np.int32), _a(mp_desc, np.uint8)]
    if ur is None:
        _check(KF.L, KF.L.orbfe_fuse_sim3(KF.h, n_mp, *[_p(x) for x in args + tail], float(th), _p(out), C.byref(n)))
    else:
        args.append(_a(ur, np.float32))
        _check(KF.L, KF.L.orbfe_fuse(KF.h, n_mp, *[_p(x) for x in args + tail], float(th), _p(out), C.byref(n)))
    return n.value, out


def SearchBySim3(KF1, KF2, side1, side2, th):
    """SearchBySim3 (orb_matcher.cpp:1081-1310); side = (valid, u, v, pred_level, mp_desc) -> (nFound, match12[N1])"""
    def pack(sd):
        return [_a(sd[0], np.uint8), _a(sd[1], np.float32), _a(sd[2], np.float32), _a(sd[3], np.int32), _a(sd[4], np.uint8)]
    a1, a2 = pack(side1), pack(side2)
    out = np.zeros(len(KF1.kps), np.int32)
    n = C.c_int()
    _check(KF1.L, KF1.L.orbfe_search_by_sim3(KF1.h, KF2.h, *[_p(x) for x in a1 + a2], float(th), _p(out), C.byref(n)))
    return n.value, out


def SearchByBoWKeyFrames(KF2, desc1, angle1, valid1, valid2, fv1, fv2, nnratio=0.8, checkOri=True):
    """SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12) (orb_matcher.cpp:499-632) -> (nmatches, matches12[N1])"""
    d1, a1, v1, v2 = _a(desc1, np.uint8), _a(angle1, np.float32), _a(valid1, np.uint8), _a(valid2, np.uint8)
    i1, s1, x1 = flatten_feature_vector(fv1)
    i2, s2, x2 = flatten_feature_vector(fv2)
    out = np.zeros(len(d1), np.int32)
    n = C.c_int()
    _check(KF2.L, KF2.L.orbfe_search_by_bow_keyframes(KF2.h, len(d1), _p(d1), _p(a1), _p(v1), _p(v2), len(i1), _p(i1), _p(s1),
                                                     _p(x1), len(i2), _p(i2), _p(s2), _p(x2), nnratio, int(checkOri), _p(out),
                                                     C.byref(n)))
    return n.value, out


def SearchForTriangulation(KF2, kps1, desc1, valid1, stereo1, valid2, fv1, fv2, F12, ex, ey, onlyStereo=False, checkOri=True):
    """SearchForTriangulation (orb_matcher.cpp:634-802) -> (nmatches, matches12[N1]); vMatchedPairs =
    [(i, matches12[i]) for i if matches12[i] >= 0]"""
    k1, d1 = _a(kps1, KP_DTYPE), _a(desc1, np.uint8)
    v1, st1, v2 = _a(valid1, np.uint8), _a(stereo1, np.uint8), _a(valid2, np.uint8)
    i1, s1, x1 = flatten_feature_vector(fv1)
    i2, s2, x2 = flatten_feature_vector(fv2)
    F = _a(np.asarray(F12, np.float32).reshape(9), np.float32)
    out = np.zeros(len(k1), np.int32)
    n = C.c_int()
    _check(KF2.L, KF2.L.orbfe_search_for_triangulation(KF2.h, len(k1), _p(k1), _p(d1), _p(v1), _p(st1), _p(v2), len(i1), _p(i1),
                                                      _p(s1), _p(x1), len(i2), _p(i2), _p(s2), _p(x2), _p(F), float(ex), float(ey),
                                                      int(onlyStereo), int(checkOri), _p(out), C.byref(n)))
    return n.value, out


class OrbVocabulary:
    """Mirror of OrbVocabulary = DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB> for the transform path.
    Build it from node arrays (`parent`, `is_leaf`, `desc`, `weight`, node 0 = root) or with loadFromTextFile (ORBvoc.txt)."""

    def __init__(self, k, L, scoring, weighting, parent, is_leaf, desc, weight, device=0, lib=None, _handle=None):
        self.L_ = lib or load()
        if _handle is not None:
            self.h = _handle
            return
        parent = np.ascontiguousarray(parent, np.int32)
        is_leaf = np.ascontiguousarray(is_leaf, np.uint8)
        desc = np.ascontiguousarray(desc, np.uint8)
        weight = np.ascontiguousarray(weight, np.float64)
        self.h = vp()
        _check(self.L_, self.L_.orbfe_vocabulary_create(device, k, L, scoring, weighting, len(parent), _p(parent), _p(is_leaf),
                                                       _p(desc), _p(weight), C.byref(self.h)))

    @classmethod
    def loadFromTextFile(cls, path, device=0, lib=None):
        L = lib or load()
        h = vp()
        _check(L, L.orbfe_vocabulary_load_text(str(path).encode(), device, C.byref(h)))
        return cls(0, 0, 0, 0, None, None, None, None, lib=L, _handle=h)

    def info(self):
        v = [C.c_int() for _ in range(6)]
        _check(self.L_, self.L_.orbfe_vocabulary_info(self.h, *[C.byref(x) for x in v]))
        return dict(zip(("k", "L", "scoring", "weighting", "nodes", "words"), [x.value for x in v]))

    def close(self):
        if getattr(self, "h", None):
            self.L_.orbfe_vocabulary_destroy(self.h)
            self.h = None

    __del__ = close

    def transform(self, desc, levelsup=4):
        """-> dict(word_id, node_id, bow = (words, values), fv = (nodes, start, idx)); Frame::ComputeBoW uses levelsup 4"""
        d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(d)
        word, node = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
        bw, bv = np.zeros(n, np.uint32), np.zeros(n, np.float64)
        fn, fs, fi = np.zeros(n, np.uint32), np.zeros(n + 1, np.int32), np.zeros(n, np.uint32)
        nb, nf = C.c_int(), C.c_int()
        _check(self.L_, self.L_.orbfe_bow_transform(self.h, n, _p(d), int(levelsup), _p(word), _p(node), _p(bw), _p(bv), C.byref(nb),
                                                   _p(fn), _p(fs), _p(fi), C.byref(nf)))
        nb, nf = nb.value, nf.value
        return dict(word_id=word, node_id=node, bow=(bw[:nb].copy(), bv[:nb].copy()),
                    fv=(fn[:nf].copy(), fs[:nf + 1].copy(), fi[:fs[nf]].copy()))


def feature_vector_dict(fv):
    """flattened FeatureVector (nodes, start, idx) -> {node: [feature indices]} as orbfe.SearchByBoW* take"""
    nodes, start, idx = fv
    return {int(nodes[k]): [int(x) for x in idx[start[k]:start[k + 1]]] for k in range(len(nodes))}


def UndistortKeyPoints(kps, fx, fy, cx, cy, dist_coeffs, device=0, lib=None):
    """Frame::UndistortKeyPoints (frame.cpp:614-641) -> undistorted_keypoints_ (same dtype as kps)"""
    L = lib or load()
    k = np.ascontiguousarray(kps, KP_DTYPE)
    d = np.ascontiguousarray(dist_coeffs, np.float32)
    out = np.zeros_like(k)
    _check(L, L.orbfe_undistort_keypoints(device, len(k), _p(k), fx, fy, cx, cy, _p(d), len(d), _p(out)))
    return out


def IsInFrustum(world, normal, min_dist, max_dist, max_dist_raw, Rcw, tcw, Ow, fx, fy, cx, cy, bf, bounds, log_scale_factor, n_levels,
                viewingCosLimit=0.5, device=0, lib=None):
    """Frame::IsInFrustum over a batch (frame.cpp:277-337) -> (n_in_view, dict of the track_* arrays)"""
    L = lib or load()
    w, nrm = _a(world, np.float32).reshape(-1, 3), _a(normal, np.float32).reshape(-1, 3)
    n = len(w)
    out = dict(in_view=np.zeros(n, np.uint8), proj_x=np.zeros(n, np.float32), proj_y=np.zeros(n, np.float32),
               proj_xr=np.zeros(n, np.float32), level=np.zeros(n, np.int32), view_cos=np.zeros(n, np.float32))
    cnt = C.c_int()
    _check(L, L.orbfe_is_in_frustum(device, n, _p(w), _p(nrm), _p(_a(min_dist, np.float32)), _p(_a(max_dist, np.float32)), _p(_a(max_dist_raw, np.float32)),
                                    _p(_a(Rcw, np.float32).reshape(9)), _p(_a(tcw, np.float32).reshape(3)),
                                    _p(_a(Ow, np.float32).reshape(3)), fx, fy, cx, cy, bf, bounds[0], bounds[1], bounds[2], bounds[3],
                                    log_scale_factor, n_levels, viewingCosLimit, _p(out["in_view"]), _p(out["proj_x"]),
                                    _p(out["proj_y"]), _p(out["proj_xr"]), _p(out["level"]), _p(out["view_cos"]), C.byref(cnt)))
    return cnt.value, out


def debug_logf(x, device=0, lib=None):
    L = lib or load()
    x = _a(x, np.float32)
    y = np.zeros_like(x)
    _check(L, L.orbfe_debug_logf(device, len(x), _p(x), _p(y)))
    return y


def SearchLocalPoints(F, world, normal, min_dist, max_dist, max_dist_raw, Rcw, tcw, Ow, fx, fy, cx, cy, bf, log_scale_factor,
                      mp_desc, has_obs, occupied, th=1, nnratio=0.8, viewingCosLimit=0.5):
    """Tracker::SearchLocalPoints (core/tracker.cpp:1196-1226): IsInFrustum chained into SearchByProjection on the device
    -> (n_in_view, n_matches, in_view[n], scale_level[n], assigned[NumKeypoints])"""
    w, nrm = _a(world, np.float32).reshape(-1, 3), _a(normal, np.float32).reshape(-1, 3)
    n = len(w)
    in_view, lvl, assigned = np.zeros(n, np.uint8), np.zeros(n, np.int32), np.zeros(len(F.kps), np.int32)
    nv, nm = C.c_int(), C.c_int()
    _check(F.L, F.L.orbfe_search_local_points(
        F.h, n, _p(w), _p(nrm), _p(_a(min_dist, np.float32)), _p(_a(max_dist, np.float32)), _p(_a(max_dist_raw, np.float32)),
        _p(_a(Rcw, np.float32).reshape(9)), _p(_a(tcw, np.float32).reshape(3)), _p(_a(Ow, np.float32).reshape(3)), fx, fy, cx, cy, bf,
        log_scale_factor, viewingCosLimit, _p(_a(mp_desc, np.uint8)), _p(_a(has_obs, np.uint8)), _p(_a(occupied, np.uint8)), int(th),
        nnratio, _p(in_view), _p(lvl), _p(assigned), C.byref(nv), C.byref(nm)))
    return nv.value, nm.value, in_view, lvl, assigned
