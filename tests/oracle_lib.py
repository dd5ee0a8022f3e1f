"""ctypes binding of the CPU oracle (oracle/liborb_oracle.so).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs, never by the
product package."""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28

_lib = None
u8p = C.POINTER(C.c_uint8)
f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int)


def build(force=False):
    so = os.path.join(ORACLE_DIR, "liborb_oracle.so")
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("orb_oracle.cpp", "orb_oracle_batch.cpp", "orb_oracle.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return so


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        vp = C.c_void_p
        L.orc_resize_linear.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, C.c_int, C.c_int]
        L.orc_border_reflect101.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_gaussian7x7.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.c_int]
        L.orc_fast9.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int]
        L.orc_cvt_gray.argtypes = [vp, C.c_size_t, C.c_int, C.c_int, vp]
        L.orc_fast_atan2.argtypes = [C.c_float, C.c_float]
        L.orc_fast_atan2.restype = C.c_float
        L.orc_descriptor_distance.argtypes = [vp, vp]
        L.orc_extractor_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.orc_extractor_create.restype = vp
        L.orc_extractor_destroy.argtypes = [vp]
        L.orc_extract.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int]
        L.orc_levels.argtypes = [vp]
        L.orc_scale_factors.argtypes = [vp, vp, vp, vp, vp]
        L.orc_features_per_level.argtypes = [vp, vp]
        L.orc_umax.argtypes = [vp, vp]
        for fn in (L.orc_pyramid_level, L.orc_pyramid_padded, L.orc_stage_blurred):
            fn.argtypes = [vp, C.c_int, i32p, i32p, i32p]
            fn.restype = vp
        L.orc_stage_candidates.argtypes = [vp, C.c_int, vp, C.c_int]
        L.orc_stage_level_keypoints.argtypes = [vp, C.c_int, vp, C.c_int]
        L.orc_distribute_octree.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int]
        L.orc_stereo_match.argtypes = [vp, vp, C.c_int, vp, vp, C.c_int, vp, vp, C.c_float, C.c_float, vp, vp]
        L.orc_frame_create.argtypes = [C.c_int, vp, vp, vp, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, vp]
        L.orc_frame_create.restype = vp
        L.orc_frame_destroy.argtypes = [vp]
        L.orc_features_in_area.argtypes = [vp, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, vp, C.c_int]
        L.orc_search_for_initialization.argtypes = [vp, vp, vp, vp, C.c_int, C.c_float, C.c_int]
        L.orc_search_by_projection_mappoints.argtypes = [vp, C.c_int] + [vp] * 9 + [C.c_int, C.c_float, vp]
        L.orc_search_by_projection_lastframe.argtypes = [vp, C.c_int] + [vp] * 8 + [C.c_float, C.c_int, C.c_int, vp,
                                                                                 C.c_float, C.c_int, vp]
        L.orc_search_by_bow.argtypes = [vp, C.c_int, vp, vp, vp, C.c_int, vp, vp, vp, C.c_int, vp, vp, vp, C.c_float, C.c_int, vp]
        L.orc_search_by_projection_sim3.argtypes = [vp, C.c_int] + [vp] * 6 + [C.c_int, vp]
        L.orc_search_by_projection_keyframe.argtypes = [vp, C.c_int] + [vp] * 7 + [C.c_float, C.c_int, C.c_int, vp]
        L.orc_fuse.argtypes = [vp, C.c_int] + [vp] * 6 + [C.c_float, vp]
        L.orc_search_by_sim3.argtypes = [vp, vp] + [vp] * 10 + [C.c_float, vp]
        L.orc_search_by_bow_keyframes.argtypes = [vp, C.c_int, vp, vp, vp, vp, C.c_int, vp, vp, vp, C.c_int, vp, vp, vp,
                                                  C.c_float, C.c_int, vp]
        L.orc_search_for_triangulation.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, C.c_int, vp, vp, vp, C.c_int, vp, vp, vp,
                                                   vp, C.c_float, C.c_float, C.c_int, C.c_int, vp]
        L.orc_vocabulary_create.argtypes = [C.c_int] * 5 + [vp] * 4
        L.orc_vocabulary_create.restype = vp
        L.orc_vocabulary_destroy.argtypes = [vp]
        L.orc_bow_transform.argtypes = [vp, C.c_int, vp, C.c_int] + [vp] * 9
        L.orc_undistort_points.argtypes = [C.c_int, vp, C.c_float, C.c_float, C.c_float, C.c_float, vp, C.c_int, vp]
        L.orc_undistort_points.restype = None
        L.orc_is_in_frustum.argtypes = [C.c_int] + [vp] * 8 + [C.c_float] * 10 + [C.c_int, C.c_float] + [vp] * 6
        L.orc_bench_stereo_batch.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int,
                                             C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, vp, vp]
        L.orc_bench_stereo_batch.restype = C.c_double
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def resize_linear(src, dw, dh):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty((dh, dw), np.uint8)
    lib().orc_resize_linear(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dw, dh, dw)
    return dst


def border_reflect101(img, b):
    h, w = img.shape
    buf = np.zeros((h + 2 * b, w + 2 * b), np.uint8)
    buf[b:b + h, b:b + w] = img
    lib().orc_border_reflect101(_p(buf), w, h, buf.strides[0], b)
    return buf


def gaussian7x7(src):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty_like(src)
    lib().orc_gaussian7x7(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dst.strides[0])
    return dst


def fast9(img, threshold, nms=True):
    img = np.ascontiguousarray(img, np.uint8)
    cap = img.size
    out = np.zeros(cap, KP_DTYPE)
    n = lib().orc_fast9(_p(img), img.shape[1], img.shape[0], img.strides[0], threshold, int(nms), _p(out), cap)
    return out[:n]


def cvt_gray(img, rgb_order=True):
    img = np.ascontiguousarray(img, np.uint8)
    out = np.empty(img.shape[:2], np.uint8)
    lib().orc_cvt_gray(_p(img), img.shape[0] * img.shape[1], img.shape[2], int(rgb_order), _p(out))
    return out


def fast_atan2(y, x):
    return lib().orc_fast_atan2(float(y), float(x))


def descriptor_distance(a, b):
    a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
    return lib().orc_descriptor_distance(_p(a), _p(b))


def distribute_octree(cand, minX, maxX, minY, maxY, N):
    cand = np.ascontiguousarray(cand, KP_DTYPE)
    out = np.zeros(max(len(cand), 1), KP_DTYPE)
    n = lib().orc_distribute_octree(_p(cand), len(cand), minX, maxX, minY, maxY, N, _p(out), len(out))
    return out[:n]


class Extractor:
    def __init__(self, nfeatures=2000, scale=1.2, nlevels=8, ini_th=20, min_th=7):
        self.h = lib().orc_extractor_create(nfeatures, scale, nlevels, ini_th, min_th)
        self.nfeatures, self.nlevels = nfeatures, nlevels

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_extractor_destroy(self.h)
            self.h = None

    def extract(self, img):
        img = np.ascontiguousarray(img, np.uint8)
        cap = self.nfeatures + 64 * self.nlevels + 64
        while True:
            kps = np.zeros(cap, KP_DTYPE)
            desc = np.zeros((cap, 32), np.uint8)
            n = lib().orc_extract(self.h, _p(img), img.shape[1], img.shape[0], img.strides[0], _p(kps), _p(desc), cap)
            if n >= 0:
                return kps[:n].copy(), desc[:n].copy()
            cap = -n + 16

    def tables(self):
        n = self.nlevels
        s, i, s2, i2 = (np.zeros(n, np.float32) for _ in range(4))
        lib().orc_scale_factors(self.h, _p(s), _p(i), _p(s2), _p(i2))
        fpl = np.zeros(n, np.int32)
        lib().orc_features_per_level(self.h, _p(fpl))
        um = np.zeros(16, np.int32)
        lib().orc_umax(self.h, _p(um))
        return dict(scale=s, inv_scale=i, sigma2=s2, inv_sigma2=i2, features_per_level=fpl, umax=um)

    def _plane(self, fn, level):
        w, h, st = C.c_int(), C.c_int(), C.c_int()
        ptr = fn(self.h, level, C.byref(w), C.byref(h), C.byref(st))
        if not ptr or w.value == 0:
            return np.zeros((0, 0), np.uint8)
        buf = (C.c_uint8 * (st.value * (h.value - 1) + w.value)).from_address(ptr)
        a = np.frombuffer(buf, np.uint8)
        return np.lib.stride_tricks.as_strided(a, (h.value, w.value), (st.value, 1)).copy()

    def pyramid_level(self, level):
        return self._plane(lib().orc_pyramid_level, level)

    def pyramid_padded(self, level):
        return self._plane(lib().orc_pyramid_padded, level)

    def blurred(self, level):
        return self._plane(lib().orc_stage_blurred, level)

    def candidates(self, level):
        out = np.zeros(1 << 20, KP_DTYPE)
        n = lib().orc_stage_candidates(self.h, level, _p(out), len(out))
        return out[:n].copy()

    def level_keypoints(self, level):
        out = np.zeros(1 << 16, KP_DTYPE)
        n = lib().orc_stage_level_keypoints(self.h, level, _p(out), len(out))
        return out[:n].copy()


def stereo_match(exL, exR, kl, dl, kr, dr, bf, baseline):
    kl = np.ascontiguousarray(kl, KP_DTYPE); kr = np.ascontiguousarray(kr, KP_DTYPE)
    dl = np.ascontiguousarray(dl, np.uint8); dr = np.ascontiguousarray(dr, np.uint8)
    ur = np.zeros(len(kl), np.float32); dp = np.zeros(len(kl), np.float32)
    n = lib().orc_stereo_match(exL.h, exR.h, len(kl), _p(kl), _p(dl), len(kr), _p(kr), _p(dr), bf, baseline,
                               _p(ur), _p(dp))
    return n, ur, dp


class Frame:
    def __init__(self, kps, desc, scale_factors, bounds, u_right=None):
        self.kps = np.ascontiguousarray(kps, KP_DTYPE)
        self.desc = np.ascontiguousarray(desc, np.uint8)
        self.scale = np.ascontiguousarray(scale_factors, np.float32)
        self.ur = None if u_right is None else np.ascontiguousarray(u_right, np.float32)
        self.bounds = bounds
        self.h = lib().orc_frame_create(len(self.kps), _p(self.kps), _p(self.desc), _p(self.ur), bounds[0], bounds[1],
                                        bounds[2], bounds[3], len(self.scale), _p(self.scale))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_frame_destroy(self.h)
            self.h = None

    def features_in_area(self, x, y, r, min_level=-1, max_level=-1):
        out = np.zeros(len(self.kps) + 1, np.int32)
        n = lib().orc_features_in_area(self.h, x, y, r, min_level, max_level, _p(out), len(out))
        return out[:n].copy()


def search_for_initialization(f1, f2, prev_matched, window=100, nnratio=0.9, check_ori=True):
    pm = np.ascontiguousarray(prev_matched, np.float32).copy()
    m12 = np.zeros(len(f1.kps), np.int32)
    n = lib().orc_search_for_initialization(f1.h, f2.h, _p(pm), _p(m12), window, nnratio, int(check_ori))
    return n, m12, pm


def search_by_projection_mappoints(f, valid, px, py, pxr, level, viewcos, mp_desc, has_obs, occupied, th=1,
                                   nnratio=0.8):
    a = lambda v, t: np.ascontiguousarray(v, t)
    assigned = np.zeros(len(f.kps), np.int32)
    n = lib().orc_search_by_projection_mappoints(
        f.h, len(valid), _p(a(valid, np.uint8)), _p(a(px, np.float32)), _p(a(py, np.float32)),
        _p(a(pxr, np.float32)), _p(a(level, np.int32)), _p(a(viewcos, np.float32)), _p(a(mp_desc, np.uint8)),
        _p(a(has_obs, np.uint8)), _p(a(occupied, np.uint8)), th, nnratio, _p(assigned))
    return n, assigned


def search_by_projection_lastframe(cur, valid, u, v, invzc, octave, angle, mp_desc, has_obs, bf, forward, backward,
                                   occupied, th, check_ori=True):
    a = lambda x, t: np.ascontiguousarray(x, t)
    assigned = np.zeros(len(cur.kps), np.int32)
    n = lib().orc_search_by_projection_lastframe(
        cur.h, len(valid), _p(a(valid, np.uint8)), _p(a(u, np.float32)), _p(a(v, np.float32)),
        _p(a(invzc, np.float32)), _p(a(octave, np.int32)), _p(a(angle, np.float32)), _p(a(mp_desc, np.uint8)),
        _p(a(has_obs, np.uint8)), bf, int(forward), int(backward), _p(a(occupied, np.uint8)), th, int(check_ori),
        _p(assigned))
    return n, assigned


def search_by_bow(f, kf_desc, kf_angle, kf_valid, kf_fv, f_fv, nnratio=0.7, check_ori=True):
    from slam_framework_b200.orbfe import flatten_feature_vector
    kd = np.ascontiguousarray(kf_desc, np.uint8); ka = np.ascontiguousarray(kf_angle, np.float32)
    kv = np.ascontiguousarray(kf_valid, np.uint8)
    ki, ks, kx = flatten_feature_vector(kf_fv)
    fi, fs, fx = flatten_feature_vector(f_fv)
    out = np.zeros(len(f.kps), np.int32)
    n = lib().orc_search_by_bow(f.h, len(kd), _p(kd), _p(ka), _p(kv), len(ki), _p(ki), _p(ks), _p(kx), len(fi), _p(fi), _p(fs),
                                _p(fx), nnratio, int(check_ori), _p(out))
    return n, out


def _a(x, t):
    return np.ascontiguousarray(x, t)


def search_by_projection_sim3(kf, valid, u, v, pred_level, mp_desc, matched_in, th):
    out = np.zeros(len(kf.kps), np.int32)
    args = [_a(valid, np.uint8), _a(u, np.float32), _a(v, np.float32), _a(pred_level, np.int32), _a(mp_desc, np.uint8),
            _a(matched_in, np.uint8)]
    n = lib().orc_search_by_projection_sim3(kf.h, len(args[0]), *[_p(x) for x in args], int(th), _p(out))
    return n, out


def search_by_projection_keyframe(cur, valid, u, v, pred_level, kf_angle, mp_desc, occupied, th, orb_dist, check_ori=True):
    out = np.zeros(len(cur.kps), np.int32)
    args = [_a(valid, np.uint8), _a(u, np.float32), _a(v, np.float32), _a(pred_level, np.int32), _a(kf_angle, np.float32),
            _a(mp_desc, np.uint8), _a(occupied, np.uint8)]
    n = lib().orc_search_by_projection_keyframe(cur.h, len(args[0]), *[_p(x) for x in args], th, int(orb_dist), int(check_ori),
                                                _p(out))
    return n, out


def fuse(kf, valid, u, v, ur, pred_level, mp_desc, th):
    out = np.zeros(len(valid), np.int32)
    urp = None if ur is None else _a(ur, np.float32)
    args = [_a(valid, np.uint8), _a(u, np.float32), _a(v, np.float32), urp, _a(pred_level, np.int32), _a(mp_desc, np.uint8)]
    n = lib().orc_fuse(kf.h, len(valid), *[_p(x) for x in args], th, _p(out))
    return n, out


def search_by_sim3(kf1, kf2, side1, side2, th):
    def pack(sd):
        return [_a(sd[0], np.uint8), _a(sd[1], np.float32), _a(sd[2], np.float32), _a(sd[3], np.int32), _a(sd[4], np.uint8)]
    out = np.zeros(len(kf1.kps), np.int32)
    n = lib().orc_search_by_sim3(kf1.h, kf2.h, *[_p(x) for x in pack(side1) + pack(side2)], th, _p(out))
    return n, out


def search_by_bow_keyframes(kf2, desc1, angle1, valid1, valid2, fv1, fv2, nnratio=0.8, check_ori=True):
    from slam_framework_b200.orbfe import flatten_feature_vector
    d1, a1, v1, v2 = _a(desc1, np.uint8), _a(angle1, np.float32), _a(valid1, np.uint8), _a(valid2, np.uint8)
    i1, s1, x1 = flatten_feature_vector(fv1)
    i2, s2, x2 = flatten_feature_vector(fv2)
    out = np.zeros(len(d1), np.int32)
    n = lib().orc_search_by_bow_keyframes(kf2.h, len(d1), _p(d1), _p(a1), _p(v1), _p(v2), len(i1), _p(i1), _p(s1), _p(x1),
                                          len(i2), _p(i2), _p(s2), _p(x2), nnratio, int(check_ori), _p(out))
    return n, out


def search_for_triangulation(kf2, kps1, desc1, valid1, stereo1, valid2, fv1, fv2, F12, ex, ey, only_stereo=False, check_ori=True):
    from slam_framework_b200.orbfe import flatten_feature_vector
    k1, d1 = _a(kps1, KP_DTYPE), _a(desc1, np.uint8)
    v1, st1, v2 = _a(valid1, np.uint8), _a(stereo1, np.uint8), _a(valid2, np.uint8)
    i1, s1, x1 = flatten_feature_vector(fv1)
    i2, s2, x2 = flatten_feature_vector(fv2)
    F = _a(np.asarray(F12, np.float32).reshape(9), np.float32)
    out = np.zeros(len(k1), np.int32)
    n = lib().orc_search_for_triangulation(kf2.h, len(k1), _p(k1), _p(d1), _p(v1), _p(st1), _p(v2), len(i1), _p(i1), _p(s1),
                                           _p(x1), len(i2), _p(i2), _p(s2), _p(x2), _p(F), ex, ey, int(only_stereo),
                                           int(check_ori), _p(out))
    return n, out


class Vocabulary:
    def __init__(self, k, L, scoring, weighting, parent, is_leaf, desc, weight):
        self.arrays = (_a(parent, np.int32), _a(is_leaf, np.uint8), _a(desc, np.uint8), _a(weight, np.float64))
        self.h = lib().orc_vocabulary_create(k, L, scoring, weighting, len(self.arrays[0]), *[_p(x) for x in self.arrays])

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_vocabulary_destroy(self.h)
            self.h = None

    def transform(self, desc, levelsup=4):
        d = _a(desc, np.uint8).reshape(-1, 32)
        n = len(d)
        word, node = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
        bw, bv = np.zeros(n, np.uint32), np.zeros(n, np.float64)
        fn, fs, fi = np.zeros(n, np.uint32), np.zeros(n + 1, np.int32), np.zeros(n, np.uint32)
        nb, nf = C.c_int(), C.c_int()
        lib().orc_bow_transform(self.h, n, _p(d), int(levelsup), _p(word), _p(node), _p(bw), _p(bv), C.byref(nb), _p(fn), _p(fs),
                                _p(fi), C.byref(nf))
        nb, nf = nb.value, nf.value
        return dict(word_id=word, node_id=node, bow=(bw[:nb].copy(), bv[:nb].copy()),
                    fv=(fn[:nf].copy(), fs[:nf + 1].copy(), fi[:fs[nf]].copy()))


def undistort_points(xy, fx, fy, cx, cy, dist):
    xy = _a(xy, np.float32).reshape(-1, 2)
    dist = _a(dist, np.float32)
    out = np.zeros_like(xy)
    lib().orc_undistort_points(len(xy), _p(xy), fx, fy, cx, cy, _p(dist), len(dist), _p(out))
    return out


def is_in_frustum(world, normal, min_dist, max_dist, max_dist_raw, Rcw, tcw, Ow, fx, fy, cx, cy, bf, bounds, log_scale_factor, n_levels,
                  viewing_cos_limit=0.5):
    w, nrm = _a(world, np.float32).reshape(-1, 3), _a(normal, np.float32).reshape(-1, 3)
    n = len(w)
    out = dict(in_view=np.zeros(n, np.uint8), proj_x=np.zeros(n, np.float32), proj_y=np.zeros(n, np.float32),
               proj_xr=np.zeros(n, np.float32), level=np.zeros(n, np.int32), view_cos=np.zeros(n, np.float32))
    cnt = lib().orc_is_in_frustum(n, _p(w), _p(nrm), _p(_a(min_dist, np.float32)), _p(_a(max_dist, np.float32)), _p(_a(max_dist_raw, np.float32)),
                                  _p(_a(Rcw, np.float32).reshape(9)), _p(_a(tcw, np.float32).reshape(3)), _p(_a(Ow, np.float32).reshape(3)),
                                  fx, fy, cx, cy, bf, bounds[0], bounds[1], bounds[2], bounds[3], log_scale_factor, n_levels,
                                  viewing_cos_limit, _p(out["in_view"]), _p(out["proj_x"]), _p(out["proj_y"]), _p(out["proj_xr"]),
                                  _p(out["level"]), _p(out["view_cos"]))
    return cnt, out
