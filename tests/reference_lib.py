"""ctypes binding of oracle/_ref/libslam_ref.so: the reference's OWN sources (src/orb_features/orb_extractor.cpp, src/data/*,
third_party/DBoW2) compiled unmodified from /root/reference against oracle/cvstub by oracle/Makefile.ref.  TEST
INFRASTRUCTURE ONLY: used to pin the oracle's restatement of the control logic against the reference's real code."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "libslam_ref.so")
REF = "/root/reference"
KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"), ("octave", "<i4"),
                     ("class_id", "<i4")])
_lib = None
vp = C.c_void_p


def available():
    return os.path.isdir(os.path.join(REF, "src", "orb_features")) or os.path.exists(LIB)


def lib():
    global _lib
    if _lib is None:
        if os.path.isdir(os.path.join(REF, "src", "orb_features")):  # (re)build from the mounted reference tree
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"])
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-f", "Makefile.ref", "-s"])
        L = C.CDLL(LIB)
        L.ref_orb_extract.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, vp, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int]
        L.ref_orb_tables.argtypes = [C.c_int, C.c_float, C.c_int, vp, vp, vp, vp]
        L.ref_orb_tables.restype = None
        L.ref_voc_load_text.argtypes = [C.c_char_p]
        L.ref_voc_load_text.restype = vp
        L.ref_voc_destroy.argtypes = [vp]
        L.ref_voc_transform.argtypes = [vp, C.c_int, vp, C.c_int] + [vp] * 7
        L.ref_frame_stereo.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int] + [C.c_float] * 6
        L.ref_frame_stereo.restype = vp
        L.ref_frame_mono.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int] + [C.c_float] * 6
        L.ref_frame_mono.restype = vp
        L.ref_frame_destroy.argtypes = [vp]
        L.ref_set_test_transform.argtypes = [vp, C.c_float]
        L.ref_set_test_transform.restype = None
        L.ref_frame_set_translation.argtypes = [vp, vp]
        for fn in (L.ref_frame_n, L.ref_frame_n_right, L.ref_frame_levels):
            fn.argtypes = [vp]
        L.ref_frame_get.argtypes = [vp] * 9
        L.ref_frame_get_right.argtypes = [vp] * 3
        L.ref_features_in_area.argtypes = [vp, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, vp, C.c_int]
        L.ref_search_for_initialization.argtypes = [vp, vp, vp, vp, C.c_int, C.c_float, C.c_int]
        L.ref_search_by_projection_mappoints.argtypes = [vp, C.c_int] + [vp] * 9 + [C.c_int, C.c_float, vp]
        L.ref_search_by_projection_lastframe.argtypes = [vp, vp] + [vp] * 6 + [C.c_float, C.c_int, C.c_int, vp]
        L.ref_set_vocabulary.argtypes = [vp]
        L.ref_frame_get_bow.argtypes = [vp] * 8
        L.ref_keyframe_create.argtypes = [vp, vp, vp, vp]
        L.ref_keyframe_create.restype = vp
        L.ref_keyframe_destroy.argtypes = [vp]
        L.ref_search_by_bow_kf_f.argtypes = [vp, vp, C.c_float, C.c_int, vp]
        L.ref_search_by_bow_kf_kf.argtypes = [vp, vp, C.c_float, C.c_int, vp]
        L.ref_search_for_triangulation.argtypes = [vp, vp, vp, C.c_int, C.c_float, C.c_int, vp]
        L.ref_is_in_frustum.argtypes = [vp, C.c_int, vp, vp, vp, C.c_float] + [vp] * 10
        L.ref_search_by_projection_sim3.argtypes = [vp, C.c_int] + [vp] * 6 + [C.c_int] + [vp] * 4
        L.ref_fuse_sim3.argtypes = [vp, C.c_int] + [vp] * 5 + [C.c_float] + [vp] * 4
        L.ref_fuse.argtypes = [vp, C.c_int] + [vp] * 5 + [C.c_float] + [vp] * 4
        L.ref_search_by_projection_keyframe.argtypes = [vp, vp, vp, vp, vp, C.c_float, C.c_int, C.c_int, vp]
        L.ref_keyframe_create_at.argtypes = [vp] * 8
        L.ref_keyframe_create_at.restype = vp
        L.ref_search_by_sim3.argtypes = [vp, vp, vp, C.c_float, vp, vp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(vp) if a is not None else None


def set_test_distortion(d4=None):
    """k1 k2 p1 p2 of every Frame built from now on (None: undistorted camera again)"""
    L = lib()
    L.ref_set_test_distortion.argtypes = [C.c_void_p]
    L.ref_set_test_distortion.restype = None
    if d4 is None:
        L.ref_set_test_distortion(None)
    else:
        d = np.ascontiguousarray(d4, np.float32)
        assert d.shape == (4,)
        L.ref_set_test_distortion(d.ctypes.data)


def set_test_transform(R=None, sim3_scale=1.0):
    """rotation (3x3) and Sim3 scale given to every pose the entry points build from a translation; None = identity"""
    if R is None:
        lib().ref_set_test_transform(None, float(sim3_scale))
    else:
        R = np.ascontiguousarray(R, np.float32).reshape(9)
        lib().ref_set_test_transform(_p(R), float(sim3_scale))


def extract(img, nfeatures=2000, scale=1.2, nlevels=8, ini_th=20, min_th=7):
    """the reference's ORBextractor::Compute (orb_extractor.cpp:985-1049)"""
    assert img.dtype == np.uint8 and img.ndim == 2 and img.strides[1] == 1
    cap = nfeatures * 4 + 4096
    k, d = np.zeros(cap, KP_DTYPE), np.zeros((cap, 32), np.uint8)
    n = lib().ref_orb_extract(nfeatures, scale, nlevels, ini_th, min_th, img.ctypes.data, img.shape[1], img.shape[0], img.strides[0],
                              _p(k), _p(d), cap)
    assert n >= 0
    return k[:n].copy(), d[:n].copy()


def tables(nfeatures=2000, scale=1.2, nlevels=8):
    out = [np.zeros(nlevels, np.float32) for _ in range(4)]
    lib().ref_orb_tables(nfeatures, scale, nlevels, *[_p(x) for x in out])
    return dict(zip(("scale", "inv_scale", "sigma2", "inv_sigma2"), out))


class Vocabulary:
    """the reference's OrbVocabulary (DBoW2 TemplatedVocabulary<FORB>) loaded with loadFromTextFile"""

    def __init__(self, path):
        self.h = lib().ref_voc_load_text(str(path).encode())
        assert self.h, "loadFromTextFile failed"

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_voc_destroy(self.h)
            self.h = None

    def transform(self, desc, levelsup=4):
        d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(d)
        bw, bv = np.zeros(n, np.uint32), np.zeros(n, np.float64)
        fn, fs, fi = np.zeros(n, np.uint32), np.zeros(n + 1, np.int32), np.zeros(n, np.uint32)
        nb, nf = C.c_int(), C.c_int()
        lib().ref_voc_transform(self.h, n, _p(d), int(levelsup), _p(bw), _p(bv), C.byref(nb), _p(fn), _p(fs), _p(fi), C.byref(nf))
        nb, nf = nb.value, nf.value
        return dict(bow=(bw[:nb].copy(), bv[:nb].copy()), fv=(fn[:nf].copy(), fs[:nf + 1].copy(), fi[:fs[nf]].copy()))


class Frame:
    """the reference's Frame built by its own stereo / monocular constructor (frame.cpp:61-111, 163-205)"""

    def __init__(self, left, right=None, nfeatures=2000, params=(1.2, 8, 20, 7), fx=718.856, fy=718.856, cx=607.1928, cy=185.2157,
                 bf=386.1448, th_depth=35.0):
        left = np.ascontiguousarray(left, np.uint8)
        h, w = left.shape
        sf, nl, ini, mn = params
        if right is None:
            self.h = lib().ref_frame_mono(left.ctypes.data, w, h, nfeatures, sf, nl, ini, mn, fx, fy, cx, cy, bf, th_depth)
        else:
            right = np.ascontiguousarray(right, np.uint8)
            self.h = lib().ref_frame_stereo(left.ctypes.data, right.ctypes.data, w, h, nfeatures, sf, nl, ini, mn, fx, fy, cx, cy, bf,
                                            th_depth)
        n, nl_ = lib().ref_frame_n(self.h), lib().ref_frame_levels(self.h)
        self.n = n
        self.kps, self.kps_un = np.zeros(n, KP_DTYPE), np.zeros(n, KP_DTYPE)
        self.desc = np.zeros((n, 32), np.uint8)
        self.u_right, self.depth = np.full(n, -1, np.float32), np.full(n, -1, np.float32)
        self.bounds, self.scale, self.misc = np.zeros(4, np.float32), np.zeros(nl_, np.float32), np.zeros(2, np.float32)
        lib().ref_frame_get(self.h, _p(self.kps), _p(self.kps_un), _p(self.desc), _p(self.u_right) if right is not None else None,
                            _p(self.depth) if right is not None else None, _p(self.bounds), _p(self.scale), _p(self.misc))
        if right is not None:
            nr = lib().ref_frame_n_right(self.h)
            self.kps_right, self.desc_right = np.zeros(nr, KP_DTYPE), np.zeros((nr, 32), np.uint8)
            lib().ref_frame_get_right(self.h, _p(self.kps_right), _p(self.desc_right))

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_frame_destroy(self.h)
            self.h = None

    def features_in_area(self, x, y, r, min_level=-1, max_level=-1):
        out = np.zeros(self.n + 1, np.int32)
        n = lib().ref_features_in_area(self.h, x, y, r, min_level, max_level, _p(out), len(out))
        return out[:n].copy()


def _a(x, t):
    return np.ascontiguousarray(x, t)


def search_for_initialization(F1, F2, prev_matched, window, nnratio, check_ori):
    pm = _a(prev_matched, np.float32).copy()
    m12 = np.zeros(F1.n, np.int32)
    n = lib().ref_search_for_initialization(F1.h, F2.h, _p(pm), _p(m12), window, nnratio, int(check_ori))
    return n, m12, pm


def search_by_projection_mappoints(F, valid, px, py, pxr, level, viewcos, mp_desc, has_obs, occupied, th, nnratio):
    assigned = np.zeros(F.n, np.int32)
    args = [_a(valid, np.uint8), _a(px, np.float32), _a(py, np.float32), _a(pxr, np.float32), _a(level, np.int32),
            _a(viewcos, np.float32), _a(mp_desc, np.uint8), _a(has_obs, np.uint8), _a(occupied, np.uint8)]
    n = lib().ref_search_by_projection_mappoints(F.h, len(args[0]), *[_p(x) for x in args], int(th), nnratio, _p(assigned))
    return n, assigned


def search_by_projection_lastframe(Cur, Last, valid, world_pos, mp_desc, has_obs, last_t, occupied, th, mono, check_ori):
    assigned = np.zeros(Cur.n, np.int32)
    n = lib().ref_search_by_projection_lastframe(Cur.h, Last.h, _p(_a(valid, np.uint8)), _p(_a(world_pos, np.float32)),
                                                 _p(_a(mp_desc, np.uint8)), _p(_a(has_obs, np.uint8)), _p(_a(last_t, np.float32)),
                                                 _p(_a(occupied, np.uint8)), th, int(mono), int(check_ori), _p(assigned))
    return n, assigned


def set_vocabulary(voc):
    """Frames constructed from now on carry this vocabulary and run Frame::ComputeBoW (None: no vocabulary)"""
    lib().ref_set_vocabulary(voc.h if voc is not None else None)


def frame_bow(F):
    n = F.n
    bw, bv = np.zeros(n, np.uint32), np.zeros(n, np.float64)
    fn, fs, fi = np.zeros(n, np.uint32), np.zeros(n + 1, np.int32), np.zeros(n, np.uint32)
    nb, nf = C.c_int(), C.c_int()
    lib().ref_frame_get_bow(F.h, _p(bw), _p(bv), C.byref(nb), _p(fn), _p(fs), _p(fi), C.byref(nf))
    nb, nf = nb.value, nf.value
    return dict(bow=(bw[:nb].copy(), bv[:nb].copy()), fv=(fn[:nf].copy(), fs[:nf + 1].copy(), fi[:fs[nf]].copy()))


class KeyFrame:
    """the reference's KeyFrame built from a reference Frame; valid[i] gives keypoint i a map point, bad[i] flags it bad"""

    def __init__(self, F, valid=None, bad=None, translation=(0.0, 0.0, 0.0)):
        self.F = F
        v = None if valid is None else _a(valid, np.uint8)
        b = None if bad is None else _a(bad, np.uint8)
        self.h = lib().ref_keyframe_create(F.h, _p(v), _p(b), _p(_a(translation, np.float32)))

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_keyframe_destroy(self.h)
            self.h = None


def search_by_bow_kf_f(KF, F, nnratio, check_ori):
    out = np.zeros(F.n, np.int32)
    n = lib().ref_search_by_bow_kf_f(KF.h, F.h, nnratio, int(check_ori), _p(out))
    return n, out


def search_by_bow_kf_kf(KF1, KF2, nnratio, check_ori):
    out = np.zeros(KF1.F.n, np.int32)
    n = lib().ref_search_by_bow_kf_kf(KF1.h, KF2.h, nnratio, int(check_ori), _p(out))
    return n, out


def search_for_triangulation(KF1, KF2, F12, only_stereo, check_ori):
    out = np.zeros(KF1.F.n, np.int32)
    F = _a(np.asarray(F12, np.float32).reshape(9), np.float32)
    n = lib().ref_search_for_triangulation(KF1.h, KF2.h, _p(F), int(only_stereo), 0.6, int(check_ori), _p(out))
    return n, out


def is_in_frustum(F, world, idx, translation, viewing_cos_limit=0.5):
    w = _a(world, np.float32).reshape(-1, 3)
    n = len(w)
    nrm, mn, mx, ow = np.zeros((n, 3), np.float32), np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros(3, np.float32)
    out = dict(in_view=np.zeros(n, np.uint8), proj_x=np.zeros(n, np.float32), proj_y=np.zeros(n, np.float32),
               proj_xr=np.zeros(n, np.float32), level=np.zeros(n, np.int32), view_cos=np.zeros(n, np.float32))
    cnt = lib().ref_is_in_frustum(F.h, n, _p(w), _p(_a(idx, np.int32)), _p(_a(translation, np.float32)), viewing_cos_limit, _p(nrm), _p(mn),
                                  _p(mx), _p(ow), _p(out["in_view"]), _p(out["proj_x"]), _p(out["proj_y"]), _p(out["proj_xr"]),
                                  _p(out["level"]), _p(out["view_cos"]))
    return cnt, out, dict(normal=nrm, min_dist=mn, max_dist=mx, Ow=ow)


def _held(n):
    return np.zeros((n, 3), np.float32), np.zeros(n, np.float32), np.zeros(n, np.float32)


def search_by_projection_sim3(KF, world, idx, desc, bad, matched_in, t, th):
    w = _a(world, np.float32).reshape(-1, 3)
    nrm, mn, mx = _held(len(w))
    out = np.zeros(KF.F.n, np.int32)
    n = lib().ref_search_by_projection_sim3(KF.h, len(w), _p(w), _p(_a(idx, np.int32)), _p(_a(desc, np.uint8)), _p(_a(bad, np.uint8)),
                                            _p(_a(matched_in, np.uint8)), _p(_a(t, np.float32)), int(th), _p(nrm), _p(mn), _p(mx), _p(out))
    return n, out, dict(normal=nrm, min_dist=mn, max_dist=mx)


def fuse_sim3(KF, world, idx, desc, bad, t, th):
    w = _a(world, np.float32).reshape(-1, 3)
    nrm, mn, mx = _held(len(w))
    out = np.zeros(len(w), np.int32)
    n = lib().ref_fuse_sim3(KF.h, len(w), _p(w), _p(_a(idx, np.int32)), _p(_a(desc, np.uint8)), _p(_a(bad, np.uint8)),
                            _p(_a(t, np.float32)), float(th), _p(nrm), _p(mn), _p(mx), _p(out))
    return n, out, dict(normal=nrm, min_dist=mn, max_dist=mx)


class KeyFrameAt(KeyFrame):
    """KeyFrame whose keypoint i holds a map point at world[i] (where valid[i]) with descriptor desc[i]"""

    def __init__(self, F, valid, world, desc, translation=(0.0, 0.0, 0.0)):
        self.F = F
        nrm, mn, mx = _held(F.n)
        self.h = lib().ref_keyframe_create_at(F.h, _p(_a(valid, np.uint8)), _p(_a(world, np.float32)), _p(_a(desc, np.uint8)),
                                              _p(_a(translation, np.float32)), _p(nrm), _p(mn), _p(mx))
        self.held = dict(normal=nrm, min_dist=mn, max_dist=mx)


def search_by_projection_keyframe(Cur, KF, already_found, cur_t, occupied, th, orb_dist, check_ori):
    out = np.zeros(Cur.n, np.int32)
    n = lib().ref_search_by_projection_keyframe(Cur.h, KF.h, _p(_a(already_found, np.uint8)), _p(_a(cur_t, np.float32)),
                                                _p(_a(occupied, np.uint8)), float(th), int(orb_dist), int(check_ori), _p(out))
    return n, out


def search_by_sim3(KF1, KF2, t12, th, pre_idx2):
    out = np.zeros(KF1.F.n, np.int32)
    n = lib().ref_search_by_sim3(KF1.h, KF2.h, _p(_a(t12, np.float32)), float(th), _p(_a(pre_idx2, np.int32)), _p(out))
    return n, out


def fuse(KF, world, idx, desc, bad, t, th):
    w = _a(world, np.float32).reshape(-1, 3)
    nrm, mn, mx = _held(len(w))
    out = np.zeros(len(w), np.int32)
    n = lib().ref_fuse(KF.h, len(w), _p(w), _p(_a(idx, np.int32)), _p(_a(desc, np.uint8)), _p(_a(bad, np.uint8)), _p(_a(t, np.float32)),
                       float(th), _p(nrm), _p(mn), _p(mx), _p(out))
    return n, out, dict(normal=nrm, min_dist=mn, max_dist=mx)
