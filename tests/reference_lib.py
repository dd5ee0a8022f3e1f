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
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(vp) if a is not None else None


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
