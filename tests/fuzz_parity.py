#!/usr/bin/env python
"""Randomised parity sweep: CUDA path (or the emulated build) against the oracle on image content chosen to stress
the order-dependent parts of the path -- exact score ties (checkerboards, gratings, piecewise-constant shapes),
plateaus under NMS, cells that fall back to minThFAST next to cells that do not (sparse dots on flat ground),
saturated and low-contrast frames, tiny and odd-sized images, random extractor parameters.

    python tests/fuzz_parity.py [--cases N] [--seed S] [--emu | --ref] [--batch] [--max-side PX]

--ref checks the ORACLE on the same cases against the reference's own code (oracle/_ref, built by oracle/Makefile.ref; needs
/root/reference at build time): the stereo Frame constructor end to end.  Cases whose top pyramid level would be smaller than
64 px get fewer levels there (the reference divides by a zero cell count on such levels), and images are landscape (its
quad-tree starts from round(width/height) root nodes: none for a portrait image).

Prints one line per failing case with the seed that reproduces it; exit status 1 if any case failed.
tests/test_gpu_parity.py::test_fuzz_* and tests/test_emu_parity.py::test_fuzz_* run a bounded slice of it."""
import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import parity_common as P  # noqa: E402
from slam_framework_b200 import orbfe, synth  # noqa: E402

KINDS = ("texture", "checker", "grating", "rects", "dots", "blobs", "lowcontrast", "saturated", "mix", "noise")


def _box_blur(a, r):
    if r <= 0:
        return a
    k = 2 * r + 1
    c = np.cumsum(np.pad(a, ((0, 0), (r + 1, r)), mode="edge"), axis=1)
    a = (c[:, k:] - c[:, :-k]) / k
    c = np.cumsum(np.pad(a, ((r + 1, r), (0, 0)), mode="edge"), axis=0)
    return (c[k:, :] - c[:-k, :]) / k


def content(kind, h, w, rng):
    """float image (h, w) in [0, 255]"""
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    if kind == "texture":
        return synth.frame(h, w, seed=int(rng.integers(1 << 30))).astype(np.float32)
    if kind == "checker":
        p = int(rng.integers(3, 24))
        ox, oy = rng.integers(0, p, 2)
        lo, hi = sorted(rng.integers(0, 256, 2))
        return np.where((((xx + ox) // p) + ((yy + oy) // p)) % 2 == 0, lo, hi).astype(np.float32)
    if kind == "grating":
        th = rng.uniform(0, np.pi)
        per = rng.uniform(4, 30)
        a = 127.5 + rng.uniform(20, 127) * np.sign(np.sin((xx * np.cos(th) + yy * np.sin(th)) * 2 * np.pi / per))
        b = 127.5 + rng.uniform(20, 127) * np.sign(np.sin((-xx * np.sin(th) + yy * np.cos(th)) * 2 * np.pi / (per * rng.uniform(0.5, 2))))
        return (a + b) / 2
    if kind == "rects":
        img = np.full((h, w), float(rng.integers(0, 256)), np.float32)
        for _ in range(int(rng.integers(5, 120))):
            x0, y0 = int(rng.integers(0, w)), int(rng.integers(0, h))
            img[y0:y0 + int(rng.integers(1, 60)), x0:x0 + int(rng.integers(1, 60))] = float(rng.integers(0, 256))
        return img
    if kind == "dots":
        img = np.full((h, w), float(rng.integers(40, 200)), np.float32)
        n = int(rng.integers(1, 80))
        xs, ys = rng.integers(0, w, n), rng.integers(0, h, n)
        for x, y in zip(xs, ys):
            r = int(rng.integers(1, 4))
            img[max(0, y - r):y + r, max(0, x - r):x + r] += float(rng.integers(-120, 120))
        return img
    if kind == "blobs":
        return _box_blur(rng.uniform(0, 255, (h, w)).astype(np.float32), int(rng.integers(1, 5))) * 3 - 255
    if kind == "lowcontrast":
        return synth.frame(h, w, seed=int(rng.integers(1 << 30))).astype(np.float32) / rng.uniform(4, 16) + rng.uniform(0, 200)
    if kind == "saturated":
        return (synth.frame(h, w, seed=int(rng.integers(1 << 30))).astype(np.float32) - 128) * rng.uniform(2, 8) + 128
    if kind == "mix":
        a = content(KINDS[int(rng.integers(0, 7))], h, w, rng)
        b = content(KINDS[int(rng.integers(0, 7))], h, w, rng)
        m = xx / max(w - 1, 1) if rng.uniform() < 0.5 else (yy > h * rng.uniform(0.2, 0.8)).astype(np.float32)
        return a * m + b * (1 - m)
    return rng.uniform(0, 255, (h, w)).astype(np.float32)


def make_case(seed, max_side=700, landscape=False):
    rng = np.random.default_rng(seed)
    kind = KINDS[int(rng.integers(len(KINDS)))]
    h = int(rng.integers(40, max(41, max_side * 2 // 3)))
    w = int(rng.integers(40, max_side))
    if landscape:  # what the reference's own code can take (see --ref)
        h, w = max(min(h, w), 100), max(h, w, 100)
    img = np.clip(np.rint(content(kind, h, w + 40, rng)), 0, 255).astype(np.uint8)
    d = int(rng.integers(0, 40))
    left, right = np.ascontiguousarray(img[:, 40:]), np.ascontiguousarray(img[:, 40 - d:w + 40 - d])
    if rng.uniform() < 0.5:  # sensor noise on the right image: descriptors differ, ties break
        right = np.clip(right.astype(np.int16) + rng.integers(-3, 4, right.shape), 0, 255).astype(np.uint8)
    nf = int(rng.choice([50, 200, 500, 1000, 2000, 4000]))
    if rng.uniform() < 0.6:
        params = (1.2, 8, 20, 7)
    else:
        ini = int(rng.integers(5, 60))
        params = (float(rng.choice([1.1, 1.2, 1.3, 1.5, 2.0])), int(rng.integers(1, 11)), ini, int(rng.integers(1, ini + 1)))
    sf, nl, ini, mn = params
    while nl > 1 and min(h, w) / sf ** (nl - 1) < 1.0:  # a level that rounds to 0 px is an error in cv::resize and in orbfe alike
        nl -= 1
    params = (sf, nl, ini, mn)
    return dict(kind=kind, left=left, right=right, nfeatures=nf, params=params)


def run_case(lib, seed, max_side=700):
    c = make_case(seed, max_side)
    P.check_extract(lib, c["left"], nfeatures=c["nfeatures"], params=c["params"])
    P.check_stereo(lib, c["left"], c["right"], nfeatures=c["nfeatures"], params=c["params"])
    return c


def run_batch_case(lib, seed, max_side=700):
    """the device-resident batch path (upload / run / run_stereo / download) on 2-5 pairs of one geometry, mixed content"""
    rng = np.random.default_rng(seed)
    h, w = int(rng.integers(80, max(81, max_side * 2 // 3))), int(rng.integers(120, max(121, max_side)))
    pairs = []
    for _ in range(int(rng.integers(2, 6))):
        img = np.clip(np.rint(content(KINDS[int(rng.integers(len(KINDS)))], h, w + 40, rng)), 0, 255).astype(np.uint8)
        d = int(rng.integers(0, 40))
        pairs.append((np.ascontiguousarray(img[:, 40:]), np.ascontiguousarray(img[:, 40 - d:w + 40 - d])))
    P.check_batch_stereo(lib, pairs, nfeatures=int(rng.choice([200, 1000, 2000])))
    return dict(kind="batch", left=pairs[0][0], nfeatures=0, params=())


def run_case_ref(seed, max_side=700):
    """oracle vs the reference's own Frame constructor; returns None when the case is skipped"""
    import oracle_lib as O
    import reference_lib as R
    c = make_case(seed, max_side, landscape=True)
    sf, nl, ini, mn = c["params"]
    h, w = c["left"].shape
    while nl > 1 and min(h, w) / sf ** (nl - 1) < 64:  # fewer levels rather than a top level the reference cannot take
        nl -= 1
    c["params"] = (sf, nl, ini, mn)
    if w < h:  # portrait: the reference's quad-tree starts from round(w/h) = 0 root nodes (UB)
        return None
    nf = c["nfeatures"]
    F = R.Frame(c["left"], c["right"], nfeatures=nf, params=c["params"])
    oL, oR = O.Extractor(nf, *c["params"]), O.Extractor(nf, *c["params"])
    okl, odl = oL.extract(c["left"])
    okr, odr = oR.extract(c["right"])
    assert len(F.kps) == len(okl) and len(F.kps_right) == len(okr), f"keypoint counts {len(F.kps)}/{len(F.kps_right)} vs oracle {len(okl)}/{len(okr)}"
    for f in okl.dtype.names:
        assert np.array_equal(F.kps[f], okl[f]), f"left {f}"
        assert np.array_equal(F.kps_right[f], okr[f]), f"right {f}"
    assert np.array_equal(F.desc, odl) and np.array_equal(F.desc_right, odr), "descriptors"
    bf, fx = np.float32(386.1448), np.float32(718.856)
    n, ur, dp = O.stereo_match(oL, oR, okl, odl, okr, odr, float(bf), float(bf / fx))
    assert np.array_equal(F.u_right, ur), "uRight"
    assert np.array_equal(F.depth, dp), "depth"
    return c


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=100)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--max-side", type=int, default=700)
    ap.add_argument("--emu", action="store_true")
    ap.add_argument("--ref", action="store_true")
    ap.add_argument("--batch", action="store_true", help="batch path: several pairs of one geometry per case")
    a = ap.parse_args()
    lib = None
    if a.ref:
        pass
    elif a.emu:
        from emu import build_emu
        lib = orbfe.load(build_emu.build(), _test_emulation=True)
    else:
        lib = orbfe.load()
    bad = skipped = 0
    for s in range(a.seed, a.seed + a.cases):
        try:
            c = run_case_ref(s, a.max_side) if a.ref else run_batch_case(lib, s, a.max_side) if a.batch else run_case(lib, s, a.max_side)
            skipped += c is None
        except AssertionError as e:
            c = make_case(s, a.max_side, landscape=a.ref)
            bad += 1
            print(f"FAIL seed={s} kind={c['kind']} shape={c['left'].shape} nf={c['nfeatures']} params={c['params']}: {str(e)[:200]}", flush=True)
        except Exception as e:  # an API error is a failure too (the oracle accepts every one of these inputs)
            c = make_case(s, a.max_side, landscape=a.ref)
            bad += 1
            print(f"ERROR seed={s} kind={c['kind']} shape={c['left'].shape} nf={c['nfeatures']} params={c['params']}: {type(e).__name__} {str(e)[:200]}", flush=True)
    print(f"fuzz: {a.cases - bad - skipped}/{a.cases - skipped} cases identical" + (f" ({skipped} skipped)" if skipped else ""))
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
