"""Shared parity checks: the C-ABI library (product build on the GPU, or the emulated TEST build on
CPU) against the CPU oracle on the same seeded inputs.  Bit-exact for keypoints (coordinates,
octave, response, size), descriptors, match indices; angles / sub-pixel disparities are compared
bit-exact too (the stated tolerance of BASELINE.json is 1e-4; we meet 0)."""
import numpy as np

import oracle_lib as O
from slam_framework_b200 import orbfe

KITTI = dict(bf=386.1448, fx=718.856)
KP_FIELDS = ("x", "y", "size", "angle", "response", "octave", "class_id")
ANGLE_TOL = 1e-4


def assert_kps_equal(got, ref, what=""):
    assert len(got) == len(ref), f"{what}: {len(got)} keypoints vs oracle {len(ref)}"
    for f in KP_FIELDS:
        if f == "angle":
            assert np.allclose(got[f], ref[f], rtol=0, atol=ANGLE_TOL), f"{what}: angle beyond {ANGLE_TOL}"
        assert np.array_equal(got[f], ref[f]), f"{what}: field {f} differs"


def check_extract(lib, img, nfeatures=2000, stages=True, params=(1.2, 8, 20, 7)):
    ex = orbfe.ORBextractor(nfeatures, *params, lib=lib)
    oe = O.Extractor(nfeatures, *params)
    kps, desc = ex.Compute(img)
    okps, odesc = oe.extract(img)
    if stages:
        for l in range(params[1]):
            assert np.array_equal(ex.pyramid_level(l), oe.pyramid_level(l)), f"pyramid level {l}"
            c, oc = ex.debug_candidates(l), oe.candidates(l)
            assert len(c) == len(oc), f"FAST candidates level {l}: {len(c)} vs {len(oc)}"
            for f in ("x", "y", "response"):
                assert np.array_equal(c[f], oc[f]), f"FAST candidates level {l} field {f}"
            k, ok = ex.debug_level_keypoints(l), oe.level_keypoints(l)
            assert len(k) == len(ok), f"quad-tree level {l}: {len(k)} vs {len(ok)}"
            assert np.array_equal(k["x"] + 16, ok["x"]) and np.array_equal(k["y"] + 16, ok["y"]), f"quad-tree level {l}"
            if len(ok):
                assert np.array_equal(ex.debug_blurred(l), oe.blurred(l)), f"blur level {l}"
    assert_kps_equal(kps, okps, "extract")
    assert np.array_equal(desc, odesc), f"descriptors differ in {(desc != odesc).any(1).sum()} rows"
    ex.close()
    return kps, desc


def check_stereo(lib, left, right, nfeatures=2000, bf=KITTI["bf"], fx=KITTI["fx"]):
    eL, eR = orbfe.ORBextractor(nfeatures, lib=lib), orbfe.ORBextractor(nfeatures, lib=lib)
    oL, oR = O.Extractor(nfeatures), O.Extractor(nfeatures)
    kl, dl = eL.Compute(left)
    kr, dr = eR.Compute(right)
    okl, odl = oL.extract(left)
    okr, odr = oR.extract(right)
    assert_kps_equal(kl, okl, "left")
    assert_kps_equal(kr, okr, "right")
    assert np.array_equal(dl, odl) and np.array_equal(dr, odr)
    n, ur, dp = orbfe.ComputeStereoMatches(eL, eR, kl, dl, kr, dr, bf, bf / fx)
    on, our, odp = O.stereo_match(oL, oR, okl, odl, okr, odr, bf, bf / fx)
    assert n == on, f"stereo matches {n} vs oracle {on}"
    assert np.array_equal(ur >= 0, our >= 0), "matched set differs"
    assert np.allclose(ur, our, rtol=0, atol=1e-4) and np.allclose(dp, odp, rtol=1e-6, atol=1e-4)
    assert np.array_equal(ur, our) and np.array_equal(dp, odp), "sub-pixel results not bit-identical"
    eL.close(); eR.close()
    return n


def check_batch_stereo(lib, pairs, nfeatures=2000, bf=KITTI["bf"], fx=KITTI["fx"]):
    """device-resident batch path == per-frame oracle, slot by slot."""
    n = 2 * len(pairs)
    ex = orbfe.ORBextractor(nfeatures, lib=lib, max_images=n)
    imgs = [im for p in pairs for im in p]
    ex.upload(imgs)
    ex.run(n)
    ex.run_stereo(len(pairs), bf, bf / fx)
    buf = ex.download(n, ex.make_buffers(n, stereo=True))
    total = 0
    for p, (l, r) in enumerate(pairs):
        oL, oR = O.Extractor(nfeatures), O.Extractor(nfeatures)
        okl, odl = oL.extract(l)
        okr, odr = oR.extract(r)
        nl, nr = buf["n"][2 * p], buf["n"][2 * p + 1]
        assert_kps_equal(buf["kps"][2 * p, :nl], okl, f"pair {p} left")
        assert_kps_equal(buf["kps"][2 * p + 1, :nr], okr, f"pair {p} right")
        assert np.array_equal(buf["desc"][2 * p, :nl], odl) and np.array_equal(buf["desc"][2 * p + 1, :nr], odr)
        on, our, odp = O.stereo_match(oL, oR, okl, odl, okr, odr, bf, bf / fx)
        assert np.array_equal(buf["ur"][2 * p, :nl], our), f"pair {p} uR"
        assert np.array_equal(buf["depth"][2 * p, :nl], odp), f"pair {p} depth"
        total += on
    ex.close()
    return total
