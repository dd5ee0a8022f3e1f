"""PINS the oracle against the reference's OWN code: oracle/_ref/libslam_ref.so is src/orb_features/orb_extractor.cpp and
third_party/DBoW2 compiled UNMODIFIED from /root/reference (oracle/Makefile.ref) against a stand-in for the OpenCV subset they
use (oracle/cvstub; its image primitives are the oracle's own, pinned to cv2 by test_oracle_primitives.py).  So these tests
check the part cv2 cannot: the reference's control logic (FAST grid loop with the threshold fallback, DistributeOctTree /
DivideNode, IC_Angle, computeOrbDescriptor, keypoint rescaling; DBoW2 tree walk, BowVector / FeatureVector assembly and
normalisation) run as written.  The one definition involved: the quad-tree's pointer tie-break is evaluated on a monotonic
heap (oracle/ref_wrap.cpp), i.e. "later-created node sorts higher", which is how DESIGN.md defines it."""
import numpy as np
import pytest

import oracle_lib as O
import parity_common as P
import reference_lib as R
from slam_framework_b200 import synth

pytestmark = pytest.mark.skipif(not R.available(), reason="neither /root/reference nor a prebuilt oracle/_ref is present")

KITTI = (1.2, 8, 20, 7)


def assert_same_extraction(img, nfeatures, params=KITTI, what=""):
    k, d = R.extract(img, nfeatures, *params)
    ok, od = O.Extractor(nfeatures, *params).extract(img)
    assert len(k) == len(ok), f"{what}: reference {len(k)} keypoints, oracle {len(ok)}"
    for f in k.dtype.names:
        assert np.array_equal(k[f], ok[f]), f"{what}: keypoint field {f} differs from the reference"
    assert np.array_equal(d, od), f"{what}: {(d != od).any(1).sum()} descriptors differ from the reference"
    return len(k)


@pytest.mark.parametrize("seed", [0, 7, 21])
def test_extractor_config1_kitti_frames(seed):
    assert assert_same_extraction(synth.frame(seed=seed), 2000, what=f"KITTI frame {seed}") > 1900


def test_extractor_stereo_pair_and_mono_4000():
    l, r = synth.stereo_pair(seed=3)
    assert_same_extraction(l, 2000, what="left")
    assert_same_extraction(r, 2000, what="right")
    assert assert_same_extraction(synth.frame(seed=5), 4000, what="mono 4000") > 3900


def test_extractor_other_geometries_and_thresholds():
    rng = np.random.default_rng(0)
    assert_same_extraction(synth.frame(200, 640, seed=1), 500, what="640x200")
    assert_same_extraction(rng.integers(0, 256, (300, 400), dtype=np.uint8), 1000, what="white noise")       # every cell saturated
    assert_same_extraction(np.full((300, 400), 77, np.uint8), 500, what="flat")                               # no keypoints at all
    assert_same_extraction(synth.frame(400, 1300, seed=11)[10:386, 20:1261], 2000, what="strided view")
    assert_same_extraction(synth.frame(seed=9), 1500, (1.5, 5, 20, 7), what="scale 1.5, 5 levels")
    assert_same_extraction(synth.frame(seed=12), 1000, (2.0, 4, 20, 7), what="scale 2.0 (INTER_AREA path), 4 levels")
    assert_same_extraction(synth.frame(seed=10), 1000, (1.2, 8, 40, 12), what="thresholds 40/12")
    low = (synth.frame(seed=13).astype(np.float32) * 0.15 + 100).astype(np.uint8)                             # minThFAST fallback everywhere
    assert_same_extraction(low, 2000, what="low contrast")


def test_extractor_config5_1080p_8000():
    assert assert_same_extraction(synth.frame(1080, 1920, seed=3), 8000, what="1080p") > 7900


def test_extractor_tables():
    for nf, sf, nl in ((2000, 1.2, 8), (4000, 1.2, 8), (1000, 1.5, 5)):
        t, o = R.tables(nf, sf, nl), O.Extractor(nf, sf, nl, 20, 7).tables()
        for key in ("scale", "inv_scale", "sigma2", "inv_sigma2"):
            assert np.array_equal(t[key], o[key]), key


def test_vocabulary_transform_against_dbow2(tmp_path):
    """DBoW2's own loadFromTextFile + transform (TemplatedVocabulary.h:1335-1422, 1124-1250) vs the oracle restatement"""
    kps, desc = O.Extractor(1500).extract(synth.frame(240, 800, seed=4))
    rng = np.random.default_rng(8)
    for (k, L, scoring, weighting, prune) in ((10, 3, 0, 0, 0.0), (10, 3, 1, 0, 0.0), (10, 3, 5, 0, 0.0), (10, 3, 0, 1, 0.0), (10, 3, 0, 2, 0.0),
                                              (10, 3, 1, 3, 0.0), (4, 5, 0, 0, 0.0), (10, 3, 0, 0, 0.08), (4, 5, 1, 1, 0.1)):
        arrays = P.synth_vocabulary(rng, k, L, prune_frac=prune, seed_desc=desc[0].copy())
        path = tmp_path / f"voc_{k}_{L}_{scoring}_{weighting}_{prune}.txt"
        P.write_vocabulary_text(str(path), k, L, scoring, weighting, arrays)
        V, OV = R.Vocabulary(path), O.Vocabulary(k, L, scoring, weighting, *arrays)
        # a leaf above nid_level leaves DBoW2's `nid` unwritten (indeterminate): with early leaves in the tree only
        # nid_level <= 1 is compared; full-depth trees are compared at every level
        for levelsup in ((0, 1, 2, L - 1, L, L + 1) if prune == 0.0 else (L - 1, L, L + 1)):
            got, ref = OV.transform(desc, levelsup), V.transform(desc, levelsup)
            assert np.array_equal(got["bow"][0], ref["bow"][0]), (k, L, scoring, weighting, levelsup)
            assert np.array_equal(got["bow"][1], ref["bow"][1]), "BowVector values are not bit-identical to DBoW2's"
            for j in range(3):
                assert np.array_equal(got["fv"][j], ref["fv"][j]), ("FeatureVector", j, levelsup)
