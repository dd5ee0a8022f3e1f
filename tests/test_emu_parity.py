"""Kernel LOGIC parity on CPU: the same CUDA sources compiled against tests/emu/cuda_emu.h (a
TEST-ONLY emulation of the CUDA subset the kernels use) are diffed against the oracle.  This is
what lets the GPU-less container catch logic errors; the parity tests proper are test_gpu_parity.py
(-m gpu), which run the real sm_100a build."""
import numpy as np
import pytest

import parity_common as P
from emu import build_emu
from slam_framework_b200 import orbfe, synth


@pytest.fixture(scope="module")
def emu():
    return orbfe.load(build_emu.build(), _test_emulation=True)


@pytest.mark.parametrize("h,w,seed", [(120, 400, 0), (97, 131, 1), (200, 640, 2)])
def test_extract_small(emu, h, w, seed):
    P.check_extract(emu, synth.frame(h, w, seed=seed), nfeatures=500)


def test_extract_kitti_frame(emu):
    kps, _ = P.check_extract(emu, synth.frame(seed=0))
    assert 1900 < len(kps) < 2100


def test_extract_strided_and_noise(emu):
    rng = np.random.default_rng(4)
    big = rng.integers(0, 256, (150, 420), dtype=np.uint8)
    P.check_extract(emu, big[5:140, 7:400], nfeatures=1000)  # non-contiguous rows, dense texture


@pytest.mark.parametrize("pct", [1, 3, 20])
def test_fast_dense_form_when_the_candidate_queue_overflows(emu, monkeypatch, pct):
    """k_fast_cells sizes its candidate queue for occupancy, not for the worst case; a tile whose pre-test passes more pixels is
    scored and NMS-tested densely.  The test hook shrinks the queue so that ordinary frames (both thresholds, fallback cells
    beside cells with keypoints) and white noise take that form; results must not change."""
    monkeypatch.setenv("ORBFE_TEST_FAST_QUEUE_PCT", str(pct))
    P.check_extract(emu, synth.frame(200, 640, seed=2), nfeatures=500)
    P.check_extract(emu, synth.frame(seed=5))
    rng = np.random.default_rng(9)
    P.check_extract(emu, rng.integers(0, 256, (140, 400), dtype=np.uint8), nfeatures=1000)
    low = (synth.frame(160, 500, seed=3) // 8 + 100).astype(np.uint8)  # low contrast: most cells go to the minThFAST round
    P.check_extract(emu, low, nfeatures=600)


def test_extract_flat_image_yields_nothing(emu):
    ex = orbfe.ORBextractor(lib=emu)
    kps, desc = ex.Compute(np.full((100, 300), 128, np.uint8))
    assert len(kps) == 0 and desc.shape == (0, 32)
    kps, desc = ex.Compute(np.zeros((0, 0), np.uint8))  # empty image => silent return (:990-991)
    assert len(kps) == 0


def test_stereo_kitti_pair(emu):
    left, right = synth.stereo_pair(seed=0)
    assert P.check_stereo(emu, left, right) > 300


def test_batch_stereo_two_pairs(emu):
    pairs = [synth.stereo_pair(160, 500, seed=s) for s in (3, 4)]
    assert P.check_batch_stereo(emu, pairs, nfeatures=600) > 50


def test_batch_of_16_small_frames_uses_wide_descriptor_path(emu):
    """>= 16 images in flight switch k_orient_describe to 32 keypoints per warp"""
    pairs = [synth.stereo_pair(100, 320, seed=s) for s in range(8)]
    assert P.check_batch_stereo(emu, pairs, nfeatures=300) > 20


@pytest.mark.parametrize("channels,rgb", [(3, True), (3, False), (4, True), (4, False)])
def test_colour_upload_matches_cvtcolor_then_extract(emu, channels, rgb):
    """N4: colour frames converted on the device == oracle cvtColor + extraction of the gray frame"""
    import oracle_lib as O
    rng = np.random.default_rng(channels)
    gray = synth.frame(101, 323, seed=6).astype(np.int32)
    col = np.clip(gray[..., None] + rng.integers(-25, 26, (101, 323, channels)), 0, 255).astype(np.uint8)
    ex = orbfe.ORBextractor(300, lib=emu, max_images=2)
    ex.upload_color([col, col[:, ::-1].copy()], rgb=rgb)
    ex.run(2)
    b = ex.download(2, ex.make_buffers(2))
    ref = O.cvt_gray(col, rgb)
    assert np.array_equal(ex.pyramid_level(0, 0), ref)
    ok, od = O.Extractor(300).extract(ref)
    P.assert_kps_equal(b["kps"][0, :b["n"][0]], ok)
    assert np.array_equal(b["desc"][0, :b["n"][0]], od)


@pytest.mark.parametrize("nf,params", [(300, (1.2, 1, 20, 7)), (600, (1.1, 12, 20, 7)), (400, (1.5, 5, 20, 7)), (400, (2.0, 4, 20, 7)),
                                       (300, (2.5, 3, 20, 7)), (500, (1.3, 6, 12, 5)), (500, (1.2, 8, 7, 7)), (60, (1.2, 8, 40, 12))])
def test_extractor_parameter_sweep(emu, nf, params):
    """other pyramids (1 and 12 levels, scale 1.1 / 1.5, exact 2x = OpenCV's INTER_AREA path, > 2x = generic resize kernel),
    thresholds (ini == min: no fallback round) and small quotas, stage by stage"""
    P.check_extract(emu, synth.frame(200, 640, seed=int(params[0] * 10) + params[1]), nfeatures=nf, params=params)


def test_fuzz_slice_tie_heavy_content_and_random_parameters(emu):
    """a bounded slice of tests/fuzz_parity.py: checkerboards, gratings, rectangles, dots, saturated / low-contrast frames,
    odd sizes and random extractor parameters (extraction stage by stage + ComputeStereoMatches)"""
    import fuzz_parity
    # 9: a pyramid level one pixel high (the host LUT builder looped for ever); 114: checkerboard, the stereo search's tie rule
    for seed in list(range(24)) + [114]:
        fuzz_parity.run_case(emu, seed, max_side=300)
