"""Run by tests/test_emu_asan.py in a subprocess with libasan preloaded: the CUDA sources compiled against
cuda_emu.h with -fsanitize=address, so that every out-of-bounds access to "device" memory (calloc'ed),
dynamic shared memory or a static __shared__ array aborts.  compute-sanitizer is closed on the GPU pool;
this is the bounds check we can run."""
import sys
import numpy as np

lib_path, root = sys.argv[1], sys.argv[2]
sys.path.insert(0, root)
sys.path.insert(0, root + "/tests")
import oracle_lib as O
import parity_common as P
from slam_framework_b200 import orbfe, synth

L = orbfe.load(lib_path, _test_emulation=True)
rng = np.random.default_rng(4)
P.check_extract(L, synth.frame(120, 400, seed=0), nfeatures=500)
P.check_extract(L, synth.frame(97, 131, seed=1), nfeatures=300)
P.check_extract(L, rng.integers(0, 256, (150, 420), dtype=np.uint8)[5:140, 7:400], nfeatures=1000)  # dense, strided
l, r = synth.stereo_pair(188, 620, seed=1)
P.check_stereo(L, l, r, nfeatures=1000)
P.check_batch_stereo(L, [synth.stereo_pair(100, 320, seed=s) for s in range(8)], nfeatures=300)
img = synth.frame(240, 800, seed=2)
kps, desc = O.Extractor(1500).extract(img)
scale = O.Extractor(1500).tables()["scale"]
ur = np.where(rng.uniform(0, 1, len(kps)) < 0.6, kps["x"] - rng.uniform(1, 60, len(kps)), -1).astype(np.float32)
P.check_search_by_projection_mappoints(L, kps, desc, scale, 800, 240, 3000, seed=5, u_right=ur)
P.check_search_by_projection_lastframe(L, kps, desc, scale, 800, 240, seed=7, u_right=ur)
a, b = synth.shifted_frame(3, 200, 640, dx=8, dy=4)
P.check_search_for_initialization(L, a, b, lambda im, nf: O.Extractor(nf).extract(im), nfeatures=1500)
# N1: the remaining OrbMatcher searches; N3: vocabulary transform
ka, da = O.Extractor(1000).extract(a)
kb, db = O.Extractor(1000).extract(b)
P.check_search_by_bow(L, kb, db, ka, da, scale, 640, 200, seed=3)
P.check_search_by_projection_sim3(L, ka, da, scale, 640, 200, 2000, seed=11)
P.check_search_by_projection_keyframe(L, ka, da, scale, 640, 200, 1200, seed=12)
P.check_fuse(L, ka, da, scale, 640, 200, 2000, seed=13, u_right=np.where(rng.uniform(0, 1, len(ka)) < 0.6, ka["x"] - 20, -1).astype(np.float32))
P.check_search_by_sim3(L, ka, da, kb, db, scale, 640, 200, seed=14, shift=(8.0, 4.0))
P.check_search_by_bow_keyframes(L, ka, da, kb, db, scale, 640, 200, seed=15)
P.check_search_for_triangulation(L, ka, da, kb, db, scale, 640, 200, seed=16)
P.check_bow_transform(L, da, seed=21, k=10, L=3)
P.check_bow_transform(L, da[:257], seed=22, k=4, L=5)
# N2: the Frame tail
P.check_undistort_keypoints(L, ka, seed=31)
P.check_is_in_frustum(L, 5000, seed=32)
P.check_search_local_points(L, ka, da, scale, 640, 200, seed=42, n_extra=700)
l2, r2 = synth.stereo_pair(120, 400, seed=8)
P.check_frame_from_extractor(L, l2, r2, nfeatures=400, seed=52)
P.check_empty_inputs(L, ka, da, scale)
print("ASAN-RUN-OK")
