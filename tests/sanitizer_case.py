import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import parity_common as P
from slam_framework_b200 import orbfe, synth
L = orbfe.load()
P.check_extract(L, synth.frame(120, 400, seed=0), nfeatures=500)
P.check_extract(L, synth.frame(97, 131, seed=1), nfeatures=300)
l, r = synth.stereo_pair(188, 620, seed=1)
print("stereo", P.check_stereo(L, l, r, nfeatures=1000))
pairs = [synth.stereo_pair(100, 320, seed=s) for s in range(8)]
print("batch", P.check_batch_stereo(L, pairs, nfeatures=300))
import oracle_lib as O
img = synth.frame(240, 800, seed=2)
kps, desc = O.Extractor(1500).extract(img)
scale = O.Extractor(1500).tables()["scale"]
rng = np.random.default_rng(1)
ur = np.where(rng.uniform(0, 1, len(kps)) < 0.6, kps["x"] - rng.uniform(1, 60, len(kps)), -1).astype(np.float32)
print("mp", P.check_search_by_projection_mappoints(L, kps, desc, scale, 800, 240, 3000, seed=5, u_right=ur))
print("lf", P.check_search_by_projection_lastframe(L, kps, desc, scale, 800, 240, seed=7, u_right=ur))
a, b = synth.shifted_frame(3, 200, 640, dx=8, dy=4)
print("init", P.check_search_for_initialization(L, a, b, lambda im, nf: O.Extractor(nf).extract(im), nfeatures=1500))
print("SANITIZER-RUN-OK")
