import sys, time, numpy as np
sys.path.insert(0,'/root/repo')
from slam_framework_b200 import orbfe, synth
l, r = synth.stereo_pair(seed=0)
for n in (1, 2, 8):
    ex = orbfe.ORBextractor(max_images=n)
    imgs = ([l, r] * n)[:n]
    ex.upload(imgs); ex.run(n); ex.sync()
    ex.set_stage_timing(True); ex.stage_summary()
    for _ in range(20):
        ex.run(n)
        if n >= 2: ex.run_stereo(n // 2, 386.1448, 386.1448/718.856)
    st, runs = ex.stage_summary()
    print(n, {k: round(1e3*v/runs, 1) for k, v in st.items()}, "us total", round(1e3*sum(st.values())/runs,1))
    # wall-clock of the sync drop-in call
    t=[]
    for _ in range(30):
        t0=time.perf_counter(); ex.Compute(l) if n==1 else None; t.append(time.perf_counter()-t0)
    if n==1: print("Compute wall p50 us", round(1e6*np.median(t),1))
