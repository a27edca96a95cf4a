"""dev: sp_permanent_compressed with a multi-device id -- leaves spread over the devices vs one device"""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
import superman_b200 as sp
from _compressed import banded
g = sp.device_count()
for n, leaf in ((44, 30), (48, 30)):
    a = banded(np.random.default_rng(100 + n), n, "real")
    for gpus, algo in ((1, 4), (g, 5)):
        st = sp.SpStats()
        sp.permanent_compressed(a, sparse=True, preprocessing=1, algo_id=algo, gpu_num=gpus, leaf_nov=leaf, stats=st)
        t = time.perf_counter()
        v = sp.permanent_compressed(a, sparse=True, preprocessing=1, algo_id=algo, gpu_num=gpus, leaf_nov=leaf, stats=st)
        dt = time.perf_counter() - t
        print("n=%d leaf_nov=%d gpus=%d: %.15e  %d leaves  wall %.1f ms  kernel %.1f ms  device ms %s" % (
            n, leaf, gpus, v, st.chunks, dt * 1e3, st.kernel_ms, [round(x, 1) for x in st.device_ms[:gpus]]), flush=True)
