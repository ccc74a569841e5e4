import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ggllm_cpp_b200.binding as b
b.init(0); L = b.lib()
e0, e1 = L.b200_event_create(), L.b200_event_create()
for K, M in [(8192, 9216), (8192, 32768), (32768, 8192)]:
    Ws = [b.Weight(12, K, M, seed=i + 1) for i in range(max(2, int(400e6 // (K * M * 0.5625)) + 1))]
    rng = np.random.default_rng(0)
    x, ra, rb, g, be = (b.DevBuf(src=rng.standard_normal(K).astype(np.float32)) for _ in range(5))
    xo, yd = b.DevBuf(K * 4), b.DevBuf(M * 4)
    A = b.ActQ(12, K, 1); A.quantize(x.ptr)
    def run(mode):
        for w in Ws:
            if mode == 0: L.b200_mul_mat_vec_q(w.h, A.h, yd.ptr, M, 0, None, None)
            elif mode == 1: L.b200_mul_mat_vec_fused(w.h, x.ptr, None, None, None, None, None, yd.ptr, 0)
            else: L.b200_mul_mat_vec_fused(w.h, x.ptr, ra.ptr, rb.ptr, g.ptr, be.ptr, xo.ptr, yd.ptr, 0)
    for mode in (0, 1, 2):
        if mode == 2 and K > 16384: continue
        run(mode); L.b200_synchronize()
        L.b200_event_record(e0, None)
        for _ in range(5): run(mode)
        L.b200_event_record(e1, None); L.b200_event_synchronize(e1)
        us = L.b200_event_elapsed_ms(e0, e1) / (5 * len(Ws)) * 1e3
        print(json.dumps(dict(K=K, M=M, mode=mode, us=round(us, 2), GBs=round(K * M * 0.5625 / us / 1e3, 1))), flush=True)
    for w in Ws: w.free()
