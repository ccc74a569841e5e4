"""Decode tok/s as a function of the context position (run on the GPU box).  usage: python tools/ctx_decode.py <n_ctx> [40b|7b]"""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ggllm_cpp_b200.binding as b
import ggllm_cpp_b200.ggcc as ggcc

n_ctx = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
b.init(0); L = b.lib()
model = sys.argv[2] if len(sys.argv) > 2 else "40b"
hp = dict(n_vocab=65024, n_embd=8192, n_head=128, n_head_kv=8, n_layer=60, falcon_type=40) if model == "40b" else \
     dict(n_vocab=65024, n_embd=4544, n_head=71, n_head_kv=1, n_layer=32, falcon_type=7)
f = b.Falcon(hp, n_ctx=n_ctx, n_batch=1)
f.set_random(ggcc.falcon_shapes(hp), 12 if model == "40b" else 2, seed=1234)
tok = b.DevBuf(src=np.array([1234], np.int32))
e0, e1 = L.b200_event_create(), L.b200_event_create()
for p in range(8): f.decode_dev(tok.ptr, p, 0)
for start in [8, 128, 512, 1024, 2040, 4088, 8184]:
    if start + 8 > n_ctx: break
    f.decode_dev(tok.ptr, start, 0)                             # (builds the decode graph of this tier outside the timed region)
    L.b200_stream_synchronize(f.stream())
    L.b200_event_record(e0, f.stream())
    for i in range(8): f.decode_dev(tok.ptr, start + i, 0)      # (the KV slots in between hold zeros: timing only)
    L.b200_event_record(e1, f.stream()); L.b200_event_synchronize(e1)
    ms = L.b200_event_elapsed_ms(e0, e1) / 8
    print(json.dumps(dict(model=model, n_past=start, ms_per_tok=round(ms, 3), tok_s=round(1e3 / ms, 1))), flush=True)
