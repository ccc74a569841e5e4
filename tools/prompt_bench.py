"""BASELINE config 3: Falcon-40B Q4_K prompt processing, n_batch = 512, 2048 synthetic tokens (4 evals at n_past 0/512/1024/1536).
usage: python tools/prompt_bench.py [layers] [40b|7b] [ggml type id]   (other models / weight types: same measurement)"""
import sys, os, json, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ggllm_cpp_b200.binding as b
import ggllm_cpp_b200.ggcc as ggcc

def main():
    layers = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    b.init(0)
    model = sys.argv[2] if len(sys.argv) > 2 else "40b"
    wtype = int(sys.argv[3]) if len(sys.argv) > 3 else 12
    hp = dict(n_vocab=65024, n_embd=8192, n_head=128, n_head_kv=8, n_layer=layers, falcon_type=40) if model == "40b" else \
         dict(n_vocab=65024, n_embd=4544, n_head=71, n_head_kv=1, n_layer=layers, falcon_type=7)
    f = b.Falcon(hp, n_ctx=2048, n_batch=512)
    f.set_random(ggcc.falcon_shapes(hp), wtype, seed=1234)
    toks = np.random.default_rng(7).integers(12, 65024, size=2048).astype(np.int32)
    f.eval(toks[:512], 0)            # warm-up
    res = []
    t0 = time.perf_counter()
    for c in range(4):
        f.eval(toks[512 * c: 512 * (c + 1)], 512 * c)
        res.append(round(f.last_ms(), 2))
    dt = time.perf_counter() - t0
    flops = 2 * sum(int(np.prod(sh)) for nm, sh in ggcc.falcon_shapes(hp).items() if len(sh) == 2 and 'word_embeddings' not in nm and 'lm_head' not in nm) * 2048
    print(json.dumps(dict(model=model, type=wtype, layers=layers, eval_ms=res, total_s=round(dt, 4), prompt_tok_s=round(2048 / dt, 1), matmul_TFLOPs=round(flops / dt / 1e12, 1), launches_last=f.last_launches())))

if __name__ == "__main__":
    main()
