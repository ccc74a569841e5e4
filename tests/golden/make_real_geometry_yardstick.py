"""How far do the reference's OWN two CPU code paths disagree at the real model widths?

The unmodified reference (oracle/_ref, AVX2 bodies of quantize_row_q8_* / vec_dot, ggml.c:1201-1237, k_quants.c:1856-1915) and the
scalar restatement (oracle/ggml_oracle.c, bit-exact codecs, scalar summation order) evaluate the same 2-layer random models with the
Falcon-40B / 7B / 180B geometry (tests/test_real_geometry_gpu.py: GEOM, random_model).  They compute the same integers per block dot
and differ only in fp32 summation order (~1e-7), but at n_embd 8192 one eval re-quantises ~57k activation values per layer to int8:
with a quantisation step of amax/127 ~ sigma/32 and element errors of ~3e-7 sigma, P(flip) ~ 2e-5 per value, i.e. about one flipped
code per layer per token, each moving that mat-mul's outputs by ~1/(32 sqrt(K)) ~ 3e-4 of their scale.  The numbers written to
real_geometry_cpu_vs_cpu.json are the yardstick the GPU-vs-oracle comparison at these widths is read against: the "tight"
(reassociation-level) bound of the tiny models cannot hold there for ANY pair of implementations, the "loose" bound does.

Run where /root/reference exists (needs oracle/_ref):  python tests/golden/make_real_geometry_yardstick.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), ROOT]
import numpy as np  # noqa: E402
import pyoracle as po  # noqa: E402
import ggllm_cpp_b200.ggcc as ggcc  # noqa: E402
from test_real_geometry_gpu import GEOM, random_model, MODEL_SEED  # noqa: E402

out = {}
for geom, wt, ft in (("40b", po.Q4_K, 15), ("40b", po.Q3_K, 12), ("7b", po.Q4_0, 2), ("180b", po.Q4_K, 15)):
    hp = GEOM[geom]
    tensors = random_model(hp, wt, seed=MODEL_SEED[geom] + wt)
    path = "/dev/shm/yardstick_%d.ggcc" % os.getpid()
    ggcc.write_ggcc(path, hp, tensors, ftype=ft)
    ref = po.RefFalcon(path, n_ctx=64, n_batch=8)
    o = po.OrcFalcon(hp, tensors, n_ctx=64)
    rows = []
    for i in range(6):
        tok = np.array([17 + i], np.int32)
        a, b = ref.eval(tok, i, n_threads=8, n_max_real_ctx=8192), o.eval(tok, i, 8192)
        S, d = float(np.abs(b).max()), np.abs(a - b)
        rows.append({"n_past": i, "max_over_S": float(d.max() / S), "median_over_S": float(np.median(d) / S)})
    ref.close()
    os.unlink(path)
    out["%s_%s" % (geom, po.TYPE_NAMES[wt])] = rows
json.dump({"what": "unmodified reference (AVX2 CPU build) vs scalar oracle restatement, decode evals of 2-layer real-geometry models", "cases": out},
          open(os.path.join(HERE, "real_geometry_cpu_vs_cpu.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
