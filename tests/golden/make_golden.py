"""Generate the committed golden vectors from the UNMODIFIED reference (oracle/_ref, built from /root/reference).

Run here (where /root/reference exists):   python tests/golden/make_golden.py
Outputs (small, committed):
  codecs.npz          for every weight type: the reference's own test vector x[i] = 0.1 + 2 cos(i) (test-quantize-fns.cpp:26-30),
                      the reference's quantised bytes, its dequantised values, its Q8 activation bytes and vec_dot result
  tiny40b_q4_K.npz    logits of falcon_eval (CPU build) for the synthetic model recipe tests/helpers.synth_model
  tiny7b_q4_0.npz
The GPU box has no /root/reference; the -m gpu tests compare against these files.
"""
import os
import sys
import tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from helpers import po, ggcc, synth_model, TINY_40B, TINY_7B  # noqa: E402


def codecs():
    r = po.ref()
    out = {}
    x = po.synth_vector(4096).reshape(4, 1024)
    a = po.synth_vector(4096, offset=1.0).reshape(4, 1024)
    out["x"], out["a"] = x, a
    for t in po.WEIGHT_TYPES:
        n = po.TYPE_NAMES[t]
        q = r.quantize(t, x)
        out[n + "_q"] = q
        out[n + "_deq"] = r.dequantize(t, q, 1024)
        aq = r.quantize_act(t, a)
        out[n + "_aq"] = aq
        out[n + "_dot"] = np.array([r.vec_dot(t, 1024, q[i], aq[i]) for i in range(4)], np.float32)
    np.savez_compressed(os.path.join(HERE, "codecs.npz"), **out)


def model(name, hp, wt, seed, n_ctx=64):
    tensors = synth_model(hp, wt, seed)
    path = os.path.join(tempfile.gettempdir(), name + ".ggcc")
    ggcc.write_ggcc(path, hp, tensors, ftype=ggcc.FTYPE_OF_TYPE[wt])
    ref = po.RefFalcon(path, n_ctx=n_ctx, n_batch=8, logits_all=True)
    prompt = np.array([11, 100, 101, 102, 103, 104, 105], np.int32)
    pl = ref.eval(prompt, 0, n_threads=4)
    dec = np.array([200, 17, 333, 42], np.int32)
    dl = np.concatenate([ref.eval(dec[i:i + 1], len(prompt) + i, n_threads=4) for i in range(len(dec))])
    ref.close()
    os.remove(path)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), prompt=prompt, prompt_logits=pl, decode_tokens=dec, decode_logits=dl,
                        wtype=wt, seed=seed, n_ctx=n_ctx, **{"hp_" + k: v for k, v in hp.items()})


if __name__ == "__main__":
    codecs()
    model("tiny40b_q4_K", TINY_40B, po.Q4_K, 1234)
    model("tiny7b_q4_0", TINY_7B, po.Q4_0, 1234)
    print("golden vectors written to", HERE)
