"""Device timeline of one decode step of the 40B Q4_K model (globaltimer stamps written by the kernels themselves)."""
import sys, os, json, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ggllm_cpp_b200.binding as b
import ggllm_cpp_b200.ggcc as ggcc

layers = int(sys.argv[1]) if len(sys.argv) > 1 else 8
model = sys.argv[2] if len(sys.argv) > 2 else "40b"
b.init(0); L = b.lib()
L.b200_trace_enable.argtypes = [C.c_int]; L.b200_trace_reset.argtypes = [C.c_void_p]; L.b200_trace_dump.argtypes = [C.c_void_p, C.c_void_p, C.c_int]; L.b200_trace_dump.restype = C.c_int
hp = dict(n_vocab=65024, n_embd=8192, n_head=128, n_head_kv=8, n_layer=layers, falcon_type=40) if model == "40b" else \
     dict(n_vocab=65024, n_embd=14848, n_head=232, n_head_kv=8, n_layer=layers, falcon_type=40) if model == "180b" else \
     dict(n_vocab=65024, n_embd=4544, n_head=71, n_head_kv=1, n_layer=layers, falcon_type=7)
n_ctx = int(os.environ.get("N_CTX", "2048")); p0 = int(os.environ.get("N_PAST", "6")) - 6          # trace the step at position N_PAST
f = b.Falcon(hp, n_ctx=n_ctx, n_batch=1)
f.set_random(ggcc.falcon_shapes(hp), 2 if model == "7b" else 12, seed=1234)
L.b200_trace_enable(2048)                      # slots are claimed while the decode graph is captured
tok = b.DevBuf(src=np.array([1234], np.int32))
for pos in range(p0, p0 + 6):
    f.decode_dev(tok.ptr, pos, 129)
L.b200_stream_synchronize(f.stream())
L.b200_trace_reset(None)
f.decode_dev(tok.ptr, p0 + 6, 129)
L.b200_stream_synchronize(f.stream())
out = np.zeros((2048, 2), np.uint64); names = C.create_string_buffer(24 * 2048)
n = L.b200_trace_dump(out.ctypes.data_as(C.c_void_p), names, 2048)
ev = [(names.raw[24 * i: 24 * i + 24].split(b"\0")[0].decode(), int(out[i, 0]), int(out[i, 1])) for i in range(n) if out[i, 1] > 0]
ev = [e for e in ev if e[1] != 2 ** 64 - 1]
t0 = min(e[1] for e in ev)
ev.sort(key=lambda e: e[1])
# the eager warm-up pass and the capture pass both claimed slots; only the graph's own slots are stamped after reset
print("events", len(ev))
for name, s, e in ev[: 8 * 4 + 2]:
    print("%-12s start %8.2f us  end %8.2f us  dur %7.2f" % (name, (s - t0) / 1e3, (e - t0) / 1e3, (e - s) / 1e3))
print("total span us", (max(e[2] for e in ev) - t0) / 1e3, "per layer", (max(e[2] for e in ev) - t0) / 1e3 / layers)
