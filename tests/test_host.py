"""CPU tests of the host-side logic: GGCC reader/writer, layer-range partition, and the 2-rank pipeline hand-off
(gloo, world_size 2) with the oracle standing in for the device stages."""
import os
import sys
import numpy as np
import pytest
import pyoracle as po
from helpers import TINY_40B, TINY_7B, synth_model, ggcc


def test_ggcc_roundtrip(tmp_path):
    hp = dict(TINY_7B)
    tensors = synth_model(hp, po.Q4_0, seed=1)
    p = str(tmp_path / "a.ggcc")
    ggcc.write_ggcc(p, hp, tensors, ftype=2)
    hp2, t2 = ggcc.read_ggcc(p)
    assert {k: hp2[k] for k in hp} == hp and hp2["ftype"] == 2
    assert list(t2) == list(tensors)
    for k in tensors:
        assert t2[k][0] == tensors[k][0] and tuple(t2[k][1]) == tuple(tensors[k][1])
        assert np.array_equal(np.asarray(t2[k][2]), np.ascontiguousarray(tensors[k][2]).view(np.uint8).reshape(-1))
        assert (t2[k][2].ctypes.data - t2[k][2].base.ctypes.data if hasattr(t2[k][2], "base") and t2[k][2].base is not None else 0) % 32 == 0


def test_ggcc_shapes_are_the_loaders():
    s = ggcc.falcon_shapes(dict(n_vocab=65024, n_embd=8192, n_head=128, n_head_kv=8, n_layer=60, falcon_type=40))
    assert s["transformer.h.59.self_attention.query_key_value.weight"] == (8192, 9216)
    assert s["transformer.h.0.mlp.dense_4h_to_h.weight"] == (32768, 8192)
    n = sum(int(np.prod(v)) for k, v in s.items() if len(v) == 2 and "word_embeddings" not in k)
    assert n == 41301311488                       # W_elems(40B) of SURVEY.md section 8a
    assert ggcc.tensor_nbytes(12, (8192, 9216)) == 8192 * 9216 // 256 * 144


def test_layer_ranges_cover_the_model():
    from ggllm_cpp_b200.binding import layer_range
    for L in (60, 80, 32, 7):
        for world in (1, 2, 4, 8):
            r = [layer_range(L, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == L
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def _pipeline_worker(rank, world, port, q):
    import torch.distributed as dist
    import torch
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ggllm_cpp_b200.binding import layer_range
    hp = dict(TINY_40B, n_layer=5)
    tensors = synth_model(hp, po.Q4_K, seed=8)
    lf, ll = layer_range(hp["n_layer"], rank, world)
    stage = po.OrcFalcon(hp, tensors, n_ctx=32)
    out = []
    for toks, n_past in ((np.array([11, 40, 41], np.int32), 0), (np.array([42], np.int32), 3)):
        resid = None
        if rank > 0:
            buf = torch.zeros(len(toks), hp["n_embd"])
            dist.recv(buf, src=rank - 1)                       # one message per boundary per eval
            resid = buf.numpy()
        r = stage.eval_range(toks, n_past, lf, ll, resid_in=resid, all_logits=True)
        if rank < world - 1:
            dist.send(torch.from_numpy(r), dst=rank + 1)
        else:
            out.append(r)
    if rank == world - 1:
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_pipeline_gloo():
    """world_size 2: rank r owns layers layer_range(L, r, 2), the residual stream crosses the boundary once per eval;
    the last rank's logits equal the single-process evaluation bit for bit."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_pipeline_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    hp = dict(TINY_40B, n_layer=5)
    whole = po.OrcFalcon(hp, synth_model(hp, po.Q4_K, seed=8), n_ctx=32)
    assert np.array_equal(got[0], whole.eval(np.array([11, 40, 41], np.int32), 0, all_logits=True))
    assert np.array_equal(got[1], whole.eval(np.array([42], np.int32), 3, all_logits=True))


def test_committed_bench_line_has_every_contract_key():
    """profiles/r1_bench_final.json is the line `python bench.py` printed on a B200 at the end of the round: the keys the
    driver and the judge read must all be there (guards bench.py's JSON contract against accidental edits)"""
    import json, os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r1_bench_final.json")
    d = json.load(open(path))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["config"]["workload"] and "model" not in d["config"]
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(d["e2e"])
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(d["roofline"])
    assert abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-9
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(d["cpu_baseline"])
    assert set(("sm_mhz", "sm_max_mhz", "reasons")) <= set(d["clocks"])
    assert d["gpu_launches"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["e2e"]["value"] <= d["value"] * 1.02          # the host round trip cannot be faster than the device-resident step
