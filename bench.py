#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on its headline config: Falcon-40B Q4_K decode (n_batch=1) tokens/s on B200.

    python bench.py --gpus N --steps K --warmup W [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (N > 1)

A "step" is one decode eval (one token, n_batch=1) of the synthetic random-init Falcon-40B Q4_K model through the
hot path.  Weights are generated directly as well-formed Q4_K blocks on the device.
  value  : tokens/s with the token id and logits resident in HBM (b200_falcon_decode_dev), timed with CUDA events
  e2e    : tokens/s through the reference-facing C-ABI call b200_falcon_eval with HOST buffers: token id H2D and
           logits D2H inside the timed region
  roofline: the dominant kernel (mmv_fast_kernel<Q4_K>) timed alone on the model's own matrices with CUDA events;
           traffic = DRAM bytes per launch from the committed ncu launch list (profiles/r1_traffic.json)
  cpu_baseline / --impl reference: the UNMODIFIED reference's CPU path (oracle/_ref falcon_eval) on the host cores,
           on a bounded sample (a 6-layer slice of the same 40B-shaped model), scaled by weight bytes
N > 1: contiguous layer ranges, one rank per GPU, the residual stream crosses each boundary by ncclSend/ncclRecv.
value / e2e use teacher-forced token ids (consecutive tokens overlap across the pipeline stages); config also reports
autoregressive_tok_s, the strict single-stream rate (the last rank's argmax is broadcast before the next step).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FALCON_40B = dict(n_vocab=65024, n_embd=8192, n_head=128, n_head_kv=8, n_layer=60, falcon_type=40)
Q4_K = 12
METRIC = "falcon40b_q4_k_decode_tokens_per_s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region (B200_PROFILING.md recipe)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, dev):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(dev), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm = [float(r[1]) for r in rows if len(r) >= 9]
        reasons = set()
        for r in rows:
            if len(r) < 9:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(rows[0][2]) if rows else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ reference arm
def reference_cpu_decode(steps, warmup, n_layer_sample=6):
    """The UNMODIFIED reference's CPU path (falcon_eval from oracle/_ref/libfalcon_ref.so) decoding with n_batch=1 on a
    40B-shaped random Q4_K GGCC file of `n_layer_sample` layers (bounded sample), all host threads ggml can use."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as po
    import ggllm_cpp_b200.ggcc as ggcc
    kind = "reference" if po.have_ref_falcon() else "port"
    hp = dict(FALCON_40B, n_layer=n_layer_sample)
    cores = os.cpu_count() or 1
    threads = max(1, min(cores, 32))           # ggml's spin-barrier pool stops scaling past a few dozen threads (README.md:137)
    rng = np.random.default_rng(1)
    shapes = ggcc.falcon_shapes(hp)
    tensors = {}
    for name, ne in shapes.items():
        if len(ne) == 1:
            tensors[name] = (0, ne, (1.0 + 0.1 * rng.standard_normal(ne[0])).astype(np.float32) if name.endswith("weight") else (0.01 * rng.standard_normal(ne[0])).astype(np.float32))
        else:
            tensors[name] = (Q4_K, ne, ggcc.random_blocks(Q4_K, ne[1], ne[0], rng))
    sample_bytes = sum(ggcc.tensor_nbytes(t, ne) for n, (t, ne, _) in tensors.items() if len(ne) == 2 and "word_embeddings" not in n)
    full_bytes = sum(ggcc.tensor_nbytes(Q4_K, ne) for n, ne in ggcc.falcon_shapes(FALCON_40B).items() if len(ne) == 2 and "word_embeddings" not in n)
    t0 = time.time()
    if kind == "reference":
        path = os.path.join(tempfile.gettempdir(), "b200_bench_sample_%d.ggcc" % os.getpid())
        ggcc.write_ggcc(path, hp, tensors, ftype=15)
        del tensors
        eng = po.RefFalcon(path, n_ctx=256, n_batch=1)
        run = lambda tok, pos: eng.eval(np.array([tok], np.int32), pos, n_threads=threads, n_max_real_ctx=129)
    else:
        eng = po.OrcFalcon(hp, tensors, n_ctx=256)
        run = lambda tok, pos: eng.eval(np.array([tok], np.int32), pos, n_ctx_rope=129, nthreads=threads)
    run(11, 0)                                  # the reference's own warm-up eval (falcon_main.cpp:662-673)
    for i in range(warmup):
        run(100 + i, 1 + i)
    t1 = time.time()
    for i in range(steps):
        run(200 + i, 1 + warmup + i)
    dt = time.time() - t1
    if kind == "reference":
        eng.close()
        os.unlink(path)
    sample_tps = steps / dt
    full_tps = sample_tps * sample_bytes / full_bytes
    return dict(value=full_tps, unit="tok/s", cores=threads, kind=kind, ms_per_step=dt / steps * 1e3 * full_bytes / sample_bytes,
                sample=("%d decode tokens (after %d warm-up) of a %d-layer slice of the same random-init Falcon-40B Q4_K GGCC model (%.2f GB of the "
                        "%.2f GB weights) through %s falcon_eval, -t %d of %d host cores; tok/s scaled by the weight-byte ratio %.4f; setup %.0f s")
                       % (steps, warmup + 1, n_layer_sample, sample_bytes / 1e9, full_bytes / 1e9,
                          "the unmodified reference's (oracle/_ref)" if kind == "reference" else "the oracle port's", threads, cores,
                          sample_bytes / full_bytes, t1 - t0))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=128)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--layers", type=int, default=0, help="debug: use fewer layers than the real model (result is then NOT a valid bench value)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(args.warmup, 3)
    config = {"workload": "Falcon-40B Q4_K decode, n_batch=1, synthetic random-init GGCC-shaped weights, teacher-forced token ids",
              "model_shape": FALCON_40B, "n_ctx": 2048, "n_ctx_rope": 129,
              "l2": "inputs (23.2 GB of weights per step) are larger than the 126 MB L2; no explicit flush needed",
              "parallelism": "layer-range pipeline x%d (contiguous layers per GPU, ncclSend/ncclRecv of the residual per boundary)" % world if world > 1 else "single GPU"}

    if args.impl == "reference":
        if rank != 0:
            return
        r = reference_cpu_decode(steps=max(1, min(args.steps, 12)), warmup=min(warmup, 3))
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "tok/s", "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                          "dtype": "int8 x int4 block dots, fp32 accumulate (CPU)", "data": "synthetic", "config": config,
                          "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                          "e2e": {"value": r["value"], "unit": "tok/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import ggllm_cpp_b200.binding as b
    import ggllm_cpp_b200.ggcc as ggcc
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl")
    b.init(local_rank)
    L = b.lib()
    hp = dict(FALCON_40B)
    if args.layers:
        hp["n_layer"] = args.layers
        config["INVALID_debug_layers"] = args.layers
    f = b.Falcon(hp, n_ctx=config["n_ctx"], n_batch=int(os.environ.get("BENCH_NBATCH", "512")), rank=rank, world=world)      # 512: the prompt leg below (BASELINE config 3); decode uses row 0
    f.set_random(ggcc.falcon_shapes(hp), Q4_K, seed=1234)
    if world > 1:
        ids = [b.Falcon.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        f.init_pipeline(ids[0])
    stream = f.stream()
    rope = config["n_ctx_rope"]
    tok_dev = b.DevBuf(src=np.array([1234], np.int32))
    logits = np.zeros((1, hp["n_vocab"]), np.float32)

    def barrier():
        L.b200_stream_synchronize(stream)
        if dist is not None:
            dist.barrier()

    # warm-up: the reference's own BOS eval at n_past = 0 (falcon_main.cpp:662-673), then W decode steps on each path
    f.eval(np.array([11], np.int32), 0, rope)
    pos = 1
    for _ in range(warmup):
        f.decode_dev(tok_dev.ptr, pos, rope)
        pos += 1
    for _ in range(3):
        f.eval(np.array([100 + pos], np.int32), pos, rope)
        pos += 1
    barrier()

    # ---- timed region 1: device-resident decode, CUDA events on the eval stream
    e0, e1 = L.b200_event_create(), L.b200_event_create()
    sampler = ClockSampler(local_rank)
    barrier()
    t_wall0 = time.perf_counter()
    L.b200_event_record(e0, stream)
    for _ in range(args.steps):
        f.decode_dev(tok_dev.ptr, pos, rope)
        pos += 1
    L.b200_event_record(e1, stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    dev_ms = L.b200_event_elapsed_ms(e0, e1)
    launches = f.last_launches() * args.steps

    # ---- timed region 2: end to end through the C ABI with host buffers (token H2D + logits D2H each step)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        f.eval(np.array([300 + i], np.int32), pos, rope)
        pos += 1
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop()

    # ---- N > 1 only: strictly autoregressive single stream.  The two regions above are teacher-forced, so rank r can
    # start token i+1 while rank r+1 still works on token i (what falcon_perplexity-style scoring allows).  Generation
    # cannot: the next token id exists only after the last stage has produced logits.  Here the last rank "samples"
    # (argmax) and broadcasts the token id to everybody before the next step may start.
    auto_ms = None
    if dist is not None:
        import torch
        tok = torch.zeros(1, dtype=torch.int32, device="cuda")
        n_auto = min(args.steps, 32)
        barrier()
        t0 = time.perf_counter()
        for i in range(n_auto):
            lg = f.eval(np.array([int(tok.item()) % hp["n_vocab"]], np.int32), pos, rope)
            if rank == world - 1:
                tok[0] = int(np.argmax(lg[0]))
            dist.broadcast(tok, src=world - 1)
            pos += 1
        barrier()
        auto_ms = (time.perf_counter() - t0) * 1e3 / n_auto

    # ---- BASELINE's second headline number: prompt processing, n_batch = 512, 2048 synthetic tokens (4 evals at n_past
    # 0 / 512 / 1024 / 1536) through the same host-buffer C-ABI call; reported beside the decode metric, not as `value`
    ptoks = np.random.default_rng(7).integers(12, hp["n_vocab"], size=2048).astype(np.int32)
    prompt_s = float("nan")
    if int(os.environ.get("BENCH_NBATCH", "512")) >= 512:
        f.eval(ptoks[:512], 0, 0)                      # warm-up (tensor maps, scratch)
        barrier()
        t0 = time.perf_counter()
        for c in range(4):
            f.eval(ptoks[512 * c: 512 * (c + 1)], 512 * c, 0)
        barrier()
        prompt_s = time.perf_counter() - t0

    # ---- dominant kernel alone (roofline): every resident mat-vec back to back, CUDA events
    mv_ms, mv_n, mv_bytes = f.profile_matvec(reps=3)

    if dist is not None:
        import torch
        t = torch.tensor([dev_ms, e2e_s * 1e3, t_wall * 1e3, auto_ms, prompt_s], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms, wall_ms, auto_ms, prompt_s = [float(v) for v in t.tolist()]
        agg = torch.tensor([float(f.weight_bytes()), float(launches)], device="cuda", dtype=torch.float64)
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
        weight_bytes, launches = [float(v) for v in agg.tolist()]      # the roofline probe stays per GPU (rank 0's own matrices)
    else:
        e2e_ms, wall_ms, weight_bytes = e2e_s * 1e3, t_wall * 1e3, float(f.weight_bytes())
    if rank != 0:
        return

    peak, peak_src = peaks()
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as tf:
            traffic = float(json.load(tf)["dram_bytes_per_matvec_launch"])
    except Exception:
        pass
    value = args.steps / (dev_ms / 1e3)
    kv_bytes = hp["n_layer"] * 2 * (pos - args.steps) * hp["n_head_kv"] * 64 * 4
    step_bytes = weight_bytes + kv_bytes
    ach = mv_bytes / (mv_ms / 1e3) / 1e9
    if auto_ms is not None:
        config["decode_dependency"] = ("value / e2e: teacher-forced token ids, so consecutive tokens overlap across pipeline stages; "
                                       "autoregressive_tok_s: strict single stream (last rank broadcasts the argmax token before the next step)")
        config["autoregressive_tok_s"] = 1e3 / auto_ms
    out = {"metric": METRIC, "value": value, "unit": "tok/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": warmup,
           "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,     # one token stream through the whole model: total work is fixed as GPUs are added
           "dtype": "int8 x int4 block dots (dp4a), fp32 accumulate; f32 KV/attention", "data": "synthetic", "config": config,
           "e2e": {"value": args.steps / (e2e_ms / 1e3), "unit": "tok/s", "h2d_bytes_per_step": 8, "d2h_bytes_per_step": hp["n_vocab"] * 4,
                   "ms_per_step": e2e_ms / args.steps, "api": "b200_falcon_eval (host token id in, host logits out)"},
           "gpu_launches": int(launches),
           "clocks": clocks,
           "roofline": {"bound": "hbm", "kernel": "mmv_fast_kernel<Q4_K> (register-resident fused dequantise + int8 dot mat-vec)", "achieved": ach, "peak": peak, "unit": "GB/s",
                        "frac": ach / peak, "peak_source": peak_src, "traffic": traffic,
                        "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum per mat-vec launch, profiles/r1_bench_launches.csv" if traffic else None,
                        "launches_timed": int(mv_n), "avg_launch_us": mv_ms * 1e3 / max(mv_n, 1), "algorithmic_bytes_per_launch": mv_bytes / max(mv_n, 1),
                        "how": "all resident mat-vecs (4 per layer + lm_head) launched back to back x3 on the eval stream, CUDA events around the region; "
                               "each launch reads a different matrix, one pass = 23.2 GB >> L2",
                        "step_achieved_GBs_per_gpu": step_bytes * value / 1e9 / world, "step_frac": step_bytes * value / 1e9 / world / peak,
                        "step_bytes": step_bytes, "step_roofline_tok_s_per_gpu": peak * 1e9 / step_bytes},
           "wall_ms_per_step": wall_ms / args.steps,
           "prompt": {"tok_s": 2048 / prompt_s, "tokens": 2048, "n_batch": 512, "seconds": prompt_s,
                      "matmul_TFLOPs": 2.0 * (weight_bytes / 0.5625) * 2048 / prompt_s / 1e12,
                      "what": "Falcon-40B Q4_K prompt (BASELINE config 3): 4 x b200_falcon_eval of 512 host tokens, tcgen05 GEMM with fused "
                              "dequantisation + tcgen05 attention, wall clock incl. H2D / D2H"}}
    if world == 1 and not args.no_cpu_baseline:
        try:
            r = reference_cpu_decode(steps=8, warmup=2)
            out["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as ex:       # the baseline is reporting only; never let it take the GPU number down
            out["cpu_baseline"] = {"value": None, "unit": "tok/s", "cores": 0, "kind": "reference", "sample": "failed: %r" % (ex,)}
    print(json.dumps(out))
    f.free()


if __name__ == "__main__":
    main()
