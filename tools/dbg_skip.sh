for v in "B200_NO_GRAPH=1" "B200_NO_GRAPH=1 B200_DBG_SKIP=attn" "B200_NO_GRAPH=1 B200_NO_PDL=1" "B200_DBG_SKIP=none"; do
  env $v timeout 200 python bench.py --steps 32 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', 'ms/tok', round(d['ms_per_step'],3), 'per-layer us', round((d['ms_per_step'])*1000/60,1))"
done
