# final 8-GPU bench line of round 2 (gpurun --gpus 8) + N = 1 on the same box
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29561 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu --no-cudnn > $O/r02p_bench_n8.json 2> $O/r02p_bench_n8.err; echo bench8 rc=$?
timeout 300 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu --no-cudnn > $O/r02p_bench_n1_samebox.json 2> /dev/null; echo bench1 rc=$?
for f in $O/r02p_bench_n8.json $O/r02p_bench_n1_samebox.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); x=d.get('extra') or {}
    print(sys.argv[1], 'n', d['n_gpus'], 'value', round(d['value'],2), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],2), d['clocks'], {k:(round(v['value'],4) if isinstance(v,dict) and 'value' in v else None) for k,v in x.items()}, x.get('rank_spread'))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
