# 8-GPU evidence of round 2 (gpurun --gpus 8): data-parallel parity at 4 and 8 ranks, the bench line at N = 8, and a CUPTI
# timeline of one eager 8-rank step (rank 0).  Every multi-rank command runs under its own timeout.
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
set -x
nvidia-smi -L | head -8 > $O/r02d_gpus.txt
timeout 300 $TR --nproc-per-node 8 --master-port 29541 tests/dp_parity.py     > $O/r02d_dp_parity_n8_dc3d.txt 2>&1; echo parity8 rc=$?
timeout 300 $TR --nproc-per-node 8 --master-port 29542 tests/dp_parity.py att > $O/r02d_dp_parity_n8_att.txt 2>&1; echo parity8att rc=$?
timeout 300 $TR --nproc-per-node 4 --master-port 29543 tests/dp_parity.py     > $O/r02d_dp_parity_n4_dc3d.txt 2>&1; echo parity4 rc=$?
timeout 300 $TR --nproc-per-node 4 --master-port 29544 tests/dp_parity.py att > $O/r02d_dp_parity_n4_att.txt 2>&1; echo parity4att rc=$?
grep -h "DP parity\|DP PARITY" $O/r02d_dp_parity_n*.txt
timeout 600 $TR --nproc-per-node 8 --master-port 29545 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu --no-cudnn > $O/r02d_bench_n8.json 2> $O/r02d_bench_n8.err; echo bench8 rc=$?
timeout 300 $TR --nproc-per-node 8 --master-port 29546 profiles/dp_timeline.py $O/r02d_dp_timeline_n8.md > $O/r02d_dp_timeline_n8.log 2>&1; echo timeline rc=$?
timeout 300 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu --no-cudnn > $O/r02d_bench_n1_samebox.json 2> /dev/null; echo bench1 rc=$?
for f in $O/r02d_bench_n8.json $O/r02d_bench_n1_samebox.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); x=d.get('extra') or {}
    print(sys.argv[1], 'n', d['n_gpus'], 'value', round(d['value'],2), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],2), d['clocks'], {k:(round(v['value'],4) if isinstance(v,dict) and 'value' in v else None) for k,v in x.items()})
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
head -12 $O/r02d_dp_timeline_n8.md
