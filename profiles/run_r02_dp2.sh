set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2d_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2d_tests.log; tail -5 gpurun_out/r2d_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
python bench.py --steps 10 --warmup 3 --no-extra --no-cpu > gpurun_out/r2d_n1.json 2> gpurun_out/r2d_n1.err; echo n1 rc=$?
$TR bench.py --gpus 2 --steps 10 --warmup 3 --no-extra > gpurun_out/r2d_n2_peer_flat.json 2> gpurun_out/r2d_n2_peer_flat.err; echo n2 rc=$?
DRAM_GRAD_OVERLAP=1 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-extra > gpurun_out/r2d_n2_peer_overlap.json 2> gpurun_out/r2d_n2_peer_overlap.err; echo n2b rc=$?
DRAM_PEER=0 DRAM_GRAD_OVERLAP=1 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-extra > gpurun_out/r2d_n2_nccl_overlap.json 2> gpurun_out/r2d_n2_nccl_overlap.err; echo n2c rc=$?
DRAM_PEER=0 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-extra > gpurun_out/r2d_n2_nccl_flat.json 2> gpurun_out/r2d_n2_nccl_flat.err; echo n2d rc=$?
$TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2d_n2_full.json 2> gpurun_out/r2d_n2_full.err; echo n2full rc=$?
for f in gpurun_out/r2d_n*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], 'n', d['n_gpus'], 'value', round(d['value'],2), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],2), d['clocks']['sm_mhz'] if d.get('clocks') else None)
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
tail -5 gpurun_out/r2d_n2_peer_flat.err
