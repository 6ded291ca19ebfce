# final 1-GPU evidence of round 2 (run under gpurun): tests, bench lines, per-layer micro, launch list with DRAM bytes
TAG=${TAG:-r02k}
O=gpurun_out
set -x
python -m pytest tests -m gpu -x -q > $O/${TAG}_tests.log 2>&1; echo tests rc=$?; tail -2 $O/${TAG}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.txt 2>&1; echo smoke rc=$?
python bench.py --steps 10 --warmup 3 > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err; echo bench rc=$?
python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_ref.json 2> /dev/null; echo ref rc=$?
python tests/micro_conv.py 8 > $O/${TAG}_conv_per_layer_micro.txt 2>&1
python tests/micro_elementwise.py > $O/${TAG}_elementwise_micro.txt 2>&1
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-extra --no-cudnn"
$CMD > $O/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2600 \
    --csv --log-file $O/${TAG}_train_b8_launches.csv $CMD > $O/${TAG}_ncu_list.log 2>&1; echo list rc=$?
python profiles/summarize_launches.py $O/${TAG}_train_b8_launches.csv $O/${TAG}_train_b8_launches_summary.md $O/${TAG}_train_b8_launches.json
head -20 $O/${TAG}_train_b8_launches_summary.md; cat $O/${TAG}_conv_per_layer_micro.txt
