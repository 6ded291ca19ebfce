# ncu launch lists (time + DRAM bytes per launch) of the inference and scan workloads with the final code
O=gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
CI="python bench.py --workload infer --steps 2 --warmup 1 --no-cpu"
$CI > $O/r02x_infer_plain.log 2>&1 &&
ncu --metrics $M --clock-control none -c 600 --csv --log-file $O/r02x_infer_b5_launches.csv $CI > $O/r02x_infer_ncu.log 2>&1; echo infer rc=$?
CS="python bench.py --workload scan --steps 1 --warmup 1 --no-cpu"
$CS > $O/r02x_scan_plain.log 2>&1 &&
ncu --metrics $M --clock-control none -c 900 --csv --log-file $O/r02x_scan_launches.csv $CS > $O/r02x_scan_ncu.log 2>&1; echo scan rc=$?
python profiles/summarize_launches.py $O/r02x_infer_b5_launches.csv $O/r02x_infer_b5_launches_summary.md $O/r02x_infer_b5_launches.json
python profiles/summarize_launches.py $O/r02x_scan_launches.csv $O/r02x_scan_launches_summary.md $O/r02x_scan_launches.json
head -14 $O/r02x_infer_b5_launches_summary.md; head -16 $O/r02x_scan_launches_summary.md
