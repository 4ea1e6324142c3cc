set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02_tests_a.log 2>&1; echo "tests rc=$?"
python bench.py > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err; echo "bench rc=$?"
python tools/l2_probe.py > gpurun_out/r02_l2_probe.jsonl 2> gpurun_out/r02_l2_probe.err; echo "probe rc=$?"
PROBE_ROWS=18944 python tools/l2_probe.py >> gpurun_out/r02_l2_probe.jsonl 2>> gpurun_out/r02_l2_probe.err
PROBE_ROWS=4864 python tools/l2_probe.py >> gpurun_out/r02_l2_probe.jsonl 2>> gpurun_out/r02_l2_probe.err
for m in default l2; do
PROBE_MODE=$m timeout 600 ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_l2_dram_$m.csv python tools/l2_probe.py > gpurun_out/r02_l2_ncu_$m.log 2>&1; echo "ncu $m rc=$?"
done
tail -3 gpurun_out/r02_tests_a.log; cat gpurun_out/r02_bench_a.json | cut -c1-600; cat gpurun_out/r02_l2_probe.jsonl; tail -5 gpurun_out/r02_l2_probe.err
