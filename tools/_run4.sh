set -x
python -m pytest tests -m gpu -q > gpurun_out/r02_tests_d.log 2>&1; echo "all tests rc=$?"; tail -5 gpurun_out/r02_tests_d.log
python tools/bench_configs.py --config 4 > gpurun_out/r02b_config4_n1.jsonl 2> gpurun_out/r02b_config4_n1.err; echo "cfg4 rc=$?"
python tools/bench_configs.py --config 4 --clouds 1 > gpurun_out/r02b_config4_n1_1cloud.jsonl 2>> gpurun_out/r02b_config4_n1.err
python tools/bench_configs.py --config 2 > gpurun_out/r02b_config2.jsonl 2> gpurun_out/r02b_config2.err; echo "cfg2 rc=$?"
cat gpurun_out/r02b_config4_n1.jsonl gpurun_out/r02b_config4_n1_1cloud.jsonl gpurun_out/r02b_config2.jsonl | cut -c1-260
