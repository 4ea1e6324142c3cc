set -x
python -m pytest tests/test_gpu_tc.py tests/test_gpu_fullsize.py -q -k "pool or fullsize or infer" > gpurun_out/r02_tests_pool2.log 2>&1; echo "pool tests rc=$?"; tail -3 gpurun_out/r02_tests_pool2.log
python tools/gemm_probe.py 2>&1 | grep -i "pool K"
python tools/gemm_probe.py 2>&1 | grep -i "pool K"
python -m pytest tests -m gpu -q > gpurun_out/r02_tests_f.log 2>&1; echo "all tests rc=$?"; tail -4 gpurun_out/r02_tests_f.log
python bench.py --no-cpu > gpurun_out/r02_bench_d.json 2> gpurun_out/r02_bench_d.err; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_d.json')); r=d['roofline']; print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'gemm', round(r['gemm_ms_per_step'],2), round(r['achieved'],1), r['without_side_jobs']['achieved'], r['without_side_jobs']['ms_per_step'])"
python tools/bench_configs.py --config 2 2>/dev/null | cut -c1-330
