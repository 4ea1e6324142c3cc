set -x
python -m pytest tests/test_gpu_tc.py -q -k "pool" > gpurun_out/r02_tests_pool.log 2>&1; echo "pool tests rc=$?"; tail -3 gpurun_out/r02_tests_pool.log
python tools/gemm_probe.py 2>&1 | tee gpurun_out/r02_gemm_probe.txt | grep -i "pool"
python -m pytest tests -m gpu -q > gpurun_out/r02_tests_e.log 2>&1; echo "all tests rc=$?"; tail -4 gpurun_out/r02_tests_e.log
python bench.py > gpurun_out/r02_bench_c.json 2> gpurun_out/r02_bench_c.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_c.json')); r=d['roofline']; print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'gemm', round(r['gemm_ms_per_step'],2), round(r['achieved'],1), r['without_side_jobs'], d['cpu_baseline']['value'], d['gpu_eager_baseline']['fp32']['value'], d['gpu_eager_baseline']['tf32']['value'])"
python tools/bench_configs.py --config 2 > gpurun_out/r02c_config2.jsonl 2> gpurun_out/r02c_config2.err; cut -c1-400 gpurun_out/r02c_config2.jsonl
