set -x
python -m pytest tests/test_gpu_tc.py -q -k "rowmlp" > gpurun_out/r02_tests_rowmlp2.log 2>&1; echo "rowmlp tests rc=$?"; tail -3 gpurun_out/r02_tests_rowmlp2.log
for c in 2 3 4; do WF_B200_SIDE_CHUNKS=$c python bench.py --no-cpu > gpurun_out/r02_bench_chunks$c.json 2> gpurun_out/r02_bench_chunks$c.err; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_chunks$c.json')); print('chunks $c', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],3), 'gemm', round(d['roofline']['gemm_ms_per_step'],2), d['clocks']['sm_mhz'], d['step_ms_first_last'])"; done
WF_B200_SIDE=0 python bench.py --no-cpu > gpurun_out/r02_bench_noside.json 2> gpurun_out/r02_bench_noside.err; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_noside.json')); print('noside', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'gemm', round(d['roofline']['gemm_ms_per_step'],2), round(d['roofline']['achieved'],1), d['clocks']['sm_mhz'])"
WF_BENCH_SETTLE=0 python bench.py --no-cpu --steps 2 --warmup 3 > gpurun_out/plain_bench.log 2>&1 && \
WF_BENCH_SETTLE=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches.csv python bench.py --no-cpu --steps 2 --warmup 3 > gpurun_out/r02_launches_run.log 2>&1; echo "launch list rc=$?"
python tools/launch_summary.py gpurun_out/r02_launches.csv 40 > gpurun_out/r02_launches_step_summary.txt 2>&1; head -12 gpurun_out/r02_launches_step_summary.txt
