set -x
python -m pytest tests/test_gpu_tc.py -q -k "rowmlp" -x > gpurun_out/r02_tests_rowmlp.log 2>&1; echo "rowmlp tests rc=$?"; tail -15 gpurun_out/r02_tests_rowmlp.log
python -m pytest tests -m gpu -q > gpurun_out/r02_tests_c.log 2>&1; echo "all tests rc=$?"; tail -5 gpurun_out/r02_tests_c.log
WF_B200_ROWMLP=0 python bench.py --no-cpu > gpurun_out/r02_bench_rowmlp0.json 2> gpurun_out/r02_bench_rowmlp0.err; cut -c1-330 gpurun_out/r02_bench_rowmlp0.json
WF_B200_ROWMLP=1 python bench.py --no-cpu > gpurun_out/r02_bench_rowmlp1.json 2> gpurun_out/r02_bench_rowmlp1.err; cut -c1-330 gpurun_out/r02_bench_rowmlp1.json
python tools/step_timeline.py 4 > gpurun_out/r02_timeline_rowmlp.txt 2>&1; head -3 gpurun_out/r02_timeline_rowmlp.txt
