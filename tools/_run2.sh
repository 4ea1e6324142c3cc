set -x
python -m pytest tests/test_gpu_tc.py -q -k "l1" > gpurun_out/r02_tests_l1.log 2>&1; echo "l1 tests rc=$?"; tail -3 gpurun_out/r02_tests_l1.log
python -m pytest tests/test_gpu_simt.py -q -k "edge" > gpurun_out/r02_tests_edge.log 2>&1; echo "edge tests rc=$?"; tail -3 gpurun_out/r02_tests_edge.log
for v in "1 1" "1 2" "0 1"; do set -- $v; echo "L1_MMA=$1 NTL=$2"; WF_B200_L1_MMA=$1 WF_B200_L1_BWD_NTL=$2 python tools/prof_enc_kernels.py 20 2>&1 | grep "l1 "; done | tee gpurun_out/r02_l1_times.txt
python -m pytest tests -m gpu -q > gpurun_out/r02_tests_b.log 2>&1; echo "all tests rc=$?"; tail -5 gpurun_out/r02_tests_b.log
python bench.py --no-eager > gpurun_out/r02_bench_b.json 2> gpurun_out/r02_bench_b.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r02_bench_b.json
