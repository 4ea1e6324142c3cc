set -x
python tools/prof_gemm.py > gpurun_out/plain_prof_gemm.log 2>&1 && \
ncu --set full --clock-control none -k regex:gemm_tc_kernel -s 20 -c 20 -o /tmp/prof_side python tools/prof_gemm.py > gpurun_out/r02_prof_gemm_side.log 2>&1; echo "ncu side rc=$?"
python tools/ncu_summary.py /tmp/prof_side.ncu-rep gpurun_out/r02_gemm_side_ncu_full.csv gemm_tc; echo "summary rc=$?"
WF_B200_SIDE=0 python tools/prof_gemm.py > gpurun_out/plain_prof_gemm0.log 2>&1 && \
WF_B200_SIDE=0 ncu --set full --clock-control none -k regex:gemm_tc_kernel -s 10 -c 10 -o /tmp/prof_noside python tools/prof_gemm.py > gpurun_out/r02_prof_gemm_noside.log 2>&1; echo "ncu noside rc=$?"
python tools/ncu_summary.py /tmp/prof_noside.ncu-rep gpurun_out/r02_gemm_noside_ncu_full.csv gemm_tc; echo "summary rc=$?"
python bench.py --no-cpu --steps 2 --warmup 3 > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches.csv python bench.py --no-cpu --steps 2 --warmup 3 > gpurun_out/r02_launches_run.log 2>&1; echo "launch list rc=$?"
du -sh gpurun_out
