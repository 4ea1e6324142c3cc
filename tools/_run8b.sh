set -x
python -m pytest tests/test_gpu_tc.py -q -k "l1" > gpurun_out/r02_tests_l1b.log 2>&1; echo "l1 tests rc=$?"; tail -3 gpurun_out/r02_tests_l1b.log
python tools/prof_enc_kernels.py 20 2>&1 | grep "l1 " | tee gpurun_out/r02_l1_times_b.txt
python tools/gemm_probe.py > gpurun_out/plain_probe.log 2>&1 && \
ncu --set full --clock-control none -k regex:gemm_tc_kernel -s 48 -c 12 -o /tmp/prof_pool python tools/gemm_probe.py > gpurun_out/r02_prof_pool.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py /tmp/prof_pool.ncu-rep gpurun_out/r02_pool_probe_ncu_full.csv gemm_tc; echo "summary rc=$?"
ncu -i /tmp/prof_pool.ncu-rep --page raw --csv 2>/dev/null | python -c "
import sys,csv
rows=list(csv.reader(sys.stdin)); hdr=rows[0]
want=[i for i,h in enumerate(hdr) if any(k in h for k in ('op_red','op_atom','_red.','inst_executed.sum','gpu__time_duration.sum','tc_cycles_active','Kernel Name','stalled'))]
for r in rows[2:]:
    print({hdr[i][-60:]: r[i][:60] for i in want})
" > gpurun_out/r02_pool_probe_metrics.txt 2>&1
tail -4 gpurun_out/r02_pool_probe_metrics.txt | cut -c1-1500
