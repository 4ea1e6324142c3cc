python tools/gemm_probe.py 2>&1 | grep -i "pool K"
WF_B200_POOL_NOPUSH=1 timeout 120 python tools/gemm_probe.py 2>&1 | grep -i "pool K\|rror"
python tools/gemm_probe.py 2>&1 | grep -i "pool K"
WF_B200_POOL_NOPUSH=1 timeout 120 python tools/gemm_probe.py 2>&1 | grep -i "pool K\|rror"
