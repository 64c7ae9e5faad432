# round 2, job ga1: GA-GM Hungarian-stage segment accounting, U-all on / off
mkdir -p gpurun_out
TTDG_GAGM_UALL=1 timeout 200 python tools/run_kernels.py gagm_bench 2 2>&1 | grep gagm_bench | cut -c1-600 > gpurun_out/r02ga1_uall1.txt
TTDG_GAGM_UALL=0 timeout 200 python tools/run_kernels.py gagm_bench 2 2>&1 | grep gagm_bench | cut -c1-600 > gpurun_out/r02ga1_uall0.txt
cat gpurun_out/r02ga1_uall1.txt gpurun_out/r02ga1_uall0.txt
