mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short > gpurun_out/test_all.log 2>&1; tail -4 gpurun_out/test_all.log; grep -E "^(FAILED|E  )" gpurun_out/test_all.log | cut -c1-250 | head -20
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 6000 --csv --log-file gpurun_out/launches_full_step_tc.csv python tools/run_kernels.py full_step 2 > gpurun_out/ncu5.log 2>&1; tail -2 gpurun_out/ncu5.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 60 -c 4 -o gpurun_out/prof_conv_tc -f python tools/run_kernels.py full_step 1 > gpurun_out/ncu6.log 2>&1; tail -2 gpurun_out/ncu6.log
