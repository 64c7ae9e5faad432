# ncu evidence for the full step at the current commit (launch list + --set full captures of the top kernels);
# reports are exported to CSV on the box and the .ncu-rep files dropped (gpurun_out/ is capped at 64 MiB)
mkdir -p gpurun_out
NCU="ncu --clock-control none"
timeout 300 python tools/run_kernels.py layers 3 60 > gpurun_out/layers_full_step.csv 2>gpurun_out/layers_err.log; head -45 gpurun_out/layers_full_step.csv | cut -c1-160
timeout 600 $NCU --metrics gpu__time_duration.sum -s 1500 -c 1700 --csv --log-file gpurun_out/launches_full_step.csv python tools/run_kernels.py full_step 2 > gpurun_out/ncu_launch.log 2>&1; tail -1 gpurun_out/ncu_launch.log
timeout 600 $NCU --set full -k regex:conv_tc_kernel -s 213 -c 48 -o /tmp/prof_conv_tc -f python tools/run_kernels.py full_step 2 > gpurun_out/ncu_conv.log 2>&1; tail -1 gpurun_out/ncu_conv.log
ncu -i /tmp/prof_conv_tc.ncu-rep --page raw --csv > gpurun_out/prof_conv_tc_raw.csv 2>/dev/null
timeout 600 $NCU --set full -k regex:wgrad_tc_kernel -s 50 -c 10 -o /tmp/prof_wgrad_tc -f python tools/run_kernels.py full_step 2 > gpurun_out/ncu_wgrad.log 2>&1; tail -1 gpurun_out/ncu_wgrad.log
ncu -i /tmp/prof_wgrad_tc.ncu-rep --page raw --csv > gpurun_out/prof_wgrad_tc_raw.csv 2>/dev/null
timeout 600 $NCU --set full --import-source on -k regex:gagm_kernel -s 1 -c 1 -o /tmp/prof_gagm_full -f python tools/run_kernels.py full_step 2 > gpurun_out/ncu_gagm.log 2>&1; tail -1 gpurun_out/ncu_gagm.log
ncu -i /tmp/prof_gagm_full.ncu-rep --page raw --csv > gpurun_out/prof_gagm_raw.csv 2>/dev/null
ncu -i /tmp/prof_gagm_full.ncu-rep --page source --csv > gpurun_out/prof_gagm_source.csv 2>/dev/null
cp /tmp/prof_gagm_full.ncu-rep gpurun_out/ 2>/dev/null
du -sh gpurun_out; ls -la gpurun_out
