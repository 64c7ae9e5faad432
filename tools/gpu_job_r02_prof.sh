# round 2, profile job: ncu launch list of one full step, ncu --set full of the conv kernels (tf32x3 and bf16) and of gagm_kernel,
# per-layer tables, the three bench configs + the reference arm
mkdir -p gpurun_out
NCU="ncu --clock-control none --profile-from-start off"
timeout 900 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02_launches_full_step.csv python tools/run_kernels.py full_step 4 > gpurun_out/r02_ncu_launch.log 2>&1; tail -1 gpurun_out/r02_ncu_launch.log; wc -l gpurun_out/r02_launches_full_step.csv
timeout 900 $NCU --set full -k regex:conv_tc_kernel -c 80 -o /tmp/prof_conv_tc -f python tools/run_kernels.py full_step 3 > gpurun_out/r02_ncu_conv.log 2>&1; tail -1 gpurun_out/r02_ncu_conv.log
ncu -i /tmp/prof_conv_tc.ncu-rep --page raw --csv > gpurun_out/r02_conv_tc_ncu_full_raw.csv 2>/dev/null; wc -c gpurun_out/r02_conv_tc_ncu_full_raw.csv
TTDG_CONV=bf16 timeout 900 $NCU --set full -k regex:conv_tc_kernel -c 80 -o /tmp/prof_conv_bf16 -f python tools/run_kernels.py full_step 3 > gpurun_out/r02_ncu_conv_bf16.log 2>&1; tail -1 gpurun_out/r02_ncu_conv_bf16.log
ncu -i /tmp/prof_conv_bf16.ncu-rep --page raw --csv > gpurun_out/r02_conv_tc_bf16_ncu_full_raw.csv 2>/dev/null; wc -c gpurun_out/r02_conv_tc_bf16_ncu_full_raw.csv
timeout 900 $NCU --set full --import-source on -k regex:gagm_kernel -c 1 -o /tmp/prof_gagm -f python tools/run_kernels.py full_step 3 > gpurun_out/r02_ncu_gagm.log 2>&1; tail -1 gpurun_out/r02_ncu_gagm.log
ncu -i /tmp/prof_gagm.ncu-rep --page raw --csv > gpurun_out/r02_gagm_ncu_full_raw.csv 2>/dev/null
ncu -i /tmp/prof_gagm.ncu-rep --page source --csv > gpurun_out/r02_gagm_ncu_source.csv 2>/dev/null; wc -c gpurun_out/r02_gagm_ncu_source.csv
timeout 900 $NCU --set full -k regex:"rpn_topk_decode_kernel|sort_candidates_kernel|gather_kept_kernel" -c 6 -o /tmp/prof_sel -f python tools/run_kernels.py full_step 3 > gpurun_out/r02_ncu_sel.log 2>&1
ncu -i /tmp/prof_sel.ncu-rep --page raw --csv > gpurun_out/r02_select_ncu_full_raw.csv 2>/dev/null
timeout 300 python tools/run_kernels.py layers 3 70 > gpurun_out/r02_layers_full_step_fp32.csv 2>/dev/null
TTDG_CONV=bf16 timeout 300 python tools/run_kernels.py layers 3 70 > gpurun_out/r02_layers_full_step_bf16.csv 2>/dev/null
timeout 300 python tools/run_kernels.py busy 3 gaps > gpurun_out/r02_busy_full_step.csv 2>/dev/null
timeout 300 python tools/run_kernels.py timing 3 > gpurun_out/r02_timing.log 2>&1; tail -2 gpurun_out/r02_timing.log | cut -c1-600
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2>gpurun_out/r02_bench_reference.err; cut -c1-200 gpurun_out/r02_bench_reference.json
for c in 1 2 3; do timeout 900 python bench.py --steps 10 --warmup 3 --config $c > gpurun_out/r02_bench_cfg$c.json 2>gpurun_out/r02_bench_cfg$c.err; cut -c1-200 gpurun_out/r02_bench_cfg$c.json; done
du -sh gpurun_out
