# round-1 closing run: tests, smoke, bench (both arms), ncu evidence exported to CSV on the box
mkdir -p gpurun_out
NCU="ncu --clock-control none"
timeout 900 python -m pytest tests -m gpu -q --tb=short > gpurun_out/test_all.log 2>&1; tail -3 gpurun_out/test_all.log; grep -E "^(FAILED|E  )" gpurun_out/test_all.log | cut -c1-300 | head -30
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final.json 2>gpurun_out/bench_final.err; cut -c1-200 gpurun_out/bench_final.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final_reference.json 2>gpurun_out/bench_final_reference.err; cut -c1-200 gpurun_out/bench_final_reference.json
timeout 300 python tools/run_kernels.py layers 3 70 > gpurun_out/layers_final.csv 2>gpurun_out/layers_final_err.log; head -4 gpurun_out/layers_final.csv | cut -c1-150
timeout 300 python tools/run_kernels.py busy 3 > gpurun_out/busy_final.csv 2>gpurun_out/busy_final_err.log; head -3 gpurun_out/busy_final.csv | cut -c1-150
timeout 600 $NCU --metrics gpu__time_duration.sum -s 1100 -c 941 --csv --log-file gpurun_out/launches_final.csv python tools/run_kernels.py full_step 3 > gpurun_out/ncu_launch.log 2>&1; tail -1 gpurun_out/ncu_launch.log
timeout 600 $NCU --set full -k regex:conv_tc_kernel -s 213 -c 48 -o /tmp/prof_conv_tc -f python tools/run_kernels.py full_step 2 > gpurun_out/ncu_conv.log 2>&1; tail -1 gpurun_out/ncu_conv.log
ncu -i /tmp/prof_conv_tc.ncu-rep --page raw --csv > gpurun_out/prof_conv_tc_final_raw.csv 2>/dev/null
timeout 600 $NCU --set full -k regex:wgrad_tc_kernel -s 50 -c 10 -o /tmp/prof_wgrad_tc -f python tools/run_kernels.py full_step 2 > gpurun_out/ncu_wgrad.log 2>&1; tail -1 gpurun_out/ncu_wgrad.log
ncu -i /tmp/prof_wgrad_tc.ncu-rep --page raw --csv > gpurun_out/prof_wgrad_tc_final_raw.csv 2>/dev/null
timeout 300 $NCU --set full -k regex:sinkhorn_stream -s 1 -c 1 -o /tmp/prof_sk -f python tools/run_kernels.py sinkhorn_stream 2 > gpurun_out/ncu_sk.log 2>&1; tail -1 gpurun_out/ncu_sk.log
ncu -i /tmp/prof_sk.ncu-rep --page raw --csv > gpurun_out/prof_sinkhorn_final_raw.csv 2>/dev/null
du -sh gpurun_out
