# round 2, job ov3: device-busy fraction and idle gaps of the overlapped schedule (CUPTI via torch.profiler)
mkdir -p gpurun_out
timeout 300 python tools/run_kernels.py busy 3 gaps overlap > gpurun_out/r02ov3_busy_overlap.csv 2>gpurun_out/r02ov3_busy_overlap.err; head -24 gpurun_out/r02ov3_busy_overlap.csv | cut -c1-200; tail -3 gpurun_out/r02ov3_busy_overlap.err
