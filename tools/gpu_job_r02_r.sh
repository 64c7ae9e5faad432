# round 2, job r: per-tile timeline of CTA 0 (clock64 at 8 points) for the short-K layers
mkdir -p gpurun_out
(TTDG_TRACE=1 timeout 120 python tools/conv_layer.py 8 128 128 64 256 1 0 1 1 1 4
TTDG_TRACE=1 timeout 120 python tools/conv_layer.py 8 128 128 64 256 1 0 1 0 1 4
TTDG_TRACE=1 timeout 120 python tools/conv_layer.py 8 128 128 256 64 1 0 1 0 1 4
TTDG_TRACE=1 TTDG_CONV=bf16 timeout 120 python tools/conv_layer.py 8 128 128 64 256 1 0 1 1 1 4
TTDG_TRACE=1 timeout 120 python tools/conv_layer.py 8 128 128 256 256 3 1 1 0 1 4) > gpurun_out/r02r_trace.txt 2>&1
tail -5 gpurun_out/r02r_trace.txt
