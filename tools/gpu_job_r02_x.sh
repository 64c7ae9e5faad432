# round 2, job x: does a 3-stage ring (instead of 4; bf16 4 instead of 6) cost the big layers anything?  (frees 48 KB for an epilogue tile)
mkdir -p gpurun_out
for lib in "" tools/build/libttdg_exp_s3.so; do
echo "== lib '$lib'"
for mode in tf32x3 bf16; do
TTDG_LIB=$lib TTDG_CONV=$mode timeout 120 python tools/conv_layer.py 8 128 128 256 256 3 1 1 0 1 6 | cut -c60-
TTDG_LIB=$lib TTDG_CONV=$mode timeout 120 python tools/conv_layer.py 800 14 14 256 256 3 1 1 0 1 6 | cut -c60-
TTDG_LIB=$lib TTDG_CONV=$mode timeout 120 python tools/conv_layer.py 8 32 32 256 256 3 1 1 0 1 6 | cut -c60-
TTDG_LIB=$lib TTDG_CONV=$mode timeout 120 python tools/conv_layer.py 8 16 16 512 512 3 1 1 0 1 6 | cut -c60-
TTDG_LIB=$lib TTDG_CONV=$mode timeout 120 python tools/conv_layer.py 8 32 32 1024 256 1 0 1 0 1 6 | cut -c60-
TTDG_LIB=$lib TTDG_CONV=$mode timeout 120 python tools/conv_layer.py 8 128 128 64 256 1 0 1 1 1 6 | cut -c60-
done
done
