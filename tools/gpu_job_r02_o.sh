# round 2, job o: source-level ncu of the short-K 1x1 layer (64 -> 256 at 128^2 with residual) with both epilogues
mkdir -p gpurun_out
NCU="ncu --clock-control none --set full --import-source on -k regex:conv_tc_kernel"
for epi in 0 1; do
TTDG_TC_EPI=$epi timeout 120 python tools/conv_layer.py 8 128 128 64 256 1 0 1 1 1 6
TTDG_TC_EPI=$epi timeout 300 $NCU -s 3 -c 1 -o /tmp/prof_l_$epi -f python tools/conv_layer.py 8 128 128 64 256 1 0 1 1 1 5 > gpurun_out/r02o_ncu_$epi.log 2>&1; tail -1 gpurun_out/r02o_ncu_$epi.log
ncu -i /tmp/prof_l_$epi.ncu-rep --page source --csv > /tmp/src_$epi.csv 2>/dev/null
python tools/ncu_top_sass.py /tmp/src_$epi.csv 70 > gpurun_out/r02o_top_sass_epi$epi.txt
ncu -i /tmp/prof_l_$epi.ncu-rep --page raw --csv > gpurun_out/r02o_raw_epi$epi.csv 2>/dev/null
done
TTDG_CONV=bf16 TTDG_TC_EPI=0 timeout 120 python tools/conv_layer.py 8 128 128 64 256 1 0 1 1 1 6
TTDG_CONV=bf16 TTDG_TC_EPI=1 timeout 120 python tools/conv_layer.py 8 128 128 64 256 1 0 1 1 1 6
TTDG_TC_EPI=1 timeout 120 python tools/conv_layer.py 8 128 128 64 256 1 0 1 0 1 6
TTDG_TC_EPI=1 timeout 120 python tools/conv_layer.py 8 128 128 256 64 1 0 1 0 1 6
