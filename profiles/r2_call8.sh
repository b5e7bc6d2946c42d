mkdir -p gpurun_out
rm -f gpurun_out/r2_call8.txt
for v in v13l0 v13 v13c v13cf32; do VC_FRAMES=2000 MSP_LIB=$PWD/build/variants/$v.so python profiles/variant_check.py 2>/dev/null | tail -1 >> gpurun_out/r2_call8.txt; done
VC_FRAMES=2000 MSP_LIB=$PWD/build/variants/v13c.so ncu --set full --clock-control none --import-source on -k regex:k_frame_cull9 -s 2 -c 1 -o gpurun_out/prof_v13c python profiles/variant_check.py > gpurun_out/prof_v13c.log 2>&1
cat gpurun_out/r2_call8.txt
