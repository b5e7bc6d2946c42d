mkdir -p gpurun_out
./build/ubench_prefetch > gpurun_out/ubench_prefetch.txt 2>&1
rm -f gpurun_out/r2_call5.txt
for v in v10 v10pf; do VC_FRAMES=2000 MSP_LIB=$PWD/build/variants/$v.so python profiles/variant_check.py 2>/dev/null | tail -1 >> gpurun_out/r2_call5.txt; done
VC_FRAMES=2000 MSP_LIB=$PWD/build/variants/v10pf.so ncu --set full --clock-control none --import-source on -k regex:k_frame_cull9 -s 2 -c 1 -o gpurun_out/prof_v10pf python profiles/variant_check.py > gpurun_out/prof_v10pf.log 2>&1
cat gpurun_out/r2_call5.txt gpurun_out/ubench_prefetch.txt
