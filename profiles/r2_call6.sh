mkdir -p gpurun_out
rm -f gpurun_out/r2_call6.txt
for v in v11 v11f32 v11ns; do VC_FRAMES=2000 MSP_LIB=$PWD/build/variants/$v.so python profiles/variant_check.py 2>/dev/null | tail -1 >> gpurun_out/r2_call6.txt; done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_v11.log 2>&1
echo pytest_v11 rc=$? >> gpurun_out/r2_call6.txt
VC_FRAMES=2000 ncu --set full --clock-control none --import-source on -k regex:k_frame_cull9 -s 2 -c 1 -o gpurun_out/prof_v11 python profiles/variant_check.py > gpurun_out/prof_v11.log 2>&1
cat gpurun_out/r2_call6.txt; tail -3 gpurun_out/r2_pytest_v11.log
