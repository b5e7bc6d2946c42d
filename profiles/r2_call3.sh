set -x
mkdir -p gpurun_out
for v in v9 v9f32; do VC_FRAMES=2000 MSP_LIB=$PWD/build/variants/$v.so python profiles/variant_check.py 2>/dev/null | tail -1 >> gpurun_out/r2_call3.txt; done
VC_FRAMES=2000 MSP_CULL_V9=0 python profiles/variant_check.py 2>/dev/null | tail -1 >> gpurun_out/r2_call3.txt
VC_FRAMES=2000 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_v9.csv python profiles/variant_check.py > /dev/null 2>&1
VC_FRAMES=2000 ncu --set full --clock-control none --import-source on -k regex:k_frame_cull9 -s 2 -c 1 -o gpurun_out/prof_v9a python profiles/variant_check.py > gpurun_out/prof_v9a.log 2>&1
cat gpurun_out/r2_call3.txt
