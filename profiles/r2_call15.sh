mkdir -p gpurun_out
rm -f gpurun_out/r2_call15.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_k2.log 2>&1
echo pytest_k2 rc=$? >> gpurun_out/r2_call15.txt
for r in "" morton; do REORDER=$r python profiles/stress_probe.py 2>/dev/null | tail -1 >> gpurun_out/r2_call15.txt; done
for r in "" morton; do MSP_CULL_V9=0 REORDER=$r python profiles/stress_probe.py 2>/dev/null | tail -1 >> gpurun_out/r2_call15.txt; done
python profiles/zbuffer_bench.py > gpurun_out/r2_zbuffer_bench.json 2>/dev/null; tail -3 gpurun_out/r2_zbuffer_bench.json >> gpurun_out/r2_call15.txt
REORDER= ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2_launches_stress_k2v9.csv python profiles/stress_probe.py > /dev/null 2>&1
REORDER= ncu --set full --clock-control none --import-source on -k regex:k_zbuf_cull9 -s 2 -c 1 -o gpurun_out/prof_k2_v9 python profiles/stress_probe.py > /dev/null 2>&1
cat gpurun_out/r2_call15.txt; tail -4 gpurun_out/r2_pytest_k2.log
