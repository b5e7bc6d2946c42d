#!/bin/bash
# Round-end GPU pass: pick the faster of the library variants under build/variants/ (if any), then run the whole GPU suite,
# the bench, one `ncu --set full` capture of K1 and the launch list with it.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
if ls build/variants/*.so >/dev/null 2>&1; then
  bash profiles/variant_sweep.sh 2>&1 | tee gpurun_out/sweep_final.txt
  best=$(sort -t= -k3 -n gpurun_out/sweep_final.txt | head -1 | cut -d' ' -f1)
  echo "chosen: $best" | tee gpurun_out/chosen.txt
  cp "$best" multi_spatialmllm_b200/libmsp.so
fi
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/final_pytest.log
python __graft_entry__.py smoke 2>&1 | tail -1 | tee -a gpurun_out/final_pytest.log
python bench.py 2>gpurun_out/final_bench.err | tail -1 > gpurun_out/final_bench.json
cut -c1-400 gpurun_out/final_bench.json
bash profiles/ncu_capture.sh "k_frame_cull.*int.0, .int.0" prof_k1_final 1 2
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_final.csv \
    python bench.py --scenes 3 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/launches_final.log 2>&1
ls -la gpurun_out
