mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_synth.py -x -q -m gpu -k "phase or 600s or prologue or visib or presets or scans" 2>&1 | tail -4 | tee gpurun_out/pytest_pq.log
: > gpurun_out/cold.log
for i in 1 2 3; do
  R4WB_PROLOGUE_TRACE=1 python bench.py --no-per-config --no-parity --no-block-api --no-cpu-baseline 2>gpurun_out/cold_err.txt | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('prologue_ms', d['prologue_ms'], 'value', d['value'], 'cold', d['value_cold'])" | tee -a gpurun_out/cold.log
  grep "r4wb prologue" gpurun_out/cold_err.txt | tail -1 | tee -a gpurun_out/cold.log
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02_x_prologue_launches.csv python tools/prof_prologue.py e1c_8prn_600s_cn34_orbital.yaml 0 3e9 > gpurun_out/ncu_pro.log 2>&1
