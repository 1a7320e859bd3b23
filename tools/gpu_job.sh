mkdir -p gpurun_out
python tools/prof_prologue.py > gpurun_out/prologue_r2j.log 2>&1; cat gpurun_out/prologue_r2j.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/prologue_launches_r2j.csv python tools/prof_prologue.py > /dev/null 2>&1; grep -E "k_phase_q|k_phase_chain" gpurun_out/prologue_launches_r2j.csv | awk -F'","' '{print $5, $NF}' | head -4
timeout 1200 python -m pytest tests/test_gpu_synth.py -x -q -m gpu -k "600s or prologue or phase or cn0 or lattice or block_loop" 2>&1 | tail -4
