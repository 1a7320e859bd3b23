cd "$(dirname "$0")/.."
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base function -k 'regex:^k_synth$' -s 1 -c 1 -f -o gpurun_out/ksynth_general python tools/prof_general.py configs/e1c_8prn_60s_cn34_orbital.yaml 5e7 3 > gpurun_out/ncu_ksynth.log 2>&1
tail -3 gpurun_out/ncu_ksynth.log
