mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 > gpurun_out/bench_r2h_n8.json 2> gpurun_out/bench_r2h_n8.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_r2h_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/d2h_bw.py > gpurun_out/d2h_nobind.json 2> gpurun_out/d2h_nobind.err; cat gpurun_out/d2h_nobind.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 tools/d2h_bw.py --bind > gpurun_out/d2h_bind.json 2> gpurun_out/d2h_bind.err; cat gpurun_out/d2h_bind.json
python tools/e2e_multi.py > gpurun_out/e2e_multi.json 2> gpurun_out/e2e_multi.err; cat gpurun_out/e2e_multi.json; tail -3 gpurun_out/e2e_multi.err
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1; lscpu | grep -E "NUMA|Socket|Model name|^CPU\(s\)" >> gpurun_out/topo.txt; cat /sys/fs/cgroup/cpuset.cpus.effective >> gpurun_out/topo.txt 2>/dev/null; nproc >> gpurun_out/topo.txt
