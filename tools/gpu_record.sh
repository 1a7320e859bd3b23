# the round's record run: GPU test-suite, smoke(), the default bench line and the reference arm (1 GPU), or the bench under
# torchrun (bash tools/gpu_record.sh N TAG)
mkdir -p gpurun_out
n=${1:-1}; tag=${2:-final}
if [ "$n" = "1" ]; then
  timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
  python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
  timeout 1200 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_$tag.err
  timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${tag}_reference.json 2> gpurun_out/bench_${tag}_reference.err; echo "reference rc=$?"
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n > gpurun_out/bench_${tag}_n$n.json 2> gpurun_out/bench_${tag}_n$n.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_${tag}_n$n.err
fi
