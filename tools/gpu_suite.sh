#!/bin/bash
# What a gpurun call of this repository runs.  Usage (from the repo root, through gpurun):
#   bash tools/gpu_suite.sh tests                 pytest -m gpu + smoke()
#   bash tools/gpu_suite.sh bench TAG [flags]     python bench.py [flags] -> gpurun_out/bench_TAG.json (1 GPU)
#   bash tools/gpu_suite.sh benchN N TAG [flags]  the same under torchrun on N GPUs
#   bash tools/gpu_suite.sh launches TAG          ncu launch list (gpu__time_duration) of a short bench run
#   bash tools/gpu_suite.sh ncu TAG REGEX CMD...  one `ncu --set full` capture of the kernels matching REGEX in CMD
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
mode=$1; shift
case "$mode" in
tests)
    timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
    python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log ;;
bench)
    tag=$1; shift
    python bench.py "$@" > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; tail -c 400 gpurun_out/bench_$tag.json ;;
benchN)
    n=$1; tag=$2; shift 2
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n "$@" \
        > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; tail -c 400 gpurun_out/bench_$tag.json ;;
launches)
    tag=$1; shift
    F="--steps 2 --warmup 1 --no-per-config --no-cpu-baseline --no-block-api --acq-snapshots 296"
    python bench.py $F > gpurun_out/plain_$tag.log 2>&1 && \
    ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_$tag.csv \
        python bench.py $F > gpurun_out/ncu_launches_$tag.log 2>&1
    echo "launch list rc=$?"; wc -l gpurun_out/launches_$tag.csv ;;
ncu)
    tag=$1; regex=$2; shift 2
    "$@" > gpurun_out/plain_$tag.log 2>&1 && \
    timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base function -k "regex:$regex" -c 3 -f \
        -o gpurun_out/prof_$tag "$@" > gpurun_out/ncu_$tag.log 2>&1
    echo "ncu rc=$?"; tail -2 gpurun_out/ncu_$tag.log ;;
*) echo "unknown mode $mode"; exit 2 ;;
esac
