cd "$(dirname "$0")/.."
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/san_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/san_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/san_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/san_racecheck.log
