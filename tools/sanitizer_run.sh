# compute-sanitizer passes over the GPU parity tests (memcheck is run separately, see profiles/r02_sanitizer.txt):
#   gpurun --timeout 2100 -- 'bash tools/sanitizer_run.sh'
timeout 1300 compute-sanitizer --tool racecheck --error-exitcode 7 --log-file gpurun_out/racecheck.log python -m pytest tests/test_step_gpu.py -x -q -m gpu -k "not 1000 and not large" > gpurun_out/racecheck_tests.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/racecheck_tests.log; grep -c "Race reported\|hazard" gpurun_out/racecheck.log; tail -5 gpurun_out/racecheck.log
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 7 --log-file gpurun_out/synccheck.log python -m pytest tests/test_step_gpu.py -x -q -m gpu -k "not 1000 and not large" > gpurun_out/synccheck_tests.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/synccheck_tests.log; tail -3 gpurun_out/synccheck.log
