cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_gputest.log 2>&1; tail -3 gpurun_out/r2i_gputest.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_fm.py -m gpu -x -q -k "not full_size and not aten_mirrors" > gpurun_out/r2i_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/r2i_memcheck.log
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_fm.py -m gpu -x -q -k "update_embedding_and_fit_bit_exact or presorted or sort_fields" > gpurun_out/r2i_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -5 gpurun_out/r2i_racecheck.log
