cd $GRAFT_REPO_ROOT
( time timeout 600 python -m pytest tests/test_fitness_gpu.py tests/test_ga_gpu.py -x -q -m gpu -k "fitness_pass or hall_of_fame or larger_population or baseline_shapes" 2>&1 | tail -15 ) > gpurun_out/r2e_tests.log 2>&1
cat gpurun_out/r2e_tests.log
timeout 300 python scripts/pass_timing.py 2>&1 | tail -6 | tee gpurun_out/r2e_pass_timing.txt
