cd $GRAFT_REPO_ROOT
( time timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 ) > gpurun_out/r02_gpu_tests_final.log 2>&1
cat gpurun_out/r02_gpu_tests_final.log
bash scripts/gpu_evidence2.sh
