cd $GRAFT_REPO_ROOT
timeout 540 python scripts/train_and_evaluate.py --episodes 300 --train-sets 8 2> gpurun_out/r02_train.err | tail -1 > gpurun_out/r02_trained_agent_quality.json
tail -c 1500 gpurun_out/r02_trained_agent_quality.json; tail -n 5 gpurun_out/r02_train.err
