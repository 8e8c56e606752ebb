timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/r02m_pytest_mg.log 2>&1
tail -3 gpurun_out/r02m_pytest_mg.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r02m_scale_n2.json 2> gpurun_out/r02m_scale_n2.err
tail -c 1500 gpurun_out/r02m_scale_n2.json; tail -3 gpurun_out/r02m_scale_n2.err
