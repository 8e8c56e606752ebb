timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "banded or sync or update_frame or headless" > gpurun_out/r02n_pytest_sync.log 2>&1
tail -5 gpurun_out/r02n_pytest_sync.log
timeout 300 python tools/sync_ab.py > gpurun_out/r02n_sync_ab.jsonl 2> gpurun_out/r02n_sync_ab.err
cat gpurun_out/r02n_sync_ab.jsonl; tail -3 gpurun_out/r02n_sync_ab.err
RT_PROGRESS_DEBUG=1 timeout 100 python tools/_dbg_progress.py 2>&1 | tail -12
