timeout 900 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/r02m_pytest_gpu.log 2>&1
tail -12 gpurun_out/r02m_pytest_gpu.log
python tools/lbvh_ab.py --sizes 3840x2160 --kernels persistent --reps 2 --tag count --lib build/variants/libcount.so --counts gpurun_out/r02_lbvh_counts_sah.json
python tools/lbvh_ab.py --sizes 3840x2160 --kernels persistent --reps 2 --tag count --builder lbvh --lib build/variants/libcount.so --counts gpurun_out/r02_lbvh_counts_karras.json
