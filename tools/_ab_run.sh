S="--sizes 1920x1080,3840x2160 --kernels persistent --reps 6"
python tools/lbvh_ab.py $S --tag sah --builder sah > gpurun_out/r02_bvh_builder_ab.jsonl
python tools/lbvh_ab.py $S --tag karras --builder lbvh >> gpurun_out/r02_bvh_builder_ab.jsonl
python tools/lbvh_ab.py $S --tag sah2 --builder sah >> gpurun_out/r02_bvh_builder_ab.jsonl
cat gpurun_out/r02_bvh_builder_ab.jsonl
timeout 600 python -m pytest tests -m gpu -x -q -k "lbvh or config5 or deep_tree or bvh_builders or refit or anyhit or large_scene" 2>&1 | tail -8 > gpurun_out/r02l_pytest_bvh.log
cat gpurun_out/r02l_pytest_bvh.log
