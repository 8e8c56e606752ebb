ncu --set full --clock-control none --import-source on -k regex:render_persistent -s 3 -c 1 -f -o gpurun_out/r02m_lbvh_persistent_4k python tools/lbvh_ab.py --one --sizes 3840x2160 --kernels persistent > gpurun_out/ncu_m.log 2>&1
tail -3 gpurun_out/ncu_m.log
(echo "--- memcheck"; timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_run.py 2>&1 | tail -4; echo "--- racecheck"; timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_run.py 2>&1 | tail -4) > gpurun_out/r02m_sanitizer.txt
cat gpurun_out/r02m_sanitizer.txt
