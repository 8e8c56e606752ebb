ncu --set full --clock-control none --import-source on -k regex:render_persistent -s 3 -c 1 -f -o gpurun_out/r02p_lbvh_persistent_4k python tools/lbvh_ab.py --one --sizes 3840x2160 --kernels persistent > gpurun_out/ncu_p.log 2>&1
tail -2 gpurun_out/ncu_p.log
