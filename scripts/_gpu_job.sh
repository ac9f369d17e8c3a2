python scripts/prof_geometry.py > gpurun_out/s16_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:knn_block --launch-skip 3 -c 1 -f -o gpurun_out/s16_knn python scripts/prof_geometry.py > gpurun_out/s16_ncu.log 2>&1
