for k in outproj qk p1; do PZ_RG_TIMELINE=$k python scripts/rowgemm_timeline.py > gpurun_out/s11_tl_$k.txt 2>&1; done
