mkdir -p gpurun_out
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step -s 26 -c 1 -o gpurun_out/r5_prof python tools/profile_step.py --preroll 24 --gens 50 --launches 4 ) > gpurun_out/r5_ncu_full.log 2>&1
tail -n 3 gpurun_out/r5_ncu_full.log
