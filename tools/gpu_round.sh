mkdir -p gpurun_out
for L in 1 8; do ( timeout 300 python tools/gen_profile.py --lanes $L ) >> gpurun_out/r1_genprof.json 2>> gpurun_out/r1_genprof.err; done
( timeout 300 python tools/gen_profile.py --lanes 8 --reuse 0 ) >> gpurun_out/r1_genprof.json 2>> gpurun_out/r1_genprof.err
( timeout 300 python tools/gen_profile.py --lanes 1 --games 262144 --preroll 12 ) >> gpurun_out/r1_genprof.json 2>> gpurun_out/r1_genprof.err
cat gpurun_out/r1_genprof.json; tail -3 gpurun_out/r1_genprof.err
