mkdir -p gpurun_out
for v in 100 50 25 0; do echo "carveout=$v" >> gpurun_out/r14_variants.json; ( B2AZ_CARVEOUT=$v timeout 300 python tools/gen_profile.py --span 410 ) >> gpurun_out/r14_variants.json 2>> gpurun_out/r14_variants.err; done
echo "ldcg" >> gpurun_out/r14_variants.json; ( B2AZ_LIB_PATH=build/variants/libb2az_ldcg.so timeout 300 python tools/gen_profile.py --span 410 ) >> gpurun_out/r14_variants.json 2>> gpurun_out/r14_variants.err
echo "ldcg carveout 100" >> gpurun_out/r14_variants.json; ( B2AZ_CARVEOUT=100 B2AZ_LIB_PATH=build/variants/libb2az_ldcg.so timeout 300 python tools/gen_profile.py --span 410 ) >> gpurun_out/r14_variants.json 2>> gpurun_out/r14_variants.err
cut -c1-140 gpurun_out/r14_variants.json; tail -n 3 gpurun_out/r14_variants.err
