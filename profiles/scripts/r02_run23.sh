set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"field_half|field_rec_kernel" --launch-skip 4 -c 2 -o gpurun_out/r02_field_half -f python profiles/prof_field_split.py 1 disc > gpurun_out/r02_field_half.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"field_half|field_rec_kernel" --launch-skip 2 -c 1 -o gpurun_out/r02_field_fused -f python profiles/prof_field_split.py 0 disc > gpurun_out/r02_field_fused.log 2>&1
python profiles/ncu_extract.py gpurun_out/r02_field_half.ncu-rep gpurun_out/r02_ncu_full_field_half.csv
python profiles/ncu_extract.py gpurun_out/r02_field_fused.ncu-rep gpurun_out/r02_ncu_full_field_fused.csv
