set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"field_rec_kernel" --launch-skip 5 -c 1 -o gpurun_out/r02_field_rec_final_disc -f python profiles/prof_field_split.py 0 disc 16777216 > gpurun_out/r02_field_rec_final_disc.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"field_rec_kernel" --launch-skip 5 -c 1 -o gpurun_out/r02_field_rec_final_halo -f python profiles/prof_field_split.py 0 halo 16777216 > gpurun_out/r02_field_rec_final_halo.log 2>&1
python profiles/ncu_extract.py gpurun_out/r02_field_rec_final_disc.ncu-rep gpurun_out/r02_ncu_full_field_rec_final_disc.csv
python profiles/ncu_extract.py gpurun_out/r02_field_rec_final_halo.ncu-rep gpurun_out/r02_ncu_full_field_rec_final_halo.csv
