set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"field_rec_kernel" --launch-skip 2 -c 1 -o gpurun_out/r02_field_rec_disc_final -f python profiles/prof_field_split.py 0 disc 4194304 > gpurun_out/r02_field_rec_disc_final.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"field_rec_kernel" --launch-skip 2 -c 1 -o gpurun_out/r02_field_rec_halo_final -f python profiles/prof_field_split.py 0 halo 4194304 > gpurun_out/r02_field_rec_halo_final.log 2>&1
python profiles/ncu_extract.py gpurun_out/r02_field_rec_disc_final.ncu-rep gpurun_out/r02_ncu_full_field_rec_disc_final.csv
python profiles/ncu_extract.py gpurun_out/r02_field_rec_halo_final.ncu-rep gpurun_out/r02_ncu_full_field_rec_halo_final.csv
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_points_final.csv python profiles/prof_field_split.py 0 halo 16777216 > gpurun_out/r02_launches_points_final.log 2>&1
