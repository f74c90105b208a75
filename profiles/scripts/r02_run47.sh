set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pt_key_kernel|key_scan_kernel|rec_scatter_kernel|field_gather_kernel" --launch-skip 8 -c 4 -o gpurun_out/r02_support -f python profiles/prof_field_split.py 0 halo 16777216 > gpurun_out/r02_support.log 2>&1
python profiles/ncu_extract.py gpurun_out/r02_support.ncu-rep gpurun_out/r02_ncu_full_support_kernels.csv
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_points_final2.csv python profiles/prof_field_split.py 0 halo 16777216 > /dev/null 2>&1
