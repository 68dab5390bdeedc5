ncu --set full --clock-control none --import-source on -k regex:'linear_bwd_fused|mean_res_bwd_tile|gat_agg_bwd_tile_pipe' -c 6 -o gpurun_out/prof_r4g_new_kernels -f python bench.py --profile-kernels > gpurun_out/r4g_ncu.log 2>&1
tail -3 gpurun_out/r4g_ncu.log | cut -c1-200
