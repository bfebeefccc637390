cd $GRAFT_REPO_ROOT
# round-2 ncu --set full captures (one GPU): particle / grid-stage / set-up kernels of substep 2, and one whole CG iteration (V-cycle) of substep 1
KA='regex:k_p2g|k_liquid_sdf|k_advect_particles|k_apply_pressure|k_cell_place|k_pcg_resident|k_visc_rows|k_visc_volumes|k_gmg_build_g'
timeout 500 ncu --set full --clock-control none --import-source on -k "$KA" -s 14 -c 14 -f -o gpurun_out/r2p_stages python tests/gpu_dev_gmg.py 256 2 2 > gpurun_out/r2p_stages.log 2>&1; echo ncuA rc=$?
ncu -i gpurun_out/r2p_stages.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/r2p_stages_raw.csv.gz
KB='regex:k_gmg_sweep_tma|k_gmg0_sweep|k_gmg0_prolong|k_gmg_restrict|k_gmg_prolong|k_gmg_dense_apply|k_visc_apply|k_cg_update|k_cg_dot|k_cg_direction'
timeout 500 ncu --set full --clock-control none --import-source on -k "$KB" -s 42 -c 44 -f -o gpurun_out/r2p_vcycle python tests/gpu_dev_gmg.py 256 2 1 > gpurun_out/r2p_vcycle.log 2>&1; echo ncuB rc=$?
ncu -i gpurun_out/r2p_vcycle.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/r2p_vcycle_raw.csv.gz
python dev/ncu_summary.py gpurun_out/r2p_stages_raw.csv.gz > gpurun_out/r2p_stages_ncu.csv
python dev/ncu_summary.py gpurun_out/r2p_vcycle_raw.csv.gz > gpurun_out/r2p_vcycle_ncu.csv
ls -la gpurun_out/r2p_*; tail -n 3 gpurun_out/r2p_stages.log gpurun_out/r2p_vcycle.log; cut -d, -f1,4 gpurun_out/r2p_stages_ncu.csv | head -20
