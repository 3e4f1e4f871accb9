cd $GRAFT_REPO_ROOT
export LDPC_B200_TUNE_CACHE=off
# h.txt, fp64 sum-product, 50 fixed iterations; LDPC_B200_PAIR presets the shape (1: two CTAs per SM x 192 threads, 0: one CTA x 384 threads)
LDPC_B200_PAIR=${1:-1} ncu --set full --import-source on --clock-control none -k regex:tile4 --launch-skip 2 -c 1 -o gpurun_out/r2_bp_ed -f python profiles/profile_cmd.py bp 37888 > gpurun_out/r2_ncu_bp_ed.log 2>&1
python profiles/ncu_summary.py gpurun_out/r2_bp_ed.ncu-rep > gpurun_out/ncu_bp_ed_summary.txt 2>&1
ncu -i gpurun_out/r2_bp_ed.ncu-rep --page source --csv > gpurun_out/r2_bp_ed_source.csv 2>/dev/null
python profiles/hot_sass.py gpurun_out/r2_bp_ed_source.csv > gpurun_out/hot_sass_bp_ed.txt 2>&1
cat gpurun_out/ncu_bp_ed_summary.txt
