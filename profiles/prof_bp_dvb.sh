cd $GRAFT_REPO_ROOT
export LDPC_B200_TUNE_CACHE=off
for sh in "1 512" "2 512" "4 512" "1 256" "2 256"; do set -- $sh; python profiles/profile_large.py codes/dvbs2_like_r12_n64800.txt 1184 1.0 $1 $2 BP 2>&1 | tail -2 | head -1 | cut -c1-200; done
ncu --set full --import-source on --clock-control none -k regex:tile4 --launch-skip 1 -c 1 -o gpurun_out/r2_bp_dvb_ed -f python profiles/profile_large.py codes/dvbs2_like_r12_n64800.txt 1184 1.0 1 512 BP > gpurun_out/r2_ncu_bp_dvb_ed.log 2>&1
python profiles/ncu_summary.py gpurun_out/r2_bp_dvb_ed.ncu-rep > gpurun_out/ncu_bp_dvb_ed_summary.txt 2>&1
ncu -i gpurun_out/r2_bp_dvb_ed.ncu-rep --page source --csv > gpurun_out/r2_bp_dvb_ed_source.csv 2>/dev/null
python profiles/hot_sass.py gpurun_out/r2_bp_dvb_ed_source.csv > gpurun_out/hot_sass_bp_dvb_ed.txt 2>&1
cat gpurun_out/ncu_bp_dvb_ed_summary.txt
