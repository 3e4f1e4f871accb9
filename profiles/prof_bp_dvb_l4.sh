cd $GRAFT_REPO_ROOT
export LDPC_B200_TUNE_CACHE=off
ncu --set full --import-source on --clock-control none -k regex:tile4 --launch-skip 1 -c 1 -o gpurun_out/r2_bp_dvb_ed_l4 -f python profiles/profile_large.py codes/dvbs2_like_r12_n64800.txt 1184 1.0 4 512 BP > gpurun_out/r2_ncu_bp_dvb_ed_l4.log 2>&1
python profiles/ncu_summary.py gpurun_out/r2_bp_dvb_ed_l4.ncu-rep > gpurun_out/ncu_bp_dvb_ed_l4_summary.txt 2>&1
ncu -i gpurun_out/r2_bp_dvb_ed_l4.ncu-rep --page source --csv > gpurun_out/r2_bp_dvb_ed_l4_source.csv 2>/dev/null
cat gpurun_out/ncu_bp_dvb_ed_l4_summary.txt
for f in 1184 2368 4736; do python profiles/profile_large.py codes/dvbs2_like_r12_n64800.txt $f 1.0 4 512 BP 2>&1 | tail -2 | head -1 | cut -c1-120; done
for f in 2368; do python profiles/profile_large.py codes/dvbs2_like_r12_n64800.txt $f 1.0 4 512 BP_MS 2>&1 | tail -2 | head -1 | cut -c1-120; python profiles/profile_large.py codes/dvbs2_like_r12_n64800.txt $f 1.0 1 512 BP_MS 2>&1 | tail -2 | head -1 | cut -c1-120; done
