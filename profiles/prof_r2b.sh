# Fresh ncu captures with the final library of round 2: headline min-sum kernel, its early-termination instantiation at +3 dB,
# min-sum on the DVB-S2-shaped code with four lanes per node (the autotuned shape).
cd $GRAFT_REPO_ROOT
export LDPC_B200_TUNE_CACHE=off
PAIR=1 LDPC_B200_PAIR=1 ncu --set full --import-source on --clock-control none -k regex:tile4 --launch-skip 2 -c 1 -o gpurun_out/r2b_ms -f python profiles/profile_cmd.py ms 4736 > gpurun_out/r2b_ncu_ms.log 2>&1
python profiles/ncu_summary.py gpurun_out/r2b_ms.ncu-rep > gpurun_out/ncu_ms_final_summary.txt 2>&1
ncu -i gpurun_out/r2b_ms.ncu-rep --page source --csv > gpurun_out/r2b_ms_source.csv 2>/dev/null
python profiles/hot_sass.py gpurun_out/r2b_ms_source.csv > gpurun_out/hot_sass_ms_final.txt 2>&1
PAIR=1 LDPC_B200_PAIR=1 SNR=3 ncu --set full --clock-control none -k regex:tile4 --launch-skip 2 -c 1 -o gpurun_out/r2b_et3 -f python profiles/profile_cmd.py et 59200 > gpurun_out/r2b_ncu_et3.log 2>&1
python profiles/ncu_summary.py gpurun_out/r2b_et3.ncu-rep > gpurun_out/ncu_ms_final_et3db_summary.txt 2>&1
ncu --set full --clock-control none -k regex:tile4 --launch-skip 1 -c 1 -o gpurun_out/r2b_dvb_ms -f python profiles/profile_large.py codes/dvbs2_like_r12_n64800.txt 1184 1.0 4 512 BP_MS > gpurun_out/r2b_ncu_dvb_ms.log 2>&1
python profiles/ncu_summary.py gpurun_out/r2b_dvb_ms.ncu-rep > gpurun_out/ncu_ms_dvb_global_l4_summary.txt 2>&1
head -12 gpurun_out/ncu_ms_final_summary.txt; head -12 gpurun_out/ncu_ms_dvb_global_l4_summary.txt
