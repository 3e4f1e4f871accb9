set -x
cd $GRAFT_REPO_ROOT
export LDPC_B200_TUNE_CACHE=off
# 1. headline kernel (two-CTA shape preset), fixed 50 iterations
PAIR=1 LDPC_B200_PAIR=1 ncu --set full --import-source on --clock-control none -k regex:tile4 --launch-skip 2 -c 1 -o gpurun_out/r2_ms_v7 -f python profiles/profile_cmd.py ms 4736 > gpurun_out/r2_ncu_ms.log 2>&1
python profiles/ncu_summary.py gpurun_out/r2_ms_v7.ncu-rep > gpurun_out/ncu_ms_v7_summary.txt 2>&1
ncu -i gpurun_out/r2_ms_v7.ncu-rep --page source --csv > gpurun_out/r2_ms_v7_source.csv 2>/dev/null
python profiles/hot_sass.py gpurun_out/r2_ms_v7_source.csv > gpurun_out/hot_sass_ms_v7.txt 2>&1
# 2. early-termination instantiation at -4.5 dB and +3 dB
PAIR=1 LDPC_B200_PAIR=1 SNR=3 ncu --set full --clock-control none -k regex:tile4 --launch-skip 2 -c 1 -o gpurun_out/r2_et3b -f python profiles/profile_cmd.py et 59200 > gpurun_out/r2_ncu_et3b.log 2>&1
python profiles/ncu_summary.py gpurun_out/r2_et3b.ncu-rep > gpurun_out/ncu_ms_v7_et3db_summary.txt 2>&1
# 3. bit-sliced erasure kernel (bank conflicts after the edge colouring)
ncu --set full --clock-control none -k regex:bec_slice --launch-skip 1 -c 1 -o gpurun_out/r2_bec -f python profiles/profile_bec.py > gpurun_out/r2_ncu_bec.log 2>&1
python profiles/ncu_summary.py gpurun_out/r2_bec.ncu-rep > gpurun_out/ncu_bec_slice_v2_summary.txt 2>&1
# 4. DVB-S2-shaped code, min-sum, global residency (autotuned shape of round 1: lanes 4 wide / lanes 1) 
ncu --set full --clock-control none -k regex:tile4 --launch-skip 1 -c 1 -o gpurun_out/r2_dvb_ms -f python profiles/profile_large.py codes/dvbs2_like_r12_n64800.txt 1184 1.0 4 512 BP_MS > gpurun_out/r2_ncu_dvb_ms.log 2>&1
python profiles/ncu_summary.py gpurun_out/r2_dvb_ms.ncu-rep > gpurun_out/ncu_ms_dvb_global_summary.txt 2>&1
# 5. launch list of the bench command
LDPC_B200_PAIR=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench_r2.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r2_ncu_bench.log 2>&1
tail -n 3 gpurun_out/ncu_ms_v7_summary.txt gpurun_out/ncu_bec_slice_v2_summary.txt gpurun_out/ncu_ms_dvb_global_summary.txt
grep -c tile4 gpurun_out/launches_bench_r2.csv
