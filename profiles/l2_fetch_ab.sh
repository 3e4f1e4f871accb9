export LDPC_B200_TUNE_CACHE=off 
for g in 64 32 128; do for sh in "1 512" "2 512" "4 512"; do set -- $sh; echo "fetch $g lanes $1: $(LDPC_B200_L2_FETCH=$g python profiles/profile_large.py codes/dvbs2_like_r12_n64800.txt 2368 1.0 $1 $2 BP 2>&1 | tail -2 | head -1 | grep -o "device_ms[^}]*")"; done; done
for g in 64 32; do echo "bg1 ms fetch $g: $(LDPC_B200_L2_FETCH=$g python profiles/profile_large.py codes/nr_bg1_like_z384.txt 4736 -0.5 1 512 BP_MS 2>&1 | tail -2 | head -1 | grep -o "device_ms[^}]*")"; done
