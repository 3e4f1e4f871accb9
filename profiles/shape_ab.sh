export LDPC_B200_TUNE_CACHE=off
for lib in libldpc_prev.so libldpc.so; do for pair in 0 1; do
echo "$lib pair=$pair: $(LDPC_B200_LIB=$PWD/libldpc_b200/$lib LDPC_B200_PAIR=$pair python profiles/profile_cmd.py bp 37888 2>&1 | tail -2 | head -1 | grep -o "device_ms.*") | ET+2: $(LDPC_B200_LIB=$PWD/libldpc_b200/$lib LDPC_B200_PAIR=$pair SNR=2 python profiles/profile_cmd.py bpet 151552 2>&1 | tail -2 | head -1 | grep -o "device_ms.*")"
done; done
