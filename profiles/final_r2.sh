# Round-2 final evidence run on one B200: GPU tests, smoke, bench line + reference arm (same box), ncu launch list of the bench command.
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/final_tests.log 2>&1; tail -3 gpurun_out/final_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
python bench.py --impl reference > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; cut -c1-160 gpurun_out/final_bench_ref.json
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -c 300 gpurun_out/final_bench.err
LDPC_B200_PAIR=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_bench_r2b.csv python bench.py --steps 2 --warmup 3 > gpurun_out/final_ncu_bench.log 2>&1
grep -c tile4 gpurun_out/launches_bench_r2b.csv
python profiles/configs_table.py gpurun_out/final_bench.json | cut -c1-200 | head -11
