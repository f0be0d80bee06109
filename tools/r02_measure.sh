# Round-2 measurement session on one B200 (run under gpurun; results land in gpurun_out/).
# The TC-off comparison runs, tools/pw_cl_bench.py and the brute-force KNN variant table were taken earlier in the round
# (profiles/r02_bench_train40960_tc_off.json, r02_pw_cl_bench.txt, r02_knn_brute_variants.txt) and are not repeated here.
set -x
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -1 gpurun_out/r02_smoke.log
python bench.py > gpurun_out/r02_final_train40960.json 2> gpurun_out/r02_final_train40960.err
cp gpurun_out/bench_train40960_n1.json gpurun_out/r02_side_train40960_n1.json
python bench.py --impl reference --steps 3 > gpurun_out/r02_final_reference.json 2>/dev/null
for w in train2500 infer16k infer64k infer256k knn1m_k16 knn1m_k32 predict160k train40960_b8; do
  python bench.py --workload $w > gpurun_out/r02_final_$w.json 2> gpurun_out/r02_final_$w.err
  cp gpurun_out/bench_${w}_n1.json gpurun_out/r02_side_${w}_n1.json
done
python bench.py --profile-steps 1 --warmup 3 > gpurun_out/r02_prof_plain3.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_train40960_final.csv python bench.py --profile-steps 1 --warmup 3 > gpurun_out/r02_ncu_ll2.log 2>&1
ncu --set full --clock-control none --profile-from-start off -k regex:"lfa_cl_(bwd|fwd|wide)" -c 26 --csv --page raw --log-file gpurun_out/r02_ncu_lfa_cl_raw.csv python bench.py --profile-steps 1 --warmup 3 > gpurun_out/r02_ncu_full2.log 2>&1
tail -2 gpurun_out/r02_ncu_full2.log
grep -h '"metric"' gpurun_out/r02_final_*.json | cut -c1-200
