# Round-2 measurement session on one B200 (run under gpurun; results land in gpurun_out/).
set -x
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -1 gpurun_out/r02_smoke.log
python bench.py > gpurun_out/r02_final_train40960.json 2> gpurun_out/r02_final_train40960.err
cp gpurun_out/bench_train40960_n1.json gpurun_out/r02_side_train40960_n1.json
python bench.py --impl reference --steps 3 > gpurun_out/r02_final_reference.json 2>/dev/null
for w in train2500 infer16k infer64k infer256k knn1m_k16 knn1m_k32 predict160k; do
  python bench.py --workload $w > gpurun_out/r02_final_$w.json 2> gpurun_out/r02_final_$w.err
  cp gpurun_out/bench_${w}_n1.json gpurun_out/r02_side_${w}_n1.json
done
R3D_TC_OFF=1 python bench.py --workload train2500 --no-cpu-baseline > gpurun_out/r02_train2500_tc_off.json 2>/dev/null
R3D_TC_OFF=1 python bench.py --no-cpu-baseline > gpurun_out/r02_train40960_tc_off.json 2>/dev/null
python tools/pw_cl_bench.py > gpurun_out/r02_pw_cl_bench.txt 2>&1
for shp in 1,1048576,1048576,16 1,1048576,1048576,32 64,40960,40960,16; do for v in 2 3; do python tools/knn_bench.py --shape $shp --variant $v --iters 2; done; done > gpurun_out/r02_knn_brute_variants.txt 2>&1
python bench.py --profile-steps 1 --warmup 3 > gpurun_out/r02_prof_plain3.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_train40960_final.csv python bench.py --profile-steps 1 --warmup 3 > gpurun_out/r02_ncu_ll2.log 2>&1
ncu --set full --clock-control none --profile-from-start off -k regex:"lfa_cl_(bwd|fwd|wide)|pc_(gemm|wgrad)" -c 90 --csv --page raw --log-file gpurun_out/r02_ncu_lfa_cl_raw.csv python bench.py --profile-steps 1 --warmup 3 > gpurun_out/r02_ncu_full2.log 2>&1
tail -2 gpurun_out/r02_ncu_full2.log
grep -h '"metric"' gpurun_out/r02_final_*.json gpurun_out/r02_train2500_tc_off.json | cut -c1-200
ls -la gpurun_out | head -40
