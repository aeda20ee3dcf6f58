set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r1g_pytest.log
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r1g_bench_ref.json 2> gpurun_out/r1g_bench_ref.err
python bench.py > gpurun_out/r1g_bench.json 2> gpurun_out/r1g_bench.err
python bench.py --steps 10 --warmup 20 --no-cpu > gpurun_out/r1g_plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r1g_launches.csv python bench.py --steps 10 --warmup 20 --no-cpu > gpurun_out/r1g_ncu_bench.log 2>&1
python scripts/profile_step.py 1000000 8 100 > gpurun_out/prof_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_repulse_pairs|k_attract_update" -s 297 -c 3 -o gpurun_out/prof_r1g -f python scripts/profile_step.py 1000000 8 100 > gpurun_out/prof_ncu.log 2>&1
cat gpurun_out/r1g_pytest.log; cat gpurun_out/r1g_bench.json | cut -c1-400
timeout 120 python scripts/gpu_ab.py c5 20 10 wembed_b200/lib/libwembed_b200.so > gpurun_out/r1g_c5.log 2>&1
timeout 110 python scripts/gpu_ab.py c4 20 60 wembed_b200/lib/libwembed_b200.so > gpurun_out/r1g_c4.log 2>&1
cat gpurun_out/r1g_c5.log gpurun_out/r1g_c4.log | tail -4
