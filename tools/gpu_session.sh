# Final verification of a round on the GPU box (usage: gpurun --timeout 3600 -- 'bash tools/gpu_session.sh'):
# the driver's own sequence -- GPU tests, smoke, both bench arms -- plus the ncu launch list of the
# same bench command.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/final_pytest.log; tail -3 gpurun_out/final_pytest.log | cut -c1-300
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1; tail -2 gpurun_out/final_smoke.log | cut -c1-300
timeout 900 python bench.py --impl reference > gpurun_out/final_bench_ref.log 2>&1; tail -c 400 gpurun_out/final_bench_ref.log
timeout 1500 python bench.py > gpurun_out/final_bench_1b.log 2>&1; tail -c 3000 gpurun_out/final_bench_1b.log | cut -c1-1500
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name 'regex:^(tables_|bins[0-9]|adc_|rank2_|rerank_|lut_|dispatch_)' -c 60 --csv --log-file gpurun_out/final_ncu_launches_1b.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --variants knn > gpurun_out/final_ncu_bench.log 2>&1
