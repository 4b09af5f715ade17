for i in 1 2 3 4 5; do python -m pytest tests/test_gpu_context.py -x -q 2>&1 | tail -1; done
python -m pytest tests -x -q -m gpu > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2k_pytest.log
python bench.py > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2k_bench.err
python bench.py --impl reference --steps 30 --warmup 3 > gpurun_out/r2k_bench_ref.json 2>> gpurun_out/r2k_bench.err; echo "ref rc=$?"
