python -m pytest tests/test_gpu_tracker.py -x -q -k "throughput_gemm_forms or cfg4_full_size" > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2e_pytest.log
rm -f gpurun_out/r2e_cfg4.jsonl
for v in "VT_B200_NO_AS_MLP=1" "VT_X=1"; do echo "$v"; env $v timeout 300 python tools/bench_configs.py cfg4 2>&1 | tail -1 | tee -a gpurun_out/r2e_cfg4.jsonl; done
timeout 300 python tools/chain_timeline.py --targets 16 --frames 8 --events > gpurun_out/r2e_timeline_new.txt 2>&1
VT_B200_NO_AS_MLP=1 timeout 300 python tools/chain_timeline.py --targets 16 --frames 8 --events > gpurun_out/r2e_timeline_nomlp.txt 2>&1
tail -12 gpurun_out/r2e_timeline_new.txt
