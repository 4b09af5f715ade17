python -m pytest tests/test_gpu_tracker.py tests/test_gpu_gemm_tc.py -x -q -k "throughput_gemm_forms or kernel_forms_agree or cfg4_full_size or multi_target_equals or gemm or teacher_forced" > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2f_pytest.log
for i in 1 2; do
for lib in tools/_ab/lib_prev.so ""; do
  echo "lib=$lib"
  VT_B200_LIB=$PWD/$lib; [ -z "$lib" ] && unset VT_B200_LIB || export VT_B200_LIB
  timeout 300 python tools/bench_configs.py cfg4 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg4', d['frames_per_s'], d['stages_ms']['vit_ms'])"
  timeout 300 python tools/bench_configs.py cfg1 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg1', d['frames_per_s'], d['stages_ms']['vit_ms'])"
done; done
