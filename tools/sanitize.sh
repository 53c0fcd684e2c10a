#!/bin/bash
# compute-sanitizer memcheck over the kernel unit tests of this round's new kernels (small shapes; ONE tool per gpurun call):
#     gpurun --timeout 1200 -- 'bash tools/sanitize.sh'
OUT=gpurun_out
mkdir -p $OUT
SEL="scatter or split_precision or ensemble or reader_device or bit_identical or module_forward_backward"
timeout 300 python -m pytest tests/test_gpu_model.py tests/test_gpu_kernels.py tests/test_gpu_reader_infer.py -m gpu -q -x -k "$SEL" > $OUT/sanitize_plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/sanitize_plain.log; exit 1; }
timeout 1000 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file $OUT/sanitize_memcheck.log \
    python -m pytest tests/test_gpu_model.py tests/test_gpu_kernels.py tests/test_gpu_reader_infer.py -m gpu -q -x -k "$SEL" > $OUT/sanitize_pytest.log 2>&1
echo "memcheck exit code $?"
tail -3 $OUT/sanitize_pytest.log
grep -c "ERROR SUMMARY" $OUT/sanitize_memcheck.log; grep "ERROR SUMMARY" $OUT/sanitize_memcheck.log | tail -2
