#!/bin/bash
# round-2, session 3x: compute-sanitizer memcheck over the new kernels (queue kernels, tiled grid kernel, LAVD slabs / separable means)
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest \
    "tests/test_gpu_queue.py::test_queue_kernel_masked_grid" "tests/test_gpu_queue.py::test_queue_kernel_point_list" \
    "tests/test_gpu_queue.py::test_grid_launch_shapes_bit_identical" \
    "tests/test_gpu_parity.py::test_lavd_time_collapsed_vorticity" "tests/test_gpu_parity.py::test_lavd_cubic_spline" \
    -m gpu -q -x > gpurun_out/r3x_memcheck.txt 2>&1
echo "exit code $?" >> gpurun_out/r3x_memcheck.txt
tail -15 gpurun_out/r3x_memcheck.txt | cut -c1-200
