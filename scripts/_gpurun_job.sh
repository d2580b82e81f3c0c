timeout 900 python -m pytest tests/test_forward_gpu.py -x -q -k "postprocess_modes" 2>&1 | tail -6
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_align_gpu.py -x -q -k "first_iterations or ragged_image or entry_window or fx_and_fy" 2>&1 | tail -6
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_forward_gpu.py tests/test_scene_ops_gpu.py -x -q -k "attention or kernel_matches" 2>&1 | tail -6
