mkdir -p gpurun_out
PT="python -m pytest -m gpu -q -x -p no:cacheprovider --timeout=300 --timeout-method=thread"
timeout 600 $PT tests/test_gpu_texture.py tests/test_gpu_boundaries.py > gpurun_out/t_tex.log 2>&1; echo "texture rc=$?"; tail -n 25 gpurun_out/t_tex.log | cut -c1-250
