mkdir -p gpurun_out
python -m pytest tests/test_geometry_surfaces_gpu.py tests/test_geometry_gpu.py -x -q > gpurun_out/t_geom.log 2>&1; tail -15 gpurun_out/t_geom.log
