echo "== tests with IVF_SLAB_DS=2 IVF_TUNE=0"
IVF_SLAB_DS=2 IVF_TUNE=0 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_i3d.py -m gpu -x -q 2>&1 | tail -5
echo "== events ds=1 (no tune)"
IVF_TUNE=0 python tools/step_events.py 8 5 > gpurun_out/r3_ev_ds1.txt 2>&1; grep -E "k4x4x4|k3x3x3" gpurun_out/r3_ev_ds1.txt | head -50 > /dev/null; tail -10 gpurun_out/r3_ev_ds1.txt | head -4
echo "== events ds=2 (no tune)"
IVF_SLAB_DS=2 IVF_TUNE=0 python tools/step_events.py 8 5 > gpurun_out/r3_ev_ds2.txt 2>&1; tail -10 gpurun_out/r3_ev_ds2.txt | head -4
paste <(grep -E "k4x4x4|k3x3x3" gpurun_out/r3_ev_ds1.txt | cut -c1-75) <(grep -E "k4x4x4|k3x3x3" gpurun_out/r3_ev_ds2.txt | awk '{print $2}')
