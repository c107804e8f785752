python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "gru or trunk" 2>&1 | tail -5
python tools/trace_gru.py 256 32 801
python tools/trace_gru.py 256 16 801
python tools/trace_gru.py 384 4 1001
python tools/trace_gru.py 384 2 1001
python tools/profile_layers.py 2>&1 | sed -n 1,40p > gpurun_out/r2p_layers.txt
grep -E "gru|dec.4.trunk|score step" gpurun_out/r2p_layers.txt
