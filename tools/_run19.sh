timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -k "odd_shapes" 2>&1 | grep -v "^  File" | tail -5
