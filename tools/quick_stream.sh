python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "first_order or fform_closed or sqp" 2>&1 | tail -12
