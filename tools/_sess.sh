set -u
timeout 600 python -m pytest tests/test_gpu_deform.py -x -q -k "six_word" 2>&1 | tail -5
