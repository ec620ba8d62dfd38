timeout 600 python -m pytest tests/test_gpu_proposals.py -x -q -k invalid_resident 2>&1 | tail -30
