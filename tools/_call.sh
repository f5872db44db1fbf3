timeout 900 python -m pytest tests/test_gpu_reference.py -m gpu -q -s 2>&1 | grep "^\[\|AssertionError\|passed\|failed\|Error" | cut -c1-300
