#!/bin/bash
python bench.py --quick --steps 5 --warmup 2 --methods elastic,fluid 2>&1 | tail -1
python -m pytest tests/test_relaxed_gpu.py tests/test_fullsize_gpu.py tests/test_configs_gpu.py tests/test_batch_gpu.py tests/test_engine_gpu.py -m gpu -q -x 2>&1 | tail -6
