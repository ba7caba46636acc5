#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_decoder.py -q -m gpu -k "partial_label" 2>&1 | tail -3
