#!/bin/bash
timeout 900 python -m pytest tests/test_metric_shape_gpu.py -m gpu -q -k "graph_replay" 2>&1 | tail -4
