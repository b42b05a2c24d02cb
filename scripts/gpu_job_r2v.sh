#!/bin/bash
timeout 200 python -m pytest tests/test_render_grad.py -x -q -m gpu -k "planes_only" 2>&1 | tail -4 | cut -c1-500
