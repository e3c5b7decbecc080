python tools/diag_k2.py 2>&1 | tail -8
