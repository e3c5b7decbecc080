python tools/api_profile.py 2>&1 | grep -E "dxb200|ms per call" | head -40
