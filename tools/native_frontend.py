"""prints the C++-adapter front-end timings (adapter/frontend_bench) -- the 'native_cpp' object of bench.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
print(json.dumps(bench.native_frontend_numbers(int(sys.argv[1]) if len(sys.argv) > 1 else 300, 20)))
