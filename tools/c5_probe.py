"""C5 probe: time-to-R-hat<1.2 for batched independent HydroModel runs (bench.convergence_c5)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, 'tests')
import torch, bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
print(json.dumps(bench.convergence_c5(0, torch, n_runs=n)))
