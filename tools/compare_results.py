"""Diff two batch result files (rsruckig_b200/batchfile.py) at north_star's bar: codes exact, durations <= 1e-9 relative.
Usage: compare_results.py a.rkb2 b.rkb2          e.g. ours (tools/dump_results.py) against tools/rsruckig_dump's."""
import json
import sys

sys.path.insert(0, ".")
from rsruckig_b200.batchfile import compare_results, read_results  # noqa: E402

r = compare_results(read_results(sys.argv[1]), read_results(sys.argv[2]))
print(json.dumps(r, indent=1))
sys.exit(0 if r["pass"] else 1)
