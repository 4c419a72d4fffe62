"""The 'leaf count <= 50' reading of C5 alone (bench.py's c5_leaf50 leg): timing, or a target for ncu."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
args = bench.parse()
import __graft_entry__
__graft_entry__.build()
print(json.dumps(bench.leaf50_leg(args, 0)), file=sys.stderr)
