import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import __graft_entry__ as g
ca = g.load_package()
sx, sy, sz, tx, tz = ca.halfbox_symmetries()
print(json.dumps(bench.hbm_case(ca, [sx * sy * sz, sz * tx * tz], 6537.6)))
