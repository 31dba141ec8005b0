import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import coregex_b200 as cg
from oracle_lib import Oracle
for n in (20000, 5000, 5290, 5291, 5292, 10581, 10582):
    hay = b"foobar" * n
    r = cg.Compile("foo|bar")
    got = r.find_all_index_array(hay)
    want = Oracle("foo|bar").find_all(np.frombuffer(hay, dtype=np.uint8))
    if got.shape != want.shape:
        gs = set(map(tuple, got.tolist()))
        miss = [w for w in want.tolist() if tuple(w) not in gs]
        print(n, len(hay), got.shape, want.shape, "missing", miss[:5], "count", r.Count(hay))
    else:
        print(n, len(hay), "ok", np.array_equal(got, want))
