"""Development aid: one small team-engine solve against the oracle (MCF_TEAM_PRICERS forces the number of pricing CTAs)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import importlib.util
spec = importlib.util.spec_from_file_location('gc', os.path.join(ROOT, 'tools', 'gpu_check.py')); gc = importlib.util.module_from_spec(spec)
sys.argv = ['x', 'none']
spec.loader.exec_module(gc)
from mincostflow_b200 import instances
for k in [int(x) for x in (os.environ.get("NP_TEST_K", "10,13")).split(",")]:
    p = instances.netgen8(k)
    for np_ in [int(x) for x in os.environ.get("NP_TEST_PRICERS", "0").split(",")]:
        gc.run(p, 2, stop=int(os.environ.get("NP_TEST_STOP", "0")) or None, lookahead=np_ or None, oracle_too=k <= 16)
