"""Large-batch roofline of the transition kernel alone (the leg bench.py reports as `transition_roofline`)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ccbs_b200.gae import GaeWeights

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=1 << 20)
ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--workload", default="c2")
a = ap.parse_args()
specs = bench.build_specs(bench.WORKLOADS[a.workload])
print(json.dumps(bench.transition_roofline(specs, GaeWeights.random(bench.GAE_SEED), 0, a.envs, a.steps)))
