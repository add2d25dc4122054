"""BASELINE metric (ii) at a shape where one thread sample is large: k=50,
10^5 compressed sites, 1 iteration, reference binary vs drop-in binary."""
import argparse
import json
import sys

sys.path.insert(0, ".")
import bench  # noqa: E402

a = argparse.Namespace(k=50, sites=100000, ntimes=20)
print(json.dumps(bench.mcmc_pair(a, 1)))
