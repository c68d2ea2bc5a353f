#!/usr/bin/env python
"""Launch lists at HEAD from the ncu captures merged back under gpurun_out/: launches_r2_net2.csv (whole-network forward,
tools/time_pwc.py --tc --B 8 --once) and launches_r2_train.csv (one training forward + backward plan,
tools/time_train.py --tc --once)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import importlib.util
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
src = open(os.path.join(ROOT, "tools", "make_profiles_r2.py")).read().split("\nnet = launches(")[0]
ns = {"__file__": os.path.join(ROOT, "tools", "make_profiles_r2.py")}
exec(compile(src, "make_profiles_r2.py", "exec"), ns)
launches, write_list = ns["launches"], ns["write_list"]


def ours(path):
    return [x for x in launches(path) if "b2f::" in x[3] or "unnamed>::" in x[3]]


net = ours(os.path.join(G, "launches_r2_net2.csv"))
step = net[len(net) - len(net) // 2:]
t1 = write_list("r02_ncu_launch_list_network.txt",
                "# ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python tools/time_pwc.py --tc --B 8 --once\n"
                "# second of two eager whole-network forwards (Ours-Hard, 8 x 9 x 448 x 1024; decoders incl. heads and the\n"
                "# stride-1 pyramid layers with >= 64 channels on tcgen05)\n", step)
tr = ours(os.path.join(G, "launches_r2_train.csv"))
# the capture holds: first forward + backward (time_train.py's warm-up through net.forward / net.backward), then --once's
# plan launches; the last half is one forward + backward plan
half = tr[len(tr) - len(tr) // 2:]
t2 = write_list("r02_ncu_launch_list_train.txt",
                "# ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 python tools/time_train.py --tc --once\n"
                "# one forward + backward plan of the training step (Ours-Hard, 8 x 9 x 320 x 640, tensor-core path), eager and\n"
                "# serialised by ncu: in the step these run as ONE graph over several streams (DESIGN 4.7)\n", half)
print("network forward %.1f us over %d kernels; training forward + backward %.1f us over %d kernels" % (t1, len(step), t2, len(half)))
