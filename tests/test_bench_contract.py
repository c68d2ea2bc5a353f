"""CPU-side checks of bench.py's contract with the driver: the call order follows the network's dataflow, and the
reference arm (the CPU restatement on the host cores) runs without a GPU and prints ONE JSON line with the keys the
driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_network_order_respects_the_dataflow_of_pwc_lua():
    sys.path.insert(0, ROOT)
    import bench
    order = bench.Workload.network_order()
    assert len(order) == 56 and len(set(order)) == 56           # 10 CV + 8 feature warps + 10 image warps, each way
    pos = {k: i for i, k in enumerate(order)}
    for d in (0, 1):
        for l in (7, 6, 5, 4, 3):
            if l > 3:   # forward: CV_l -> feature warp L(l-1) -> CV_(l-1)   (pwc.lua:247-263, 402-408)
                assert pos[("f", ("cv", l, d))] < pos[("f", ("fw", l - 1, d))] < pos[("f", ("cv", l - 1, d))]
            assert pos[("f", ("cv", l, d))] < pos[("f", ("iw", l, d))]            # image warp needs the level's flow
            # backward mirrors it: image-warp backward of level l before CV_l backward, CV_l before its feature warp
            assert pos[("b", ("iw", l, d))] < pos[("b", ("cv", l, d))]
            if l <= 6:
                assert pos[("b", ("cv", l, d))] < pos[("b", ("fw", l, d))] < pos[("b", ("cv", l + 1, d))]
    assert max(pos[k] for k in pos if k[0] == "f") < min(pos[k] for k in pos if k[0] == "b")


def test_reference_arm_prints_one_json_line_without_a_gpu():
    # OMP_NUM_THREADS=1 is what torchrun exports for N > 1: the arm must still use every core it may run on
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, env=env, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "frame_triplets_per_sec_1024x448" and d["unit"] == "triplets/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert d["e2e"] == {"value": d["value"], "unit": "triplets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_gpu_local_cpu_list_parsing_is_defensive():
    """bench._gpu_local_cpus returns None (no binding) when there is no GPU / no sysfs entry instead of raising."""
    sys.path.insert(0, ROOT)
    import bench
    import torch
    assert bench._gpu_local_cpus(torch, torch.device("cuda:0")) is None or isinstance(
        bench._gpu_local_cpus(torch, torch.device("cuda:0")), set)
    with bench._NumaLocal(torch, torch.device("cuda:0")) as n:
        assert n.bound >= 0
    assert os.sched_getaffinity(0)   # affinity restored / untouched
