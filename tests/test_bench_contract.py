"""bench.py's reference arm runs on the host cores, so its side of the measurement contract can be
checked without a GPU: one JSON line, the keys the driver reads, the same metric/config as the
B200 arm, and `--gpus N` ranks other than 0 exiting quietly."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(args, env=None):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                       timeout=600, env=dict(os.environ, **(env or {})))
    assert r.returncode == 0, r.stderr[-2000:]
    return [l for l in r.stdout.splitlines() if l.strip()]


def test_reference_arm_prints_one_contract_line():
    lines = run(["--impl", "reference", "--steps", "1", "--warmup", "0"])
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "volume_render_mrays_per_s" and d["unit"] == "Mrays/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["config"]["workload"].startswith("c2:") and d["config"]["image"] == [1920, 1080]
    assert d["value"] > 0 and abs(d["value"] - 1920 * 1080 / (d["ms_per_step"] * 1e-3) / 1e6) < 1e-6 * d["value"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and d["dtype"] == "f32"
    # the same config object as the B200 arm prints, steps/warmup honoured as given
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.config_of(bench.workload_c2(), 1) and d["warmup"] == 0


def test_reference_arm_imports_nothing_of_the_product_and_uses_all_cores():
    """the reference arm must not load libvr_b200.so, and must not be throttled to one
    thread by the launcher's OMP_NUM_THREADS=1."""
    code = ("import sys, json, os; sys.argv=['bench.py','--impl','reference','--steps','1','--warmup','0'];"
            "import runpy; runpy.run_path(%r, run_name='__main__');"
            "bad=[m for m in sys.modules if m.startswith('ascent_b200')];"
            "maps=open('/proc/self/maps').read();"
            "print('MODS', bad, 'libvr' in maps)" % os.path.join(ROOT, "bench.py"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    d = json.loads(lines[0])
    assert lines[1] == "MODS [] False", lines[1]
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))


def test_workload_follows_the_launch_style():
    """plain --gpus 1 is config 2; anything torchrun starts (N = 1 included) is config 3, so that a
    1/2/4/8 sweep is one workload"""
    sys.path.insert(0, ROOT)
    import bench
    env = {k: os.environ.pop(k) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "TORCHELASTIC_RUN_ID") if k in os.environ}
    try:
        assert not bench.under_torchrun()
        os.environ.update(RANK="0", WORLD_SIZE="1", LOCAL_RANK="0")
        assert bench.under_torchrun()
    finally:
        for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
            os.environ.pop(k, None)
        os.environ.update(env)
    c = bench.config_of(bench.workload_c3(), 8)
    assert c["path"].startswith("A") and bench.config_of(bench.workload_c3(), 4)["path"].startswith("B")


def test_reference_arm_ranks_other_than_zero_exit_without_work():
    assert run(["--impl", "reference", "--gpus", "2", "--steps", "1"], env={"RANK": "1", "WORLD_SIZE": "2"}) == []
