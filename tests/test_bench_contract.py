"""bench.py's reference arm runs on CPU (the numpy oracle on a bounded sample): check the
one-JSON-line contract there, and the clock sampler's parsing, without a GPU."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True,
                         timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True
    assert d["unit"] == "Gcell.channel/s" and d["value"] > 0 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and d["dtype"] == "f64"


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--gpus", "2", "--steps", "1", "--warmup", "0"], capture_output=True,
                         text=True, timeout=120, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_clock_sampler_keeps_only_the_timed_window(tmp_path):
    sys.path.insert(0, ROOT)
    import bench
    s = bench.ClockSampler.__new__(bench.ClockSampler)
    s.path = str(tmp_path / "clocks.csv")
    now = time.time()

    def stamp(t):
        return time.strftime("%Y/%m/%d %H:%M:%S", time.localtime(t)) + f".{int((t % 1) * 1000):03d}"

    rows = [(now - 5.0, 500, "Not Active"), (now + 0.2, 1950, "Not Active"),
            (now + 0.4, 1965, "Active"), (now + 9.0, 300, "Not Active")]
    with open(s.path, "wt") as f:
        for t, clk, cap in rows:
            f.write(f"{stamp(t)}, {clk}, 1965, 700.0, Not Active, Not Active, Not Active, {cap}\n")
    s.f = open(s.path)           # stop() closes it
    s.t0, s.t1 = now, now + 1.0

    class Done:
        def terminate(self):
            pass

        def wait(self, timeout=None):
            return 0
    s.proc = Done()
    out = s.stop()
    assert out["samples"] == 2 and out["samples_whole_run"] == 4
    assert out["sm_mhz"] == 1957.5 and out["sm_max_mhz"] == 1965.0
    assert out["reasons"] == ["sw_power_cap"]


def test_clock_sampler_reports_the_slowest_gpu(tmp_path):
    """N-rank runs: rank 0 watches every GPU; a GPU whose clock sags shows up by index."""
    sys.path.insert(0, ROOT)
    import bench
    s = bench.ClockSampler.__new__(bench.ClockSampler)
    s.path = str(tmp_path / "clocks.csv")
    now = time.time()

    def stamp(t):
        return time.strftime("%Y/%m/%d %H:%M:%S", time.localtime(t)) + f".{int((t % 1) * 1000):03d}"

    with open(s.path, "wt") as f:
        for k in range(4):
            for gpu, clk, therm in ((0, 1965, "Not Active"), (1, 1500, "Active")):
                f.write(f"{stamp(now + 0.1 + 0.1 * k)}, {clk}, 1965, 700.0, Not Active, "
                        f"Not Active, {therm}, Not Active, {gpu}\n")
    s.f = open(s.path)
    s.t0, s.t1 = now, now + 1.0

    class Done:
        def terminate(self):
            pass

        def wait(self, timeout=None):
            return 0
    s.proc = Done()
    out = s.stop()
    assert out["samples"] == 8
    assert out["per_gpu_sm_mhz"] == {"0": 1965.0, "1": 1500.0}
    assert out["sm_mhz_slowest_gpu"] == 1500.0 and out["sm_max_mhz"] == 1965.0
    assert out["reasons"] == ["sw_thermal_slowdown"]
