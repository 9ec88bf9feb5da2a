"""CPU tests of bench.py's multi-process plumbing (gloo, world_size 2) and of its workload definitions."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(nproc, *extra, env=None):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(ROOT, "bench.py"), *extra]
    e = dict(os.environ); e.update(env or {})
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=e)


def test_two_rank_max_reduction_and_weak_scaling_value():
    r = _torchrun(2, "--selftest-dist", "--steps", "100")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, "exactly one JSON line, printed by rank 0"
    d = json.loads(lines[0])
    assert d["n_gpus"] == 2 and d["ms"] == 11.0            # MAX over ranks of (10 + rank)
    assert abs(d["value"] - 2 * 100 / 0.011) < 1e-6         # both sessions' frames over the slower rank's time
    assert d["scaling"] == "weak"


def test_single_process_selftest():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--selftest-dist", "--steps", "50"], capture_output=True, text=True, cwd=ROOT)
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][0])
    assert d["n_gpus"] == 1 and d["ms"] == 10.0 and abs(d["value"] - 5000.0) < 1e-9


def test_workload_definition_matches_baseline_config():
    sys.path.insert(0, ROOT)
    import bench
    assert (bench.W, bench.H, bench.NFEAT, bench.WIN, bench.MAXLEVEL, bench.TRAIL) == (752, 480, 150, 31, 3, 20)
    assert (bench.CHECKS, bench.UPDATES, bench.PREDICTS) == (20, 5, 10)
    assert bench.PYR_BYTES == 2_397_000 and bench.LK_BYTES == 3_688_200
    assert [bench.ekf_rows(c) for c in range(4)] == [(8, 34), (20, 55), (40, 90), (84, 160)]
    seen = [bench.frame_index(k) for k in range(1, 600)]
    assert all(abs(a - b) == 1 for a, b in zip(seen, seen[1:]))               # consecutive steps are consecutive frames
    assert min(seen) == 0 and max(seen) == bench.POOL_FRAMES - 1
    pool_mb = bench.POOL_FRAMES * 2 * bench.W * bench.H / 1e6
    assert pool_mb + 40 > 126                                                   # inputs larger than L2


def test_reference_arm_json_contract():
    """`bench.py --impl reference` (the reference's own CPU path from oracle/_ref, or the C port where that is missing) runs on host
    cores only and prints the contract's line: same metric / unit / config as the CUDA arm, impl = reference, e2e with zero copies."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, cwd=ROOT, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("stereo frames/sec") and d["unit"] == "frames/s"
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] >= 3 and d["higher_is_better"] is True and d["value"] > 0
    assert "workload" in d["config"] and d["data"] == "synthetic" and d["vs_baseline"] is None
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["sample"] and abs(cb["value"] - d["value"]) < 1e-6 * d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["unit"] == "frames/s"


def test_next_row_report_cannot_take_the_bench_line_down():
    """bench.py attaches the track-model measurement from a separate process; without a GPU (here) that process fails, and the hook
    must come back with an error entry instead of raising."""
    sys.path.insert(0, ROOT)
    import bench
    r = bench.next_row_track_model()
    assert isinstance(r, dict) and (("error" in r and "hv_ctx_create" in r["error"]) or "kernel" in r)      # "kernel": a GPU was present after all


def test_reference_arm_under_torchrun_prints_one_line():
    """N > 1: rank 0 alone runs the reference arm and prints it, the other ranks exit 0 without work."""
    r = _torchrun(2, "--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0


def test_clock_sampler_polls_nvml_in_process_and_reports_throttle_reasons(monkeypatch):
    """The timed region (~100 ms) is shorter than the start-up of an nvidia-smi child: the sampler must poll NVML itself.
    A stand-in NVML module provides the clock and the reason bit mask."""
    import time
    import types
    import pynvml as real
    sys.path.insert(0, ROOT)
    import bench
    fake = types.ModuleType("pynvml")
    for n in dir(real):
        if n.startswith("nvmlClocks") or n.startswith("NVML_CLOCK"):
            setattr(fake, n, getattr(real, n))
    seen = {}
    fake.nvmlInit = lambda: None
    fake.nvmlDeviceGetHandleByUUID = lambda u: seen.setdefault("uuid", u) and "h"
    fake.nvmlDeviceGetHandleByIndex = lambda i: "h"
    fake.nvmlDeviceGetClockInfo = lambda h, c: 1920
    fake.nvmlDeviceGetMaxClockInfo = lambda h, c: 1965
    fake.nvmlDeviceGetCurrentClocksEventReasons = lambda h: real.nvmlClocksThrottleReasonSwPowerCap | real.nvmlClocksThrottleReasonGpuIdle
    monkeypatch.setitem(sys.modules, "pynvml", fake)
    c = bench.ClockSampler(0, "1234-abcd")
    time.sleep(0.08)
    d = c.stop()
    assert seen["uuid"] == "GPU-1234-abcd"
    assert d["sm_mhz"] == 1920.0 and d["sm_max_mhz"] == 1965.0 and d["samples"] >= 3
    assert d["reasons"] == ["sw_power_cap"]                 # gpu_idle is not a slowdown reason


def test_clock_sampler_degrades_without_nvml_and_nvidia_smi(monkeypatch):
    sys.path.insert(0, ROOT)
    import bench
    monkeypatch.setitem(sys.modules, "pynvml", None)        # import fails
    monkeypatch.setenv("PATH", "/nonexistent")
    d = bench.ClockSampler(0).stop()
    assert d["sm_mhz"] is None and d["samples"] == 0 and d["reasons"]
