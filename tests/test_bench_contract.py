"""CPU: the parts of bench.py that run without a GPU -- the reference arm's JSON line (driver contract) and the
clock sampler's parsing of nvidia-smi rows."""
import datetime
import importlib.util
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench_module():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_reference_arm_line():
    env = dict(os.environ, OMP_NUM_THREADS="4")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "3"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "tem_train_samples_per_s" and line["unit"] == "samples/s"
    assert line["higher_is_better"] is True and line["steps"] == 1 and line["warmup"] == 3 and line["value"] > 0
    assert line["config"]["batch_per_gpu"] == 384 and "workload" in line["config"]
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_are_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "3"], capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_clock_sampler_summary_uses_the_timed_window():
    b = _bench_module()
    c = b.ClockSampler(None).start()            # no poller on this rank: parsing only
    now = datetime.datetime.now()
    c.t0, c.t1 = now, now + datetime.timedelta(milliseconds=20)

    def row(dt_ms, sm, cap="Not Active", thermal="Not Active"):
        ts = (now + datetime.timedelta(milliseconds=dt_ms)).strftime("%Y/%m/%d %H:%M:%S.%f")[:-3]
        return [ts, str(sm), "1965", "Not Active", thermal, "Not Active", cap]
    c.rows = [row(-3000, 345, thermal="Active"), row(-40, 1965), row(10, 1950, cap="Active"), row(60, 1965),
              row(5000, 210), ["garbage"]]
    s = c.summary()
    assert s["samples"] == 3 and s["sm_mhz"] == 1965.0 and s["sm_max_mhz"] == 1965.0
    assert s["reasons"] == ["sw_power_cap"]      # the idle-time thermal flag 3 s earlier is outside the window
    c.rows = []
    assert c.summary()["reasons"] == ["unavailable"]
    c.stop()


def test_guarded_section_prints_the_line_and_exits_zero_when_it_hangs():
    """bench.guarded wraps the 16M-row sections that run after the headline numbers are final: a section that never
    returns must still end in ONE printed line and exit status 0; results and exceptions pass through otherwise."""
    b = _bench_module()
    assert b.guarded(lambda: 7, 30, lambda: None) == 7
    try:
        b.guarded(lambda: 1 // 0, 30, lambda: None)
        raise AssertionError("the section's exception must reach the caller")
    except ZeroDivisionError:
        pass
    prog = ("import importlib.util, json, time\n"
            "spec = importlib.util.spec_from_file_location('bench_mod', %r)\n"
            "b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)\n"
            "line = {'value': 1.0, 'extra': None}\n"
            "def give_up():\n"
            "    line['extra'] = {'unavailable': 'watchdog'}\n"
            "    print(json.dumps(line))\n"
            "def hang():\n"
            "    while True:\n"
            "        time.sleep(0.05)\n"
            "b.guarded(hang, 1, give_up)\n"
            "print('not reached')\n") % os.path.join(ROOT, "bench.py")
    out = subprocess.run([sys.executable, "-c", prog], capture_output=True, text=True, timeout=120, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = out.stdout.strip().splitlines()
    assert len(lines) == 1 and json.loads(lines[0]) == {"value": 1.0, "extra": {"unavailable": "watchdog"}}


def test_cpu_reference_others_shape_of_the_entry():
    """cpu_baseline.others (PV step, catalog ranking literal / GEMM + topk, gather / scatter-add on the
    host cores), at 1/50 of the bench's table sizes: every entry carries value / unit / sample and a positive number."""
    b = _bench_module()
    out = b.cpu_reference_others(scale=50)
    assert out["cores"] >= 1
    for key, unit in (("pv_train_step", "reviews/s"), ("catalog_rank_1M_literal", "queries/s"),
                      ("catalog_rank_1M_gemm_topk", "queries/s"), ("gather_rows", "GB/s"), ("scatter_add_rows", "GB/s")):
        assert out[key]["unit"] == unit and out[key]["value"] > 0 and out[key]["sample"], key


def test_variant_side_measurement_parses_the_check_scripts_lines(monkeypatch):
    """extra.G5_catalog_topk_1M_variant3: the bench runs profiles/check_tc16_v2.py in a subprocess and folds its JSON
    lines (per-run lines with a digest, per-(M, variant) comparison lines) into one labelled entry."""
    b = _bench_module()
    fake = "\n".join([
        json.dumps({"n_items": 1000000, "m": 384, "epi": "1", "max_mt": 4, "ms": 0.34, "tflops": 290.0, "sha": "aa"}),
        json.dumps({"n_items": 1000000, "m": 384, "epi": "3", "max_mt": 4, "ms": 0.30, "tflops": 329.0, "sha": "aa"}),
        json.dumps({"n_items": 1000000, "m": 384, "variant": "3/4", "identical": True, "v1_ms": 0.34, "ms": 0.30, "speedup": 1.133}),
        "VERDICT: every variant returns v1's lists bit for bit"])

    class R(object):
        returncode, stdout, stderr = 0, fake, ""
    monkeypatch.setattr(b.subprocess, "run", lambda *a_, **k_: R())
    out = b.variant3_side_measurement({"bf16": 1645.0})
    assert out["exit"] == 0 and out["errors"] == []
    (run,) = out["runs"]
    assert run["identical_lists"] is True and run["m"] == 384 and run["variant_tflops"] == 329.0
    assert abs(run["variant_queries_per_s"] - 384 / 0.30e-3) < 1e-6 and abs(run["variant_frac_of_tensor_peak"] - 0.2) < 1e-9
    monkeypatch.setattr(b.subprocess, "run", lambda *a_, **k_: (_ for _ in ()).throw(OSError("no such file")))
    assert "unavailable" in b.variant3_side_measurement({"bf16": 1645.0})
