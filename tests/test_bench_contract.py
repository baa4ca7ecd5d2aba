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


def test_summarize_regimes_puts_the_fractions_under_roofline():
    """The compact per-kernel table lives under ``roofline`` (the key the driver keeps), built from extra.*."""
    b = _bench_module()
    line = {"roofline": {"kernel": "tail_bwd_kernel"}, "extra": {
        "bandwidth_regime": {"G1_gather_rows": {"frac": 0.94, "achieved": 6166.0},
                             "G5_catalog_topk_1M": {"tcgen05_f16_m384": {"ms": 0.3, "frac_of_tensor_peak": 0.2, "frac_of_hbm_peak": 0.1}},
                             "G5_catalog_topk_16M": {"ms": 16.3, "tflops": 1027.0, "frac_of_tensor_peak": 0.61}},
        "train_16M": {"rowsparse": {"ms_per_step": 0.4}, "dense": {"unavailable": "x"}},
        "rtm_configs2": {"pvc": {"ms_per_step": 9.0, "meanpool_kernel": {"frac": 0.8, "achieved": 5000.0}}}}}
    b.summarize_regimes(line)
    tab = line["roofline"]["regimes"]
    assert tab["G1_gather_rows"]["frac_of_hbm_peak"] == 0.94 and tab["G5_16M_m4096"]["frac_of_tensor_peak"] == 0.61
    assert tab["train_16M_rowsparse"]["ms_per_step"] == 0.4 and "train_16M_dense" not in tab
    assert tab["rtm_pvc"]["meanpool_kernel_requested_over_hbm_peak_L2_resident"] == 0.8
    b.summarize_regimes({"roofline": None, "extra": None})          # nothing to do, nothing raised
