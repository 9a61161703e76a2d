"""CPU: the parts of bench.py that need no GPU -- the wall-clock planner and the --impl reference arm (the oracle timed as
the CPU baseline; its JSON line must name the size that was really timed)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plan_steps_keeps_three_warmups_when_they_fit_and_one_timed_step_always():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.plan_steps(5, 20, 30.0) == (5, 20)          # everything fits
    assert bench.plan_steps(5, 20, 6.2) == (3, 4)             # N = 50000 on one GPU inside the default budget
    assert bench.plan_steps(5, 20, 3.1) == (3, 1)
    assert bench.plan_steps(5, 20, 0.4) == (1, 1)             # a solve so long that nothing else fits: still one timed step
    assert bench.plan_steps(1, 1, 10.0) == (1, 1)             # never more than requested
    w, k = bench.plan_steps(3, 100, 50.0)
    assert w == 3 and 1 <= k <= 48


def test_workload_tags_name_the_baseline_configs():
    sys.path.insert(0, ROOT)
    import bench
    assert "configs[3]" in bench.workload_name("eigen_s", 50000)
    assert "configs[4]" in bench.workload_name("eigen_sx", 100000)
    assert "configs[3]" not in bench.workload_name("eigen_s", 8000).split("(")[1].split(")")[0].replace("same family as BASELINE configs[3]", "")


def test_reference_arm_line_is_honest_about_its_sample():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "2",
                        "--warmup", "1", "--cpu-n", "800", "--ref-budget-s", "20"], capture_output=True, text=True, timeout=300,
                       cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["unit"] == "TFLOP/s" and line["higher_is_better"] is True
    assert line["config"]["n"] == 800 and line["config"]["extrapolated_to"] == 50000     # the timed size, not the headline size
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert 1 <= line["steps"] <= 2 and line["requested_steps"] == 2
    assert line["value"] > 0


def test_reference_arm_other_ranks_do_no_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=60, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
