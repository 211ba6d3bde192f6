"""CPU guards for the two driver entry points that only ever run on the GPU box: bench.py and
__graft_entry__.py must not reference names that are never defined (an edit once dropped a helper
that only the GPU arm calls), and the clock sampler must parse nvidia-smi's output."""
import ast
import builtins
import os
import stat
import sys
import time

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _undefined_names(path):
    tree = ast.parse(open(path).read())
    defined = set(dir(builtins)) | {"__file__", "__name__"}
    for node in ast.walk(tree):
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)):
            defined.add(node.name)
        elif isinstance(node, ast.Name) and isinstance(node.ctx, (ast.Store, ast.Del)):
            defined.add(node.id)
        elif isinstance(node, ast.arg):
            defined.add(node.arg)
        elif isinstance(node, (ast.Import, ast.ImportFrom)):
            for a in node.names:
                defined.add((a.asname or a.name).split(".")[0])
        elif isinstance(node, ast.ExceptHandler) and node.name:
            defined.add(node.name)
    used = {n.id for n in ast.walk(tree) if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load)}
    return sorted(used - defined)


@pytest.mark.parametrize("rel", ["bench.py", "__graft_entry__.py", "msmbuilder_b200/_kernels.py",
                                 "msmbuilder_b200/parallel.py", "msmbuilder_b200/_device.py",
                                 "msmbuilder_b200/cluster/kcenters.py", "tools/check_parallel.py"])
def test_no_undefined_names(rel):
    assert _undefined_names(os.path.join(ROOT, rel)) == []


def test_clock_sampler_windows(tmp_path, monkeypatch):
    fake = tmp_path / "nvidia-smi"
    fake.write_text('''#!/usr/bin/env python3
import time, datetime
time.sleep(0.25)
while True:
    t = datetime.datetime.now().strftime("%Y/%m/%d %H:%M:%S.%f")[:-3]
    print("%s, 1550, 1965, 812.3, 0x0000000000000004, Not Active, Not Active, Not Active, Active" % t, flush=True)
    time.sleep(0.05)
''')
    fake.chmod(fake.stat().st_mode | stat.S_IEXEC)
    monkeypatch.setenv("PATH", str(tmp_path) + os.pathsep + os.environ["PATH"])
    sys.path.insert(0, ROOT)
    import bench
    s = bench.ClockSampler(0)
    s.start()
    time.sleep(0.6)
    s.mark_begin()
    time.sleep(0.3)
    s.mark_end()
    out = s.stop()
    assert out["window"] == "timed region" and 3 <= out["samples"] <= 8
    assert out["sm_mhz"] == 1550.0 and out["sm_max_mhz"] == 1965.0 and out["reasons"] == ["sw_power_cap"]
    s = bench.ClockSampler(0)       # a timed region shorter than the sampling period falls back to the warm-up
    s.start()
    time.sleep(0.6)
    s.mark_begin()
    s.mark_end()
    out = s.stop()
    assert out["samples"] >= 1 and out["window"].startswith("warm-up")
    assert bench.ClockSampler(0).stop()["samples"] == 0      # never started: empty, no exception


def test_reference_arm_of_the_assign_workload_runs_on_cpu():
    # bench.py --impl reference --workload assign: the reference's assign.hpp (oracle/_ref, or the port) on a
    # bounded sample; one JSON line with the keys the driver reads
    import json
    import subprocess
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload",
                          "assign", "--steps", "1", "--warmup", "0", "--assign-frames", "20000"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["metric"] == "frames/sec assign_nearest"
    assert line["value"] > 0 and line["unit"] == "frames/s" and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] in ("reference", "port")
    assert line["e2e"]["h2d_bytes_per_step"] == 0
