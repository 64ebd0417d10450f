"""bench.py contract checks that need no GPU: the reference arm prints exactly ONE JSON line on stdout (library chatter goes
to stderr), honours --steps / --warmup, never maps libsmslu.so, and uses all host cores even under torchrun's OMP_NUM_THREADS=1."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def _run(extra_env=None, args=()):
    env = dict(os.environ)
    env.update(extra_env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "lap3d_96",
                        "--ref-size", "16", "--steps", "2", "--warmup", "1", *args], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return r


def test_reference_arm_prints_one_json_line():
    r = _run()
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["steps"] == 2 and d["warmup"] == 1
    assert d["metric"] == "lu_refactorize_plus_solve_per_sec" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "3D 7-point Laplacian 96^3" in d["config"]["workload"]


def test_reference_arm_non_zero_ranks_print_nothing_and_rank0_keeps_its_cores():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.stdout.strip() == ""
    r0 = _run({"RANK": "0", "WORLD_SIZE": "2", "LOCAL_RANK": "0", "OMP_NUM_THREADS": "1"})
    d = json.loads(r0.stdout.strip().splitlines()[-1])
    assert d["cpu_baseline"]["cores"] == (os.cpu_count() if os.cpu_count() <= 64 else d["cpu_baseline"]["cores"])


def test_reference_arm_does_not_load_the_product_library():
    code = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--config', 'lap3d_96', '--ref-size', '12', '--steps', '1', '--warmup', '1'];"
            "runpy.run_path(%r, run_name='__main__');"
            "maps = open('/proc/self/maps').read(); sys.stderr.write('LOADED' if 'libsmslu.so' in maps else 'CLEAN')") % os.path.join(ROOT, "bench.py")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stderr.rstrip().endswith("CLEAN"), r.stderr[-500:]
