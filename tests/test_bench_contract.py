"""CPU test of the bench.py contract for the reference arm (`--impl reference` needs no GPU): one JSON
line with the keys the driver reads, the oracle port timed on the host cores and, where oracle/_ref
exists, the translated reference's serial code next to it."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "gx3",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["dtype"] == "f64"
    assert d["unit"] == "grid-cell-subcycles/s" and d["value"] > 1e5
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"]
    assert "subcycles" in cb["sample"]
    assert "workload" in d["config"] and "model" not in d["config"]
    from oracle import oracle as O
    if O.ref_available() and os.path.exists(os.path.join(O.REF_DIR, "libevp_ref_cice4_fast.so")):
        # the line's value is the reference's own code; the port and the serial build are timed next to it
        assert cb["kind"] == "reference"
        assert cb["port"]["kind"] == "port" and cb["port"]["value"] > 1e5
        rs = cb["reference_serial"]
        assert rs["kind"] == "reference" and rs["cores"] == 1 and rs["value"] > 1e5
    else:
        assert cb["kind"] == "port"
    # a step is a full ndte loop of the SAME workload, dt included, as the b200 arm prints it
    sys.path.insert(0, ROOT)
    import bench
    case = bench.build_case("gx3", False)
    assert d["config"]["workload"] == bench.workload_string(case, 120, False)
    assert "dt=3600s" in d["config"]["workload"] and "120 subcycles" in cb["sample"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--workload", "gx3", "--steps", "1", "--warmup", "0"], capture_output=True, text=True,
                       timeout=120, env=env, cwd=ROOT)
    assert r.returncode == 0 and not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
