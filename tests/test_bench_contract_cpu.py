"""bench.py's reference arm (`--impl reference`: the UNMODIFIED reference from oracle/_ref when oracle/fetch_ref.py has
populated it, else the numpy oracle port -- the one place outside tests/ and smoke() that may execute oracle/) runs
without a GPU: check the JSON line the driver parses, that its `config` is the one the GPU arm prints, and that under a
multi-rank launch only rank 0 works and prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra):
    env = dict(os.environ, **env_extra)
    env.pop("CUDA_VISIBLE_DEVICES", None)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus",
                          env_extra.get("WORLD_SIZE", "1"), "--steps", "1", "--warmup", "1", "--batch", "4096"],
                         capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    return [l for l in out.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_prints_the_contract_line():
    lines = _run({})
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "LID-VAE train samples/s" and d["unit"] == "samples/s"
    assert d["higher_is_better"] is True and d["steps"] == 1 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["config"]["workload"].startswith("configs[1]")
    sys.path.insert(0, ROOT)
    import bench
    from oracle import fetch_ref
    assert d["config"] == bench.make_config(4096, 1)          # same config object as the GPU arm's line
    assert d["config"]["encoder"] == "included" and "Adam" in d["config"]["optimizer"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == ("reference" if fetch_ref.available() else "port")
    assert cb["cores"] >= 1 and cb["value"] == d["value"] and "4096" in cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    assert _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}) == []
